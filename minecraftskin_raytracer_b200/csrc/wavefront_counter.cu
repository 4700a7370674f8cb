// The kernels of wavefront.cu with counter-based random streams (McConfig::rng_mode 1; namespace mcskin::counter):
// see dev_types.cuh and dev_mt19937.cuh.
#define MCSKIN_COUNTER_RNG 1
#include "wavefront.cu"
