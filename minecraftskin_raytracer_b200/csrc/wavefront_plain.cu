// The kernels of wavefront.cu for scenes without posed boxes (namespace mcskin::plain): see dev_types.cuh.
#define MCSKIN_POSED 0
#include "wavefront.cu"
