// dev_stage.cuh — staging the scene blob into shared memory, once per CTA, with one
// bulk asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier;
// SASS: UBLKCP + SYNCS).  The blob is a few KB (12 boxes = 2.1 KB), so a single
// elected thread issues it and every thread waits on the barrier's phase 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_intersect.cuh"

namespace mcskin {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// dst: 16-byte aligned shared memory, bytes: multiple of 16.  All threads of the CTA call.
__device__ __forceinline__ void stage_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t barAddr = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(barAddr)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(barAddr)
            : "memory");
    }
}

// Scene view over a blob that already sits at `blob` (shared or global memory).
__device__ __forceinline__ SceneView scene_view(const unsigned char* blob, const float4* texels, const DevFrame& fr) {
    const int nBoxes = fr.n_boxes;
    const SceneBlobLayout lay(nBoxes);
    SceneView sc;
    sc.lo = reinterpret_cast<const float4*>(blob + lay.loOffset());
    sc.hi = reinterpret_cast<const float4*>(blob + lay.hiOffset());
    sc.rect = fr.box_rects_valid ? reinterpret_cast<const int4*>(blob + lay.rectOffset()) : nullptr;
    sc.boxes = reinterpret_cast<const DevBox*>(blob + lay.boxOffset());
    sc.texels = texels;
    sc.n_boxes = nBoxes;
    sc.posed_mask = fr.posed_mask;
    sc.usable_mask = fr.usable_mask;
    sc.opaque_mask = fr.opaque_mask;
    sc.rotated_mask = fr.rotated_mask;
    sc.opaque_posed_mask = fr.opaque_posed_mask;
    sc.root_mask = fr.root_mask;
    return sc;
}

}  // namespace mcskin
