// wavefront.cu — see wavefront.cuh for the design.  Every arithmetic step is the same
// device code the megakernel uses (dev_intersect.cuh / dev_shade.cuh); only the order in
// which threads meet the work differs, so results are bit-identical between the two.
#include "wavefront.cuh"

#include <algorithm>

#include "dev_sample.cuh"

namespace mcskin {

namespace {

constexpr int kWfThreads = 256;
// Minimum resident blocks per SM asked of the compiler (i.e. register caps), from a sweep on B200:
// 4 blocks (64 registers) for the primary-hit and the queued shade kernels (+1.3 % frame rate),
// compiler's choice for the shadow kernel (59 registers; capping it at 48 was slower) and for
// the in-thread tail form of the shade kernel (128).
#ifndef MCSKIN_WF_HIT0_MIN_BLOCKS
#define MCSKIN_WF_HIT0_MIN_BLOCKS 4
#endif
#ifndef MCSKIN_WF_SHADE_MIN_BLOCKS
#define MCSKIN_WF_SHADE_MIN_BLOCKS 4
#endif
#define WF_HIT0_BOUNDS __launch_bounds__(kWfThreads, MCSKIN_WF_HIT0_MIN_BLOCKS)
#ifdef MCSKIN_WF_SHADOW_MIN_BLOCKS
#define WF_SHADOW_BOUNDS __launch_bounds__(kWfThreads, MCSKIN_WF_SHADOW_MIN_BLOCKS)
#else
#define WF_SHADOW_BOUNDS __launch_bounds__(kWfThreads)
#endif
#define WF_SHADE_BOUNDS __launch_bounds__(kWfThreads, QUEUED ? MCSKIN_WF_SHADE_MIN_BLOCKS : 2)

__device__ __forceinline__ unsigned int pack_hit(const Hit& h) {
    return static_cast<unsigned int>(h.box & 0xffff) | (static_cast<unsigned int>(h.face) << 16) |
           (h.flip ? (1u << 24) : 0u);
}
__device__ __forceinline__ Hit unpack_hit(float4 geo, float4 org) {
    Hit h;
    const unsigned int k = __float_as_uint(geo.w);
    h.p = mk3(geo.x, geo.y, geo.z);
    h.box = static_cast<int>(k & 0xffffu);
    h.face = static_cast<int>((k >> 16) & 0xffu);
    h.flip = (k >> 24) & 1u;
    h.texel = __float_as_int(org.w);
    h.t = 0.0f;
    return h;
}

// Appends the hits of a warp to a queue with one atomic; all 32 lanes must call.
__device__ __forceinline__ void enqueue_hit(const HitQueueView& q, unsigned int* counter, unsigned int capacity,
                                            bool isHit, const Hit& h, const Ray& ray, unsigned int path) {
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int m = __ballot_sync(0xffffffffu, isHit);
    if (m == 0u) return;
    unsigned int base = 0u;
    if (lane == 0u) base = atomicAdd(counter, static_cast<unsigned int>(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (isHit) {
        const unsigned int i = base + __popc(m & ((1u << lane) - 1u));
        if (i < capacity) {
            q.geo[i] = make_float4(h.p.x, h.p.y, h.p.z, __uint_as_float(pack_hit(h)));
            q.org[i] = make_float4(ray.o.x, ray.o.y, ray.o.z, __int_as_float(h.texel));
            q.dir[i] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(path));
        }
    }
}


// ---------------------------------------------------------------- primary hits
template <bool BATCH>
__global__ void WF_HIT0_BOUNDS
k_wf_hit0(const DevFrame fr, const FramePointers fp_, const ActiveList list_, const WaveView wv_,
          const BatchSlice* __restrict__ batch) {
    __shared__ __align__(8) uint64_t stageBar;
    const FramePointers& fp = BATCH ? batch[blockIdx.y].fp : fp_;
    const ActiveList& list = BATCH ? batch[blockIdx.y].list : list_;
    const WaveView& wv = BATCH ? batch[blockIdx.y].wave : wv_;
    unsigned int count = *list.count;
    if (count > list.capacity) count = list.capacity;
    if (count > wv.slotCapacity) count = wv.slotCapacity;
    const unsigned long long nPaths = static_cast<unsigned long long>(count) * fr.spp;
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kWfThreads;
    if (static_cast<unsigned long long>(blockIdx.x) * kWfThreads >= nPaths) return;
    stage_bulk(g_sceneSmem, fp.blob, fp.blob_bytes, &stageBar);
    const SceneView sc = scene_view(g_sceneSmem, fp.texels, fr);
    const int spp = fr.spp, dps = fr.draws_per_sample;

    for (unsigned long long p0 = static_cast<unsigned long long>(blockIdx.x) * kWfThreads + (threadIdx.x & ~31u);
         p0 < nPaths; p0 += stride) {
        const unsigned long long p = p0 + (threadIdx.x & 31u);
        bool isHit = false;
        Hit hit;
        hit.box = -1; hit.face = 0; hit.texel = 0; hit.flip = false; hit.t = 0.0f; hit.p = mk3(0.f, 0.f, 0.f);
        Ray ray;
        ray.o = mk3(0.f, 0.f, 0.f);
        ray.d = mk3(0.f, 0.f, 0.f);
        if (p < nPaths) {
            const unsigned int slot = static_cast<unsigned int>(p / spp);
            const int s = static_cast<int>(p - static_cast<unsigned long long>(slot) * spp);
            const uint2 sp = list.slot_pixel[slot];
            if (sp.x != kUnusedSlot) {
                const int px = static_cast<int>(sp.y & 0xffffu), py = static_cast<int>(sp.y >> 16);
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
                const float* rec = list.records + (static_cast<size_t>(slot) * spp + s) * dps;
                if (dps == 2) {
                    const float2 r = *reinterpret_cast<const float2*>(rec);
                    d0 = r.x; d1 = r.y;
                } else if (dps == 4) {
                    const float4 r = *reinterpret_cast<const float4*>(rec);
                    d0 = r.x; d1 = r.y; d2 = r.z; d3 = r.w;
                }
                const SampleDraws sd = assign_draws(fr, d0, d1, d2, d3);
                float u, v;
                sample_uv(fr, px, py, sd, &u, &v);
                ray = primary_ray(fr, u, v, sd);
                // (restricting this query to the boxes whose screen rectangle holds the pixel, as the primary
                // pass does, measured slower here: the branch-free reject pass over all boxes is cheaper than
                // the mask plus a bit loop when nearly every ray hits)
                hit = closest_hit(sc, ray);
                if (fr.max_bounces < 0) {
                    // traceRay returns at once (depth 0 > maxBounces, raytracer.cpp:86-90); the tile
                    // renderer's re-test replaces misses by the gradient (tile_renderer.cpp:111-114)
                    wv.tail[p] = hit.box < 0 ? config_background(fr, u, v) : config_background(fr, 0.5f, 0.5f);
                    wv.top[p] = 0;
                } else if (hit.box < 0) {
                    wv.tail[p] = config_background(fr, u, v);
                    wv.top[p] = 0;
                } else {
                    isHit = true;
                }
            } else {
                wv.top[p] = -1;  // unused slot
            }
        }
        enqueue_hit(wv.q[0], &wv.qCount[0], wv.pathCapacity, isHit, hit, ray, static_cast<unsigned int>(p));
    }
}

// ---------------------------------------------------------------- soft shadows of one queue level
// computeSoftShadow (shading.cpp:28-60) for every hit of a queue, in ONE kernel.  A block works through
// chunks of 256 hits, claimed from a device counter, in three block-synchronous phases:
//   A  thread = hit: the common origin of the hit's shadow rays and the boxes their bundle can reach.
//      Nothing in reach -> all N rays are lit by construction, the hit is done.  The others are
//      COMPACTED into a ring in shared memory, so that phase B runs with full warps however the
//      "nothing in reach" hits are spread over the queue (they were a third of the lanes of the
//      seeding loop when it ran per queue entry).
//   B  thread = pending hit, once 256 of them are waiting (or the queue is exhausted): the fresh
//      std::mt19937 of the hit (raytracer.cpp:110-113) — 396 dependent LCG steps — and its next
//      8 points on the light's disk, into shared memory.
//   C  thread = (pending hit, sample): one shadow ray; unoccluded rays are counted per hit.
// The light points never leave the SM (they were 96 B per hit written to and read back from HBM), and
// the engine runs only for hits that cast rays.  N > 8 samples: phases B and C repeat in rounds of 8,
// the engine state waiting in shared memory in between.
constexpr int kPendCap = 2 * kWfThreads;    // ring of hits waiting for their engine (power of two)
constexpr int kShadowRound = 8;             // light samples per hit and round
constexpr int kPtStride = kShadowRound * 3 + 1;  // floats per hit in the point table (odd: no bank conflicts)
struct ShadowSmem {
    float4 pendP[kPendCap];                 // hit point xyz, w = index in the queue
    float4 pendO[kPendCap];                 // shadow-ray origin xyz, w = box mask of the bundle
    float pts[kWfThreads * kPtStride];      // this round's sample points
    uint4 engine[kWfThreads];               // FreshStream between rounds (N > kShadowRound only)
    unsigned int lit[kWfThreads];
    unsigned int head, tail;                // monotonic: ring entries [head, tail) are pending
    unsigned int chunk;                     // the chunk claimed for this trip
};
static_assert((kPendCap & (kPendCap - 1)) == 0, "ring size must be a power of two");

#ifndef MCSKIN_WF_SOFT_MIN_BLOCKS
#define MCSKIN_WF_SOFT_MIN_BLOCKS 4
#endif
template <bool BATCH>
__global__ void __launch_bounds__(kWfThreads, MCSKIN_WF_SOFT_MIN_BLOCKS)
k_wf_softshadow(const DevFrame fr, const FramePointers fp_, const WaveView wv_, const int which, const int depth,
                const BatchSlice* __restrict__ batch) {
    __shared__ __align__(8) uint64_t stageBar;
    const FramePointers& fp = BATCH ? batch[blockIdx.y].fp : fp_;
    const WaveView& wv = BATCH ? batch[blockIdx.y].wave : wv_;
    unsigned int n = wv.qCount[depth];
    if (n > wv.pathCapacity) n = wv.pathCapacity;
    const unsigned int nChunks = (n + kWfThreads - 1) / kWfThreads;
    if (blockIdx.x >= nChunks) return;
    stage_bulk(g_sceneSmem, fp.blob, fp.blob_bytes, &stageBar);
    const SceneView sc = scene_view(g_sceneSmem, fp.texels, fr);
    ShadowSmem* sm = reinterpret_cast<ShadowSmem*>(g_sceneSmem + ((fp.blob_bytes + 15u) & ~15u));
    const int tid = threadIdx.x, lane = tid & 31;
    const HitQueueView q = wv.q[which];
    const int N = fr.shadow_samples;
    const int rounds = (N + kShadowRound - 1) / kShadowRound;
    const V3 lightCentre = ld3(fr.light_pos);
    const uint32_t one = fr.spp > 0 ? 1u : 0u;  // a 1 the compiler cannot see through
    unsigned int* chunkCounter = wv.qCount + (wv.levels + 3) + depth;
    if (tid == 0) {
        sm->head = 0u;
        sm->tail = 0u;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) sm->chunk = atomicAdd(chunkCounter, 1u);
        __syncthreads();
        const unsigned int c = sm->chunk;
        const bool last = c >= nChunks;  // the queue is exhausted: flush what is pending and leave
        if (!last) {  // ---- phase A
            const unsigned int i = c * kWfThreads + tid;
            bool pend = false;
            float4 eP = make_float4(0.f, 0.f, 0.f, 0.f), eO = eP;
            if (i < n) {
                const Hit h = unpack_hit(q.geo[i], make_float4(0.f, 0.f, 0.f, 0.f));
                // P + n*eps (isInShadow, shading.cpp:17); computeSoftShadow hands it the raw hit normal
                // (shading.cpp:54, raytracer.cpp:113)
                const V3 origin = h.p + hit_normal(sc, h) * kShadowEpsilon;
                const uint32_t allow = bundle_box_mask(sc, origin, lightCentre, fr.light_radius);
                if (allow == 0u && sc.n_boxes <= 32) {
                    wv.lit[i] = static_cast<unsigned int>(N);
                } else {
                    pend = true;
                    eP = make_float4(h.p.x, h.p.y, h.p.z, __uint_as_float(i));
                    eO = make_float4(origin.x, origin.y, origin.z, __uint_as_float(allow));
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, pend);
            if (m) {
                unsigned int base = 0u;
                if (lane == 0) base = atomicAdd(&sm->tail, static_cast<unsigned int>(__popc(m)));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pend) {
                    const unsigned int pos = (base + __popc(m & ((1u << lane) - 1u))) & (kPendCap - 1);
                    sm->pendP[pos] = eP;
                    sm->pendO[pos] = eO;
                }
            }
            __syncthreads();
        }
        for (;;) {  // ---- drain the ring, 256 hits at a time
            const unsigned int head = sm->head;
            const unsigned int avail = sm->tail - head;
            if (avail == 0u || (!last && avail < static_cast<unsigned int>(kWfThreads))) break;
            const unsigned int cnt = avail < static_cast<unsigned int>(kWfThreads) ? avail : static_cast<unsigned int>(kWfThreads);
            const bool mine = static_cast<unsigned int>(tid) < cnt;
            if (mine) sm->lit[tid] = 0u;
            for (int round = 0; round < rounds; ++round) {
                const int ns = min(kShadowRound, N - round * kShadowRound);
                if (mine) {  // ---- phase B
                    const float4 eP = sm->pendP[(head + tid) & (kPendCap - 1)];
                    const V3 P = mk3(eP.x, eP.y, eP.z);
                    FreshStream rng;
                    if (round == 0) {
                        rng.seed_balanced(shadow_seed(P, depth), one);
                    } else {
                        const uint4 e = sm->engine[tid];
                        rng.cur = e.x; rng.nxt = e.y; rng.far = e.z; rng.j = e.w;
                    }
                    soft_shadow_positions(fr, P, ns, rng, sm->pts + tid * kPtStride);
                    if (round + 1 < rounds) sm->engine[tid] = make_uint4(rng.cur, rng.nxt, rng.far, rng.j);
                }
                __syncthreads();
                const unsigned int nRays = cnt * static_cast<unsigned int>(ns);  // ---- phase C
                for (unsigned int r = tid; r < nRays; r += kWfThreads) {
                    const unsigned int e = ns == kShadowRound ? r >> 3 : r / static_cast<unsigned int>(ns);
                    const unsigned int k = r - e * static_cast<unsigned int>(ns);
                    const float4 eO = sm->pendO[(head + e) & (kPendCap - 1)];
                    const float* pt = sm->pts + e * kPtStride + k * 3u;
                    if (!in_shadow_from(sc, mk3(eO.x, eO.y, eO.z), mk3(pt[0], pt[1], pt[2]), __float_as_uint(eO.w)))
                        atomicAdd(&sm->lit[e], 1u);
                }
                __syncthreads();
            }
            if (mine) wv.lit[__float_as_uint(sm->pendP[(head + tid) & (kPendCap - 1)].w)] = sm->lit[tid];
            __syncthreads();
            if (tid == 0) sm->head = head + cnt;
            __syncthreads();
        }
        if (last) break;
    }
}

// hard shadows: one ray from every hit to the light's centre (isInShadow as shade() calls it, with the
// normalised normal: shading.cpp:69,78)
template <bool BATCH>
__global__ void __launch_bounds__(kWfThreads)
k_wf_hardshadow(const DevFrame fr, const FramePointers fp_, const WaveView wv_, const int which, const int depth,
                const BatchSlice* __restrict__ batch) {
    __shared__ __align__(8) uint64_t stageBar;
    const FramePointers& fp = BATCH ? batch[blockIdx.y].fp : fp_;
    const WaveView& wv = BATCH ? batch[blockIdx.y].wave : wv_;
    unsigned int n = wv.qCount[depth];
    if (n > wv.pathCapacity) n = wv.pathCapacity;
    if (blockIdx.x * kWfThreads >= n) return;
    stage_bulk(g_sceneSmem, fp.blob, fp.blob_bytes, &stageBar);
    const SceneView sc = scene_view(g_sceneSmem, fp.texels, fr);
    const HitQueueView q = wv.q[which];
    const V3 lightCentre = ld3(fr.light_pos);
    for (unsigned int i = blockIdx.x * kWfThreads + threadIdx.x; i < n; i += gridDim.x * kWfThreads) {
        const Hit h = unpack_hit(q.geo[i], make_float4(0.f, 0.f, 0.f, 0.f));
        // (a box mask for the single ray, as for soft-shadow bundles, costs more than the one ray it can save)
        wv.lit[i] = in_shadow(sc, h.p, normalize3(hit_normal(sc, h)), lightCentre) ? 0u : 1u;
    }
}

// ---------------------------------------------------------------- shade + bounce
// tail == 0: a queue level whose shadow rays ran in k_wf_softshadow / k_wf_hardshadow; bounce hits go to the next queue.
// tail == 1: the last queue (depth == wv.queueLevels).  By then only ~2 % of the paths are left,
//            too few to be worth three launches per level: each thread takes one queued hit and
//            follows its chain to the end, evaluating shadow rays in place.  The queue is compact,
//            so warps start full and only thin out at the deepest, rarest levels.
// QUEUED: the visibility comes from the shadow kernel's counters and bounce hits go to the next queue
// (the lean form: no shadow code at all); otherwise shadows are evaluated in place, and with
// tail != 0 the whole remaining chain is.
template <bool QUEUED, bool BATCH>
__global__ void WF_SHADE_BOUNDS
k_wf_shade(const DevFrame fr, const FramePointers fp_, const WaveView wv_, const int which, const int depth,
           const int tailArg, const BatchSlice* __restrict__ batch) {
    const int tail = QUEUED ? 0 : tailArg;
    const FramePointers& fp = BATCH ? batch[blockIdx.y].fp : fp_;
    const WaveView& wv = BATCH ? batch[blockIdx.y].wave : wv_;
    __shared__ __align__(8) uint64_t stageBar;
    unsigned int n = wv.qCount[depth];
    if (n > wv.pathCapacity) n = wv.pathCapacity;
    if (blockIdx.x * kWfThreads >= n) return;
    stage_bulk(g_sceneSmem, fp.blob, fp.blob_bytes, &stageBar);
    const SceneView sc = scene_view(g_sceneSmem, fp.texels, fr);
    const HitQueueView q = wv.q[which];
    const HitQueueView qNext = wv.q[which ^ 1];
    const bool cfg = fr.use_config != 0;
    const size_t cap = wv.pathCapacity;

    for (unsigned int i0 = blockIdx.x * kWfThreads + (threadIdx.x & ~31u); i0 < n; i0 += gridDim.x * kWfThreads) {
        const unsigned int i = i0 + (threadIdx.x & 31u);
        bool bounceHit = false;
        Hit next;
        next.box = -1; next.face = 0; next.texel = 0; next.flip = false; next.t = 0.0f; next.p = mk3(0.f, 0.f, 0.f);
        Ray ray;
        ray.o = mk3(0.f, 0.f, 0.f);
        ray.d = mk3(0.f, 0.f, 0.f);
        unsigned int path = 0u;
        if (i < n) {
            const float4 g = q.geo[i], o = q.org[i], dd = q.dir[i];
            Hit h = unpack_hit(g, o);
            path = __float_as_uint(dd.w);
            V3 rayO = mk3(o.x, o.y, o.z), rayD = mk3(dd.x, dd.y, dd.z);
            int d = depth;
            for (;;) {
                const V3 P = h.p;
                const V3 nrm = hit_normal(sc, h);
                const float4 tex = hit_texel(sc, h);
                const V3 viewDir = normalize3(rayO - P);
                float vis;
                if (QUEUED) {
                    vis = static_cast<float>(wv.lit[i]) / static_cast<float>(wv.shadowRays);
                    if (wv.shadowMode == kShadowHard) vis = wv.lit[i] ? 1.0f : 0.0f;
                } else if (cfg && fr.soft_on) {
                    vis = soft_shadow(sc, fr, P, nrm, fr.shadow_samples, shadow_seed(P, d));
                } else {
                    vis = in_shadow(sc, P, normalize3(nrm), ld3(fr.light_pos)) ? 0.0f : 1.0f;
                }
                float4 shaded = shade_lit(fr, P, nrm, tex, viewDir, vis);
                const float alpha = shaded.w;
                if (cfg && fr.ao_on && d == 0) {
                    const float ao = ambient_occlusion(sc, P, nrm, fr.ao_samples, fr.ao_radius, ao_seed(P));
                    const float f = 1.0f - fr.ao_intensity * (1.0f - ao);
                    shaded.x *= f;
                    shaded.y *= f;
                    shaded.z *= f;
                }
                if (d < fr.max_bounces && d < wv.levels) {
                    wv.stack[static_cast<size_t>(d) * cap + path] = make_float4(shaded.x, shaded.y, shaded.z, alpha);
                    const V3 N = normalize3(nrm);
                    const V3 D = normalize3(rayD);
                    V3 Rd = D - N * (2.0f * dot3(D, N));
                    Rd = normalize3(Rd);
                    ray.o = P + N * kReflectEpsilon;
                    ray.d = Rd;
                    next = closest_hit(sc, ray);
                    if (next.box < 0) {
                        wv.tail[path] = flat_background(fr);  // bounced rays see the flat colour (raytracer.cpp:101)
                        wv.top[path] = d + 1;
                        break;
                    }
                    if (!tail) {  // hand the next level to the queues
                        bounceHit = true;
                        break;
                    }
                    h = next;
                    rayO = ray.o;
                    rayD = ray.d;
                    ++d;
                    continue;
                }
                shaded.w = alpha;
                wv.tail[path] = clamp4(shaded);
                wv.top[path] = d;
                break;
            }
        }
        enqueue_hit(qNext, &wv.qCount[depth + 1], wv.pathCapacity, bounceHit, next, ray, path);
    }
}

// ---------------------------------------------------------------- fold + ordered average
__device__ __forceinline__ float4 fold_path(const WaveView& wv, size_t path) {
    float4 c = wv.tail[path];
    const int top = wv.top[path];
    if (top > 0) {
        const size_t cap = wv.pathCapacity;
        const float alpha0 = wv.stack[path].w;
        const float keep = 1.0f - kReflectivity;
        for (int k = top - 1; k >= 0; --k) {  // raytracer.cpp:143-147, innermost level first
            const float4 s = wv.stack[static_cast<size_t>(k) * cap + path];
            float4 m;
            m.x = s.x * keep + c.x * kReflectivity;
            m.y = s.y * keep + c.y * kReflectivity;
            m.z = s.z * keep + c.z * kReflectivity;
            m.w = alpha0;
            c = clamp4(m);
        }
    }
    return c;
}

template <bool BATCH>
__global__ void __launch_bounds__(kWfThreads)
k_wf_resolve_warp(const DevFrame fr, const BandView band_, const ActiveList list_, const WaveView wv_, const int lgSpp,
                  const BatchSlice* __restrict__ batch) {
    const BandView& band = BATCH ? batch[blockIdx.y].band : band_;
    const ActiveList& list = BATCH ? batch[blockIdx.y].list : list_;
    const WaveView& wv = BATCH ? batch[blockIdx.y].wave : wv_;
    __shared__ __align__(16) float stageAll[kWfThreads / 32][kWarpStageFloats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int count = *list.count;
    if (count > list.capacity) count = list.capacity;
    if (count > wv.slotCapacity) count = wv.slotCapacity;
    const int spp = fr.spp;
    const int pixPerGroup = 32 >> lgSpp;
    const unsigned int nGroups = (count + pixPerGroup - 1) / pixPerGroup;
    const int pix = lane >> lgSpp, s = lane & (spp - 1);
    float* stageW = stageAll[warp];
    for (unsigned int g = blockIdx.x * (kWfThreads / 32) + warp; g < nGroups; g += gridDim.x * (kWfThreads / 32)) {
        const unsigned int slot = g * pixPerGroup + pix;
        uint2 sp = make_uint2(kUnusedSlot, 0u);
        if (slot < count) sp = list.slot_pixel[slot];
        const bool on = sp.x != kUnusedSlot;
        float4 colour = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) colour = fold_path(wv, static_cast<size_t>(slot) * spp + s);
        const unsigned int leaders = __ballot_sync(0xffffffffu, on && s == 0);
        unsigned int resolveMask = 0u;
        {
            unsigned int m = leaders;
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1u;
                resolveMask |= 1u << (l >> lgSpp);
            }
        }
        warp_resolve(fr, band, stageW, lane, spp, lgSpp, colour, sp.x, resolveMask);
    }
}

// any spp: one thread sums a pixel's folded samples in order
template <bool BATCH>
__global__ void __launch_bounds__(kWfThreads)
k_wf_resolve_pixel(const DevFrame fr, const BandView band_, const ActiveList list_, const WaveView wv_,
                   const BatchSlice* __restrict__ batch) {
    const BandView& band = BATCH ? batch[blockIdx.y].band : band_;
    const ActiveList& list = BATCH ? batch[blockIdx.y].list : list_;
    const WaveView& wv = BATCH ? batch[blockIdx.y].wave : wv_;
    unsigned int count = *list.count;
    if (count > list.capacity) count = list.capacity;
    if (count > wv.slotCapacity) count = wv.slotCapacity;
    const int spp = fr.spp;
    for (unsigned int slot = blockIdx.x * kWfThreads + threadIdx.x; slot < count; slot += gridDim.x * kWfThreads) {
        const uint2 sp = list.slot_pixel[slot];
        if (sp.x == kUnusedSlot) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < spp; ++s) acc = add4(acc, fold_path(wv, static_cast<size_t>(slot) * spp + s));
        store_pixel(band, sp.x, scale4(acc, fr.inv_spp));
    }
}

// zeroes the active-pixel counter and the queue counters of every scene of a batch
__global__ void k_batch_reset(const BatchSlice* __restrict__ batch, const int nScenes, const int nQueueCounters) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nScenes) return;
    *batch[s].list.count = 0u;
    for (int i = 0; i < nQueueCounters; ++i) batch[s].wave.qCount[i] = 0u;
}

int log2_pow2_le32(int v) {
    for (int lg = 0; lg <= 5; ++lg)
        if (v == (1 << lg)) return lg;
    return -1;
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

int shadow_mode_of(const DevFrame& fr) {
    if (!(fr.use_config && fr.soft_on)) return kShadowHard;
    if (fr.shadow_samples <= 1 || fr.light_radius < 1e-4f || 2 * fr.shadow_samples > kFreshStreamMaxDraws)
        return kShadowInThread;
    return kShadowSoft;
}
int stack_levels_of(const DevFrame& fr) {
    return fr.max_bounces < 0 ? 0 : (fr.max_bounces > kMaxStackDepth ? kMaxStackDepth : fr.max_bounces);
}

}  // namespace

size_t wavefront_bytes_per_path(const DevFrame& fr) {
    return 2 * 3 * sizeof(float4)            // two hit queues
           + sizeof(unsigned)                // lit counters
           + sizeof(float4) + sizeof(int)    // tail, top
           + sizeof(float4) * stack_levels_of(fr);
}

size_t wavefront_fixed_bytes(const DevFrame& fr) {
    return 256 * 16 + 2 * sizeof(unsigned int) * (stack_levels_of(fr) + 3);
}

bool wavefront_carve(const DevFrame& fr, void* base, size_t bytes, unsigned int pathCapacity, int gridBlocks,
                     WaveView* out) {
    WaveView w{};
    w.pathCapacity = pathCapacity;
    w.slotCapacity = pathCapacity / static_cast<unsigned int>(fr.spp);
    w.levels = stack_levels_of(fr);
    w.shadowMode = shadow_mode_of(fr);
    w.shadowRays = w.shadowMode == kShadowSoft ? fr.shadow_samples : (w.shadowMode == kShadowHard ? 1 : 0);
    w.gridBlocks = gridBlocks;
    w.queueLevels = 3;
    w.deepGridDiv = 1;
    w.softGrid = gridBlocks;
    unsigned char* p = static_cast<unsigned char*>(base);
    size_t off = 0;
    auto take = [&](size_t n) {
        void* r = p + off;
        off = align256(off + n);
        return r;
    };
    const size_t cap = pathCapacity;
    for (int k = 0; k < 2; ++k) {
        w.q[k].geo = static_cast<float4*>(take(cap * sizeof(float4)));
        w.q[k].org = static_cast<float4*>(take(cap * sizeof(float4)));
        w.q[k].dir = static_cast<float4*>(take(cap * sizeof(float4)));
    }
    w.lit = static_cast<unsigned int*>(take(cap * sizeof(unsigned int)));
    w.tail = static_cast<float4*>(take(cap * sizeof(float4)));
    w.top = static_cast<int*>(take(cap * sizeof(int)));
    w.stack = static_cast<float4*>(take(std::max<size_t>(16, cap * sizeof(float4) * w.levels)));
    // queue sizes per depth, then the chunk counters of the shadow kernel per depth (zeroed together)
    w.qCount = static_cast<unsigned int*>(take(2 * sizeof(unsigned int) * (w.levels + 3)));
    if (off > bytes) return false;
    *out = w;
    return true;
}

void launch_batch_reset(const BatchSlice* batch, int nScenes, int levels, cudaStream_t stream) {
    if (nScenes > 0) k_batch_reset<<<(nScenes + 127) / 128, 128, 0, stream>>>(batch, nScenes, 2 * (levels + 3));
}

void launch_wavefront(const DevFrame& fr, const FramePointers& fp, const BandView& band, const ActiveList& list,
                      const WaveView& wv, unsigned int* groupCounter, cudaStream_t stream, int* launches,
                      const BatchSlice* batch, int nScenes) {
    int n = 0;
    const int grid = wv.gridBlocks;
    const unsigned int ny = batch ? static_cast<unsigned int>(nScenes) : 1u;
    if (wv.shadowMode == kShadowSoft) {  // ring + point table + scene blob exceed the default 48 KB
        static SmemOptIn optIn;
        static const void* const fns[] = {reinterpret_cast<const void*>(k_wf_softshadow<false>),
                                          reinterpret_cast<const void*>(k_wf_softshadow<true>)};
        optIn.limit(fns, 2);
    }
    // a batch's queue counters are zeroed by the caller (one memset over all scenes)
    if (!batch) cudaMemsetAsync(wv.qCount, 0, 2 * sizeof(unsigned int) * (wv.levels + 3), stream);
    if (batch) k_wf_hit0<true><<<dim3(grid, ny), kWfThreads, fp.blob_bytes, stream>>>(fr, fp, list, wv, batch); else k_wf_hit0<false><<<dim3(grid, ny), kWfThreads, fp.blob_bytes, stream>>>(fr, fp, list, wv, batch);
    ++n;
    if (fr.max_bounces >= 0) {
        // depths 0 .. queueLevels-1: seed / shadow / shade over the queue of that depth;
        // depth queueLevels (if bounces go that deep): one tail launch that finishes every chain
        const int queued = std::min(wv.levels + 1, std::max(1, wv.queueLevels));
        for (int depth = 0; depth < queued; ++depth) {
            const int which = depth & 1;
            // Deeper queues are much shorter (~8 % of the primary hits at depth 1, then ~75 % of the
            // previous level), but blocks beyond a queue's end return at once, and a short queue
            // spread over every SM finishes sooner than one packed into a few resident blocks.
            const dim3 g(depth == 0 ? grid : std::max(1, grid / std::max(1, wv.deepGridDiv)), ny);
            if (wv.shadowMode == kShadowSoft) {
                // persistent blocks that claim chunks of 256 hits: no more of them than can be resident
                const dim3 gs(std::max(1, std::min<int>(g.x, wv.softGrid)), ny);
                const size_t smem = ((fp.blob_bytes + 15u) & ~15u) + sizeof(ShadowSmem);
                if (batch) k_wf_softshadow<true><<<gs, kWfThreads, smem, stream>>>(fr, fp, wv, which, depth, batch);
                else k_wf_softshadow<false><<<gs, kWfThreads, smem, stream>>>(fr, fp, wv, which, depth, batch);
                ++n;
            } else if (wv.shadowMode == kShadowHard) {
                if (batch) k_wf_hardshadow<true><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, which, depth, batch);
                else k_wf_hardshadow<false><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, which, depth, batch);
                ++n;
            }
            if (wv.shadowMode != kShadowInThread) {
                if (batch) k_wf_shade<true, true><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, which, depth, 0, batch); else k_wf_shade<true, false><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, which, depth, 0, batch);
            } else {
                if (batch) k_wf_shade<false, true><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, which, depth, 0, batch); else k_wf_shade<false, false><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, which, depth, 0, batch);
            }
            ++n;
        }
        if (queued <= wv.levels) {
            const dim3 g(std::max(1, grid / std::max(1, wv.deepGridDiv)), ny);
            if (batch) k_wf_shade<false, true><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, queued & 1, queued, 1, batch); else k_wf_shade<false, false><<<g, kWfThreads, fp.blob_bytes, stream>>>(fr, fp, wv, queued & 1, queued, 1, batch);
            ++n;
        }
    }
    const int lg = log2_pow2_le32(fr.spp);
    if (lg >= 0) {
        if (batch) k_wf_resolve_warp<true><<<dim3(grid, ny), kWfThreads, 0, stream>>>(fr, band, list, wv, lg, batch); else k_wf_resolve_warp<false><<<dim3(grid, ny), kWfThreads, 0, stream>>>(fr, band, list, wv, lg, batch);
    } else {
        if (batch) k_wf_resolve_pixel<true><<<dim3(grid, ny), kWfThreads, 0, stream>>>(fr, band, list, wv, batch); else k_wf_resolve_pixel<false><<<dim3(grid, ny), kWfThreads, 0, stream>>>(fr, band, list, wv, batch);
    }
    ++n;
    // pixels the queues could not take: megakernel, starting at the first slot beyond them
    // (never in a batch: its queues are sized for every pixel of a scene)
    if (!batch && wv.slotCapacity < list.capacity) {
        launch_shade(fr, fp, band, list, grid, groupCounter, wv.slotCapacity, stream);
        ++n;
    }
    if (launches) *launches += n;
}

}  // namespace mcskin
