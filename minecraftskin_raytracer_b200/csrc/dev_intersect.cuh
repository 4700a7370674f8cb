// dev_intersect.cuh — ray vs. the scene's boxes: intersectScene / intersectMesh /
// intersectAABB of the reference (src/raytracer/intersection.cpp:200-421) over
// pre-digested box records held in shared memory.
//
// What is kept bit-for-bit: slab order x,y,z; the 1e-8
// parallel test; strict comparisons so ties go to the lowest axis; the
// origin-inside case taking the exit face with its OUTWARD normal; face table, UV
// orientation and clamping; truncating nearest-texel lookup; alpha == 0 pass-through
// for inner boxes and the exit-face fallback (flipped normal) for outer boxes;
// world->local = undo rotZ then undo rotX about the pivot with the pivot subtracted
// and re-added around EACH step; renormalised local direction; t recomputed as
// dot(p_world - o, d) for posed boxes; closest hit on strict t < best in box order.
//
// What is restructured.  Every query runs in two phases:
//   1. a branch-free REJECT pass over all boxes: the three slab intervals via FMNMX,
//      tmin = max of the near values, tmax = min of the far values, and the reference's
//      own rejection predicate (tmin > tmax || tmax < 0).  For unposed boxes the values are
//      the same floats the reference compares (min/max of the same two products), so the
//      set of surviving boxes is identical; it only skips the bookkeeping of which axis
//      won.  Posed boxes are tested against a world-space box around their rotated corners,
//      inflated far beyond rounding, i.e. conservatively.  Survivors (typically 0-2 of 12)
//      go into a bit mask.
//   2. the EXACT evaluation (entry/exit axis with the reference's tie rules, face, UV,
//      texel, alpha rules, pose back-transform) for the boxes in the mask, in box order.
// Bounds, pose sines/cosines and face windows are read from the record instead of being
// recomputed per ray; the winner's normal and texel colour are materialised once after
// the loop; shadow rays stop at the first occluder (same boolean as "closest hit closer
// than the light") and reject boxes that start beyond the light.
#pragma once
#include "dev_math.cuh"
#include "dev_types.cuh"

namespace mcskin {

struct Ray {
    V3 o, d;
};

// Result of the closest-hit query (HitResult, src/scene/triangle.h:19-26, plus ids).
struct Hit {
    float t;
    V3 p;       // world-space hit point
    int box;    // -1 = miss
    int face;   // reference face index 0..5
    int texel;  // index into the texel pool
    bool flip;  // outer-layer exit-face hit: normal = -faceNormal, isOuterLayer forced true
};

// Scene as the kernels see it: box records in shared memory (see SceneBlob in
// dev_types.cuh), texels in global memory through the read-only path.
struct SceneView {
    const float4* __restrict__ lo;      // reject-pass bounds: xyz = min (the box, or a world box around a posed one), w = flags
    const float4* __restrict__ hi;      // xyz = max
    const int4* __restrict__ rect;      // per box: pixels a pinhole ray must pass through to reach it (or null)
    const DevBox* __restrict__ boxes;   // full records
    const float4* __restrict__ texels;
    int n_boxes;
    uint32_t posed_mask;   // among boxes 0..31: never pre-rejected (none today; posed boxes carry world bounds)
    uint32_t usable_mask;  // among boxes 0..31: exist and have triangles
    uint32_t opaque_mask;  // among boxes 0..31: unposed and without see-through texels (kBoxOpaque)
    uint32_t rotated_mask; // among boxes 0..31: posed (kBoxRotated); their lo / hi are loose world-space bounds
    uint32_t opaque_posed_mask;  // among boxes 0..31: posed and without see-through texels
    uint32_t root_mask;    // among boxes 0..31: not enclosed by another box (0: no two-level reject pass); hi[i].w of a
                           // root = the boxes it encloses
};

__device__ __forceinline__ V3 face_normal(int face) {
    // intersection.cpp:86-122
    switch (face) {
        case 0: return mk3(0.0f, 0.0f, -1.0f);
        case 1: return mk3(0.0f, 0.0f, 1.0f);
        case 2: return mk3(1.0f, 0.0f, 0.0f);
        case 3: return mk3(-1.0f, 0.0f, 0.0f);
        case 4: return mk3(0.0f, 1.0f, 0.0f);
        default: return mk3(0.0f, -1.0f, 0.0f);
    }
}
__device__ __forceinline__ int face_index(int axis, bool negSide) {
    if (axis == 2) return negSide ? 0 : 1;
    if (axis == 0) return negSide ? 3 : 2;
    return negSide ? 5 : 4;
}

// rotatePoint (intersection.cpp:12-37) with host-evaluated cos/sin.
__device__ __forceinline__ V3 rotate_about(V3 point, V3 pivot, bool doX, float cX, float sX, bool doZ, float cZ,
                                           float sZ) {
    V3 p = point - pivot;
    if (doX) {
        const float ny = p.y * cX - p.z * sX;
        const float nz = p.y * sX + p.z * cX;
        p.y = ny;
        p.z = nz;
    }
    if (doZ) {
        const float nx = p.x * cZ - p.y * sZ;
        const float ny = p.x * sZ + p.y * cZ;
        p.x = nx;
        p.y = ny;
    }
    return p + pivot;
}

// One slab axis (intersection.cpp:221-250 and the exit-face recomputation :265-285).
struct Slab {
    float tmin, tmax;
    int axis, exitAxis;
    bool neg, exitNeg;
};
// inv = 1.0f / d, evaluated once per ray (the reference divides per box; same operands, same result)
template <int AXIS>
__device__ __forceinline__ bool slab_axis(Slab& s, float o, float d, float inv, float lo, float hi) {
    if (fabsf(d) < 1e-8f) return !(o < lo || o > hi);
    const float t0 = (lo - o) * inv;
    const float t1 = (hi - o) * inv;
    const bool swapped = t0 > t1;
    const float tn = swapped ? t1 : t0;
    const float tf = swapped ? t0 : t1;
    if (tn > s.tmin) {
        s.tmin = tn;
        s.axis = AXIS;
        s.neg = !swapped;
    }
    if (tf < s.tmax) {  // std::min(tmax, t1) and the strict '<' of the exit-face loops
        s.tmax = tf;
        s.exitAxis = AXIS;
        s.exitNeg = swapped;
    }
    return true;
}

// computeFaceUV + TextureRegion::sample addressing (intersection.cpp:136-196,
// texture_region.h:19-26) -> texel pool index.  The two in-plane coordinates are picked
// first so the (IEEE) divisions exist once: Z faces use (x, y), X faces (z, y), Y faces (x, z).
// a / b rounded to nearest from r = RN(1/b): two Newton steps on the quotient with exact (fused)
// residuals.  After the first, the quotient is within a rounding of a/b; the second is then the
// correctly rounded a/b (Markstein's theorem; the host clears kBoxRecip for the one excluded
// family of divisors, significands of all ones).  Inputs here are box-local coordinates of
// order 1..100: no overflow, and a residual of an exact zero stays zero.
__device__ __forceinline__ float div_exact(float a, float b, float r) {
    float q = a * r;
    q = fmaf(fmaf(-q, b, a), r, q);
    return fmaf(fmaf(-q, b, a), r, q);
}
__device__ __forceinline__ int face_texel(const DevBox& bx, V3 p, int axis, bool negSide, int face) {
    const float pa = (axis == 0) ? p.z : p.x;
    const float la0 = (axis == 0) ? bx.lo[2] : bx.lo[0];
    const float sa = (axis == 0) ? bx.size[2] : bx.size[0];
    const float ra = (axis == 0) ? bx.inv_size_z : bx.inv_size_x;
    const float pb = (axis == 1) ? p.z : p.y;
    const float lb0 = (axis == 1) ? bx.lo[2] : bx.lo[1];
    const float sb = (axis == 1) ? bx.size[2] : bx.size[1];
    const float rb = (axis == 1) ? bx.inv_size_z : bx.inv_size_y;
    float la, lb;   // localX (or localZ on X faces), localY (or localZ on Y faces)
    if (bx.flags & kBoxRecip) {
        la = div_exact(pa - la0, sa, ra);
        lb = div_exact(pb - lb0, sb, rb);
    } else {
        la = (pa - la0) / sa;
        lb = (pb - lb0) / sb;
    }
    float u, v;
    if (axis == 2) {
        u = negSide ? 1.0f - la : la;
        v = 1.0f - lb;
    } else if (axis == 0) {
        u = !negSide ? 1.0f - la : la;
        v = 1.0f - lb;
    } else {
        u = la;
        v = !negSide ? lb : 1.0f - lb;
    }
    u = clamp01(u);
    v = clamp01(v);
    const int2 ft = bx.face[face];
    const int w = ft.y & 0xffff;
    const int h = ft.y >> 16;
    int x = __float2int_rz(u * static_cast<float>(w));
    int y = __float2int_rz(v * static_cast<float>(h));
    x = max(0, min(x, w - 1));
    y = max(0, min(y, h - 1));
    return ft.x + y * w + x;
}

// ---- phase 1: candidate mask ------------------------------------------------------
// Per-ray constants of the reject pass.
struct RayPre {
    V3 inv;          // 1/d per axis (IEEE division, like the reference's invD)
    bool parallel;   // some |d_i| < 1e-8: the reject pass is skipped, every box is a candidate
};
__device__ __forceinline__ RayPre ray_pre(const Ray& r) {
    RayPre p;
    p.parallel = fabsf(r.d.x) < 1e-8f || fabsf(r.d.y) < 1e-8f || fabsf(r.d.z) < 1e-8f;
    p.inv = mk3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    return p;
}

struct BoxHit {
    float t;
    V3 p;  // in the space of the ray handed to box_test
    int face, texel;
    bool flip;
};

__device__ __forceinline__ float rcp_fast(float x) {  // approximate reciprocal: the margins absorb its error
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// World -> box space of a posed box for CONSERVATIVE tests only (the exact path is mesh_test): undo rotZ, then rotX,
// about the pivot.  Same formulas, but nothing downstream depends on their last bits — every user adds margins
// four orders of magnitude above their rounding.
__device__ __forceinline__ V3 to_box_space(const DevBox& bx, V3 p, bool isPoint) {
    const V3 pivot = isPoint ? ld3(bx.pivot) : mk3(0.0f, 0.0f, 0.0f);
    V3 q = p - pivot;
    if (bx.flags & kBoxRotZ) {
        const float nx = q.x * bx.inv_cz - q.y * bx.inv_sz, ny = q.x * bx.inv_sz + q.y * bx.inv_cz;
        q.x = nx;
        q.y = ny;
    }
    if (bx.flags & kBoxRotX) {
        const float ny = q.y * bx.inv_cx - q.z * bx.inv_sx, nz = q.y * bx.inv_sx + q.z * bx.inv_cx;
        q.y = ny;
        q.z = nz;
    }
    return q + pivot;
}

// A cheap second opinion on a posed box that survived the test against its (loose) world-space bounds: the slab
// test again in the box's own space, in approximate arithmetic (the transform and the approximate reciprocals are
// good to ~1e-5 here), against the box GROWN by 1e-2 — a ray that misses that surely misses the box — and, for
// occlusion queries on boxes without see-through texels, against the box SHRUNK by 1e-2: a ray that passes through
// that, and leaves even the grown box before `limit`, surely hits the box before the light, whatever the exact
// arithmetic's last bits.  Everything in between is left to the exact evaluation.
enum : int { kPosedUnknown = 0, kPosedMissed = 1, kPosedHitBeforeLimit = 2 };
template <bool LIMITED>
__device__ __forceinline__ int posed_box_verdict(const DevBox& bx, const Ray& ray, float limit, bool opaque) {
    constexpr float kPad = 1e-2f;
    const V3 o = to_box_space(bx, ray.o, true);
    const V3 d = to_box_space(bx, ray.d, false);  // |d| = 1 up to rounding: distances along it are world distances
    float tmin = -FLT_MAX, tmax = FLT_MAX;        // grown box
    float smin = -FLT_MAX, smax = FLT_MAX;        // shrunk box
    bool sure = LIMITED && opaque;
    const float os[3] = {o.x, o.y, o.z}, ds[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = bx.lo[k] - kPad, hi = bx.hi[k] + kPad;
        if (fabsf(ds[k]) < 1e-6f) {
            if (os[k] < lo || os[k] > hi) return kPosedMissed;
            sure = sure && os[k] > lo + 2.0f * kPad && os[k] < hi - 2.0f * kPad;
            continue;
        }
        const float inv = rcp_fast(ds[k]);
        const float a = (lo - os[k]) * inv, b = (hi - os[k]) * inv;
        tmin = fmaxf(tmin, fminf(a, b));
        tmax = fminf(tmax, fmaxf(a, b));
        if (LIMITED) {
            const float shift = 2.0f * kPad * fabsf(inv);  // the shrunk slab starts that much later and ends that much earlier
            smin = fmaxf(smin, fminf(a, b) + shift);
            smax = fminf(smax, fmaxf(a, b) - shift);
        }
    }
    // relative slack on the distances for the approximate reciprocal (2^-22) and |d| != 1
    const float slack = 1e-4f * (fabsf(tmin) + fabsf(tmax)) + 1e-3f;
    if (fmaxf(tmin, 0.0f) > tmax + slack) return kPosedMissed;
    if (LIMITED && !(tmin - slack < limit)) return kPosedMissed;
    // through the shrunk box in front of the origin, and out of the grown one before the light: the hit distance
    // (entry, or exit when the origin is inside) lies below tmax
    if (LIMITED && sure && fmaxf(smin, 0.0f) + slack < smax && tmax + slack < limit) return kPosedHitBeforeLimit;
    return kPosedUnknown;
}
template <bool LIMITED>
__device__ __forceinline__ bool posed_box_missed(const DevBox& bx, const Ray& ray, float limit) {
    return posed_box_verdict<LIMITED>(bx, ray, limit, false) == kPosedMissed;
}

// Clears from `mask` (boxes 0..31 that survived the world-space reject pass) the posed boxes the ray surely misses.
template <bool LIMITED>
__device__ __forceinline__ uint32_t drop_missed_posed(const SceneView& sc, const Ray& ray, uint32_t mask, float limit) {
    if (!kPosedScenes) return mask;  // (this build of the kernels is only ever given scenes without poses)
    uint32_t todo = mask & sc.rotated_mask;
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1u;
        if (posed_box_missed<LIMITED>(sc.boxes[i], ray, limit)) mask &= ~(1u << i);
    }
    return mask;
}


// intersectAABB (intersection.cpp:200-371) against one box, ray already in box space.
__device__ __forceinline__ bool box_test(const SceneView& sc, const DevBox& bx, V3 o, V3 d, V3 inv, BoxHit& out) {
    Slab s;
    s.tmin = -FLT_MAX;
    s.tmax = FLT_MAX;
    s.axis = 0;
    s.exitAxis = 0;
    s.neg = false;
    s.exitNeg = false;
    if (!slab_axis<0>(s, o.x, d.x, inv.x, bx.lo[0], bx.hi[0])) return false;
    if (!slab_axis<1>(s, o.y, d.y, inv.y, bx.lo[1], bx.hi[1])) return false;
    if (!slab_axis<2>(s, o.z, d.z, inv.z, bx.lo[2], bx.hi[2])) return false;
    // the reference tests this after every axis; tmin only grows and tmax only shrinks,
    // so testing once at the end rejects exactly the same rays
    if (s.tmin > s.tmax || s.tmax < 0.0f) return false;

    float tHit = s.tmin;
    int axis = s.axis;
    bool negSide = s.neg;
    if (tHit < 0.0f) {  // origin inside: leave through the exit face (intersection.cpp:254-288)
        tHit = s.tmax;
        axis = s.exitAxis;
        negSide = s.exitNeg;
    }
    // First the face found above; if its texel is fully transparent and the box is an outer
    // layer, once more for the exit face at tmax with the normal flipped (intersection.cpp:311-361).
    // Written as a two-trip loop so the face/UV/texel code exists once.
    bool flip = false;
    for (;;) {
        const V3 p = o + d * tHit;
        const int face = face_index(axis, negSide);
        const int texel = face_texel(bx, p, axis, negSide, face);
        const float alpha = __ldg(&sc.texels[texel].w);
        const bool opaque = flip ? (alpha > 0.0f) : !(alpha == 0.0f);
        if (opaque) {
            out.t = tHit;
            out.p = p;
            out.face = face;
            out.texel = texel;
            out.flip = flip;
            return true;
        }
        if (flip || !(bx.flags & kBoxOuter) || !(s.tmax > tHit)) return false;
        flip = true;
        tHit = s.tmax;
        axis = s.exitAxis;
        negSide = s.exitNeg;
    }
}

// intersectMesh (intersection.cpp:373-406): hit distance and point in WORLD space.
__device__ __forceinline__ bool mesh_test(const SceneView& sc, const int box, const Ray& ray, const RayPre& pre, BoxHit& out) {
    const DevBox& bx = sc.boxes[box];
    const uint32_t flags = bx.flags;
    if (flags & kBoxEmpty) return false;
    const bool rotated = kPosedScenes && (flags & kBoxRotated);
    const V3 pivot = ld3(bx.pivot);
    const bool doX = flags & kBoxRotX, doZ = flags & kBoxRotZ;
    V3 o = ray.o, d = ray.d, inv = pre.inv;
    if (rotated) {
        const V3 zero = mk3(0.0f, 0.0f, 0.0f);
        // rotatePoint(o, pivot, 0, -rotZ) then rotatePoint(., pivot, -rotX, 0); same for the direction about 0
        o = rotate_about(o, pivot, false, 0.0f, 0.0f, doZ, bx.inv_cz, bx.inv_sz);
        o = rotate_about(o, pivot, doX, bx.inv_cx, bx.inv_sx, false, 0.0f, 0.0f);
        d = rotate_about(d, zero, false, 0.0f, 0.0f, doZ, bx.inv_cz, bx.inv_sz);
        d = rotate_about(d, zero, doX, bx.inv_cx, bx.inv_sx, false, 0.0f, 0.0f);
        d = normalize3(d);
        inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    }
    if (!box_test(sc, bx, o, d, inv, out)) return false;
    if (rotated) {
        out.p = rotate_about(out.p, pivot, doX, bx.fwd_cx, bx.fwd_sx, doZ, bx.fwd_cz, bx.fwd_sz);
        out.t = dot3(out.p - ray.o, ray.d);
    }
    return true;
}

// Bit b set = box (base + b) may be hit: it survives the reference's slab rejection
// (unposed boxes) or is posed / not pre-testable.  With LIMITED, boxes whose slab entry is at
// or beyond `limit` are dropped too (their hit distance is >= the entry distance).
// The rejection predicate `tmin > tmax || tmax < 0` is evaluated as max(tmin, 0) > tmax, which
// is the same boolean for every non-NaN pair.
template <bool LIMITED>
__device__ __forceinline__ uint32_t candidate_mask(const SceneView& sc, const Ray& ray, const RayPre& pre, int base,
                                                   float limit) {
    const int n = min(32, sc.n_boxes - base);
    const uint32_t all = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
    uint32_t posed = sc.posed_mask, usable = sc.usable_mask;
    if (base != 0) {  // scenes with more than 32 meshes: derive the masks of this chunk
        posed = 0u;
        usable = 0u;
        for (int i = 0; i < n; ++i) {
            const uint32_t flags = __float_as_uint(sc.lo[base + i].w);
            if (!(flags & kBoxEmpty)) usable |= 1u << i;
        }
    }
    if (pre.parallel) return all & usable;
    if (base == 0 && sc.root_mask != 0u) {
        // Two levels: a box that lies well inside another one (an inner body part inside its outer layer) can only
        // be hit by rays that pass the enclosing box's test — the host records for every enclosing box ("root")
        // which boxes it contains (hi.w) — so the roots are tested first and the enclosed boxes only where their
        // root survives: 6 + ~1.5 tests per ray on a skin with every outer layer instead of 12.
        uint32_t survivors = 0u, enclosed = 0u;
        auto test = [&](int i, uint32_t* children) {
            const float4 L = sc.lo[i];
            const float4 H = sc.hi[i];
            const float ax = (L.x - ray.o.x) * pre.inv.x, bx = (H.x - ray.o.x) * pre.inv.x;
            const float ay = (L.y - ray.o.y) * pre.inv.y, by = (H.y - ray.o.y) * pre.inv.y;
            const float az = (L.z - ray.o.z) * pre.inv.z, bz = (H.z - ray.o.z) * pre.inv.z;
            const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            bool reject = fmaxf(tmin, 0.0f) > tmax;
            if (LIMITED) reject = reject || !(tmin < limit);
            *children = __float_as_uint(H.w);
            return !reject;
        };
        for (uint32_t m = sc.root_mask & usable & all; m; m &= m - 1u) {
            const int i = __ffs(m) - 1;
            uint32_t children;
            if (test(i, &children) || ((posed >> i) & 1u)) {
                survivors |= 1u << i;
                enclosed |= children;
            }
        }
        for (uint32_t m = enclosed & usable & all; m; m &= m - 1u) {
            const int i = __ffs(m) - 1;
            uint32_t unused;
            if (test(i, &unused) || ((posed >> i) & 1u)) survivors |= 1u << i;
        }
        return drop_missed_posed<LIMITED>(sc, ray, survivors, limit);
    }
    uint32_t rejected = 0u;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const float4 L = sc.lo[base + i];
        const float4 H = sc.hi[base + i];
        // same products the reference forms: (min - o) * invD and (max - o) * invD
        const float ax = (L.x - ray.o.x) * pre.inv.x, bx = (H.x - ray.o.x) * pre.inv.x;
        const float ay = (L.y - ray.o.y) * pre.inv.y, by = (H.y - ray.o.y) * pre.inv.y;
        const float az = (L.z - ray.o.z) * pre.inv.z, bz = (H.z - ray.o.z) * pre.inv.z;
        const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        bool reject = fmaxf(tmin, 0.0f) > tmax;
        if (LIMITED) reject = reject || !(tmin < limit);
        if (reject) rejected |= 1u << i;
    }
    uint32_t mask = (~rejected | posed) & usable & all;
    // posed boxes were tested against loose world-space bounds: ask again in their own space before the exact
    // evaluation (nothing to do — one uniform branch — in a scene without poses)
    if (base == 0) mask = drop_missed_posed<LIMITED>(sc, ray, mask, limit);
    return mask;
}

// The boxes a pinhole ray through pixel (px, py) can reach: those whose screen rectangle holds the
// pixel (conservative: projected bounds plus a 2-pixel margin, host_prep.cpp).  All boxes when the
// scene carries no rectangles (depth of field, a box behind the camera, more than 32 boxes).
__device__ __forceinline__ uint32_t pixel_box_mask(const SceneView& sc, int px, int py) {
    if (sc.rect == nullptr) return 0xffffffffu;
    uint32_t mask = 0u;
    const int n = min(32, sc.n_boxes);
    for (int i = 0; i < n; ++i) {
        const int4 r = sc.rect[i];
        if (px >= r.x && px <= r.z && py >= r.y && py <= r.w) mask |= 1u << i;
    }
    return mask;
}

// The reject pass over the boxes of `allow` only (first 32 boxes): survivors of the reference's slab rejection.
__device__ __forceinline__ uint32_t candidate_mask_among(const SceneView& sc, const Ray& ray, const RayPre& pre, uint32_t allow) {
    uint32_t todo = allow & sc.usable_mask;
    if (sc.n_boxes < 32) todo &= (1u << sc.n_boxes) - 1u;
    if (pre.parallel) return todo;
    uint32_t mask = todo, rejected = 0u;
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1u;
        const float4 L = sc.lo[i];
        const float4 H = sc.hi[i];
        const float ax = (L.x - ray.o.x) * pre.inv.x, bx = (H.x - ray.o.x) * pre.inv.x;
        const float ay = (L.y - ray.o.y) * pre.inv.y, by = (H.y - ray.o.y) * pre.inv.y;
        const float az = (L.z - ray.o.z) * pre.inv.z, bz = (H.z - ray.o.z) * pre.inv.z;
        const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        if (fmaxf(tmin, 0.0f) > tmax) rejected |= 1u << i;
    }
    return drop_missed_posed<false>(sc, ray, mask & (~rejected | sc.posed_mask), 0.0f);
}

// any_hit for a ray known to reach only the boxes of `allow` (see pixel_box_mask): a box outside the mask cannot be hit
__device__ __forceinline__ bool any_hit_among(const SceneView& sc, const Ray& ray, uint32_t allow) {
    const RayPre pre = ray_pre(ray);
    uint32_t mask = candidate_mask_among(sc, ray, pre, allow);
    // a surviving box without see-through texels is a hit (see occluded_among)
    if (!pre.parallel && (mask & sc.opaque_mask)) return true;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1u;
        BoxHit h;
        if (mesh_test(sc, b, ray, pre, h)) return true;
    }
    return false;
}

// intersectScene (intersection.cpp:408-421).
__device__ __forceinline__ Hit closest_hit(const SceneView& sc, const Ray& ray) {
    Hit best;
    best.t = FLT_MAX;
    best.box = -1;
    best.face = 0;
    best.texel = 0;
    best.flip = false;
    best.p = mk3(0.0f, 0.0f, 0.0f);
    const RayPre pre = ray_pre(ray);
    for (int base = 0; base < sc.n_boxes; base += 32) {
        uint32_t mask = candidate_mask<false>(sc, ray, pre, base, FLT_MAX);
        while (mask) {  // increasing box index: strict '<' keeps the reference's tie order
            const int b = base + __ffs(mask) - 1;
            mask &= mask - 1u;
            BoxHit h;
            if (mesh_test(sc, b, ray, pre, h) && h.t < best.t) {
                best.t = h.t;
                best.p = h.p;
                best.box = b;
                best.face = h.face;
                best.texel = h.texel;
                best.flip = h.flip;
            }
        }
    }
    return best;
}

// intersectMesh of one box (the reference's unit tests call it directly).
__device__ __forceinline__ Hit single_box_hit(const SceneView& sc, int b, const Ray& ray) {
    Hit best;
    best.t = 0.0f;
    best.box = -1;
    best.face = 0;
    best.texel = 0;
    best.flip = false;
    best.p = mk3(0.0f, 0.0f, 0.0f);
    BoxHit h;
    const RayPre pre = ray_pre(ray);
    if (b >= 0 && b < sc.n_boxes && mesh_test(sc, b, ray, pre, h)) {
        best.t = h.t;
        best.p = h.p;
        best.box = b;
        best.face = h.face;
        best.texel = h.texel;
        best.flip = h.flip;
    }
    return best;
}

// hit.hit of intersectScene, stopping at the first box that reports a hit.
__device__ __forceinline__ bool any_hit(const SceneView& sc, const Ray& ray) {
    const RayPre pre = ray_pre(ray);
    for (int base = 0; base < sc.n_boxes; base += 32) {
        uint32_t mask = candidate_mask<false>(sc, ray, pre, base, FLT_MAX);
        while (mask) {
            const int b = base + __ffs(mask) - 1;
            mask &= mask - 1u;
            BoxHit h;
            if (mesh_test(sc, b, ray, pre, h)) return true;
        }
    }
    return false;
}

// isInShadow's `hit.hit && hit.t < distToLight` (shading.cpp:23-25): the closest hit is
// nearer than the light iff some box reports a hit nearer than the light.  `allow` limits
// the first 32 boxes to a subset the caller has shown to be sufficient (bundle_box_mask).
__device__ __forceinline__ bool occluded_among(const SceneView& sc, const Ray& ray, const float dist, const uint32_t allow) {
    const RayPre pre = ray_pre(ray);
    for (int base = 0; base < sc.n_boxes; base += 32) {
        uint32_t mask;
        if (base == 0 && allow != 0xffffffffu) {
            // reject pass over the allowed boxes only (typically 1-3 of them)
            uint32_t todo = allow & sc.usable_mask;
            if (sc.n_boxes < 32) todo &= (1u << sc.n_boxes) - 1u;
            mask = todo;
            if (!pre.parallel) {
                uint32_t rejected = 0u;
                while (todo) {
                    const int i = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const float4 L = sc.lo[i];
                    const float4 H = sc.hi[i];
                    const float ax = (L.x - ray.o.x) * pre.inv.x, bx = (H.x - ray.o.x) * pre.inv.x;
                    const float ay = (L.y - ray.o.y) * pre.inv.y, by = (H.y - ray.o.y) * pre.inv.y;
                    const float az = (L.z - ray.o.z) * pre.inv.z, bz = (H.z - ray.o.z) * pre.inv.z;
                    const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
                    const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
                    if (fmaxf(tmin, 0.0f) > tmax || !(tmin < dist)) {
                        rejected |= 1u << i;
                    } else if ((sc.opaque_mask >> i) & 1u) {
                        // An unposed box with no see-through texel: intersectAABB reports a hit for every ray
                        // that passes its slab test, at tmin — or at tmax when the origin is inside
                        // (intersection.cpp:246-288); tmin / tmax above are the reference's own values.
                        if ((tmin < 0.0f ? tmax : tmin) < dist) return true;
                        rejected |= 1u << i;  // hit, but beyond the light
                    }
                }
                mask &= ~rejected | sc.posed_mask;
                if (kPosedScenes) {  // posed survivors: missed / surely hit before the light / ask the exact test
                    uint32_t posedTodo = mask & sc.rotated_mask;
                    while (posedTodo) {
                        const int i = __ffs(posedTodo) - 1;
                        posedTodo &= posedTodo - 1u;
                        const int verdict = posed_box_verdict<true>(sc.boxes[i], ray, dist, (sc.opaque_posed_mask >> i) & 1u);
                        if (verdict == kPosedHitBeforeLimit) return true;
                        if (verdict == kPosedMissed) mask &= ~(1u << i);
                    }
                }
            }
        } else {
            mask = candidate_mask<true>(sc, ray, pre, base, dist);
        }
        while (mask) {
            const int b = base + __ffs(mask) - 1;
            mask &= mask - 1u;
            BoxHit h;
            if (mesh_test(sc, b, ray, pre, h) && h.t < dist) return true;
        }
    }
    return false;
}
__device__ __forceinline__ bool occluded(const SceneView& sc, const Ray& ray, float dist) {
    return occluded_among(sc, ray, dist, 0xffffffffu);
}

// Which of the first 32 boxes can be touched by ANY segment from `from` to a point within
// `radius` of `to`?  (All shadow rays of one hit form such a bundle: same origin, end points
// on the light's disk.)  A point at fraction s of such a segment lies within s*radius of the
// point at fraction s of the CENTRE segment.  So if a segment of the bundle touches a box at
// fraction s, the centre segment is, at that s, inside the box grown by s*radius <= radius:
//   pass 1: clip the centre segment against the box grown by `radius` -> fractions [s0, s1];
//           empty -> no segment of the bundle reaches the box;
//   pass 2: every touching fraction is <= s1, so the box grown by only s1*radius must be
//           crossed as well — near the hit point (small s1) that is a much tighter test.
// Own arithmetic with generous margins (1 % + 0.01 on the growth, 1e-3 on the fractions):
// the mask only ever drops boxes no shadow ray of the hit can reach, so restricting
// occluded_among() to it cannot change a result.  Typical skins: ~2 of 12 boxes survive.
// Clips the segment a + s*d, s in [0,1], against the box [lo-grow, hi+grow]; inv = 1/d.
__device__ __forceinline__ bool bundle_clip(V3 a, V3 d, V3 inv, V3 lo, V3 hi, float grow, float* sEnd) {
    const float ax0 = (lo.x - grow - a.x) * inv.x, ax1 = (hi.x + grow - a.x) * inv.x;
    const float ay0 = (lo.y - grow - a.y) * inv.y, ay1 = (hi.y + grow - a.y) * inv.y;
    const float az0 = (lo.z - grow - a.z) * inv.z, az1 = (hi.z + grow - a.z) * inv.z;
    // an axis along which the segment does not move: inside the slab -> no constraint
    const bool px = fabsf(d.x) < 1e-6f, py = fabsf(d.y) < 1e-6f, pz = fabsf(d.z) < 1e-6f;
    const bool outX = px && (a.x < lo.x - grow || a.x > hi.x + grow);
    const bool outY = py && (a.y < lo.y - grow || a.y > hi.y + grow);
    const bool outZ = pz && (a.z < lo.z - grow || a.z > hi.z + grow);
    float s0 = 0.0f, s1 = 1.0f;
    if (!px) { s0 = fmaxf(s0, fminf(ax0, ax1)); s1 = fminf(s1, fmaxf(ax0, ax1)); }
    if (!py) { s0 = fmaxf(s0, fminf(ay0, ay1)); s1 = fminf(s1, fmaxf(ay0, ay1)); }
    if (!pz) { s0 = fmaxf(s0, fminf(az0, az1)); s1 = fminf(s1, fmaxf(az0, az1)); }
    *sEnd = s1;
    return !(outX || outY || outZ) && s0 <= s1 + 1e-3f;
}

// Can any segment from `from` to a point within `radius` of from + d touch the box [lo, hi]?  (See bundle_box_mask.)
// pad: how far the vectors may be off (0 for world-space boxes, whose bounds are the reference's own floats; the
// rounding allowance of to_box_space otherwise): the box is grown by it for the clips, and the release rule asks
// for an origin that much beyond a slab.
__device__ __forceinline__ bool bundle_reaches(V3 from, V3 d, V3 lo, V3 hi, float radius, float pad) {
    const float growFull = radius * 1.01f + 0.01f + pad;
    const V3 inv = mk3(rcp_fast(d.x), rcp_fast(d.y), rcp_fast(d.z));
    if (fabsf(d.x) < 1e-6f || fabsf(d.y) < 1e-6f || fabsf(d.z) < 1e-6f) {
        // the centre segment does not move along some axis (rare): the general clip
        float sEnd;
        if (!bundle_clip(from, d, inv, lo, hi, growFull, &sEnd)) return false;
        const float reach = fminf(fmaxf(sEnd + 1e-3f, 0.0f), 1.0f);
        float unused;
        return bundle_clip(from, d, inv, lo, hi, reach * radius * 1.01f + 0.01f + pad, &unused);
    }
    // A ray that starts beyond a slab of the box and moves further away along that axis cannot
    // enter the box (the reference's own slab test gives tmax < 0 for it).  Every ray of the
    // bundle has a direction within growFull of d, so this holds for all of them at once.  It is
    // what releases the box a hit lies ON (origin = hit point + normal * 1e-3, just outside it):
    // without it that box would survive every distance-based clip.
    if ((from.x > hi.x + pad && d.x > growFull) || (from.x < lo.x - pad && d.x < -growFull) ||
        (from.y > hi.y + pad && d.y > growFull) || (from.y < lo.y - pad && d.y < -growFull) ||
        (from.z > hi.z + pad && d.z > growFull) || (from.z < lo.z - pad && d.z < -growFull))
        return false;
    // Every axis moves: growing a slab by g widens its parameter interval by g*|1/d| at both ends,
    // so the slab fractions of the ungrown box are computed once and serve both passes.
    const V3 ainv = mk3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
    const V3 g1 = ainv * growFull;
    const float ax = (lo.x - from.x) * inv.x, bx = (hi.x - from.x) * inv.x;
    const float ay = (lo.y - from.y) * inv.y, by = (hi.y - from.y) * inv.y;
    const float az = (lo.z - from.z) * inv.z, bz = (hi.z - from.z) * inv.z;
    const float nx = fminf(ax, bx), fx = fmaxf(ax, bx);
    const float ny = fminf(ay, by), fy = fmaxf(ay, by);
    const float nz = fminf(az, bz), fz = fmaxf(az, bz);
    const float s0 = fmaxf(fmaxf(nx - g1.x, ny - g1.y), fmaxf(nz - g1.z, 0.0f));
    const float s1 = fminf(fminf(fx + g1.x, fy + g1.y), fminf(fz + g1.z, 1.0f));
    if (!(s0 <= s1 + 1e-3f)) return false;
    const float reach = fminf(fmaxf(s1 + 1e-3f, 0.0f), 1.0f);
    const V3 g2 = ainv * (reach * radius * 1.01f + 0.01f + pad);
    const float t0 = fmaxf(fmaxf(nx - g2.x, ny - g2.y), fmaxf(nz - g2.z, 0.0f));
    const float t1 = fminf(fminf(fx + g2.x, fy + g2.y), fminf(fz + g2.z, 1.0f));
    return t0 <= t1 + 1e-3f;
}

// A posed box again in its own space, where it is tight (sc.lo / sc.hi hold loose world-space bounds of it): a
// rotation keeps distances, so the bundle is the same bundle there.  Above all this releases the box the hit lies on,
// which its world-space bounds — the hit is inside them — never do: without it every shadow ray of a hit on a posed
// limb ran the limb's full pose transform and slab test, and its outer layer's.
// (the shadow origin sits 1e-3 off the face it left, the rounding of the transform is ~1e-5: allow 1e-4)
__device__ __forceinline__ bool posed_bundle_reaches(const DevBox& pb, V3 from, V3 d, float radius) {
    return bundle_reaches(to_box_space(pb, from, true), to_box_space(pb, d, false), ld3(pb.lo), ld3(pb.hi), radius, 1e-4f);
}

__device__ __forceinline__ uint32_t drop_unreached_posed(const SceneView& sc, V3 from, V3 d, float radius, uint32_t mask) {
    if (!kPosedScenes) return mask;
    uint32_t todo = mask & sc.rotated_mask;
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1u;
        if (!posed_bundle_reaches(sc.boxes[i], from, d, radius)) mask &= ~(1u << i);
    }
    return mask;
}


__device__ __forceinline__ uint32_t bundle_box_mask(const SceneView& sc, V3 from, V3 to, float radius) {
    const int n = min(32, sc.n_boxes);
    const V3 d = to - from;
    const V3 inv = mk3(rcp_fast(d.x), rcp_fast(d.y), rcp_fast(d.z));
    const float growFull = radius * 1.01f + 0.01f;
    uint32_t mask = 0u;
    if (fabsf(d.x) < 1e-6f || fabsf(d.y) < 1e-6f || fabsf(d.z) < 1e-6f) {
        // the centre segment does not move along some axis (rare): the general clip
        for (int i = 0; i < n; ++i) {
            // sc.lo / sc.hi: the box itself, or for a posed box a world-space box that contains it
            const float4 L = sc.lo[i];
            const float4 H = sc.hi[i];
            if (__float_as_uint(L.w) & kBoxEmpty) continue;
            const V3 lo = mk3(L.x, L.y, L.z), hi = mk3(H.x, H.y, H.z);
            float sEnd;
            if (!bundle_clip(from, d, inv, lo, hi, growFull, &sEnd)) continue;
            const float reach = fminf(fmaxf(sEnd + 1e-3f, 0.0f), 1.0f);
            float unused;
            if (!bundle_clip(from, d, inv, lo, hi, reach * radius * 1.01f + 0.01f, &unused)) continue;
            mask |= 1u << i;
        }
        return drop_unreached_posed(sc, from, d, radius, mask);
    }
    // Every axis moves: growing a slab by g widens its parameter interval by g*|1/d| at both ends,
    // so the slab fractions of the ungrown box are computed once and serve both passes.
    const V3 ainv = mk3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
    const V3 g1 = ainv * growFull;
    auto reaches = [&](int i, uint32_t* children) {
        const float4 L = sc.lo[i];
        const float4 H = sc.hi[i];
        *children = __float_as_uint(H.w);
        if (__float_as_uint(L.w) & kBoxEmpty) return false;
        // A ray that starts beyond a slab of the box and moves further away along that axis cannot
        // enter the box (the reference's own slab test gives tmax < 0 for it).  Every ray of the
        // bundle has a direction within growFull of d, so this holds for all of them at once.  It is
        // what releases the box a hit lies ON (origin = hit point + normal * 1e-3, just outside it):
        // without it that box would survive every distance-based clip.
        if ((from.x > H.x && d.x > growFull) || (from.x < L.x && d.x < -growFull) ||
            (from.y > H.y && d.y > growFull) || (from.y < L.y && d.y < -growFull) ||
            (from.z > H.z && d.z > growFull) || (from.z < L.z && d.z < -growFull))
            return false;
        const float ax = (L.x - from.x) * inv.x, bx = (H.x - from.x) * inv.x;
        const float ay = (L.y - from.y) * inv.y, by = (H.y - from.y) * inv.y;
        const float az = (L.z - from.z) * inv.z, bz = (H.z - from.z) * inv.z;
        const float nx = fminf(ax, bx), fx = fmaxf(ax, bx);
        const float ny = fminf(ay, by), fy = fmaxf(ay, by);
        const float nz = fminf(az, bz), fz = fmaxf(az, bz);
        const float s0 = fmaxf(fmaxf(nx - g1.x, ny - g1.y), fmaxf(nz - g1.z, 0.0f));
        const float s1 = fminf(fminf(fx + g1.x, fy + g1.y), fminf(fz + g1.z, 1.0f));
        if (!(s0 <= s1 + 1e-3f)) return false;
        const float reach = fminf(fmaxf(s1 + 1e-3f, 0.0f), 1.0f);
        const V3 g2 = ainv * (reach * radius * 1.01f + 0.01f);
        const float t0 = fmaxf(fmaxf(nx - g2.x, ny - g2.y), fmaxf(nz - g2.z, 0.0f));
        const float t1 = fminf(fminf(fx + g2.x, fy + g2.y), fminf(fz + g2.z, 1.0f));
        return t0 <= t1 + 1e-3f;
    };
    if (sc.root_mask != 0u) {
        // Two levels, as in candidate_mask: a box enclosed by another one (an inner body part inside its outer layer, with
        // room to spare on every side) lies inside everything this test grows the enclosing box to, and beyond every
        // slab the bundle leaves the enclosing box behind — a bundle that cannot reach the root cannot reach it either.
        const uint32_t all = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
        uint32_t enclosed = 0u;
        for (uint32_t m = sc.root_mask & all; m; m &= m - 1u) {
            const int i = __ffs(m) - 1;
            uint32_t children;
            if (reaches(i, &children)) {
                mask |= 1u << i;
                enclosed |= children;
            }
        }
        for (uint32_t m = enclosed & all; m; m &= m - 1u) {
            const int i = __ffs(m) - 1;
            uint32_t unused;
            if (reaches(i, &unused)) mask |= 1u << i;
        }
        return drop_unreached_posed(sc, from, d, radius, mask);
    }
    for (int i = 0; i < n; ++i) {
        uint32_t unused;
        if (reaches(i, &unused)) mask |= 1u << i;
    }
    return drop_unreached_posed(sc, from, d, radius, mask);
}


// Normal of the winning hit (intersection.cpp:355,366,399-401).
__device__ __forceinline__ V3 hit_normal(const SceneView& sc, const Hit& h) {
    V3 n = face_normal(h.face);
    if (h.flip) n = n * -1.0f;
    const DevBox& bx = sc.boxes[h.box];
    if (kPosedScenes && (bx.flags & kBoxRotated)) {
        n = rotate_about(n, mk3(0.0f, 0.0f, 0.0f), bx.flags & kBoxRotX, bx.fwd_cx, bx.fwd_sx, bx.flags & kBoxRotZ,
                         bx.fwd_cz, bx.fwd_sz);
        n = normalize3(n);
    }
    return n;
}
__device__ __forceinline__ float4 hit_texel(const SceneView& sc, const Hit& h) { return __ldg(&sc.texels[h.texel]); }
__device__ __forceinline__ bool hit_is_outer(const SceneView& sc, const Hit& h) {
    return h.flip || (sc.boxes[h.box].flags & kBoxOuter);
}

// Conservative reject against the inflated bounds of the whole figure.
__device__ __forceinline__ bool misses_cull_box(const DevFrame& fr, const Ray& ray) {
    if (!fr.cull_valid) return false;
    Slab s;
    s.tmin = -FLT_MAX;
    s.tmax = FLT_MAX;
    s.axis = s.exitAxis = 0;
    s.neg = s.exitNeg = false;
    if (!slab_axis<0>(s, ray.o.x, ray.d.x, 1.0f / ray.d.x, fr.cull_lo[0], fr.cull_hi[0])) return true;
    if (!slab_axis<1>(s, ray.o.y, ray.d.y, 1.0f / ray.d.y, fr.cull_lo[1], fr.cull_hi[1])) return true;
    if (!slab_axis<2>(s, ray.o.z, ray.d.z, 1.0f / ray.d.z, fr.cull_lo[2], fr.cull_hi[2])) return true;
    return s.tmin > s.tmax || s.tmax < 0.0f;
}

}  // namespace mcskin
