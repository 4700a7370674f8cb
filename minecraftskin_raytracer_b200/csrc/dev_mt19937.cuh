// dev_mt19937.cuh — std::mt19937 + libstdc++'s uniform_real_distribution<float>(0,1)
// on the device, in the two access patterns the reference has:
//
//  (a) one engine per TILE, seeded tile.y*width + tile.x, consumed sequentially in
//      pixel-row-major, sample-minor order (tile_renderer.cpp:78-104).  A CTA owns a
//      tile and regenerates the stream 624 words at a time with a three-phase
//      parallel twist, dropping canonical floats into a shared-memory ring in step
//      with the sample loop (TileStream).
//  (b) a FRESH engine per shaded hit / AO query whose first 2N outputs are used
//      (shading.cpp:43-50, raytracer.cpp:50-56).  Output j < 227 of a fresh engine
//      only needs words j, j+1 and j+397 of the Knuth-LCG seeding sequence, so two
//      LCG cursors 397 words apart stream them out of registers with no state array
//      (FreshStream); more than 227 outputs fall back to a full state in local memory.
//
// Float mapping (libstdc++ 13 bits/random.tcc:3349-3381, generate_canonical<float,24>
// with a 32-bit engine): one engine call, float(u32) * 2^-32 with the conversion
// rounded to nearest, and a result >= 1 replaced by nextafterf(1,0) = 0x3f7fffff.
#pragma once
#include "dev_math.cuh"

namespace mcskin {

constexpr int kMtN = 624;
constexpr int kMtM = 397;

__device__ __forceinline__ uint32_t mt_lcg(uint32_t prev, uint32_t i) {
    return 1812433253u * (prev ^ (prev >> 30)) + i;
}
// The twist: far ^ (y >> 1) ^ (y odd ? 0x9908b0df : 0), y = top bit of cur | low 31 bits of next.
// Spelled so that it compiles to five instructions (bit-select LOP3, AND, IMAD for the conditional
// constant — on the FMA pipe —, SHF, three-input XOR) instead of the seven of the plain C form.
__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t next, uint32_t far) {
    uint32_t y, mag;
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xE4;" : "=r"(y) : "r"(cur), "r"(next));  // (cur & m) | (next & ~m)
    asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, 1;\n\tmul.lo.u32 %0, t, 0x9908b0df;\n\t}" : "=r"(mag) : "r"(next));
    return far ^ (y >> 1) ^ mag;
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__device__ __forceinline__ float mt_canonical(uint32_t u) {
    const float r = __uint2float_rn(u) * 2.3283064365386963e-10f;  // exact scaling by 2^-32
    return fminf(r, __uint_as_float(0x3f7fffffu));  // r <= 1; only r == 1 is replaced
}

// ---- counter-based streams (McConfig::rng_mode 1; the MCSKIN_COUNTER_RNG build) -----------------
// Word k of the stream seeded s is lowbias32(lowbias32(s ^ 0x9e3779b9) + k) (mc_rng_counter_word in mcskin_cuda.h):
// no engine state, no seeding recurrence; seeds, number and order of draws and the float mapping stay the reference's.
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t counter_key(uint32_t seed) { return lowbias32(seed ^ 0x9e3779b9u); }

// ---- (b) fresh engine, first outputs only -------------------------------------
// Word 397 of the seeding sequence, given word 1: 396 dependent LCG steps.  Kept as a
// short rolled loop: fully unrolling it (19 KB of straight-line code) measured 20 % slower
// on B200 — instruction-cache misses cost more than the loop counter.
__device__ __forceinline__ uint32_t mt_seed_word397(uint32_t word1) {
    uint32_t x = word1;
#pragma unroll 12
    for (uint32_t i = 2u; i <= static_cast<uint32_t>(kMtM); ++i) x = mt_lcg(x, i);
    return x;
}

// Word 397 for the seeding kernel, as straight-line code: the step index becomes an immediate addend of the
// IMAD, so a step is SHF + LOP3 + IMAD (19 KB of code, but every warp of the kernel walks it in step).
// `one` must be 1 at run time: it keeps the multiplier in a register, so the compiler emits the
// register-times-register IMAD with an immediate addend.  Measured on B200 against: the rolled loop with a
// register index (4 instructions per step, ALU pipe saturated), and forms that move the shift to the FMA
// pipe as IMAD.HI (slower: IMAD.HI issues at a lower rate than it relieves the ALU pipe).
__device__ __forceinline__ uint32_t mt_seed_word397_balanced(uint32_t word1, uint32_t one) {
    uint32_t x = word1;
    const uint32_t a = 1812433253u * one;
#pragma unroll
    for (int k = 2; k <= kMtM; ++k) {
        const uint32_t y = x ^ (x >> 30);
        x = y * a + static_cast<uint32_t>(k);
    }
    return x;
}

struct FreshStream {
    uint32_t cur, nxt, far;
    uint32_t j;

    // `one` must be 1 at run time (see mt_seed_word397_balanced)
    __device__ __forceinline__ void seed_balanced(uint32_t s, uint32_t one) {
        if (kCounterRng) {  // cur = the stream's key, j = the next word
            cur = counter_key(s);
            nxt = far = 0u;
            j = 0u;
            return;
        }
        cur = s;
        nxt = mt_lcg(s, 1u);
        far = mt_seed_word397_balanced(nxt, one);
        j = 0u;
    }

    // The same with a memo of the seeding recurrence.  Word 397 of the seeded state is a function of the seed alone, and
    // traceRay's shadow seeds (raytracer.cpp:110-112: a weighted sum of the hit's coordinates cast to unsigned) are
    // small integers — a figure's hits fall into a few million of them, every one used again and again, by the next
    // sample, the next frame, the next scene.  cache: kSeedMemoEntries pairs (seed ^ kSeedMemoTag, word 397), zeroed
    // once (the tag is outside the window, so a zero entry matches no seed); an 8-byte entry is stored and loaded
    // whole (one STG.E.64 / LDG.E.64 at L2, checked in the SASS), and every writer of an entry writes the same
    // pair, so no ordering is needed.  Seeds outside the window, or a null cache: the recurrence.
    __device__ __forceinline__ void seed_memo(uint32_t s, uint32_t one, uint2* cache) {
        if (kCounterRng || !cache) {
            seed_balanced(s, one);
            return;
        }
        const uint32_t idx = s + kSeedMemoOffset;  // (wraps: seeds just below zero come from negative sums)
        if (idx < kSeedMemoEntries) {
            const uint2 e = __ldcg(cache + idx);
            if (e.x == (s ^ kSeedMemoTag)) {
                cur = s;
                nxt = mt_lcg(s, 1u);
                far = e.y;
                j = 0u;
                return;
            }
        }
        seed_balanced(s, one);
        if (idx < kSeedMemoEntries) __stcg(cache + idx, make_uint2(s ^ kSeedMemoTag, far));
    }

    __device__ __forceinline__ void seed(uint32_t s) {
        if (kCounterRng) {
            cur = counter_key(s);
            nxt = far = 0u;
            j = 0u;
            return;
        }
        cur = s;
        nxt = mt_lcg(s, 1u);
        far = mt_seed_word397(nxt);
        j = 0u;
    }
    // valid for the first kMtN - kMtM = 227 calls
    __device__ __forceinline__ float next() {
        if (kCounterRng) return mt_canonical(lowbias32(cur + j++));
        const uint32_t v = mt_mix(cur, nxt, far);
        cur = nxt;
        nxt = mt_lcg(nxt, j + 2u);
        far = mt_lcg(far, j + static_cast<uint32_t>(kMtM) + 1u);
        ++j;
        return mt_canonical(mt_temper(v));
    }
};
constexpr int kFreshStreamMaxDraws = kMtN - kMtM;

// Full engine in local memory: only for shadowSamples/aoSamples > 113.
struct LocalEngine {
    uint32_t s[kMtN];
    int idx;
    __device__ void seed(uint32_t v) {
        if (kCounterRng) {
            s[0] = counter_key(v);
            idx = 0;
            return;
        }
        s[0] = v;
        for (int i = 1; i < kMtN; ++i) s[i] = mt_lcg(s[i - 1], static_cast<uint32_t>(i));
        idx = kMtN;
    }
    __device__ float next() {
        if (kCounterRng) return mt_canonical(lowbias32(s[0] + static_cast<uint32_t>(idx++)));
        if (idx >= kMtN) {
            for (int i = 0; i < kMtN; ++i) {
                const int i1 = (i + 1 == kMtN) ? 0 : i + 1;
                const int im = (i + kMtM >= kMtN) ? i + kMtM - kMtN : i + kMtM;
                s[i] = mt_mix(s[i], s[i1], s[im]);
            }
            idx = 0;
        }
        return mt_canonical(mt_temper(s[idx++]));
    }
};

// ---- (a) per-tile stream, CTA-cooperative ----------------------------------------
constexpr int kRingSize = 2048;  // floats; >= 4 draws * 256 samples + one 624-word block
constexpr int kRingMask = kRingSize - 1;

struct TileStreamSmem {
    uint32_t state[2][kMtN];  // ping-pong so a phase never overwrites a word another thread still reads
    float ring[kRingSize];    // canonical floats, absolute stream index & kRingMask
};

struct TileStream {
    TileStreamSmem* sm;
    int which;               // state[which] is the current block
    long long produced;      // stream words generated so far (multiple of 624)

    // All threads call; thread 0 runs the 623-step seeding recurrence.
    __device__ void seed(TileStreamSmem* smem, uint32_t seedValue) {
        sm = smem;
        which = 0;
        produced = 0;
        if (threadIdx.x == 0) {
            uint32_t x = seedValue;
            sm->state[0][0] = x;
            for (uint32_t i = 1u; !kCounterRng && i < static_cast<uint32_t>(kMtN); ++i) {
                x = mt_lcg(x, i);
                sm->state[0][i] = x;
            }
        }
        __syncthreads();
    }

    // Generates the next 624 words into the ring.  Uniform call; ends with a barrier.
    __device__ void produce_block() {
        if (kCounterRng) {  // word k of the stream needs nothing but the seed and k
            const uint32_t key = counter_key(sm->state[0][0]);
            for (int i = threadIdx.x; i < kMtN; i += blockDim.x)
                sm->ring[static_cast<int>((produced + i) & kRingMask)] = mt_canonical(lowbias32(key + static_cast<uint32_t>(produced + i)));
            __syncthreads();
            produced += kMtN;
            return;
        }
        const uint32_t* a = sm->state[which];
        uint32_t* b = sm->state[which ^ 1];
        const int base = static_cast<int>(produced & kRingMask);
        // phase 1: words [0, 227) depend on the old block only
        for (int i = threadIdx.x; i < kMtN - kMtM; i += blockDim.x) {
            const uint32_t v = mt_mix(a[i], a[i + 1], a[i + kMtM]);
            b[i] = v;
            sm->ring[(base + i) & kRingMask] = mt_canonical(mt_temper(v));
        }
        __syncthreads();
        // phase 2: words [227, 454) use new words [0, 227)
        for (int i = (kMtN - kMtM) + threadIdx.x; i < 2 * (kMtN - kMtM); i += blockDim.x) {
            const uint32_t v = mt_mix(a[i], a[i + 1], b[i - (kMtN - kMtM)]);
            b[i] = v;
            sm->ring[(base + i) & kRingMask] = mt_canonical(mt_temper(v));
        }
        __syncthreads();
        // phase 3: words [454, 624) use new words [227, 397); the last one wraps to new word 0
        for (int i = 2 * (kMtN - kMtM) + threadIdx.x; i < kMtN; i += blockDim.x) {
            const uint32_t nextWord = (i + 1 == kMtN) ? b[0] : a[i + 1];
            const uint32_t v = mt_mix(a[i], nextWord, b[i - (kMtN - kMtM)]);
            b[i] = v;
            sm->ring[(base + i) & kRingMask] = mt_canonical(mt_temper(v));
        }
        __syncthreads();
        which ^= 1;
        produced += kMtN;
    }

    // Makes every stream word below `end` available (words older than
    // kRingSize - 624 behind `end` may already be overwritten).
    __device__ void ensure(long long end) {
        while (produced < end) produce_block();
    }
    __device__ __forceinline__ float at(long long index) const {
        return sm->ring[static_cast<int>(index & kRingMask)];
    }
};

}  // namespace mcskin
