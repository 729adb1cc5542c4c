// satmc_sampler.cuh -- counter-based Gaussian pose sampler (sm_100a).
//
// Replaces the reference's per-thread XORWOW state in global memory (curandState, setup_kernel
// utils.cu:111-117, curand_normal x5 per sample utils.cu:146-150: 18 LDG + 24 STG per sample) with a
// stateless generator: normals of sample s of stream (pair) p under seed k are a pure function of
// (k, p, s), so samples never touch HBM and any partition of the work gives identical results.
//
//   Samples are drawn in groups of four consecutive sample indices so that every Philox word and every
//   Box-Muller output is used (IMAD.WIDE runs at quarter rate on sm_100a: Philox is the scarce resource):
//     group g = s >> 2 holds samples 4g..4g+3, each needing D normals (D = 3: x,y,theta; D = 5: +w,h)
//     Philox4x32-10, key = (seed_lo, seed_hi), counter = (g_lo, g_hi, p, j), j = 0..D-1
//       -> 4 words -> two Box-Muller pairs -> normals n[4j..4j+3] = (cos0, sin0, cos1, sin1)
//     sample 4g+t uses n[D*t .. D*t+D-1] in the order x, y, theta[, w, h].
//   Box-Muller on words (a, b):
//     U     = RN(RN(a) * 2^-32 + 2^-33)         (float in [2^-33, 1]; exact for small a, so the tail has 32-bit
//                                                 resolution: largest radius sqrt(2*33*ln2) = 6.76, cuRAND's 6.66)
//     R     = sqrt(-2 ln U) = sqrt(-2 ln2 * log2 U)
//     phi   = 2 pi * ((b & 0x7fffff) + 0.5) * 2^-23
//     (R cos phi, R sin phi)
// log2/sqrt/sin/cos are the MUFU approximations: the normals are "native RNG" values, specified
// statistically (N(0,1) to ~1e-6), not bitwise; the CPU oracle restates the same formula with libm.
// The hit count for a given set of normals is bit-exact (satmc_fused_normals exposes them).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace satmc {

#define SATMC_PHILOX_M0 0xD2511F53u
#define SATMC_PHILOX_M1 0xCD9E8D57u
#define SATMC_PHILOX_W0 0x9E3779B9u
#define SATMC_PHILOX_W1 0xBB67AE85u

// The ten round keys (k0 + r*W0, k1 + r*W1) are precomputed on the host and live in the kernel
// parameter (constant) bank, so a round is 2 IMAD.WIDE + 2 LOP3 with a constant operand.
struct PhiloxKeys { uint32_t rk[20]; };

__host__ __device__ inline void philox_expand_key(uint32_t k0, uint32_t k1, PhiloxKeys& K)
{
    for (int r = 0; r < 10; r++) {
        K.rk[2 * r] = k0 + (uint32_t)r * SATMC_PHILOX_W0;
        K.rk[2 * r + 1] = k1 + (uint32_t)r * SATMC_PHILOX_W1;
    }
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& K, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)SATMC_PHILOX_M0 * c0;      // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)SATMC_PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.rk[2 * r];         // LOP3 with constant operand
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.rk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float mufu_lg2(float x)  { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x)  { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x)  { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// radius of the Box-Muller pair from one 32-bit word:  U = RN(RN(a) * 2^-32 + 2^-33) in (0,1], R = sqrt(-2 ln U)
__device__ __forceinline__ float bm_radius(uint32_t a)
{
    // I2FP.F32.U32 runs on the ALU pipe at half rate on sm_100a (tools/ubench.cu), cheaper than assembling the
    // float from 16-bit halves on the FMA pipe.  Small a convert exactly, so the tail keeps 32-bit resolution
    // (largest radius sqrt(2*33*ln2) = 6.76; cuRAND's is 6.66); near U = 1 (small radii) the spacing is the
    // float spacing 2^-24, where lg2.approx is accurate to 2^-22 absolute -- the same resolution as cuRAND's
    // float uniform + logf.
    const float U = __fmaf_rn(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float r2 = __fmul_rn(mufu_lg2(U), -1.3862943611198906f);           // -2 ln U >= 0
    return mufu_sqrt(fabsf(r2));
}

__device__ __forceinline__ float bm_angle(uint32_t b)
{
    const float f = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);        // [1,2)
    return __fmaf_rn(f, 6.283185307179586f, -6.283184932672558f);              // 2pi*(f - 1 + 2^-24)
}

// both normals of one Box-Muller pair
__device__ __forceinline__ void bm_pair(uint32_t a, uint32_t b, float& n_cos, float& n_sin)
{
    const float r = bm_radius(a), ang = bm_angle(b);
    n_cos = __fmul_rn(r, mufu_cos(ang));
    n_sin = __fmul_rn(r, mufu_sin(ang));
}

// The two Box-Muller pairs of one Philox block at once, FP32 arithmetic in packed form (sm_100a FFMA2 / FMUL2: the same
// IEEE operations lane by lane as bm_pair, so the normals are bit-identical to the scalar formulation; 6 packed
// instructions instead of 10 scalar ones, and packed FP32 overlaps with the Philox LOP3 / IMAD.WIDE stream around it).
#ifndef SATMC_BM_PACKED
#define SATMC_BM_PACKED 0     // measured: 3-DoF loop 3.23 -> 3.34 ms (16 extra register moves to form the pairs), 5-DoF +1 %
#endif
__device__ __forceinline__ void bm_two_pairs(const uint32_t w[4], float& c0, float& s0, float& c1, float& s1)
{
#if SATMC_BM_PACKED
    typedef unsigned long long f2;
    auto pk = [](float lo, float hi) { f2 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; };
    auto lo = [](f2 v) { return __uint_as_float((unsigned)v); };
    auto hi = [](f2 v) { return __uint_as_float((unsigned)(v >> 32)); };
    f2 U, r2, ang;
    const f2 a = pk(__uint2float_rn(w[0]), __uint2float_rn(w[2]));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(U) : "l"(a), "l"(pk(2.3283064365386963e-10f, 2.3283064365386963e-10f)),
        "l"(pk(1.1641532182693481e-10f, 1.1641532182693481e-10f)));
    const f2 lg = pk(mufu_lg2(lo(U)), mufu_lg2(hi(U)));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r2) : "l"(lg), "l"(pk(-1.3862943611198906f, -1.3862943611198906f)));
    const float ra = mufu_sqrt(fabsf(lo(r2))), rb = mufu_sqrt(fabsf(hi(r2)));
    const f2 f = pk(__uint_as_float((w[1] & 0x007fffffu) | 0x3f800000u), __uint_as_float((w[3] & 0x007fffffu) | 0x3f800000u));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ang) : "l"(f), "l"(pk(6.283185307179586f, 6.283185307179586f)),
        "l"(pk(-6.283184932672558f, -6.283184932672558f)));
    const f2 r = pk(ra, rb);
    f2 nc, ns;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(nc) : "l"(r), "l"(pk(mufu_cos(lo(ang)), mufu_cos(hi(ang)))));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ns) : "l"(r), "l"(pk(mufu_sin(lo(ang)), mufu_sin(hi(ang)))));
    c0 = lo(nc); c1 = hi(nc); s0 = lo(ns); s1 = hi(ns);
#else
    bm_pair(w[0], w[1], c0, s0);
    bm_pair(w[2], w[3], c1, s1);
#endif
}

// the 4*D normals of group (g_lo, g_hi) of stream p
template <int D>
__device__ __forceinline__ void group_normals(uint32_t g_lo, uint32_t g_hi, uint32_t p, const PhiloxKeys& K, float n[4 * D])
{
#pragma unroll
    for (int j = 0; j < D; j++) {
        uint32_t w[4];
        philox4x32_10(g_lo, g_hi, p, (uint32_t)j, K, w);
        bm_two_pairs(w, n[4 * j], n[4 * j + 1], n[4 * j + 2], n[4 * j + 3]);
    }
}

}  // namespace satmc
