// satmc_geom.cuh -- device geometry for the Monte Carlo SAT path (sm_100a).
//
// Two evaluators of the same decision "does the sampled obstacle overlap the robot?":
//
//   exact_*  : the reference arithmetic, operation for operation, as nvcc 12.9 compiles
//              sample_rectangle (utils.cu:144-157) and convex_collide (utils.cu:159-184) inside the
//              reference kernel (ztest.cu:151-155).  Every rounding is pinned with
//              __fmaf_rn/__fmul_rn/__fadd_rn so the compiler can neither add nor remove a
//              contraction.  This is the arithmetic contract of the library.
//
//   screen_* : a ~35-instruction oriented-box separating-axis test in centre/half-extent form
//              (4 axes, approximate sin/cos) that returns the largest normalised signed gap m.
//              |m| > eps  => the sign of m provably equals the exact decision (DESIGN.md section 4
//              derives eps); otherwise the sample is re-evaluated with exact_*.  Results are
//              therefore bit-identical to the exact path for every input, including NaN/Inf.
//
// Compile WITHOUT --use_fast_math: cosf/sinf below must be the precise libdevice versions the
// reference calls (utils.cu:133-134).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

// -DSATMC_DEBUG (libsatmc_debug.so): every indexed device-side write is bounds-checked; a violation prints
// the site and traps, which the host sees as a launch failure.  compute-sanitizer is not available on the
// GPU pool this library is developed on, so this build is run through the whole GPU test suite instead.
#ifdef SATMC_DEBUG
#include <cstdio>
#define SATMC_ASSERT(cond) do { if (!(cond)) { printf("SATMC_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                                                        (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define SATMC_ASSERT(cond) ((void)0)
#endif

namespace satmc {

// Bound on |z| under which the screening threshold is valid.  The fused sampler cannot exceed 6.77
// (Box-Muller radius of the smallest uniform); streamed samples are checked against it.
#define SATMC_Z_BOUND 8.0f

// ---------------------------------------------------------------------------------------------
// packed FP32 (sm_100a FFMA2 / FMUL2 / FADD2: two IEEE round-to-nearest operations per instruction, lane by lane
// identical to the scalar ones).  One packed instruction takes one issue slot for two FMA-pipe cycles, and -- unlike two
// scalar FFMAs -- overlaps with the half-rate ALU-pipe instructions around it (tools/ubench.cu: 8 FFMA2 + 16 LOP3 take
// 34 clk, 16 FFMA + 16 LOP3 take 45), which is what the screening loops are short of.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; }
__device__ __forceinline__ f32x2 dup2(float a) { return pack2(a, a); }
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// min(a, |b|, |c|) that returns NaN if any input is NaN (FMNMX3.NAN): a running minimum that cannot lose a NaN
__device__ __forceinline__ float min3_nan_abs(float a, float b, float c)
{
    float d;
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c)));
    return d;
}

// max(|a|, |b|, |c|) that returns NaN if any input is NaN (FMNMX3.NAN)
__device__ __forceinline__ float max3_nan_abs(float a, float b, float c)
{
    float d;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c)));
    return d;
}

// ---------------------------------------------------------------------------------------------
// exact reference arithmetic
// ---------------------------------------------------------------------------------------------

// create_rect(w,h) [utils.cu:119-130]: corners counter-clockwise from (-w/2, -h/2), AoS x0,y0..x3,y3
__device__ __forceinline__ void rect_base(float w, float h, float b[8])
{
    const float hx = w / 2, hy = h / 2;
    b[0] = -hx; b[1] = -hy; b[2] = hx; b[3] = -hy; b[4] = hx; b[5] = hy; b[6] = -hx; b[7] = hy;
}

// rot_trans_rectangle(pos, theta) [utils.cu:132-142] applied to the robot base quad (ztest.cu:148-149,297):
//   x' = FADD(FFMA(x,c,-FMUL(y,s)), px)   y' = FADD(FFMA(x,s,FMUL(y,c)), py)
// c, s = cosf(pose.theta), sinf(pose.theta): the precise libdevice values (PairConst::ca, sa)
__device__ __forceinline__ void exact_robot_corners(float px, float py, float c, float s, const float base[8], float r[8])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r[2 * i]     = __fadd_rn(__fmaf_rn(base[2 * i], c, -__fmul_rn(base[2 * i + 1], s)), px);
        r[2 * i + 1] = __fadd_rn(__fmaf_rn(base[2 * i], s, __fmul_rn(base[2 * i + 1], c)), py);
    }
}

// sample_rectangle with the normals supplied, as compiled inside the reference kernel:
//   dt = FMUL(z2,sd_t)  dw = FMUL(z3,sd_w)  dh = FMUL(z4,sd_h)   q = FFMA(dw|dh, -+0.5, base)
//   x' = FFMA(z0, sd_x, FFMA(qx, c, -FMUL(qy, s)))   y' = FFMA(z1, sd_y, FFMA(qx, s, FMUL(qy, c)))
__device__ __forceinline__ void exact_sample_corners(float ow, float oh, float sd_x, float sd_y, float sd_t,
                                                     float sd_w, float sd_h, float z0, float z1, float z2,
                                                     float z3, float z4, float o[8])
{
    const float hx = ow / 2, hy = oh / 2;
    const float dt = __fmul_rn(z2, sd_t);
    const float dw = __fmul_rn(z3, sd_w);
    const float dh = __fmul_rn(z4, sd_h);
    const float c = cosf(dt), s = sinf(dt);
    const float bx[4] = {-hx, hx, hx, -hx};
    const float by[4] = {-hy, -hy, hy, hy};
    const float hs_x[4] = {-0.5f, 0.5f, 0.5f, -0.5f};
    const float hs_y[4] = {-0.5f, -0.5f, 0.5f, 0.5f};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float qx = __fmaf_rn(dw, hs_x[i], bx[i]);
        const float qy = __fmaf_rn(dh, hs_y[i], by[i]);
        o[2 * i]     = __fmaf_rn(z0, sd_x, __fmaf_rn(qx, c, -__fmul_rn(qy, s)));
        o[2 * i + 1] = __fmaf_rn(z1, sd_y, __fmaf_rn(qx, s, __fmul_rn(qy, c)));
    }
}

// convex_collide [utils.cu:159-184]: 8 axes (edge vectors of r1 then r2),
//   n = FADD(r[i+1], -r[i])   p(q) = FFMA(n0, q.x, FMUL(n1, q.y))
// min/max with thrust's sequential minmax_element semantics (strict <, element 0 seeds both; NaNs
// never replace), strict separation test, no early exit.
__device__ __forceinline__ int exact_convex_collide(const float r1[8], const float r2[8])
{
    int collide = 1;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const float* r = j ? r2 : r1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float n0 = __fadd_rn(r[((i + 1) * 2) % 8], -r[i * 2]);
            const float n1 = __fadd_rn(r[((i + 1) * 2 + 1) % 8], -r[i * 2 + 1]);
            float min1, max1, min2, max2;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float p1 = __fmaf_rn(n0, r1[2 * k], __fmul_rn(n1, r1[2 * k + 1]));
                const float p2 = __fmaf_rn(n0, r2[2 * k], __fmul_rn(n1, r2[2 * k + 1]));
                if (k == 0) { min1 = max1 = p1; min2 = max2 = p2; }
                else {
                    if (p1 < min1) min1 = p1;
                    if (max1 < p1) max1 = p1;
                    if (p2 < min2) min2 = p2;
                    if (max2 < p2) max2 = p2;
                }
            }
            if (max1 < min2 || max2 < min1) collide = 0;
        }
    }
    return collide;
}

// ---------------------------------------------------------------------------------------------
// per-pair constants
// ---------------------------------------------------------------------------------------------
struct PairConst {
    // screening pass (robot: centre P, axes A0=(ca,sa), A1=(-sa,ca), half extents a0,a1;
    //                 obstacle: centre d = (sd_x z0, sd_y z1), angle dt = sd_t z2, half extents b0,b1)
    float ca, sa;                    // cosf/sinf(robot heading), precise (also used for the exact robot corners)
    float pa0, pa1;                  // P.A0, P.A1 (robot centre in the robot's own frame)
    float nkx0, nky0, kx1, nky1;     // u.A0 = pa0 + nkx0 z0 + nky0 z1 ; u.A1 = pa1 + kx1 z0 + nky1 z1,  u = P - d
    float th, nst;                   // phi = th + nst*z2 = (robot heading reduced to [-pi,pi]) - sd_t*z2
    float a0, a1, b0, b1;
    float hw, hh;                    // 0.5*sd_w, 0.5*sd_h  (5-DoF half-extent perturbation)
    float eps;                       // 3-DoF screening threshold eps_a + eps_b / min(b0,b1); +inf => always exact
    float eps_a, eps_b;              // 5-DoF: decided iff hmin > 0 and (|m| - eps_a) * hmin > eps_b, hmin = min(hx, hy)
    // exact pass
    float ow, oh, sd_x, sd_y, sd_t, sd_w, sd_h;
};

// Screening threshold.  u = 2^-24.  All quantities are upper bounds over every
// sample with |z_k| <= SATMC_Z_BOUND:
//   E_ref  : |exact-arithmetic normalised gap of the reference's rounded quads on its rounded edge
//             axis  -  gap of the ideal rectangles on the ideal axis|
//   E_fast : |screening value  -  gap of the ideal rectangles|
// The bound has the form eps = eps_a + eps_b / hmin, hmin = the smaller obstacle half extent (the
// obstacle's edge axes are known to relative accuracy ~ corner error / edge length).  With shape
// variance hmin changes per sample, so the two parts are kept separate.
__device__ __forceinline__ void screen_eps(float px, float py, float theta, float a0, float a1, float b0, float b1,
                                           float sd_x, float sd_y, float sd_t, float sd_w, float sd_h,
                                           float& eps_a, float& eps_b)
{
    const float u = 5.9604645e-8f;
    const float Z = SATMC_Z_BOUND;
    const float dmx = Z * fabsf(sd_x), dmy = Z * fabsf(sd_y), dmt = Z * fabsf(sd_t);
    const float dmw = 0.5f * Z * fabsf(sd_w), dmh = 0.5f * Z * fabsf(sd_h);
    const float hx_max = fabsf(b0) + dmw, hy_max = fabsf(b1) + dmh;
    const float r1a = a0 + a1, r1b = hx_max + hy_max;
    const float pn = fabsf(px) + fabsf(py), dn = dmx + dmy;
    const float un = pn + dn;
    const float M = un + r1a + r1b;
    const float dS = u * (7.0f * r1b + fmaxf(dmx, dmy));
    const float dR = u * (7.0f * r1a + pn);
    const float LA = 2.0f * fminf(a0, a1);
    // relative angle phi = theta_r - dt, |phi| <= pi + dmt.  |__sinf/__cosf - sin/cos|(x) <= 2^-20 + 4u|x| (measured on
    // B200, profiles/r1_trig_err.log); the float phi differs from the real theta - dt by <= 2u(|phi| + dmt + |theta|)
    // (rounding of the fma, of dt in the exact pass, and of the reduction of theta)
    const float phimax = 3.1415927f + dmt;
    const float e_m = 9.5367432e-7f + 4.0f * u * phimax + 2.0f * u * (phimax + dmt + fabsf(theta));
    const float e_fast = (un + 2.0f * (r1a + r1b)) * e_m + 24.0f * u * M;
    const float e_ref_a = 5.5f * u * M + 1.5f * (dS + dR) + 3.0f * M * __fdividef(dR, LA);   // 2-ulp division: inside the 6 % slack
    float ea = 1.0625f * (e_ref_a + e_fast);
    float eb = 1.0625f * 1.5f * M * dS;                               // 3 M dS / LB, LB = 2 hmin
    // outside the validated domain of the bound (degenerate robot, huge angles, non-finite or
    // astronomically large input): never trust the screening pass
    const bool ok = (LA > 0.0f) && (dmt <= 64.0f) && (fabsf(theta) <= 1.0e4f) && (ea == ea) && (eb == eb) && (M < 1.0e18f);
    eps_a = ok ? ea : CUDART_INF_F;
    eps_b = ok ? eb : CUDART_INF_F;
}

__device__ __forceinline__ void pair_const_init(PairConst& P, float rx, float ry, float rtheta, float rw,
                                                float rh, float ow, float oh, float sd_x, float sd_y,
                                                float sd_t, float sd_w, float sd_h)
{
    const float ca = cosf(rtheta), sa = sinf(rtheta);
    P.ca = ca; P.sa = sa;
    P.pa0 = fmaf(rx, ca, ry * sa);
    P.pa1 = fmaf(ry, ca, -(rx * sa));
    P.nkx0 = -(sd_x * ca); P.nky0 = -(sd_y * sa);
    P.kx1 = sd_x * sa;     P.nky1 = -(sd_y * ca);
    const float k = rintf(rtheta * 0.15915494f);                     // heading reduced to [-pi, pi] (2 pi = hi + lo)
    P.th = fmaf(-k, -1.7484555e-7f, fmaf(-k, 6.2831855f, rtheta));
    P.nst = -sd_t;
    P.a0 = 0.5f * fabsf(rw); P.a1 = 0.5f * fabsf(rh);
    P.b0 = 0.5f * fabsf(ow); P.b1 = 0.5f * fabsf(oh);
    P.hw = 0.5f * sd_w;      P.hh = 0.5f * sd_h;
    P.ow = ow; P.oh = oh; P.sd_x = sd_x; P.sd_y = sd_y; P.sd_t = sd_t; P.sd_w = sd_w; P.sd_h = sd_h;
    screen_eps(rx, ry, rtheta, P.a0, P.a1, P.b0, P.b1, sd_x, sd_y, sd_t, sd_w, sd_h, P.eps_a, P.eps_b);
    const float hmin = fminf(P.b0, P.b1);
    const float e3 = P.eps_a + __fdividef(P.eps_b, hmin);
    P.eps = (hmin > 0.0f && e3 == e3) ? e3 : CUDART_INF_F;
}

// ---------------------------------------------------------------------------------------------
// screening pass: largest normalised signed gap over the 4 box axes (m > 0 separated, m < 0 overlap)
// 22 FP32 instructions + 2 MUFU per sample (3-DoF)
// ---------------------------------------------------------------------------------------------
// The screening value in two steps, so that callers evaluating one sample under several settings with the same
// sd_theta can keep the trigonometric part (covariance sweep).  screen_gap is exactly screen_trig + screen_gap_sc.
__device__ __forceinline__ void screen_trig(float nst, float th, float z2, float& s, float& c)
{
    // relative angle phi = theta - dt of the robot frame against the sampled obstacle frame
    const float phi = fmaf(nst, z2, th);
    s = __sinf(phi); c = __cosf(phi);
}

template <int NDOF>
__device__ __forceinline__ float screen_gap_sc(const PairConst& P, float z0, float z1, float s, float c, float z3, float z4,
                                               float& hmin)
{
    // u in the robot frame, then rotated by the relative angle phi = theta - dt into the obstacle frame:
    // u.B0 = cos(phi) u.A0 - sin(phi) u.A1,  u.B1 = sin(phi) u.A0 + cos(phi) u.A1;  |A_i.B_j| = |cos phi|, |sin phi|
    const float ua0 = fmaf(P.nkx0, z0, fmaf(P.nky0, z1, P.pa0));
    const float ua1 = fmaf(P.kx1, z0, fmaf(P.nky1, z1, P.pa1));
    const float ub0 = fmaf(c, ua0, -(s * ua1));
    const float ub1 = fmaf(s, ua0, c * ua1);
    const float C = fabsf(c);
    const float S = fabsf(s);
    float hx = P.b0, hy = P.b1;
    // a perturbation past -width flips the corner order but spans the same rectangle: |half extent|
    if (NDOF == 5) { hx = fabsf(fmaf(z3, P.hw, hx)); hy = fabsf(fmaf(z4, P.hh, hy)); }
    hmin = fminf(hx, hy);
    const float tb0 = fabsf(ub0) - fmaf(P.a0, C, fmaf(P.a1, S, hx));
    const float tb1 = fabsf(ub1) - fmaf(P.a0, S, fmaf(P.a1, C, hy));
    const float ta0 = fabsf(ua0) - fmaf(hx, C, fmaf(hy, S, P.a0));
    const float ta1 = fabsf(ua1) - fmaf(hx, S, fmaf(hy, C, P.a1));
    return fmaxf(fmaxf(tb0, tb1), fmaxf(ta0, ta1));
}

template <int NDOF>
__device__ __forceinline__ float screen_gap(const PairConst& P, float z0, float z1, float z2, float z3, float z4,
                                            float& hmin)
{
    float s, c;
    screen_trig(P.nst, P.th, z2, s, c);
    return screen_gap_sc<NDOF>(P, z0, z1, s, c, z3, z4, hmin);
}

// Two samples at once in packed FP32: the same operations with the same roundings as two screen_gap<NDOF> calls
// (fma.rn.f32x2 is IEEE round-to-nearest per lane), a quarter fewer instructions -- and the packed ones overlap with the
// ALU-pipe instructions around them (fused cfg 3: 3.34 -> 3.23 ms).  The four gaps |u| - e stay scalar (operand |.| has
// no packed form).  z3 / z4 are read only for NDOF = 5.
template <int NDOF>
__device__ __forceinline__ void screen_gap_pair(const PairConst& P, float z0a, float z1a, float z2a, float z3a, float z4a,
                                                float z0b, float z1b, float z2b, float z3b, float z4b,
                                                float& m_a, float& m_b, float& hmin_a, float& hmin_b)
{
    const f32x2 phi = fma2(dup2(P.nst), pack2(z2a, z2b), dup2(P.th));
    const float sa = __sinf(lo2(phi)), ca = __cosf(lo2(phi)), sb = __sinf(hi2(phi)), cb = __cosf(hi2(phi));
    const f32x2 c2 = pack2(ca, cb), s2 = pack2(sa, sb), ns2 = pack2(-sa, -sb);
    const f32x2 z0 = pack2(z0a, z0b), z1 = pack2(z1a, z1b);
    const f32x2 ua0 = fma2(dup2(P.nkx0), z0, fma2(dup2(P.nky0), z1, dup2(P.pa0)));
    const f32x2 ua1 = fma2(dup2(P.kx1), z0, fma2(dup2(P.nky1), z1, dup2(P.pa1)));
    const f32x2 ub0 = fma2(c2, ua0, mul2(ns2, ua1));
    const f32x2 ub1 = fma2(s2, ua0, mul2(c2, ua1));
    const f32x2 C = pack2(fabsf(ca), fabsf(cb)), S = pack2(fabsf(sa), fabsf(sb));
    const f32x2 a0 = dup2(P.a0), a1 = dup2(P.a1);
    f32x2 hx = dup2(P.b0), hy = dup2(P.b1);
    hmin_a = hmin_b = fminf(P.b0, P.b1);
    if (NDOF == 5) {                              // a perturbation past -width flips the corner order but spans the same rectangle
        const f32x2 qx = fma2(pack2(z3a, z3b), dup2(P.hw), hx), qy = fma2(pack2(z4a, z4b), dup2(P.hh), hy);
        hx = pack2(fabsf(lo2(qx)), fabsf(hi2(qx))); hy = pack2(fabsf(lo2(qy)), fabsf(hi2(qy)));
        hmin_a = fminf(lo2(hx), lo2(hy)); hmin_b = fminf(hi2(hx), hi2(hy));
    }
    const f32x2 eb0 = fma2(a0, C, fma2(a1, S, hx));
    const f32x2 eb1 = fma2(a0, S, fma2(a1, C, hy));
    const f32x2 ea0 = fma2(hx, C, fma2(hy, S, a0));
    const f32x2 ea1 = fma2(hx, S, fma2(hy, C, a1));
    m_a = fmaxf(fmaxf(fabsf(lo2(ub0)) - lo2(eb0), fabsf(lo2(ub1)) - lo2(eb1)), fmaxf(fabsf(lo2(ua0)) - lo2(ea0), fabsf(lo2(ua1)) - lo2(ea1)));
    m_b = fmaxf(fmaxf(fabsf(hi2(ub0)) - hi2(eb0), fabsf(hi2(ub1)) - hi2(eb1)), fmaxf(fabsf(hi2(ua0)) - hi2(ea0), fabsf(hi2(ua1)) - hi2(ea1)));
}

__device__ __forceinline__ void screen_gap_pair3(const PairConst& P, float z0a, float z1a, float z2a, float z0b, float z1b, float z2b,
                                                 float& m_a, float& m_b)
{
    float ha, hb;
    screen_gap_pair<3>(P, z0a, z1a, z2a, 0.f, 0.f, z0b, z1b, z2b, 0.f, 0.f, m_a, m_b, ha, hb);
}

// true iff the sign of m is provably the exact decision (false for NaN anywhere)
template <int NDOF>
__device__ __forceinline__ bool screen_decided(const PairConst& P, float m, float hmin)
{
    if (NDOF == 5) return (hmin > 0.0f) && ((fabsf(m) - P.eps_a) * hmin > P.eps_b);
    return fabsf(m) > P.eps;
}

// Exact decision for one sample given the robot corners (kept in shared memory by the caller).
__device__ __noinline__ int exact_decide(const float* __restrict__ robot, float ow, float oh, float sd_x, float sd_y,
                                         float sd_t, float sd_w, float sd_h, float z0, float z1, float z2,
                                         float z3, float z4)
{
    float r[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = robot[k];
    exact_sample_corners(ow, oh, sd_x, sd_y, sd_t, sd_w, sd_h, z0, z1, z2, z3, z4, o);
    return exact_convex_collide(r, o);
}

}  // namespace satmc
