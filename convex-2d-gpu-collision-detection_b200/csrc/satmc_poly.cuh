// satmc_poly.cuh -- Monte Carlo SAT for general convex polygons.
//
// The reference handles rectangles only and notes that SAT "can easily be extended" (README.md:3); its
// convex_collide projects on edge DIRECTIONS (utils.cu:170-171), which is valid only because a rectangle's edges
// are mutually perpendicular.  Here the axes are the true edge normals n = (e.y, -e.x), e = V[i+1] - V[i], of both
// polygons; everything else follows the reference's conventions:
//   robot vertex    x' = FADD(FFMA(x, c, -FMUL(y, s)), px)          (rot_trans_rectangle as compiled, utils.cu:132-142)
//   obstacle vertex x' = FFMA(z0, sd_x, FFMA(x, c, -FMUL(y, s)))    (sample_rectangle as compiled, utils.cu:144-157)
//   projection      p(q) = FFMA(n.x, q.x, FMUL(n.y, q.y))           (utils.cu:173-174)
//   min/max = fminf/fmaxf over the vertices; separated iff max1 < min2 || max2 < min1 (strict); ties collide
// Every rounding is pinned with intrinsics, so a CPU restatement of the same sequence (the test suite has one)
// gives bit-identical decisions.  Vertices are counter-clockwise, 1..8 per polygon.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace satmc {

constexpr int kPolyMax = 8;

// per-warp, in shared memory: everything that does not depend on the sample
struct PolyPairShared {
    float rob_x[kPolyMax], rob_y[kPolyMax];      // robot vertices, world (= nominal obstacle) frame
    float rnx[kPolyMax], rny[kPolyMax];          // robot edge normals
    float rmin[kPolyMax], rmax[kPolyMax];        // robot projected on its own normals
    float obs_x[kPolyMax], obs_y[kPolyMax];      // obstacle vertices, nominal
    float sd_x, sd_y, sd_t;
    int nr, no;
    // screening pass (see poly_screen): robot-normal thresholds against the obstacle's bounding circle, and the
    // inscribed-circle test
    float scr_kx[kPolyMax], scr_ky[kPolyMax];    // d_i = scr_kx[i] z0 + scr_ky[i] z1 = n_i . (sampled position of the obstacle's origin)
    float scr_mid[kPolyMax], scr_half[kPolyMax]; // |d_i - scr_mid[i]| > scr_half[i]  =>  separated on robot normal i
    float crx, cry, t2;                          // |centre - (crx, cry)|^2 < t2  =>  the inscribed circles overlap deeply
    // second level (poly_fast): the separating-axis test itself in fast arithmetic, see poly_fast_prologue
    float unx[kPolyMax], uny[kPolyMax], usup[kPolyMax];     // robot: unit outward normals (world frame), own support on them
    float vnx[kPolyMax], vny[kPolyMax], vsup[kPolyMax];     // obstacle: unit outward normals (its local frame), own support
    float eps2;                                  // |G| > eps2 => the sign of G is the exact decision; +inf: never decides
};

// Inscribed radius about (cx, cy) of the vertex chain (x, y)[0..n): the smallest signed distance to an edge line, counter-
// clockwise positive, clamped at 0.  A disc of that radius lies inside every edge half-plane, so the chain winds around
// it and it lies in the convex hull of the vertices -- for any input, convex or not; 0 for clockwise or degenerate input.
__device__ __forceinline__ float poly_inradius(const float* x, const float* y, int n, float cx, float cy)
{
    if (n < 3) return 0.f;
    float r = CUDART_INF_F;
    for (int i = 0; i < n; i++) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const float ex = x[j] - x[i], ey = y[j] - y[i];
        const float el = sqrtf(fmaf(ex, ex, ey * ey));
        const float d = ((x[i] - cx) * ey - (y[i] - cy) * ex) / el;
        if (!(d > 0.f)) return 0.f;                                   // also NaN (zero-length edge)
        r = fminf(r, d);
    }
    return r * 0.99999f;
}

// Screening constants (lane 0, after the exact constants).  With c = (z0 sd_x, z1 sd_y) the sampled position of the
// obstacle's local origin, every sampled obstacle vertex lies within rho of c whatever the rotation, so on robot
// normal n_i the obstacle projects inside [n_i.c - rho |n_i|, n_i.c + rho |n_i|]: if that interval clears the robot's
// own extent [rmin_i, rmax_i] by more than the rounding of both evaluations, the exact pass finds axis i separating.
// Conversely, if the circle of radius rin_o about c overlaps the robot's inscribed circle by more than that rounding,
// the projections overlap on every axis and the exact pass reports a collision.  Rounding: every quantity involved is
// a sum of at most four products of magnitudes <= M carrying at most six roundings of 2^-24 each (precise sinf/cosf:
// 2 ulp), i.e. below 2^-20 M; the margins use eta = 2^-16 on the sum of the magnitudes, with |z0|, |z1| <= 8 (samples
// beyond that are not screened).  Anything non-finite makes a comparison false, which means "not decided".
__device__ __forceinline__ void poly_screen_prologue(PolyPairShared& S, bool enable)
{
    const float eta = 1.52587890625e-05f, up = 1.000002f;
    const int nr = S.nr, no = S.no;
    float rho = 0.f;
    for (int k = 0; k < no; k++) rho = fmaxf(rho, sqrtf(fmaf(S.obs_x[k], S.obs_x[k], S.obs_y[k] * S.obs_y[k])));
    rho *= up;
    const float X = 8.0f * fabsf(S.sd_x), Y = 8.0f * fabsf(S.sd_y);
    for (int i = 0; i < nr; i++) {
        const float nx = S.rnx[i], ny = S.rny[i];
        const float ext = rho * (sqrtf(fmaf(nx, nx, ny * ny)) * up);
        const float m = eta * (fabsf(nx) * X + fabsf(ny) * Y + ext + fmaxf(fabsf(S.rmin[i]), fabsf(S.rmax[i])));
        float hi = S.rmax[i] + ext + m, lo = S.rmin[i] - ext - m;
        hi += fabsf(hi) * 2e-6f; lo -= fabsf(lo) * 2e-6f;
        // the same interval as centre and half width, widened by the rounding of that conversion and of d_i being formed
        // from the pre-multiplied coefficients n_i sd (two more roundings, far inside the margin m)
        const float mid = 0.5f * (hi + lo);
        const float half = 0.5f * (hi - lo) * 1.000001f + 4e-7f * (fabsf(hi) + fabsf(lo));
        S.scr_kx[i] = nx * S.sd_x; S.scr_ky[i] = ny * S.sd_y;
        S.scr_mid[i] = mid;
        S.scr_half[i] = (enable && half == half) ? half : CUDART_INF_F;
    }
    float cx = 0.f, cy = 0.f;
    for (int k = 0; k < nr; k++) { cx += S.rob_x[k]; cy += S.rob_y[k]; }
    cx /= (float)nr; cy /= (float)nr;
    float rho_r = 0.f;
    for (int k = 0; k < nr; k++) rho_r = fmaxf(rho_r, sqrtf(fmaf(S.rob_x[k] - cx, S.rob_x[k] - cx, (S.rob_y[k] - cy) * (S.rob_y[k] - cy))));
    // both centres must lie strictly inside their polygon (inradius 0 = centre outside, on the boundary, or degenerate)
    const float rin_r = poly_inradius(S.rob_x, S.rob_y, nr, cx, cy), rin_o = poly_inradius(S.obs_x, S.obs_y, no, 0.f, 0.f);
    const float t = rin_r + rin_o - eta * (fabsf(cx) + fabsf(cy) + X + Y + rho + rho_r);
    S.crx = cx; S.cry = cy;
    S.t2 = (enable && rin_r > 0.f && rin_o > 0.f && t > 0.f) ? t * t * 0.99999f : -1.0f;
}

// Second screening level: the separating-axis test itself, rotation included, in fast arithmetic.
//
// For convex counter-clockwise polygons A (robot) and B (sampled obstacle) let, for every edge of A with unit outward
// normal u, gap_u = min over B's vertices of u.b - max over A's vertices of u.a, and likewise for B's edges against A's
// vertices (evaluated in B's own frame, where B's normals and supports are per-pair constants: the robot's vertices are
// carried into that frame instead).  G = the largest of these nr + no one-sided gaps.  In exact real arithmetic G > 0 iff
// the polygons are disjoint (some edge line separates them), and for G < 0 its magnitude is the penetration depth: the
// smallest overlap of the two projections over ALL directions, so every one of the two-sided interval tests of the exact
// pass (poly_collide) overlaps by at least -G.  Hence with eps2 >= |G computed here - G of the ideal polygons| + the
// rounding of the exact pass's own comparisons in the same (length) units,
//     G >  eps2  =>  poly_collide finds a separating axis        G < -eps2  =>  poly_collide finds none,
// and only |G| <= eps2 (and NaN) goes on to the exact pass.  eps2 has the form of the rectangle bound,
// with the box sizes replaced by the vertex radii rho (1-norm) and the shortest edge L of each polygon:
//     E_exact = 5.5 u M + 1.5 (dS + dR) + 3 M (dR / L_r + dS / L_o),  dS = u (7 rho_o + dmax),  dR = u (7 rho_r + |P|_1)
//     E_fast  = 2 M e_m + 32 u M,   e_m = 2^-20 + 4 u |dt|max   (|__sinf/__cosf - sin/cos|, profiles/r1_trig_err.log)
//     M = |P|_1 + |d|max_1 + rho_r + rho_o,   eps2 = 1.0625 (E_exact + E_fast)
// Disabled (eps2 = +inf) unless both polygons are strictly convex and counter-clockwise with edges that are not tiny
// against their size, 8 sd_theta <= 64 and everything is finite: then every sample takes the exact pass as before.
__device__ __forceinline__ void poly_fast_prologue(PolyPairShared& S, float px, float py, const float* __restrict__ rob_local, bool enable)
{
    const float u = 5.9604645e-8f, Z = 8.0f;
    const int nr = S.nr, no = S.no;
    bool ok = enable && nr >= 3 && no >= 3;
    float rho_r = 0.f, rho_o = 0.f, Lr = CUDART_INF_F, Lo = CUDART_INF_F;
    for (int k = 0; k < nr; k++) rho_r = fmaxf(rho_r, fabsf(rob_local[2 * k]) + fabsf(rob_local[2 * k + 1]));
    for (int k = 0; k < no; k++) rho_o = fmaxf(rho_o, fabsf(S.obs_x[k]) + fabsf(S.obs_y[k]));
    for (int i = 0; i < nr; i++) {                                     // robot: world-frame vertices (already rotated and translated)
        const int j = (i + 1 == nr) ? 0 : i + 1, l = (j + 1 == nr) ? 0 : j + 1;
        const float ex = S.rob_x[j] - S.rob_x[i], ey = S.rob_y[j] - S.rob_y[i];
        const float fx = S.rob_x[l] - S.rob_x[j], fy = S.rob_y[l] - S.rob_y[j];
        const float len = sqrtf(fmaf(ex, ex, ey * ey)), len2 = sqrtf(fmaf(fx, fx, fy * fy));
        const float nx = ey / len, ny = -ex / len;
        float sup = -CUDART_INF_F;
        for (int k = 0; k < nr; k++) sup = fmaxf(sup, fmaf(nx, S.rob_x[k], ny * S.rob_y[k]));
        S.unx[i] = nx; S.uny[i] = ny; S.usup[i] = sup;
        Lr = fminf(Lr, len);
        ok = ok && (ex * fy - ey * fx > 1e-3f * len * len2);            // strictly convex turn, counter-clockwise
    }
    for (int i = 0; i < no; i++) {                                     // obstacle: nominal (local) vertices
        const int j = (i + 1 == no) ? 0 : i + 1, l = (j + 1 == no) ? 0 : j + 1;
        const float ex = S.obs_x[j] - S.obs_x[i], ey = S.obs_y[j] - S.obs_y[i];
        const float fx = S.obs_x[l] - S.obs_x[j], fy = S.obs_y[l] - S.obs_y[j];
        const float len = sqrtf(fmaf(ex, ex, ey * ey)), len2 = sqrtf(fmaf(fx, fx, fy * fy));
        const float nx = ey / len, ny = -ex / len;
        float sup = -CUDART_INF_F;
        for (int k = 0; k < no; k++) sup = fmaxf(sup, fmaf(nx, S.obs_x[k], ny * S.obs_y[k]));
        S.vnx[i] = nx; S.vny[i] = ny; S.vsup[i] = sup;
        Lo = fminf(Lo, len);
        ok = ok && (ex * fy - ey * fx > 1e-3f * len * len2);
    }
    const float dmx = Z * fabsf(S.sd_x), dmy = Z * fabsf(S.sd_y), dmt = Z * fabsf(S.sd_t);
    const float pn = fabsf(px) + fabsf(py);
    const float M = pn + dmx + dmy + rho_r + rho_o;
    const float dS = u * (7.0f * rho_o + fmaxf(dmx, dmy)), dR = u * (7.0f * rho_r + pn);
    const float e_m = 9.5367432e-7f + 4.0f * u * dmt;
    const float e_exact = 5.5f * u * M + 1.5f * (dS + dR) + 3.0f * M * (dR / Lr + dS / Lo);
    const float e_fast = 2.0f * M * e_m + 32.0f * u * M;
    const float eps = 1.0625f * (e_exact + e_fast);
    ok = ok && (dmt <= 64.0f) && (eps == eps) && (M < 1.0e18f) && (Lr > 1e-4f * rho_r) && (Lo > 1e-4f * rho_o) && (eps > 1.0e-30f);
    S.eps2 = ok ? eps : CUDART_INF_F;
}

// robot vertices of the current pair in registers (loaded once per work item)
struct PolyRobotRegs { float x[kPolyMax], y[kPolyMax]; };

__device__ __forceinline__ void poly_load_robot(const PolyPairShared& S, PolyRobotRegs& R)
{
#pragma unroll
    for (int k = 0; k < kPolyMax; k++) { R.x[k] = (k < S.nr) ? S.rob_x[k] : 0.f; R.y[k] = (k < S.nr) ? S.rob_y[k] : 0.f; }
}

// 0 = not decided (exact pass), 1 = separated, 2 = overlapping.  NR, NO as in poly_collide.
template <int NR, int NO>
__device__ __forceinline__ int poly_fast(const PolyPairShared& S, const PolyRobotRegs& R, float z0, float z1, float z2)
{
    const int nr = NR ? NR : S.nr, no = NO ? NO : S.no;
    const float dx = z0 * S.sd_x, dy = z1 * S.sd_y, dt = z2 * S.sd_t;
    const float c = __cosf(dt), s = __sinf(dt);
    float G = -CUDART_INF_F;
    // the sampled obstacle's vertices against the robot's edge lines
    float ox[kPolyMax], oy[kPolyMax];
#pragma unroll
    for (int k = 0; k < kPolyMax; k++) {
        if (k >= no) break;
        ox[k] = fmaf(c, S.obs_x[k], fmaf(-s, S.obs_y[k], dx));
        oy[k] = fmaf(s, S.obs_x[k], fmaf(c, S.obs_y[k], dy));
    }
#pragma unroll
    for (int i = 0; i < kPolyMax; i++) {
        if (i >= nr) break;
        const float nx = S.unx[i], ny = S.uny[i];
        float mn = fmaf(nx, ox[0], ny * oy[0]);
#pragma unroll
        for (int k = 1; k < kPolyMax; k++) {
            if (k >= no) break;
            mn = fminf(mn, fmaf(nx, ox[k], ny * oy[k]));
        }
        G = fmaxf(G, mn - S.usup[i]);
    }
    // the robot's vertices, carried into the obstacle's frame, against the obstacle's edge lines
    float wx[kPolyMax], wy[kPolyMax];
#pragma unroll
    for (int k = 0; k < kPolyMax; k++) {
        if (k >= nr) break;
        const float ux = R.x[k] - dx, uy = R.y[k] - dy;
        wx[k] = fmaf(c, ux, s * uy);
        wy[k] = fmaf(c, uy, -(s * ux));
    }
#pragma unroll
    for (int i = 0; i < kPolyMax; i++) {
        if (i >= no) break;
        const float nx = S.vnx[i], ny = S.vny[i];
        float mn = fmaf(nx, wx[0], ny * wy[0]);
#pragma unroll
        for (int k = 1; k < kPolyMax; k++) {
            if (k >= nr) break;
            mn = fminf(mn, fmaf(nx, wx[k], ny * wy[k]));
        }
        G = fmaxf(G, mn - S.vsup[i]);
    }
    // |z| beyond the bound the threshold was derived for, or anything non-finite: comparisons false => not decided
    const bool in_range = (fabsf(z0) <= 8.0f) && (fabsf(z1) <= 8.0f) && (fabsf(z2) <= 8.0f);
    if (!in_range) return 0;
    return (G > S.eps2) ? 1 : ((G < -S.eps2) ? 2 : 0);
}

// The screening constants of the current pair in registers (NR > 0) or left in shared memory (NR = 0: general loop).
template <int NR>
struct PolyScreenRegs {
    float kx[NR ? NR : 1], ky[NR ? NR : 1], mid[NR ? NR : 1], half[NR ? NR : 1];
    float sd_x, sd_y, sd_t, crx, cry, t2;
};

template <int NR>
__device__ __forceinline__ void poly_load_screen(const PolyPairShared& S, PolyScreenRegs<NR>& C)
{
#pragma unroll
    for (int i = 0; i < NR; i++) { C.kx[i] = S.scr_kx[i]; C.ky[i] = S.scr_ky[i]; C.mid[i] = S.scr_mid[i]; C.half[i] = S.scr_half[i]; }
    C.sd_x = S.sd_x; C.sd_y = S.sd_y; C.sd_t = S.sd_t; C.crx = S.crx; C.cry = S.cry; C.t2 = S.t2;
}

// Screening of one sample: 0 = not decided, 1 = separated, 2 = overlapping.  NR as in poly_collide.  RANGE: check the
// normals against the bound the margins were derived for (streamed samples; the fused sampler cannot exceed 6.77).
template <int NR, bool RANGE>
__device__ __forceinline__ int poly_screen(const PolyPairShared& S, const PolyScreenRegs<NR>& C, float z0, float z1, float z2)
{
    bool sep = false;
    if (NR) {
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const float d = fmaf(C.kx[i], z0, C.ky[i] * z1);
            sep = sep || (fabsf(d - C.mid[i]) > C.half[i]);
        }
    } else {
        const int nr = S.nr;
#pragma unroll
        for (int i = 0; i < kPolyMax; i++) {
            if (i >= nr) break;
            const float d = fmaf(S.scr_kx[i], z0, S.scr_ky[i] * z1);
            sep = sep || (fabsf(d - S.scr_mid[i]) > S.scr_half[i]);
        }
    }
    const float dx = fmaf(z0, C.sd_x, -C.crx), dy = fmaf(z1, C.sd_y, -C.cry);
    const bool col = fmaf(dx, dx, dy * dy) < C.t2;
    const int r = sep ? 1 : (col ? 2 : 0);
    if (!RANGE) return r;
    // the rotation angle itself does not matter (bounding / inscribed circles), but a non-finite one makes every exact
    // projection NaN, which the exact pass counts as a collision like the reference does: not screened
    const bool in_range = (fabsf(z0) <= 8.0f) && (fabsf(z1) <= 8.0f) && (fabsf(z2 * C.sd_t) <= 3.0e38f);
    return in_range ? r : 0;
}

// lane 0 of the warp fills the shared block from the descriptor (160 bytes: see satmc_poly_pair in satmc.h)
__device__ __forceinline__ void poly_prologue(PolyPairShared& S, const float* __restrict__ d)
{
    const float px = d[0], py = d[1], th = d[2];
    S.sd_x = d[3]; S.sd_y = d[4]; S.sd_t = d[5];
    int nr = (int)__float_as_uint(d[6]), no = (int)__float_as_uint(d[7]);
    nr = nr < 1 ? 1 : (nr > kPolyMax ? kPolyMax : nr);
    no = no < 1 ? 1 : (no > kPolyMax ? kPolyMax : no);
    S.nr = nr; S.no = no;
    const float c = cosf(th), s = sinf(th);
    for (int k = 0; k < nr; k++) {
        const float x = d[8 + 2 * k], y = d[8 + 2 * k + 1];
        S.rob_x[k] = __fadd_rn(__fmaf_rn(x, c, -__fmul_rn(y, s)), px);
        S.rob_y[k] = __fadd_rn(__fmaf_rn(x, s, __fmul_rn(y, c)), py);
    }
    for (int k = 0; k < no; k++) { S.obs_x[k] = d[24 + 2 * k]; S.obs_y[k] = d[24 + 2 * k + 1]; }
    for (int i = 0; i < nr; i++) {
        const int j = (i + 1 == nr) ? 0 : i + 1;
        const float ex = __fadd_rn(S.rob_x[j], -S.rob_x[i]), ey = __fadd_rn(S.rob_y[j], -S.rob_y[i]);
        const float nx = ey, ny = -ex;
        float mn = 0.f, mx = 0.f;
        for (int k = 0; k < nr; k++) {
            const float p = __fmaf_rn(nx, S.rob_x[k], __fmul_rn(ny, S.rob_y[k]));
            if (k == 0) { mn = mx = p; } else { mn = fminf(mn, p); mx = fmaxf(mx, p); }
        }
        S.rnx[i] = nx; S.rny[i] = ny; S.rmin[i] = mn; S.rmax[i] = mx;
    }
}

// decision for one sample: 1 = overlap.  NR, NO > 0: vertex counts known at compile time (straight-line code; the kernel
// uses <4, 4> for quadrilaterals).  NR = NO = 0: counts read from S; they are warp-uniform, so the `break`s below are
// uniform branches and a quad-quad pair does half the work of an octagon-octagon pair.
template <int NR, int NO>
__device__ __forceinline__ unsigned poly_collide(const PolyPairShared& S, const PolyRobotRegs& R, float z0, float z1, float z2)
{
    const int nr = NR ? NR : S.nr, no = NO ? NO : S.no;
    const float dt = __fmul_rn(z2, S.sd_t);
    const float c = cosf(dt), s = sinf(dt);
    float ox[kPolyMax], oy[kPolyMax];
#pragma unroll
    for (int k = 0; k < kPolyMax; k++) {
        if (k >= no) break;
        const float x = S.obs_x[k], y = S.obs_y[k];
        ox[k] = __fmaf_rn(z0, S.sd_x, __fmaf_rn(x, c, -__fmul_rn(y, s)));
        oy[k] = __fmaf_rn(z1, S.sd_y, __fmaf_rn(x, s, __fmul_rn(y, c)));
    }
    bool sep = false;
    // robot normals: the robot's own extent is a per-pair constant
#pragma unroll
    for (int i = 0; i < kPolyMax; i++) {
        if (i >= nr) break;
        const float nx = S.rnx[i], ny = S.rny[i];
        float mn = __fmaf_rn(nx, ox[0], __fmul_rn(ny, oy[0])), mx = mn;
#pragma unroll
        for (int k = 1; k < kPolyMax; k++) {
            if (k >= no) break;
            const float p = __fmaf_rn(nx, ox[k], __fmul_rn(ny, oy[k]));
            mn = fminf(mn, p); mx = fmaxf(mx, p);
        }
        sep = sep || (S.rmax[i] < mn) || (mx < S.rmin[i]);
    }
    // obstacle normals
#pragma unroll
    for (int i = 0; i < kPolyMax; i++) {
        if (i >= no) break;
        const float jx = (i + 1 < kPolyMax && i + 1 < no) ? ox[(i + 1) % kPolyMax] : ox[0];
        const float jy = (i + 1 < kPolyMax && i + 1 < no) ? oy[(i + 1) % kPolyMax] : oy[0];
        const float ex = __fadd_rn(jx, -ox[i]), ey = __fadd_rn(jy, -oy[i]);
        const float nx = ey, ny = -ex;
        float mn1 = __fmaf_rn(nx, R.x[0], __fmul_rn(ny, R.y[0])), mx1 = mn1;
#pragma unroll
        for (int k = 1; k < kPolyMax; k++) {
            if (k >= nr) break;
            const float p = __fmaf_rn(nx, R.x[k], __fmul_rn(ny, R.y[k]));
            mn1 = fminf(mn1, p); mx1 = fmaxf(mx1, p);
        }
        float mn2 = __fmaf_rn(nx, ox[0], __fmul_rn(ny, oy[0])), mx2 = mn2;
#pragma unroll
        for (int k = 1; k < kPolyMax; k++) {
            if (k >= no) break;
            const float p = __fmaf_rn(nx, ox[k], __fmul_rn(ny, oy[k]));
            mn2 = fminf(mn2, p); mx2 = fmaxf(mx2, p);
        }
        sep = sep || (mx1 < mn2) || (mx2 < mn1);
    }
    return sep ? 0u : 1u;
}

}  // namespace satmc
