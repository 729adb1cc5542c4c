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
};

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

// robot vertices of the current pair in registers (loaded once per work item)
struct PolyRobotRegs { float x[kPolyMax], y[kPolyMax]; };

__device__ __forceinline__ void poly_load_robot(const PolyPairShared& S, PolyRobotRegs& R)
{
#pragma unroll
    for (int k = 0; k < kPolyMax; k++) { R.x[k] = (k < S.nr) ? S.rob_x[k] : 0.f; R.y[k] = (k < S.nr) ? S.rob_y[k] : 0.f; }
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
