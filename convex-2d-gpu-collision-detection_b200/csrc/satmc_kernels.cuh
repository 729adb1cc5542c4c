// satmc_kernels.cuh -- the device side of libsatmc: counting kernels for the Monte Carlo SAT path (sm_100a).
//
// Work decomposition: a work item is (pair, sample chunk); one warp owns one
// item at a time (grid-stride over items), its 32 lanes stride over the samples of the chunk, each
// lane keeps a private hit counter, and the item ends with one REDUX warp reduction and one store
// or one atomic.  When every warp of a block works on the same pair the warp sums are combined in
// shared memory first and the block issues a single 64-bit atomic.
//
// The reference does the opposite (one thread = one pair, serial over samples, RNG state in global
// memory, ztest.cu:122-155), which starves the GPU whenever pairs < resident threads.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/satmc.h"
#include "satmc_geom.cuh"
#include "satmc_poly.cuh"
#include "satmc_sampler.cuh"

namespace satmc {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
// resident blocks per SM the register allocator aims for: the fused loop is FMA-pipe bound and wants
// registers (2 x 256 threads at 128 regs), the streamed loop is latency bound and wants warps
// (measured on B200: fused 254 vs 246 Gtests/s, streamed 3-DoF 3.8 vs 4.9 TB/s for 2 vs 4 blocks)
#ifndef SATMC_MIN_BLOCKS_FUSED
#define SATMC_MIN_BLOCKS_FUSED 2
#endif
#ifndef SATMC_MIN_BLOCKS_STREAMED
#define SATMC_MIN_BLOCKS_STREAMED 4
#endif
// bulk-tensor streamed kernel (measured: 3-DoF 5.58 / 5.73 / 5.26 TB/s, 5-DoF 6.62 / 5.88 / 5.74 TB/s at 2 / 3 / 4 blocks)
#ifndef SATMC_MIN_BLOCKS_TMA3
#define SATMC_MIN_BLOCKS_TMA3 2
#endif
#ifndef SATMC_MIN_BLOCKS_TMA5
#define SATMC_MIN_BLOCKS_TMA5 2
#endif

// ---------------------------------------------------------------------------------------------
// pair sources
// ---------------------------------------------------------------------------------------------
struct DirectSrc {
    const satmc_pair* pairs;
    __device__ __forceinline__ uint64_t element(uint64_t slot) const { return slot; }
    __device__ __forceinline__ uint32_t stream_id(uint64_t elem) const { return (uint32_t)elem; }
    // the robot is create_rect(rw, rh) by construction: the screening pass always applies
    __device__ __forceinline__ bool robot_is_centred_rect() const { return true; }
    __device__ __forceinline__ void robot_base8(const float v[12], float b[8]) const { rect_base(v[3], v[4], b); }
    __device__ __forceinline__ void load(uint64_t i, float v[12]) const {
        const float4* p = reinterpret_cast<const float4*>(pairs + i);   // 48 B, 16-B aligned
        const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y;
        v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
    }
};

// The reference's indirect layout (ztest.cu:135-140): per live pair a position and two float indices
// into the pose and std-dev tables; the robot is create_rect(robot_w, robot_h).
struct IndirectSrc {
    const float* robot_base; const float* poses; const float* std_devs;
    const float* pose_idxs; const float* std_dev_idxs; const float* positions;
    uint32_t n_poses, n_std;
    const int* live;               // optional: slot -> pair id (device-side work list of unfinished pairs)
    uint32_t stream_stride;        // Philox stream of element e = pair_id_offset + e * stream_stride (rows dealt round-robin to GPUs)
    __device__ __forceinline__ uint64_t element(uint64_t slot) const { return live ? (uint64_t)__ldg(live + slot) : slot; }
    __device__ __forceinline__ uint32_t stream_id(uint64_t elem) const { return (uint32_t)elem * stream_stride; }
    // The reference kernel transforms whatever 8 floats robot_base holds (ztest.cu:148-149).  Its mains always upload
    // create_rect(robot_w, robot_h) (ztest.cu:297), which is what the screening pass assumes (centre / half extents);
    // any other quad is honoured by evaluating every sample of the launch with the exact arithmetic on the 8 corners
    // as given (slow, same results as the reference).
    __device__ __forceinline__ bool robot_is_centred_rect() const {
        const float x0 = __ldg(robot_base), y0 = __ldg(robot_base + 1), x1 = __ldg(robot_base + 2), y1 = __ldg(robot_base + 3);
        const float x2 = __ldg(robot_base + 4), y2 = __ldg(robot_base + 5), x3 = __ldg(robot_base + 6), y3 = __ldg(robot_base + 7);
        return x0 == -x1 && y0 == -y2 && y1 == y0 && x2 == x1 && x3 == x0 && y3 == y2 && x1 >= 0.0f && y2 >= 0.0f;
    }
    __device__ __forceinline__ void robot_base8(const float*, float b[8]) const {
#pragma unroll
        for (int k = 0; k < 8; k++) b[k] = __ldg(robot_base + k);
    }
    __device__ __forceinline__ void load(uint64_t i, float v[12]) const {
        uint32_t pi = (uint32_t)(int)__ldg(pose_idxs + i);
        uint32_t si = (uint32_t)(int)__ldg(std_dev_idxs + i);
        pi = pi < n_poses ? pi : n_poses - 1;          // the reference would read out of bounds
        si = si < n_std ? si : n_std - 1;
        v[0] = __ldg(positions + 2 * i); v[1] = __ldg(positions + 2 * i + 1);
        v[2] = __ldg(poses + 3 * (size_t)pi + 2);
        v[3] = 2.0f * __ldg(robot_base + 2);           // create_rect: r[2] = w/2, r[5] = h/2 (exact)
        v[4] = 2.0f * __ldg(robot_base + 5);
        v[5] = __ldg(poses + 3 * (size_t)pi); v[6] = __ldg(poses + 3 * (size_t)pi + 1);
#pragma unroll
        for (int k = 0; k < 5; k++) v[7 + k] = __ldg(std_devs + 5 * (size_t)si + k);
    }
};

struct CountParams {
    uint64_t n_pairs;
    uint64_t n_samples;        // per pair
    uint64_t sample_offset;    // fused: first sample index
    uint64_t chunk;            // samples per work item (multiple of 128)
    uint64_t n_items;          // n_pairs * n_chunks
    uint32_t n_chunks;
    uint32_t pair_id_offset;
    uint32_t flags;            // SATMC_ACCUMULATE | SATMC_EXACT_ONLY
    uint32_t block_uniform;    // all warps of a block share a pair -> block reduction
    unsigned long long* hits;
    unsigned long long* exact_evals;
    // streamed
    const float* z; uint64_t ldz; uint64_t z_pair_stride; int ndof; int vec_ok;
    PhiloxKeys keys;           // fused: round keys, read straight from the constant bank
    // dynamic work distribution (null = static grid-stride): a warp's first item is its global warp index, every further
    // one is drawn from a device counter that only ever grows; ticket_base is its value when this launch starts
    unsigned long long* ticket; unsigned long long ticket_base;
    // Several work items per counter (n_chunks > 1): contributions are added atomically to `acc`, a scratch array of
    // the context that holds zeros between launches; the totals are moved to `hits` and `acc` is left zero again by the
    // kernel itself -- one launch, no memset.  With a known number of contributions per counter (`arrivals` != 0) the
    // last contribution does it (packed_arrive); otherwise the last block to finish does (finalize_counters_slow,
    // deferred-queue kernel only).  Counter i of the launch lives at offset (i / fin_inner) * fin_stride +
    // i % fin_inner of both arrays (sweep: one slice of the settings).
    unsigned long long* acc; unsigned* blocks_done;
    uint64_t n_counters, fin_inner, fin_stride;
    uint64_t hits_len;         // counters behind `hits` (and `acc`): bounds for the -DSATMC_DEBUG build
    uint32_t arrivals;         // packed scheme: contributions every counter receives (0: irregular, tail pass instead)
};

// where the atomics of a launch go
__device__ __forceinline__ unsigned long long* counter_base(const CountParams& p) { return p.acc ? p.acc : p.hits; }

// Packed variant for launches whose every counter receives a known number of contributions (everything but the
// deferred-queue kernel): arrival count in the top 24 bits of the accumulator, hits in the low 40.  The
// contribution that completes a counter sees the sum of all earlier ones in the value its own atomic returns, so it
// moves the total out and clears the accumulator on the spot -- one L2 round trip, no fence, no second counter, no
// barrier at the end of the kernel (a cfg 2 call is ~15 us in all, of which the fence + arrival + exchange chain of
// finalize_counters was ~3).  The host guarantees arrivals < 2^24 and hits per counter < 2^40.
constexpr int kPackShift = 40;
// internal flag (never accepted from callers): the totals are ADDED to p.hits with system-scope atomics, because p.hits
// is one array in the memory of a peer GPU that several devices finish into (satmc_group_count_fused_host, sample-range
// sharding inside one process: the compute kernel's own epilogue is the reduction over NVLink, no collective launch).
#define SATMC_PEER_ATOMIC_OUT 0x80000000u
__device__ __forceinline__ void packed_arrive(const CountParams& p, uint64_t counter, unsigned long long c)
{
    const unsigned long long add = (1ull << kPackShift) | c;
    const unsigned long long old = atomicAdd(p.acc + counter, add);
    if ((unsigned)(old >> kPackShift) + 1u == p.arrivals) {
        const unsigned long long total = (old + add) & ((1ull << kPackShift) - 1ull);
        SATMC_ASSERT(counter < p.hits_len);
        if (p.flags & SATMC_PEER_ATOMIC_OUT) atomicAdd_system(p.hits + counter, total);
        else if (p.flags & SATMC_ACCUMULATE) p.hits[counter] += total;
        else p.hits[counter] = total;
        p.acc[counter] = 0ull;
    }
}

// One contribution to a counter that several work items share (p.acc != nullptr).
__device__ __forceinline__ void contribute(const CountParams& p, uint64_t counter, unsigned long long c)
{
    if (p.arrivals != 0u) packed_arrive(p, counter, c);
    else atomicAdd(counter_base(p) + counter, c);
}

// End of every counting kernel.  All threads of the block must call it (it contains barriers).
// Out of line on purpose: inlined, its barriers and the extra live launch parameters changed ptxas's schedule of the
// fused hot loop (same instruction mix, 2 % slower; profiles/r2_codegen_experiments.log).
__device__ __noinline__ void finalize_counters_slow(const CountParams& p)
{
    __shared__ unsigned s_last;
    __syncthreads();                                                  // this block's atomics are issued
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(p.blocks_done, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();                                                  // every other block's atomics are visible
    // plain L2 loads and stores (nobody else touches the scratch any more), four counters in flight per thread: a loop
    // of dependent atomic exchanges took 0.8 us per counter and thread (56 us for 16 384 counters)
    constexpr int kInFlight = 4;
    for (uint64_t i0 = threadIdx.x; i0 < p.n_counters; i0 += (uint64_t)kInFlight * blockDim.x) {
        uint64_t off[kInFlight];
        unsigned long long v[kInFlight];
#pragma unroll
        for (int k = 0; k < kInFlight; k++) {
            const uint64_t i = i0 + (uint64_t)k * blockDim.x;
            off[k] = (i / p.fin_inner) * p.fin_stride + i % p.fin_inner;
            v[k] = (i < p.n_counters) ? __ldcg(p.acc + off[k]) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < kInFlight; k++) {
            if (i0 + (uint64_t)k * blockDim.x >= p.n_counters) break;
            SATMC_ASSERT(off[k] < p.hits_len);
            __stcg(p.acc + off[k], 0ull);
            if (p.flags & SATMC_PEER_ATOMIC_OUT) atomicAdd_system(p.hits + off[k], v[k]);
            else if (p.flags & SATMC_ACCUMULATE) p.hits[off[k]] += v[k];
            else p.hits[off[k]] = v[k];
        }
    }
    if (threadIdx.x == 0) *p.blocks_done = 0u;
}

__device__ __forceinline__ void finalize_counters(const CountParams& p)
{
    if (p.acc != nullptr && p.arrivals == 0u) finalize_counters_slow(p);   // launch-uniform; packed counters finish themselves
}

// Next work item of a warp.  `drawn` is the ticket lane 0 took while the warp was busy with the current item.
__device__ __forceinline__ uint64_t next_item(const CountParams& p, uint64_t item, uint64_t stride, unsigned long long drawn)
{
    if (p.ticket == nullptr) return item + stride;
    const unsigned long long t = __shfl_sync(0xffffffffu, drawn, 0);
    return (uint64_t)(t - p.ticket_base) + stride;
}
__device__ __forceinline__ unsigned long long draw_ticket(const CountParams& p, int lane)
{
    return (p.ticket != nullptr && lane == 0) ? atomicAdd(p.ticket, 1ull) : 0ull;
}

// Slow, general evaluation of the slots of one sample group selected by slot_mask (bit t = sample
// 4g+t): used for the ragged ends of a chunk and whenever the hot loop meets a sample the screening
// pass cannot decide.  The normals are regenerated from the counter, so the hot loop keeps nothing
// alive for it.
template <int D>
__device__ __noinline__ unsigned fused_group_slow(const PairConst& P, const float* robot, uint64_t g, unsigned slot_mask,
                                                  uint32_t pid, const PhiloxKeys& K, unsigned long long* exact_evals)
{
    float n[4 * D];
    group_normals<D>((uint32_t)g, (uint32_t)(g >> 32), pid, K, n);
    unsigned cnt = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        if (!(slot_mask & (1u << t))) continue;
        const float z3 = (D == 5) ? n[D * t + 3] : 0.0f, z4 = (D == 5) ? n[D * t + 4] : 0.0f;
        float hmin;
        const float m = screen_gap<D>(P, n[D * t], n[D * t + 1], n[D * t + 2], z3, z4, hmin);
        unsigned hit = __float_as_uint(m) >> 31;
        if (!screen_decided<D>(P, m, hmin)) {
            hit = (unsigned)exact_decide(robot, P.ow, P.oh, P.sd_x, P.sd_y, P.sd_t, P.sd_w, P.sd_h, n[D * t], n[D * t + 1],
                                         n[D * t + 2], z3, z4);
            if (exact_evals) atomicAdd(exact_evals, 1ull);
        }
        cnt += hit;
    }
    return cnt;
}

// Deferred cold work of one warp.  A group the screening pass cannot decide is evaluated by ONE lane while the other
// 31 wait (about 1 300 warp-instructions: regenerate the normals, screen again, precise sincos, 8-axis SAT).  Instead
// of doing that on the spot the hot loop only records (output slot, group) here; the warp works the list off 32
// entries at a time, all lanes busy, after an item or at the end of the kernel (cold_flush in k_count).
constexpr unsigned kColdCap = 256;
struct ColdQueue {
    unsigned n;                   // entries pushed (may run past kColdCap: those were evaluated on the spot)
    unsigned pair[kColdCap];      // output slot (index into hits / the live list)
    uint64_t g[kColdCap];         // sample group
};

// rare branch of the hot loop: queue the group (contributes 0 now, its exact count is added by cold_flush) or, when
// there is no queue (SATMC_EXACT_ONLY) or it is full, evaluate it here
template <int D>
__device__ __noinline__ unsigned fused_group_cold(ColdQueue* Q, unsigned pair_slot, const PairConst& Pcold, const float* robot,
                                                  uint64_t g, uint32_t pid, const PhiloxKeys& K, unsigned long long* exact_evals)
{
    if (Q != nullptr) {
        const unsigned slot = atomicAdd(&Q->n, 1u);
        if (slot < kColdCap) { SATMC_ASSERT(pair_slot < 0xffffffffu); Q->pair[slot] = pair_slot; Q->g[slot] = g; return 0u; }
    }
    return fused_group_slow<D>(Pcold, robot, g, 0xFu, pid, K, exact_evals);
}

// packed FP32 screening in the fused hot loop (3-DoF: cfg 3 3.34 -> 3.23 ms, single pair 293 -> 314 Gtests/s); the
// packed Box-Muller arithmetic (SATMC_BM_PACKED, satmc_sampler.cuh) costs more register moves than it saves
#ifndef SATMC_FUSED_PACKED
#define SATMC_FUSED_PACKED 1
#endif
#ifndef SATMC_FUSED_PACKED5
#define SATMC_FUSED_PACKED5 1
#endif
// hot path: all four samples of group g
template <int D, bool DEFER>
__device__ __forceinline__ unsigned fused_group(const PairConst& P, const PairConst& Pcold, const float* robot, uint64_t g,
                                                uint32_t pid, const PhiloxKeys& K, unsigned long long* exact_evals,
                                                ColdQueue* Q, unsigned pair_slot)
{
    float n[4 * D];
    group_normals<D>((uint32_t)g, (uint32_t)(g >> 32), pid, K, n);
    unsigned cnt = 0;
    bool decided = true;
    if constexpr (SATMC_FUSED_PACKED != 0 && D == 3) {              // samples in pairs, packed FP32 (screen_gap_pair)
        float m0, m1, m2, m3;
        screen_gap_pair3(P, n[0], n[1], n[2], n[3], n[4], n[5], m0, m1);
        screen_gap_pair3(P, n[6], n[7], n[8], n[9], n[10], n[11], m2, m3);
        cnt = (__float_as_uint(m0) >> 31) + (__float_as_uint(m1) >> 31) + (__float_as_uint(m2) >> 31) + (__float_as_uint(m3) >> 31);
        decided = min3_nan_abs(min3_nan_abs(CUDART_INF_F, m0, m1), m2, m3) > P.eps;    // NaN-propagating: false for NaN
    } else if constexpr (SATMC_FUSED_PACKED != 0 && SATMC_FUSED_PACKED5 != 0 && D == 5) {
        float m[4], h[4];
        screen_gap_pair<5>(P, n[0], n[1], n[2], n[3], n[4], n[5], n[6], n[7], n[8], n[9], m[0], m[1], h[0], h[1]);
        screen_gap_pair<5>(P, n[10], n[11], n[12], n[13], n[14], n[15], n[16], n[17], n[18], n[19], m[2], m[3], h[2], h[3]);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            cnt += __float_as_uint(m[t]) >> 31;
            decided = decided && screen_decided<5>(P, m[t], h[t]);
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const float z3 = (D == 5) ? n[D * t + 3] : 0.0f, z4 = (D == 5) ? n[D * t + 4] : 0.0f;
            float hmin;
            const float m = screen_gap<D>(P, n[D * t], n[D * t + 1], n[D * t + 2], z3, z4, hmin);
            cnt += __float_as_uint(m) >> 31;                        // m < 0 (m = -0 / NaN are undecided anyway)
            decided = decided && screen_decided<D>(P, m, hmin);
        }
    }
    if (!decided)                                                   // rare: redo the group, now or (DEFER) in cold_flush
        cnt = DEFER ? fused_group_cold<D>(Q, pair_slot, Pcold, robot, g, pid, K, exact_evals)
                    : fused_group_slow<D>(Pcold, robot, g, 0xFu, pid, K, exact_evals);
    return cnt;
}

// one sample of the streamed path (normals supplied)
template <int NDOF>
__device__ __noinline__ unsigned streamed_sample(const PairConst& P, const float* robot, float z0, float z1,
                                                    float z2, float z3, float z4, unsigned long long* exact_evals)
{
    float hmin;
    const float m = screen_gap<NDOF>(P, z0, z1, z2, z3, z4, hmin);
    unsigned hit = __float_as_uint(m) >> 31;
    // the screening bound assumes |z| <= SATMC_Z_BOUND; "<=" is false for NaN, so NaN/Inf go exact
    bool ok = screen_decided<NDOF>(P, m, hmin);
    ok = ok && (fabsf(z0) <= SATMC_Z_BOUND) && (fabsf(z1) <= SATMC_Z_BOUND) && (fabsf(z2) <= SATMC_Z_BOUND);
    if (NDOF == 5) ok = ok && (fabsf(z3) <= SATMC_Z_BOUND) && (fabsf(z4) <= SATMC_Z_BOUND);
    if (!ok) {
        hit = (unsigned)exact_decide(robot, P.ow, P.oh, P.sd_x, P.sd_y, P.sd_t, P.sd_w, P.sd_h, z0, z1, z2, z3, z4);
        if (exact_evals) atomicAdd(exact_evals, 1ull);
    }
    return hit;
}

// samples [b, e) (absolute indices) of one pair: full groups go through the hot loop, the at most two
// ragged groups at the ends through the slow path on lanes 0 and 1
// P lives in registers for the hot loop; Pcold is the same data in shared memory, handed to the out-of-line
// cold functions so that P's address is never taken (otherwise the compiler homes P on the local stack).
template <int D, bool DEFER>
__device__ __forceinline__ unsigned fused_chunk(const PairConst& P, const PairConst& Pcold, const float* robot, uint64_t b,
                                                uint64_t e, uint32_t pid, const PhiloxKeys& K, int lane,
                                                unsigned long long* exact_evals, ColdQueue* Q, unsigned pair_slot)
{
    unsigned cnt = 0;
    const uint64_t g_lo = (b + 3) >> 2, g_hi = e >> 2;
    if (g_hi < g_lo) {                                              // whole range inside one group
        if (lane == 0) {
            const unsigned mask = (0xFu << (unsigned)(b & 3)) & (0xFu >> (4 - (unsigned)(e & 3))) & 0xFu;
            cnt = fused_group_slow<D>(Pcold, robot, b >> 2, mask, pid, K, exact_evals);
        }
        return cnt;
    }
    for (uint64_t g = g_lo + (uint64_t)lane; g < g_hi; g += 32)
        cnt += fused_group<D, DEFER>(P, Pcold, robot, g, pid, K, exact_evals, Q, pair_slot);
    if (lane == 0 && (b & 3))
        cnt += fused_group_slow<D>(Pcold, robot, b >> 2, (0xFu << (unsigned)(b & 3)) & 0xFu, pid, K, exact_evals);
    if (lane == 1 && (e & 3))
        cnt += fused_group_slow<D>(Pcold, robot, g_hi, 0xFu >> (4 - (unsigned)(e & 3)), pid, K, exact_evals);
    return cnt;
}

// screening value of one streamed sample and whether it is decided (|z| guard included)
template <int NDOF>
__device__ __forceinline__ bool streamed_screen(const PairConst& P, float z0, float z1, float z2, float z3, float z4,
                                                unsigned& hit)
{
    float hmin;
    const float m = screen_gap<NDOF>(P, z0, z1, z2, z3, z4, hmin);
    hit = __float_as_uint(m) >> 31;
    bool ok = screen_decided<NDOF>(P, m, hmin);
    ok = ok & (fabsf(z0) <= SATMC_Z_BOUND) & (fabsf(z1) <= SATMC_Z_BOUND) & (fabsf(z2) <= SATMC_Z_BOUND);
    if (NDOF == 5) ok = ok & (fabsf(z3) <= SATMC_Z_BOUND) & (fabsf(z4) <= SATMC_Z_BOUND);
    return ok;
}

// Cold path of the vector loop: re-read the four samples starting at z[i] (they are still in L1/L2) and decide
// each with screening + exact fallback; the hot loop keeps nothing alive for it.
template <int NDOF>
__device__ __noinline__ unsigned streamed_quad_slow(const PairConst& P, const float* robot, const float* __restrict__ z,
                                                    uint64_t ldz, uint64_t i, unsigned long long* exact_evals)
{
    unsigned cnt = 0;
    for (int t = 0; t < 4; t++) {
        const float z0 = z[i + t], z1 = z[ldz + i + t], z2 = z[2 * ldz + i + t];
        const float z3 = (NDOF == 5) ? z[3 * ldz + i + t] : 0.0f, z4 = (NDOF == 5) ? z[4 * ldz + i + t] : 0.0f;
        cnt += streamed_sample<NDOF>(P, robot, z0, z1, z2, z3, z4, exact_evals);
    }
    return cnt;
}

template <int NDOF>
struct ZQuad { float4 a, b, c, d, e; };

template <int NDOF>
__device__ __forceinline__ void load_quad(ZQuad<NDOF>& q, const float* __restrict__ z, uint64_t ldz, uint64_t v)
{
    q.a = __ldg(reinterpret_cast<const float4*>(z) + v);
    q.b = __ldg(reinterpret_cast<const float4*>(z + ldz) + v);
    q.c = __ldg(reinterpret_cast<const float4*>(z + 2 * ldz) + v);
    if (NDOF == 5) {
        q.d = __ldg(reinterpret_cast<const float4*>(z + 3 * ldz) + v);
        q.e = __ldg(reinterpret_cast<const float4*>(z + 4 * ldz) + v);
    } else {
        q.d = make_float4(0.f, 0.f, 0.f, 0.f); q.e = q.d;
    }
}

template <int NDOF>
__device__ __forceinline__ unsigned streamed_chunk(const PairConst& P, const PairConst& Pcold, const float* robot,
                                                   const float* __restrict__ z, uint64_t ldz, uint64_t len, int vec_ok,
                                                   int lane, unsigned long long* exact_evals)
{
    unsigned cnt = 0;
    uint64_t done = 0;
    if (vec_ok) {
        // 4 samples per lane per trip from one LDG.128 per plane; the next trip's loads are issued before this
        // trip's arithmetic (register double buffering) so that each warp keeps 2 x ndof x 512 B in flight
        const uint64_t nvec = len / 4;
        uint64_t v = (uint64_t)lane;
        ZQuad<NDOF> cur, nxt;
        if (v < nvec) load_quad<NDOF>(cur, z, ldz, v);
        for (; v < nvec; v += 32) {
            const uint64_t vn = v + 32;
            if (vn < nvec) load_quad<NDOF>(nxt, z, ldz, vn);
            unsigned h0, h1, h2, h3;
            bool ok = streamed_screen<NDOF>(P, cur.a.x, cur.b.x, cur.c.x, cur.d.x, cur.e.x, h0);
            ok = ok & streamed_screen<NDOF>(P, cur.a.y, cur.b.y, cur.c.y, cur.d.y, cur.e.y, h1);
            ok = ok & streamed_screen<NDOF>(P, cur.a.z, cur.b.z, cur.c.z, cur.d.z, cur.e.z, h2);
            ok = ok & streamed_screen<NDOF>(P, cur.a.w, cur.b.w, cur.c.w, cur.d.w, cur.e.w, h3);
            unsigned c4 = h0 + h1 + h2 + h3;
            if (!ok) c4 = streamed_quad_slow<NDOF>(Pcold, robot, z, ldz, 4 * v, exact_evals);   // rare: redo the four
            cnt += c4;
            cur = nxt;
        }
        done = nvec * 4;
    }
    for (uint64_t i = done + (uint64_t)lane; i < len; i += 32) {
        const float z0 = __ldg(z + i), z1 = __ldg(z + ldz + i), z2 = __ldg(z + 2 * ldz + i);
        float z3 = 0.f, z4 = 0.f;
        if (NDOF == 5) { z3 = __ldg(z + 3 * ldz + i); z4 = __ldg(z + 4 * ldz + i); }
        cnt += streamed_sample<NDOF>(Pcold, robot, z0, z1, z2, z3, z4, exact_evals);
    }
    return cnt;
}

// Works off the warp's deferred groups, one entry per lane: rebuild the pair's constants, evaluate the group with the
// same code the immediate path uses, add the count to the pair's counter.  Entries may belong to earlier items of
// this warp; their counters were stored before (plain store or atomic, ordered by __syncwarp), so an atomic add is right.
template <class Src>
__device__ __noinline__ void cold_flush(ColdQueue* Qp, const Src& src, const CountParams& p, int lane, unsigned long long* counters)
{
    ColdQueue& Q = *Qp;
    __syncwarp();
    const unsigned n = Q.n < kColdCap ? Q.n : kColdCap;
    for (unsigned i = (unsigned)lane; i < n; i += 32) {
        const unsigned slot = Q.pair[i];
        const uint64_t g = Q.g[i];
        float v[12];
        const uint64_t elem = src.element(slot);
        src.load(elem, v);
        PairConst P;
        pair_const_init(P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
        float robot[8], base[8];
        src.robot_base8(v, base);
        exact_robot_corners(v[0], v[1], P.ca, P.sa, base, robot);
        if (!src.robot_is_centred_rect()) { P.eps = CUDART_INF_F; P.eps_b = CUDART_INF_F; }
        const uint32_t pid = p.pair_id_offset + src.stream_id(elem);
        const unsigned c = (v[10] == 0.0f && v[11] == 0.0f) ? fused_group_slow<3>(P, robot, g, 0xFu, pid, p.keys, p.exact_evals)
                                                           : fused_group_slow<5>(P, robot, g, 0xFu, pid, p.keys, p.exact_evals);
        SATMC_ASSERT(slot < p.hits_len);
        if (c) atomicAdd(counters + slot, (unsigned long long)c);
    }
    __syncwarp();
    if (lane == 0) Q.n = 0;
    __syncwarp();
}

// MULTI: several work items per counter (n_chunks > 1) -- atomics into the zero-invariant scratch + finalize_counters.
// A separate instantiation, so that the one-item-per-pair kernels (cfg 3, the adaptive loop) carry none of that code:
// its mere presence changes ptxas's schedule of the fused hot loop (same instruction mix, 1-2 % slower;
// profiles/r2_codegen_experiments.log).
template <class Src, bool STREAMED, bool DEFER = false, bool MULTI = false>
__global__ void __launch_bounds__(kThreads, STREAMED ? SATMC_MIN_BLOCKS_STREAMED : SATMC_MIN_BLOCKS_FUSED) k_count(const __grid_constant__ Src src, const __grid_constant__ CountParams p)
{
    unsigned long long* const counters = MULTI ? p.acc : p.hits;
    constexpr bool PACKED = MULTI && !DEFER;                          // see packed_arrive
    __shared__ float s_robot[kWarps][8];
    __shared__ PairConst s_pair[kWarps];
    __shared__ unsigned s_part[kWarps];
    __shared__ ColdQueue s_cold[DEFER ? kWarps : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * kWarps;
    ColdQueue* const Q = (!DEFER || (p.flags & SATMC_EXACT_ONLY)) ? nullptr : &s_cold[DEFER ? warp : 0];
    if (DEFER && Q != nullptr) { if (lane == 0) Q->n = 0; __syncwarp(); }
    // items are laid out pair-major; with block_uniform all 8 warps of a block walk the loop in step (static order)
    for (uint64_t item = (uint64_t)blockIdx.x * kWarps + warp; item < p.n_items;) {
        const unsigned long long drawn = draw_ticket(p, lane);       // in flight while this item is processed
        const uint64_t pair = item / p.n_chunks;
        const uint32_t chunk_id = (uint32_t)(item - pair * p.n_chunks);
        float v[12];
        const uint64_t elem = src.element(pair);                    // array index and Philox stream of this slot
        src.load(elem, v);
        PairConst P;
        pair_const_init(P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
        if ((p.flags & SATMC_EXACT_ONLY) || !src.robot_is_centred_rect()) { P.eps = CUDART_INF_F; P.eps_b = CUDART_INF_F; }
        __syncwarp();
        if (lane == 0) {
            float base[8];
            src.robot_base8(v, base);
            exact_robot_corners(v[0], v[1], P.ca, P.sa, base, s_robot[warp]);
            s_pair[warp] = P;
        }
        __syncwarp();
        const PairConst& Pc = s_pair[warp];
        const uint64_t c_begin = (uint64_t)chunk_id * p.chunk;
        const uint64_t c_len = (c_begin + p.chunk <= p.n_samples) ? p.chunk : (p.n_samples - c_begin);
        unsigned long long* ev = (p.flags & SATMC_EXACT_ONLY) ? nullptr : p.exact_evals;
        SATMC_ASSERT(pair < p.hits_len && c_begin < p.n_samples);
        unsigned cnt;
        if (STREAMED) {
            const float* z = p.z + pair * p.z_pair_stride + c_begin;
            cnt = (p.ndof == 5) ? streamed_chunk<5>(P, Pc, s_robot[warp], z, p.ldz, c_len, p.vec_ok, lane, ev)
                                : streamed_chunk<3>(P, Pc, s_robot[warp], z, p.ldz, c_len, p.vec_ok, lane, ev);
        } else {
            const uint32_t pid = p.pair_id_offset + src.stream_id(elem);
            const uint64_t s_begin = p.sample_offset + c_begin;
            const bool dof3 = (v[10] == 0.0f) && (v[11] == 0.0f);
            cnt = dof3 ? fused_chunk<3, DEFER>(P, Pc, s_robot[warp], s_begin, s_begin + c_len, pid, p.keys, lane, ev, Q, (unsigned)pair)
                       : fused_chunk<5, DEFER>(P, Pc, s_robot[warp], s_begin, s_begin + c_len, pid, p.keys, lane, ev, Q, (unsigned)pair);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        // (the non-MULTI instantiation keeps both branches although the host never sets block_uniform or n_chunks > 1
        // for it: without them ptxas schedules the fused hot loop differently and it runs 2 % slower)
        if (p.block_uniform) {
            if (lane == 0) s_part[warp] = cnt;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long t = 0;
#pragma unroll
                for (int w = 0; w < kWarps; w++) t += s_part[w];
                if (PACKED) packed_arrive(p, pair, t);
                else atomicAdd(counters + pair, t);                // one atomic per block
            }
            __syncthreads();
        } else if (lane == 0) {
            if (p.n_chunks == 1) {
                if (p.flags & SATMC_ACCUMULATE) p.hits[pair] += cnt; else p.hits[pair] = cnt;
            } else if (PACKED) {
                packed_arrive(p, pair, (unsigned long long)cnt);
            } else {
                atomicAdd(counters + pair, (unsigned long long)cnt);
            }
        }
        if (DEFER && Q != nullptr) {
            __syncwarp();
            if (Q->n >= 32u) cold_flush(Q, src, p, lane, counters);  // enough for a full pass
        }
        item = next_item(p, item, stride, drawn);
    }
    if (DEFER && Q != nullptr) cold_flush(Q, src, p, lane, counters);
    if (MULTI && !PACKED) finalize_counters_slow(p);
}

// ---------------------------------------------------------------------------------------------
// streamed path, bulk-copy staged (the default when the sample bank is 16-byte aligned)
//
// The sample bank is described to the TMA unit as a 2-D tensor [ndof planes][ldz samples].  Each warp runs a
// private 2-stage ring in shared memory: lane 0 issues ONE cp.async.bulk.tensor.2d (UTMALDG) for the next
// tile of 128 samples x ndof planes, completion is signalled on an mbarrier with expect_tx, and the 32
// lanes read their 4 samples per plane with one conflict-free LDS.128.  The bytes in flight live in shared
// memory instead of registers (32 warps x 2 stages x ndof x 512 B per SM) and the hot loop has no LDG and
// no address arithmetic.
// ---------------------------------------------------------------------------------------------
#ifndef SATMC_STREAMED_PACKED
#define SATMC_STREAMED_PACKED 1
#endif
// Ring geometry per sample layout and bank kind (measured on B200, profiles/r2_streamed_ring_experiments.log), all at
// 2 blocks per SM.  The 3-DoF loop is short (40 instructions per test), so the per-tile bookkeeping and the bytes in
// flight per warp decide: 4 stages of 256 samples (no spills at 128 registers).  The 5-DoF loop streams private banks
// from HBM best with 2 stages of 128 samples (6.9 vs 6.4 TB/s with 256), while on a shared, L2-resident bank it is
// issue bound and 256-sample tiles halve the per-tile bookkeeping (526 vs 465 Gtests/s).
#ifndef SATMC_TMA_TILE3
#define SATMC_TMA_TILE3 256
#endif
#ifndef SATMC_TMA_STAGES3
#define SATMC_TMA_STAGES3 4
#endif
#ifndef SATMC_TMA_TILE5
#define SATMC_TMA_TILE5 128
#endif
#ifndef SATMC_TMA_TILE5_SHARED
#define SATMC_TMA_TILE5_SHARED 256
#endif
#ifndef SATMC_TMA_STAGES5
#define SATMC_TMA_STAGES5 2
#endif
// samples per tile per plane (tile / 128 sub-tiles of 32 lanes x float4) and ring depth
__host__ __device__ constexpr int tma_tile(int ndof, bool shared_bank) { return ndof == 5 ? (shared_bank ? SATMC_TMA_TILE5_SHARED : SATMC_TMA_TILE5) : SATMC_TMA_TILE3; }
__host__ __device__ constexpr int tma_stages(int ndof) { return ndof == 5 ? SATMC_TMA_STAGES5 : SATMC_TMA_STAGES3; }
static_assert(tma_tile(3, false) % 128 == 0 && tma_tile(3, false) <= 256 && tma_tile(5, false) % 128 == 0 && tma_tile(5, false) <= 256 &&
              tma_tile(5, true) % 128 == 0 && tma_tile(5, true) <= 256, "a bulk-tensor box row is at most 256 elements");
static_assert((tma_stages(3) & (tma_stages(3) - 1)) == 0 && (tma_stages(5) & (tma_stages(5) - 1)) == 0, "ring depth: a power of two");

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_addr(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_addr(bar)) : "memory");
}

template <int NDOF, int kTile>
__global__ void __launch_bounds__(kThreads, NDOF == 5 ? SATMC_MIN_BLOCKS_TMA5 : SATMC_MIN_BLOCKS_TMA3)
k_count_streamed_tma(const __grid_constant__ DirectSrc src, const __grid_constant__ CountParams p,
                     const __grid_constant__ CUtensorMap zmap)
{
    constexpr int kStages = tma_stages(NDOF);
    extern __shared__ __align__(128) float s_tiles[];                 // [kWarps][kStages][NDOF][kTile]
    __shared__ __align__(8) uint64_t s_bar[kWarps][kStages];
    __shared__ float s_robot[kWarps][8];
    __shared__ PairConst s_pair[kWarps];
    __shared__ unsigned s_part[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* my_tiles = s_tiles + (size_t)warp * kStages * NDOF * kTile;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kStages; st++) mbar_init(&s_bar[warp][st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t tiles_done = 0;                                          // running tile counter: stage and parity
    const uint64_t stride = (uint64_t)gridDim.x * kWarps;
    // The ring runs across work items: while a warp consumes the last tiles of an item, lane 0 already requests the
    // first tiles of the warp's next item, so an item switch costs no memory round trip and items can be short (a
    // short item keeps the addresses the resident warps stream from close together).  Lane 0 only:
    int nx_x0 = 0;                                                    // tensor coordinate of the next item's chunk
    uint32_t nx_tiles = 0;                                            // its whole tiles
    bool nx_known = false;
    uint32_t pre = 0;                                                 // tiles of the CURRENT item requested while the previous one ran
    for (uint64_t item = (uint64_t)blockIdx.x * kWarps + warp; item < p.n_items;) {
        const unsigned long long drawn = draw_ticket(p, lane);
        const uint64_t pair = item / p.n_chunks;
        const uint32_t chunk_id = (uint32_t)(item - pair * p.n_chunks);
        const uint64_t c_begin = (uint64_t)chunk_id * p.chunk;
        const uint64_t c_len = (c_begin + p.chunk <= p.n_samples) ? p.chunk : (p.n_samples - c_begin);
        const uint32_t n_tiles = (uint32_t)(c_len / kTile);
        const int x0 = (int)(pair * p.z_pair_stride + c_begin);       // tensor coordinate of the chunk (< 2^31, host-checked)
        // request number u of this item's stream: its own tile u, or tile u - n_tiles of the warp's next item
        auto issue = [&](uint32_t u) {                                // lane 0 only
            int x;
            if (u < n_tiles) {
                x = x0 + (int)(u * kTile);
            } else {
                if (!nx_known) {
                    const uint64_t nx = (p.ticket != nullptr ? (uint64_t)(drawn - p.ticket_base) : item) + stride;
                    nx_tiles = 0;
                    if (nx < p.n_items) {
                        const uint64_t npair = nx / p.n_chunks;
                        const uint64_t nb = (nx - npair * p.n_chunks) * p.chunk;
                        const uint64_t nl = (nb + p.chunk <= p.n_samples) ? p.chunk : (p.n_samples - nb);
                        nx_tiles = (uint32_t)(nl / kTile);
                        nx_x0 = (int)(npair * p.z_pair_stride + nb);
                    }
                    nx_known = true;
                }
                const uint32_t w = u - n_tiles;
                if (w >= nx_tiles) return;                            // (no look-ahead beyond the next item)
                x = nx_x0 + (int)(w * kTile);
                pre = w + 1;
            }
            const uint32_t st = (tiles_done + u) % kStages;
            mbar_expect_tx(&s_bar[warp][st], NDOF * kTile * 4);
            tma_load_2d(my_tiles + (size_t)st * NDOF * kTile, &zmap, x, 0, &s_bar[warp][st]);
        };
        if (lane == 0) {
            const uint32_t have = pre;
            pre = 0; nx_known = false;
            for (uint32_t u = have; u < (uint32_t)kStages; u++) issue(u);
        }
        float v[12];
        src.load(pair, v);
        PairConst P;
        pair_const_init(P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
        if (p.flags & SATMC_EXACT_ONLY) { P.eps = CUDART_INF_F; P.eps_b = CUDART_INF_F; }
        __syncwarp();
        if (lane == 0) {
            float base[8];
            src.robot_base8(v, base);
            exact_robot_corners(v[0], v[1], P.ca, P.sa, base, s_robot[warp]);
            s_pair[warp] = P;
        }
        __syncwarp();
        const PairConst& Pc = s_pair[warp];
        unsigned long long* ev = (p.flags & SATMC_EXACT_ONLY) ? nullptr : p.exact_evals;
        SATMC_ASSERT(pair < p.hits_len && c_begin < p.n_samples);
        const float* z = p.z + pair * p.z_pair_stride + c_begin;
        unsigned cnt = 0;
        for (uint32_t t = 0; t < n_tiles; t++) {
            const uint32_t seq = tiles_done + t, st = seq % kStages;
            mbar_wait(&s_bar[warp][st], (seq / kStages) & 1u);
            const float4* tp = reinterpret_cast<const float4*>(my_tiles + (size_t)st * NDOF * kTile) + lane;
#pragma unroll
            for (int sub = 0; sub < kTile / 128; sub++, tp += 32) {
                const float4 a = tp[0], b = tp[kTile / 4], c = tp[2 * (kTile / 4)];
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f), e = d;
                if (NDOF == 5) { d = tp[3 * (kTile / 4)]; e = tp[4 * (kTile / 4)]; }
                if (sub == kTile / 128 - 1) {
                    __syncwarp();                                         // every lane has its samples: the stage is free
                    if (lane == 0) issue(t + kStages);                // __syncwarp ordered the reads before this write
                }
                unsigned c4;
                bool ok;
                if (NDOF == 3 && SATMC_STREAMED_PACKED) {                  // two samples per packed-FP32 evaluation
                    float m0, m1, m2, m3;
                    screen_gap_pair3(P, a.x, b.x, c.x, a.y, b.y, c.y, m0, m1);
                    screen_gap_pair3(P, a.z, b.z, c.z, a.w, b.w, c.w, m2, m3);
                    c4 = (__float_as_uint(m0) >> 31) + (__float_as_uint(m1) >> 31) + (__float_as_uint(m2) >> 31) + (__float_as_uint(m3) >> 31);
                    // decided: every |m| above the threshold (NaN-propagating minimum) and every normal inside the bound
                    const float mn = min3_nan_abs(min3_nan_abs(CUDART_INF_F, m0, m1), m2, m3);
                    // largest |normal| of the four samples, NaN-propagating: "<= bound" is false for NaN, +-Inf exceeds it
                    const float zmax = max3_nan_abs(max3_nan_abs(max3_nan_abs(a.x, a.y, a.z), max3_nan_abs(a.w, b.x, b.y), max3_nan_abs(b.z, b.w, c.x)),
                                                    c.y, max3_nan_abs(c.z, c.w, 0.0f));
                    ok = (mn > P.eps) & (zmax <= SATMC_Z_BOUND);
                } else {
                    unsigned h0, h1, h2, h3;
                    ok = streamed_screen<NDOF>(P, a.x, b.x, c.x, d.x, e.x, h0);
                    ok = ok & streamed_screen<NDOF>(P, a.y, b.y, c.y, d.y, e.y, h1);
                    ok = ok & streamed_screen<NDOF>(P, a.z, b.z, c.z, d.z, e.z, h2);
                    ok = ok & streamed_screen<NDOF>(P, a.w, b.w, c.w, d.w, e.w, h3);
                    c4 = h0 + h1 + h2 + h3;
                }
                if (!ok) c4 = streamed_quad_slow<NDOF>(Pc, s_robot[warp], z, p.ldz, (uint64_t)t * kTile + 128 * sub + 4 * lane, ev);
                cnt += c4;
            }
        }
        tiles_done += n_tiles;
        for (uint64_t i = (uint64_t)n_tiles * kTile + (uint64_t)lane; i < c_len; i += 32) {      // ragged tail
            const float z0 = __ldg(z + i), z1 = __ldg(z + p.ldz + i), z2 = __ldg(z + 2 * p.ldz + i);
            float z3 = 0.f, z4 = 0.f;
            if (NDOF == 5) { z3 = __ldg(z + 3 * p.ldz + i); z4 = __ldg(z + 4 * p.ldz + i); }
            cnt += streamed_sample<NDOF>(Pc, s_robot[warp], z0, z1, z2, z3, z4, ev);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (p.block_uniform) {
            if (lane == 0) s_part[warp] = cnt;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long tsum = 0;
#pragma unroll
                for (int w = 0; w < kWarps; w++) tsum += s_part[w];
                contribute(p, pair, tsum);
            }
            __syncthreads();
        } else if (lane == 0) {
            if (p.n_chunks == 1) {
                if (p.flags & SATMC_ACCUMULATE) p.hits[pair] += cnt; else p.hits[pair] = cnt;
            } else {
                contribute(p, pair, (unsigned long long)cnt);
            }
        }
        item = next_item(p, item, stride, drawn);
    }
    finalize_counters(p);
}

// ---------------------------------------------------------------------------------------------
// covariance sweep with common random numbers (BASELINE cfg 5, the ztest-style variance sweep)
//
// Every pair is evaluated under n_cov pose-covariance settings (sd_x, sd_y, sd_theta) on the SAME normals: setting c
// of pair p sees exactly the samples satmc_count_fused would give a pair with that sigma and stream id p.  The
// sampler (47 % of the issue slots of the plain fused loop) is then paid once per sample instead of once per
// (sample, setting): a lane draws two 4-sample groups (24 normals) and runs the screening test for all settings.
//
// The settings arrive as a launch parameter, sorted by sd_theta on the host (SweepSettings).  Settings with the same
// sd_theta give every sample the same relative angle, hence the same sine and cosine AND the same four projected
// extents (they depend on the angle and the box sizes only): the kernel recomputes those 2 MUFU + 9 FP32 per sample
// only when sd_theta changes (a 4x4x4 grid: 4 times per 64 settings) and is left with 13 FP32 + 1 shift per
// (sample, setting): two centre projections (4 FFMA), their rotation into the obstacle frame (1 FMUL + 2 FFMA + 1 FADD
// folded), four gaps (4 FADD), the maximum (FMNMX3 x 2), the hit bit shifted into a per-lane mask (SHF) and the
// "decided" predicate (FSETP).  Hits per setting are popcounted, warp-reduced and added to shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kSweepMax = 64;

struct SweepSettings {                // by value in the kernel-parameter bank; order = ascending sd_theta bit pattern
    float sx[kSweepMax], sy[kSweepMax], st[kSweepMax];
    unsigned char orig[kSweepMax];    // sorted position -> position in the caller's array (the output column)
    int n;
};

struct SweepSetting {             // 32 bytes: one pointer walks the settings, fields at immediate offsets
    float4 k;                     // nkx0, nky0, kx1, nky1 of the pair under sorted setting r
    float2 e;                     // nst, eps
    unsigned cnt, pad;
};
struct SweepShared {
    SweepSetting s[kSweepMax];
    float robot[8];
    PairConst base;               // setting-independent constants, for the out-of-line paths
};

// one sample of one setting, screening + exact fallback (edges and undecided samples)
__device__ __noinline__ unsigned sweep_sample_slow(const SweepShared& S, const SweepSettings& W, int r, float z0, float z1, float z2,
                                                   unsigned long long* exact_evals)
{
    PairConst Q = S.base;
    const float4 k = S.s[r].k; const float2 e = S.s[r].e;
    Q.nkx0 = k.x; Q.nky0 = k.y; Q.kx1 = k.z; Q.nky1 = k.w; Q.nst = e.x; Q.eps = e.y;
    float hmin;
    const float m = screen_gap<3>(Q, z0, z1, z2, 0.f, 0.f, hmin);
    unsigned hit = __float_as_uint(m) >> 31;
    if (!screen_decided<3>(Q, m, hmin)) {
        hit = (unsigned)exact_decide(S.robot, Q.ow, Q.oh, W.sx[r], W.sy[r], W.st[r], 0.f, 0.f, z0, z1, z2, 0.f, 0.f);
        if (exact_evals) atomicAdd(exact_evals, 1ull);
    }
    return hit;
}

// all samples of super-group q (groups G q .. G q + G - 1) for setting r, normals regenerated: rare path of the main loop
__device__ __noinline__ unsigned sweep_super_slow(const SweepShared& S, const SweepSettings& W, int r, uint64_t q, int G, uint32_t pid,
                                                  const PhiloxKeys& K, unsigned long long* exact_evals)
{
    unsigned cnt = 0;
    for (int h = 0; h < G; h++) {
        const uint64_t g = (uint64_t)G * q + h;
        float n[12];
        group_normals<3>((uint32_t)g, (uint32_t)(g >> 32), pid, K, n);
        for (int t = 0; t < 4; t++) cnt += sweep_sample_slow(S, W, r, n[3 * t], n[3 * t + 1], n[3 * t + 2], exact_evals);
    }
    return cnt;
}

// the angle-dependent part of screen_gap_sc: the four projected extents (same arithmetic, same roundings)
__device__ __forceinline__ void screen_extents(const PairConst& P, float s, float c, float& eb0, float& eb1, float& ea0, float& ea1)
{
    const float C = fabsf(c), S = fabsf(s);
    eb0 = fmaf(P.a0, C, fmaf(P.a1, S, P.b0));
    eb1 = fmaf(P.a0, S, fmaf(P.a1, C, P.b1));
    ea0 = fmaf(P.b0, C, fmaf(P.b1, S, P.a0));
    ea1 = fmaf(P.b0, S, fmaf(P.b1, C, P.a1));
}

// G = 4-sample groups per lane and trip (NS = 4 G samples); BPS = resident blocks per SM the register budget is cut for;
// WARPS per block.  Measured on B200, cfg 5 (profiles/r2_sweep_experiments.log): 8 samples per trip in 6-warp blocks at
// 168 registers (nothing spilled in the loop) 1027 Gtests/s; 4 samples in 8-warp blocks at 128 registers 958; 8 samples
// at 128 registers (spills) 955; 8 samples at 232 registers, one block per SM 909.
#ifndef SATMC_SWEEP_G
#define SATMC_SWEEP_G 2
#endif
#ifndef SATMC_SWEEP_BPS
#define SATMC_SWEEP_BPS 2
#endif
#ifndef SATMC_SWEEP_WARPS
#define SATMC_SWEEP_WARPS 6
#endif
constexpr int kSweepWarps = SATMC_SWEEP_WARPS;
template <int G>
__global__ void __launch_bounds__(32 * kSweepWarps, SATMC_SWEEP_BPS) k_count_sweep(const satmc_pair* __restrict__ pairs, const __grid_constant__ SweepSettings W,
                                                             uint64_t hits_stride, const __grid_constant__ CountParams p)
{
    constexpr int NS = 4 * G;
    constexpr int LG = (G == 1) ? 2 : 3;                              // log2(NS)
    static_assert(G == 1 || G == 2, "one or two groups per trip");
    __shared__ SweepShared s_sw[kSweepWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_cov = W.n;
    SweepShared& S = s_sw[warp];
    const uint64_t stride = (uint64_t)gridDim.x * kSweepWarps;
    for (uint64_t item = (uint64_t)blockIdx.x * kSweepWarps + warp; item < p.n_items; item += stride) {
        const uint64_t pair = item / p.n_chunks;
        const uint32_t chunk_id = (uint32_t)(item - pair * p.n_chunks);
        float v[12];
        DirectSrc src{pairs};
        src.load(pair, v);
        PairConst P;                                                   // setting-independent part (sigma fields overridden below)
        pair_const_init(P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], 0.f, 0.f, 0.f, 0.f, 0.f);
        SATMC_ASSERT(n_cov >= 0 && n_cov <= kSweepMax && pair < p.n_pairs);
        __syncwarp();
        for (int r = lane; r < n_cov; r += 32) {                       // per-setting constants, two settings per lane
            const float sx = W.sx[r], sy = W.sy[r], st = W.st[r];
            float ea, eb;
            screen_eps(v[0], v[1], v[2], P.a0, P.a1, P.b0, P.b1, sx, sy, st, 0.f, 0.f, ea, eb);
            const float hmin = fminf(P.b0, P.b1);
            const float e3 = ea + __fdividef(eb, hmin);
            const bool ok = hmin > 0.0f && e3 == e3 && !(p.flags & SATMC_EXACT_ONLY);
            S.s[r].k = make_float4(-(sx * P.ca), -(sy * P.sa), sx * P.sa, -(sy * P.ca));
            S.s[r].e = make_float2(-st, ok ? e3 : CUDART_INF_F);
            S.s[r].cnt = 0;
        }
        if (lane == 0) {
            float base[8];
            rect_base(v[3], v[4], base);
            exact_robot_corners(v[0], v[1], P.ca, P.sa, base, S.robot);
            S.base = P;
        }
        __syncwarp();
        const uint64_t c_begin = (uint64_t)chunk_id * p.chunk;
        const uint64_t c_len = (c_begin + p.chunk <= p.n_samples) ? p.chunk : (p.n_samples - c_begin);
        unsigned long long* ev = (p.flags & SATMC_EXACT_ONLY) ? nullptr : p.exact_evals;
        const uint32_t pid = p.pair_id_offset + (uint32_t)pair;
        const uint64_t b = p.sample_offset + c_begin, e = b + c_len;
        const uint64_t q_lo = (b + NS - 1) >> LG, q_hi = e >> LG;       // full NS-sample super-groups [q_lo, q_hi)
        for (uint64_t q0 = q_lo; q0 < q_hi; q0 += 32) {                 // all 32 lanes stay in step; idle lanes count nothing
            const uint64_t q = q0 + (uint64_t)lane;
            const bool mine = q < q_hi;
            float n[3 * NS];
#pragma unroll
            for (int h = 0; h < G; h++) {
                const uint64_t g = (uint64_t)G * q + h;
                group_normals<3>((uint32_t)g, (uint32_t)(g >> 32), pid, p.keys, n + 12 * h);
            }
            // samples in pairs (2j, 2j+1): the centre projections and their rotation run as packed FP32
            f32x2 z0p[NS / 2], z1p[NS / 2], cs2[NS / 2], sn2[NS / 2], ns2[NS / 2];
            float eb0[NS], eb1[NS], ea0[NS], ea1[NS];
#pragma unroll
            for (int j = 0; j < NS / 2; j++) { z0p[j] = pack2(n[6 * j], n[6 * j + 3]); z1p[j] = pack2(n[6 * j + 1], n[6 * j + 4]); }
            const f32x2 pa0_2 = dup2(P.pa0), pa1_2 = dup2(P.pa1);
            unsigned mine_mask = mine ? 0xffffffffu : 0u;
            asm volatile("" : "+r"(mine_mask));                          // keep it in a register (else it is re-derived per setting)
            unsigned long long und = 0ull;
            uint32_t prev = 0u;
            SweepSetting* sp = S.s;
            for (int r = 0; r < n_cov; r++, sp++) {
                const float4 k = sp->k; const float2 ee = sp->e;
                if (r == 0 || __float_as_uint(ee.x) != prev) {           // warp-uniform: sd_theta changes (sorted: once per value)
                    prev = __float_as_uint(ee.x);
#pragma unroll
                    for (int j = 0; j < NS / 2; j++) {
                        float s0, c0, s1, c1;
                        screen_trig(ee.x, P.th, n[6 * j + 2], s0, c0);
                        screen_trig(ee.x, P.th, n[6 * j + 5], s1, c1);
                        screen_extents(P, s0, c0, eb0[2 * j], eb1[2 * j], ea0[2 * j], ea1[2 * j]);
                        screen_extents(P, s1, c1, eb0[2 * j + 1], eb1[2 * j + 1], ea0[2 * j + 1], ea1[2 * j + 1]);
                        cs2[j] = pack2(c0, c1); sn2[j] = pack2(s0, s1); ns2[j] = pack2(-s0, -s1);
                    }
                }
                const f32x2 kx = dup2(k.x), ky = dup2(k.y), kz = dup2(k.z), kw = dup2(k.w);
                unsigned hits = 0u;
                float mn = CUDART_INF_F;                                 // min |m| over the samples (NaN-propagating)
#pragma unroll
                for (int j = 0; j < NS / 2; j++) {
                    const f32x2 ua0 = fma2(kx, z0p[j], fma2(ky, z1p[j], pa0_2));
                    const f32x2 ua1 = fma2(kz, z0p[j], fma2(kw, z1p[j], pa1_2));
                    const f32x2 ub0 = fma2(cs2[j], ua0, mul2(ns2[j], ua1));
                    const f32x2 ub1 = fma2(sn2[j], ua0, mul2(cs2[j], ua1));
                    const float m0 = fmaxf(fmaxf(fabsf(lo2(ub0)) - eb0[2 * j], fabsf(lo2(ub1)) - eb1[2 * j]),
                                           fmaxf(fabsf(lo2(ua0)) - ea0[2 * j], fabsf(lo2(ua1)) - ea1[2 * j]));
                    const float m1 = fmaxf(fmaxf(fabsf(hi2(ub0)) - eb0[2 * j + 1], fabsf(hi2(ub1)) - eb1[2 * j + 1]),
                                           fmaxf(fabsf(hi2(ua0)) - ea0[2 * j + 1], fabsf(hi2(ua1)) - ea1[2 * j + 1]));
                    hits += __float_as_uint(m0) >> 31;                      // sign bit of m = hit (m = -0 / NaN are undecided anyway)
                    hits += __float_as_uint(m1) >> 31;
                    mn = min3_nan_abs(mn, m0, m1);
                }
                // A lane with an undecided sample (also NaN) under this setting contributes a marker instead of its hits; its
                // samples are redone exactly after the loop.  The marker surfaces in the warp total, so the bookkeeping sits
                // behind a branch on a warp-uniform value that is taken about once in 200 iterations.
                const bool decided = mn > ee.y;
                const unsigned tot = __reduce_add_sync(0xffffffffu, decided ? (hits & mine_mask) : (mine_mask & 0x10000u));
                if (tot >= 0x10000u) {
                    const unsigned who = __ballot_sync(0xffffffffu, !decided && mine_mask != 0u);
                    if ((who >> lane) & 1u) und |= 1ull << r;
                }
                sp->cnt += tot & 0xffffu;                                // every lane writes the same value: no branch, no atomic
            }
            // rare: this lane's samples under the settings that had an undecided one, exactly (kept out of the loop above so
            // that the loop contains no call: with one, ptxas keeps the per-sample state in local memory)
            __syncwarp();
            if (mine && und != 0ull) {
                for (int r = 0; r < n_cov; r++)
                    if ((und >> r) & 1ull) atomicAdd(&S.s[r].cnt, sweep_super_slow(S, W, r, q, G, pid, p.keys, ev));
            }
            __syncwarp();
        }
        __syncwarp();
        // ragged ends: samples of [b, e) outside the full super-groups; lanes parallelise over settings
        const uint64_t head_end = (q_lo <= q_hi) ? ((q_lo << LG) < e ? (q_lo << LG) : e) : e;
        const uint64_t tail_begin = (q_lo <= q_hi) ? (q_hi << LG) : e;
        for (int part = 0; part < 2; part++) {
            const uint64_t lo = part ? (tail_begin > head_end ? tail_begin : head_end) : b;
            const uint64_t hi = part ? e : head_end;
            for (uint64_t sidx = lo; sidx < hi; sidx++) {
                float n[12];
                group_normals<3>((uint32_t)(sidx >> 2), (uint32_t)((sidx >> 2) >> 32), pid, p.keys, n);
                const int t = (int)(sidx & 3);
                float z0 = n[0], z1 = n[1], z2 = n[2];
                if (t == 1) { z0 = n[3]; z1 = n[4]; z2 = n[5]; }
                if (t == 2) { z0 = n[6]; z1 = n[7]; z2 = n[8]; }
                if (t == 3) { z0 = n[9]; z1 = n[10]; z2 = n[11]; }
                for (int r = lane; r < n_cov; r += 32) S.s[r].cnt += sweep_sample_slow(S, W, r, z0, z1, z2, ev);
            }
        }
        __syncwarp();
        for (int r = lane; r < n_cov; r += 32) {
            const unsigned long long tot = S.s[r].cnt;
            const uint64_t off = pair * hits_stride + (uint64_t)W.orig[r];
            SATMC_ASSERT(off < p.hits_len);
            if (p.n_chunks == 1) {
                if (p.flags & SATMC_ACCUMULATE) p.hits[off] += tot; else p.hits[off] = tot;
            } else {
                contribute(p, off, tot);
            }
        }
        __syncwarp();
    }
    finalize_counters(p);
}

// ---------------------------------------------------------------------------------------------
// general convex polygons: same work decomposition, sampler and counting as k_count;
// every sample is evaluated with the exact polygon SAT of satmc_poly.cuh (no screening pass yet)
// ---------------------------------------------------------------------------------------------
// Undecided samples of a warp, waiting for the exact pass: the screening pass decides most samples, and running the
// exact SAT for the rest lane by lane would leave the warp mostly idle.  Samples are appended with a ballot (warp-
// uniform fill count) and evaluated 32 at a time with all lanes busy -- by ONE copy of the exact code per vertex-count
// variant (poly_queue_drain), called once per trip of the sample loop.
// The first-level queue is drained when it holds kPolyDrainAt samples: several passes per (out-of-line) call, and the
// sample loop itself contains no call (with one, ptxas keeps the loop's state in local memory: 130-190 spill
// instructions per trip).  Capacity = what is left below the threshold + 4 x 32 appended by one more trip.
constexpr unsigned kPolyDrainAt = 160;
constexpr unsigned kPolyQueueCap = kPolyDrainAt - 1 + 128 + 1;
constexpr unsigned kPolyQueue2Cap = 64;                               // second level: 31 left over + 32 appended per pass
struct PolyQueue { float z[3][kPolyQueueCap]; float y[3][kPolyQueue2Cap]; };

// The exact polygon SAT of one sample: ONE out-of-line copy with run-time vertex counts (about 1 % of the samples get
// here; the straight-line variants of round 1 made the kernel 27 000 instructions long and the hot loop miss the
// instruction cache).
__device__ __noinline__ unsigned poly_exact_sample(const PolyPairShared& S, float z0, float z1, float z2)
{
    PolyRobotRegs R;
    poly_load_robot(S, R);
    return poly_collide<0, 0>(S, R, z0, z1, z2);
}

// Second level (poly_fast) on the samples the circle tests left open, 32 at a time with all lanes busy; the few it cannot
// decide either (|G| within the rounding bound of the true gap) move on to a second queue for the exact pass.  Out of
// line: one copy per vertex-count variant, and the sample loop's registers are not shared with it.  Returns the hits;
// fill / fill2 (warp-uniform) travel packed in `state` = fill | fill2 << 16.
template <int NR, int NO>
__device__ __noinline__ unsigned poly_queue_drain_fast(const PolyPairShared& S, PolyQueue& Q, unsigned& state, bool all, int lane,
                                                       unsigned long long* exact_evals)
{
    PolyRobotRegs R;
    poly_load_robot(S, R);
    unsigned fill = state & 0xffffu, fill2 = state >> 16;
    unsigned hit = 0;
    __syncwarp();
    while (fill >= 32u || (all && fill > 0u)) {
        const unsigned n = fill < 32u ? fill : 32u;
        fill -= n;
        const bool mine = (unsigned)lane < n;
        float z0 = 0.f, z1 = 0.f, z2 = 0.f;
        int r = 1;
        if (mine) {
            z0 = Q.z[0][fill + lane]; z1 = Q.z[1][fill + lane]; z2 = Q.z[2][fill + lane];
            r = poly_fast<NR, NO>(S, R, z0, z1, z2);
        }
        hit += (r == 2) ? 1u : 0u;
        const bool und = mine && r == 0;
        const unsigned mask = __ballot_sync(0xffffffffu, und);
        if (und) {
            const unsigned pos = fill2 + __popc(mask & ((1u << lane) - 1u));
            SATMC_ASSERT(pos < kPolyQueue2Cap);
            Q.y[0][pos] = z0; Q.y[1][pos] = z1; Q.y[2][pos] = z2;
        }
        fill2 += __popc(mask);
        __syncwarp();
        if (fill2 >= 32u) {                                             // a full pass of the exact SAT
            fill2 -= 32u;
            hit += poly_exact_sample(S, Q.y[0][fill2 + lane], Q.y[1][fill2 + lane], Q.y[2][fill2 + lane]);
            if (exact_evals && lane == 0) atomicAdd(exact_evals, 32ull);
            __syncwarp();
        }
    }
    if (all && fill2 > 0u) {
        if ((unsigned)lane < fill2) hit += poly_exact_sample(S, Q.y[0][lane], Q.y[1][lane], Q.y[2][lane]);
        if (exact_evals && lane == 0) atomicAdd(exact_evals, (unsigned long long)fill2);
        fill2 = 0u;
    }
    __syncwarp();
    state = fill | (fill2 << 16);
    return hit;
}

// one sample through the first screening level; undecided ones are queued.  `fill` is warp-uniform.
template <int NR, bool RANGE>
__device__ __forceinline__ unsigned poly_sample(const PolyPairShared& S, const PolyScreenRegs<NR>& C, PolyQueue& Q, unsigned& fill,
                                                bool valid, float z0, float z1, float z2, int lane)
{
    const int r = valid ? poly_screen<NR, RANGE>(S, C, z0, z1, z2) : 1;
    const bool und = r == 0;
    const unsigned mask = __ballot_sync(0xffffffffu, und);
    if (und) {
        const unsigned pos = fill + __popc(mask & ((1u << lane) - 1u));
        SATMC_ASSERT(pos < kPolyQueueCap);
        Q.z[0][pos] = z0; Q.z[1][pos] = z1; Q.z[2][pos] = z2;
    }
    fill += __popc(mask);
    return (valid && r == 2) ? 1u : 0u;
}

// the second level, one out-of-line copy per vertex-count variant
__device__ __forceinline__ unsigned poly_drain(const PolyPairShared& S, PolyQueue& Q, unsigned& fill, unsigned& fill2, bool all, int lane,
                                               unsigned long long* ev)
{
    unsigned state = fill | (fill2 << 16);                              // both warp-uniform
    const int shape = (S.nr == S.no) ? S.nr : 0;
    unsigned h;
    if (shape == 4) h = poly_queue_drain_fast<4, 4>(S, Q, state, all, lane, ev);
    else if (shape == 3) h = poly_queue_drain_fast<3, 3>(S, Q, state, all, lane, ev);
    else if (shape == 6) h = poly_queue_drain_fast<6, 6>(S, Q, state, all, lane, ev);
    else if (shape == 8) h = poly_queue_drain_fast<8, 8>(S, Q, state, all, lane, ev);
    else h = poly_queue_drain_fast<0, 0>(S, Q, state, all, lane, ev);
    fill = state & 0xffffu; fill2 = state >> 16;
    return h;
}

// samples of one work item; NR = robot vertex count known at compile time (0: read from S).  All 32 lanes walk the loops
// in step (ballots inside).
template <int NR, bool STREAMED>
__device__ __forceinline__ unsigned poly_chunk(const PolyPairShared& S, PolyQueue& Q, const CountParams& p,
                                               uint64_t pair, uint64_t c_begin, uint64_t c_len, int lane)
{
    unsigned cnt = 0, fill = 0, fill2 = 0;
    PolyScreenRegs<NR> C;
    poly_load_screen<NR>(S, C);
    unsigned long long* ev = (p.flags & SATMC_EXACT_ONLY) ? nullptr : p.exact_evals;
    if (STREAMED) {
        const float* z = p.z + pair * p.z_pair_stride + c_begin;
        uint64_t i0 = 0;
        while (i0 < c_len) {
            for (; i0 < c_len && fill < kPolyDrainAt; i0 += 32) {         // call-free inner loop
                const uint64_t i = i0 + (uint64_t)lane;
                const bool valid = i < c_len;
                const float z0 = valid ? __ldg(z + i) : 0.f, z1 = valid ? __ldg(z + p.ldz + i) : 0.f, z2 = valid ? __ldg(z + 2 * p.ldz + i) : 0.f;
                cnt += poly_sample<NR, true>(S, C, Q, fill, valid, z0, z1, z2, lane);
            }
            if (fill >= kPolyDrainAt) cnt += poly_drain(S, Q, fill, fill2, false, lane, ev);
        }
    } else {
        const uint32_t pid = p.pair_id_offset + (uint32_t)pair;
        const uint64_t b = p.sample_offset + c_begin, e = b + c_len;
        const uint64_t g_end = (e + 3) >> 2;
        uint64_t g0 = b >> 2;
        while (g0 < g_end) {
            for (; g0 < g_end && fill < kPolyDrainAt; g0 += 32) {         // 4-sample groups; call-free inner loop
                const uint64_t g = g0 + (uint64_t)lane;
                float n[12];
                group_normals<3>((uint32_t)g, (uint32_t)(g >> 32), pid, p.keys, n);
                if (4 * g0 >= b && 4 * (g0 + 31) + 3 < e) {              // warp-uniform: every sample of the trip is inside [b, e)
#pragma unroll
                    for (int t = 0; t < 4; t++) cnt += poly_sample<NR, false>(S, C, Q, fill, true, n[3 * t], n[3 * t + 1], n[3 * t + 2], lane);
                } else {                                                // ragged ends masked
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint64_t sidx = 4 * g + t;
                        cnt += poly_sample<NR, false>(S, C, Q, fill, sidx >= b && sidx < e, n[3 * t], n[3 * t + 1], n[3 * t + 2], lane);
                    }
                }
            }
            if (fill >= kPolyDrainAt) cnt += poly_drain(S, Q, fill, fill2, false, lane, ev);
        }
    }
    cnt += poly_drain(S, Q, fill, fill2, true, lane, ev);
    return cnt;
}

template <bool STREAMED>
__global__ void __launch_bounds__(kThreads, 2) k_count_poly(const float* __restrict__ pairs, const __grid_constant__ CountParams p)
{
    __shared__ PolyPairShared s_poly[kWarps];
    __shared__ PolyQueue s_queue[kWarps];
    __shared__ unsigned s_part[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * kWarps;
    for (uint64_t item = (uint64_t)blockIdx.x * kWarps + warp; item < p.n_items;) {
        const unsigned long long drawn = draw_ticket(p, lane);
        const uint64_t pair = item / p.n_chunks;
        const uint32_t chunk_id = (uint32_t)(item - pair * p.n_chunks);
        SATMC_ASSERT(pair < p.hits_len);
        __syncwarp();
        if (lane == 0) {
            const float* d = pairs + pair * 40;                                   // 160-byte descriptors
            poly_prologue(s_poly[warp], d);
            poly_screen_prologue(s_poly[warp], !(p.flags & SATMC_EXACT_ONLY));
            poly_fast_prologue(s_poly[warp], d[0], d[1], d + 8, !(p.flags & SATMC_EXACT_ONLY));
        }
        __syncwarp();
        const PolyPairShared& S = s_poly[warp];
        const uint64_t c_begin = (uint64_t)chunk_id * p.chunk;
        const uint64_t c_len = (c_begin + p.chunk <= p.n_samples) ? p.chunk : (p.n_samples - c_begin);
        unsigned cnt;                                                  // first level with the robot's normals in registers
        if (S.nr == 4) cnt = poly_chunk<4, STREAMED>(S, s_queue[warp], p, pair, c_begin, c_len, lane);
        else if (S.nr == 3) cnt = poly_chunk<3, STREAMED>(S, s_queue[warp], p, pair, c_begin, c_len, lane);
        else if (S.nr == 6) cnt = poly_chunk<6, STREAMED>(S, s_queue[warp], p, pair, c_begin, c_len, lane);
        else if (S.nr == 8) cnt = poly_chunk<8, STREAMED>(S, s_queue[warp], p, pair, c_begin, c_len, lane);
        else cnt = poly_chunk<0, STREAMED>(S, s_queue[warp], p, pair, c_begin, c_len, lane);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (p.block_uniform) {
            if (lane == 0) s_part[warp] = cnt;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long t = 0;
#pragma unroll
                for (int w = 0; w < kWarps; w++) t += s_part[w];
                contribute(p, pair, t);
            }
            __syncthreads();
        } else if (lane == 0) {
            if (p.n_chunks == 1) {
                if (p.flags & SATMC_ACCUMULATE) p.hits[pair] += cnt; else p.hits[pair] = cnt;
            } else {
                contribute(p, pair, (unsigned long long)cnt);
            }
        }
        item = next_item(p, item, stride, drawn);
    }
    finalize_counters(p);
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_decide(satmc_pair const* pair, const float* __restrict__ z, uint64_t ldz, int ndof, uint64_t n,
                         uint8_t* out, uint32_t flags, unsigned long long* exact_evals)
{
    __shared__ float s_robot[8];
    float v[12];
    DirectSrc src{pair};
    src.load(0, v);
    PairConst P;
    pair_const_init(P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
    if (flags & SATMC_EXACT_ONLY) { P.eps = CUDART_INF_F; P.eps_b = CUDART_INF_F; }
    if (threadIdx.x == 0) {
        float base[8];
        rect_base(v[3], v[4], base);
        exact_robot_corners(v[0], v[1], P.ca, P.sa, base, s_robot);
    }
    __syncthreads();
    unsigned long long* ev = (flags & SATMC_EXACT_ONLY) ? nullptr : exact_evals;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float z0 = z[i], z1 = z[ldz + i], z2 = z[2 * ldz + i];
        unsigned h;
        if (ndof == 5) h = streamed_sample<5>(P, s_robot, z0, z1, z2, z[3 * ldz + i], z[4 * ldz + i], ev);
        else           h = streamed_sample<3>(P, s_robot, z0, z1, z2, 0.f, 0.f, ev);
        out[i] = (uint8_t)h;
    }
}

// normals of samples [offset, offset+n) of stream pid, D planes (D = 3 or 5); one thread per group
// diagnostics: the screening value m and the threshold it is compared with, per streamed sample of one pair
__global__ void k_screen_debug(satmc_pair const* pair, const float* __restrict__ z, uint64_t ldz, int ndof, uint64_t n,
                               float* m_out, float* eps_out)
{
    float v[12];
    DirectSrc src{pair};
    src.load(0, v);
    PairConst P;
    pair_const_init(P, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float hmin;
        if (ndof == 5) {
            m_out[i] = screen_gap<5>(P, z[i], z[ldz + i], z[2 * ldz + i], z[3 * ldz + i], z[4 * ldz + i], hmin);
            eps_out[i] = (hmin > 0.0f) ? P.eps_a + P.eps_b / hmin : CUDART_INF_F;
        } else {
            m_out[i] = screen_gap<3>(P, z[i], z[ldz + i], z[2 * ldz + i], 0.f, 0.f, hmin);
            eps_out[i] = P.eps;
        }
    }
}

template <int D>
__global__ void k_fused_normals(const __grid_constant__ PhiloxKeys K, uint32_t pid, uint64_t offset, uint64_t n, float* z, uint64_t ldz)
{
    const uint64_t g0 = offset >> 2, g1 = (offset + n + 3) >> 2;
    for (uint64_t g = g0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < g1; g += (uint64_t)gridDim.x * blockDim.x) {
        float nn[4 * D];
        group_normals<D>((uint32_t)g, (uint32_t)(g >> 32), pid, K, nn);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint64_t s = 4 * g + t;
            if (s < offset || s >= offset + n) continue;
#pragma unroll
            for (int k = 0; k < D; k++) { SATMC_ASSERT(s - offset < ldz); z[k * ldz + (s - offset)] = nn[D * t + k]; }
        }
    }
}

__global__ void k_philox(const uint32_t* ctr, uint64_t n, const __grid_constant__ PhiloxKeys K, uint32_t* out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], K, w);
    out[4 * i] = w[0]; out[4 * i + 1] = w[1]; out[4 * i + 2] = w[2]; out[4 * i + 3] = w[3];
}

__global__ void k_sat_corners(const float* __restrict__ r1, const float* __restrict__ r2, uint64_t n, uint8_t* out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[8], b[8];
    const float4* pa = reinterpret_cast<const float4*>(r1 + 8 * i);
    const float4* pb = reinterpret_cast<const float4*>(r2 + 8 * i);
    const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    out[i] = (uint8_t)exact_convex_collide(a, b);
}

// calcSlack / getBin / done flag of the reference kernel tail (ztest.cu:156-165, utils.cu:186-207).
// k*k is formed in 64 bits (the reference's int32 product wraps for k > 46340); bins are read in bounds.
__device__ __forceinline__ float calc_slack(int n, int k)
{
    const float z = 1.96;
    const float alpha = 0.025;
    if (k == n || k == 0) return (float)(log(1.0 / alpha) / n);
    const float kk = (float)((long long)k * (long long)k);
    return z / n * sqrtf((float)k - kk / (float)n);
}

__global__ void k_ztest_tail(const unsigned long long* __restrict__ hits, float* cps, const float* __restrict__ bins,
                             const float* __restrict__ bin_acc, int n_bins, int* done, int n_samples, int num_left,
                             const int* __restrict__ live)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= num_left) return;
    const int e = live ? live[g] : g;
    SATMC_ASSERT(e >= 0 && (live != nullptr || e < num_left));
    const int k = (int)cps[e] + (int)hits[g];
    const float slack = calc_slack(n_samples, k);
    const float p = (float)k / (float)n_samples;
    int bin = 0;
    for (int i = 0; i + 1 < n_bins; i++)
        if (p >= bins[i] && p <= bins[i + 1]) bin = i;
    done[e] = (slack <= bin_acc[bin]) ? 1 : 0;
    cps[e] = (float)k;
}

// ---- adaptive scheduler (replaces thrust::count + sort_by_key + tail copies, ztest.cu:359-371) ----
__global__ void k_iota(int* a, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// Tail of one adaptive iteration in one kernel: the reference kernel's z-test tail (ztest.cu:156-165) for the pair of
// every live slot, then the split of the live list: a finished pair gets its probability written in input order
// (count / n_samples, the arithmetic of write_collision_probability utils.cu:210-215) and leaves the list, the others
// are appended to live_out -- warp-aggregated, one atomic per warp (the reference: thrust::count + sort_by_key + 4-5
// blocking copies, ztest.cu:359-371).  n_out must be zero on entry; n_next (the other of the two counters, used by
// the next iteration) is cleared here, so the loop needs no memset.
__global__ void k_ztest_compact(const unsigned long long* __restrict__ hits, float* counts, const float* __restrict__ bins,
                                const float* __restrict__ bin_acc, int n_bins, int n_samples, int num_left,
                                const int* __restrict__ live_in, float* cp_out, int* n_samples_out, int* live_out, int* n_out,
                                int* n_next, int n_pairs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_next = 0;
    const bool valid = i < num_left;
    int e = 0; bool keep = false;
    if (valid) {
        e = live_in[i];
        SATMC_ASSERT(e >= 0 && e < n_pairs);
        const int k = (int)counts[e] + (int)hits[i];
        const float slack = calc_slack(n_samples, k);
        const float p = (float)k / (float)n_samples;
        int bin = 0;
        for (int b = 0; b + 1 < n_bins; b++)
            if (p >= bins[b] && p <= bins[b + 1]) bin = b;
        const bool fin = slack <= bin_acc[bin];
        counts[e] = (float)k;
        if (fin) {
            cp_out[e] = (float)k / (float)n_samples;
            if (n_samples_out) n_samples_out[e] = n_samples;
        }
        keep = !fin;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(n_out, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (keep) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        SATMC_ASSERT(pos >= 0 && pos < num_left);
        live_out[pos] = e;
    }
}

// pairs still live when the loop stops at max_samples: probability from the counts as they stand (ztest.cu:376-385)
__global__ void k_compact_live(const int* __restrict__ live_in, int num_left, const float* __restrict__ counts, int n_samples,
                               float* cp_out, int* n_samples_out, int n_pairs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_left) return;
    const int e = live_in[i];
    SATMC_ASSERT(e >= 0 && e < n_pairs);
    cp_out[e] = counts[e] / (float)n_samples;
    if (n_samples_out) n_samples_out[e] = n_samples;
}

// iteration-0 draw of generate_dataset (generate_dataset.cu:207-219): pose index, std-dev index and a
// robot position on the ring prior around the obstacle.  Philox stream = stream_id_offset + g, counter
// block 0xffffffff (disjoint from the sample groups).
__global__ void k_sample_positions(const __grid_constant__ PhiloxKeys K, const float* __restrict__ poses, uint32_t n_poses,
                                   const float* __restrict__ std_devs, uint32_t n_std, int n, float r_offset, float spread,
                                   uint32_t stream_id_offset, float* positions, float* pose_idxs, float* std_dev_idxs)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t w[4], w2[4];
    philox4x32_10(0xffffffffu, 0xffffffffu, stream_id_offset + (uint32_t)g, 0u, K, w);
    philox4x32_10(0xffffffffu, 0xffffffffu, stream_id_offset + (uint32_t)g, 1u, K, w2);
    const uint32_t pi = w[0] % n_poses, si = w[1] % n_std;                    // curand() % n   (:208-209)
    const float pw = poses[3 * (size_t)pi], ph = poses[3 * (size_t)pi + 1];
    const float sx = std_devs[5 * (size_t)si], sy = std_devs[5 * (size_t)si + 1];
    const float uni = ((float)(w[2] >> 8) + 1.0f) * 5.9604645e-8f;            // curand_uniform: (0, 1]   (:213)
    const float theta = (float)(uni * 2 * 3.14159265358979323846);
    float nz, unused;
    bm_pair(w2[0], w2[1], nz, unused);                                        // curand_normal            (:214)
    const float shift = nz * ((sy + sx) / 2) * spread;
    positions[2 * (size_t)g]     = (float)(cosf(theta) * ((pw / 2 + r_offset + 2.35 + sx) + shift));   // (:215)
    positions[2 * (size_t)g + 1] = (float)(sinf(theta) * ((ph / 2 + r_offset + 2.35 + sy) + shift));   // (:216)
    pose_idxs[g] = (float)pi;
    std_dev_idxs[g] = (float)si;
}

__global__ void k_write_cp(float* counts, int n_done, int n_samples)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_done) counts[g] = counts[g] / (float)n_samples;
}

__global__ void k_hits_to_cp(const unsigned long long* __restrict__ hits, uint64_t n, uint64_t n_samples, float* cp)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cp[i] = (float)hits[i] / (float)n_samples;
}

}  // namespace satmc
