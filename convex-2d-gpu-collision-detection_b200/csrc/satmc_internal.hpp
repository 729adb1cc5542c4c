// satmc_internal.hpp -- shared between the translation units of libsatmc.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/satmc.h"

struct satmc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int blocks_per_sm = 0;
    int blocks_per_sm_streamed = 0;
    int blocks_per_sm_tma[2][2] = {{0, 0}, {0, 0}};   // bulk-copy staged streamed kernel, [ndof 3 / 5][private / shared bank]
    char err[512] = {0};
    uint64_t launches = 0;
    unsigned long long* d_exact_evals = nullptr;
    // work-item counters of the dynamically scheduled kernels: never reset, the host mirrors their values (every processed
    // item draws exactly one ticket).  Two of them: launches on the auxiliary stream (pipelined host calls) may run
    // concurrently with launches on the main stream and must not share a counter.
    unsigned long long* d_ticket = nullptr;
    uint64_t ticket_next[2] = {0, 0};
    int ticket_sel = 0;
    cudaStream_t aux = nullptr;              // pipelined host calls: second slice (created on first use)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    bool last_ms_valid = false;
    bool events_by_caller = false;           // a pipelined host call brackets both of its launches itself
    // grow-only device scratch
    void* d_scratch[3] = {nullptr, nullptr, nullptr};
    size_t scratch_cap[3] = {0, 0, 0};
    // zero-invariant accumulators of launches with several work items per counter (CountParams::acc), one set per
    // stream selector like the tickets; d_blocks_done sits behind the two tickets
    unsigned long long* d_acc[2] = {nullptr, nullptr};
    size_t acc_cap[2] = {0, 0};
    unsigned* d_blocks_done = nullptr;
    int* h_word = nullptr;                   // pinned: the adaptive loop's "pairs left" comes back here
    uint64_t tune_min_chunk = 2048, tune_tiny_chunk = 256;   // planner: samples per work item (see plan_items)
    uint64_t tune_stream_chunk[2] = {2048, 8192};            // planner: most samples per work item, bulk-tensor path on private banks (3-DoF, 5-DoF)
    uint64_t tune_fused_max_chunk = 1ull << 18;              // planner: most samples per work item of the fused kernels (see launch_count)
    int tune_tiny_bps = 0;                                   // planner: blocks per SM a tiny launch is cut for (0: the resident number)
    int tune_stream_ipw = 8;                                 // planner: work items per resident warp, streamed bulk-tensor path
};

// State of one adaptive z-test loop (satmc_adaptive_run) on one context, in steps: begin, then while pending
// { enqueue; synchronise the stream; collect }, then finish.
struct AdaptiveRun {
    satmc_ctx* ctx;
    const float *d_robot_base, *d_poses, *d_std_devs, *d_pose_idxs, *d_std_dev_idxs, *d_positions, *d_bins, *d_bin_acc;
    uint32_t n_poses, n_std;
    int n_bins, n_pairs, max_samples, n_batch_small, switch_at, n_batch_large;
    uint64_t seed;
    uint32_t stream_id_offset, stream_id_stride;      // pair i draws Philox stream offset + i * stride
    float* d_cp_out; int* d_n_samples_out;
    // state
    float* d_counts; int* d_live[2]; int* d_n;
    int cur, num_left, n_samples, iter;
    long long drawn;
};
int satmc_adaptive_begin(AdaptiveRun& ar);
bool satmc_adaptive_pending(const AdaptiveRun& ar);
int satmc_adaptive_enqueue(AdaptiveRun& ar);
void satmc_adaptive_collect(AdaptiveRun& ar);
int satmc_adaptive_finish(AdaptiveRun& ar);



// text of the last error of calls made without a context (satmc_create failures, NULL ctx)
char* satmc_thread_error();

inline int satmc_fail(satmc_ctx* ctx, int code, const char* fmt, ...)
{
    char* dst = ctx ? ctx->err : satmc_thread_error();
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define fail satmc_fail

#define CU(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return satmc_fail((ctx), SATMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// internal entry points shared with the group code (satmc_group.cu)
#define SATMC_PEER_ATOMIC_OUT 0x80000000u
int satmc_count_fused_impl(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                           uint64_t sample_offset, uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags);
bool satmc_fused_is_multi(satmc_ctx* ctx, uint64_t n_pairs, uint64_t n_samples);
