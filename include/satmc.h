/*
 * satmc.h -- C ABI of the B200-native Monte Carlo SAT collision-probability path.
 *
 * Drop-in boundary for the one hot path of beautifulv0id/Convex-2D-GPU-Collision-Detection:
 * "for each (robot rectangle, Gaussian-uncertain obstacle rectangle) pair draw N perturbed obstacle
 * poses, run the separating-axis test on each, count the hits".  The reference has no library API;
 * its three programs launch one kernel, so every entry point below names the reference code it
 * replaces (paths relative to the reference repository root).  INTEGRATION.md shows the edit a
 * reference maintainer makes in each main().
 *
 * Conventions
 *   - plain C, no CUDA/torch types: device pointers are `void*`-compatible raw addresses, the
 *     stream is passed as `void*` (a cudaStream_t) at context creation.
 *   - every function returns SATMC_OK (0) or a negative satmc_status; nothing inside the library
 *     calls exit() (the reference's gpuAssert does, utils.cu:59-67).  satmc_last_error() gives text.
 *   - "d_" parameters are device pointers on the context's device, "h_" are host pointers.
 *     The *_host variants copy H2D/D2H on the context's stream and synchronise it before returning.
 *   - all work is enqueued on the context's stream; device-pointer entry points are asynchronous.
 *   - one context per host thread / per GPU; contexts share no state.
 *   - there is no CPU fallback: with no usable sm_100 device satmc_create() fails with
 *     SATMC_ERR_NO_DEVICE.
 *
 * Arithmetic contract: for given normals z the collide decision of every
 * sample is bit-identical to what the reference's compiled kernel (nvcc 12.9, default flags,
 * sm_100a) decides: sample_rectangle utils.cu:144-157 -> convex_collide utils.cu:159-184, 8 axes,
 * strict <, NaN => collide.
 */
#ifndef SATMC_H
#define SATMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SATMC_VERSION_MAJOR 0
#define SATMC_VERSION_MINOR 2

typedef enum {
    SATMC_OK              =  0,
    SATMC_ERR_INVALID     = -1,   /* bad argument (null pointer, ndof not 3/5, misaligned ...)      */
    SATMC_ERR_NO_DEVICE   = -2,   /* no CUDA device / not compute capability 10.x                    */
    SATMC_ERR_CUDA        = -3,   /* a CUDA runtime call failed; see satmc_last_error()              */
    SATMC_ERR_NOMEM       = -4,   /* host or device allocation failed                                */
    SATMC_ERR_NCCL        = -5    /* NCCL is unavailable (libnccl.so.2 not loadable) or an NCCL call failed    */
} satmc_status;

/* One (robot, uncertain obstacle) pair in direct form: 12 packed float32 = 48 bytes.
 * Coordinates are in the obstacle's nominal frame, as in the reference: the obstacle is an
 * ow x oh rectangle centred at the origin with heading 0 (create_rect, ztest.cu:143-144), the robot
 * an rw x rh rectangle with heading rtheta centred at (rx, ry) (ztest.cu:148-149,297).
 * sd_* are the STANDARD DEVIATIONS of the obstacle pose/shape perturbation (StdDev, utils.cu:86-89,
 * 107; the .npy files hold variances, the reference takes sqrt at load, ztest.cu:245-251). */
typedef struct satmc_pair {
    float rx, ry;        /* Position            utils.cu:74-77   */
    float rtheta;        /* Pose.theta          utils.cu:91-94   */
    float rw, rh;        /* --robot_width / --robot_height, defaults 4.07 x 1.74 */
    float ow, oh;        /* Pose.width / .height                 */
    float sd_x, sd_y, sd_theta, sd_w, sd_h;
} satmc_pair;

typedef struct satmc_ctx satmc_ctx;

/* Flags for the counting entry points. */
#define SATMC_ACCUMULATE  0x1u   /* add to d_hits instead of overwriting it                          */
#define SATMC_EXACT_ONLY  0x2u   /* evaluate every sample with the exact 8-axis reference arithmetic */
                                 /* (no screening pass); same results, used for A/B verification     */

/* ---- context ------------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device`, enqueueing on `stream` (a cudaStream_t cast to void*;
 * NULL = the legacy default stream, which is what the reference uses throughout). */
int satmc_create(int device, void* stream, satmc_ctx** out);
int satmc_destroy(satmc_ctx* ctx);
/* Blocks until everything enqueued through `ctx` has finished. */
int satmc_synchronize(satmc_ctx* ctx);
/* Text of the last error on this context ("" if none); never NULL.  ctx may be NULL. */
const char* satmc_last_error(const satmc_ctx* ctx);
const char* satmc_version(void);
/* Number of kernels this context has launched so far (bench.py's gpu_launches claim). */
uint64_t satmc_launch_count(const satmc_ctx* ctx);
/* Time in milliseconds the device spent in the most recent counting kernel launched by a *_host
 * call or by a call made while profiling is enabled (CUDA events on the context's stream). */
int satmc_set_profiling(satmc_ctx* ctx, int enabled);
float satmc_last_kernel_ms(const satmc_ctx* ctx);

/* ---- the hot path -------------------------------------------------------------------------- */

/* Fused path: counter-based Philox4x32-10 -> Box-Muller normals -> perturbed obstacle -> SAT ->
 * hit count, samples never touch HBM.
 *   replaces: setup_kernel utils.cu:111-117 + the sample loop of
 *             monte_carlo_sample_collision_dataset_uniform ztest.cu:151-155
 *             (= compute_collision_probability.cu:135-139, generate_dataset.cu:238-242)
 * d_hits[i] = #{ s in [sample_offset, sample_offset + n_samples) : pair i collides on sample s }.
 * The normals of sample s of pair i depend only on (seed, pair_id_offset + i, s) and on whether the
 * pair has shape variance: any split of the
 * pair range or the sample range over calls, streams, GPUs or ranks gives bit-identical totals.
 * Pairs with sd_w == sd_h == 0 take a 3-normal path (x, y, theta), others draw all five. */
int satmc_count_fused(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs,
                      uint64_t n_samples, uint64_t seed, uint64_t sample_offset,
                      uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags);

/* Covariance sweep with common random numbers (the ztest-style variance sweep, BASELINE config 5): every pair is
 * evaluated under n_cov pose-covariance settings h_sigmas[c] = (sd_x, sd_y, sd_theta) on the same normals (64 settings
 * per kernel launch).  The settings are a HOST array (3 * n_cov floats, read before the call returns): they travel to
 * the kernels as launch parameters.
 * d_hits[i*n_cov + c] equals what satmc_count_fused returns for pair i with sd_* = d_sigmas[c], sd_w = sd_h = 0 and
 * the same (seed, pair id, sample range) -- bit for bit -- but the sampler runs once per sample instead of once
 * per (sample, setting).  The sd_* fields of d_pairs are ignored. */
int satmc_count_fused_sweep(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs, const float* h_sigmas,
                            uint32_t n_cov, uint64_t n_samples, uint64_t seed, uint64_t sample_offset,
                            uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags);

/* Streamed path: the normals are supplied by the caller (verification against the reference on
 * shared samples; HBM-bound).  d_z is SoA: plane k (k = 0..ndof-1 = x, y, theta[, w, h]) of sample
 * s is d_z[k*ldz + s].  Pair i consumes samples [i*z_pair_stride, i*z_pair_stride + n_samples) of
 * every plane (z_pair_stride = 0: all pairs share one bank -- common random numbers).  ndof = 3
 * treats dw = dh = 0.  Fastest when d_z is 16-byte aligned and ldz, z_pair_stride are multiples of 4. */
int satmc_count_streamed(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs,
                         const float* d_z, uint64_t ldz, uint64_t z_pair_stride, int ndof,
                         uint64_t n_samples, uint64_t* d_hits, uint32_t flags);

/* Per-sample decisions of ONE pair on supplied normals: d_out[s] = 1 collide / 0 separated.
 * (what `n_samplestrue += convex_collide(...)` adds at ztest.cu:154, sample by sample) */
int satmc_decide_streamed(satmc_ctx* ctx, const satmc_pair* d_pair, const float* d_z, uint64_t ldz,
                          int ndof, uint64_t n_samples, uint8_t* d_out, uint32_t flags);

/* The normals the fused path uses for samples [sample_offset, sample_offset+n) of pair `pair_id`,
 * written as ndof SoA planes d_z[k*ldz + s].  ndof must be the pair's own: 3 if its sd_w == sd_h == 0,
 * else 5 (the two kinds of pair consume the Philox stream differently).  Feeding the planes to
 * satmc_count_streamed / the oracle must reproduce satmc_count_fused's count exactly. */
int satmc_fused_normals(satmc_ctx* ctx, uint64_t seed, uint32_t pair_id, uint64_t sample_offset,
                        uint64_t n, int ndof, float* d_z, uint64_t ldz);

/* Raw Philox4x32-10 blocks (known-answer tests): d_out[4*i..4*i+3] = philox(ctr = d_ctr[4*i..], key). */
int satmc_philox_blocks(satmc_ctx* ctx, const uint32_t* d_ctr, uint64_t n, uint32_t key0, uint32_t key1,
                        uint32_t* d_out);

/* SAT on explicit corner sets (AoS x0,y0..x3,y3, [n][8] each): d_out[i] = convex_collide(r1_i, r2_i).
 *   replaces: convex_collide utils.cu:159-184 (BASELINE config 1 on the GPU) */
int satmc_sat_corners(satmc_ctx* ctx, const float* d_r1, const float* d_r2, uint64_t n, uint8_t* d_out);

/* Diagnostics: the screening value m (largest normalised signed gap) and the threshold it is
 * compared with, for every supplied sample of ONE pair.  A sample is decided by the screening pass iff
 * |m| > eps (and its normals are within the bound); tests use this to measure the safety margin of eps. */
int satmc_screen_debug(satmc_ctx* ctx, const satmc_pair* d_pair, const float* d_z, uint64_t ldz, int ndof,
                       uint64_t n_samples, float* d_m_out, float* d_eps_out);

/* Diagnostics: number of samples that the screening pass could not decide and that were re-evaluated
 * with the exact arithmetic, accumulated over all counting calls since the last reset. */
int satmc_exact_evals(satmc_ctx* ctx, uint64_t* out, int reset);

/* Diagnostics: how a call would be cut into work items (kind 0 = fused rectangles, 1 = streamed, 2 = polygons, 3 = sweep):
 * samples per item and items per pair.  An item's hits are summed in 32 bits, so no item exceeds 2^31 samples. */
int satmc_plan_debug(satmc_ctx* ctx, int kind, uint64_t n_pairs, uint64_t n_samples, uint64_t* chunk_out, uint64_t* n_chunks_out);

/* ---- general convex polygons ---------------------------------------------------------------- */

/* The reference handles rectangles only and notes that SAT "can easily be extended" to other convex shapes
 * (README.md:3).  These entry points are that extension: robot and obstacle are convex polygons with 1..8
 * counter-clockwise vertices; the obstacle's pose is perturbed by z0*sd_x, z1*sd_y, z2*sd_theta exactly as
 * sample_rectangle does (utils.cu:144-157, no shape perturbation); the separating axes are the true edge normals
 * of both polygons (convex_collide's edge directions, utils.cu:170-171, are valid axes only for rectangles);
 * projections, strict-< separation and tie/NaN behaviour follow utils.cu:172-181.  160 bytes, 16-byte aligned. */
#define SATMC_POLY_MAX 8
typedef struct satmc_poly_pair {
    float rx, ry, rtheta;              /* robot pose in the obstacle's nominal frame                      */
    float sd_x, sd_y, sd_theta;        /* standard deviations of the obstacle pose perturbation            */
    uint32_t n_robot, n_obstacle;      /* vertex counts, 1..SATMC_POLY_MAX                                 */
    float robot[2 * SATMC_POLY_MAX];   /* robot vertices x0,y0,.. in the robot's own frame                 */
    float obstacle[2 * SATMC_POLY_MAX];/* obstacle vertices in its nominal frame                           */
} satmc_poly_pair;

/* Same semantics as satmc_count_fused / satmc_count_streamed (3 normals per sample: x, y, theta).
 * SATMC_EXACT_ONLY bypasses the polygon screening pass (every sample through the exact SAT); counts are identical. */
int satmc_count_fused_polygons(satmc_ctx* ctx, const satmc_poly_pair* d_pairs, uint64_t n_pairs,
                               uint64_t n_samples, uint64_t seed, uint64_t sample_offset,
                               uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags);
int satmc_count_streamed_polygons(satmc_ctx* ctx, const satmc_poly_pair* d_pairs, uint64_t n_pairs,
                                  const float* d_z, uint64_t ldz, uint64_t z_pair_stride,
                                  uint64_t n_samples, uint64_t* d_hits, uint32_t flags);

/* ---- reference-compatible Monte Carlo step ------------------------------------------------- */

/* One launch of the reference kernel, same argument meaning as
 *   monte_carlo_sample_collision_dataset_uniform  ztest.cu:106-121
 *   (= compute_collision_probability.cu:90-105; generate_dataset.cu:175-194 without its iteration-0
 *   position sampling, which satmc_sample_positions provides)
 * d_robot_base: 8 floats from create_rect(robot_w, robot_h); d_poses: [n_poses][3] (width, height,
 * theta); d_std_devs: [n_std][5]; d_pose_idxs / d_std_dev_idxs: float indices per live pair;
 * d_positions: [num_left][2]; d_cps: running hit COUNT as float, in/out (ztest.cu:135,165);
 * d_accuracy_bins [n_accuracy_bins], d_bin_accuracy [n_accuracy_bins-1]; d_done: out flags.
 * n_samples = cumulative samples AFTER this step, n_batch = samples added now, first num_left
 * entries are live.  The curandState* of the reference is replaced by (seed, stream_id_offset):
 * pair g draws the Philox stream of id stream_id_offset + g at sample indices
 * [n_samples - n_batch, n_samples).  Out-of-range bin reads of the reference (utils.cu:202) are not
 * replicated: the last bin edge is treated as closed and nothing is read past the arrays. */
int satmc_mc_step(satmc_ctx* ctx, const float* d_robot_base, const float* d_poses, uint32_t n_poses,
                  const float* d_std_devs, uint32_t n_std, const float* d_pose_idxs,
                  const float* d_std_dev_idxs, const float* d_positions, float* d_cps,
                  const float* d_accuracy_bins, const float* d_bin_accuracy, int n_accuracy_bins,
                  int* d_done, int iteration, int n_samples, int n_batch, int num_left,
                  uint64_t seed, uint32_t stream_id_offset);

/* count -> probability on a finished tail.   replaces: write_collision_probability utils.cu:210-215 */
int satmc_write_collision_probability(satmc_ctx* ctx, float* d_counts, int n_done, int n_samples);

/* The whole adaptive-sampling loop of the three programs in one call
 *   replaces: ztest.cu:328-388, compute_collision_probability.cu:276-333, generate_dataset.cu:420-479
 *             (while loop: kernel launch, thrust::count, thrust::sort_by_key compaction,
 *              write_collision_probability on the finished tail, 4-5 blocking D2H copies per iteration)
 * Tables and per-pair arrays as in satmc_mc_step (n_pairs entries, never permuted).  Every iteration
 * adds n_batch = (n_samples < switch_at ? n_batch_small : n_batch_large) samples to the unfinished
 * pairs (generate_dataset.cu:427-431: 1000 / 20000 / 100000; ztest.cu:332: 10000 fixed), evaluates the
 * z-test stop rule, and retires finished pairs through a device-side work list -- nothing is sorted
 * and nothing but one int per iteration crosses PCIe.  Stops at max_samples.
 * d_cp_out[i] = hits_i / samples_i (float division, utils.cu:214), in input order (the reference needs
 * an index array and a host-side un-permute for that, ztest.cu:390-406).  d_n_samples_out (optional)
 * receives the samples each pair used.  Pair i always draws Philox stream stream_id_offset + i, so the
 * result does not depend on the order in which pairs finish. */
int satmc_adaptive_run(satmc_ctx* ctx, const float* d_robot_base, const float* d_poses, uint32_t n_poses,
                       const float* d_std_devs, uint32_t n_std, const float* d_pose_idxs,
                       const float* d_std_dev_idxs, const float* d_positions, int n_pairs,
                       const float* d_accuracy_bins, const float* d_bin_accuracy, int n_accuracy_bins,
                       int max_samples, int n_batch_small, int switch_at, int n_batch_large,
                       uint64_t seed, uint32_t stream_id_offset, float* d_cp_out, int* d_n_samples_out,
                       int* iterations_out, long long* samples_drawn_out);

/* Iteration-0 draw of generate_dataset: pose index, std-dev index and a robot position on the ring
 * prior around the obstacle, for n data points.
 *   replaces: generate_dataset.cu:207-219 (curand() % n, curand_uniform, curand_normal -> Philox) */
int satmc_sample_positions(satmc_ctx* ctx, const float* d_poses, uint32_t n_poses, const float* d_std_devs,
                           uint32_t n_std, int n, float r_offset, float spread, uint64_t seed,
                           uint32_t stream_id_offset, float* d_positions, float* d_pose_idxs,
                           float* d_std_dev_idxs);

/* Device memory helpers so that host programs need no CUDA headers (cudaMalloc / cudaFree / blocking
 * cudaMemcpy on the context's device and stream; the reference calls these directly, ztest.cu:269-304). */
int satmc_device_alloc(satmc_ctx* ctx, void** out, size_t bytes);
int satmc_device_free(satmc_ctx* ctx, void* p);
int satmc_upload(satmc_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int satmc_download(satmc_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* Enqueues the copy and returns; h_dst should be pinned (satmc_host_alloc) and is valid after satmc_synchronize. */
int satmc_download_async(satmc_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* ---- groups: the multi-GPU split of the path ------------------------------------------------------ */

/* The reference is single-GPU (no cudaSetDevice, streams or collectives anywhere; a single pair with N = 1e11 is
 * impossible there: `int n_samples`, float counter, ztest.cu:331,135,165).  The path shards with no data dependence
 * (SURVEY.md section 8e), and because the normals of sample s of pair p depend only on (seed, p, s) every split returns
 * the single-GPU counts bit for bit:
 *   SATMC_SHARD_BY_PAIR          rank r owns pairs [r*c, (r+1)*c), c = ceil(n_pairs / world); outputs are disjoint, the
 *                                only exchange is an all-gather of the counters (none inside a single process)
 *   SATMC_SHARD_BY_SAMPLE_RANGE  rank r owns a range of the sample indices of every pair (cut at multiples of 4); ONE
 *                                ncclAllReduce(ncclUint64, ncclSum) of the n_pairs counters over NVLink is the exchange
 *   SATMC_SHARD_INTERLEAVED      rows r, r + world, ...: the adaptive z-test, whose work per row varies ~400x
 *                                (generate_dataset.cu:53,427-431)
 * A group is either one process driving n devices (satmc_group_create: ncclCommInitAll, one context and stream per
 * device, launches enqueued from the calling thread) or one process per GPU (satmc_group_create_rank: rank 0 calls
 * satmc_group_unique_id, the launcher distributes the 128 bytes, every rank joins with ncclCommInitRank).  NCCL is bound
 * at run time (dlopen of libnccl.so.2); a group of world size 1 needs none. */
#define SATMC_SHARD_BY_PAIR          0
#define SATMC_SHARD_BY_SAMPLE_RANGE  1
#define SATMC_SHARD_INTERLEAVED      2
#define SATMC_UNIQUE_ID_BYTES        128

typedef struct satmc_group satmc_group;

/* Slice of `rank` (host arithmetic only, no GPU needed).  BY_PAIR / BY_SAMPLE_RANGE: units [*lo, *hi).
 * INTERLEAVED: *lo = first row, *hi = number of rows (rows *lo, *lo + world, ...). */
int satmc_shard_range(int shard_mode, uint64_t n_units, int world, int rank, uint64_t* lo, uint64_t* hi);

int satmc_group_create(const int* devices /* NULL = 0..n_dev-1 */, int n_dev, satmc_group** out);
int satmc_group_unique_id(void* out /* SATMC_UNIQUE_ID_BYTES */);
/* `stream`: the cudaStream_t this rank's work and collectives are enqueued on (NULL = the legacy default stream). */
int satmc_group_create_rank(const void* unique_id, int world, int rank, int device, void* stream, satmc_group** out);
int satmc_group_destroy(satmc_group* g);
int satmc_group_synchronize(satmc_group* g);
int satmc_group_world(const satmc_group* g);
int satmc_group_local_count(const satmc_group* g);          /* devices driven by this process */
int satmc_group_rank(const satmc_group* g, int local);      /* rank of local device `local` */
satmc_ctx* satmc_group_context(satmc_group* g, int local);  /* its context (owned by the group) */
const char* satmc_group_last_error(const satmc_group* g);
int satmc_group_nccl_version(void);                         /* 0 if NCCL cannot be loaded */

/* Counters a d_hits array must hold for n_pairs pairs: world * ceil(n_pairs / world) (the in-place all-gather layout). */
uint64_t satmc_group_hits_capacity(const satmc_group* g, uint64_t n_pairs);

/* satmc_count_fused over the group, inputs resident: d_pairs[l] = the FULL pair array on local device l, d_hits[l] =
 * satmc_group_hits_capacity() counters there.  Every rank launches its shard and joins the collective on its stream;
 * afterwards every d_hits[l][0..n_pairs) holds the counts of all pairs -- identical to satmc_count_fused on one GPU.
 * Asynchronous (satmc_group_synchronize, or stream order on the streams given to satmc_group_create_rank). */
int satmc_group_count_fused(satmc_group* g, const satmc_pair* const* d_pairs, uint64_t n_pairs, uint64_t n_samples,
                            uint64_t seed, uint64_t sample_offset, uint32_t pair_id_offset, int shard_mode,
                            uint64_t* const* d_hits, uint32_t flags);
/* Same from host memory: every rank passes the same h_pairs, h_hits receives all n_pairs counts.  Blocking. */
int satmc_group_count_fused_host(satmc_group* g, const satmc_pair* h_pairs, uint64_t n_pairs, uint64_t n_samples,
                                 uint64_t seed, uint64_t sample_offset, uint32_t pair_id_offset, int shard_mode,
                                 uint64_t* h_hits, uint32_t flags);
/* How the last satmc_group_count_fused_host call combined the shards: nothing to combine (one device, or disjoint
 * slices read back directly), an NCCL collective, or -- sample ranges inside one process with few pairs -- no collective
 * at all: every device's counting kernel finishes into one counter array in device 0's memory with system-scope
 * atomics over NVLink (the reduction is the compute kernel's own epilogue).  Chosen automatically from 6 devices on
 * (measured per small call: 121 us against 147 us through NCCL at 8 GPUs, 89 against 80 at 4);
 * satmc_group_set_peer_reduce(g, 1 / 0 / -1) forces it on / off / back to automatic.  Same counts either way. */
#define SATMC_EXCHANGE_NONE          0
#define SATMC_EXCHANGE_NCCL          1
#define SATMC_EXCHANGE_PEER_ATOMICS  2
int satmc_group_last_exchange(const satmc_group* g);
int satmc_group_set_peer_reduce(satmc_group* g, int enabled);

/* With timing enabled satmc_group_count_fused brackets local device 0's kernel and the collective with CUDA events
 * (and waits for them): the split of a step into compute and exchange that bench.py reports (allreduce_us). */
int satmc_group_set_timing(satmc_group* g, int enabled);
int satmc_group_last_times(const satmc_group* g, float* kernel_ms, float* collective_ms);

/* The adaptive z-test loop (satmc_adaptive_run) over host rows dealt round-robin to the ranks.  The tables are made
 * resident on every local device once (they are ~0.5 GB at the reference's defaults); row i always draws Philox stream
 * stream_id_offset + i, so h_cp_out does not depend on the number of GPUs.
 *   replaces: the per-file body of compute_collision_probability.cu:259-360 and ztest.cu:262-406 on n GPUs */
int satmc_group_set_tables(satmc_group* g, const float* h_robot_base, const float* h_poses, uint32_t n_poses,
                           const float* h_std_devs, uint32_t n_std, const float* h_accuracy_bins,
                           const float* h_bin_accuracy, int n_accuracy_bins);
int satmc_group_adaptive_run_host(satmc_group* g, const float* h_pose_idxs, const float* h_std_dev_idxs,
                                  const float* h_positions, int n_rows, int max_samples, int n_batch_small, int switch_at,
                                  int n_batch_large, uint64_t seed, uint32_t stream_id_offset, float* h_cp_out,
                                  int* iterations_out, long long* samples_drawn_out);

/* ---- host-buffer convenience (H2D + kernel + D2H inside the call) --------------------------- */

int satmc_count_fused_host(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs,
                           uint64_t n_samples, uint64_t seed, uint64_t sample_offset,
                           uint32_t pair_id_offset, uint64_t* h_hits, uint32_t flags);
int satmc_count_streamed_host(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs,
                              const float* h_z, uint64_t ldz, uint64_t z_pair_stride, int ndof,
                              uint64_t n_samples, uint64_t* h_hits, uint32_t flags);
/* probabilities = hits / n_samples as float32 (the `cp` column of the dataset rows,
 * generate_dataset.cu:485-494) */
int satmc_collision_probability_host(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs,
                                     uint64_t n_samples, uint64_t seed, float* h_cp);

/* Pinned host memory for the *_host paths (optional; pageable memory works, slower). */
int satmc_host_alloc(void** out, size_t bytes);
int satmc_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SATMC_H */
