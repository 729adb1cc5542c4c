// satmc.cu -- host side and C ABI of the B200-native Monte Carlo SAT path (include/satmc.h).
// The kernels live in satmc_kernels.cuh; geometry, polygon SAT and the sampler in satmc_geom.cuh,
// satmc_poly.cuh and satmc_sampler.cuh.  One translation unit, built by ../Makefile into libsatmc.so.
#include "../../include/satmc.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <new>

#include "satmc_kernels.cuh"
#include "satmc_internal.hpp"

// =============================================================================================
// host side / C ABI
// =============================================================================================
using namespace satmc;

char* satmc_thread_error()
{
    static thread_local char e[512] = {0};
    return e;
}

static inline size_t tma_smem_bytes(int ndof, bool shared_bank) { return (size_t)kWarps * tma_stages(ndof) * ndof * tma_tile(ndof, shared_bank) * sizeof(float); }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tensor_map_encode_fn get_tensor_map_encode()
{
    // initialised once, thread-safely (contexts are driven by one host thread each and may make their first
    // streamed call at the same time)
    static const tensor_map_encode_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        const bool ok = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
                        q == cudaDriverEntryPointSuccess;
        cudaGetLastError();
        return ok ? reinterpret_cast<tensor_map_encode_fn>(f) : nullptr;
    }();
    return fn;
}

// 2-D tensor [ndof][ldz] of float32 over the sample bank, box = [ndof][kTile]
static bool make_z_tensor_map(CUtensorMap* map, const float* z, uint64_t ldz, int ndof, int tile)
{
    tensor_map_encode_fn enc = get_tensor_map_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ldz, (cuuint64_t)ndof};
    const cuuint64_t strides[1] = {(cuuint64_t)ldz * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)tile, (cuuint32_t)ndof};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(z), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int scratch(satmc_ctx* ctx, int slot, size_t bytes, void** out)
{
    if (bytes > ctx->scratch_cap[slot]) {
        if (ctx->d_scratch[slot]) { CU(ctx, cudaFree(ctx->d_scratch[slot])); ctx->d_scratch[slot] = nullptr; ctx->scratch_cap[slot] = 0; }
        size_t cap = bytes + bytes / 4 + 256;
        if (cudaMalloc(&ctx->d_scratch[slot], cap) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, SATMC_ERR_NOMEM, "cudaMalloc of %zu bytes failed", cap);
        }
        ctx->scratch_cap[slot] = cap;
    }
    *out = ctx->d_scratch[slot];
    return SATMC_OK;
}

extern "C" {

#ifdef SATMC_DEBUG
const char* satmc_version(void) { return "satmc-b200 0.2 (sm_100a) DEBUG: device-side bounds asserts"; }
#else
const char* satmc_version(void) { return "satmc-b200 0.2 (sm_100a)"; }
#endif

const char* satmc_last_error(const satmc_ctx* ctx) { return ctx ? ctx->err : satmc_thread_error(); }

int satmc_create(int device, void* stream, satmc_ctx** out)
{
    if (!out) return fail(nullptr, SATMC_ERR_INVALID, "satmc_create: out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, SATMC_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    }
    if (device < 0 || device >= n) return fail(nullptr, SATMC_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, SATMC_ERR_CUDA, "cudaGetDeviceProperties failed");
    }
    if (prop.major != 10)
        return fail(nullptr, SATMC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    satmc_ctx* ctx = new (std::nothrow) satmc_ctx();
    if (!ctx) return fail(nullptr, SATMC_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->stream = (cudaStream_t)stream;
    ctx->sm_count = prop.multiProcessorCount;
    DeviceGuard g(device);
    int bps = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_count<DirectSrc, false>, kThreads, 0);
    if (e != cudaSuccess || bps < 1) bps = SATMC_MIN_BLOCKS_FUSED;
    ctx->blocks_per_sm = bps;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_count<DirectSrc, true>, kThreads, 0);
    if (e != cudaSuccess || bps < 1) bps = SATMC_MIN_BLOCKS_STREAMED;
    ctx->blocks_per_sm_streamed = bps;
    // dynamic + static shared memory beyond 48 KB is opt-in (deeper TMA rings, larger tiles)
    auto tma_setup = [&](auto kernel, int ndof, bool shared_bank, int& out) {
        const size_t bytes = tma_smem_bytes(ndof, shared_bank);
        if (bytes > 40 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        int n = 0;
        const cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, bytes);
        out = (err == cudaSuccess && n >= 1) ? n : 1;
    };
    tma_setup(k_count_streamed_tma<3, tma_tile(3, false)>, 3, false, ctx->blocks_per_sm_tma[0][0]);
    tma_setup(k_count_streamed_tma<3, tma_tile(3, true)>, 3, true, ctx->blocks_per_sm_tma[0][1]);
    tma_setup(k_count_streamed_tma<5, tma_tile(5, false)>, 5, false, ctx->blocks_per_sm_tma[1][0]);
    tma_setup(k_count_streamed_tma<5, tma_tile(5, true)>, 5, true, ctx->blocks_per_sm_tma[1][1]);
    cudaGetLastError();
    if (cudaMalloc(&ctx->d_exact_evals, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_exact_evals, 0, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_ticket, 3 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_ticket, 0, 3 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_word, 64) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        int rc = fail(nullptr, SATMC_ERR_CUDA, "context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return rc;
    }
    ctx->d_blocks_done = reinterpret_cast<unsigned*>(ctx->d_ticket + 2);
    // development knobs (multiples of 128 samples)
    if (const char* e = getenv("SATMC_MIN_CHUNK")) { const long v = atol(e); if (v >= 128) ctx->tune_min_chunk = (uint64_t)v / 128 * 128; }
    if (const char* e = getenv("SATMC_STREAM_CHUNK3")) { const long v = atol(e); if (v >= 256) ctx->tune_stream_chunk[0] = (uint64_t)v / 256 * 256; }
    if (const char* e = getenv("SATMC_STREAM_CHUNK5")) { const long v = atol(e); if (v >= 256) ctx->tune_stream_chunk[1] = (uint64_t)v / 256 * 256; }
    if (const char* e = getenv("SATMC_FUSED_MAX_CHUNK")) { const long long v = atoll(e); if (v >= 32768 && v <= (1ll << 31)) ctx->tune_fused_max_chunk = (uint64_t)v / 128 * 128; }
    if (const char* e = getenv("SATMC_TINY_BPS")) { const long v = atol(e); if (v >= 0 && v <= 8) ctx->tune_tiny_bps = (int)v; }
    if (const char* e = getenv("SATMC_STREAM_IPW")) { const long v = atol(e); if (v >= 1 && v <= 4096) ctx->tune_stream_ipw = (int)v; }
    if (const char* e = getenv("SATMC_TINY_CHUNK")) { const long v = atol(e); if (v >= 128) ctx->tune_tiny_chunk = (uint64_t)v / 128 * 128; }
    *out = ctx;
    return SATMC_OK;
}

int satmc_destroy(satmc_ctx* ctx)
{
    if (!ctx) return SATMC_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 3; i++) if (ctx->d_scratch[i]) cudaFree(ctx->d_scratch[i]);
    for (int i = 0; i < 2; i++) if (ctx->d_acc[i]) cudaFree(ctx->d_acc[i]);
    if (ctx->d_exact_evals) cudaFree(ctx->d_exact_evals);
    if (ctx->d_ticket) cudaFree(ctx->d_ticket);
    if (ctx->h_word) cudaFreeHost(ctx->h_word);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    delete ctx;
    return SATMC_OK;
}

int satmc_synchronize(satmc_ctx* ctx)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    DeviceGuard g(ctx->device);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SATMC_OK;
}

uint64_t satmc_launch_count(const satmc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int satmc_set_profiling(satmc_ctx* ctx, int enabled)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    ctx->profiling = enabled != 0;
    return SATMC_OK;
}

float satmc_last_kernel_ms(const satmc_ctx* ctx)
{
    if (!ctx || !ctx->last_ms_valid) return -1.0f;
    satmc_ctx* c = const_cast<satmc_ctx*>(ctx);
    DeviceGuard g(c->device);
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int satmc_plan_debug(satmc_ctx* ctx, int kind, uint64_t n_pairs, uint64_t n_samples, uint64_t* chunk_out, uint64_t* n_chunks_out);

int satmc_exact_evals(satmc_ctx* ctx, uint64_t* out, int reset)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    DeviceGuard g(ctx->device);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long v = 0;
    CU(ctx, cudaMemcpy(&v, ctx->d_exact_evals, sizeof(v), cudaMemcpyDeviceToHost));
    if (out) *out = (uint64_t)v;
    if (reset) CU(ctx, cudaMemset(ctx->d_exact_evals, 0, sizeof(v)));
    return SATMC_OK;
}

}  // extern "C"

// Chooses the chunking of the sample range (work item = (pair, chunk), one warp per item) and the grid size.
// items_per_warp: how many items a resident warp should get at least (when the sample range allows); the static
// grid-stride assignment loses up to 1/items_per_warp of the run to the last, partly filled round.
// max_chunk: upper bound on the samples of one item.  An item's hits are summed in 32 bits (per-lane counters, the
// REDUX warp total, the sweep's per-setting shared-memory counters), so no item may exceed 2^31 samples.
constexpr uint64_t kMaxChunk = 1ull << 31;
static int plan_items(satmc_ctx* ctx, CountParams& p, int bps, uint64_t& blocks, int items_per_warp = 8,
                      uint64_t max_chunk = kMaxChunk, int warps_per_block = kWarps, uint64_t granule = 128,
                      bool block_sums = true)
{
    const uint64_t kWarps = (uint64_t)warps_per_block;                // (shadows the default block shape)
    const uint64_t resident_warps = (uint64_t)ctx->sm_count * bps * kWarps;
    const uint64_t target_items = resident_warps * (uint64_t)items_per_warp;
    const uint64_t min_chunk = ctx->tune_min_chunk;               // 64 samples per lane: amortises the pair prologue
    const uint64_t tiny_chunk = ctx->tune_tiny_chunk;             // small problems: parallelism matters more than the prologue
    uint64_t n_chunks = 1;
    if (p.n_pairs < target_items) {
        n_chunks = (target_items + p.n_pairs - 1) / p.n_pairs;
        const uint64_t max_chunks = (p.n_samples + min_chunk - 1) / min_chunk;
        if (n_chunks > max_chunks) {
            n_chunks = max_chunks;
            // the shortest items allowed do not make 8 per warp: at least fill whole rounds of the resident warps
            // (one pair x 8e6: 3 677 items on 2 368 warps took two rounds, 53 us; 2 315 items take one)
            const uint64_t rounds = p.n_pairs * n_chunks / resident_warps;
            if (rounds >= 1 && rounds * resident_warps / p.n_pairs >= 1) n_chunks = rounds * resident_warps / p.n_pairs;
        }
        if (n_chunks < 1) n_chunks = 1;
        if (p.n_pairs * n_chunks < resident_warps) {              // not even one item per resident warp: cut finer
            // A launch this small is over when the busiest SM is: cutting for the resident number of blocks per SM and
            // then rounding the items to whole sample groups can leave some SMs with 2 blocks and others with 1 (cfg 2:
            // 245 blocks on 148 SMs).  Candidates: cut for b = bps ... 1 blocks per SM; cost of a candidate = blocks on
            // the busiest SM x (pair prologue + sample groups per item), with a penalty for running below the resident
            // occupancy (fewer warps to hide latency behind).  Measured on one pair x 1e6: 20.5 -> 18.1 us by events;
            // 2e6 and 4e6 samples keep the resident cut (22.2 / 28.9 us against 23.8 / 33.0 when cut for one block).
            const uint64_t max_tiny = (p.n_samples + tiny_chunk - 1) / tiny_chunk;
            uint64_t best = n_chunks;
            double best_cost = 0.0;
            for (int b = bps; b >= 1; b--) {
                if (ctx->tune_tiny_bps > 0 && b != (ctx->tune_tiny_bps < bps ? ctx->tune_tiny_bps : bps)) continue;
                const uint64_t fill_warps = (uint64_t)ctx->sm_count * (uint64_t)b * kWarps;
                uint64_t nc = (fill_warps + p.n_pairs - 1) / p.n_pairs;
                if (nc > max_tiny) nc = max_tiny;
                if (nc < n_chunks) nc = n_chunks;
                if (block_sums && nc >= kWarps) nc = (nc / kWarps) * kWarps;
                uint64_t ch = (p.n_samples + nc - 1) / nc;
                ch = ((ch + granule - 1) / granule) * granule;
                const uint64_t items = p.n_pairs * ((p.n_samples + ch - 1) / ch);
                const uint64_t blk = (items + kWarps - 1) / kWarps;
                const uint64_t on_busiest = (blk + (uint64_t)ctx->sm_count - 1) / (uint64_t)ctx->sm_count;
                const double cost = (double)on_busiest * (1.43 + (double)ch / 128.0) * (on_busiest < (uint64_t)bps ? 1.15 : 1.0);
                if (best_cost == 0.0 || cost < best_cost) { best_cost = cost; best = nc; }
            }
            n_chunks = best;
        }
    }
    if (max_chunk > kMaxChunk) max_chunk = kMaxChunk;
    const uint64_t per_chunk = (p.n_samples + n_chunks - 1) / n_chunks;
    if (per_chunk > max_chunk) n_chunks *= (per_chunk + max_chunk - 1) / max_chunk;   // a multiple: the rounds stay full
    if (block_sums && n_chunks >= (uint64_t)kWarps) n_chunks = (n_chunks / kWarps) * kWarps;    // block-uniform pairs
    uint64_t chunk = (p.n_samples + n_chunks - 1) / n_chunks;
    chunk = ((chunk + granule - 1) / granule) * granule;          // whole sample groups / whole bulk-tensor tiles
    if (chunk > kMaxChunk) chunk = kMaxChunk;                         // (a multiple of 128 and of 256)
    n_chunks = (p.n_samples + chunk - 1) / chunk;
    if (n_chunks > 0xffffffffull) return fail(ctx, SATMC_ERR_INVALID, "too many chunks");
    p.chunk = chunk;
    p.n_chunks = (uint32_t)n_chunks;
    p.n_items = p.n_pairs * n_chunks;
    p.block_uniform = (block_sums && n_chunks % kWarps == 0) ? 1u : 0u;
    blocks = (p.n_items + kWarps - 1) / kWarps;
    const uint64_t max_blocks = (uint64_t)ctx->sm_count * bps;
    if (blocks > max_blocks) blocks = max_blocks;
    return SATMC_OK;
}

// Diagnostics: the chunking the planner picks for a call of the given kind (0 fused rectangles, 1 streamed, 2 polygons,
// 3 sweep) -- tests assert that no item exceeds 2^31 samples (32-bit per-item hit sums) without running 2^32-sample items.
extern "C" int satmc_plan_debug(satmc_ctx* ctx, int kind, uint64_t n_pairs, uint64_t n_samples, uint64_t* chunk_out, uint64_t* n_chunks_out)
{
    if (!ctx || !chunk_out || !n_chunks_out) return fail(ctx, SATMC_ERR_INVALID, "null argument");
    CountParams p{}; p.n_pairs = n_pairs; p.n_samples = n_samples;
    uint64_t blocks = 0;
    int rc;
    if (kind == 0) rc = plan_items(ctx, p, ctx->blocks_per_sm, blocks, 8, ctx->tune_fused_max_chunk);
    else if (kind == 1) rc = plan_items(ctx, p, ctx->blocks_per_sm_streamed, blocks);
    else if (kind == 2) rc = plan_items(ctx, p, 2, blocks);
    else rc = plan_items(ctx, p, SATMC_SWEEP_BPS, blocks, 32, kMaxChunk, kSweepWarps);
    if (rc) return rc;
    *chunk_out = p.chunk; *n_chunks_out = p.n_chunks;
    return SATMC_OK;
}

// Switches a launch to dynamic work distribution (see CountParams::ticket) when warps get more than one item and need
// not walk the items in step.  Every processed item draws one ticket, so the counter advances by n_items per launch.
static void use_tickets(satmc_ctx* ctx, CountParams& p, uint64_t blocks)
{
    if (p.block_uniform || p.n_items <= blocks * (uint64_t)kWarps) return;
    // ticket_base is a launch parameter: a captured launch replayed from a CUDA graph would see a stale one.  While the
    // stream is capturing, keep the static order (same counts, a few per cent slower).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
    p.ticket = ctx->d_ticket + ctx->ticket_sel;
    p.ticket_base = ctx->ticket_next[ctx->ticket_sel];
    ctx->ticket_next[ctx->ticket_sel] += p.n_items;
}

// With several work items per counter the kernels accumulate with atomics.  Instead of clearing the caller's counters
// first (a second launch: 2-3 us for a memset node plus the gap to the kernel, as much as a whole cfg 2 call), the
// atomics go to a scratch array of the context that holds zeros between launches; the last block to finish moves the
// totals to the caller's array (or adds them, SATMC_ACCUMULATE) and zeroes the scratch again (finalize_counters).
// `span` = counters behind p.hits, of which this launch owns `counters`, laid out as fin_inner consecutive ones every
// fin_stride.
// `arrivals` = contributions every counter receives when that number is the same for all of them (0: irregular, e.g.
// the deferred queue): the packed scheme of packed_arrive then finishes each counter with its last contribution and
// the kernel needs no tail; it is kept only if the counts fit the packing (else p.arrivals stays 0: tail pass).
static int prepare_counters(satmc_ctx* ctx, CountParams& p, uint64_t span, uint64_t counters, uint32_t arrivals, uint64_t fin_inner = 1,
                            uint64_t fin_stride = 1, uint64_t acc_offset = 0)
{
    p.hits_len = span - acc_offset;
    p.acc = nullptr; p.blocks_done = nullptr; p.arrivals = 0;
    if (p.n_chunks <= 1) return SATMC_OK;
    if (arrivals < (1u << 24) && p.chunk * (uint64_t)p.n_chunks < (1ull << kPackShift)) p.arrivals = arrivals;
    const int sel = ctx->ticket_sel;
    if (span > ctx->acc_cap[sel]) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(ctx->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            return fail(ctx, SATMC_ERR_INVALID, "the accumulator scratch must grow, which cannot be captured: make one call of "
                                                "this size before capturing");
        }
        // launches that still use the old array must finish before it is freed
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->d_acc[sel]) { CU(ctx, cudaFree(ctx->d_acc[sel])); ctx->d_acc[sel] = nullptr; ctx->acc_cap[sel] = 0; }
        const size_t cap_n = (size_t)(span + span / 4 + 64);
        if (cudaMalloc(&ctx->d_acc[sel], cap_n * sizeof(unsigned long long)) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, SATMC_ERR_NOMEM, "cudaMalloc of %zu bytes failed", cap_n * sizeof(unsigned long long));
        }
        CU(ctx, cudaMemset(ctx->d_acc[sel], 0, cap_n * sizeof(unsigned long long)));
        ctx->acc_cap[sel] = cap_n;
    }
    p.acc = ctx->d_acc[sel] + acc_offset;
    p.blocks_done = ctx->d_blocks_done + sel;
    p.n_counters = counters; p.fin_inner = fin_inner; p.fin_stride = fin_stride;
    return SATMC_OK;
}

// Launches the rectangle counting kernel.
template <class Src, bool STREAMED>
static int launch_count(satmc_ctx* ctx, const Src& src, CountParams p, bool time_it)
{
    if (p.n_pairs == 0 || p.n_samples == 0) {
        if (p.n_pairs && !(p.flags & SATMC_ACCUMULATE))
            CU(ctx, cudaMemsetAsync(p.hits, 0, p.n_pairs * sizeof(unsigned long long), ctx->stream));
        return SATMC_OK;
    }
    // bulk-tensor path: aligned bank, coordinates that fit the TMA's 32-bit signed indices, a driver that encodes the map
    CUtensorMap zmap;
    const bool shared_bank = STREAMED && p.z_pair_stride == 0;        // every pair reads the same samples (common random numbers)
    const int tile = tma_tile((int)p.ndof, shared_bank);
    bool tma = STREAMED && p.vec_ok && p.ldz < (1ull << 31) && p.ldz >= (uint64_t)tile;
    if (tma) tma = make_z_tensor_map(&zmap, p.z, p.ldz, p.ndof, tile);
    const int bps = tma ? ctx->blocks_per_sm_tma[p.ndof == 5][shared_bank] : (STREAMED ? ctx->blocks_per_sm_streamed : ctx->blocks_per_sm);
    uint64_t blocks = 0;
    // Long items (few pairs, many samples each -- cfg 4, cfg 5): undecided sample groups are queued per warp and worked
    // off 32 at a time (ColdQueue, k_count<.., DEFER = true>).  Items are capped at 2^18 samples: the queue then holds an
    // item's undecided groups at the rates seen in practice (<= 2.4e-4 per test), and a warp gets enough items for the
    // dynamic distribution to even out the speed differences between SMs (a cfg 4 share of 1.25e10 samples: 8 items of
    // 5 ms per warp at a cap of 2^20 took 39.8 ms, 24 items per warp take 38.4).  Short items (< 32768 samples) keep the
    // immediate path: there the queue's bookkeeping costs more than the few undecided groups of an item.
    // Bulk-tensor kernel on private banks (HBM bound): short items keep the addresses the resident warps stream from
    // close together; the ring runs across items, so a switch costs only the pair prologue.  Measured on five bank
    // shapes (profiles/r2_streamed_ring_experiments.log): 3-DoF 6.0-6.3 -> 6.6-6.8 TB/s with items of 2048 samples,
    // 5-DoF 6.6-6.9 -> 6.8-6.95 TB/s with 8192.  A shared (L2-resident) bank is issue bound: items stay long.
    uint64_t max_chunk = STREAMED ? (1ull << 36) : ctx->tune_fused_max_chunk;
    if (tma && !shared_bank) max_chunk = ctx->tune_stream_chunk[p.ndof == 5];
    int rc = plan_items(ctx, p, bps, blocks, tma ? ctx->tune_stream_ipw : 8, max_chunk, kWarps, tma ? (uint64_t)tile : 128, !tma);
    if (rc) return rc;
    const bool defer = !STREAMED && p.chunk >= 32768;
    // (the bulk-tensor kernel's warps never meet at a block barrier -- block_sums = false above: one packed atomic per
    // item is cheap, a barrier per item keeps eight private rings in lockstep)
    // without the deferred queue every counter receives a known number of contributions (packed_arrive)
    rc = prepare_counters(ctx, p, p.n_pairs, p.n_pairs, defer ? 0u : (p.block_uniform ? p.n_chunks / kWarps : p.n_chunks));
    if (rc) return rc;
    if (p.acc && !defer && !tma && p.arrivals == 0u)                  // (k_count's packed scheme is a template parameter)
        return fail(ctx, SATMC_ERR_INVALID, "sample range too long for one call (%llu samples per pair)", (unsigned long long)p.n_samples);
    const uint64_t ticket_before = ctx->ticket_next[ctx->ticket_sel];
    // bulk-copy staged kernel: dynamic order only when all pairs read one shared bank (L2 resident, issue bound: +5 %);
    // with private banks the kernel is HBM bound and the static order streams DRAM slightly better (measured -1..2 %)
    if (!tma || p.z_pair_stride == 0) use_tickets(ctx, p, blocks);
    if (time_it && !ctx->events_by_caller) CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if constexpr (STREAMED) {
        if (tma && p.ndof == 5) {
            if (shared_bank) k_count_streamed_tma<5, tma_tile(5, true)><<<(unsigned)blocks, kThreads, tma_smem_bytes(5, true), ctx->stream>>>(src, p, zmap);
            else k_count_streamed_tma<5, tma_tile(5, false)><<<(unsigned)blocks, kThreads, tma_smem_bytes(5, false), ctx->stream>>>(src, p, zmap);
        }
        else if (tma)
            k_count_streamed_tma<3, tma_tile(3, false)><<<(unsigned)blocks, kThreads, tma_smem_bytes(3, shared_bank), ctx->stream>>>(src, p, zmap);
        else
            if (p.acc) k_count<Src, true, false, true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(src, p);
            else k_count<Src, true, false, false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(src, p);
    } else {
        const bool multi = p.acc != nullptr;                         // several work items per counter (prepare_counters)
        if (defer && multi) k_count<Src, false, true, true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(src, p);
        else if (defer) k_count<Src, false, true, false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(src, p);
        else if (multi) k_count<Src, false, false, true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(src, p);
        else k_count<Src, false, false, false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(src, p);
    }
    if (cudaPeekAtLastError() != cudaSuccess) ctx->ticket_next[ctx->ticket_sel] = ticket_before;       // nothing ran: no ticket was drawn
    CU(ctx, cudaGetLastError());
    if (time_it && !ctx->events_by_caller) { CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream)); ctx->last_ms_valid = true; }
    ctx->launches++;
    return SATMC_OK;
}

static int check_streamed_args(satmc_ctx* ctx, const void* pairs, const void* z, uint64_t ldz, int ndof, uint64_t n_samples,
                               uint64_t n_pairs, uint64_t z_pair_stride, const void* out)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!pairs || !out) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (ndof != 3 && ndof != 5) return fail(ctx, SATMC_ERR_INVALID, "ndof must be 3 or 5, got %d", ndof);
    if (n_samples && !z) return fail(ctx, SATMC_ERR_INVALID, "d_z is NULL");
    if (n_pairs && n_samples && (n_pairs - 1) * z_pair_stride + n_samples > ldz)
        return fail(ctx, SATMC_ERR_INVALID, "sample planes too short: need %llu samples per plane, ldz = %llu",
                    (unsigned long long)((n_pairs - 1) * z_pair_stride + n_samples), (unsigned long long)ldz);
    return SATMC_OK;
}

extern "C" {

}  // extern "C"

// satmc_count_fused with the internal flags allowed (SATMC_PEER_ATOMIC_OUT, used by the group code)
int satmc_count_fused_impl(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                           uint64_t sample_offset, uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if ((!d_pairs || !d_hits) && n_pairs) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_pairs > 0xffffffffull - pair_id_offset) return fail(ctx, SATMC_ERR_INVALID, "pair ids exceed 32 bits");
    if (((uintptr_t)d_pairs & 15u) != 0) return fail(ctx, SATMC_ERR_INVALID, "d_pairs must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    CountParams p{};
    p.n_pairs = n_pairs; p.n_samples = n_samples; p.sample_offset = sample_offset; p.pair_id_offset = pair_id_offset;
    philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32), p.keys); p.flags = flags;
    p.hits = reinterpret_cast<unsigned long long*>(d_hits); p.exact_evals = ctx->d_exact_evals;
    return launch_count<DirectSrc, false>(ctx, DirectSrc{d_pairs}, p, ctx->profiling);
}

// would a fused call of this shape be cut into several work items per pair (the kernels with the accumulate / finalize
// epilogue, which is where SATMC_PEER_ATOMIC_OUT is honoured)?
bool satmc_fused_is_multi(satmc_ctx* ctx, uint64_t n_pairs, uint64_t n_samples)
{
    CountParams p{}; p.n_pairs = n_pairs; p.n_samples = n_samples;
    uint64_t blocks = 0;
    if (n_pairs == 0 || n_samples == 0 || plan_items(ctx, p, ctx->blocks_per_sm, blocks, 8, ctx->tune_fused_max_chunk)) return false;
    return p.n_chunks > 1;
}

extern "C" {

int satmc_count_fused(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                      uint64_t sample_offset, uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags)
{
    return satmc_count_fused_impl(ctx, d_pairs, n_pairs, n_samples, seed, sample_offset, pair_id_offset, d_hits,
                                  flags & (SATMC_ACCUMULATE | SATMC_EXACT_ONLY));
}

int satmc_count_streamed(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs, const float* d_z, uint64_t ldz,
                         uint64_t z_pair_stride, int ndof, uint64_t n_samples, uint64_t* d_hits, uint32_t flags)
{
    int rc = check_streamed_args(ctx, d_pairs, d_z, ldz, ndof, n_samples, n_pairs, z_pair_stride, d_hits);
    if (rc) return rc;
    if (((uintptr_t)d_pairs & 15u) != 0) return fail(ctx, SATMC_ERR_INVALID, "d_pairs must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    CountParams p{};
    p.n_pairs = n_pairs; p.n_samples = n_samples; p.flags = flags;
    p.hits = reinterpret_cast<unsigned long long*>(d_hits); p.exact_evals = ctx->d_exact_evals;
    p.z = d_z; p.ldz = ldz; p.z_pair_stride = z_pair_stride; p.ndof = ndof;
    p.vec_ok = (((uintptr_t)d_z & 15u) == 0 && ldz % 4 == 0 && z_pair_stride % 4 == 0) ? 1 : 0;
    return launch_count<DirectSrc, true>(ctx, DirectSrc{d_pairs}, p, ctx->profiling);
}

int satmc_decide_streamed(satmc_ctx* ctx, const satmc_pair* d_pair, const float* d_z, uint64_t ldz, int ndof,
                          uint64_t n_samples, uint8_t* d_out, uint32_t flags)
{
    int rc = check_streamed_args(ctx, d_pair, d_z, ldz, ndof, n_samples, 1, 0, d_out);
    if (rc) return rc;
    if (n_samples == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    uint64_t blocks = (n_samples + 255) / 256;
    if (blocks > (uint64_t)ctx->sm_count * 8) blocks = (uint64_t)ctx->sm_count * 8;
    k_decide<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_pair, d_z, ldz, ndof, n_samples, d_out, flags, ctx->d_exact_evals);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_screen_debug(satmc_ctx* ctx, const satmc_pair* d_pair, const float* d_z, uint64_t ldz, int ndof, uint64_t n_samples,
                       float* d_m_out, float* d_eps_out)
{
    int rc = check_streamed_args(ctx, d_pair, d_z, ldz, ndof, n_samples, 1, 0, d_m_out);
    if (rc) return rc;
    if (!d_eps_out) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_samples == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    uint64_t blocks = (n_samples + 255) / 256;
    if (blocks > (uint64_t)ctx->sm_count * 8) blocks = (uint64_t)ctx->sm_count * 8;
    k_screen_debug<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_pair, d_z, ldz, ndof, n_samples, d_m_out, d_eps_out);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_fused_normals(satmc_ctx* ctx, uint64_t seed, uint32_t pair_id, uint64_t sample_offset, uint64_t n, int ndof,
                        float* d_z, uint64_t ldz)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (ndof != 3 && ndof != 5) return fail(ctx, SATMC_ERR_INVALID, "ndof must be 3 or 5, got %d", ndof);
    if (!d_z || ldz < n) return fail(ctx, SATMC_ERR_INVALID, "bad output plane (null or ldz < n)");
    if (n == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    uint64_t blocks = (n / 4 + 256) / 256;
    if (blocks > (uint64_t)ctx->sm_count * 8) blocks = (uint64_t)ctx->sm_count * 8;
    PhiloxKeys K; philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32), K);
    if (ndof == 3) k_fused_normals<3><<<(unsigned)blocks, 256, 0, ctx->stream>>>(K, pair_id, sample_offset, n, d_z, ldz);
    else           k_fused_normals<5><<<(unsigned)blocks, 256, 0, ctx->stream>>>(K, pair_id, sample_offset, n, d_z, ldz);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_philox_blocks(satmc_ctx* ctx, const uint32_t* d_ctr, uint64_t n, uint32_t key0, uint32_t key1, uint32_t* d_out)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!d_ctr || !d_out) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    PhiloxKeys K; philox_expand_key(key0, key1, K);
    k_philox<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_ctr, n, K, d_out);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_sat_corners(satmc_ctx* ctx, const float* d_r1, const float* d_r2, uint64_t n, uint8_t* d_out)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!d_r1 || !d_r2 || !d_out) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if ((((uintptr_t)d_r1 | (uintptr_t)d_r2) & 15u) != 0) return fail(ctx, SATMC_ERR_INVALID, "corner arrays must be 16-byte aligned");
    if (n == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    k_sat_corners<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_r1, d_r2, n, d_out);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_count_fused_polygons(satmc_ctx* ctx, const satmc_poly_pair* d_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                               uint64_t sample_offset, uint32_t pair_id_offset, uint64_t* d_hits, uint32_t flags)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if ((!d_pairs || !d_hits) && n_pairs) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_pairs > 0xffffffffull - pair_id_offset) return fail(ctx, SATMC_ERR_INVALID, "pair ids exceed 32 bits");
    DeviceGuard g(ctx->device);
    CountParams p{};
    p.n_pairs = n_pairs; p.n_samples = n_samples; p.sample_offset = sample_offset; p.pair_id_offset = pair_id_offset;
    philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32), p.keys); p.flags = flags & (SATMC_ACCUMULATE | SATMC_EXACT_ONLY); p.exact_evals = ctx->d_exact_evals;
    p.hits = reinterpret_cast<unsigned long long*>(d_hits);
    if (n_pairs == 0 || n_samples == 0) {
        if (n_pairs && !(flags & SATMC_ACCUMULATE)) CU(ctx, cudaMemsetAsync(d_hits, 0, n_pairs * sizeof(uint64_t), ctx->stream));
        return SATMC_OK;
    }
    uint64_t blocks = 0;
    int rc = plan_items(ctx, p, 2, blocks);
    if (rc) return rc;
    rc = prepare_counters(ctx, p, n_pairs, n_pairs, p.block_uniform ? p.n_chunks / kWarps : p.n_chunks);
    if (rc) return rc;
    const uint64_t ticket_before = ctx->ticket_next[ctx->ticket_sel];
    use_tickets(ctx, p, blocks);
    k_count_poly<false><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(reinterpret_cast<const float*>(d_pairs), p);
    if (cudaPeekAtLastError() != cudaSuccess) ctx->ticket_next[ctx->ticket_sel] = ticket_before;       // nothing ran: no ticket was drawn
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_count_streamed_polygons(satmc_ctx* ctx, const satmc_poly_pair* d_pairs, uint64_t n_pairs, const float* d_z, uint64_t ldz,
                                  uint64_t z_pair_stride, uint64_t n_samples, uint64_t* d_hits, uint32_t flags)
{
    int rc = check_streamed_args(ctx, d_pairs, d_z, ldz, 3, n_samples, n_pairs, z_pair_stride, d_hits);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    CountParams p{};
    p.n_pairs = n_pairs; p.n_samples = n_samples; p.flags = flags & (SATMC_ACCUMULATE | SATMC_EXACT_ONLY); p.exact_evals = ctx->d_exact_evals;
    p.hits = reinterpret_cast<unsigned long long*>(d_hits);
    p.z = d_z; p.ldz = ldz; p.z_pair_stride = z_pair_stride; p.ndof = 3;
    if (n_pairs == 0 || n_samples == 0) {
        if (n_pairs && !(flags & SATMC_ACCUMULATE)) CU(ctx, cudaMemsetAsync(d_hits, 0, n_pairs * sizeof(uint64_t), ctx->stream));
        return SATMC_OK;
    }
    uint64_t blocks = 0;
    rc = plan_items(ctx, p, 2, blocks);
    if (rc) return rc;
    rc = prepare_counters(ctx, p, n_pairs, n_pairs, p.block_uniform ? p.n_chunks / kWarps : p.n_chunks);
    if (rc) return rc;
    const uint64_t ticket_before = ctx->ticket_next[ctx->ticket_sel];
    use_tickets(ctx, p, blocks);
    k_count_poly<true><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>(reinterpret_cast<const float*>(d_pairs), p);
    if (cudaPeekAtLastError() != cudaSuccess) ctx->ticket_next[ctx->ticket_sel] = ticket_before;       // nothing ran: no ticket was drawn
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

int satmc_count_fused_sweep(satmc_ctx* ctx, const satmc_pair* d_pairs, uint64_t n_pairs, const float* h_sigmas, uint32_t n_cov,
                            uint64_t n_samples, uint64_t seed, uint64_t sample_offset, uint32_t pair_id_offset, uint64_t* d_hits,
                            uint32_t flags)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if ((!d_pairs || !d_hits || !h_sigmas) && n_pairs && n_cov) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_pairs > 0xffffffffull - pair_id_offset) return fail(ctx, SATMC_ERR_INVALID, "pair ids exceed 32 bits");
    if (((uintptr_t)d_pairs & 15u) != 0) return fail(ctx, SATMC_ERR_INVALID, "d_pairs must be 16-byte aligned");
    if (n_pairs == 0 || n_cov == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    if (n_samples == 0) {
        if (!(flags & SATMC_ACCUMULATE)) CU(ctx, cudaMemsetAsync(d_hits, 0, n_pairs * n_cov * sizeof(uint64_t), ctx->stream));
        return SATMC_OK;
    }
    CountParams p{};
    p.n_pairs = n_pairs; p.n_samples = n_samples; p.sample_offset = sample_offset; p.pair_id_offset = pair_id_offset;
    philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32), p.keys);
    p.flags = flags;
    p.hits = reinterpret_cast<unsigned long long*>(d_hits); p.exact_evals = ctx->d_exact_evals;
    uint64_t blocks = 0;
    int rc = plan_items(ctx, p, SATMC_SWEEP_BPS, blocks, 32, kMaxChunk, kSweepWarps);   // an item costs n_cov x a plain one: cut finer (+7 % on cfg5)
    if (rc) return rc;
    // settings are processed kSweepMax at a time; every slice sees the same normals.  Within a slice the settings are
    // sorted by sd_theta (bit pattern; any total order will do) so that equal values are neighbours: the kernel keeps
    // sine, cosine and the projected extents across them.
    for (uint32_t c0 = 0; c0 < n_cov; c0 += kSweepMax) {
        const int nc = (int)((n_cov - c0 < (uint32_t)kSweepMax) ? n_cov - c0 : kSweepMax);
        SweepSettings W{};
        W.n = nc;
        int order[kSweepMax];
        for (int i = 0; i < nc; i++) order[i] = i;
        const float* sig = h_sigmas + 3 * (size_t)c0;
        auto key = [&](int i) { uint32_t u; memcpy(&u, &sig[3 * i + 2], 4); return u; };
        std::stable_sort(order, order + nc, [&](int x, int y) { return key(x) < key(y); });
        for (int r = 0; r < nc; r++) {
            const int i = order[r];
            W.sx[r] = sig[3 * i]; W.sy[r] = sig[3 * i + 1]; W.st[r] = sig[3 * i + 2];
            W.orig[r] = (unsigned char)i;
        }
        CountParams q = p;
        q.hits = p.hits + c0;
        rc = prepare_counters(ctx, q, n_pairs * n_cov, n_pairs * (uint64_t)nc, q.n_chunks, (uint64_t)nc, (uint64_t)n_cov, c0);
        if (rc) return rc;
        k_count_sweep<SATMC_SWEEP_G><<<(unsigned)blocks, 32 * kSweepWarps, 0, ctx->stream>>>(d_pairs, W, (uint64_t)n_cov, q);
        CU(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return SATMC_OK;
}

}  // extern "C"

// One step of the reference kernel contract: count n_batch more samples for the first num_left slots, then the
// z-test tail.  With `ar` (adaptive run) the tail also retires finished pairs and compacts the live list in the same
// kernel; without it the tail is the reference kernel's own (done flags, running counts).
static int mc_step_impl(satmc_ctx* ctx, const float* d_robot_base, const float* d_poses, uint32_t n_poses,
                        const float* d_std_devs, uint32_t n_std, const float* d_pose_idxs, const float* d_std_dev_idxs,
                        const float* d_positions, float* d_cps, const float* d_accuracy_bins, const float* d_bin_accuracy,
                        int n_accuracy_bins, int* d_done, int n_samples, int n_batch, int num_left, uint64_t seed,
                        uint32_t stream_id_offset, uint32_t stream_id_stride, const int* d_live, AdaptiveRun* ar)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!d_robot_base || !d_poses || !d_std_devs || !d_pose_idxs || !d_std_dev_idxs || !d_positions || !d_cps ||
        !d_accuracy_bins || !d_bin_accuracy || !d_done)
        return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_batch < 0 || n_samples < n_batch || num_left < 0 || n_accuracy_bins < 2 || n_poses == 0 || n_std == 0)
        return fail(ctx, SATMC_ERR_INVALID, "bad sizes (n_batch %d, n_samples %d, num_left %d, bins %d)", n_batch, n_samples,
                    num_left, n_accuracy_bins);
    if (num_left == 0) return SATMC_OK;
    // Philox stream ids are 32 bits: ids of the slots must not wrap (two rows 2^32 apart would share a stream)
    const uint64_t span = ar ? (uint64_t)ar->n_pairs : (uint64_t)num_left;
    if ((span - 1) * (uint64_t)stream_id_stride > 0xffffffffull - stream_id_offset)
        return fail(ctx, SATMC_ERR_INVALID, "Philox stream ids exceed 32 bits (offset %u, %llu slots, stride %u)", stream_id_offset,
                    (unsigned long long)span, stream_id_stride);
    DeviceGuard g(ctx->device);
    void* d_hits = nullptr;
    int rc = scratch(ctx, 0, (size_t)num_left * sizeof(unsigned long long), &d_hits);
    if (rc) return rc;
    CountParams p{};
    p.n_pairs = (uint64_t)num_left; p.n_samples = (uint64_t)n_batch; p.sample_offset = (uint64_t)(n_samples - n_batch);
    p.pair_id_offset = stream_id_offset; philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32), p.keys); p.flags = 0;
    p.hits = reinterpret_cast<unsigned long long*>(d_hits); p.exact_evals = ctx->d_exact_evals;
    IndirectSrc src{d_robot_base, d_poses, d_std_devs, d_pose_idxs, d_std_dev_idxs, d_positions, n_poses, n_std, d_live, stream_id_stride};
    rc = launch_count<IndirectSrc, false>(ctx, src, p, ctx->profiling);
    if (rc) return rc;
    if (ar) {
        k_ztest_compact<<<(num_left + 255) / 256, 256, 0, ctx->stream>>>(
            reinterpret_cast<unsigned long long*>(d_hits), d_cps, d_accuracy_bins, d_bin_accuracy, n_accuracy_bins, n_samples, num_left,
            d_live, ar->d_cp_out, ar->d_n_samples_out, ar->d_live[ar->cur ^ 1], ar->d_n + (ar->iter & 1), ar->d_n + ((ar->iter + 1) & 1),
            ar->n_pairs);
    } else {
        k_ztest_tail<<<(num_left + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<unsigned long long*>(d_hits), d_cps,
                                                                       d_accuracy_bins, d_bin_accuracy, n_accuracy_bins, d_done,
                                                                       n_samples, num_left, d_live);
    }
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

// ---- the adaptive loop in steps, so that a group can drive several devices in lockstep from one host thread ----
int satmc_adaptive_begin(AdaptiveRun& ar)
{
    satmc_ctx* ctx = ar.ctx;
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!ar.d_cp_out) return fail(ctx, SATMC_ERR_INVALID, "d_cp_out is NULL");
    if (ar.n_pairs < 0 || ar.n_batch_small <= 0 || ar.n_batch_large <= 0 || ar.max_samples <= 0 || ar.stream_id_stride == 0)
        return fail(ctx, SATMC_ERR_INVALID, "bad schedule (n_pairs %d, n_batch %d/%d, max_samples %d)", ar.n_pairs, ar.n_batch_small,
                    ar.n_batch_large, ar.max_samples);
    ar.num_left = ar.n_pairs; ar.n_samples = 0; ar.iter = 0; ar.cur = 0; ar.drawn = 0;
    if (ar.n_pairs == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    // work buffers: counts (float, as the reference keeps them), two live lists, two counters
    const size_t n = (size_t)ar.n_pairs;
    void* base = nullptr;
    int rc = scratch(ctx, 1, n * (sizeof(float) + 2 * sizeof(int)) + 256, &base);
    if (rc) return rc;
    ar.d_counts = reinterpret_cast<float*>(base);
    ar.d_live[0] = reinterpret_cast<int*>(ar.d_counts + n);
    ar.d_live[1] = ar.d_live[0] + n;
    ar.d_n = ar.d_live[1] + n;
    CU(ctx, cudaMemsetAsync(ar.d_counts, 0, n * sizeof(float), ctx->stream));
    CU(ctx, cudaMemsetAsync(ar.d_n, 0, 2 * sizeof(int), ctx->stream));
    k_iota<<<(ar.n_pairs + 255) / 256, 256, 0, ctx->stream>>>(ar.d_live[0], ar.n_pairs);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

bool satmc_adaptive_pending(const AdaptiveRun& ar) { return ar.num_left > 0 && ar.n_samples < ar.max_samples; }

// one iteration, asynchronous; the number of pairs still unfinished arrives in the context's pinned word
int satmc_adaptive_enqueue(AdaptiveRun& ar)
{
    satmc_ctx* ctx = ar.ctx;
    const int n_batch = (ar.n_samples < ar.switch_at) ? ar.n_batch_small : ar.n_batch_large;    // generate_dataset.cu:427-430
    ar.n_samples += n_batch;
    int rc = mc_step_impl(ctx, ar.d_robot_base, ar.d_poses, ar.n_poses, ar.d_std_devs, ar.n_std, ar.d_pose_idxs, ar.d_std_dev_idxs,
                          ar.d_positions, ar.d_counts, ar.d_bins, ar.d_bin_acc, ar.n_bins, /*d_done (unused)*/ ar.d_live[0],
                          ar.n_samples, n_batch, ar.num_left, ar.seed, ar.stream_id_offset, ar.stream_id_stride, ar.d_live[ar.cur], &ar);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    ar.drawn += (long long)ar.num_left * n_batch;
    CU(ctx, cudaMemcpyAsync(ctx->h_word, ar.d_n + (ar.iter & 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return SATMC_OK;
}

// after the stream has been synchronised
void satmc_adaptive_collect(AdaptiveRun& ar)
{
    ar.num_left = *ar.ctx->h_word;
    ar.cur ^= 1;
    ar.iter++;
}

// pairs that hit max_samples: count / n_samples as they stand (ztest.cu:376-385)
int satmc_adaptive_finish(AdaptiveRun& ar)
{
    satmc_ctx* ctx = ar.ctx;
    if (ar.num_left <= 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    k_compact_live<<<(ar.num_left + 255) / 256, 256, 0, ctx->stream>>>(ar.d_live[ar.cur], ar.num_left, ar.d_counts, ar.n_samples,
                                                                          ar.d_cp_out, ar.d_n_samples_out, ar.n_pairs);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    ar.num_left = 0;
    return SATMC_OK;
}

extern "C" {

int satmc_mc_step(satmc_ctx* ctx, const float* d_robot_base, const float* d_poses, uint32_t n_poses, const float* d_std_devs,
                  uint32_t n_std, const float* d_pose_idxs, const float* d_std_dev_idxs, const float* d_positions,
                  float* d_cps, const float* d_accuracy_bins, const float* d_bin_accuracy, int n_accuracy_bins,
                  int* d_done, int iteration, int n_samples, int n_batch, int num_left, uint64_t seed,
                  uint32_t stream_id_offset)
{
    (void)iteration;
    return mc_step_impl(ctx, d_robot_base, d_poses, n_poses, d_std_devs, n_std, d_pose_idxs, d_std_dev_idxs, d_positions,
                        d_cps, d_accuracy_bins, d_bin_accuracy, n_accuracy_bins, d_done, n_samples, n_batch, num_left, seed,
                        stream_id_offset, 1u, nullptr, nullptr);
}

int satmc_adaptive_run(satmc_ctx* ctx, const float* d_robot_base, const float* d_poses, uint32_t n_poses,
                       const float* d_std_devs, uint32_t n_std, const float* d_pose_idxs, const float* d_std_dev_idxs,
                       const float* d_positions, int n_pairs, const float* d_accuracy_bins, const float* d_bin_accuracy,
                       int n_accuracy_bins, int max_samples, int n_batch_small, int switch_at, int n_batch_large,
                       uint64_t seed, uint32_t stream_id_offset, float* d_cp_out, int* d_n_samples_out,
                       int* iterations_out, long long* samples_drawn_out)
{
    if (iterations_out) *iterations_out = 0;
    if (samples_drawn_out) *samples_drawn_out = 0;
    AdaptiveRun ar{};
    ar.ctx = ctx; ar.d_robot_base = d_robot_base; ar.d_poses = d_poses; ar.n_poses = n_poses; ar.d_std_devs = d_std_devs; ar.n_std = n_std;
    ar.d_pose_idxs = d_pose_idxs; ar.d_std_dev_idxs = d_std_dev_idxs; ar.d_positions = d_positions; ar.n_pairs = n_pairs;
    ar.d_bins = d_accuracy_bins; ar.d_bin_acc = d_bin_accuracy; ar.n_bins = n_accuracy_bins; ar.max_samples = max_samples;
    ar.n_batch_small = n_batch_small; ar.switch_at = switch_at; ar.n_batch_large = n_batch_large; ar.seed = seed;
    ar.stream_id_offset = stream_id_offset; ar.stream_id_stride = 1; ar.d_cp_out = d_cp_out; ar.d_n_samples_out = d_n_samples_out;
    int rc = satmc_adaptive_begin(ar);
    if (rc) return rc;
    while (satmc_adaptive_pending(ar)) {                                    // ztest.cu:328, generate_dataset.cu:425
        rc = satmc_adaptive_enqueue(ar);
        if (rc) return rc;
        DeviceGuard g(ctx->device);
        CU(ctx, cudaStreamSynchronize(ctx->stream));                       // the loop condition needs the count
        satmc_adaptive_collect(ar);
    }
    rc = satmc_adaptive_finish(ar);
    if (rc) return rc;
    if (iterations_out) *iterations_out = ar.iter;
    if (samples_drawn_out) *samples_drawn_out = ar.drawn;
    return SATMC_OK;
}

int satmc_sample_positions(satmc_ctx* ctx, const float* d_poses, uint32_t n_poses, const float* d_std_devs, uint32_t n_std,
                           int n, float r_offset, float spread, uint64_t seed, uint32_t stream_id_offset,
                           float* d_positions, float* d_pose_idxs, float* d_std_dev_idxs)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!d_poses || !d_std_devs || !d_positions || !d_pose_idxs || !d_std_dev_idxs || n_poses == 0 || n_std == 0)
        return fail(ctx, SATMC_ERR_INVALID, "null pointer or empty table");
    if (n <= 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    PhiloxKeys K; philox_expand_key((uint32_t)seed, (uint32_t)(seed >> 32), K);
    k_sample_positions<<<(n + 255) / 256, 256, 0, ctx->stream>>>(K, d_poses, n_poses, d_std_devs, n_std, n, r_offset, spread,
                                                                 stream_id_offset, d_positions, d_pose_idxs, d_std_dev_idxs);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

// ---- device memory helpers so that host programs need no CUDA headers -------------------------------
int satmc_device_alloc(satmc_ctx* ctx, void** out, size_t bytes)
{
    if (!ctx || !out) return fail(ctx, SATMC_ERR_INVALID, "null argument");
    DeviceGuard g(ctx->device);
    if (cudaMalloc(out, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError(); *out = nullptr;
        return fail(ctx, SATMC_ERR_NOMEM, "cudaMalloc of %zu bytes failed", bytes);
    }
    return SATMC_OK;
}

int satmc_device_free(satmc_ctx* ctx, void* p)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (p) CU(ctx, cudaFree(p));
    return SATMC_OK;
}

int satmc_upload(satmc_ctx* ctx, void* d_dst, const void* h_src, size_t bytes)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (bytes == 0) return SATMC_OK;
    if (!d_dst || !h_src) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    DeviceGuard g(ctx->device);
    CU(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SATMC_OK;
}

int satmc_download(satmc_ctx* ctx, void* h_dst, const void* d_src, size_t bytes)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (bytes == 0) return SATMC_OK;
    if (!h_dst || !d_src) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    DeviceGuard g(ctx->device);
    CU(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SATMC_OK;
}

int satmc_download_async(satmc_ctx* ctx, void* h_dst, const void* d_src, size_t bytes)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (bytes == 0) return SATMC_OK;
    if (!h_dst || !d_src) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    DeviceGuard g(ctx->device);
    CU(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SATMC_OK;
}

int satmc_write_collision_probability(satmc_ctx* ctx, float* d_counts, int n_done, int n_samples)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if (!d_counts && n_done) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_done <= 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    k_write_cp<<<(n_done + 255) / 256, 256, 0, ctx->stream>>>(d_counts, n_done, n_samples);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    return SATMC_OK;
}

// ---- host-buffer variants ---------------------------------------------------------------------

// Large batches from host memory in two slices: a lead slice of one item per planned work item (18 944 pairs) on the
// context's stream, the rest on an auxiliary stream.  The rest's host->device copy runs under the lead kernel, the
// lead's counters go back under the second kernel, and the second kernel's blocks move in as the lead's retire, so
// only the lead's copy in and the rest's copy out remain exposed.
static int count_fused_host_pipelined(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs, uint64_t lead,
                                      uint64_t n_samples, uint64_t seed, uint64_t sample_offset, uint32_t pair_id_offset,
                                      uint64_t* h_hits, uint32_t flags, satmc_pair* d_pairs, uint64_t* d_hits)
{
    if (!ctx->aux) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    cudaStream_t A = ctx->stream, B = ctx->aux;
    const uint64_t rest = n_pairs - lead;
    CU(ctx, cudaEventRecord(ctx->ev_fork, A));                       // the call stays ordered after earlier work on A
    CU(ctx, cudaStreamWaitEvent(B, ctx->ev_fork, 0));
    CU(ctx, cudaMemcpyAsync(d_pairs, h_pairs, lead * sizeof(satmc_pair), cudaMemcpyHostToDevice, A));
    if (flags & SATMC_ACCUMULATE) CU(ctx, cudaMemcpyAsync(d_hits, h_hits, lead * sizeof(uint64_t), cudaMemcpyHostToDevice, A));
    CU(ctx, cudaMemcpyAsync(d_pairs + lead, h_pairs + lead, rest * sizeof(satmc_pair), cudaMemcpyHostToDevice, B));
    if (flags & SATMC_ACCUMULATE)
        CU(ctx, cudaMemcpyAsync(d_hits + lead, h_hits + lead, rest * sizeof(uint64_t), cudaMemcpyHostToDevice, B));
    ctx->events_by_caller = true;
    int rc = SATMC_OK;
    cudaError_t ce = cudaEventRecord(ctx->ev0, A);
    if (ce != cudaSuccess) rc = fail(ctx, SATMC_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(ce));
    if (rc == SATMC_OK) rc = satmc_count_fused(ctx, d_pairs, lead, n_samples, seed, sample_offset, pair_id_offset, d_hits, flags);
    if (rc == SATMC_OK) {
        ce = cudaMemcpyAsync(h_hits, d_hits, lead * sizeof(uint64_t), cudaMemcpyDeviceToHost, A);
        if (ce != cudaSuccess) rc = fail(ctx, SATMC_ERR_CUDA, "cudaMemcpyAsync (lead counters) failed: %s", cudaGetErrorString(ce));
    }
    if (rc == SATMC_OK) {
        ctx->stream = B; ctx->ticket_sel = 1;
        rc = satmc_count_fused(ctx, d_pairs + lead, rest, n_samples, seed, sample_offset, pair_id_offset + (uint32_t)lead,
                               d_hits + lead, flags);
        ctx->stream = A; ctx->ticket_sel = 0;
    }
    ctx->events_by_caller = false;
    if (rc == SATMC_OK) {
        ce = cudaMemcpyAsync(h_hits + lead, d_hits + lead, rest * sizeof(uint64_t), cudaMemcpyDeviceToHost, B);
        if (ce != cudaSuccess) rc = fail(ctx, SATMC_ERR_CUDA, "cudaMemcpyAsync (rest counters) failed: %s", cudaGetErrorString(ce));
    }
    // always rejoin the two streams, also after an error; a failure here is reported unless an earlier one already is
    cudaError_t cj = cudaEventRecord(ctx->ev_join, B);
    if (cj == cudaSuccess) cj = cudaStreamWaitEvent(A, ctx->ev_join, 0);
    if (cj == cudaSuccess) cj = cudaEventRecord(ctx->ev1, A);
    if (cj != cudaSuccess && rc == SATMC_OK) rc = fail(ctx, SATMC_ERR_CUDA, "rejoining the streams failed: %s", cudaGetErrorString(cj));
    ctx->last_ms_valid = (rc == SATMC_OK);
    if (rc) { cudaStreamSynchronize(B); cudaStreamSynchronize(A); cudaGetLastError(); return rc; }
    CU(ctx, cudaStreamSynchronize(A));
    CU(ctx, cudaGetLastError());
    return SATMC_OK;
}

int satmc_count_fused_host(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                           uint64_t sample_offset, uint32_t pair_id_offset, uint64_t* h_hits, uint32_t flags)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if ((!h_pairs || !h_hits) && n_pairs) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_pairs == 0) return SATMC_OK;
    if (n_pairs > 0xffffffffull - pair_id_offset) return fail(ctx, SATMC_ERR_INVALID, "pair ids exceed 32 bits");
    DeviceGuard g(ctx->device);
    void *d_pairs = nullptr, *d_hits = nullptr;
    int rc = scratch(ctx, 1, n_pairs * sizeof(satmc_pair), &d_pairs); if (rc) return rc;
    rc = scratch(ctx, 2, n_pairs * sizeof(uint64_t), &d_hits); if (rc) return rc;
    const uint64_t lead = (uint64_t)ctx->sm_count * ctx->blocks_per_sm * kWarps * 8;     // = plan_items' target: one chunk per pair
    if (n_samples > 0 && n_pairs >= 2 * lead)
        return count_fused_host_pipelined(ctx, h_pairs, n_pairs, lead, n_samples, seed, sample_offset, pair_id_offset, h_hits,
                                          flags, (satmc_pair*)d_pairs, (uint64_t*)d_hits);
    CU(ctx, cudaMemcpyAsync(d_pairs, h_pairs, n_pairs * sizeof(satmc_pair), cudaMemcpyHostToDevice, ctx->stream));
    if (flags & SATMC_ACCUMULATE)
        CU(ctx, cudaMemcpyAsync(d_hits, h_hits, n_pairs * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    const bool prof = ctx->profiling; ctx->profiling = true;
    rc = satmc_count_fused(ctx, (const satmc_pair*)d_pairs, n_pairs, n_samples, seed, sample_offset, pair_id_offset,
                           (uint64_t*)d_hits, flags);
    ctx->profiling = prof;
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(h_hits, d_hits, n_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SATMC_OK;
}

int satmc_count_streamed_host(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs, const float* h_z, uint64_t ldz,
                              uint64_t z_pair_stride, int ndof, uint64_t n_samples, uint64_t* h_hits, uint32_t flags)
{
    int rc = check_streamed_args(ctx, h_pairs, h_z, ldz, ndof, n_samples, n_pairs, z_pair_stride, h_hits);
    if (rc) return rc;
    if (n_pairs == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    void *d_pairs = nullptr, *d_hits = nullptr, *d_z = nullptr;
    const uint64_t ldz_dev = (ldz + 3) / 4 * 4;
    rc = scratch(ctx, 1, n_pairs * sizeof(satmc_pair), &d_pairs); if (rc) return rc;
    rc = scratch(ctx, 2, n_pairs * sizeof(uint64_t), &d_hits); if (rc) return rc;
    rc = scratch(ctx, 0, (size_t)ndof * ldz_dev * sizeof(float), &d_z); if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d_pairs, h_pairs, n_pairs * sizeof(satmc_pair), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpy2DAsync(d_z, ldz_dev * sizeof(float), h_z, ldz * sizeof(float), ldz * sizeof(float), (size_t)ndof,
                              cudaMemcpyHostToDevice, ctx->stream));
    if (flags & SATMC_ACCUMULATE)
        CU(ctx, cudaMemcpyAsync(d_hits, h_hits, n_pairs * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    const bool prof = ctx->profiling; ctx->profiling = true;
    rc = satmc_count_streamed(ctx, (const satmc_pair*)d_pairs, n_pairs, (const float*)d_z, ldz_dev, z_pair_stride, ndof,
                              n_samples, (uint64_t*)d_hits, flags);
    ctx->profiling = prof;
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(h_hits, d_hits, n_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SATMC_OK;
}

int satmc_collision_probability_host(satmc_ctx* ctx, const satmc_pair* h_pairs, uint64_t n_pairs, uint64_t n_samples,
                                     uint64_t seed, float* h_cp)
{
    if (!ctx) return fail(nullptr, SATMC_ERR_INVALID, "ctx is NULL");
    if ((!h_pairs || !h_cp) && n_pairs) return fail(ctx, SATMC_ERR_INVALID, "null pointer argument");
    if (n_samples == 0) return fail(ctx, SATMC_ERR_INVALID, "n_samples must be > 0");
    if (n_pairs == 0) return SATMC_OK;
    DeviceGuard g(ctx->device);
    void *d_pairs = nullptr, *d_hits = nullptr, *d_cp = nullptr;
    int rc = scratch(ctx, 1, n_pairs * sizeof(satmc_pair), &d_pairs); if (rc) return rc;
    rc = scratch(ctx, 2, n_pairs * sizeof(uint64_t), &d_hits); if (rc) return rc;
    rc = scratch(ctx, 0, n_pairs * sizeof(float), &d_cp); if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(d_pairs, h_pairs, n_pairs * sizeof(satmc_pair), cudaMemcpyHostToDevice, ctx->stream));
    const bool prof = ctx->profiling; ctx->profiling = true;
    rc = satmc_count_fused(ctx, (const satmc_pair*)d_pairs, n_pairs, n_samples, seed, 0, 0, (uint64_t*)d_hits, 0);
    ctx->profiling = prof;
    if (rc) return rc;
    k_hits_to_cp<<<(unsigned)((n_pairs + 255) / 256), 256, 0, ctx->stream>>>((const unsigned long long*)d_hits, n_pairs, n_samples,
                                                                           (float*)d_cp);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(h_cp, d_cp, n_pairs * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SATMC_OK;
}

int satmc_host_alloc(void** out, size_t bytes)
{
    if (!out) return fail(nullptr, SATMC_ERR_INVALID, "out is NULL");
    if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(nullptr, SATMC_ERR_NOMEM, "cudaMallocHost of %zu bytes failed", bytes);
    }
    return SATMC_OK;
}

int satmc_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, SATMC_ERR_CUDA, "cudaFreeHost failed"); }
    return SATMC_OK;
}

}  // extern "C"

