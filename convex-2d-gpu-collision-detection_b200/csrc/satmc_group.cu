// satmc_group.cu -- the multi-GPU split of the Monte Carlo SAT path behind the C ABI (include/satmc.h, "groups").
//
// The path has no data dependence between units: it shards by pair (disjoint outputs, an
// all-gather at most) or by sample range of every pair (one all-reduce(SUM) of the 64-bit hit counters -- 8 bytes per
// pair, latency bound).  Because the normals of sample s of pair p depend only on (seed, p, s), every split returns
// the counts of the single-GPU call bit for bit.
//
// Two ways to form a group, same entry points afterwards:
//   satmc_group_create       one process drives n devices (ncclCommInitAll), one context + stream per device; all
//                            launches of a call are enqueued from the calling thread, device by device, and run
//                            concurrently; the collective is one ncclGroupStart/End over the local communicators.
//   satmc_group_create_rank  one process per GPU (torchrun / MPI style): rank 0 obtains an id with
//                            satmc_group_unique_id, the launcher distributes its 128 bytes, every rank joins with
//                            ncclCommInitRank.
// NCCL is bound at run time (dlopen "libnccl.so.2"): single-GPU users of libsatmc.so need no NCCL, and inside a
// PyTorch process the already loaded libnccl is reused.  A single-process group on one device needs no NCCL at all.
// The reference has no multi-GPU code (one device, default stream; a single pair with N = 1e11 is impossible there:
// `int n_samples`, float counter, ztest.cu:331,135,165).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "satmc_internal.hpp"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    char why[256] = {0};
    bool ok = false;
};

NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { snprintf(api.why, sizeof(api.why), "cannot load libnccl.so.2: %s", dlerror()); return; }
        bool all = true;
        auto bind = [&](auto& fn, const char* sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, sym));
            if (!fn) { all = false; snprintf(api.why, sizeof(api.why), "libnccl lacks %s", sym); }
        };
        bind(api.GetUniqueId, "ncclGetUniqueId"); bind(api.CommInitRank, "ncclCommInitRank"); bind(api.CommInitAll, "ncclCommInitAll");
        bind(api.CommDestroy, "ncclCommDestroy"); bind(api.AllReduce, "ncclAllReduce"); bind(api.AllGather, "ncclAllGather");
        bind(api.GroupStart, "ncclGroupStart"); bind(api.GroupEnd, "ncclGroupEnd"); bind(api.GetErrorString, "ncclGetErrorString");
        bind(api.GetVersion, "ncclGetVersion");
        api.ok = all;
    });
    return &api;
}

struct RestoreDevice {                           // the caller's current device survives every exit path
    int prev = -1;
    RestoreDevice() { cudaGetDevice(&prev); }
    ~RestoreDevice() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Local {                                   // one local device of a group
    satmc_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    ncclComm_t comm = nullptr;
    void* d_buf[4] = {nullptr, nullptr, nullptr, nullptr};   // grow-only: pairs, counters / cps, per-row inputs, spare
    size_t cap[4] = {0, 0, 0, 0};
    void* h_buf[2] = {nullptr, nullptr};         // pinned staging (interleaved rows in, padded results out)
    size_t h_cap[2] = {0, 0};
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    // resident tables of the adaptive path (satmc_group_set_tables)
    float *d_robot = nullptr, *d_poses = nullptr, *d_sds = nullptr, *d_bins = nullptr, *d_acc = nullptr;
};

}  // namespace

struct satmc_group {
    int world = 1, rank0 = 0;                    // ranks of this group's local devices: rank0 .. rank0 + n_local - 1
    std::vector<Local> dev;
    char err[512] = {0};
    uint32_t n_poses = 0, n_std = 0; int n_bins = 0;
    float kernel_ms = -1.f, collective_ms = -1.f;
    bool timing = false;
    int p2p = 0;                                 // 0 not probed, 1 every local device can reach device 0's memory, -1 no
    int use_p2p = -1;                            // -1 automatic (world >= 6: measured 121 vs 147 us per small call at 8 GPUs, 89 vs 80 at 4), 0 never, 1 whenever possible
    int last_exchange = SATMC_EXCHANGE_NONE;
    bool peer_acc_dirty = false;                 // a call failed between the kernels and the re-clearing of the accumulator
};

namespace {

int gfail(satmc_group* g, int code, const char* fmt, ...)
{
    char* dst = g ? g->err : satmc_thread_error();
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define GCU(g, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return gfail((g), SATMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define GNCCL(g, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return gfail((g), SATMC_ERR_NCCL, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); } while (0)
#define GSAT(g, l, call) do { int rc_ = (call); if (rc_ != SATMC_OK) \
    return gfail((g), rc_, "device %d: %s", (g)->dev[l].ctx->device, satmc_last_error((g)->dev[l].ctx)); } while (0)

int dev_buf(satmc_group* g, Local& L, int slot, size_t bytes, void** out)
{
    if (bytes > L.cap[slot]) {
        GCU(g, cudaStreamSynchronize(L.stream));
        if (L.d_buf[slot]) { GCU(g, cudaFree(L.d_buf[slot])); L.d_buf[slot] = nullptr; L.cap[slot] = 0; }
        const size_t cap = bytes + bytes / 4 + 256;
        if (cudaMalloc(&L.d_buf[slot], cap) != cudaSuccess) { cudaGetLastError(); return gfail(g, SATMC_ERR_NOMEM, "cudaMalloc of %zu bytes failed", cap); }
        L.cap[slot] = cap;
    }
    *out = L.d_buf[slot];
    return SATMC_OK;
}

int host_buf(satmc_group* g, Local& L, int slot, size_t bytes, void** out)
{
    if (bytes > L.h_cap[slot]) {
        GCU(g, cudaStreamSynchronize(L.stream));
        if (L.h_buf[slot]) { GCU(g, cudaFreeHost(L.h_buf[slot])); L.h_buf[slot] = nullptr; L.h_cap[slot] = 0; }
        const size_t cap = bytes + bytes / 4 + 256;
        if (cudaMallocHost(&L.h_buf[slot], cap) != cudaSuccess) { cudaGetLastError(); return gfail(g, SATMC_ERR_NOMEM, "cudaMallocHost of %zu bytes failed", cap); }
        L.h_cap[slot] = cap;
    }
    *out = L.h_buf[slot];
    return SATMC_OK;
}

int add_local(satmc_group* g, int device, void* stream, bool own)
{
    Local L;
    cudaStream_t st = (cudaStream_t)stream;
    GCU(g, cudaSetDevice(device));
    if (own) GCU(g, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    L.stream = st; L.own_stream = own;
    int rc = satmc_create(device, st, &L.ctx);
    if (rc != SATMC_OK) {
        if (own) cudaStreamDestroy(st);
        return gfail(g, rc, "device %d: %s", device, satmc_last_error(nullptr));
    }
    for (auto& e : L.ev) GCU(g, cudaEventCreate(&e));
    g->dev.push_back(L);
    return SATMC_OK;
}

void free_group(satmc_group* g)
{
    if (!g) return;
    RestoreDevice restore;
    for (Local& L : g->dev) {
        if (!L.ctx) continue;
        cudaSetDevice(L.ctx->device);
        cudaStreamSynchronize(L.stream);
        if (L.comm) nccl_api()->CommDestroy(L.comm);
        for (void* p : L.d_buf) if (p) cudaFree(p);
        for (void* p : L.h_buf) if (p) cudaFreeHost(p);
        for (float* p : {L.d_robot, L.d_poses, L.d_sds, L.d_bins, L.d_acc}) if (p) cudaFree(p);
        for (cudaEvent_t e : L.ev) if (e) cudaEventDestroy(e);
        satmc_destroy(L.ctx);
        if (L.own_stream) cudaStreamDestroy(L.stream);
    }
    delete g;
}

// single-process groups: ncclCommInitAll over the local devices, on first use
int ensure_comms(satmc_group* g)
{
    if (g->world == 1 || g->dev[0].comm != nullptr) return SATMC_OK;
    NcclApi* N = nccl_api();
    if (!N->ok) return gfail(g, SATMC_ERR_NCCL, "NCCL unavailable: %s", N->why);
    const int n = (int)g->dev.size();
    if (n != g->world) return gfail(g, SATMC_ERR_NCCL, "internal: communicator missing");   // create_rank initialises eagerly
    std::vector<ncclComm_t> comms(n);
    std::vector<int> ids(n);
    for (int i = 0; i < n; i++) ids[i] = g->dev[i].ctx->device;
    ncclResult_t r = N->CommInitAll(comms.data(), n, ids.data());
    if (r != ncclSuccess) return gfail(g, SATMC_ERR_NCCL, "ncclCommInitAll failed: %s", N->GetErrorString(r));
    for (int i = 0; i < n; i++) g->dev[i].comm = comms[i];
    return SATMC_OK;
}

// all-reduce(SUM) of n 64-bit counters, in place, on every local device (sample-range sharding)
int all_reduce_u64(satmc_group* g, std::vector<void*>& bufs, size_t n)
{
    if (g->world == 1) return SATMC_OK;
    int rc = ensure_comms(g);
    if (rc) return rc;
    NcclApi* N = nccl_api();
    GNCCL(g, N->GroupStart());
    for (size_t l = 0; l < g->dev.size(); l++)
        GNCCL(g, N->AllReduce(bufs[l], bufs[l], n, ncclUint64, ncclSum, g->dev[l].comm, g->dev[l].stream));
    GNCCL(g, N->GroupEnd());
    return SATMC_OK;
}

// in-place all-gather: rank r's `chunk` elements sit at offset r * chunk of every buffer
int all_gather(satmc_group* g, std::vector<void*>& bufs, size_t chunk, size_t elem_bytes)
{
    if (g->world == 1) return SATMC_OK;
    int rc = ensure_comms(g);
    if (rc) return rc;
    NcclApi* N = nccl_api();
    GNCCL(g, N->GroupStart());
    for (size_t l = 0; l < g->dev.size(); l++) {
        char* base = static_cast<char*>(bufs[l]);
        GNCCL(g, N->AllGather(base + (size_t)(g->rank0 + (int)l) * chunk * elem_bytes, base, chunk * elem_bytes, ncclUint8,
                              g->dev[l].comm, g->dev[l].stream));
    }
    GNCCL(g, N->GroupEnd());
    return SATMC_OK;
}

// Peer access from every local device to local device 0 (NVLink / NVSwitch): probed and enabled once.
bool p2p_ready(satmc_group* g)
{
    if (g->p2p != 0) return g->p2p > 0;
    g->p2p = -1;
    const int root = g->dev[0].ctx->device;
    for (size_t l = 1; l < g->dev.size(); l++) {
        int can = 0;
        const int d = g->dev[l].ctx->device;
        if (cudaDeviceCanAccessPeer(&can, d, root) != cudaSuccess || !can) { cudaGetLastError(); return false; }
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); return false; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(root, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
        cudaGetLastError();
    }
    g->p2p = 1;
    return true;
}

int sync_all(satmc_group* g)
{
    for (Local& L : g->dev) { GCU(g, cudaSetDevice(L.ctx->device)); GCU(g, cudaStreamSynchronize(L.stream)); }
    return SATMC_OK;
}

}  // namespace

extern "C" {

int satmc_shard_range(int shard_mode, uint64_t n_units, int world, int rank, uint64_t* lo, uint64_t* hi)
{
    if (!lo || !hi || world < 1 || rank < 0 || rank >= world) return gfail(nullptr, SATMC_ERR_INVALID, "satmc_shard_range: bad arguments");
    if (shard_mode == SATMC_SHARD_BY_PAIR) {
        // equal chunks of ceil(n / world): rank r's slice starts at r * chunk, which is also where an in-place
        // all-gather expects it; the last ranks may get a short or empty slice
        const uint64_t chunk = (n_units + (uint64_t)world - 1) / (uint64_t)world;
        uint64_t a = (uint64_t)rank * chunk, b = a + chunk;
        *lo = a < n_units ? a : n_units; *hi = b < n_units ? b : n_units;
        return SATMC_OK;
    }
    if (shard_mode == SATMC_SHARD_BY_SAMPLE_RANGE) {
        // blocks of 4 samples (the sampler's group size), dealt as evenly as possible
        const uint64_t blocks = (n_units + 3) / 4, base = blocks / (uint64_t)world, rem = blocks % (uint64_t)world;
        const uint64_t r = (uint64_t)rank;
        const uint64_t a = (r * base + (r < rem ? r : rem)) * 4, b = a + (base + (r < rem ? 1 : 0)) * 4;
        *lo = a < n_units ? a : n_units; *hi = b < n_units ? b : n_units;
        return SATMC_OK;
    }
    if (shard_mode == SATMC_SHARD_INTERLEAVED) {
        // rows rank, rank + world, ...: *lo = first row, *hi = number of rows of this rank
        *lo = (uint64_t)rank < n_units ? (uint64_t)rank : n_units;
        *hi = (n_units + (uint64_t)world - 1 - (uint64_t)rank) / (uint64_t)world;
        return SATMC_OK;
    }
    return gfail(nullptr, SATMC_ERR_INVALID, "satmc_shard_range: unknown shard mode %d", shard_mode);
}

int satmc_group_create(const int* devices, int n_dev, satmc_group** out)
{
    if (!out) return gfail(nullptr, SATMC_ERR_INVALID, "satmc_group_create: out is NULL");
    *out = nullptr;
    if (n_dev < 1 || n_dev > 64) return gfail(nullptr, SATMC_ERR_INVALID, "satmc_group_create: n_dev %d out of range", n_dev);
    satmc_group* g = new (std::nothrow) satmc_group();
    if (!g) return gfail(nullptr, SATMC_ERR_NOMEM, "out of host memory");
    RestoreDevice restore;
    g->world = n_dev; g->rank0 = 0;
    int rc = SATMC_OK;
    for (int i = 0; i < n_dev && rc == SATMC_OK; i++) rc = add_local(g, devices ? devices[i] : i, nullptr, true);
    // the communicators are created by the first call that needs a collective (ensure_comms): the adaptive path and
    // host-buffer calls sharded by pair read every slice from the device that computed it and never pay for NCCL
    if (rc != SATMC_OK) { snprintf(satmc_thread_error(), 512, "%s", g->err); free_group(g); return rc; }
    *out = g;
    return SATMC_OK;
}

int satmc_group_unique_id(void* out128)
{
    if (!out128) return gfail(nullptr, SATMC_ERR_INVALID, "satmc_group_unique_id: out is NULL");
    NcclApi* N = nccl_api();
    if (!N->ok) return gfail(nullptr, SATMC_ERR_NCCL, "NCCL unavailable: %s", N->why);
    ncclUniqueId id;
    ncclResult_t r = N->GetUniqueId(&id);
    if (r != ncclSuccess) return gfail(nullptr, SATMC_ERR_NCCL, "ncclGetUniqueId failed: %s", N->GetErrorString(r));
    static_assert(sizeof(id) == SATMC_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(out128, &id, sizeof(id));
    return SATMC_OK;
}

int satmc_group_create_rank(const void* unique_id, int world, int rank, int device, void* stream, satmc_group** out)
{
    if (!out) return gfail(nullptr, SATMC_ERR_INVALID, "satmc_group_create_rank: out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return gfail(nullptr, SATMC_ERR_INVALID, "bad rank %d of %d", rank, world);
    if (world > 1 && !unique_id) return gfail(nullptr, SATMC_ERR_INVALID, "unique_id is NULL");
    satmc_group* g = new (std::nothrow) satmc_group();
    if (!g) return gfail(nullptr, SATMC_ERR_NOMEM, "out of host memory");
    RestoreDevice restore;
    g->world = world; g->rank0 = rank;
    int rc = add_local(g, device, stream, false);      // NULL = the legacy default stream, as in satmc_create
    if (rc == SATMC_OK && world > 1) {
        NcclApi* N = nccl_api();
        if (!N->ok) rc = gfail(g, SATMC_ERR_NCCL, "NCCL unavailable: %s", N->why);
        else {
            ncclUniqueId id;
            memcpy(&id, unique_id, sizeof(id));
            ncclResult_t r = N->CommInitRank(&g->dev[0].comm, world, id, rank);
            if (r != ncclSuccess) rc = gfail(g, SATMC_ERR_NCCL, "ncclCommInitRank failed: %s", N->GetErrorString(r));
        }
    }
    if (rc != SATMC_OK) { snprintf(satmc_thread_error(), 512, "%s", g->err); free_group(g); return rc; }
    *out = g;
    return SATMC_OK;
}

int satmc_group_destroy(satmc_group* g) { free_group(g); return SATMC_OK; }
int satmc_group_world(const satmc_group* g) { return g ? g->world : 0; }
int satmc_group_local_count(const satmc_group* g) { return g ? (int)g->dev.size() : 0; }
int satmc_group_rank(const satmc_group* g, int local) { return (g && local >= 0 && local < (int)g->dev.size()) ? g->rank0 + local : -1; }
satmc_ctx* satmc_group_context(satmc_group* g, int local) { return (g && local >= 0 && local < (int)g->dev.size()) ? g->dev[local].ctx : nullptr; }
const char* satmc_group_last_error(const satmc_group* g) { return g ? g->err : satmc_thread_error(); }

int satmc_group_nccl_version(void)
{
    NcclApi* N = nccl_api();
    int v = 0;
    if (!N->ok || N->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int satmc_group_set_peer_reduce(satmc_group* g, int enabled)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    g->use_p2p = enabled < 0 ? -1 : (enabled != 0 ? 1 : 0);
    return SATMC_OK;
}

int satmc_group_last_exchange(const satmc_group* g) { return g ? g->last_exchange : SATMC_EXCHANGE_NONE; }

int satmc_group_set_timing(satmc_group* g, int enabled)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    g->timing = enabled != 0;
    return SATMC_OK;
}

int satmc_group_last_times(const satmc_group* g, float* kernel_ms, float* collective_ms)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    if (kernel_ms) *kernel_ms = g->kernel_ms;
    if (collective_ms) *collective_ms = g->collective_ms;
    return SATMC_OK;
}

uint64_t satmc_group_hits_capacity(const satmc_group* g, uint64_t n_pairs)
{
    if (!g) return 0;
    const uint64_t chunk = (n_pairs + (uint64_t)g->world - 1) / (uint64_t)g->world;
    return chunk * (uint64_t)g->world;
}

// The counting step on resident inputs: d_pairs[l] is the FULL pair array on local device l, d_hits[l] an array of
// satmc_group_hits_capacity() counters there.  On return (asynchronously, on the group's streams) every d_hits[l]
// holds the counts of all n_pairs pairs.
int satmc_group_count_fused(satmc_group* g, const satmc_pair* const* d_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                            uint64_t sample_offset, uint32_t pair_id_offset, int shard_mode, uint64_t* const* d_hits, uint32_t flags)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    if (!d_pairs || !d_hits) return gfail(g, SATMC_ERR_INVALID, "null pointer argument");
    if (shard_mode != SATMC_SHARD_BY_PAIR && shard_mode != SATMC_SHARD_BY_SAMPLE_RANGE)
        return gfail(g, SATMC_ERR_INVALID, "shard mode %d not valid for counting", shard_mode);
    if (flags & SATMC_ACCUMULATE) return gfail(g, SATMC_ERR_INVALID, "SATMC_ACCUMULATE is not defined for group calls");
    if (n_pairs == 0) return SATMC_OK;
    if (n_pairs > 0xffffffffull - pair_id_offset) return gfail(g, SATMC_ERR_INVALID, "pair ids exceed 32 bits");
    const size_t nl = g->dev.size();
    std::vector<void*> bufs(nl);
    RestoreDevice restore;
    const uint64_t chunk = (n_pairs + (uint64_t)g->world - 1) / (uint64_t)g->world;
    for (size_t l = 0; l < nl; l++) {
        Local& L = g->dev[l];
        if (!d_pairs[l] || !d_hits[l]) return gfail(g, SATMC_ERR_INVALID, "null device pointer for local device %zu", l);
        GCU(g, cudaSetDevice(L.ctx->device));
        uint64_t lo = 0, hi = 0;
        bufs[l] = d_hits[l];
        if (g->timing && l == 0) GCU(g, cudaEventRecord(L.ev[0], L.stream));
        if (shard_mode == SATMC_SHARD_BY_PAIR) {
            satmc_shard_range(shard_mode, n_pairs, g->world, g->rank0 + (int)l, &lo, &hi);
            if (hi > lo)
                GSAT(g, l, satmc_count_fused(L.ctx, d_pairs[l] + lo, hi - lo, n_samples, seed, sample_offset, pair_id_offset + (uint32_t)lo,
                                             d_hits[l] + lo, flags));
            (void)chunk;
        } else {
            satmc_shard_range(shard_mode, n_samples, g->world, g->rank0 + (int)l, &lo, &hi);
            GSAT(g, l, satmc_count_fused(L.ctx, d_pairs[l], n_pairs, hi - lo, seed, sample_offset + lo, pair_id_offset, d_hits[l], flags));
        }
        if (g->timing && l == 0) GCU(g, cudaEventRecord(L.ev[1], L.stream));
    }
    int rc = (shard_mode == SATMC_SHARD_BY_PAIR) ? all_gather(g, bufs, chunk, sizeof(uint64_t)) : all_reduce_u64(g, bufs, n_pairs);
    if (rc == SATMC_OK && g->timing) {
        Local& L = g->dev[0];
        cudaSetDevice(L.ctx->device);
        GCU(g, cudaEventRecord(L.ev[2], L.stream));
        GCU(g, cudaEventSynchronize(L.ev[2]));
        GCU(g, cudaEventElapsedTime(&g->kernel_ms, L.ev[0], L.ev[1]));
        GCU(g, cudaEventElapsedTime(&g->collective_ms, L.ev[1], L.ev[2]));
    }
    return rc;
}

int satmc_group_synchronize(satmc_group* g)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    RestoreDevice restore;
    int rc = sync_all(g);
    return rc;
}

// Host buffers: every rank passes the same h_pairs; h_hits receives all n_pairs counts on every rank.
int satmc_group_count_fused_host(satmc_group* g, const satmc_pair* h_pairs, uint64_t n_pairs, uint64_t n_samples, uint64_t seed,
                                 uint64_t sample_offset, uint32_t pair_id_offset, int shard_mode, uint64_t* h_hits, uint32_t flags)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    if ((!h_pairs || !h_hits) && n_pairs) return gfail(g, SATMC_ERR_INVALID, "null pointer argument");
    if (n_pairs == 0) return SATMC_OK;
    if (shard_mode != SATMC_SHARD_BY_PAIR && shard_mode != SATMC_SHARD_BY_SAMPLE_RANGE)
        return gfail(g, SATMC_ERR_INVALID, "shard mode %d not valid for counting", shard_mode);
    const size_t nl = g->dev.size();
    const uint64_t cap = satmc_group_hits_capacity(g, n_pairs);
    std::vector<const satmc_pair*> dp(nl);
    std::vector<uint64_t*> dh(nl);
    RestoreDevice restore;
    for (size_t l = 0; l < nl; l++) {
        Local& L = g->dev[l];
        GCU(g, cudaSetDevice(L.ctx->device));
        void *p = nullptr, *h = nullptr;
        int rc = dev_buf(g, L, 0, n_pairs * sizeof(satmc_pair), &p); if (rc) return rc;
        rc = dev_buf(g, L, 1, cap * sizeof(uint64_t), &h); if (rc) return rc;
        dp[l] = static_cast<const satmc_pair*>(p); dh[l] = static_cast<uint64_t*>(h);
        uint64_t lo = 0, hi = n_pairs;                       // by pair: only this rank's slice needs to travel
        if (shard_mode == SATMC_SHARD_BY_PAIR) satmc_shard_range(shard_mode, n_pairs, g->world, g->rank0 + (int)l, &lo, &hi);
        if (hi > lo)
            GCU(g, cudaMemcpyAsync(static_cast<satmc_pair*>(p) + lo, h_pairs + lo, (hi - lo) * sizeof(satmc_pair), cudaMemcpyHostToDevice, L.stream));
    }
    // a single process reads every slice from the device that computed it; no collective is needed by pair
    const bool direct = (shard_mode == SATMC_SHARD_BY_PAIR) && (int)nl == g->world;
    // Sample ranges inside one process, few pairs (every shard is cut into several work items per pair): the compute
    // kernels themselves finish into ONE counter array in device 0's memory with system-scope atomics over NVLink --
    // the reduction is the kernels' own epilogue, no collective is launched.
    const bool want_peer = g->use_p2p > 0 || (g->use_p2p < 0 && g->world >= 6);
    bool peer = (shard_mode == SATMC_SHARD_BY_SAMPLE_RANGE) && (int)nl == g->world && g->world > 1 && want_peer && !(flags & SATMC_ACCUMULATE);
    if (peer) {
        for (size_t l = 0; l < nl && peer; l++) {
            uint64_t lo = 0, hi = 0;
            satmc_shard_range(shard_mode, n_samples, g->world, g->rank0 + (int)l, &lo, &hi);
            peer = hi == lo || satmc_fused_is_multi(g->dev[l].ctx, n_pairs, hi - lo);
        }
        peer = peer && p2p_ready(g);
    }
    int rc = SATMC_OK;
    g->last_exchange = g->world == 1 || direct ? SATMC_EXCHANGE_NONE : (peer ? SATMC_EXCHANGE_PEER_ATOMICS : SATMC_EXCHANGE_NCCL);
    if (peer) {
        // The accumulator in device 0's memory holds zeros between calls (cleared when allocated and again right after
        // every read-out, which the end-of-call synchronisation covers), so every device's kernel starts at once: no
        // memset and no event in front of them.  Device 0 only waits for the others' kernels before it reads the totals.
        Local& R = g->dev[0];
        GCU(g, cudaSetDevice(R.ctx->device));
        const size_t had = R.cap[3];
        void* accv = nullptr;
        rc = dev_buf(g, R, 3, n_pairs * sizeof(uint64_t), &accv);
        if (rc) return rc;
        if (R.cap[3] != had || g->peer_acc_dirty) { GCU(g, cudaMemsetAsync(accv, 0, R.cap[3], R.stream)); GCU(g, cudaStreamSynchronize(R.stream)); }
        g->peer_acc_dirty = true;
        uint64_t* acc = static_cast<uint64_t*>(accv);
        for (size_t l = 0; l < nl; l++) {
            Local& L = g->dev[l];
            GCU(g, cudaSetDevice(L.ctx->device));
            uint64_t lo = 0, hi = 0;
            satmc_shard_range(shard_mode, n_samples, g->world, g->rank0 + (int)l, &lo, &hi);
            if (hi > lo)
                GSAT(g, l, satmc_count_fused_impl(L.ctx, dp[l], n_pairs, hi - lo, seed, sample_offset + lo, pair_id_offset, acc,
                                                  (flags & SATMC_EXACT_ONLY) | SATMC_PEER_ATOMIC_OUT));
            if (l > 0) GCU(g, cudaEventRecord(L.ev[1], L.stream));
        }
        GCU(g, cudaSetDevice(R.ctx->device));
        for (size_t l = 1; l < nl; l++) GCU(g, cudaStreamWaitEvent(R.stream, g->dev[l].ev[1], 0));
        GCU(g, cudaMemcpyAsync(h_hits, acc, n_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, R.stream));
        GCU(g, cudaMemsetAsync(acc, 0, n_pairs * sizeof(uint64_t), R.stream));
        g->peer_acc_dirty = false;
    } else if (direct) {
        for (size_t l = 0; l < nl && rc == SATMC_OK; l++) {
            Local& L = g->dev[l];
            GCU(g, cudaSetDevice(L.ctx->device));
            uint64_t lo = 0, hi = 0;
            satmc_shard_range(shard_mode, n_pairs, g->world, g->rank0 + (int)l, &lo, &hi);
            if (hi <= lo) continue;
            GSAT(g, l, satmc_count_fused(L.ctx, dp[l] + lo, hi - lo, n_samples, seed, sample_offset, pair_id_offset + (uint32_t)lo, dh[l] + lo, flags));
            GCU(g, cudaMemcpyAsync(h_hits + lo, dh[l] + lo, (hi - lo) * sizeof(uint64_t), cudaMemcpyDeviceToHost, L.stream));
        }
    } else {
        rc = satmc_group_count_fused(g, dp.data(), n_pairs, n_samples, seed, sample_offset, pair_id_offset, shard_mode, dh.data(), flags);
        if (rc == SATMC_OK) {
            Local& L = g->dev[0];
            GCU(g, cudaSetDevice(L.ctx->device));
            GCU(g, cudaMemcpyAsync(h_hits, dh[0], n_pairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, L.stream));
        }
    }
    if (rc == SATMC_OK) rc = sync_all(g);
    return rc;
}

// ---- adaptive z-test over rows dealt round-robin to the ranks ----------------------------------------------------

int satmc_group_set_tables(satmc_group* g, const float* h_robot_base, const float* h_poses, uint32_t n_poses, const float* h_std_devs,
                           uint32_t n_std, const float* h_accuracy_bins, const float* h_bin_accuracy, int n_accuracy_bins)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    if (!h_robot_base || !h_poses || !h_std_devs || !h_accuracy_bins || !h_bin_accuracy || n_poses == 0 || n_std == 0 || n_accuracy_bins < 2)
        return gfail(g, SATMC_ERR_INVALID, "null pointer, empty table or fewer than two bin edges");
    RestoreDevice restore;
    for (Local& L : g->dev) {
        GCU(g, cudaSetDevice(L.ctx->device));
        GCU(g, cudaStreamSynchronize(L.stream));
        for (float** p : {&L.d_robot, &L.d_poses, &L.d_sds, &L.d_bins, &L.d_acc}) if (*p) { cudaFree(*p); *p = nullptr; }
        struct { float** d; const float* h; size_t n; } up[] = {
            {&L.d_robot, h_robot_base, 8}, {&L.d_poses, h_poses, 3 * (size_t)n_poses}, {&L.d_sds, h_std_devs, 5 * (size_t)n_std},
            {&L.d_bins, h_accuracy_bins, (size_t)n_accuracy_bins}, {&L.d_acc, h_bin_accuracy, (size_t)n_accuracy_bins - 1}};
        for (auto& u : up) {
            if (cudaMalloc(u.d, u.n * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return gfail(g, SATMC_ERR_NOMEM, "cudaMalloc of a table failed"); }
            GCU(g, cudaMemcpyAsync(*u.d, u.h, u.n * sizeof(float), cudaMemcpyHostToDevice, L.stream));
        }
    }
    int rc = sync_all(g);
    g->n_poses = n_poses; g->n_std = n_std; g->n_bins = n_accuracy_bins;
    return rc;
}

int satmc_group_adaptive_run_host(satmc_group* g, const float* h_pose_idxs, const float* h_std_dev_idxs, const float* h_positions,
                                  int n_rows, int max_samples, int n_batch_small, int switch_at, int n_batch_large, uint64_t seed,
                                  uint32_t stream_id_offset, float* h_cp_out, int* iterations_out, long long* samples_drawn_out)
{
    if (!g) return gfail(nullptr, SATMC_ERR_INVALID, "group is NULL");
    if (iterations_out) *iterations_out = 0;
    if (samples_drawn_out) *samples_drawn_out = 0;
    if (g->n_poses == 0) return gfail(g, SATMC_ERR_INVALID, "satmc_group_set_tables has not been called");
    if (n_rows < 0 || ((!h_pose_idxs || !h_std_dev_idxs || !h_positions || !h_cp_out) && n_rows))
        return gfail(g, SATMC_ERR_INVALID, "null pointer argument or negative row count");
    if (n_rows == 0) return SATMC_OK;
    if ((uint64_t)(n_rows - 1) > 0xffffffffull - stream_id_offset) return gfail(g, SATMC_ERR_INVALID, "Philox stream ids exceed 32 bits");
    const size_t nl = g->dev.size();
    const int world = g->world;
    const size_t chunk = ((size_t)n_rows + world - 1) / world;         // rows per rank, padded
    std::vector<AdaptiveRun> runs(nl);
    std::vector<void*> cps(nl);
    RestoreDevice restore;
    // work per row varies ~400x (1e4 .. 4e6 samples, generate_dataset.cu:53,427-431) and neighbouring rows of a file are
    // often alike, so rank r takes rows r, r + world, ...: Philox stream of local row e = offset + r + e * world
    for (size_t l = 0; l < nl; l++) {
        Local& L = g->dev[l];
        const int r = g->rank0 + (int)l;
        GCU(g, cudaSetDevice(L.ctx->device));
        uint64_t first = 0, k = 0;
        satmc_shard_range(SATMC_SHARD_INTERLEAVED, (uint64_t)n_rows, world, r, &first, &k);
        void *h_in = nullptr, *d_in = nullptr, *d_cp = nullptr;
        int rc = host_buf(g, L, 0, 4 * chunk * sizeof(float), &h_in); if (rc) return rc;
        rc = dev_buf(g, L, 2, 4 * chunk * sizeof(float), &d_in); if (rc) return rc;
        rc = dev_buf(g, L, 1, chunk * (size_t)world * sizeof(float), &d_cp); if (rc) return rc;
        float* hp = static_cast<float*>(h_in);                          // [pose_idx | sd_idx | positions(2)] of this rank's rows
        for (size_t e = 0; e < k; e++) {
            const size_t row = (size_t)r + e * (size_t)world;
            hp[e] = h_pose_idxs[row]; hp[chunk + e] = h_std_dev_idxs[row];
            hp[2 * chunk + 2 * e] = h_positions[2 * row]; hp[2 * chunk + 2 * e + 1] = h_positions[2 * row + 1];
        }
        GCU(g, cudaMemcpyAsync(d_in, h_in, 4 * chunk * sizeof(float), cudaMemcpyHostToDevice, L.stream));
        float* di = static_cast<float*>(d_in);
        cps[l] = d_cp;
        AdaptiveRun& ar = runs[l];
        ar = AdaptiveRun{};
        ar.ctx = L.ctx; ar.d_robot_base = L.d_robot; ar.d_poses = L.d_poses; ar.n_poses = g->n_poses; ar.d_std_devs = L.d_sds; ar.n_std = g->n_std;
        ar.d_pose_idxs = di; ar.d_std_dev_idxs = di + chunk; ar.d_positions = di + 2 * chunk; ar.n_pairs = (int)k;
        ar.d_bins = L.d_bins; ar.d_bin_acc = L.d_acc; ar.n_bins = g->n_bins; ar.max_samples = max_samples;
        ar.n_batch_small = n_batch_small; ar.switch_at = switch_at; ar.n_batch_large = n_batch_large; ar.seed = seed;
        ar.stream_id_offset = stream_id_offset + (uint32_t)r; ar.stream_id_stride = (uint32_t)world;
        ar.d_cp_out = static_cast<float*>(d_cp) + (size_t)r * chunk; ar.d_n_samples_out = nullptr;
        GSAT(g, l, satmc_adaptive_begin(ar));
    }
    // lockstep: one iteration on every device that still has live rows, then one synchronisation each
    std::vector<char> busy(nl);
    for (;;) {
        bool any = false;
        for (size_t l = 0; l < nl; l++) {
            busy[l] = satmc_adaptive_pending(runs[l]) ? 1 : 0;
            if (!busy[l]) continue;
            any = true;
            GSAT(g, l, satmc_adaptive_enqueue(runs[l]));
        }
        if (!any) break;
        for (size_t l = 0; l < nl; l++) {                              // (enqueue advanced n_samples: go by what was enqueued)
            if (!busy[l]) continue;
            GCU(g, cudaSetDevice(g->dev[l].ctx->device));
            GCU(g, cudaStreamSynchronize(g->dev[l].stream));
            satmc_adaptive_collect(runs[l]);
        }
    }
    int iters = 0; long long drawn = 0;
    for (size_t l = 0; l < nl; l++) {
        GSAT(g, l, satmc_adaptive_finish(runs[l]));
        if (runs[l].iter > iters) iters = runs[l].iter;
        drawn += runs[l].drawn;
    }
    // results: rank r's chunk sits at r * chunk of every buffer after the all-gather; then undo the interleaving
    const bool direct = (int)nl == world;
    int rc = direct ? SATMC_OK : all_gather(g, cps, chunk, sizeof(float));
    if (rc) return rc;
    std::vector<float*> h_out(nl, nullptr);
    for (size_t l = 0; l < nl; l++) {
        Local& L = g->dev[l];
        if (!direct && l > 0) break;
        GCU(g, cudaSetDevice(L.ctx->device));
        void* h = nullptr;
        rc = host_buf(g, L, 1, chunk * (size_t)world * sizeof(float), &h); if (rc) return rc;
        h_out[l] = static_cast<float*>(h);
        if (direct) {
            const size_t off = (size_t)(g->rank0 + (int)l) * chunk;
            GCU(g, cudaMemcpyAsync(h_out[l] + off, static_cast<float*>(cps[l]) + off, chunk * sizeof(float), cudaMemcpyDeviceToHost, L.stream));
        } else {
            GCU(g, cudaMemcpyAsync(h_out[l], cps[l], chunk * (size_t)world * sizeof(float), cudaMemcpyDeviceToHost, L.stream));
        }
    }
    rc = sync_all(g);
    if (rc) return rc;
    for (int r = 0; r < world; r++) {
        const float* src = (direct ? h_out[(size_t)(r - g->rank0)] : h_out[0]) + (size_t)r * chunk;
        size_t e = 0;
        for (size_t row = (size_t)r; row < (size_t)n_rows; row += (size_t)world, e++) h_cp_out[row] = src[e];
    }
    if (iterations_out) *iterations_out = iters;
    if (samples_drawn_out) *samples_drawn_out = drawn;       // of the local devices
    return SATMC_OK;
}

}  // extern "C"
