// oracle/ref_gpu.cu -- GPU oracle: the UNMODIFIED reference, compiled for sm_100a.
//
// TEST INFRASTRUCTURE ONLY (see oracle/sat_oracle.h).  This translation unit textually includes
// /root/reference/ztest.cu (which includes utils.cu) from where it lies -- no reference source is
// copied into this repository -- and exports thin extern "C" launchers around the reference's own
// __device__ functions and its Monte Carlo kernel.  The output (oracle/_ref/libref_gpu.so) is
// git-ignored and travels to the GPU box with gpurun.  Built by oracle/build_ref.sh.
//
// What each launcher pins:
//   ref_convex_collide   -> convex_collide            utils.cu:159-184
//   ref_rot_trans        -> rot_trans_rectangle       utils.cu:132-142
//   ref_sample_record    -> sample_rectangle          utils.cu:144-157  (normals recorded by replaying
//                           curand_normal on a copy of the state, then the reference's own function
//                           runs on the original state: shared samples drawn by the reference itself)
//   ref_mc_run           -> setup_kernel + monte_carlo_sample_collision_dataset_uniform
//                           utils.cu:111-117, ztest.cu:106-166 (optionally records every normal the
//                           kernel is about to draw so another implementation can consume the same ones)
//   ref_write_cp         -> write_collision_probability  utils.cu:210-215
//   ref_mc_time          -> the reference kernel timed with CUDA events (reference-GPU baseline)
//   ref_adaptive_batch   -> the reference's adaptive loop (kernel + thrust::count + thrust::sort_by_key + tail copies)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <numeric>
#include <cstring>
#include <chrono>
#include <cmath>

#define main ztest_reference_main
#include "ztest.cu"
#undef main

#define RG_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "ref_gpu: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while (0)

namespace {

__global__ void k_collide(const float* r1, const float* r2, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[8], b[8];
    for (int k = 0; k < 8; k++) { a[k] = r1[8 * (size_t)i + k]; b[k] = r2[8 * (size_t)i + k]; }
    out[i] = convex_collide(a, b);
}

__global__ void k_rot_trans(float* r, const float* dx, const float* dy, const float* dt, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[8];
    for (int k = 0; k < 8; k++) a[k] = r[8 * (size_t)i + k];
    rot_trans_rectangle(a, dx[i], dy[i], dt[i]);
    for (int k = 0; k < 8; k++) r[8 * (size_t)i + k] = a[k];
}

// one thread per base rectangle; each draws n_per samples with the reference's sample_rectangle.
__global__ void k_sample_record(curandState* state, const float* r_in, const StdDev* sd, int n, int n_per,
                                float* z_out, float* corners_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float base[8], out[8];
    for (int k = 0; k < 8; k++) base[k] = r_in[8 * (size_t)i + k];
    StdDev s = sd[i];
    size_t total = (size_t)n * n_per;
    for (int j = 0; j < n_per; j++) {
        size_t idx = (size_t)i * n_per + j;
        curandState replay = state[i];
        for (int k = 0; k < 5; k++) z_out[k * total + idx] = curand_normal(&replay);
        sample_rectangle(base, out, s, &state[i]);
        for (int k = 0; k < 8; k++) corners_out[8 * idx + k] = out[k];
    }
}

// records the normals thread g of the MC kernel will draw next (5 per sample, n_batch samples).
__global__ void k_record_normals(const curandState* state, int num_left, int n_batch, float* z_out, size_t ldz) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= num_left) return;
    curandState replay = state[g];
    for (int j = 0; j < n_batch; j++)
        for (int k = 0; k < 5; k++) z_out[k * ldz + (size_t)g * n_batch + j] = curand_normal(&replay);
}


// libdevice probes: precise sinf/cosf (what utils.cu:133-134 call) and the MUFU approximations the
// screening pass uses, so the oracle's restatement and the eps budget can be checked on the device.
__global__ void k_dev_sincos(const float* x, int n, float* s, float* c, int fast) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fast) { s[i] = __sinf(x[i]); c[i] = __cosf(x[i]); }
    else      { s[i] = sinf(x[i]);   c[i] = cosf(x[i]); }
}

// max |__sinf(x) - sin(x)|, |__cosf(x) - cos(x)| (double reference) over every float in [lo, hi] by bit pattern
__global__ void k_fast_trig_err(unsigned lo_bits, unsigned count, float* max_err_sin, float* max_err_cos) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    float es = 0.f, ec = 0.f;
    for (unsigned k = i; k < count; k += gridDim.x * blockDim.x) {
        float x = __uint_as_float(lo_bits + k);
        float e1 = fabsf((float)((double)__sinf(x) - sin((double)x)));
        float e2 = fabsf((float)((double)__cosf(x) - cos((double)x)));
        float e3 = fabsf((float)((double)__sinf(-x) - sin(-(double)x)));
        float e4 = fabsf((float)((double)__cosf(-x) - cos(-(double)x)));
        es = fmaxf(es, fmaxf(e1, e3)); ec = fmaxf(ec, fmaxf(e2, e4));
    }
    atomicMax((int*)max_err_sin, __float_as_int(es));
    atomicMax((int*)max_err_cos, __float_as_int(ec));
}

template <class T> struct DevBuf {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" {

int ref_convex_collide(const float* r1, const float* r2, int n, int* out) {
    DevBuf<float> a, b; DevBuf<int> o;
    RG_CHECK(a.alloc(8 * (size_t)n)); RG_CHECK(b.alloc(8 * (size_t)n)); RG_CHECK(o.alloc(n));
    RG_CHECK(cudaMemcpy(a.p, r1, 8 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(b.p, r2, 8 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    k_collide<<<(n + 255) / 256, 256>>>(a.p, b.p, n, o.p);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaMemcpy(out, o.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int ref_rot_trans(float* r, const float* dx, const float* dy, const float* dt, int n) {
    DevBuf<float> a, x, y, t;
    RG_CHECK(a.alloc(8 * (size_t)n)); RG_CHECK(x.alloc(n)); RG_CHECK(y.alloc(n)); RG_CHECK(t.alloc(n));
    RG_CHECK(cudaMemcpy(a.p, r, 8 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(x.p, dx, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(y.p, dy, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(t.p, dt, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    k_rot_trans<<<(n + 255) / 256, 256>>>(a.p, x.p, y.p, t.p, n);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaMemcpy(r, a.p, 8 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// r_in [n][8], sd [n][5] (std-devs), seed -> z_out [5][n*n_per] (SoA), corners_out [n*n_per][8]
int ref_sample_record(const float* r_in, const float* sd, int n, int n_per, int seed,
                      float* z_out, float* corners_out) {
    int padded = ((n + THREADS - 1) / THREADS) * THREADS;
    size_t total = (size_t)n * n_per;
    DevBuf<curandState> st; DevBuf<float> a, z, c; DevBuf<StdDev> s;
    RG_CHECK(st.alloc(padded)); RG_CHECK(a.alloc(8 * (size_t)n)); RG_CHECK(s.alloc(n));
    RG_CHECK(z.alloc(5 * total)); RG_CHECK(c.alloc(8 * total));
    RG_CHECK(cudaMemcpy(a.p, r_in, 8 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(s.p, sd, (size_t)n * sizeof(StdDev), cudaMemcpyHostToDevice));
    setup_kernel<<<padded / THREADS, THREADS>>>(st.p, seed);
    RG_CHECK(cudaGetLastError());
    k_sample_record<<<(n + 127) / 128, 128>>>(st.p, a.p, s.p, n, n_per, z.p, c.p);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaMemcpy(z_out, z.p, 5 * total * sizeof(float), cudaMemcpyDeviceToHost));
    RG_CHECK(cudaMemcpy(corners_out, c.p, 8 * total * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// One launch of the reference MC kernel exactly as ztest.cu:340-357 issues it.
// cps is in/out (running COUNT as float), done is out, z_record (optional) is [5][num_left*n_batch].
int ref_mc_run(const float* robot_base, const float* poses, int n_poses, const float* std_devs, int n_sd,
               const float* pose_idxs, const float* sd_idxs, const float* positions, float* cps,
               const float* accuracy_bins, int n_bins, const float* bin_accuracy, int* done,
               int n_samples, int n_batch, int num_left, int seed, float* z_record) {
    int padded = ((num_left + THREADS - 1) / THREADS) * THREADS;
    DevBuf<float> d_robot, d_pi, d_si, d_cp, d_bins, d_acc, d_z;
    DevBuf<Pose> d_poses; DevBuf<StdDev> d_sd; DevBuf<Position> d_pos; DevBuf<int> d_done;
    DevBuf<curandState> st;
    RG_CHECK(d_robot.alloc(8)); RG_CHECK(d_poses.alloc(n_poses)); RG_CHECK(d_sd.alloc(n_sd));
    RG_CHECK(d_pi.alloc(num_left)); RG_CHECK(d_si.alloc(num_left)); RG_CHECK(d_pos.alloc(num_left));
    RG_CHECK(d_cp.alloc(num_left)); RG_CHECK(d_done.alloc(num_left));
    RG_CHECK(d_bins.alloc(n_bins + 1)); RG_CHECK(d_acc.alloc(n_bins + 1)); RG_CHECK(st.alloc(padded));
    RG_CHECK(cudaMemset(d_bins.p, 0, (n_bins + 1) * sizeof(float)));
    RG_CHECK(cudaMemset(d_acc.p, 0, (n_bins + 1) * sizeof(float)));
    RG_CHECK(cudaMemcpy(d_robot.p, robot_base, 8 * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_poses.p, poses, (size_t)n_poses * sizeof(Pose), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_sd.p, std_devs, (size_t)n_sd * sizeof(StdDev), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_pi.p, pose_idxs, (size_t)num_left * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_si.p, sd_idxs, (size_t)num_left * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_pos.p, positions, (size_t)num_left * sizeof(Position), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_cp.p, cps, (size_t)num_left * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_bins.p, accuracy_bins, (size_t)n_bins * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_acc.p, bin_accuracy, (size_t)(n_bins - 1) * sizeof(float), cudaMemcpyHostToDevice));
    setup_kernel<<<padded / THREADS, THREADS>>>(st.p, seed);
    RG_CHECK(cudaGetLastError());
    if (z_record) {
        size_t ldz = (size_t)num_left * n_batch;
        RG_CHECK(d_z.alloc(5 * ldz));
        k_record_normals<<<padded / THREADS, THREADS>>>(st.p, num_left, n_batch, d_z.p, ldz);
        RG_CHECK(cudaGetLastError());
        RG_CHECK(cudaMemcpy(z_record, d_z.p, 5 * ldz * sizeof(float), cudaMemcpyDeviceToHost));
    }
    monte_carlo_sample_collision_dataset_uniform<<<padded / THREADS, THREADS>>>(
        d_robot.p, d_poses.p, d_sd.p, d_pi.p, d_si.p, d_pos.p, d_cp.p, d_bins.p, d_acc.p, n_bins, d_done.p,
        0, n_samples, n_batch, num_left, st.p);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaDeviceSynchronize());
    RG_CHECK(cudaMemcpy(cps, d_cp.p, (size_t)num_left * sizeof(float), cudaMemcpyDeviceToHost));
    RG_CHECK(cudaMemcpy(done, d_done.p, (size_t)num_left * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}


int ref_dev_sincos(const float* x, int n, float* s, float* c, int fast) {
    DevBuf<float> dx, ds, dc;
    RG_CHECK(dx.alloc(n)); RG_CHECK(ds.alloc(n)); RG_CHECK(dc.alloc(n));
    RG_CHECK(cudaMemcpy(dx.p, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    k_dev_sincos<<<(n + 255) / 256, 256>>>(dx.p, n, ds.p, dc.p, fast);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaMemcpy(s, ds.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    RG_CHECK(cudaMemcpy(c, dc.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// every float x with lo <= x <= hi (0 < lo < hi), both signs
int ref_fast_trig_err(float lo, float hi, float* err_sin, float* err_cos) {
    unsigned lb, hb; memcpy(&lb, &lo, 4); memcpy(&hb, &hi, 4);
    DevBuf<float> e; RG_CHECK(e.alloc(2));
    RG_CHECK(cudaMemset(e.p, 0, 2 * sizeof(float)));
    k_fast_trig_err<<<148 * 8, 256>>>(lb, hb - lb + 1, e.p, e.p + 1);
    RG_CHECK(cudaGetLastError());
    float h[2]; RG_CHECK(cudaMemcpy(h, e.p, sizeof(h), cudaMemcpyDeviceToHost));
    *err_sin = h[0]; *err_cos = h[1];
    return 0;
}

int ref_write_cp(float* counts, int n, int n_samples) {
    DevBuf<float> d;
    RG_CHECK(d.alloc(n));
    RG_CHECK(cudaMemcpy(d.p, counts, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    write_collision_probability<<<(n + THREADS - 1) / THREADS, THREADS>>>(d.p, n, n_samples);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaMemcpy(counts, d.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// Times `launches` back-to-back launches of the reference kernel (n_batch samples each) over
// num_left pairs in the reference's own indirect layout, bin_accuracy = 0 so nothing stops early
// (BASELINE.md 4b).  Returns total milliseconds in *ms_out (CUDA events), tests = num_left*n_batch*launches.
int ref_mc_time(const float* robot_base, const float* poses, int n_poses, const float* std_devs, int n_sd,
                const float* pose_idxs, const float* sd_idxs, const float* positions,
                int num_left, int n_batch, int launches, int seed, float* ms_out, float* cps_out) {
    int padded = ((num_left + THREADS - 1) / THREADS) * THREADS;
    DevBuf<float> d_robot, d_pi, d_si, d_cp, d_bins, d_acc;
    DevBuf<Pose> d_poses; DevBuf<StdDev> d_sd; DevBuf<Position> d_pos; DevBuf<int> d_done;
    DevBuf<curandState> st;
    const float bins[5] = {0.f, 0.01f, 0.1f, 1.f, 0.f};
    RG_CHECK(d_robot.alloc(8)); RG_CHECK(d_poses.alloc(n_poses)); RG_CHECK(d_sd.alloc(n_sd));
    RG_CHECK(d_pi.alloc(num_left)); RG_CHECK(d_si.alloc(num_left)); RG_CHECK(d_pos.alloc(num_left));
    RG_CHECK(d_cp.alloc(num_left)); RG_CHECK(d_done.alloc(num_left));
    RG_CHECK(d_bins.alloc(5)); RG_CHECK(d_acc.alloc(5)); RG_CHECK(st.alloc(padded));
    RG_CHECK(cudaMemset(d_acc.p, 0, 5 * sizeof(float)));
    RG_CHECK(cudaMemset(d_cp.p, 0, (size_t)num_left * sizeof(float)));
    RG_CHECK(cudaMemcpy(d_bins.p, bins, 5 * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_robot.p, robot_base, 8 * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_poses.p, poses, (size_t)n_poses * sizeof(Pose), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_sd.p, std_devs, (size_t)n_sd * sizeof(StdDev), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_pi.p, pose_idxs, (size_t)num_left * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_si.p, sd_idxs, (size_t)num_left * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_pos.p, positions, (size_t)num_left * sizeof(Position), cudaMemcpyHostToDevice));
    setup_kernel<<<padded / THREADS, THREADS>>>(st.p, seed);
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    RG_CHECK(cudaEventCreate(&e0)); RG_CHECK(cudaEventCreate(&e1));
    RG_CHECK(cudaEventRecord(e0));
    int n_samples = 0;
    for (int l = 0; l < launches; l++) {
        n_samples += n_batch;
        monte_carlo_sample_collision_dataset_uniform<<<padded / THREADS, THREADS>>>(
            d_robot.p, d_poses.p, d_sd.p, d_pi.p, d_si.p, d_pos.p, d_cp.p, d_bins.p, d_acc.p, 4, d_done.p,
            l, n_samples, n_batch, num_left, st.p);
    }
    RG_CHECK(cudaEventRecord(e1));
    RG_CHECK(cudaEventSynchronize(e1));
    RG_CHECK(cudaGetLastError());
    RG_CHECK(cudaEventElapsedTime(ms_out, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (cps_out) RG_CHECK(cudaMemcpy(cps_out, d_cp.p, (size_t)num_left * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// The reference's adaptive loop around its own kernel, as compute_collision_probability.cu:276-333 /
// generate_dataset.cu:420-479 run it (schedule 1000 -> 100000 at 20000 samples, thrust::count, stable
// thrust::sort_by_key compaction over a zip of (position, cp, var_idx, pose_idx, index),
// write_collision_probability on the finished tail, blocking D2H copies of the tail), timed end to end with a
// host clock.  The loop body is re-typed here because the reference's main() also does file I/O and needs boost;
// the kernel, the thrust calls and their order are the reference's.  Returns probabilities in input order.
int ref_adaptive_batch(const float* robot_base, const float* poses, int n_poses, const float* std_devs, int n_sd,
                       const float* pose_idxs, const float* sd_idxs, const float* positions, int n,
                       const float* accuracy_bins, int n_bins, const float* bin_accuracy, int max_samples, int seed,
                       float* cp_out, float* ms_out, long long* samples_out) {
    int padded = ((n + THREADS - 1) / THREADS) * THREADS;
    DevBuf<float> d_robot, d_pi, d_si, d_cp, d_bins, d_acc;
    DevBuf<Pose> d_poses; DevBuf<StdDev> d_sd; DevBuf<Position> d_pos; DevBuf<int> d_done, d_index;
    DevBuf<curandState> st;
    RG_CHECK(d_robot.alloc(8)); RG_CHECK(d_poses.alloc(n_poses)); RG_CHECK(d_sd.alloc(n_sd));
    RG_CHECK(d_pi.alloc(n)); RG_CHECK(d_si.alloc(n)); RG_CHECK(d_pos.alloc(n)); RG_CHECK(d_cp.alloc(n));
    RG_CHECK(d_done.alloc(n)); RG_CHECK(d_index.alloc(n)); RG_CHECK(d_bins.alloc(n_bins + 1)); RG_CHECK(d_acc.alloc(n_bins + 1));
    RG_CHECK(st.alloc(padded));
    RG_CHECK(cudaMemset(d_bins.p, 0, (n_bins + 1) * sizeof(float))); RG_CHECK(cudaMemset(d_acc.p, 0, (n_bins + 1) * sizeof(float)));
    RG_CHECK(cudaMemset(d_cp.p, 0, (size_t)n * sizeof(float)));
    std::vector<int> index(n); std::iota(index.begin(), index.end(), 0);
    std::vector<float> cp(n), vi(n), pi(n); std::vector<Position> pos(n);
    RG_CHECK(cudaMemcpy(d_robot.p, robot_base, 8 * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_poses.p, poses, (size_t)n_poses * sizeof(Pose), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_sd.p, std_devs, (size_t)n_sd * sizeof(StdDev), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_bins.p, accuracy_bins, (size_t)n_bins * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_acc.p, bin_accuracy, (size_t)(n_bins - 1) * sizeof(float), cudaMemcpyHostToDevice));
    setup_kernel<<<padded / THREADS, THREADS>>>(st.p, seed);
    RG_CHECK(cudaDeviceSynchronize());
    auto t0 = std::chrono::steady_clock::now();
    RG_CHECK(cudaMemcpy(d_index.p, index.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_pos.p, positions, (size_t)n * sizeof(Position), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_pi.p, pose_idxs, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_si.p, sd_idxs, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    DeviceZipIterator d_iter(thrust::make_tuple(thrust::device_pointer_cast(d_pos.p), thrust::device_pointer_cast(d_cp.p),
                                                thrust::device_pointer_cast(d_si.p), thrust::device_pointer_cast(d_pi.p),
                                                thrust::device_pointer_cast(d_index.p)));
    int num_left = n, n_samples = 0, iteration = 0;
    long long drawn = 0;
    while (num_left > 0 && n_samples < max_samples) {
        int blocks = (int)ceil((float)num_left / THREADS);
        int n_batch = n_samples < 20000 ? 1000 : 100000;
        n_samples += n_batch;
        drawn += (long long)num_left * n_batch;
        monte_carlo_sample_collision_dataset_uniform<<<blocks, THREADS>>>(d_robot.p, d_poses.p, d_sd.p, d_pi.p, d_si.p, d_pos.p, d_cp.p,
                                                                          d_bins.p, d_acc.p, n_bins, d_done.p, iteration, n_samples,
                                                                          n_batch, num_left, st.p);
        int batch_done = thrust::count(thrust::device, thrust::device_pointer_cast(d_done.p), thrust::device_pointer_cast(d_done.p + num_left), 1);
        if (batch_done > 0) {
            thrust::sort_by_key(thrust::device_pointer_cast(d_done.p), thrust::device_pointer_cast(d_done.p + num_left), d_iter);
            num_left -= batch_done;
            write_collision_probability<<<(int)ceil((float)batch_done / THREADS), THREADS>>>(d_cp.p + num_left, batch_done, n_samples);
            cudaMemcpy(pos.data() + num_left, d_pos.p + num_left, sizeof(Position) * batch_done, cudaMemcpyDeviceToHost);
            cudaMemcpy(cp.data() + num_left, d_cp.p + num_left, sizeof(float) * batch_done, cudaMemcpyDeviceToHost);
            cudaMemcpy(vi.data() + num_left, d_si.p + num_left, sizeof(float) * batch_done, cudaMemcpyDeviceToHost);
            cudaMemcpy(pi.data() + num_left, d_pi.p + num_left, sizeof(float) * batch_done, cudaMemcpyDeviceToHost);
            cudaMemcpy(index.data() + num_left, d_index.p + num_left, sizeof(int) * batch_done, cudaMemcpyDeviceToHost);
        }
        iteration++;
    }
    if (num_left > 0) {
        write_collision_probability<<<(int)ceil((float)num_left / THREADS), THREADS>>>(d_cp.p, num_left, n_samples);
        cudaMemcpy(cp.data(), d_cp.p, sizeof(float) * num_left, cudaMemcpyDeviceToHost);
        cudaMemcpy(index.data(), d_index.p, sizeof(int) * num_left, cudaMemcpyDeviceToHost);
    }
    RG_CHECK(cudaDeviceSynchronize());
    auto t1 = std::chrono::steady_clock::now();
    RG_CHECK(cudaGetLastError());
    for (int j = 0; j < n; j++) cp_out[index[j]] = cp[j];
    *ms_out = std::chrono::duration<float, std::milli>(t1 - t0).count();
    *samples_out = drawn;
    return 0;
}

}  // extern "C"
