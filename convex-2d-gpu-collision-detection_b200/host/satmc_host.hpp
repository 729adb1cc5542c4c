// satmc_host.hpp -- C++ convenience layer of the host programs over the C ABI (include/satmc.h).
// No CUDA headers: device memory goes through satmc_device_alloc / satmc_upload / satmc_download.
#pragma once
#include <cmath>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/satmc.h"

namespace satmc_host {

// The reference's PODs (utils.cu:74-109): all packed float32, indices stored as floats.
struct Position { float x, y; };
struct PositionWithVarAndPoseIdx { float x, y, var_idx, pose_idx; };
struct Variance { float x, y, theta, width, height; };
typedef Variance StdDev;
struct Pose { float width, height, theta; };
struct PoseCPVarAndPoseIdx { float x, y, cp, var_idx, pose_idx; };

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Context {
public:
    explicit Context(int device = 0) {
        int rc = satmc_create(device, nullptr, &h_);
        if (rc != SATMC_OK) throw Error(rc, std::string("satmc_create: ") + satmc_last_error(nullptr));
    }
    ~Context() { satmc_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    satmc_ctx* get() const { return h_; }
    void check(int rc, const char* what) const {
        if (rc != SATMC_OK) throw Error(rc, std::string(what) + ": " + satmc_last_error(h_));
    }
private:
    satmc_ctx* h_ = nullptr;
};

template <class T>
class DeviceArray {
public:
    DeviceArray(Context& ctx, size_t n) : ctx_(ctx), n_(n) {
        ctx_.check(satmc_device_alloc(ctx_.get(), reinterpret_cast<void**>(&p_), n * sizeof(T)), "satmc_device_alloc");
    }
    DeviceArray(Context& ctx, const std::vector<T>& host) : DeviceArray(ctx, host.size()) { upload(host.data(), host.size()); }
    ~DeviceArray() { satmc_device_free(ctx_.get(), p_); }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
    T* get() const { return p_; }
    size_t size() const { return n_; }
    void upload(const T* src, size_t n) { ctx_.check(satmc_upload(ctx_.get(), p_, src, n * sizeof(T)), "satmc_upload"); }
    void download(T* dst, size_t n) const { ctx_.check(satmc_download(ctx_.get(), dst, p_, n * sizeof(T)), "satmc_download"); }
    std::vector<T> to_host() const { std::vector<T> v(n_); download(v.data(), n_); return v; }
private:
    Context& ctx_;
    T* p_ = nullptr;
    size_t n_;
};

// std_dev = sqrt(variance), component-wise (generate_dataset.cu:309-317, ztest.cu:245-251)
inline std::vector<StdDev> to_std_devs(const std::vector<Variance>& v) {
    std::vector<StdDev> s(v.size());
    for (size_t i = 0; i < v.size(); i++) {
        s[i].x = std::sqrt(v[i].x); s[i].y = std::sqrt(v[i].y); s[i].theta = std::sqrt(v[i].theta);
        s[i].width = std::sqrt(v[i].width); s[i].height = std::sqrt(v[i].height);
    }
    return s;
}

// create_rect (utils.cu:119-130), host side: the robot base the reference uploads (ztest.cu:297-300)
inline std::vector<float> create_rect(float w, float h) {
    return {-w / 2, -h / 2, w / 2, -h / 2, w / 2, h / 2, -w / 2, h / 2};
}

// Sampling schedule of the adaptive loop.
struct Schedule {
    int n_batch_small, switch_at, n_batch_large, max_samples;
    static Schedule dataset(int max_samples) { return {1000, 20000, 100000, max_samples}; }   // generate_dataset.cu:427-431
    static Schedule ztest(int max_samples) { return {10000, 0, 10000, max_samples}; }         // ztest.cu:332
};

// Tables resident on the device for the lifetime of a run + the adaptive z-test loop over one batch.
class MonteCarlo {
public:
    MonteCarlo(Context& ctx, float robot_w, float robot_h, const std::vector<Pose>& poses, const std::vector<StdDev>& sds,
               const std::vector<float>& accuracy_bins, const std::vector<float>& bin_accuracy)
        : ctx_(ctx), n_poses_((uint32_t)poses.size()), n_std_((uint32_t)sds.size()), n_bins_((int)accuracy_bins.size()),
          d_robot_(ctx, create_rect(robot_w, robot_h)),
          d_poses_(ctx, flatten(poses)), d_sds_(ctx, flatten(sds)), d_bins_(ctx, accuracy_bins),
          d_acc_(ctx, padded(bin_accuracy, accuracy_bins.size())) {
        if (poses.empty() || sds.empty()) throw Error(SATMC_ERR_INVALID, "empty pose or variance table");
        if (accuracy_bins.size() < 2 || bin_accuracy.size() + 1 < accuracy_bins.size())
            throw Error(SATMC_ERR_INVALID, "need n accuracy_bins (n >= 2) and n-1 bin_accuracy values");
    }

    // generate_dataset iteration 0: draws (pose_idx, var_idx, position) for n data points
    void sample_positions(int n, float r_offset, float spread, uint64_t seed, uint32_t stream_offset,
                          DeviceArray<float>& d_pos, DeviceArray<float>& d_pose_idx, DeviceArray<float>& d_var_idx) {
        ctx_.check(satmc_sample_positions(ctx_.get(), d_poses_.get(), n_poses_, d_sds_.get(), n_std_, n, r_offset, spread, seed,
                                          stream_offset, d_pos.get(), d_pose_idx.get(), d_var_idx.get()), "satmc_sample_positions");
    }

    // the whole while-loop of the reference mains; cp_out in input order
    void run(const DeviceArray<float>& d_pos, const DeviceArray<float>& d_pose_idx, const DeviceArray<float>& d_var_idx, int n,
             const Schedule& s, uint64_t seed, uint32_t stream_offset, DeviceArray<float>& d_cp, int* iterations = nullptr,
             long long* samples = nullptr) {
        ctx_.check(satmc_adaptive_run(ctx_.get(), d_robot_.get(), d_poses_.get(), n_poses_, d_sds_.get(), n_std_, d_pose_idx.get(),
                                      d_var_idx.get(), d_pos.get(), n, d_bins_.get(), d_acc_.get(), n_bins_, s.max_samples,
                                      s.n_batch_small, s.switch_at, s.n_batch_large, seed, stream_offset, d_cp.get(), nullptr,
                                      iterations, samples), "satmc_adaptive_run");
    }

private:
    template <class T> static std::vector<float> flatten(const std::vector<T>& v) {
        const float* p = reinterpret_cast<const float*>(v.data());
        return std::vector<float>(p, p + v.size() * (sizeof(T) / sizeof(float)));
    }
    static std::vector<float> padded(std::vector<float> v, size_t n) { v.resize(n > v.size() ? n : v.size(), 0.0f); return v; }

    Context& ctx_;
    uint32_t n_poses_, n_std_;
    int n_bins_;
    DeviceArray<float> d_robot_, d_poses_, d_sds_, d_bins_, d_acc_;
};

// The tables resident on `gpus` devices starting at `first_device` (one context per GPU, created once) and the
// adaptive z-test over n host-side rows sharded across them: one host thread per GPU, contiguous row ranges.  Row i
// always draws Philox stream stream_offset + i, so cp[] does not depend on the number of GPUs.
class ShardedMonteCarlo {
public:
    ShardedMonteCarlo(int first_device, int gpus, float robot_w, float robot_h, const std::vector<Pose>& poses,
                      const std::vector<StdDev>& sds, const std::vector<float>& accuracy_bins, const std::vector<float>& bin_accuracy) {
        if (gpus < 1) gpus = 1;
        for (int w = 0; w < gpus; w++) {
            ctx_.emplace_back(new Context(first_device + w));
            mc_.emplace_back(new MonteCarlo(*ctx_.back(), robot_w, robot_h, poses, sds, accuracy_bins, bin_accuracy));
        }
    }
    ~ShardedMonteCarlo() {
        for (MonteCarlo* m : mc_) delete m;
        for (Context* c : ctx_) delete c;
    }
    ShardedMonteCarlo(const ShardedMonteCarlo&) = delete;
    ShardedMonteCarlo& operator=(const ShardedMonteCarlo&) = delete;

    std::vector<float> run_rows(const std::vector<float>& pos, const std::vector<float>& pose_idx, const std::vector<float>& var_idx,
                                const Schedule& schedule, uint64_t seed, uint32_t stream_offset, int* iterations = nullptr,
                                long long* samples = nullptr) {
        const size_t n = pose_idx.size();
        const size_t gpus = ctx_.size();
        std::vector<float> cp(n);
        std::mutex m;
        std::string failure;
        int max_iter = 0; long long total = 0;
        auto worker = [&](size_t w) {
            try {
                const size_t lo = n * w / gpus, hi = n * (w + 1) / gpus;
                if (hi == lo) return;
                Context& ctx = *ctx_[w];
                const size_t k = hi - lo;
                DeviceArray<float> d_pos(ctx, 2 * k), d_pi(ctx, k), d_vi(ctx, k), d_cp(ctx, k);
                d_pos.upload(pos.data() + 2 * lo, 2 * k); d_pi.upload(pose_idx.data() + lo, k); d_vi.upload(var_idx.data() + lo, k);
                int it = 0; long long smp = 0;
                mc_[w]->run(d_pos, d_pi, d_vi, (int)k, schedule, seed, stream_offset + (uint32_t)lo, d_cp, &it, &smp);
                d_cp.download(cp.data() + lo, k);
                std::lock_guard<std::mutex> lock(m);
                if (it > max_iter) max_iter = it;
                total += smp;
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lock(m);
                if (failure.empty()) failure = std::string("GPU shard ") + std::to_string(w) + ": " + e.what();
            }
        };
        std::vector<std::thread> threads;
        for (size_t w = 1; w < gpus; w++) threads.emplace_back(worker, w);
        worker(0);
        for (std::thread& t : threads) t.join();
        if (!failure.empty()) throw Error(SATMC_ERR_CUDA, failure);
        if (iterations) *iterations = max_iter;
        if (samples) *samples = total;
        return cp;
    }

private:
    std::vector<Context*> ctx_;
    std::vector<MonteCarlo*> mc_;
};

}  // namespace satmc_host
