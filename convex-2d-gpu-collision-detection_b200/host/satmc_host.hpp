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
    // enqueue only (pinned destination); complete after satmc_synchronize
    void download_async(T* dst, size_t n) const { ctx_.check(satmc_download_async(ctx_.get(), dst, p_, n * sizeof(T)), "satmc_download_async"); }
    std::vector<T> to_host() const { std::vector<T> v(n_); download(v.data(), n_); return v; }
private:
    Context& ctx_;
    T* p_ = nullptr;
    size_t n_;
};

// std_dev = sqrt(variance), component-wise (generate_dataset.cu:309-317, ztest.cu:245-251)
inline std::vector<StdDev> to_std_devs(const std::vector<Variance>& v) {
    std::vector<StdDev> s(v.size());
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            s[i].x = std::sqrt(v[i].x); s[i].y = std::sqrt(v[i].y); s[i].theta = std::sqrt(v[i].theta);
            s[i].width = std::sqrt(v[i].width); s[i].height = std::sqrt(v[i].height);
        }
    };
    unsigned threads = v.size() >= (1u << 16) ? std::thread::hardware_concurrency() : 1;
    if (threads == 0) threads = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; t++) pool.emplace_back(work, v.size() * t / threads, v.size() * (t + 1) / threads);
    work(0, v.size() / threads);
    for (std::thread& t : pool) t.join();
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
        : ctx_(validated(ctx, poses.size(), sds.size(), accuracy_bins.size(), bin_accuracy.size())),
          n_poses_((uint32_t)poses.size()), n_std_((uint32_t)sds.size()), n_bins_((int)accuracy_bins.size()),
          d_robot_(ctx, create_rect(robot_w, robot_h)),
          d_poses_(ctx, flatten(poses)), d_sds_(ctx, flatten(sds)), d_bins_(ctx, accuracy_bins),
          d_acc_(ctx, padded(bin_accuracy, accuracy_bins.size())) {}

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
    // argument checks run before the first member allocates or uploads anything
    static Context& validated(Context& ctx, size_t n_poses, size_t n_sds, size_t n_bins, size_t n_acc) {
        if (n_poses == 0 || n_sds == 0) throw Error(SATMC_ERR_INVALID, "empty pose or variance table");
        if (n_bins < 2 || n_acc + 1 < n_bins) throw Error(SATMC_ERR_INVALID, "need n accuracy_bins (n >= 2) and n-1 bin_accuracy values");
        return ctx;
    }
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

// Pinned host memory (satmc_host_alloc): download target of the double-buffered batch loop.
template <class T>
class PinnedArray {
public:
    explicit PinnedArray(size_t n) : n_(n) {
        if (satmc_host_alloc(reinterpret_cast<void**>(&p_), n * sizeof(T)) != SATMC_OK)
            throw Error(SATMC_ERR_NOMEM, std::string("satmc_host_alloc: ") + satmc_last_error(nullptr));
    }
    ~PinnedArray() { satmc_host_free(p_); }
    PinnedArray(const PinnedArray&) = delete;
    PinnedArray& operator=(const PinnedArray&) = delete;
    T* get() const { return p_; }
    size_t size() const { return n_; }
    const T& operator[](size_t i) const { return p_[i]; }
private:
    T* p_ = nullptr;
    size_t n_;
};

// The adaptive z-test over host rows on `gpus` devices starting at `first_device`, through the library's group entry
// points (satmc_group_*): the tables are made resident on every device once, rows are dealt round-robin to the GPUs
// (work per row varies ~400x, generate_dataset.cu:53,427-431), the devices advance in lockstep from this thread.
// Row i always draws Philox stream stream_offset + i, so cp[] does not depend on the number of GPUs.
class ShardedMonteCarlo {
public:
    ShardedMonteCarlo(int first_device, int gpus, float robot_w, float robot_h, const std::vector<Pose>& poses,
                      const std::vector<StdDev>& sds, const std::vector<float>& accuracy_bins, const std::vector<float>& bin_accuracy) {
        if (gpus < 1) gpus = 1;
        if (poses.empty() || sds.empty()) throw Error(SATMC_ERR_INVALID, "empty pose or variance table");
        if (accuracy_bins.size() < 2 || bin_accuracy.size() + 1 < accuracy_bins.size())
            throw Error(SATMC_ERR_INVALID, "need n accuracy_bins (n >= 2) and n-1 bin_accuracy values");
        std::vector<int> devices;
        for (int w = 0; w < gpus; w++) devices.push_back(first_device + w);
        int rc = satmc_group_create(devices.data(), gpus, &g_);
        if (rc != SATMC_OK) throw Error(rc, std::string("satmc_group_create: ") + satmc_group_last_error(nullptr));
        const std::vector<float> robot = create_rect(robot_w, robot_h);
        rc = satmc_group_set_tables(g_, robot.data(), reinterpret_cast<const float*>(poses.data()), (uint32_t)poses.size(),
                                    reinterpret_cast<const float*>(sds.data()), (uint32_t)sds.size(), accuracy_bins.data(),
                                    bin_accuracy.data(), (int)accuracy_bins.size());
        if (rc != SATMC_OK) {
            const std::string msg = std::string("satmc_group_set_tables: ") + satmc_group_last_error(g_);
            satmc_group_destroy(g_);
            throw Error(rc, msg);
        }
    }
    ~ShardedMonteCarlo() { satmc_group_destroy(g_); }
    ShardedMonteCarlo(const ShardedMonteCarlo&) = delete;
    ShardedMonteCarlo& operator=(const ShardedMonteCarlo&) = delete;

    std::vector<float> run_rows(const std::vector<float>& pos, const std::vector<float>& pose_idx, const std::vector<float>& var_idx,
                                const Schedule& schedule, uint64_t seed, uint32_t stream_offset, int* iterations = nullptr,
                                long long* samples = nullptr) {
        std::vector<float> cp(pose_idx.size());
        const int rc = satmc_group_adaptive_run_host(g_, pose_idx.data(), var_idx.data(), pos.data(), (int)pose_idx.size(),
                                                     schedule.max_samples, schedule.n_batch_small, schedule.switch_at,
                                                     schedule.n_batch_large, seed, stream_offset, cp.data(), iterations, samples);
        if (rc != SATMC_OK) throw Error(rc, std::string("satmc_group_adaptive_run_host: ") + satmc_group_last_error(g_));
        return cp;
    }

private:
    satmc_group* g_ = nullptr;
};

// Philox stream ids are 32 bits.  Programs that number their rows with a 64-bit running index fold the high part into
// the seed ("epoch") instead of letting ids wrap: rows 2^32 apart would otherwise share a stream at the same sample
// indices and produce perfectly correlated estimates.  A block of rows never straddles an epoch.
struct StreamCursor {
    uint64_t epoch = 0, offset = 0;
    // reserves n consecutive ids; returns the first
    uint32_t take(uint64_t n) {
        if (offset + n > 0x100000000ull) { epoch++; offset = 0; }
        const uint32_t first = (uint32_t)offset;
        offset += n;
        return first;
    }
    uint64_t seed(uint64_t base) const { return base ^ (epoch * 0x9E3779B97F4A7C15ull); }
};

}  // namespace satmc_host
