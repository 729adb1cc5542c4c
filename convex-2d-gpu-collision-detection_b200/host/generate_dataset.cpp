// generate_dataset -- drop-in for the reference program of the same name (generate_dataset.cu:255-524):
// same flags (generate_dataset.cu:66-89, plus argparser.h's --variance_dim/--pose_dim), same files
// (SURVEY.md appendix C):  <data_dir>/poses.npy [P,3], variances.npy [V,5], meta/accuracy_bins.npy,
// meta/bin_accuracy.npy, and one <start_batch_count + b>.npy [B,5] = (x, y, cp, var_idx, pose_idx) per batch,
// rows shuffled with std::default_random_engine(0) as upstream (:496).
// The Monte Carlo work goes through the satmc C ABI (no CUDA code here).  Differences from upstream, all
// deliberate: the data directory is created before the first file is written (upstream writes
// variances.npy first and fails if the directory is missing, :303 vs :343-345); --seed makes runs
// reproducible (upstream seeds from time(0), :406); --device selects the GPU; --gpus N spreads the batches
// round-robin over N GPUs (one host thread and one satmc context per GPU; a batch's random streams depend
// only on its number, so the files do not depend on N).
#include <sys/stat.h>

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <atomic>
#include <iostream>
#include <mutex>
#include <random>
#include <thread>

#include "cli.hpp"
#include "npy.hpp"
#include "satmc_host.hpp"

using namespace satmc_host;

struct Arguments {
    std::string data_dir = "./data/", pose_dir = "", variance_dir = "";
    int num_batches = 100, batch_size = 100000, start_batch_count = 0;
    int num_poses = 64 * 64 * 64 * 64, num_variances = 64 * 64 * 64 * 64, max_samples = 4000000;
    std::vector<float> min_variance = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, max_variance = {0.3f, 0.3f, 0.3f, 0.3f, 0.3f};
    std::vector<float> min_pose = {0.1f, 0.1f, 0.0f}, max_pose = {5, 5, (float)(2 * M_PI)};
    std::vector<float> accuracy_bins = {0.0f, 0.01f, 0.1f, 1.0f}, bin_accuracy = {0.0001f, 0.001f, 0.01f};
    float robot_width = 4.07f, robot_height = 1.74f, spread = 4;
    bool shape_variance = false;
    long long seed = -1;
    int device = 0, gpus = 1;
};

static Arguments parse_args(int argc, char** argv) {
    using cli::Kind;
    Arguments a;
    cli::Parser p("Allowed options");
    p.add("help", Kind::Switch, "produce help message")
     .add("data_dir", Kind::String, "where to store the data")
     .add("num_batches", Kind::Int, "number of batches", 'n')
     .add("batch_size", Kind::Int, "number of samples per batch", 'b')
     .add("start_batch_count", Kind::Int, "start value for batches", 's')
     .add("num_poses", Kind::Int, "number of poses")
     .add("num_variances", Kind::Int, "number of variances")
     .add("variance_dim", Kind::Int, "dimension of variance (accepted for argparser.h compatibility; always 5)")
     .add("pose_dim", Kind::Int, "dimension of pose (accepted for argparser.h compatibility; always 3)")
     .add("shape_variance", Kind::Switch, "whether or not to have shape variance")
     .add("max_samples", Kind::Int, "maximum number of samples for z-test")
     .add("accuracy_bins", Kind::FloatList, "accuracy bins e.g. [0.0001 0.001 0.01 0]")
     .add("bin_accuracy", Kind::FloatList, "accuracy for each bin e.g. [0.0001, 0.001, 0.01]")
     .add("min_variance", Kind::FloatList, "min variance for each dimension e.g. [0.0, 0.0, 0.0, 0.0, 0.0]")
     .add("max_variance", Kind::FloatList, "max variance for each dimension e.g. [0.3, 0.3, 0.3, 0.3, 0.3]")
     .add("min_pose", Kind::FloatList, "min pose for each dimension e.g. [0.1, 0.1, 0.0]")
     .add("max_pose", Kind::FloatList, "max pose for each dimension e.g. [5, 5, 2*M_PI]")
     .add("robot_width", Kind::Float, "robot width", 'w')
     .add("robot_height", Kind::Float, "robot height", 'h')
     .add("spread", Kind::Float, "spread of poses")
     .add("pose_dir", Kind::String, "path of a poses .npy file to load instead of sampling")
     .add("variance_dir", Kind::String, "path of a variances .npy file to load instead of sampling")
     .add("seed", Kind::Int, "RNG seed (default: from the clock, as upstream)")
     .add("device", Kind::Int, "CUDA device index (first device when --gpus > 1)")
     .add("gpus", Kind::Int, "number of GPUs to spread the batches over");
    p.parse(argc, argv);
    if (p.count("help")) { p.print_help(std::cout); std::cout << "\n"; exit(1); }
    if (p.count("data_dir")) a.data_dir = p.str("data_dir");
    if (p.count("num_batches")) a.num_batches = p.integer("num_batches");
    if (p.count("batch_size")) a.batch_size = p.integer("batch_size");
    if (p.count("start_batch_count")) a.start_batch_count = p.integer("start_batch_count");
    if (p.count("num_poses")) a.num_poses = p.integer("num_poses");
    if (p.count("num_variances")) a.num_variances = p.integer("num_variances");
    if (p.count("max_samples")) a.max_samples = p.integer("max_samples");
    if (p.count("accuracy_bins")) a.accuracy_bins = p.reals("accuracy_bins");
    if (p.count("bin_accuracy")) a.bin_accuracy = p.reals("bin_accuracy");
    auto fixed = [&](const char* name, std::vector<float>& dst, size_t n) {
        if (!p.count(name)) return;
        std::vector<float> v = p.reals(name);
        if (v.size() != n) throw std::runtime_error(std::string("--") + name + " needs exactly " + std::to_string(n) + " values");
        dst = v;
    };
    fixed("min_variance", a.min_variance, 5); fixed("max_variance", a.max_variance, 5);
    fixed("min_pose", a.min_pose, 3); fixed("max_pose", a.max_pose, 3);
    if (p.count("robot_width")) a.robot_width = p.real("robot_width");
    if (p.count("robot_height")) a.robot_height = p.real("robot_height");
    if (p.count("spread")) a.spread = p.real("spread");
    if (p.count("shape_variance")) a.shape_variance = true;
    if (p.count("pose_dir")) a.pose_dir = p.str("pose_dir");
    if (p.count("variance_dir")) a.variance_dir = p.str("variance_dir");
    if (p.count("seed")) a.seed = p.integer("seed");
    if (p.count("device")) a.device = p.integer("device");
    if (p.count("gpus")) a.gpus = p.integer("gpus");
    if (a.gpus < 1) throw std::runtime_error("--gpus must be >= 1");
    return a;
}

static void make_dir(const std::string& d) {
    struct stat st;
    if (stat(d.c_str(), &st) == -1) mkdir(d.c_str(), 0700);
}

template <class T> static std::vector<T> load_rows(const std::string& file, size_t cols) {
    npyio::Array a = npyio::load_f32(file);
    if (a.data.size() % cols) throw std::runtime_error(file + ": size is not a multiple of " + std::to_string(cols));
    std::vector<T> v(a.data.size() / cols);
    std::memcpy(v.data(), a.data.data(), a.data.size() * sizeof(float));
    return v;
}

int main(int argc, char* argv[]) try {
    Arguments args = parse_args(argc, argv);
    const std::string data_dir = args.data_dir;
    std::cout << "data dir: " << data_dir << std::endl;
    std::cout << "num batches: " << args.num_batches << std::endl;
    std::cout << "num batch: " << args.batch_size << std::endl;
    std::cout << "start batch count: " << args.start_batch_count << std::endl;
    make_dir(data_dir);
    make_dir(data_dir + "/meta");

    // pose / variance tables: uniform in [min, max] per dimension from one default-constructed engine,
    // variances first (generate_dataset.cu:279-303,319-332) -- the same draws as upstream
    std::default_random_engine generator;
    std::vector<Variance> variances;
    std::vector<Pose> poses;
    if (args.variance_dir.empty()) {
        if (!args.shape_variance) {
            args.min_variance[3] = args.max_variance[3] = 0.0f;
            args.min_variance[4] = args.max_variance[4] = 0.0f;
        }
        std::vector<std::uniform_real_distribution<float>> u;
        for (int i = 0; i < 5; i++) u.emplace_back(args.min_variance[i], args.max_variance[i]);
        variances.resize(args.num_variances);
        for (Variance& v : variances) {
            v.x = u[0](generator); v.y = u[1](generator); v.theta = u[2](generator);
            v.width = u[3](generator); v.height = u[4](generator);
        }
        npyio::save_f32(data_dir + "/variances.npy", {variances.size(), 5}, reinterpret_cast<const float*>(variances.data()));
    } else {
        variances = load_rows<Variance>(args.variance_dir, 5);
    }
    if (args.pose_dir.empty()) {
        std::vector<std::uniform_real_distribution<float>> u;
        for (int i = 0; i < 3; i++) u.emplace_back(args.min_pose[i], args.max_pose[i]);
        poses.resize(args.num_poses);
        for (Pose& q : poses) { q.width = u[0](generator); q.height = u[1](generator); q.theta = u[2](generator); }
        npyio::save_f32(data_dir + "/poses.npy", {poses.size(), 3}, reinterpret_cast<const float*>(poses.data()));
    } else {
        poses = load_rows<Pose>(args.pose_dir, 3);
    }
    std::cout << "num poses: " << poses.size() << std::endl;
    std::cout << "num variances: " << variances.size() << std::endl;
    npyio::save_f32(data_dir + "/meta/accuracy_bins.npy", {args.accuracy_bins.size()}, args.accuracy_bins);
    npyio::save_f32(data_dir + "/meta/bin_accuracy.npy", {args.bin_accuracy.size()}, args.bin_accuracy);

    const std::vector<StdDev> std_devs = to_std_devs(variances);
    const int B = args.batch_size;
    const float r_offset = (args.robot_width + args.robot_height) / 4;                    // :398
    const uint64_t seed = args.seed >= 0 ? (uint64_t)args.seed : (uint64_t)std::time(nullptr);

    auto begin = std::chrono::steady_clock::now();
    std::cout << "Total number of configurations: " << (long long)B * args.num_batches << std::endl;
    std::cout << "Begin computation..." << std::endl;
    std::atomic<int> counter{0};
    std::mutex io;
    std::string failure;
    printf("batches generated: %i/%i", 0, args.num_batches);
    fflush(stdout);

    // one worker per GPU; worker w takes batches w, w + gpus, w + 2 gpus, ...
    auto worker = [&](int w) {
        try {
            Context ctx(args.device + w);
            MonteCarlo mc(ctx, args.robot_width, args.robot_height, poses, std_devs, args.accuracy_bins, args.bin_accuracy);
            DeviceArray<float> d_pos(ctx, 2 * (size_t)B), d_pose_idx(ctx, B), d_var_idx(ctx, B), d_cp(ctx, B);
            std::vector<PoseCPVarAndPoseIdx> dataset(B);
            for (int b = w; b < args.num_batches; b += args.gpus) {
                const uint32_t stream = (uint32_t)(((uint64_t)(args.start_batch_count + b) * (uint64_t)B) & 0xffffffffu);
                mc.sample_positions(B, r_offset, args.spread, seed, stream, d_pos, d_pose_idx, d_var_idx);
                mc.run(d_pos, d_pose_idx, d_var_idx, B, Schedule::dataset(args.max_samples), seed, stream, d_cp);
                std::vector<float> pos = d_pos.to_host(), pi = d_pose_idx.to_host(), vi = d_var_idx.to_host(), cp = d_cp.to_host();
                for (int j = 0; j < B; j++) dataset[j] = {pos[2 * j], pos[2 * j + 1], cp[j], vi[j], pi[j]};   // :485-494
                std::shuffle(dataset.begin(), dataset.end(), std::default_random_engine(0));                  // :496
                npyio::save_f32(data_dir + "/" + std::to_string(args.start_batch_count + b) + ".npy", {(size_t)B, 5},
                                reinterpret_cast<const float*>(dataset.data()));
                const int done = ++counter;
                std::lock_guard<std::mutex> lock(io);
                auto now = std::chrono::steady_clock::now();
                printf("\33[2K\r");
                printf("batches generated: %i/%i, Time: %i [min]", done, args.num_batches,
                       (int)std::chrono::duration_cast<std::chrono::minutes>(now - begin).count());
                fflush(stdout);
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lock(io);
            if (failure.empty()) failure = std::string("GPU ") + std::to_string(args.device + w) + ": " + e.what();
        }
    };
    std::vector<std::thread> threads;
    for (int w = 1; w < args.gpus; w++) threads.emplace_back(worker, w);
    worker(0);
    for (std::thread& t : threads) t.join();
    if (!failure.empty()) throw std::runtime_error(failure);
    std::cout << std::endl;
    auto end = std::chrono::steady_clock::now();
    std::cout << "Finished computation" << std::endl;
    std::cout << "Elapsed time: " << std::chrono::duration_cast<std::chrono::minutes>(end - begin).count() << " [min]" << std::endl;
    std::cout << "Done." << std::endl;
    return 0;
} catch (const std::exception& e) {
    std::cerr << "generate_dataset: " << e.what() << std::endl;
    return 2;
}
