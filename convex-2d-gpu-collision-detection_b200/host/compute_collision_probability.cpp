// compute_collision_probability -- drop-in for the reference program of the same name
// (compute_collision_probability.cu:152-379): for every numbered <data_in>/<b>.npy ([N,4] rows of
// (x, y, var_idx, pose_idx)) estimate the collision probabilities with the generate_dataset schedule
// (1000 samples per iteration up to 20 000, then 100 000, :281-286) and write
// <data_out>/<start + b>.npy ([N,5] = (x, y, cp, var_idx, pose_idx)), start = the number of numbered
// files already in data_out (:157).  Tables and meta are read from data_out, as upstream (:162-166).
// Flags as upstream (:44-56) plus --seed / --device / --gpus; rows are written in input order and then shuffled
// with std::default_random_engine(0) when --shuffle is true (:346-349).
// Differences from upstream: every batch file may have its own number of rows (upstream sizes all
// buffers from 0.npy, :164,174); upstream's seed is glibc's first rand() (srand commented out, :249).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <future>
#include <iostream>
#include <random>

#include "cli.hpp"
#include "npy.hpp"
#include "satmc_host.hpp"

using namespace satmc_host;
namespace fs = std::filesystem;

struct Arguments {
    std::string data_in = "./data_in/", data_out = "./data_out/";
    int max_samples = 4000000;
    float robot_width = 4.07f, robot_height = 1.74f;
    bool shuffle = true;
    uint64_t seed = 0;
    bool has_seed = false, stats = false;
    int device = 0, gpus = 1;
};

static Arguments parse_args(int argc, char** argv) {
    using cli::Kind;
    Arguments a;
    cli::Parser p("Allowed options");
    p.add("help", Kind::Switch, "produce help message")
     .add("data_in", Kind::String, "where to read the data")
     .add("data_out", Kind::String, "where to write the data")
     .add("max_samples", Kind::Int, "maximum number of samples for z-test")
     .add("robot_width", Kind::Float, "robot width", 'w')
     .add("robot_height", Kind::Float, "robot height", 'h')
     .add("shuffle", Kind::Bool, "whether or not to shuffle data")
     .add("seed", Kind::String, "RNG seed (default: 1804289383, glibc's first rand(), as upstream)")
     .add("stats", Kind::Switch, "print a one-line timing summary at the end")
     .add("device", Kind::Int, "CUDA device index (first device when --gpus > 1)")
     .add("gpus", Kind::Int, "number of GPUs to shard the rows of each file over");
    p.parse(argc, argv);
    if (p.count("help")) { p.print_help(std::cout); std::cout << "\n"; exit(1); }
    if (p.count("data_in")) a.data_in = p.str("data_in");
    if (p.count("data_out")) a.data_out = p.str("data_out");
    if (p.count("max_samples")) a.max_samples = p.integer("max_samples");
    if (p.count("robot_width")) a.robot_width = p.real("robot_width");
    if (p.count("robot_height")) a.robot_height = p.real("robot_height");
    if (p.count("shuffle")) a.shuffle = p.boolean("shuffle");
    if (p.count("seed")) { a.seed = p.unsigned64("seed"); a.has_seed = true; }
    if (p.count("stats")) a.stats = true;
    if (p.count("device")) a.device = p.integer("device");
    if (p.count("gpus")) a.gpus = p.integer("gpus");
    return a;
}

// number of regular files "<int>.npy" in a directory (get_num_batches_in_dir, utils.cu:36-56)
static int num_batches_in_dir(const std::string& dir) {
    int n = 0;
    for (const auto& e : fs::directory_iterator(dir)) {
        if (!fs::is_regular_file(e) || e.path().extension() != ".npy") continue;
        try { (void)std::stoi(e.path().filename().string()); n++; } catch (...) {}
    }
    return n;
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
    const std::string data_in = args.data_in, data_out = args.data_out;
    const int start_batch_count = num_batches_in_dir(data_out);
    const int num_batches = num_batches_in_dir(data_in);
    std::cout << "Reading data..." << std::endl;
    std::vector<Pose> poses = load_rows<Pose>(data_out + "/poses.npy", 3);
    std::vector<Variance> variances = load_rows<Variance>(data_out + "/variances.npy", 5);
    std::vector<float> accuracy_bins = npyio::load_f32(data_out + "/meta/accuracy_bins.npy").data;
    std::vector<float> bin_accuracy = npyio::load_f32(data_out + "/meta/bin_accuracy.npy").data;
    std::cout << "num poses: " << poses.size() << std::endl;
    std::cout << "num variances: " << variances.size() << std::endl;

    ShardedMonteCarlo mc(args.device, args.gpus, args.robot_width, args.robot_height, poses, to_std_devs(variances), accuracy_bins,
                         bin_accuracy);
    const uint64_t seed = args.has_seed ? args.seed : 1804289383ull;
    auto begin = std::chrono::steady_clock::now();
    std::cout << "Begin computation..." << std::endl;
    int counter = 0;
    printf("batches generated: %i/%i\n", counter, num_batches);
    // Three stages in flight: file b+1 is read and unpacked, and file b-1 assembled, shuffled and written, by helper
    // threads while the GPUs run the adaptive loop of file b (upstream does the three in sequence, :259-360).
    struct Input {
        std::vector<PositionWithVarAndPoseIdx> rows;
        std::vector<float> pos, var_idx, pose_idx;
    };
    auto load = [&](int b) {
        Input in;
        in.rows = load_rows<PositionWithVarAndPoseIdx>(data_in + "/" + std::to_string(b) + ".npy", 4);
        const size_t n = in.rows.size();
        in.pos.resize(2 * n); in.var_idx.resize(n); in.pose_idx.resize(n);
        for (size_t i = 0; i < n; i++) {                                              // :262-268
            in.pos[2 * i] = in.rows[i].x; in.pos[2 * i + 1] = in.rows[i].y;
            in.var_idx[i] = in.rows[i].var_idx; in.pose_idx[i] = in.rows[i].pose_idx;
        }
        return in;
    };
    auto store = [&](int b, std::vector<PositionWithVarAndPoseIdx> rows, std::vector<float> cp) {
        const size_t n = rows.size();
        std::vector<PoseCPVarAndPoseIdx> dataset(n);
        for (size_t j = 0; j < n; j++) dataset[j] = {rows[j].x, rows[j].y, cp[j], rows[j].var_idx, rows[j].pose_idx};   // :337-344
        if (args.shuffle) std::shuffle(dataset.begin(), dataset.end(), std::default_random_engine(0));                  // :346-349
        npyio::save_f32(data_out + "/" + std::to_string(start_batch_count + b) + ".npy", {n, 5},
                        reinterpret_cast<const float*>(dataset.data()));
    };
    StreamCursor cursor;                                 // 64-bit running row index: 32-bit stream id + epoch in the seed
    double gpu_s = 0;
    std::future<Input> next;
    std::future<void> pending_write;
    if (num_batches > 0) next = std::async(std::launch::async, load, 0);
    for (int b = 0; b < num_batches; b++) {
        Input in = next.get();
        if (b + 1 < num_batches) next = std::async(std::launch::async, load, b + 1);
        const int n = (int)in.rows.size();
        if (b == 0) std::cout << "num data points: " << n << std::endl;
        const uint32_t first_stream = cursor.take((uint64_t)n);
        const auto t_gpu = std::chrono::steady_clock::now();
        std::vector<float> cp = mc.run_rows(in.pos, in.pose_idx, in.var_idx, Schedule::dataset(args.max_samples), cursor.seed(seed),
                                            first_stream);
        gpu_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_gpu).count();
        if (pending_write.valid()) pending_write.get();                                // at most one file behind
        pending_write = std::async(std::launch::async, store, b, std::move(in.rows), std::move(cp));
        auto end = std::chrono::steady_clock::now();
        printf("\33[2K\r");
        printf("batches generated: %i/%i, Time: %i [min]", ++counter, num_batches,
               (int)std::chrono::duration_cast<std::chrono::minutes>(end - begin).count());
        fflush(stdout);
    }
    if (pending_write.valid()) pending_write.get();
    if (args.stats) {
        const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - begin).count();
        printf("\nstats: {\"files\": %d, \"total_s\": %.3f, \"gpu_loop_busy_s\": %.3f, \"gpu_busy_frac\": %.3f, \"gpus\": %d}", num_batches,
               total, gpu_s, total > 0 ? gpu_s / total : 0.0, args.gpus);
    }
    std::cout << std::endl;
    auto end = std::chrono::steady_clock::now();
    std::cout << "Finished computation" << std::endl;
    std::cout << "Elapsed time: " << std::chrono::duration_cast<std::chrono::minutes>(end - begin).count() << " [min]" << std::endl;
    std::cout << "Done." << std::endl;
    return 0;
} catch (const std::exception& e) {
    std::cerr << "compute_collision_probability: " << e.what() << std::endl;
    return 2;
}
