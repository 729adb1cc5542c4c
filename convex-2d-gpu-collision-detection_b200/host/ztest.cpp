// ztest -- drop-in for the reference program of the same name (ztest.cu:168-444): reads
// <data_dir>/poses.npy, variances.npy, meta/*.npy and an [N,4] file of (x, y, var_idx, pose_idx) rows,
// estimates the collision probability of every row with the adaptive z-test loop (10 000 samples per
// iteration, ztest.cu:332) and writes [N,5] = (x, y, cp, var_idx, pose_idx) or, with --cps_only 1, [N] cp.
// Flags as upstream (ztest.cu:49-63) plus --seed / --device / --gpus (rows sharded over GPUs; same output for any count).
// Upstream quirks kept in effect, not in mechanism: the output is in input order -- upstream's shuffle
// branches are inverted (ztest.cu:408-414 shuffle the array that is not written), so --shuffle never
// changes the file; --meta_dir only suppresses writing the default meta files, the bins are always read
// from <data_dir>/meta (ztest.cu:186-194,221-222).
#include <chrono>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <iostream>

#include "cli.hpp"
#include "npy.hpp"
#include "satmc_host.hpp"

using namespace satmc_host;
namespace fs = std::filesystem;

struct Arguments {
    std::string data_dir = "./data/", data_file_in = "", data_file_out = "", meta_dir = "";
    int max_samples = 4000000;
    float robot_width = 4.07f, robot_height = 1.74f;
    bool shuffle = true, cps_only = false;
    uint64_t seed = 0;
    bool has_seed = false;
    int device = 0, gpus = 1;
};

static Arguments parse_args(int argc, char** argv) {
    using cli::Kind;
    Arguments a;
    cli::Parser p("Allowed options");
    p.add("help", Kind::Switch, "produce help message")
     .add("data_dir", Kind::String, "where to read the data")
     .add("data_file_in", Kind::String, "where to read the data")
     .add("data_file_out", Kind::String, "where to write the data")
     .add("max_samples", Kind::Int, "maximum number of samples for z-test")
     .add("robot_width", Kind::Float, "robot width", 'w')
     .add("robot_height", Kind::Float, "robot height", 'h')
     .add("shuffle", Kind::Bool, "whether or not to shuffle data")
     .add("cps_only", Kind::Bool, "whether or not to only compute collision probabilities")
     .add("meta_dir", Kind::String, "path to meta folder containing accuracy_bins.npy and bin_accuracy.npy")
     .add("seed", Kind::String, "RNG seed (default: from the clock, as upstream)")
     .add("device", Kind::Int, "CUDA device index (first device when --gpus > 1)")
     .add("gpus", Kind::Int, "number of GPUs to shard the rows over");
    p.parse(argc, argv);
    if (p.count("help")) { p.print_help(std::cout); std::cout << "\n"; exit(1); }
    if (p.count("data_dir")) a.data_dir = p.str("data_dir");
    if (p.count("data_file_in")) a.data_file_in = p.str("data_file_in");
    if (p.count("data_file_out")) a.data_file_out = p.str("data_file_out");
    if (p.count("max_samples")) a.max_samples = p.integer("max_samples");
    if (p.count("robot_width")) a.robot_width = p.real("robot_width");
    if (p.count("robot_height")) a.robot_height = p.real("robot_height");
    if (p.count("shuffle")) a.shuffle = p.boolean("shuffle");
    if (p.count("cps_only")) a.cps_only = p.boolean("cps_only");
    if (p.count("meta_dir")) a.meta_dir = p.str("meta_dir");
    if (p.count("seed")) { a.seed = p.unsigned64("seed"); a.has_seed = true; }
    if (p.count("device")) a.device = p.integer("device");
    if (p.count("gpus")) a.gpus = p.integer("gpus");
    return a;
}

template <class T> static std::vector<T> load_rows(const fs::path& file, size_t cols) {
    npyio::Array a = npyio::load_f32(file.string());
    if (a.data.size() % cols) throw std::runtime_error(file.string() + ": size is not a multiple of " + std::to_string(cols));
    std::vector<T> v(a.data.size() / cols);
    std::memcpy(v.data(), a.data.data(), a.data.size() * sizeof(float));
    return v;
}

int main(int argc, char* argv[]) try {
    Arguments args = parse_args(argc, argv);
    fs::path data_dir = args.data_dir, data_file_in = args.data_file_in, data_file_out = args.data_file_out;
    if (!fs::exists(data_dir)) { std::cout << "Error: data_dir " << data_dir << " does not exist." << std::endl; return 1; }
    if (args.meta_dir.empty()) {                                                     // ztest.cu:186-194
        fs::create_directory(data_dir / "meta");
        npyio::save_f32((data_dir / "meta/accuracy_bins.npy").string(), {4}, std::vector<float>{0.0f, 0.01f, 0.1f, 1.0f});
        npyio::save_f32((data_dir / "meta/bin_accuracy.npy").string(), {3}, std::vector<float>{0.0001f, 0.001f, 0.01f});
    }
    if (data_file_in.empty()) {
        fs::create_directories(data_dir / "tmp");
        data_file_in = data_dir / "tmp/0.npy";
        std::cout << "Using default input file: " << data_file_in << std::endl;
    }
    if (data_file_out.empty()) {
        data_file_out = data_dir / "0.npy";
        std::cout << "Using default output file: " << data_file_out << std::endl;
    }
    if (fs::exists(data_file_out)) std::cout << "Warning: " << data_file_out << " already exists, will be overwritten" << std::endl;
    for (const char* f : {"poses.npy", "variances.npy"})
        if (!fs::exists(data_dir / f)) { std::cout << "Error: " << data_dir / f << " does not exist." << std::endl; return 1; }

    std::cout << "Reading data..." << std::endl;
    std::vector<Pose> poses; std::vector<Variance> variances; std::vector<PositionWithVarAndPoseIdx> rows;
    std::vector<float> accuracy_bins, bin_accuracy;
    try {
        poses = load_rows<Pose>(data_dir / "poses.npy", 3);
        variances = load_rows<Variance>(data_dir / "variances.npy", 5);
        rows = load_rows<PositionWithVarAndPoseIdx>(data_file_in, 4);
        accuracy_bins = npyio::load_f32((data_dir / "meta/accuracy_bins.npy").string()).data;
        bin_accuracy = npyio::load_f32((data_dir / "meta/bin_accuracy.npy").string()).data;
    } catch (const std::exception& e) {
        std::cout << "Error while reading numpy arrays" << std::endl << e.what() << std::endl;
        return 1;
    }
    const int n = (int)rows.size();
    std::cout << "num poses: " << poses.size() << std::endl;
    std::cout << "num variances: " << variances.size() << std::endl;
    std::cout << "num data points: " << n << std::endl;

    std::vector<float> pos(2 * (size_t)n), var_idx(n), pose_idx(n);
    for (int i = 0; i < n; i++) {                                                     // ztest.cu:262-268
        pos[2 * i] = rows[i].x; pos[2 * i + 1] = rows[i].y; var_idx[i] = rows[i].var_idx; pose_idx[i] = rows[i].pose_idx;
    }
    const uint64_t seed = args.has_seed ? args.seed : (uint64_t)std::time(nullptr);

    auto begin = std::chrono::steady_clock::now();
    std::cout << "Total number of configurations: " << n << std::endl;
    std::cout << "Begin computation..." << std::endl;
    int iterations = 0; long long samples = 0;
    ShardedMonteCarlo mc(args.device, args.gpus, args.robot_width, args.robot_height, poses, to_std_devs(variances), accuracy_bins,
                         bin_accuracy);
    std::vector<float> cp = mc.run_rows(pos, pose_idx, var_idx, Schedule::ztest(args.max_samples), seed, 0, &iterations, &samples);

    if (args.cps_only) {
        npyio::save_f32(data_file_out.string(), {(size_t)n}, cp);                    // ztest.cu:418-420
    } else {
        std::vector<PoseCPVarAndPoseIdx> dataset(n);
        for (int j = 0; j < n; j++) dataset[j] = {rows[j].x, rows[j].y, cp[j], rows[j].var_idx, rows[j].pose_idx};
        npyio::save_f32(data_file_out.string(), {(size_t)n, 5}, reinterpret_cast<const float*>(dataset.data()));
    }
    auto end = std::chrono::steady_clock::now();
    std::cout << "Finished computation (" << iterations << " iterations, " << samples << " samples)" << std::endl;
    std::cout << "Elapsed time: " << std::chrono::duration_cast<std::chrono::minutes>(end - begin).count() << " [min]" << std::endl;
    std::cout << "Done." << std::endl;
    return 0;
} catch (const std::exception& e) {
    std::cerr << "ztest: " << e.what() << std::endl;
    return 2;
}
