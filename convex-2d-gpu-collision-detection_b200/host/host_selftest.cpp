// host_selftest -- exercises the host-side pieces that need no GPU (used by tests/test_host_cpu.py).
//   host_selftest npy-write <file> <rows> <cols>     writes a[i][j] = i + j/16
//   host_selftest npy-read <file>                     prints shape and a checksum
//   host_selftest cli <args...>                       parses generate_dataset-style flags, prints what it saw
//   host_selftest tables <n_variances> <n_poses>      parallel table generation == the sequential std:: loop
//   host_selftest seed <text>                         64-bit --seed parsing
#include <cstdio>
#include <cstring>
#include <iostream>

#include "cli.hpp"
#include "npy.hpp"
#include "tables.hpp"

int main(int argc, char** argv) try {
    if (argc < 2) return 64;
    const std::string mode = argv[1];
    if (mode == "npy-write" && argc == 5) {
        size_t r = std::stoul(argv[3]), c = std::stoul(argv[4]);
        std::vector<float> a(r * c);
        for (size_t i = 0; i < r; i++) for (size_t j = 0; j < c; j++) a[i * c + j] = (float)i + (float)j / 16.0f;
        if (c == 1) npyio::save_f32(argv[2], {r}, a); else npyio::save_f32(argv[2], {r, c}, a);
        return 0;
    }
    if (mode == "npy-read" && argc == 3) {
        npyio::Array a = npyio::load_f32(argv[2]);
        double sum = 0;
        for (float v : a.data) sum += v;
        std::printf("ndim %zu shape", a.shape.size());
        for (size_t s : a.shape) std::printf(" %zu", s);
        std::printf(" sum %.6f first %.6f last %.6f\n", sum, a.data.empty() ? 0.0 : a.data.front(), a.data.empty() ? 0.0 : a.data.back());
        return 0;
    }
    if (mode == "tables" && argc == 4) {
        // parallel skip-ahead generation == upstream's sequential std:: loop, bit for bit (variances then poses)
        const size_t nv = std::stoul(argv[2]), np_ = std::stoul(argv[3]);
        const float vmin[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, vmax[5] = {.3f, .3f, .3f, 0.f, 0.f};
        const float pmin[3] = {0.1f, 0.1f, 0.f}, pmax[3] = {5.f, 5.f, 6.2831855f};
        std::vector<float> v_ref(5 * nv), p_ref(3 * np_), v(5 * nv), q(3 * np_);
        std::default_random_engine gen;
        tables::fill_uniform_rows_std(v_ref.data(), nv, 5, vmin, vmax, gen);
        tables::fill_uniform_rows_std(p_ref.data(), np_, 3, pmin, pmax, gen);
        size_t bad = 0;
        for (unsigned threads : {1u, 3u, 8u}) {
            tables::fill_uniform_rows(v.data(), nv, 5, vmin, vmax, 0, threads);
            tables::fill_uniform_rows(q.data(), np_, 3, pmin, pmax, 5ull * nv, threads);
            for (size_t i = 0; i < v.size(); i++) bad += std::memcmp(&v[i], &v_ref[i], 4) != 0;
            for (size_t i = 0; i < q.size(); i++) bad += std::memcmp(&q[i], &p_ref[i], 4) != 0;
        }
        std::printf("mismatches %zu\n", bad);
        return bad ? 3 : 0;
    }
    if (mode == "seed" && argc == 3) {
        cli::Parser p("o");
        p.add("seed", cli::Kind::String, "seed");
        char* av[] = {argv[0], (char*)"--seed", argv[2]};
        p.parse(3, av);
        std::printf("%llu\n", p.unsigned64("seed"));
        return 0;
    }
    if (mode == "cli") {
        using cli::Kind;
        cli::Parser p("Allowed options");
        p.add("help", Kind::Switch, "produce help message").add("data_dir", Kind::String, "dir")
         .add("num_batches", Kind::Int, "n", 'n').add("batch_size", Kind::Int, "b", 'b').add("start_batch_count", Kind::Int, "s", 's')
         .add("shape_variance", Kind::Switch, "sv").add("max_variance", Kind::FloatList, "mv").add("min_pose", Kind::FloatList, "mp")
         .add("robot_width", Kind::Float, "w", 'w').add("robot_height", Kind::Float, "h", 'h').add("shuffle", Kind::Bool, "sh");
        p.parse(argc - 1, argv + 1);
        if (p.count("help")) { p.print_help(std::cout); return 1; }
        for (const char* k : {"data_dir"}) if (p.count(k)) std::printf("%s=%s\n", k, p.str(k).c_str());
        for (const char* k : {"num_batches", "batch_size", "start_batch_count"}) if (p.count(k)) std::printf("%s=%d\n", k, p.integer(k));
        for (const char* k : {"robot_width", "robot_height"}) if (p.count(k)) std::printf("%s=%g\n", k, p.real(k));
        for (const char* k : {"max_variance", "min_pose"}) if (p.count(k)) { std::printf("%s=", k); for (float v : p.reals(k)) std::printf("%g,", v); std::printf("\n"); }
        if (p.count("shape_variance")) std::printf("shape_variance=1\n");
        if (p.count("shuffle")) std::printf("shuffle=%d\n", (int)p.boolean("shuffle"));
        return 0;
    }
    return 64;
} catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 2;
}
