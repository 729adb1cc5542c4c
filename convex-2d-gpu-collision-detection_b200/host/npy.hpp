// npy.hpp -- minimal float32 .npy reader/writer for the dataset files (header-only, no dependencies).
//
// The reference does its file I/O through the third-party libnpy (npy::LoadArrayFromNumpy /
// npy::SaveArrayAsNumpy, utils.cu:217-224, generate_dataset.cu:303,332,351-352,500), which is not
// vendored.  Every file the three programs exchange is little-endian float32, C order, format
// version 1.0; that is all this header supports, and it says so loudly otherwise.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace npyio {

struct Array {
    std::vector<size_t> shape;
    std::vector<float> data;
    size_t rows() const { return shape.empty() ? 0 : shape[0]; }
    size_t cols() const { return shape.size() < 2 ? 1 : shape[1]; }
};

inline size_t count(const std::vector<size_t>& shape) {
    size_t n = 1;
    for (size_t s : shape) n *= s;
    return n;
}

inline void save_f32(const std::string& path, const std::vector<size_t>& shape, const float* data) {
    std::string dict = "{'descr': '<f4', 'fortran_order': False, 'shape': (";
    for (size_t i = 0; i < shape.size(); i++) {
        dict += std::to_string(shape[i]);
        if (shape.size() == 1 || i + 1 < shape.size()) dict += ",";
        if (i + 1 < shape.size()) dict += " ";
    }
    dict += "), }";
    // magic(6) + version(2) + header_len(2) + dict + padding + '\n' must be a multiple of 64
    size_t unpadded = 10 + dict.size() + 1;
    size_t pad = (64 - unpadded % 64) % 64;
    dict.append(pad, ' ');
    dict.push_back('\n');
    if (dict.size() > 65535) throw std::runtime_error("npy: header too long for format 1.0");
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("npy: cannot open " + path + " for writing");
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    f.write(reinterpret_cast<const char*>(magic), 8);
    const unsigned char len[2] = {(unsigned char)(dict.size() & 0xff), (unsigned char)(dict.size() >> 8)};
    f.write(reinterpret_cast<const char*>(len), 2);
    f.write(dict.data(), (std::streamsize)dict.size());
    f.write(reinterpret_cast<const char*>(data), (std::streamsize)(count(shape) * sizeof(float)));
    if (!f) throw std::runtime_error("npy: write to " + path + " failed");
}

inline void save_f32(const std::string& path, const std::vector<size_t>& shape, const std::vector<float>& data) {
    if (data.size() != count(shape)) throw std::runtime_error("npy: shape does not match data size for " + path);
    save_f32(path, shape, data.data());
}

inline Array load_f32(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("npy: cannot open " + path);
    unsigned char head[10];
    f.read(reinterpret_cast<char*>(head), 8);
    if (!f || std::memcmp(head, "\x93NUMPY", 6) != 0) throw std::runtime_error("npy: " + path + " is not a .npy file");
    size_t hlen = 0;
    if (head[6] == 1) {
        f.read(reinterpret_cast<char*>(head + 8), 2);
        hlen = head[8] | (size_t)head[9] << 8;
    } else if (head[6] == 2 || head[6] == 3) {
        unsigned char l4[4];
        f.read(reinterpret_cast<char*>(l4), 4);
        hlen = l4[0] | (size_t)l4[1] << 8 | (size_t)l4[2] << 16 | (size_t)l4[3] << 24;
    } else {
        throw std::runtime_error("npy: unsupported format version in " + path);
    }
    std::string dict(hlen, '\0');
    f.read(&dict[0], (std::streamsize)hlen);
    if (!f) throw std::runtime_error("npy: truncated header in " + path);
    auto value_of = [&](const std::string& key) -> std::string {
        size_t k = dict.find("'" + key + "'");
        if (k == std::string::npos) throw std::runtime_error("npy: key " + key + " missing in " + path);
        size_t c = dict.find(':', k);
        size_t b = dict.find_first_not_of(" ", c + 1);
        size_t e;
        if (dict[b] == '(') e = dict.find(')', b) + 1;
        else if (dict[b] == '\'') e = dict.find('\'', b + 1) + 1;
        else e = dict.find_first_of(",}", b);
        return dict.substr(b, e - b);
    };
    const std::string descr = value_of("descr");
    if (descr != "'<f4'" && descr != "'=f4'" && descr != "'|f4'")
        throw std::runtime_error("npy: " + path + " has dtype " + descr + ", expected little-endian float32 ('<f4')");
    if (value_of("fortran_order").rfind("False", 0) != 0)
        throw std::runtime_error("npy: " + path + " is Fortran-ordered; C order expected");
    Array a;
    const std::string shp = value_of("shape");
    size_t i = 1;
    while (i < shp.size()) {
        while (i < shp.size() && (shp[i] < '0' || shp[i] > '9')) i++;
        if (i >= shp.size()) break;
        size_t v = 0;
        while (i < shp.size() && shp[i] >= '0' && shp[i] <= '9') v = v * 10 + (size_t)(shp[i++] - '0');
        a.shape.push_back(v);
    }
    a.data.resize(count(a.shape));
    f.read(reinterpret_cast<char*>(a.data.data()), (std::streamsize)(a.data.size() * sizeof(float)));
    if ((size_t)f.gcount() != a.data.size() * sizeof(float)) throw std::runtime_error("npy: truncated data in " + path);
    return a;
}

}  // namespace npyio
