// pgsgd_arrays.hpp — tiny named-array container ("PGSGDARR" v1) used for flattened graphs, layouts
// and golden vectors.  Written by the C++ host tools, read/written by odgi_b200/arrays.py (numpy).
//
//   magic "PGSGDARR" | u32 version=1 | u32 n_arrays
//   per array: u32 name_len | name | u32 dtype | u64 count | raw little-endian data | pad to 8 B
//
// dtype codes: 0=u8 1=u32 2=u64 3=f32 4=f64 5=i64
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pgsgd {

enum ArrDtype : uint32_t { ARR_U8 = 0, ARR_U32 = 1, ARR_U64 = 2, ARR_F32 = 3, ARR_F64 = 4, ARR_I64 = 5 };

inline size_t arr_dtype_size(uint32_t dt) {
    switch (dt) {
        case ARR_U8: return 1;
        case ARR_U32: case ARR_F32: return 4;
        case ARR_U64: case ARR_F64: case ARR_I64: return 8;
    }
    throw std::runtime_error("pgsgd_arrays: bad dtype");
}

template <typename T> struct arr_dtype_of;
template <> struct arr_dtype_of<uint8_t>  { static constexpr uint32_t v = ARR_U8; };
template <> struct arr_dtype_of<uint32_t> { static constexpr uint32_t v = ARR_U32; };
template <> struct arr_dtype_of<uint64_t> { static constexpr uint32_t v = ARR_U64; };
template <> struct arr_dtype_of<float>    { static constexpr uint32_t v = ARR_F32; };
template <> struct arr_dtype_of<double>   { static constexpr uint32_t v = ARR_F64; };
template <> struct arr_dtype_of<int64_t>  { static constexpr uint32_t v = ARR_I64; };

class ArrayWriter {
public:
    explicit ArrayWriter(const std::string& path) : f_(std::fopen(path.c_str(), "wb")) {
        if (!f_) throw std::runtime_error("pgsgd_arrays: cannot open " + path + " for writing");
        std::fwrite("PGSGDARR", 1, 8, f_);
        uint32_t hdr[2] = {1u, 0u};
        std::fwrite(hdr, 4, 2, f_);
    }
    ~ArrayWriter() { close(); }
    template <typename T>
    void add(const std::string& name, const T* data, uint64_t count) {
        uint32_t nl = (uint32_t) name.size();
        uint32_t dt = arr_dtype_of<T>::v;
        std::fwrite(&nl, 4, 1, f_);
        std::fwrite(name.data(), 1, nl, f_);
        std::fwrite(&dt, 4, 1, f_);
        std::fwrite(&count, 8, 1, f_);
        if (count) std::fwrite(data, sizeof(T), count, f_);
        long pos = std::ftell(f_);
        static const char zeros[8] = {0};
        if (pos % 8) std::fwrite(zeros, 1, 8 - (pos % 8), f_);
        ++n_;
    }
    template <typename T>
    void add(const std::string& name, const std::vector<T>& v) { add(name, v.data(), (uint64_t) v.size()); }
    template <typename T>
    void add_scalar(const std::string& name, T v) { add(name, &v, 1); }
    void close() {
        if (!f_) return;
        std::fseek(f_, 12, SEEK_SET);
        std::fwrite(&n_, 4, 1, f_);
        std::fclose(f_);
        f_ = nullptr;
    }
private:
    FILE* f_;
    uint32_t n_ = 0;
};

struct ArrayEntry {
    uint32_t dtype = 0;
    uint64_t count = 0;
    std::vector<uint8_t> bytes;
    template <typename T> const T* as() const {
        if (arr_dtype_of<T>::v != dtype) throw std::runtime_error("pgsgd_arrays: dtype mismatch");
        return reinterpret_cast<const T*>(bytes.data());
    }
    template <typename T> std::vector<T> vec() const { const T* p = as<T>(); return std::vector<T>(p, p + count); }
};

inline std::map<std::string, ArrayEntry> read_arrays(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("pgsgd_arrays: cannot open " + path);
    char magic[8];
    uint32_t hdr[2];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "PGSGDARR", 8) != 0 || std::fread(hdr, 4, 2, f) != 2 || hdr[0] != 1) {
        std::fclose(f);
        throw std::runtime_error("pgsgd_arrays: " + path + " is not a PGSGDARR v1 file");
    }
    std::map<std::string, ArrayEntry> out;
    for (uint32_t i = 0; i < hdr[1]; ++i) {
        uint32_t nl = 0;
        if (std::fread(&nl, 4, 1, f) != 1) break;
        std::string name(nl, '\0');
        ArrayEntry e;
        bool ok = std::fread(&name[0], 1, nl, f) == nl && std::fread(&e.dtype, 4, 1, f) == 1 && std::fread(&e.count, 8, 1, f) == 1;
        if (!ok) { std::fclose(f); throw std::runtime_error("pgsgd_arrays: truncated " + path); }
        e.bytes.resize(e.count * arr_dtype_size(e.dtype));
        if (e.count && std::fread(e.bytes.data(), 1, e.bytes.size(), f) != e.bytes.size()) {
            std::fclose(f);
            throw std::runtime_error("pgsgd_arrays: truncated " + path);
        }
        long pos = std::ftell(f);
        if (pos % 8) std::fseek(f, 8 - (pos % 8), SEEK_CUR);
        out.emplace(std::move(name), std::move(e));
    }
    std::fclose(f);
    return out;
}

}  // namespace pgsgd
