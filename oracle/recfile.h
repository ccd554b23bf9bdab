// TEST INFRASTRUCTURE ONLY (oracle/). Tiny tagged-array container shared by the oracle
// tools and the python tests (tests/recfile.py reads and writes the same layout).
//
//   file   := "SDRR0001" record*
//   record := u32 name_len | name bytes | u32 dtype | u64 count | payload (count * size(dtype))
//   dtype  := 0 u8, 1 i16, 2 i32, 3 f32, 4 f64, 5 u64
#ifndef SDR_ORACLE_RECFILE_H
#define SDR_ORACLE_RECFILE_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rec {

enum DType : uint32_t { U8 = 0, I16 = 1, I32 = 2, F32 = 3, F64 = 4, U64 = 5 };

inline size_t dtype_size(uint32_t d) {
    static const size_t sz[] = {1, 2, 4, 4, 8, 8};
    if (d > 5) throw std::runtime_error("recfile: bad dtype");
    return sz[d];
}

struct Array {
    uint32_t dtype = U8;
    std::vector<uint8_t> bytes;
    size_t count() const { return bytes.size() / dtype_size(dtype); }
    template <typename T> const T* as() const { return reinterpret_cast<const T*>(bytes.data()); }
    template <typename T> std::vector<T> vec() const {
        const T* p = as<T>();
        return std::vector<T>(p, p + bytes.size() / sizeof(T));
    }
};

class Writer {
public:
    explicit Writer(const std::string& path) : f_(std::fopen(path.c_str(), "wb")) {
        if (!f_) throw std::runtime_error("recfile: cannot open " + path);
        std::fwrite("SDRR0001", 1, 8, f_);
    }
    ~Writer() { if (f_) std::fclose(f_); }
    void put(const std::string& name, uint32_t dtype, const void* data, uint64_t count) {
        uint32_t nl = static_cast<uint32_t>(name.size());
        std::fwrite(&nl, 4, 1, f_);
        std::fwrite(name.data(), 1, nl, f_);
        std::fwrite(&dtype, 4, 1, f_);
        std::fwrite(&count, 8, 1, f_);
        if (count) std::fwrite(data, dtype_size(dtype), count, f_);
    }
    void put(const std::string& n, const std::vector<float>& v) { put(n, F32, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<double>& v) { put(n, F64, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<int32_t>& v) { put(n, I32, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<int16_t>& v) { put(n, I16, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<uint8_t>& v) { put(n, U8, v.data(), v.size()); }
    void put(const std::string& n, const std::vector<uint64_t>& v) { put(n, U64, v.data(), v.size()); }
    void put(const std::string& n, const std::string& s) { put(n, U8, s.data(), s.size()); }
private:
    std::FILE* f_;
};

inline std::map<std::string, Array> read_all(const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("recfile: cannot open " + path);
    char magic[8];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "SDRR0001", 8) != 0) {
        std::fclose(f);
        throw std::runtime_error("recfile: bad magic in " + path);
    }
    std::map<std::string, Array> out;
    for (;;) {
        uint32_t nl;
        if (std::fread(&nl, 4, 1, f) != 1) break;
        std::string name(nl, '\0');
        if (nl && std::fread(&name[0], 1, nl, f) != nl) break;
        Array a;
        uint64_t count;
        if (std::fread(&a.dtype, 4, 1, f) != 1 || std::fread(&count, 8, 1, f) != 1) break;
        a.bytes.resize(count * dtype_size(a.dtype));
        if (count && std::fread(a.bytes.data(), 1, a.bytes.size(), f) != a.bytes.size()) break;
        out[name] = std::move(a);
    }
    std::fclose(f);
    return out;
}

}  // namespace rec

#endif  // SDR_ORACLE_RECFILE_H
