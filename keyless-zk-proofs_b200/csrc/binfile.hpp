// iden3 "binfile" container, groth16 zkey header and wtns header — host-side thin C++.
//
// Behavioural mirror of the reference loaders (rust-rapidsnark/rapidsnark/src):
//   container  binfile_utils.cpp:13-58   magic(4) | u32 version | u32 nSections | {u32 id, u64 size, payload}*
//   mmap       fileloader.hpp:20-47      open/fstat/mmap failures are "file load" errors
//   zkey hdr   zkey_utils.hpp:48-87      section 1 protocol==1 (groth16); section 2 primes, sizes, vk points
//   wtns hdr   wtns_utils.hpp:28-43      section 1: n8, prime, nWitness; section 2: values
// Error classes keep the reference's split because FullProver's constructor maps them to different
// FullProverState values (fullprover.cpp:80-101): FormatError <-> std::invalid_argument
// (UNSUPPORTED_ZKEY_CURVE), LoadError <-> std::system_error (ZKEY_FILE_LOAD_ERROR). Unlike the
// reference (which asserts, binfile_utils.cpp:21,143), every read is bounds-checked and reported.
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace kzp
{

struct LoadError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};
struct FormatError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

class MappedFile
{
    uint8_t* addr_ = nullptr;
    size_t   size_ = 0;
    int      fd_   = -1;

public:
    explicit MappedFile(const std::string& path)
    {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0)
            throw LoadError("open failed: " + path);
        struct stat sb;
        if (::fstat(fd, &sb) != 0)
        {
            ::close(fd);
            throw LoadError("fstat failed: " + path);
        }
        size_ = (size_t)sb.st_size;
        if (size_ == 0)
        {
            ::close(fd);
            throw LoadError("empty file: " + path);
        }
        void* m = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED)
        {
            ::close(fd);
            throw LoadError("mmap failed: " + path);
        }
        addr_ = static_cast<uint8_t*>(m);
        fd_   = fd;
    }
    ~MappedFile()
    {
        if (addr_)
            ::munmap(addr_, size_);
        if (fd_ >= 0)
            ::close(fd_);
    }
    int fd() const { return fd_; } // stays open for bulk pread() (cheaper than faulting the mapping in)
    MappedFile(const MappedFile&)            = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    const uint8_t* data() const { return addr_; }
    size_t         size() const { return size_; }
};

struct Section
{
    const uint8_t* data;
    uint64_t       size;
};

// Parsed view over a mapped (or caller-owned) buffer. Does not own the bytes.
class BinView
{
    std::map<uint32_t, std::vector<Section>> sections_;

public:
    BinView(const uint8_t* buf, size_t size, const char* magic, uint32_t max_version)
    {
        if (size < 12)
            throw FormatError("file too short");
        if (memcmp(buf, magic, 4) != 0)
            throw FormatError(std::string("invalid file type, expected ") + magic);
        uint32_t version, n_sections;
        memcpy(&version, buf + 4, 4);
        memcpy(&n_sections, buf + 8, 4);
        if (version > max_version)
            throw FormatError("invalid version");
        size_t pos = 12;
        for (uint32_t i = 0; i < n_sections; i++)
        {
            if (pos + 12 > size)
                throw FormatError("truncated section table");
            uint32_t id;
            uint64_t ssize;
            memcpy(&id, buf + pos, 4);
            memcpy(&ssize, buf + pos + 4, 8);
            pos += 12;
            if (ssize > size - pos)
                throw FormatError("section exceeds file size");
            sections_[id].push_back(Section{buf + pos, ssize});
            pos += ssize;
        }
    }

    bool has(uint32_t id) const { return sections_.count(id) != 0; }

    const Section& section(uint32_t id, uint32_t which = 0) const
    {
        auto it = sections_.find(id);
        if (it == sections_.end() || which >= it->second.size())
            throw FormatError("section does not exist: " + std::to_string(id));
        return it->second[which];
    }
};

// BN254 primes as little-endian bytes (fullprover.cpp:140-143 for r; fq_raw_generic.cpp:6 for q)
static const uint8_t kBn254R[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9,
                                    0x79, 0x48, 0xe8, 0x33, 0x28, 0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45,
                                    0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
static const uint8_t kBn254Q[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71,
                                    0x68, 0x91, 0x6a, 0x81, 0x97, 0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45,
                                    0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

struct ZkeyHeader
{
    uint32_t       n8q = 0, n8r = 0;
    bool           r_is_bn254 = false;
    bool           q_is_bn254 = false;
    uint32_t       n_vars = 0, n_public = 0, domain_size = 0;
    uint64_t       n_coefs = 0;
    const uint8_t* alpha1 = nullptr; // G1, 64 B
    const uint8_t* beta1  = nullptr; // G1
    const uint8_t* beta2  = nullptr; // G2, 128 B
    const uint8_t* gamma2 = nullptr; // G2
    const uint8_t* delta1 = nullptr; // G1
    const uint8_t* delta2 = nullptr; // G2
    const uint8_t* coefs  = nullptr; // section 4 payload AFTER the 4-byte count (groth16.cpp:33)
    const uint8_t* points_a = nullptr;
    const uint8_t* points_b1 = nullptr;
    const uint8_t* points_b2 = nullptr;
    const uint8_t* points_c = nullptr;
    const uint8_t* points_h = nullptr;
};

static inline bool prime_equals(const uint8_t* p, uint32_t n8, const uint8_t (&ref)[32])
{
    if (n8 < 32)
        return false;
    if (memcmp(p, ref, 32) != 0)
        return false;
    for (uint32_t i = 32; i < n8; i++)
        if (p[i])
            return false;
    return true;
}

static inline ZkeyHeader parse_zkey(const BinView& bin)
{
    ZkeyHeader h;
    {
        const Section& s = bin.section(1);
        if (s.size < 4)
            throw FormatError("zkey section 1 too short");
        uint32_t protocol;
        memcpy(&protocol, s.data, 4);
        if (protocol != 1)
            throw FormatError("zkey file is not groth16");
    }
    const Section& s   = bin.section(2);
    size_t         pos = 0;
    auto           need = [&](size_t n) {
        if (pos + n > s.size)
            throw FormatError("zkey header truncated");
    };
    need(4);
    memcpy(&h.n8q, s.data + pos, 4);
    pos += 4;
    need(h.n8q);
    h.q_is_bn254 = prime_equals(s.data + pos, h.n8q, kBn254Q);
    pos += h.n8q;
    need(4);
    memcpy(&h.n8r, s.data + pos, 4);
    pos += 4;
    need(h.n8r);
    h.r_is_bn254 = prime_equals(s.data + pos, h.n8r, kBn254R);
    pos += h.n8r;
    need(12);
    memcpy(&h.n_vars, s.data + pos, 4);
    memcpy(&h.n_public, s.data + pos + 4, 4);
    memcpy(&h.domain_size, s.data + pos + 8, 4);
    pos += 12;
    // the curve check happens before any section pointer is trusted (fullprover.cpp:154-158)
    if (!h.r_is_bn254 || !h.q_is_bn254 || h.n8q != 32 || h.n8r != 32)
        throw FormatError("zkey curve not supported");
    need(64 * 3 + 128 * 3);
    h.alpha1 = s.data + pos; pos += 64;
    h.beta1  = s.data + pos; pos += 64;
    h.beta2  = s.data + pos; pos += 128;
    h.gamma2 = s.data + pos; pos += 128;
    h.delta1 = s.data + pos; pos += 64;
    h.delta2 = s.data + pos; pos += 128;

    if (h.domain_size == 0 || (h.domain_size & (h.domain_size - 1)) != 0)
        throw FormatError("zkey domain size is not a power of two");
    if ((uint64_t)h.n_public + 1 > (uint64_t)h.n_vars) // 64-bit: nPublic = 0xffffffff must not wrap to 0
        throw FormatError("zkey nPublic exceeds nVars");

    const Section& c4 = bin.section(4);
    h.n_coefs         = c4.size / (12 + h.n8r); // zkey_utils.hpp:84
    if (c4.size < 4 || 4 + h.n_coefs * 44 > c4.size)
        throw FormatError("zkey coefficient section truncated");
    h.coefs = c4.data + 4;

    auto pts = [&](uint32_t id, uint64_t count, uint64_t bytes_each) {
        const Section& ps = bin.section(id);
        if (ps.size < count * bytes_each)
            throw FormatError("zkey point section too short: " + std::to_string(id));
        return ps.data;
    };
    h.points_a  = pts(5, h.n_vars, 64);
    h.points_b1 = pts(6, h.n_vars, 64);
    h.points_b2 = pts(7, h.n_vars, 128);
    h.points_c  = pts(8, (uint64_t)h.n_vars - h.n_public - 1, 64);
    h.points_h  = pts(9, h.domain_size, 64);
    return h;
}

struct WtnsHeader
{
    uint32_t       n8 = 0;
    bool           prime_is_bn254_r = false;
    uint32_t       n_witness = 0;
    const uint8_t* values = nullptr; // n_witness x 32 B canonical little-endian
    uint64_t       values_bytes = 0;
};

static inline WtnsHeader parse_wtns(const BinView& bin)
{
    WtnsHeader     h;
    const Section& s = bin.section(1);
    if (s.size < 4)
        throw FormatError("wtns header too short");
    memcpy(&h.n8, s.data, 4);
    if (s.size < 4 + (uint64_t)h.n8 + 4)
        throw FormatError("wtns header truncated");
    h.prime_is_bn254_r = (h.n8 == 32) && prime_equals(s.data + 4, h.n8, kBn254R);
    memcpy(&h.n_witness, s.data + 4 + h.n8, 4);
    const Section& d = bin.section(2);
    h.values         = d.data;
    h.values_bytes   = d.size;
    return h;
}

} // namespace kzp
