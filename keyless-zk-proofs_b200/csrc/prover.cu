// Host-side driver of one GPU-resident Groth16 prover (thin C++: parse, upload once, launch, assemble).
// Stage order follows Prover::prove in the reference (rust-rapidsnark/rapidsnark/src/groth16.cpp:43-360);
// what the reference runs as std::async futures on CPU threads runs here on two CUDA streams.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <unistd.h>
#include <cerrno>
#include <sys/random.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "binfile.hpp"
#include "device.hpp"
#include "hostff.hpp"
#include "prover.hpp"

namespace kzp
{

typedef Fp2T<HFq>       HFq2;
typedef XyzzT<HFq>      HG1;
typedef XyzzT<HFq2>     HG2;
typedef AffineT<HFq>    HG1Affine;
typedef AffineT<HFq2>   HG2Affine;

static_assert(sizeof(HG1) == 128 && sizeof(HG2) == 256, "host point layout");
static_assert(sizeof(G1Xyzz) == 128 && sizeof(G2Xyzz) == 256, "device point layout");
static_assert(sizeof(G1Affine) == 64 && sizeof(G2Affine) == 128, "affine layout = zkey layout");

namespace
{

double now_ms()
{
    return std::chrono::duration<double, std::milli>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

template <class XY>
void scalar_mul(XY& out, const XY& base, const uint8_t* k32)
{
    XY acc;
    XY::set_inf(acc);
    int top = 255;
    while (top >= 0 && !((k32[top >> 3] >> (top & 7)) & 1))
        top--;
    for (int i = top; i >= 0; i--)
    {
        XY t = acc;
        XY::dbl(acc, t);
        if ((k32[i >> 3] >> (i & 7)) & 1)
            XY::add(acc, base);
    }
    out = acc;
}

// Fr bytes >= r ?
bool fr_bytes_geq_modulus(const uint8_t* b)
{
    HFr t;
    memcpy(&t, b, 32);
    return HFr::geq_p(t);
}

// 32 bytes from the kernel CSPRNG (the reference uses libsodium's randombytes_buf, random_generator.hpp:4-8, which
// reads the same source). The zero-knowledge of every proof rests on r and s: no fallback to a userspace PRNG; if
// the kernel cannot deliver, the proof fails with a CudaError-class (service-side) error.
void csprng_bytes(uint8_t* out, size_t n)
{
    size_t got = 0;
    while (got < n)
    {
        ssize_t k = ::getrandom(out + got, n - got, 0);
        if (k < 0)
        {
            if (errno == EINTR)
                continue;
            throw CudaError(std::string("getrandom failed: ") + strerror(errno));
        }
        got += (size_t)k;
    }
}

void sample_blinding(uint8_t* out32)
{
    // groth16.cpp:296-316: 32 random bytes, clear the top two bits, reject values >= r
    for (;;)
    {
        csprng_bytes(out32, 32);
        out32[31] &= 0x3f;
        if (!fr_bytes_geq_modulus(out32))
            return;
    }
}

template <class F>
void append_decimal(std::string& s, const F& mont)
{
    s += '"';
    s += F::to_decimal(mont);
    s += '"';
}

} // namespace

// r/s-dependent terms of the proof that do not depend on the MSM results (groth16.cpp:296-316, 328-347): computed
// on the host while the GPU works.
struct BlindTerms
{
    uint8_t r[32], s[32];
    HG1     r_delta1, s_delta1, rs_delta1;
    HG2     s_delta2;
};

static void compute_blind_terms(const HostVk& vk, const uint8_t* r32, const uint8_t* s32, BlindTerms& bt)
{
    HG1Affine delta1;
    HG2Affine delta2;
    memcpy(&delta1, vk.delta1, 64);
    memcpy(&delta2, vk.delta2, 128);
    if (r32 && s32)
    {
        memcpy(bt.r, r32, 32);
        memcpy(bt.s, s32, 32);
    }
    else
    {
        sample_blinding(bt.r);
        sample_blinding(bt.s);
    }
    // rs = r*s mod r_modulus, canonical (groth16.cpp:346-347)
    uint8_t rs[32];
    {
        HFr fr, fs, t;
        memcpy(&fr, bt.r, 32);
        memcpy(&fs, bt.s, 32);
        while (HFr::geq_p(fr))
            HFr::sub_p(fr);
        while (HFr::geq_p(fs))
            HFr::sub_p(fs);
        HFr::to_mont(fr, fr);
        HFr::mul(t, fr, fs); // (r R)(s) R^-1 = r s
        memcpy(rs, &t, 32);
    }
    HG1 d1;
    HG2 d2;
    HG1::from_affine(d1, delta1);
    HG2::from_affine(d2, delta2);
    scalar_mul(bt.r_delta1, d1, bt.r);
    scalar_mul(bt.s_delta1, d1, bt.s);
    scalar_mul(bt.rs_delta1, d1, rs);
    scalar_mul(bt.s_delta2, d2, bt.s);
}

// Sums the shards' partial MSM results, blinds and prints the proof. Host-only arithmetic (a few thousand field
// multiplications); mirrors groth16.cpp:318-357 + Proof::toJson (:379-410).
static void artefacts_from_sums(const ShardPartials& sums, MsmArtefacts& art)
{
    HG1 A, B1, C, H;
    HG2 B2;
    memcpy(&A, sums.bytes + 0, 128);
    memcpy(&B1, sums.bytes + 128, 128);
    memcpy(&C, sums.bytes + 256, 128);
    memcpy(&H, sums.bytes + 384, 128);
    memcpy(&B2, sums.bytes + 512, 256);
    auto put1 = [&](uint8_t* out, const HG1& p) {
        HG1Affine a;
        HG1::to_affine(a, p);
        HFq t;
        HFq::from_mont(t, a.x);
        memcpy(out, &t, 32);
        HFq::from_mont(t, a.y);
        memcpy(out + 32, &t, 32);
    };
    put1(art.bytes + 0, A);
    put1(art.bytes + 64, B1);
    HG2Affine b2;
    HG2::to_affine(b2, B2);
    HFq t;
    HFq::from_mont(t, b2.x.a);
    memcpy(art.bytes + 128, &t, 32);
    HFq::from_mont(t, b2.x.b);
    memcpy(art.bytes + 160, &t, 32);
    HFq::from_mont(t, b2.y.a);
    memcpy(art.bytes + 192, &t, 32);
    HFq::from_mont(t, b2.y.b);
    memcpy(art.bytes + 224, &t, 32);
    put1(art.bytes + 256, C);
    put1(art.bytes + 320, H);
}

// Proof assembly in three steps, each as soon as its inputs are on the host, so that everything that does not need
// the H MSM result runs while the GPU is still busy with it: g1 = pi_a and pi_c without its H term (needs A, B1, C:
// two 254-bit scalar multiplications), g2 = pi_b (needs B2), final = + H.
// pi_c = C + H + s*pi_a + r*pib1 - rs*delta1 (groth16.cpp:340-352) is a sum in an abelian group, so adding H last
// gives the same affine point.
struct EarlyProof
{
    HG1         A, B1, C;
    HG2         B2;
    HG1         pi_c_partial; // C + s*pi_a + r*pib1 - rs*delta1
    std::string pi_a, pi_b;   // their JSON fragments
};

static void assemble_early_g1(const HostVk& vk, const BlindTerms& bt, const HG1& A, const HG1& B1, const HG1& C, EarlyProof& ep)
{
    HG1Affine alpha1, beta1;
    memcpy(&alpha1, vk.alpha1, 64);
    memcpy(&beta1, vk.beta1, 64);
    ep.A  = A;
    ep.B1 = B1;
    ep.C  = C;
    HG1 al, be1;
    HG1::from_affine(al, alpha1);
    HG1::from_affine(be1, beta1);
    // pi_a = A + alpha1 + r*delta1            (groth16.cpp:328-330)
    HG1 pi_a = A;
    HG1::add(pi_a, al);
    HG1::add(pi_a, bt.r_delta1);
    // pib1 = B1 + beta1 + s*delta1            (:336-338)
    HG1 pib1 = B1;
    HG1::add(pib1, be1);
    HG1::add(pib1, bt.s_delta1);
    // pi_c - H = C + s*pi_a + r*pib1 - rs*delta1   (:340-352)
    HG1 pc = C, p1;
    scalar_mul(p1, pi_a, bt.s);
    HG1::add(pc, p1);
    scalar_mul(p1, pib1, bt.r);
    HG1::add(pc, p1);
    HG1 np1;
    HG1::neg(np1, bt.rs_delta1);
    HG1::add(pc, np1);
    ep.pi_c_partial = pc;
    HG1Affine a_aff;
    HG1::to_affine(a_aff, pi_a);
    std::string& j = ep.pi_a;
    j.clear();
    append_decimal(j, a_aff.x);
    j += ',';
    append_decimal(j, a_aff.y);
}

static void assemble_early_g2(const HostVk& vk, const BlindTerms& bt, const HG2& B2, EarlyProof& ep)
{
    HG2Affine beta2;
    memcpy(&beta2, vk.beta2, 128);
    ep.B2 = B2;
    HG2 be2;
    HG2::from_affine(be2, beta2);
    // pi_b = B2 + beta2 + s*delta2            (:332-334)
    HG2 pi_b = B2;
    HG2::add(pi_b, be2);
    HG2::add(pi_b, bt.s_delta2);
    HG2Affine b_aff;
    HG2::to_affine(b_aff, pi_b);
    std::string& j = ep.pi_b;
    j.clear();
    j += '[';
    append_decimal(j, b_aff.x.a);
    j += ',';
    append_decimal(j, b_aff.x.b);
    j += "],[";
    append_decimal(j, b_aff.y.a);
    j += ',';
    append_decimal(j, b_aff.y.b);
    j += ']';
}

static void assemble_early(const HostVk& vk, const BlindTerms& bt, const HG1& A, const HG1& B1, const HG1& C,
                           const HG2& B2, EarlyProof& ep)
{
    assemble_early_g1(vk, bt, A, B1, C, ep);
    assemble_early_g2(vk, bt, B2, ep);
}

// sums_out (optional) receives the five summed MSM results (XYZZ) for the parity artefacts.
static std::string assemble_final(const EarlyProof& ep, const HG1& H, ShardPartials* sums_out)
{
    HG1 pi_c = ep.pi_c_partial;
    HG1::add(pi_c, H);
    if (sums_out)
    {
        memcpy(sums_out->bytes + 0, &ep.A, 128);
        memcpy(sums_out->bytes + 128, &ep.B1, 128);
        memcpy(sums_out->bytes + 256, &ep.C, 128);
        memcpy(sums_out->bytes + 384, &H, 128);
        memcpy(sums_out->bytes + 512, &ep.B2, 256);
    }
    HG1Affine c_aff;
    HG1::to_affine(c_aff, pi_c);
    // compact JSON, keys in sorted order, exactly what nlohmann::json::dump() prints for
    // Proof::toJson (groth16.cpp:379-410, fullprover.cpp:246)
    std::string j;
    j.reserve(900);
    j += "{\"pi_a\":[";
    j += ep.pi_a;
    j += ",\"1\"],\"pi_b\":[";
    j += ep.pi_b;
    j += ",[\"1\",\"0\"]],\"pi_c\":[";
    append_decimal(j, c_aff.x);
    j += ',';
    append_decimal(j, c_aff.y);
    j += ",\"1\"],\"protocol\":\"groth16\"}";
    return j;
}

static std::string assemble_with_terms(const HostVk& vk, const ShardPartials* ps, int count, const BlindTerms& bt,
                                       ShardPartials* sums_out)
{
    HG1 A, B1, C, H;
    HG2 B2;
    HG1::set_inf(A);
    HG1::set_inf(B1);
    HG1::set_inf(C);
    HG1::set_inf(H);
    HG2::set_inf(B2);
    for (int k = 0; k < count; k++)
    {
        HG1 t;
        HG2 t2;
        memcpy(&t, ps[k].bytes + 0, 128);
        HG1::add(A, t);
        memcpy(&t, ps[k].bytes + 128, 128);
        HG1::add(B1, t);
        memcpy(&t, ps[k].bytes + 256, 128);
        HG1::add(C, t);
        memcpy(&t, ps[k].bytes + 384, 128);
        HG1::add(H, t);
        memcpy(&t2, ps[k].bytes + 512, 256);
        HG2::add(B2, t2);
    }
    EarlyProof ep;
    assemble_early(vk, bt, A, B1, C, B2, ep);
    return assemble_final(ep, H, sums_out);
}

std::string assemble_proof(const HostVk& vk, const ShardPartials* ps, int count, const uint8_t* r32,
                           const uint8_t* s32, MsmArtefacts* art_out)
{
    BlindTerms bt;
    compute_blind_terms(vk, r32, s32, bt);
    ShardPartials sums;
    std::string   j = assemble_with_terms(vk, ps, count, bt, art_out ? &sums : nullptr);
    if (art_out)
        artefacts_from_sums(sums, *art_out);
    return j;
}

// ---- packed witness transfer ------------------------------------------------------------------------------
// A circom witness is mostly bits and bytes (keyless: ~96 % of the wires are below 256), so shipping 32 bytes per
// wire over PCIe is the slowest part of the end-to-end path (43 MB, ~1.1 ms DMA-bound). The staging workers classify
// every value while they copy it: one byte per wire (the value when it is below 256), one flag bit per wire, and the
// full 32 bytes only for the others. A slice of kPackWires wires becomes [small bytes][flag words][full values...]
// and only that prefix crosses the bus; k_witness_expand rebuilds the plain n_vars x 32-byte vector in HBM.
constexpr uint32_t kPackWires  = 32768;                                   // wires per slice (1 MiB of plain values)
constexpr size_t   kPackSmall  = kPackWires;                              // 1 byte per wire
constexpr size_t   kPackFlags  = kPackWires / 8;                          // 1 bit per wire
constexpr size_t   kPackHead   = kPackSmall + kPackFlags;
constexpr size_t   kPackStride = kPackHead + (size_t)kPackWires * 32;     // worst case: every value is full width
constexpr uint32_t kPackGroup  = 128;                                     // wires classified per inner step

// Classifies `count` values (count a multiple of kPackGroup; the caller zero-pads) that start at wire `first` of
// the slice whose packed image begins at `slice`; *n_full is the slice's running count of full-width values.
void pack_values(uint8_t* slice, uint32_t first, const uint8_t* vals, uint32_t count, uint32_t* n_full)
{
    uint8_t*  small = slice + first;
    uint8_t*  flags = slice + kPackSmall + first / 8;
    uint8_t*  full  = slice + kPackHead;
    uint32_t  nf    = *n_full;
    const __m128i zero     = _mm_setzero_si128();
    const __m128i not_byte0 = _mm_set_epi32(-1, -1, -1, (int)0xffffff00u);
    for (uint32_t g = 0; g < count; g += kPackGroup)
    {
        alignas(16) uint8_t  sm[kPackGroup];
        alignas(16) uint32_t fl[kPackGroup / 32];
        for (uint32_t w = 0; w < kPackGroup / 32; w++)
        {
            uint32_t bits = 0;
            for (uint32_t j = 0; j < 32; j++)
            {
                const uint8_t* v  = vals + (size_t)(g + 32 * w + j) * 32;
                __m128i        lo = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v));
                __m128i        hi = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v + 16));
                __m128i        x  = _mm_or_si128(hi, _mm_and_si128(lo, not_byte0));
                bool           is_small = _mm_movemask_epi8(_mm_cmpeq_epi8(x, zero)) == 0xffff;
                if (is_small)
                    sm[32 * w + j] = v[0];
                else
                {
                    sm[32 * w + j] = 0;
                    bits |= 1u << j;
                    __m128i* dst = reinterpret_cast<__m128i*>(full + (size_t)nf * 32);
                    _mm_stream_si128(dst, lo);
                    _mm_stream_si128(dst + 1, hi);
                    nf++;
                }
            }
            fl[w] = bits;
        }
        for (uint32_t k = 0; k < kPackGroup / 16; k++)
            _mm_stream_si128(reinterpret_cast<__m128i*>(small + g) + k, _mm_load_si128(reinterpret_cast<const __m128i*>(sm) + k));
        _mm_stream_si128(reinterpret_cast<__m128i*>(flags + g / 8), _mm_load_si128(reinterpret_cast<const __m128i*>(fl)));
    }
    _mm_sfence();
    *n_full = nf;
}

// One CTA per slice, 1024 threads: thread t owns flag word t (wires 32 t .. 32 t + 31 of the slice). A block-wide
// exclusive scan of the popcounts gives every word the index of its first full-width value; the warp then walks its
// 32 words with the lanes on consecutive wires, so the 32-byte stores are coalesced.
__global__ void __launch_bounds__(1024) k_witness_expand(const uint8_t* __restrict__ pack, uint4* __restrict__ w, uint32_t n_vars)
{
    __shared__ uint32_t warp_tot[32];
    const uint8_t*  slice = pack + (size_t)blockIdx.x * kPackStride;
    const uint8_t*  small = slice;
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(slice + kPackSmall);
    const uint4*    full  = reinterpret_cast<const uint4*>(slice + kPackHead);
    uint32_t        tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t        word = flags[tid];
    uint32_t        cnt  = __popc(word);
    uint32_t        inc  = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o)
            inc += t;
    }
    if (lane == 31)
        warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0)
    {
        uint32_t v = warp_tot[lane], x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o)
                x += t;
        }
        warp_tot[lane] = x - v;
    }
    __syncthreads();
    uint32_t base = warp_tot[wid] + inc - cnt; // index of this word's first full-width value
    for (uint32_t j = 0; j < 32; j++)
    {
        uint32_t wj   = __shfl_sync(0xffffffffu, word, j);
        uint32_t bj   = __shfl_sync(0xffffffffu, base, j);
        uint32_t loc  = (wid * 32 + j) * 32 + lane; // wire within the slice
        uint32_t wire = blockIdx.x * kPackWires + loc;
        if (wire >= n_vars)
            continue;
        uint4 lo, hi;
        if ((wj >> lane) & 1u)
        {
            uint32_t r = bj + __popc(wj & ((1u << lane) - 1u));
            lo         = full[2 * (size_t)r];
            hi         = full[2 * (size_t)r + 1];
        }
        else
        {
            lo = make_uint4(small[loc], 0, 0, 0);
            hi = make_uint4(0, 0, 0, 0);
        }
        w[2 * (size_t)wire]     = lo;
        w[2 * (size_t)wire + 1] = hi;
    }
}

// The two sources of witness values (see WitnessPacker::begin): the caller's memory, or an open .wtns file
inline auto witness_reader_mem(const uint8_t* values)
{
    return [values](uint8_t*, uint64_t first, uint32_t) -> const uint8_t* { return values + first * 32; };
}
inline auto witness_reader_fd(int fd, uint64_t file_offset)
{
    return [fd, file_offset](uint8_t* bounce, uint64_t first, uint32_t count) -> const uint8_t* {
        uint8_t* dst = bounce;
        size_t   len = (size_t)count * 32;
        off_t    off = (off_t)(file_offset + first * 32);
        while (len > 0)
        {
            ssize_t got = ::pread(fd, dst, len, off);
            if (got < 0 && errno == EINTR)
                continue;
            if (got <= 0) // error, or the file is shorter than its section table says
                return nullptr;
            dst += got;
            off += got;
            len -= (size_t)got;
        }
        return bounce;
    };
}

// Persistent staging workers (spawning threads per proof costs more than the copy they do).
class SlicePool
{
    std::vector<std::thread>           threads_;
    std::mutex                         mu_;
    std::condition_variable            cv_;
    std::function<void(size_t)>        job_;
    size_t                             n_slices_ = 0;
    std::atomic<size_t>                next_{0};
    uint64_t                           generation_ = 0;
    int                                active_     = 0; // workers currently inside drain()
    bool                               stop_       = false;

    void loop()
    {
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_)
                    return;
                seen = generation_;
                active_++;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(mu_);
                active_--;
            }
            cv_.notify_all();
        }
    }

public:
    explicit SlicePool(int n)
    {
        for (int i = 0; i < n; i++)
            threads_.emplace_back([this] { loop(); });
    }
    ~SlicePool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_)
            t.join();
    }
    // workers call job(k) for every k < n_slices exactly once; the job must stay valid until the caller has seen
    // every slice complete (the job signals that itself)
    void start(size_t n_slices, std::function<void(size_t)> job)
    {
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return active_ == 0; }); // nobody is still looking at the previous job
            job_      = std::move(job);
            n_slices_ = n_slices;
            next_.store(0);
            generation_++;
        }
        cv_.notify_all();
    }
    void drain()
    {
        for (;;)
        {
            size_t k = next_.fetch_add(1);
            if (k >= n_slices_)
                return;
            job_(k);
        }
    }
    size_t size() const { return threads_.size(); }
};

// Host half of the packed witness transfer: a pinned buffer of n_slices x kPackStride bytes and the workers that
// classify and pack slices into it. One per prover; a group of shards (one proof over several GPUs) shares one, so
// the witness is read and packed once and every GPU copies the same pinned bytes over its own PCIe link.
struct PackState
{
    std::vector<std::atomic<int>> ready;
    std::vector<uint32_t>         n_full;
    std::atomic<int>              failed{0};
    explicit PackState(size_t n_slices)
        : ready(n_slices)
        , n_full(n_slices, 0)
    {
        for (auto& r : ready)
            r.store(0, std::memory_order_relaxed);
    }
    void wait_all()
    {
        for (auto& r : ready)
            while (!r.load(std::memory_order_acquire))
                std::this_thread::yield();
    }
};

class WitnessPacker
{
public:
    uint8_t*                   pinned   = nullptr; // packed slices (kPackStride each)
    size_t                     n_slices = 0;
    uint32_t                   n_vars   = 0;
    std::unique_ptr<SlicePool> pool;
    std::function<void(size_t)> job_; // kept alive until the next begin()

    // the calling thread must have a current CUDA device (the allocation is portable: valid for every device)
    void create(uint32_t n_vars_)
    {
        n_vars   = n_vars_;
        n_slices = ((size_t)n_vars + kPackWires - 1) / kPackWires;
        KZP_CUDA_CHECK(cudaHostAlloc(&pinned, std::max<size_t>(1, n_slices) * kPackStride, cudaHostAllocPortable));
        const char* te = getenv("KZP_UPLOAD_THREADS");
        int         nt = te ? atoi(te) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        pool.reset(new SlicePool(std::max(nt, 0)));
    }
    void destroy()
    {
        pool.reset();
        if (pinned)
            cudaFreeHost(pinned);
        pinned = nullptr;
    }

    // read(bounce, value_index, count) returns a pointer to `count` plain 32-byte values starting at wire
    // `value_index` (either `bounce`, filled by the call, or the caller's own memory), nullptr on failure.
    // Returns at once when there are workers; `read` and `st` must stay valid until every ready[k] is set.
    template <class Read>
    void begin(Read& read, PackState& st)
    {
        job_ = [this, &read, &st](size_t k) {
            constexpr uint32_t          kChunk = 4096; // values per read: 128 KiB, stays in the reading core's cache
            static thread_local uint8_t bounce[(size_t)kChunk * 32] __attribute__((aligned(64)));
            uint32_t first_wire = (uint32_t)(k * kPackWires);
            uint32_t in_slice   = std::min<uint32_t>(kPackWires, n_vars - first_wire);
            uint8_t* slice      = pinned + k * kPackStride;
            uint32_t nf         = 0;
            for (uint32_t done = 0; done < in_slice && !st.failed.load(std::memory_order_relaxed); done += kChunk)
            {
                uint32_t cnt = std::min(kChunk, in_slice - done);
                const uint8_t* src = read(bounce, (uint64_t)first_wire + done, cnt);
                if (!src)
                {
                    st.failed.store(1);
                    break;
                }
                uint32_t padded = (cnt + kPackGroup - 1) / kPackGroup * kPackGroup;
                if (padded != cnt)
                {
                    if (src != bounce)
                        memcpy(bounce, src, (size_t)cnt * 32);
                    memset(bounce + (size_t)cnt * 32, 0, (size_t)(padded - cnt) * 32);
                    src = bounce;
                }
                pack_values(slice, done, src, padded, &nf);
            }
            st.n_full[k] = nf;
            st.ready[k].store(1, std::memory_order_release);
        };
        if (pool && pool->size() > 0 && n_slices > 1)
            pool->start(n_slices, job_);
        else
            for (size_t k = 0; k < n_slices; k++)
                job_(k);
    }
};

// A fixed crew of host threads, one per shard of a multi-GPU proof: run(fn) calls fn(r) on thread r for every r and
// returns when all are done (kernel launches of different devices must not queue behind one another on one thread).
class Crew
{
    std::vector<std::thread>  threads_;
    std::mutex                mu_;
    std::condition_variable   cv_;
    std::function<void(int)>  fn_;
    uint64_t                  generation_ = 0;
    int                       pending_    = 0;
    bool                      stop_       = false;
    std::atomic<uint32_t>     bar_count_{0};
    std::atomic<uint32_t>     bar_gen_{0};

    void loop(int r)
    {
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_)
                    return;
                seen = generation_;
            }
            fn_(r);
            {
                std::lock_guard<std::mutex> lk(mu_);
                pending_--;
            }
            cv_.notify_all();
        }
    }

public:
    explicit Crew(int n)
    {
        for (int r = 0; r < n; r++)
            threads_.emplace_back([this, r] { loop(r); });
    }
    ~Crew()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_)
            t.join();
    }
    int size() const { return (int)threads_.size(); }
    // fn must not throw
    void run(std::function<void(int)> fn)
    {
        std::unique_lock<std::mutex> lk(mu_);
        fn_      = std::move(fn);
        pending_ = (int)threads_.size();
        generation_++;
        cv_.notify_all();
        cv_.wait(lk, [&] { return pending_ == 0; });
    }
    // to be called by every crew thread from inside fn: returns when all have arrived (they are all busy launching,
    // so the wait is short: spin)
    void barrier()
    {
        uint32_t gen = bar_gen_.load(std::memory_order_acquire);
        if (bar_count_.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)threads_.size())
        {
            bar_count_.store(0, std::memory_order_relaxed);
            bar_gen_.store(gen + 1, std::memory_order_release);
        }
        else
            while (bar_gen_.load(std::memory_order_acquire) == gen)
                std::this_thread::yield();
    }
};

class DeviceProverImpl;
// What a shard of a multi-GPU proof knows about the others (set by ProverGroup once every shard is built).
struct GroupLinks
{
    std::vector<DeviceProverImpl*> shard;
    int  owner_of[3]   = {0, 0, 0}; // shard that computes a, b, c (SpMV + coset-NTT chain) when !dist_ntt
    bool fused_scatter = false;     // the chain's last level stores every shard's slice straight into that shard's buffer
    // Every chain spread over all 2^dist_k shards (ntt_coset_chain_phase): each shard runs 1/2^k of every level of
    // a, b and c; the two transposes between the low-bit and the top-bit partition and the delivery of the H slices
    // are peer stores of the level kernels. Needs a power-of-two shard count, a batched domain size and peer access.
    bool     dist_ntt = false;
    uint32_t dist_k   = 0;
};

class DeviceProverImpl
{
public:
    int      device     = 0;
    int      rank       = 0;
    int      world      = 1;
    uint32_t n_vars     = 0;
    uint32_t n_public   = 0;
    uint32_t domain     = 0;
    uint32_t log_domain = 0;
    uint64_t n_coefs    = 0;

    // st_h: SpMV -> NTT chain -> H MSM.  st_w: witness digit sort + the three G1 witness MSMs (one batch).
    // st_w2: the G2 witness MSM (joins st_w after the sort).  st_copy: witness upload.
    cudaStream_t st_h = nullptr, st_w = nullptr, st_w2 = nullptr, st_copy = nullptr;
    enum
    {
        EV_H2D0, EV_H2D1, EV_H0, EV_SPMV, EV_NTT, EV_HMSM, EV_W0, EV_WSORT, EV_WG1, EV_WG2_0, EV_WG2, EV_XCHG, EV_X1, EV_X2, EV_COUNT
    };
    cudaEvent_t ev[EV_COUNT] = {};

    CoefCsr   csr;
    NttDomain ntt;
    Fr *      d_w = nullptr, *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_h = nullptr;
    Fr *      d_keep_a = nullptr, *d_keep_b = nullptr;
    bool      keep_ab  = false;
    WitnessPacker  own_packer;        // single-GPU prover: its own pinned staging buffer and workers
    WitnessPacker* packer = nullptr;  // the one in use (a group's shards share the group's)
    uint8_t*  d_pack   = nullptr; // device image of the packed slices
    size_t    n_pack_slices = 0;
    // one proof over several GPUs (ProverGroup): which of a, b, c this shard computes (bit 0 / 1 / 2), its slice of
    // the H domain, and the other shards
    uint32_t          own_mask = 7;
    uint64_t          h0 = 0, h1 = 0;
    const GroupLinks* links = nullptr;
    uint64_t  h2d_bytes = 0;     // bytes the last upload moved over PCIe
    uint8_t*  pinned_out = nullptr; // 5 result points

    MsmSort            sort_w, sort_h;
    MsmBases<G1Xyzz>   bases_a, bases_b1, bases_c, bases_h;
    MsmBases<G2Xyzz>   bases_b2;
    MsmScratch<G1Xyzz> sc_a, sc_b1, sc_c, sc_h;
    MsmScratch<G2Xyzz> sc_b2;

    HostVk vk;

    ShardPartials parts;
    ShardPartials sums;       // the five MSM results of the last assemble(), summed over shards
    bool          art_valid = false;
    MsmArtefacts  art;
    ProveTimings  tm;
    bool          witness_resident = false;
    bool          gpu_in_flight    = false;
    bool          tm_pending       = false;
    uint32_t      launches_        = 0;

    void set_device() const { KZP_CUDA_CHECK(cudaSetDevice(device)); }

    void build_csr(const ZkeyHeader& zh)
    {
        // zkey section 4: {u32 m, u32 c, u32 s, Fr coef} x nCoefs (groth16.hpp:33-42). Re-bucketed by
        // (row, matrix) once so the SpMV is a gather with no atomics or locks (the reference scatters
        // under 1024 striped spinlocks, groth16.cpp:137-155).
        uint32_t              N = domain;
        std::vector<uint32_t> ptr(2 * (size_t)N + 1, 0);
        const uint8_t*        p = zh.coefs;
        for (uint64_t i = 0; i < n_coefs; i++)
        {
            uint32_t m, c, s;
            memcpy(&m, p + 44 * i, 4);
            memcpy(&c, p + 44 * i + 4, 4);
            memcpy(&s, p + 44 * i + 8, 4);
            if (c >= N || s >= n_vars)
                throw FormatError("zkey coefficient out of range");
            ptr[2 * (size_t)c + (m ? 1 : 0) + 1]++;
        }
        for (size_t k = 0; k < 2 * (size_t)N; k++)
            ptr[k + 1] += ptr[k];
        std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
        std::vector<uint32_t> wire(std::max<uint64_t>(n_coefs, 1));
        std::vector<uint8_t>  coef(std::max<uint64_t>(n_coefs, 1) * 32);
        for (uint64_t i = 0; i < n_coefs; i++)
        {
            uint32_t m, c, s;
            memcpy(&m, p + 44 * i, 4);
            memcpy(&c, p + 44 * i + 4, 4);
            memcpy(&s, p + 44 * i + 8, 4);
            uint32_t pos = fill[2 * (size_t)c + (m ? 1 : 0)]++;
            wire[pos]    = s;
            memcpy(&coef[(size_t)pos * 32], p + 44 * i + 12, 32);
        }
        csr.n_rows = N;
        csr.nnz    = n_coefs;
        KZP_CUDA_CHECK(cudaMalloc(&csr.row_ptr, ptr.size() * 4));
        KZP_CUDA_CHECK(cudaMalloc(&csr.wire, wire.size() * 4));
        KZP_CUDA_CHECK(cudaMalloc(&csr.coef, coef.size()));
        KZP_CUDA_CHECK(cudaMemcpy(csr.row_ptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice));
        KZP_CUDA_CHECK(cudaMemcpy(csr.wire, wire.data(), wire.size() * 4, cudaMemcpyHostToDevice));
        KZP_CUDA_CHECK(cudaMemcpy(csr.coef, coef.data(), coef.size(), cudaMemcpyHostToDevice));
    }

    // Base columns [first, last) of a zkey point section whose point k belongs to scalar index k + lead
    // (lead = nPublic + 1 for section 8, which holds private wires only; 0 otherwise). Columns without a point
    // are infinity (all-zero), so that every MSM over the witness shares one digit sort.
    template <class XY>
    void make_bases(MsmBases<XY>& b, const uint8_t* sec, uint64_t sec_count, uint32_t lead, uint64_t first,
                    uint64_t last, uint32_t window_bits = 16)
    {
        const size_t         psz = sizeof(typename XY::Affine);
        std::vector<uint8_t> cols((size_t)(last - first) * psz, 0);
        for (uint64_t col = first; col < last; col++)
        {
            if (col < lead || col - lead >= sec_count)
                continue;
            memcpy(&cols[(size_t)(col - first) * psz], sec + (col - lead) * psz, psz);
        }
        msm_bases_create<XY>(b, cols.data(), last - first, false, st_h, window_bits);
    }

    // own_mask_: 7 for a self-contained prover (also the one-process-per-GPU shards, which recompute a, b, c);
    // shared_packer: the staging buffer of the group this shard belongs to (nullptr: make one)
    DeviceProverImpl(const std::string& path, int dev, int rank_, int world_, uint32_t own_mask_ = 7,
                     WitnessPacker* shared_packer = nullptr)
        : device(dev)
        , rank(rank_)
        , world(world_)
        , packer(shared_packer)
        , own_mask(own_mask_)
    {
        if (world < 1 || rank < 0 || rank >= world)
            throw FormatError("invalid shard rank/world");
        // a constructor that throws half-way (CSR out of range, log_n > 27, out of memory) must not leak the streams,
        // events and device buffers it already made: the destructor does not run for a partly built object
        try
        {
            construct(path);
        }
        catch (...)
        {
            teardown();
            throw;
        }
    }

    void construct(const std::string& path)
    {
        MappedFile file(path);
        BinView    bin(file.data(), file.size(), "zkey", 1);
        ZkeyHeader zh = parse_zkey(bin);
        n_vars        = zh.n_vars;
        n_public      = zh.n_public;
        domain        = zh.domain_size;
        n_coefs       = zh.n_coefs;
        while ((1u << log_domain) < domain)
            log_domain++;

        int n_dev = 0;
        KZP_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
        if (device < 0 || device >= n_dev)
            throw CudaError("CUDA device " + std::to_string(device) + " not present");
        set_device();
        // Stream priorities (KZP_PRIO=1: witness streams high, -1: H stream high, 0: equal). Measured on B200: equal
        // priorities are best (12.7 ms vs 13.0 ms); the NTT CTAs use the whole register file, so nothing co-resides
        // with them whatever the priority.
        int prio_lo = 0, prio_hi = 0;
        KZP_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        // A shard of a multi-GPU proof runs the witness streams at high priority: its share of the H MSM no longer
        // hides them (2 GPUs, same box: 6.52 -> 6.25 ms per proof), and the host needs A, B1, C early for its part.
        const char* pe   = getenv("KZP_PRIO");
        int         mode = pe ? atoi(pe) : (packer != nullptr ? 1 : 0);
        int         ph   = mode == 1 ? prio_lo : (mode == -1 ? prio_hi : prio_lo);
        int         pw   = mode == 1 ? prio_hi : (mode == -1 ? prio_lo : prio_lo);
        KZP_CUDA_CHECK(cudaStreamCreateWithPriority(&st_h, cudaStreamNonBlocking, ph));
        KZP_CUDA_CHECK(cudaStreamCreateWithPriority(&st_w, cudaStreamNonBlocking, pw));
        KZP_CUDA_CHECK(cudaStreamCreateWithPriority(&st_w2, cudaStreamNonBlocking, pw));
        KZP_CUDA_CHECK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        for (auto& e : ev)
            KZP_CUDA_CHECK(cudaEventCreate(&e));

        memcpy(vk.alpha1, zh.alpha1, 64);
        memcpy(vk.beta1, zh.beta1, 64);
        memcpy(vk.delta1, zh.delta1, 64);
        memcpy(vk.beta2, zh.beta2, 128);
        memcpy(vk.delta2, zh.delta2, 128);

        if (own_mask) // a shard that computes none of a, b, c needs neither the matrices nor the twiddles
        {
            build_csr(zh);
            ntt_domain_create(ntt, log_domain);
        }
        size_t vec = (size_t)domain * 32;
        KZP_CUDA_CHECK(cudaMalloc(&d_w, (size_t)n_vars * 32));
        KZP_CUDA_CHECK(cudaMalloc(&d_a, vec));
        KZP_CUDA_CHECK(cudaMalloc(&d_b, vec));
        KZP_CUDA_CHECK(cudaMalloc(&d_c, vec));
        KZP_CUDA_CHECK(cudaMalloc(&d_h, vec));
        n_pack_slices = ((size_t)n_vars + kPackWires - 1) / kPackWires;
        if (!packer)
        {
            own_packer.create(n_vars);
            packer = &own_packer;
        }
        KZP_CUDA_CHECK(cudaMalloc(&d_pack, std::max<size_t>(1, n_pack_slices) * kPackStride));
        KZP_CUDA_CHECK(cudaMallocHost(&pinned_out, sizeof(ShardPartials)));

        // this shard's base ranges: wires [w0, w1) of sections 5-8, points [h0, h1) of section 9
        uint64_t w0 = (uint64_t)rank * n_vars / (uint64_t)world, w1 = (uint64_t)(rank + 1) * n_vars / (uint64_t)world;
        h0 = (uint64_t)rank * domain / (uint64_t)world;
        h1 = (uint64_t)(rank + 1) * domain / (uint64_t)world;
        make_bases(bases_a, zh.points_a, n_vars, 0, w0, w1);
        make_bases(bases_b1, zh.points_b1, n_vars, 0, w0, w1);
        make_bases(bases_b2, zh.points_b2, n_vars, 0, w0, w1);
        make_bases(bases_c, zh.points_c, n_vars - n_public - 1, n_public + 1, w0, w1);
        // H scalars are uniform 254-bit values: with the window table resident, c = 20 means 13 mixed additions per
        // scalar instead of 16 (and 2^19 buckets, about 52 entries each at 2^21). Small domains keep c = 16: their
        // cost is the bucket reduction, not the accumulation.   KZP_H_WINDOW overrides (16..22).
        auto envu = [](const char* n, uint32_t d) { const char* e = getenv(n); return e ? (uint32_t)atoi(e) : d; };
        // Measured alone (scripts/window_sweep.py, uniform scalars; c = 16 / 17 / 20): 2^18 points 1.11 / 1.16 / 1.39 ms,
        // 2^19 1.77 / 1.76 / 1.95 ms, 2^20 3.12 / 3.02 / 3.01 ms: the wide window pays from 2^20 points (the slices of
        // a proof sharded over 4 or 8 GPUs stay at c = 16).
        const uint32_t h_window = envu("KZP_H_WINDOW", (h1 - h0) >= (1u << 20) ? 20u : 16u);
        make_bases(bases_h, zh.points_h, domain, 0, h0, h1, h_window);
        msm_sort_create(sort_w, (uint32_t)(w1 - w0), nullptr, (uint32_t)w0);
        msm_sort_create(sort_h, (uint32_t)(h1 - h0), nullptr, (uint32_t)h0, h_window);
        // chunk = sorted entries per accumulate thread. A witness has few non-trivial digits (mostly bits and
        // bytes): small chunks keep enough threads in flight; G2 additions are 3x as long, so smaller still.
        // measured under the final schedule (same box): G1 32 -> 64 and G2 8 -> 16 take the proof from 12.05 to 11.80 ms
        // H with wide windows: about 52 entries per bucket, 64-entry chunks (measured 128 / 64 / 32: 10.90 / 10.70 / 10.89 ms per proof)
        const uint32_t ch_g1 = envu("KZP_CHUNK_G1", 64), ch_g2 = envu("KZP_CHUNK_G2", 16), ch_h = envu("KZP_CHUNK_H", h_window >= 19 ? ((h1 - h0) >= (1u << 21) ? 64 : 32) : 0);
        msm_scratch_create(sc_a, sort_w, ch_g1);
        msm_scratch_create(sc_b1, sort_w, ch_g1);
        msm_scratch_create(sc_c, sort_w, ch_g1);
        msm_scratch_create(sc_b2, sort_w, ch_g2);
        msm_scratch_create(sc_h, sort_h, ch_h);
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
    }

    ~DeviceProverImpl() { teardown(); }

    void teardown()
    {
        if (cudaSetDevice(device) != cudaSuccess)
        {
            cudaGetLastError();
            own_packer.pool.reset();
            return; // no such device: nothing was allocated
        }
        cudaDeviceSynchronize();
        own_packer.destroy();
        msm_bases_destroy(bases_a);
        msm_bases_destroy(bases_b1);
        msm_bases_destroy(bases_b2);
        msm_bases_destroy(bases_c);
        msm_bases_destroy(bases_h);
        msm_scratch_destroy(sc_a);
        msm_scratch_destroy(sc_b1);
        msm_scratch_destroy(sc_b2);
        msm_scratch_destroy(sc_c);
        msm_scratch_destroy(sc_h);
        msm_sort_destroy(sort_w);
        msm_sort_destroy(sort_h);
        ntt_domain_destroy(ntt);
        cudaFree(csr.row_ptr);
        cudaFree(csr.wire);
        cudaFree(csr.coef);
        cudaFree(d_w);
        cudaFree(d_pack);
        cudaFree(d_a);
        cudaFree(d_b);
        cudaFree(d_c);
        cudaFree(d_h);
        cudaFree(d_keep_a);
        cudaFree(d_keep_b);
        cudaFreeHost(pinned_out);
        for (auto& e : ev)
            if (e)
                cudaEventDestroy(e);
        for (cudaStream_t* s : {&st_h, &st_w, &st_w2, &st_copy})
            if (*s)
                cudaStreamDestroy(*s);
        cudaGetLastError();
    }

    // Host -> pinned staging -> device, pipelined: worker threads classify and pack slices of kPackWires values into
    // the pinned buffer (reading them from memory, or with pread() straight from the witness file: no page faults
    // on a fresh mapping) while the calling thread hands every finished slice's packed prefix to the copy engine;
    // k_witness_expand then rebuilds the plain vector in HBM.   KZP_UPLOAD_THREADS (default min(16, cores)).
    // Measured on the B200 host (16 cores), 43 MB keyless witness: plain copy 1.15 ms (DMA-bound at ~38 GB/s);
    // packed: 3.7 MB over the bus, 0.5 ms with 16 workers (pread-bound), end to end 13.6 -> 13.0 ms.

    // Device half: copies every slice of packer->pinned as soon as it is packed, then expands. Safe to run for several
    // shards at once on different threads (they read the same pinned bytes).
    void enqueue_packed(PackState& ps)
    {
        set_device();
        const size_t n_slices = n_pack_slices;
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_H2D0], st_copy));
        cudaError_t err   = cudaSuccess;
        uint64_t    moved = 0;
        for (size_t k = 0; k < n_slices; k++)
        {
            while (!ps.ready[k].load(std::memory_order_acquire))
                std::this_thread::yield();
            size_t len = kPackHead + (size_t)ps.n_full[k] * 32;
            if (err == cudaSuccess && !ps.failed.load())
                err = cudaMemcpyAsync(d_pack + k * kPackStride, packer->pinned + k * kPackStride, len, cudaMemcpyHostToDevice, st_copy);
            moved += len;
        }
        KZP_CUDA_CHECK(err);
        if (ps.failed.load())
            throw LoadError("reading the witness failed");
        if (n_slices > 0)
        {
            k_witness_expand<<<(unsigned int)n_slices, 1024, 0, st_copy>>>(d_pack, reinterpret_cast<uint4*>(d_w), n_vars);
            KZP_CUDA_CHECK(cudaGetLastError());
        }
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_H2D1], st_copy));
        h2d_bytes        = moved;
        witness_resident = true;
    }

    template <class Read>
    void upload_with(Read&& read)
    {
        set_device();
        PackState ps(n_pack_slices);
        double    dbg_t0 = now_ms();
        packer->begin(read, ps);
        try
        {
            enqueue_packed(ps);
        }
        catch (...)
        {
            ps.wait_all(); // the workers still use `read` and `ps`
            throw;
        }
        if (getenv("KZP_DEBUG_UPLOAD"))
        {
            double t1 = now_ms();
            cudaStreamSynchronize(st_copy);
            fprintf(stderr, "[kzp upload] staged+enqueued %.3f ms, DMA+expand drained +%.3f ms, %zu slices, %zu workers, %.2f MB moved\n",
                    t1 - dbg_t0, now_ms() - t1, n_pack_slices, packer->pool ? packer->pool->size() : 0, h2d_bytes / 1e6);
        }
    }

    void upload(const uint8_t* values, uint64_t n)
    {
        if (n < n_vars)
            throw FormatError("witness has fewer values than the zkey has variables");
        upload_with(witness_reader_mem(values));
    }

    void upload_fd(int fd, uint64_t file_offset, uint64_t n)
    {
        if (n < n_vars)
            throw FormatError("witness has fewer values than the zkey has variables");
        upload_with(witness_reader_fd(fd, file_offset));
    }

    Fr* vec_ptr(int i) const { return i == 0 ? d_a : (i == 1 ? d_b : d_c); }

    // Enqueues the whole GPU part of one proof (no host synchronisation).
    void launch_gpu()
    {
        launch_front();
        launch_witness();
        launch_h();
    }

    void begin_proof()
    {
        if (!witness_resident)
            throw FormatError("no witness uploaded");
        set_device();
        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_h, ev[EV_H2D1], 0));
        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_w, ev[EV_H2D1], 0));
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_H0], st_h));
        launches_ = 0;
    }

    // SpMV and the coset-NTT chains of the vectors this prover owns (stream H, up to EV_NTT).
    // In a group (links != nullptr) the chain output is handed to the other shards: either the chain's last level
    // stores each shard's slice [h0, h1) of the H domain straight into that shard's buffer over NVLink (fused_scatter),
    // or the finished vector is cut up with peer copies.
    void launch_front()
    {
        begin_proof();
        uint32_t ntt_kernels = 0;
        if (own_mask)
        {
            spmv_abc(csr, d_w, d_a, d_b, d_c, st_h, own_mask);
            if (keep_ab)
            {
                KZP_CUDA_CHECK(cudaMemcpyAsync(d_keep_a, d_a, (size_t)domain * 32, cudaMemcpyDeviceToDevice, st_h));
                KZP_CUDA_CHECK(cudaMemcpyAsync(d_keep_b, d_b, (size_t)domain * 32, cudaMemcpyDeviceToDevice, st_h));
            }
        }
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_SPMV], st_h));
        if (own_mask)
        {
            Fr* vecs[3];
            int owned[3], cnt = 0;
            for (int i = 0; i < 3; i++)
                if (own_mask & (1u << i))
                {
                    owned[cnt]  = i;
                    vecs[cnt++] = vec_ptr(i);
                }
            if (links && links->fused_scatter)
            {
                NttRoute rt = {};
                rt.store    = kNttStoreBounds;
                rt.world    = (int)links->shard.size();
                for (int r = 0; r < rt.world; r++)
                {
                    rt.bound[r] = (uint32_t)links->shard[r]->h0;
                    for (int i = 0; i < cnt; i++)
                        rt.dst[i][r] = links->shard[r]->vec_ptr(owned[i]);
                }
                rt.bound[rt.world] = domain;
                ntt_kernels        = ntt_coset_chain(ntt, vecs, cnt, st_h, &rt);
            }
            else
            {
                ntt_kernels = ntt_coset_chain(ntt, vecs, cnt, st_h);
                if (links)
                    for (size_t r = 0; r < links->shard.size(); r++)
                    {
                        const DeviceProverImpl* peer = links->shard[r];
                        if (peer == this || peer->h1 == peer->h0)
                            continue;
                        for (int i = 0; i < cnt; i++)
                            KZP_CUDA_CHECK(cudaMemcpyPeerAsync(peer->vec_ptr(owned[i]) + peer->h0, peer->device,
                                                               vecs[i] + peer->h0, device, (size_t)(peer->h1 - peer->h0) * 32, st_h));
                    }
            }
        }
        if (links)
            KZP_CUDA_CHECK(cudaEventRecord(ev[EV_XCHG], st_h)); // this shard's vectors have reached every other shard
        else
            h_pointwise(d_a, d_b, d_c, d_h, domain, st_h);
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_NTT], st_h));
        launches_ += (own_mask ? 1 : 0) + ntt_kernels;
    }

    // The distributed front (links->dist_ntt): phase 0 = this shard's rows of the SpMV + the inverse levels, phase 1 =
    // the fused middle level, phase 2 = the forward levels. Every phase ends in peer stores into the other shards'
    // buffers and an event; the next phase waits for ALL shards' events, which ProverGroup guarantees to have been
    // recorded for this proof (host barrier between the phases).
    void launch_dist(int phase)
    {
        const uint32_t k = links->dist_k;
        Fr* const      vecs[3] = {d_a, d_b, d_c};
        Fr*            dst[3][kNttMaxShards] = {};
        for (size_t r = 0; r < links->shard.size(); r++)
            for (int i = 0; i < 3; i++)
                dst[i][r] = links->shard[r]->vec_ptr(i);
        if (phase == 0)
        {
            begin_proof();
            spmv_abc(csr, d_w, d_a, d_b, d_c, st_h, 7, k, (uint32_t)rank);
            KZP_CUDA_CHECK(cudaEventRecord(ev[EV_SPMV], st_h));
            launches_ += 1;
        }
        else
        {
            set_device();
            for (const DeviceProverImpl* peer : links->shard)
                if (peer != this)
                    KZP_CUDA_CHECK(cudaStreamWaitEvent(st_h, peer->ev[phase == 1 ? EV_X1 : EV_X2], 0));
        }
        launches_ += ntt_coset_chain_phase(ntt, vecs, 3, st_h, phase, k, (uint32_t)rank, dst);
        KZP_CUDA_CHECK(cudaEventRecord(ev[phase == 0 ? EV_X1 : (phase == 1 ? EV_X2 : EV_XCHG)], st_h));
        if (phase == 2)
            KZP_CUDA_CHECK(cudaEventRecord(ev[EV_NTT], st_h));
    }

    // Everything on the witness streams: one digit sort of the witness, then A, B1, C as one G1 batch; B2 on its own
    // stream. Call after EV_NTT has been recorded for this proof.
    void launch_witness()
    {
        set_device();
        const uint32_t* w = reinterpret_cast<const uint32_t*>(d_w);
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_W0], st_w));
        msm_sort_run(sort_w, w, st_w);
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_WSORT], st_w));
        // KZP_WDELAY: hold witness bucket work back until the NTT chain is done, so that it overlaps the H digit
        // sort (integer pipe idle) and the H MSM instead of sharing the pipe with the NTT. 0 = nothing held, 1 = all
        // four, 2 = only B2, 3 = only A/B1/C (default). Measured, same box: 12.32 / 12.76 / 12.52 / 12.00 ms per proof.
        static const int wdelay = getenv("KZP_WDELAY") ? atoi(getenv("KZP_WDELAY")) : 3;
        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_w2, ev[EV_WSORT], 0));
        if (wdelay == 1 || wdelay == 2)
            KZP_CUDA_CHECK(cudaStreamWaitEvent(st_w2, ev[EV_NTT], 0));
        if (wdelay == 1 || wdelay == 3)
            KZP_CUDA_CHECK(cudaStreamWaitEvent(st_w, ev[EV_NTT], 0));
        {
            const MsmBases<G2Xyzz>* b[1] = {&bases_b2};
            MsmScratch<G2Xyzz>*     s[1] = {&sc_b2};
            KZP_CUDA_CHECK(cudaEventRecord(ev[EV_WG2_0], st_w2));
            msm_reduce_batch<G2Xyzz>(sort_w, b, s, 1, st_w2);
            KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 512, sc_b2.result, 256, cudaMemcpyDeviceToHost, st_w2));
            KZP_CUDA_CHECK(cudaEventRecord(ev[EV_WG2], st_w2));
        }
        {
            const MsmBases<G1Xyzz>* b[3] = {&bases_a, &bases_b1, &bases_c};
            MsmScratch<G1Xyzz>*     s[3] = {&sc_a, &sc_b1, &sc_c};
            msm_reduce_batch<G1Xyzz>(sort_w, b, s, 3, st_w);
            KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 0, sc_a.result, 128, cudaMemcpyDeviceToHost, st_w));
            KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 128, sc_b1.result, 128, cudaMemcpyDeviceToHost, st_w));
            KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 256, sc_c.result, 128, cudaMemcpyDeviceToHost, st_w));
            KZP_CUDA_CHECK(cudaEventRecord(ev[EV_WG1], st_w));
        }
        launches_ += 2 * kMsmSortLaunches + 3 * kMsmReduceLaunches; // + 1 per upload (expand)
    }

    // The H MSM. In a group: first wait until the owners of a, b, c have delivered this shard's slices (their EV_XCHG
    // must have been recorded for THIS proof before this call: ProverGroup puts a host barrier between the two
    // halves), then the pointwise step on the slice only.
    void launch_h()
    {
        set_device();
        if (links)
        {
            if (links->dist_ntt)
            {
                for (const DeviceProverImpl* peer : links->shard)
                    if (peer != this)
                        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_h, peer->ev[EV_XCHG], 0));
            }
            else
                for (int i = 0; i < 3; i++)
                {
                    const DeviceProverImpl* o = links->shard[links->owner_of[i]];
                    bool seen = o == this;
                    for (int j = 0; j < i; j++)
                        seen = seen || links->shard[links->owner_of[j]] == o;
                    if (!seen)
                        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_h, o->ev[EV_XCHG], 0));
                }
            if (h1 > h0)
                h_pointwise(d_a + h0, d_b + h0, d_c + h0, d_h + h0, h1 - h0, st_h);
        }
        {
            const MsmBases<G1Xyzz>* b[1] = {&bases_h};
            MsmScratch<G1Xyzz>*     s[1] = {&sc_h};
            msm_sort_run(sort_h, reinterpret_cast<const uint32_t*>(d_h), st_h);
            msm_reduce_batch<G1Xyzz>(sort_h, b, s, 1, st_h);
        }
        KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 384, sc_h.result, 128, cudaMemcpyDeviceToHost, st_h));
        KZP_CUDA_CHECK(cudaEventRecord(ev[EV_HMSM], st_h));
        launches_ += 1; // pointwise
        gpu_in_flight = true;
    }

    // witness-side MSM results are on the host after these (A, B1, C / B2); the H stream may still be running
    void wait_witness_g1()
    {
        if (!gpu_in_flight)
            throw FormatError("no proof in flight");
        set_device();
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_w));
    }
    void wait_witness_g2()
    {
        if (!gpu_in_flight)
            throw FormatError("no proof in flight");
        set_device();
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_w2));
    }

    void wait_gpu()
    {
        if (!gpu_in_flight)
            throw FormatError("no proof in flight");
        set_device();
        gpu_in_flight = false;
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_w));
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_w2));
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_h));
        memcpy(parts.bytes, pinned_out, sizeof(parts.bytes));
        tm_pending = true; // the event queries cost a few microseconds each: made when somebody asks (timings())
    }

    const ProveTimings& timings()
    {
        if (!tm_pending)
            return tm;
        tm_pending = false;
        set_device();
        auto el = [&](int a, int b) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[a], ev[b]);
            return ms;
        };
        tm.h2d_ms       = el(EV_H2D0, EV_H2D1);
        tm.h2d_mbytes   = (float)(h2d_bytes / 1e6);
        tm.spmv_ms      = el(EV_H0, EV_SPMV);
        tm.ntt_ms       = el(EV_SPMV, EV_NTT);
        tm.msm_h_ms     = el(EV_NTT, EV_HMSM);
        tm.msm_wsort_ms = el(EV_W0, EV_WSORT);
        tm.msm_wg1_ms   = el(EV_WSORT, EV_WG1);
        tm.msm_wg2_ms   = el(EV_WG2_0, EV_WG2);
        tm.gpu_ms       = std::max(std::max(el(EV_H0, EV_HMSM), el(EV_H0, EV_WG1)), el(EV_H0, EV_WG2));
        tm.kernel_launches = launches_;
        return tm;
    }

    void run_gpu()
    {
        launch_gpu();
        wait_gpu();
    }

    std::string assemble(const ShardPartials* ps, int count, const uint8_t* r32, const uint8_t* s32)
    {
        double     t0 = now_ms();
        BlindTerms bt;
        compute_blind_terms(vk, r32, s32, bt);
        std::string j       = assemble_with_terms(vk, ps, count, bt, &sums);
        art_valid           = false;
        tm.assemble_host_ms = (float)(now_ms() - t0);
        return j;
    }

    // upload + GPU + assembly. Host work is ordered by what it depends on: the r/s-only terms right after the
    // launch, everything that needs A, B1, C, B2 (two 254-bit scalar multiplications, two affine conversions,
    // six decimal strings) as soon as the witness streams are done, and only "+ H, to affine, print" after the H MSM.
    template <class Upload>
    std::string prove_with(Upload&& do_upload, const uint8_t* r32, const uint8_t* s32)
    {
        static const bool timeline = getenv("KZP_DEBUG_TIMELINE") != nullptr;
        double t0 = now_ms();
        do_upload();
        launch_gpu();
        double     ta = now_ms();
        BlindTerms bt;
        compute_blind_terms(vk, r32, s32, bt);
        double tb = now_ms();
        EarlyProof ep;
        wait_witness_g1();
        double tc = now_ms();
        {
            HG1 A, B1, C;
            memcpy(&A, pinned_out + 0, 128);
            memcpy(&B1, pinned_out + 128, 128);
            memcpy(&C, pinned_out + 256, 128);
            assemble_early_g1(vk, bt, A, B1, C, ep);
        }
        wait_witness_g2();
        {
            HG2 B2;
            memcpy(&B2, pinned_out + 512, 256);
            assemble_early_g2(vk, bt, B2, ep);
        }
        double td = now_ms();
        wait_gpu();
        double t1 = now_ms();
        HG1    H;
        memcpy(&H, parts.bytes + 384, 128);
        std::string j       = assemble_final(ep, H, &sums);
        art_valid           = false;
        tm.assemble_host_ms = (float)(now_ms() - t1);
        tm.total_host_ms    = (float)(now_ms() - t0);
        if (timeline)
            fprintf(stderr, "[kzp timeline] upload+launch %.3f | blind terms +%.3f | A, B1, C waited +%.3f | early assembly (incl. waiting for B2) +%.3f | "
                            "H waited +%.3f | final +%.3f = %.3f ms\n",
                    ta - t0, tb - ta, tc - tb, td - tc, t1 - td, now_ms() - t1, now_ms() - t0);
        return j;
    }

    const MsmArtefacts& artefacts()
    {
        if (!art_valid)
        {
            artefacts_from_sums(sums, art);
            art_valid = true;
        }
        return art;
    }
};

// ------------------------------------------------------------------ one proof over several GPUs, one process
// the reference's only entry is FullProver::prove under one mutex (prover-service/src/request_handler/
// prover_handler.rs:266-269), so a drop-in has to shard INSIDE that call. Shard r lives on devices[r] and holds base
// range r of every MSM section. The three coset-NTT chains are independent (the reference runs them as three
// std::async tasks, groth16.cpp:172-262): each is computed by one shard and its output is delivered slice by slice to
// the shards that need it for their part of the H MSM (peer stores fused into the chain's last level, or peer copies).
// The 768-byte partial results meet on the host, which sums them: no communicator, no collective library.
class ProverGroup
{
public:
    std::vector<std::unique_ptr<DeviceProverImpl>> sh;
    std::vector<int>      devices;
    GroupLinks            links;
    WitnessPacker         packer;
    std::unique_ptr<Crew> crew;
    ShardPartials         parts; // sum over the shards, in one shard's format
    ShardPartials         sums;
    bool                  art_valid = false;
    MsmArtefacts          art;
    ProveTimings          tm;
    bool                  gpu_in_flight = false;
    bool                  tm_pending    = false;

    static void owners_for(int world, int (&owner_of)[3])
    {
        if (world == 1)
            owner_of[0] = owner_of[1] = owner_of[2] = 0;
        else if (world == 2)
            owner_of[0] = 0, owner_of[1] = owner_of[2] = 1; // c = a o b needs both row sums anyway: b rides along
        else
            owner_of[0] = 0, owner_of[1] = 1, owner_of[2] = 2;
    }

    ProverGroup(const std::string& path, const std::vector<int>& devs)
        : devices(devs)
    {
        const int world = (int)devices.size();
        if (world < 1 || world > kNttMaxShards)
            throw FormatError("a prover group has 1 to 8 shards");
        int n_dev = 0;
        KZP_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
        for (int d : devices)
            if (d < 0 || d >= n_dev)
                throw CudaError("CUDA device " + std::to_string(d) + " not present");
        uint32_t n_vars = 0;
        {
            MappedFile file(path);
            BinView    bin(file.data(), file.size(), "zkey", 1);
            n_vars = parse_zkey(bin).n_vars;
        }
        owners_for(world, links.owner_of);
        // KZP_GROUP_NTT=dist (default when possible) | chain: every chain spread over all shards, or one chain per shard
        const char* ne       = getenv("KZP_GROUP_NTT");
        const bool  want_dist = !(ne && strcmp(ne, "chain") == 0) && world >= 2 && (world & (world - 1)) == 0;
        try
        {
            KZP_CUDA_CHECK(cudaSetDevice(devices[0]));
            packer.create(n_vars);
            crew.reset(new Crew(world));
            sh.resize(world);
            std::vector<std::exception_ptr> err(world);
            crew->run([&](int r) {
                try
                {
                    // which vectors this shard must be able to compute: all three when the chains may be spread over
                    // the shards (decided below, once the domain size and the peer topology are known)
                    uint32_t mask = want_dist ? 7u : 0u;
                    for (int i = 0; i < 3; i++)
                        if (links.owner_of[i] == r)
                            mask |= 1u << i;
                    sh[r].reset(new DeviceProverImpl(path, devices[r], r, world, mask, &packer));
                }
                catch (...)
                {
                    err[r] = std::current_exception();
                }
            });
            for (auto& e : err)
                if (e)
                    std::rethrow_exception(e);
            // peer mappings for the slice exchange
            bool all_peers = true;
            for (int i = 0; i < world; i++)
                for (int j = 0; j < world; j++)
                {
                    if (devices[i] == devices[j])
                        continue;
                    int can = 0;
                    KZP_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
                    if (!can)
                    {
                        all_peers = false;
                        continue;
                    }
                    KZP_CUDA_CHECK(cudaSetDevice(devices[i]));
                    cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled)
                        cudaGetLastError();
                    else
                        KZP_CUDA_CHECK(e);
                }
            for (auto& p : sh)
                links.shard.push_back(p.get());
            const char* se      = getenv("KZP_GROUP_SCATTER");
            links.fused_scatter = (se ? atoi(se) != 0 : true) && all_peers && ntt_chain_is_batched(sh[0]->log_domain);
            uint32_t k = 0;
            while ((1 << k) < world)
                k++;
            // the top-bit partition must coincide with the H slices: 2^k | domain, and every slice at least one tile
            links.dist_ntt = want_dist && all_peers && ntt_chain_is_batched(sh[0]->log_domain) && k <= 3 &&
                             sh[0]->log_domain >= 11 + k;
            links.dist_k = k;
            if (!links.dist_ntt)
                for (int r = 0; r < world; r++) // back to one chain per owner
                {
                    uint32_t mask = 0;
                    for (int i = 0; i < 3; i++)
                        if (links.owner_of[i] == r)
                            mask |= 1u << i;
                    sh[r]->own_mask = mask;
                }
            for (auto& p : sh)
                p->links = &links;
        }
        catch (...)
        {
            destroy();
            throw;
        }
    }

    ~ProverGroup() { destroy(); }

    void destroy()
    {
        sh.clear(); // each shard synchronises its device first
        crew.reset();
        if (!devices.empty() && cudaSetDevice(devices[0]) == cudaSuccess)
            packer.destroy();
        else
            packer.pool.reset();
        cudaGetLastError();
    }

    int world() const { return (int)sh.size(); }

    // One wake-up of the crew per proof: thread r copies + expands the witness on its GPU (when ps != nullptr), launches
    // its prefix, meets the others (every EV_XCHG of this proof is recorded), launches its H MSM.
    void run_crew(PackState* ps)
    {
        const int                       n = world();
        std::vector<std::exception_ptr> err(n);
        // failed[s]: some shard threw while launching stage s. A thread looks only at the flags of stages that ended
        // before the barrier it has just left — those are final, so every thread takes the same decision and nobody is
        // left waiting at the next barrier for a thread that has already given up.
        std::atomic<int> failed[4];
        for (auto& f : failed)
            f.store(0);
        crew->run([&](int r) {
            try
            {
                if (ps)
                    sh[r]->enqueue_packed(*ps);
                if (links.dist_ntt)
                    sh[r]->launch_dist(0);
                else
                {
                    sh[r]->launch_front();
                    sh[r]->launch_witness();
                }
            }
            catch (...)
            {
                err[r] = std::current_exception();
                failed[0].store(1);
            }
            // every barrier: the events the next stage waits on have been recorded by all shards for this proof
            for (int stage = 1; stage <= (links.dist_ntt ? 3 : 1); stage++)
            {
                crew->barrier();
                bool give_up = false;
                for (int done = 0; done < stage; done++)
                    give_up = give_up || failed[done].load() != 0;
                if (give_up)
                    return;
                try
                {
                    if (!links.dist_ntt || stage == 3)
                        sh[r]->launch_h();
                    else
                    {
                        sh[r]->launch_dist(stage);
                        if (stage == 2)
                            sh[r]->launch_witness();
                    }
                }
                catch (...)
                {
                    if (!err[r])
                        err[r] = std::current_exception();
                    failed[stage].store(1);
                }
            }
        });
        if (ps)
            ps->wait_all();
        for (auto& e : err)
            if (e)
            {
                for (auto& p : sh) // leave no work in flight behind a failed proof
                    if (cudaSetDevice(p->device) == cudaSuccess)
                        cudaDeviceSynchronize();
                for (auto& p : sh)
                    p->gpu_in_flight = false;
                std::rethrow_exception(e);
            }
        gpu_in_flight = true;
    }

    template <class Read>
    void upload_with(Read&& read)
    {
        PackState ps(packer.n_slices);
        packer.begin(read, ps);
        std::vector<std::exception_ptr> err(world());
        crew->run([&](int r) {
            try
            {
                sh[r]->enqueue_packed(ps);
            }
            catch (...)
            {
                err[r] = std::current_exception();
            }
        });
        ps.wait_all();
        for (auto& e : err)
            if (e)
                std::rethrow_exception(e);
    }
    void check_n(uint64_t n) const
    {
        if (n < sh[0]->n_vars)
            throw FormatError("witness has fewer values than the zkey has variables");
    }

    void wait_witness_g1(HG1& A, HG1& B1, HG1& C)
    {
        if (!gpu_in_flight)
            throw FormatError("no proof in flight");
        HG1::set_inf(A);
        HG1::set_inf(B1);
        HG1::set_inf(C);
        for (auto& p : sh)
        {
            p->wait_witness_g1();
            HG1 t;
            memcpy(&t, p->pinned_out + 0, 128);
            HG1::add(A, t);
            memcpy(&t, p->pinned_out + 128, 128);
            HG1::add(B1, t);
            memcpy(&t, p->pinned_out + 256, 128);
            HG1::add(C, t);
        }
    }
    void wait_witness_g2(HG2& B2)
    {
        HG2::set_inf(B2);
        for (auto& p : sh)
        {
            p->wait_witness_g2();
            HG2 t2;
            memcpy(&t2, p->pinned_out + 512, 256);
            HG2::add(B2, t2);
        }
    }

    // all shards done; parts = the sums
    void wait_gpu()
    {
        if (!gpu_in_flight)
            throw FormatError("no proof in flight");
        gpu_in_flight = false;
        HG1 A, B1, C, H;
        HG2 B2;
        HG1::set_inf(A);
        HG1::set_inf(B1);
        HG1::set_inf(C);
        HG1::set_inf(H);
        HG2::set_inf(B2);
        for (auto& p : sh)
        {
            p->wait_gpu();
            HG1 g;
            HG2 g2;
            memcpy(&g, p->parts.bytes + 0, 128);
            HG1::add(A, g);
            memcpy(&g, p->parts.bytes + 128, 128);
            HG1::add(B1, g);
            memcpy(&g, p->parts.bytes + 256, 128);
            HG1::add(C, g);
            memcpy(&g, p->parts.bytes + 384, 128);
            HG1::add(H, g);
            memcpy(&g2, p->parts.bytes + 512, 256);
            HG2::add(B2, g2);
        }
        memcpy(parts.bytes + 0, &A, 128);
        memcpy(parts.bytes + 128, &B1, 128);
        memcpy(parts.bytes + 256, &C, 128);
        memcpy(parts.bytes + 384, &H, 128);
        memcpy(parts.bytes + 512, &B2, 256);
        tm_pending = true;
    }

    const ProveTimings& timings()
    {
        if (!tm_pending)
            return tm;
        tm_pending = false;
        ProveTimings t;
        for (auto& p : sh)
        {
            const ProveTimings& q = p->timings();
            t.h2d_ms       = std::max(t.h2d_ms, q.h2d_ms);
            t.spmv_ms      = std::max(t.spmv_ms, q.spmv_ms);
            t.ntt_ms       = std::max(t.ntt_ms, q.ntt_ms);
            t.msm_h_ms     = std::max(t.msm_h_ms, q.msm_h_ms);
            t.msm_wsort_ms = std::max(t.msm_wsort_ms, q.msm_wsort_ms);
            t.msm_wg1_ms   = std::max(t.msm_wg1_ms, q.msm_wg1_ms);
            t.msm_wg2_ms   = std::max(t.msm_wg2_ms, q.msm_wg2_ms);
            t.gpu_ms       = std::max(t.gpu_ms, q.gpu_ms);
            t.h2d_mbytes += q.h2d_mbytes;
            t.kernel_launches += q.kernel_launches;
        }
        t.assemble_host_ms = tm.assemble_host_ms;
        t.total_host_ms    = tm.total_host_ms;
        tm                 = t;
        return tm;
    }

    void run_gpu()
    {
        run_crew(nullptr);
        wait_gpu();
    }

    std::string assemble(const ShardPartials* ps, int count, const uint8_t* r32, const uint8_t* s32)
    {
        double     t0 = now_ms();
        BlindTerms bt;
        compute_blind_terms(sh[0]->vk, r32, s32, bt);
        std::string j       = assemble_with_terms(sh[0]->vk, ps, count, bt, &sums);
        art_valid           = false;
        tm.assemble_host_ms = (float)(now_ms() - t0);
        return j;
    }

    // same overlap of host work with the GPUs as DeviceProverImpl::prove_with
    template <class Start>
    std::string prove_with(Start&& start, const uint8_t* r32, const uint8_t* s32)
    {
        static const bool timeline = getenv("KZP_DEBUG_TIMELINE") != nullptr;
        double t0 = now_ms();
        start();
        double     ta = now_ms();
        BlindTerms bt;
        compute_blind_terms(sh[0]->vk, r32, s32, bt);
        double     tb = now_ms(), tc;
        EarlyProof ep;
        {
            HG1 A, B1, C;
            HG2 B2;
            wait_witness_g1(A, B1, C);
            tc = now_ms();
            assemble_early_g1(sh[0]->vk, bt, A, B1, C, ep);
            wait_witness_g2(B2);
            assemble_early_g2(sh[0]->vk, bt, B2, ep);
        }
        double td = now_ms();
        wait_gpu();
        double t1 = now_ms();
        if (timeline)
            fprintf(stderr, "[kzp group timeline] upload+launch %.3f | blind terms +%.3f | A, B1, C waited +%.3f | early assembly (incl. waiting for B2) "
                            "+%.3f | H waited +%.3f\n",
                    ta - t0, tb - ta, tc - tb, td - tc, t1 - td);
        HG1    H;
        memcpy(&H, parts.bytes + 384, 128);
        std::string j       = assemble_final(ep, H, &sums);
        art_valid           = false;
        tm.assemble_host_ms = (float)(now_ms() - t1);
        tm.total_host_ms    = (float)(now_ms() - t0);
        return j;
    }

    template <class Read>
    std::string prove_upload(Read&& read, const uint8_t* r32, const uint8_t* s32)
    {
        PackState ps(packer.n_slices);
        return prove_with(
            [&] {
                packer.begin(read, ps);
                try
                {
                    run_crew(&ps);
                }
                catch (...)
                {
                    ps.wait_all();
                    throw;
                }
            },
            r32, s32);
    }

    const MsmArtefacts& artefacts()
    {
        if (!art_valid)
        {
            artefacts_from_sums(sums, art);
            art_valid = true;
        }
        return art;
    }

    // every shard holds its own slice of the H coefficients
    void copy_h(uint8_t* out) const
    {
        for (auto& p : sh)
        {
            p->set_device();
            if (p->h1 > p->h0)
                KZP_CUDA_CHECK(cudaMemcpy(out + p->h0 * 32, p->d_h + p->h0, (size_t)(p->h1 - p->h0) * 32, cudaMemcpyDeviceToHost));
        }
    }
};

// ------------------------------------------------------------------ facade
DeviceProver::DeviceProver(const std::string& zkey_path, int device, int shard_rank, int shard_world)
    : impl_(new DeviceProverImpl(zkey_path, device, shard_rank, shard_world))
{
}
DeviceProver::DeviceProver(const std::string& zkey_path, const std::vector<int>& devices)
    : group_(new ProverGroup(zkey_path, devices))
{
}
DeviceProver::~DeviceProver() {}
DeviceProverImpl& DeviceProver::first() const { return group_ ? *group_->sh[0] : *impl_; }
uint32_t DeviceProver::n_vars() const { return first().n_vars; }
uint32_t DeviceProver::n_public() const { return first().n_public; }
uint32_t DeviceProver::domain_size() const { return first().domain; }
uint64_t DeviceProver::n_coefs() const { return first().n_coefs; }
int      DeviceProver::device() const { return first().device; }
int      DeviceProver::group_size() const { return group_ ? group_->world() : 1; }
bool     DeviceProver::group_fused_exchange() const { return group_ && (group_->links.fused_scatter || group_->links.dist_ntt); }
bool     DeviceProver::group_distributed_ntt() const { return group_ && group_->links.dist_ntt; }
void     DeviceProver::upload_witness(const uint8_t* values, uint64_t n)
{
    if (group_)
    {
        group_->check_n(n);
        group_->upload_with(witness_reader_mem(values));
    }
    else
        impl_->upload(values, n);
}
void DeviceProver::upload_witness_fd(int fd, uint64_t file_offset, uint64_t n)
{
    if (group_)
    {
        group_->check_n(n);
        group_->upload_with(witness_reader_fd(fd, file_offset));
    }
    else
        impl_->upload_fd(fd, file_offset, n);
}
void DeviceProver::run_gpu()
{
    if (group_)
        group_->run_gpu();
    else
        impl_->run_gpu();
}
std::string DeviceProver::prove_resident(const uint8_t* r32, const uint8_t* s32)
{
    if (group_)
        return group_->prove_with([&] { group_->run_crew(nullptr); }, r32, s32);
    return impl_->prove_with([] {}, r32, s32);
}
const ShardPartials& DeviceProver::partials() const { return group_ ? group_->parts : impl_->parts; }
std::string DeviceProver::assemble(const ShardPartials* parts, int count, const uint8_t* r32,
                                   const uint8_t* s32)
{
    return group_ ? group_->assemble(parts, count, r32, s32) : impl_->assemble(parts, count, r32, s32);
}
std::string DeviceProver::prove(const uint8_t* values, uint64_t n, const uint8_t* r32, const uint8_t* s32)
{
    if (group_)
    {
        group_->check_n(n);
        return group_->prove_upload(witness_reader_mem(values), r32, s32);
    }
    return impl_->prove_with([&] { impl_->upload(values, n); }, r32, s32);
}
std::string DeviceProver::prove_fd(int fd, uint64_t file_offset, uint64_t n, const uint8_t* r32, const uint8_t* s32)
{
    if (group_)
    {
        group_->check_n(n);
        return group_->prove_upload(witness_reader_fd(fd, file_offset), r32, s32);
    }
    return impl_->prove_with([&] { impl_->upload_fd(fd, file_offset, n); }, r32, s32);
}
const ProveTimings& DeviceProver::timings() const { return group_ ? group_->timings() : impl_->timings(); }
const ProveTimings& DeviceProver::shard_timings(int shard) const
{
    if (group_ && shard >= 0 && shard < group_->world())
        return group_->sh[shard]->timings();
    return first().timings();
}
void DeviceProver::msm_profile(int which, float* ms, uint64_t* entries) const
{
    DeviceProverImpl& p = first();
    p.set_device();
    switch (which)
    {
    case 0: // A, B1 and C run as one batched launch over the shared witness sort
    case 1:
    case 3: msm_last_accumulate(p.sort_w, p.sc_a, ms, entries); break;
    case 2: msm_last_accumulate(p.sort_w, p.sc_b2, ms, entries); break;
    case 4: msm_last_accumulate(p.sort_h, p.sc_h, ms, entries); break;
    default: throw FormatError("msm index out of range");
    }
}
const MsmArtefacts& DeviceProver::msm_artefacts() const { return group_ ? group_->artefacts() : impl_->artefacts(); }
void DeviceProver::copy_h(uint8_t* out) const
{
    if (group_)
        return group_->copy_h(out);
    impl_->set_device();
    KZP_CUDA_CHECK(cudaMemcpy(out, impl_->d_h, (size_t)impl_->domain * 32, cudaMemcpyDeviceToHost));
}
void DeviceProver::set_keep_ab(bool on)
{
    if (group_)
        throw FormatError("keep_ab is not available on a prover group (a and b live on different GPUs)");
    impl_->set_device();
    if (on && !impl_->d_keep_a)
    {
        KZP_CUDA_CHECK(cudaMalloc(&impl_->d_keep_a, (size_t)impl_->domain * 32));
        KZP_CUDA_CHECK(cudaMalloc(&impl_->d_keep_b, (size_t)impl_->domain * 32));
    }
    impl_->keep_ab = on;
}
void DeviceProver::copy_ab(uint8_t* out) const
{
    if (group_ || !impl_->d_keep_a)
        throw FormatError("set_keep_ab(true) was not called before the proof");
    impl_->set_device();
    size_t vec = (size_t)impl_->domain * 32;
    KZP_CUDA_CHECK(cudaMemcpy(out, impl_->d_keep_a, vec, cudaMemcpyDeviceToHost));
    KZP_CUDA_CHECK(cudaMemcpy(out + vec, impl_->d_keep_b, vec, cudaMemcpyDeviceToHost));
}

size_t pack_slice_capacity() { return kPackStride; }

size_t pack_witness_slice(const uint8_t* values, uint32_t count, uint8_t* out, uint32_t* n_full)
{
    if (count > kPackWires)
        throw FormatError("a slice holds at most 32768 values");
    constexpr uint32_t   kChunk = 4096;
    std::vector<uint8_t> bounce((size_t)kChunk * 32 + 64);
    uint8_t*             b  = bounce.data() + ((64 - ((uintptr_t)bounce.data() & 63)) & 63);
    uint32_t             nf = 0;
    for (uint32_t done = 0; done < count; done += kChunk)
    {
        uint32_t       cnt    = std::min(kChunk, count - done);
        uint32_t       padded = (cnt + kPackGroup - 1) / kPackGroup * kPackGroup;
        const uint8_t* src    = values + (size_t)done * 32;
        if (padded != cnt)
        {
            memcpy(b, src, (size_t)cnt * 32);
            memset(b + (size_t)cnt * 32, 0, (size_t)(padded - cnt) * 32);
            src = b;
        }
        pack_values(out, done, src, padded, &nf);
    }
    if (n_full)
        *n_full = nf;
    return kPackHead + (size_t)nf * 32;
}

} // namespace kzp
