// Host-side driver of one GPU-resident Groth16 prover (thin C++: parse, upload once, launch, assemble).
// Stage order follows Prover::prove in the reference (rust-rapidsnark/rapidsnark/src/groth16.cpp:43-360);
// what the reference runs as std::async futures on CPU threads runs here on two CUDA streams.
#include <chrono>
#include <cstring>
#include <random>
#include <thread>

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

void sample_blinding(uint8_t* out32)
{
    // groth16.cpp:296-316: 32 random bytes, clear the top two bits, reject values >= r
    std::random_device rd;
    for (;;)
    {
        for (int i = 0; i < 8; i++)
        {
            uint32_t v = rd();
            memcpy(out32 + 4 * i, &v, 4);
        }
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

// Sums the shards' partial MSM results, blinds with (r, s) and prints the proof. Host-only arithmetic (a few
// thousand field multiplications); mirrors groth16.cpp:296-357 + Proof::toJson (:379-410).
std::string assemble_proof(const HostVk& vk, const ShardPartials* ps, int count, const uint8_t* r32,
                           const uint8_t* s32, MsmArtefacts* art_out)
{
    HG1Affine alpha1, beta1, delta1;
    HG2Affine beta2, delta2;
    memcpy(&alpha1, vk.alpha1, 64);
    memcpy(&beta1, vk.beta1, 64);
    memcpy(&delta1, vk.delta1, 64);
    memcpy(&beta2, vk.beta2, 128);
    memcpy(&delta2, vk.delta2, 128);
    MsmArtefacts art_local;
    MsmArtefacts& art = art_out ? *art_out : art_local;
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
    // parity artefacts (affine, canonical)
    {
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

    uint8_t r[32], s[32];
    if (r32 && s32)
    {
        memcpy(r, r32, 32);
        memcpy(s, s32, 32);
    }
    else
    {
        sample_blinding(r);
        sample_blinding(s);
    }
    // rs = r*s mod r_modulus, canonical (groth16.cpp:346-347)
    uint8_t rs[32];
    {
        HFr fr, fs, t;
        memcpy(&fr, r, 32);
        memcpy(&fs, s, 32);
        while (HFr::geq_p(fr))
            HFr::sub_p(fr);
        while (HFr::geq_p(fs))
            HFr::sub_p(fs);
        HFr::to_mont(fr, fr);
        HFr::mul(t, fr, fs); // (r R)(s) R^-1 = r s
        memcpy(rs, &t, 32);
    }
    HG1 d1, al, be1, p1;
    HG2 d2, be2, p2;
    HG1::from_affine(d1, delta1);
    HG1::from_affine(al, alpha1);
    HG1::from_affine(be1, beta1);
    HG2::from_affine(d2, delta2);
    HG2::from_affine(be2, beta2);

    // pi_a = A + alpha1 + r*delta1            (groth16.cpp:328-330)
    HG1 pi_a = A;
    HG1::add(pi_a, al);
    scalar_mul(p1, d1, r);
    HG1::add(pi_a, p1);
    // pi_b = B2 + beta2 + s*delta2            (:332-334)
    HG2 pi_b = B2;
    HG2::add(pi_b, be2);
    scalar_mul(p2, d2, s);
    HG2::add(pi_b, p2);
    // pib1 = B1 + beta1 + s*delta1            (:336-338)
    HG1 pib1 = B1;
    HG1::add(pib1, be1);
    scalar_mul(p1, d1, s);
    HG1::add(pib1, p1);
    // pi_c = C + H + s*pi_a + r*pib1 - rs*delta1   (:340-352)
    HG1 pi_c = C;
    HG1::add(pi_c, H);
    scalar_mul(p1, pi_a, s);
    HG1::add(pi_c, p1);
    scalar_mul(p1, pib1, r);
    HG1::add(pi_c, p1);
    scalar_mul(p1, d1, rs);
    HG1 np1;
    HG1::neg(np1, p1);
    HG1::add(pi_c, np1);

    HG1Affine a_aff, c_aff;
    HG2Affine b_aff;
    HG1::to_affine(a_aff, pi_a);
    HG2::to_affine(b_aff, pi_b);
    HG1::to_affine(c_aff, pi_c);

    // compact JSON, keys in sorted order, exactly what nlohmann::json::dump() prints for
    // Proof::toJson (groth16.cpp:379-410, fullprover.cpp:246)
    std::string j;
    j.reserve(900);
    j += "{\"pi_a\":[";
    append_decimal(j, a_aff.x);
    j += ',';
    append_decimal(j, a_aff.y);
    j += ",\"1\"],\"pi_b\":[[";
    append_decimal(j, b_aff.x.a);
    j += ',';
    append_decimal(j, b_aff.x.b);
    j += "],[";
    append_decimal(j, b_aff.y.a);
    j += ',';
    append_decimal(j, b_aff.y.b);
    j += "],[\"1\",\"0\"]],\"pi_c\":[";
    append_decimal(j, c_aff.x);
    j += ',';
    append_decimal(j, c_aff.y);
    j += ",\"1\"],\"protocol\":\"groth16\"}";
    return j;
}

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

    cudaStream_t st_h = nullptr, st_w = nullptr, st_copy = nullptr;
    cudaEvent_t  ev[16] = {};

    CoefCsr   csr;
    NttDomain ntt;
    Fr *      d_w = nullptr, *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_h = nullptr;
    Fr *      d_keep_a = nullptr, *d_keep_b = nullptr;
    bool      keep_ab  = false;
    uint8_t*  pinned_w = nullptr;
    uint8_t*  pinned_out = nullptr; // 5 result points

    MsmBases<G1Xyzz>   bases_a, bases_b1, bases_c, bases_h;
    MsmBases<G2Xyzz>   bases_b2;
    MsmScratch<G1Xyzz> sc_a, sc_b1, sc_c, sc_h;
    MsmScratch<G2Xyzz> sc_b2;

    HostVk vk;

    ShardPartials parts;
    MsmArtefacts  art;
    ProveTimings  tm;
    bool          witness_resident = false;

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

    template <class XY>
    void make_bases(MsmBases<XY>& b, MsmScratch<XY>& sc, const uint8_t* sec, uint64_t count,
                    uint32_t scalar_base)
    {
        uint64_t first = (uint64_t)rank * count / (uint64_t)world;
        uint64_t last  = (uint64_t)(rank + 1) * count / (uint64_t)world;
        msm_bases_create<XY>(b, sec, first, last - first, scalar_base + (uint32_t)first, st_h);
        msm_scratch_create<XY>(sc, b.n);
    }

    DeviceProverImpl(const std::string& path, int dev, int rank_, int world_)
        : device(dev)
        , rank(rank_)
        , world(world_)
    {
        if (world < 1 || rank < 0 || rank >= world)
            throw FormatError("invalid shard rank/world");
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
        KZP_CUDA_CHECK(cudaStreamCreateWithFlags(&st_h, cudaStreamNonBlocking));
        KZP_CUDA_CHECK(cudaStreamCreateWithFlags(&st_w, cudaStreamNonBlocking));
        KZP_CUDA_CHECK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        for (auto& e : ev)
            KZP_CUDA_CHECK(cudaEventCreate(&e));

        memcpy(vk.alpha1, zh.alpha1, 64);
        memcpy(vk.beta1, zh.beta1, 64);
        memcpy(vk.delta1, zh.delta1, 64);
        memcpy(vk.beta2, zh.beta2, 128);
        memcpy(vk.delta2, zh.delta2, 128);

        build_csr(zh);
        ntt_domain_create(ntt, log_domain);
        size_t vec = (size_t)domain * 32;
        KZP_CUDA_CHECK(cudaMalloc(&d_w, (size_t)n_vars * 32));
        KZP_CUDA_CHECK(cudaMalloc(&d_a, vec));
        KZP_CUDA_CHECK(cudaMalloc(&d_b, vec));
        KZP_CUDA_CHECK(cudaMalloc(&d_c, vec));
        KZP_CUDA_CHECK(cudaMalloc(&d_h, vec));
        KZP_CUDA_CHECK(cudaMallocHost(&pinned_w, (size_t)n_vars * 32));
        KZP_CUDA_CHECK(cudaMallocHost(&pinned_out, sizeof(ShardPartials)));

        make_bases(bases_a, sc_a, zh.points_a, n_vars, 0);
        make_bases(bases_b1, sc_b1, zh.points_b1, n_vars, 0);
        make_bases(bases_b2, sc_b2, zh.points_b2, n_vars, 0);
        make_bases(bases_c, sc_c, zh.points_c, n_vars - n_public - 1, n_public + 1);
        make_bases(bases_h, sc_h, zh.points_h, domain, 0);
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
    }

    ~DeviceProverImpl()
    {
        cudaSetDevice(device);
        cudaDeviceSynchronize();
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
        ntt_domain_destroy(ntt);
        cudaFree(csr.row_ptr);
        cudaFree(csr.wire);
        cudaFree(csr.coef);
        cudaFree(d_w);
        cudaFree(d_a);
        cudaFree(d_b);
        cudaFree(d_c);
        cudaFree(d_h);
        cudaFree(d_keep_a);
        cudaFree(d_keep_b);
        cudaFreeHost(pinned_w);
        cudaFreeHost(pinned_out);
        for (auto& e : ev)
            cudaEventDestroy(e);
        cudaStreamDestroy(st_h);
        cudaStreamDestroy(st_w);
        cudaStreamDestroy(st_copy);
    }

    void upload(const uint8_t* values, uint64_t n)
    {
        if (n < n_vars)
            throw FormatError("witness has fewer values than the zkey has variables");
        set_device();
        // stage through pinned memory in slices so the host copy of slice k+1 overlaps the DMA of slice k
        const size_t total = (size_t)n_vars * 32;
        const size_t slice = 4u << 20;
        KZP_CUDA_CHECK(cudaEventRecord(ev[0], st_copy));
        for (size_t off = 0; off < total; off += slice)
        {
            size_t len = std::min(slice, total - off);
            memcpy(pinned_w + off, values + off, len);
            KZP_CUDA_CHECK(cudaMemcpyAsync((uint8_t*)d_w + off, pinned_w + off, len,
                                           cudaMemcpyHostToDevice, st_copy));
        }
        KZP_CUDA_CHECK(cudaEventRecord(ev[1], st_copy));
        witness_resident = true;
    }

    void run_gpu()
    {
        if (!witness_resident)
            throw FormatError("no witness uploaded");
        set_device();
        const uint32_t* w = reinterpret_cast<const uint32_t*>(d_w);
        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_h, ev[1], 0));
        KZP_CUDA_CHECK(cudaStreamWaitEvent(st_w, ev[1], 0));

        // ---- stream H: SpMV -> 3 x (iNTT, coset, NTT) -> pointwise -> MSM H
        KZP_CUDA_CHECK(cudaEventRecord(ev[2], st_h));
        spmv_abc(csr, d_w, d_a, d_b, d_c, st_h);
        if (keep_ab)
        {
            KZP_CUDA_CHECK(cudaMemcpyAsync(d_keep_a, d_a, (size_t)domain * 32, cudaMemcpyDeviceToDevice, st_h));
            KZP_CUDA_CHECK(cudaMemcpyAsync(d_keep_b, d_b, (size_t)domain * 32, cudaMemcpyDeviceToDevice, st_h));
        }
        KZP_CUDA_CHECK(cudaEventRecord(ev[3], st_h));
        Fr* vecs[3] = {d_a, d_b, d_c};
        for (Fr* x : vecs)
        {
            ntt_inverse_dif(ntt, x, ntt.coset_br, st_h);
            ntt_forward_dit(ntt, x, st_h);
        }
        h_pointwise(d_a, d_b, d_c, d_h, domain, st_h);
        KZP_CUDA_CHECK(cudaEventRecord(ev[4], st_h));
        msm_run(bases_h, sc_h, reinterpret_cast<const uint32_t*>(d_h), st_h);
        KZP_CUDA_CHECK(cudaEventRecord(ev[5], st_h));

        // ---- stream W: the four witness MSMs
        KZP_CUDA_CHECK(cudaEventRecord(ev[6], st_w));
        msm_run(bases_a, sc_a, w, st_w);
        KZP_CUDA_CHECK(cudaEventRecord(ev[7], st_w));
        msm_run(bases_b1, sc_b1, w, st_w);
        KZP_CUDA_CHECK(cudaEventRecord(ev[8], st_w));
        msm_run(bases_b2, sc_b2, w, st_w);
        KZP_CUDA_CHECK(cudaEventRecord(ev[9], st_w));
        msm_run(bases_c, sc_c, w, st_w);
        KZP_CUDA_CHECK(cudaEventRecord(ev[10], st_w));

        // results -> pinned host
        KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 0, sc_a.result, 128, cudaMemcpyDeviceToHost, st_w));
        KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 128, sc_b1.result, 128, cudaMemcpyDeviceToHost, st_w));
        KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 256, sc_c.result, 128, cudaMemcpyDeviceToHost, st_w));
        KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 512, sc_b2.result, 256, cudaMemcpyDeviceToHost, st_w));
        KZP_CUDA_CHECK(cudaMemcpyAsync(pinned_out + 384, sc_h.result, 128, cudaMemcpyDeviceToHost, st_h));
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_w));
        KZP_CUDA_CHECK(cudaStreamSynchronize(st_h));
        memcpy(parts.bytes, pinned_out, sizeof(parts.bytes));

        auto el = [&](int a, int b) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[a], ev[b]);
            return ms;
        };
        tm.h2d_ms    = el(0, 1);
        tm.spmv_ms   = el(2, 3);
        tm.ntt_ms    = el(3, 4);
        tm.msm_h_ms  = el(4, 5);
        tm.msm_a_ms  = el(6, 7);
        tm.msm_b1_ms = el(7, 8);
        tm.msm_b2_ms = el(8, 9);
        tm.msm_c_ms  = el(9, 10);
        tm.gpu_ms    = std::max(el(2, 5), el(2, 10));
        uint32_t per_msm   = 8;
        tm.kernel_launches = 1 + 3 * 2 * log_domain + 1 + 5 * per_msm;
    }

    std::string assemble(const ShardPartials* ps, int count, const uint8_t* r32, const uint8_t* s32)
    {
        double      t0 = now_ms();
        std::string j  = assemble_proof(vk, ps, count, r32, s32, &art);
        tm.assemble_host_ms = (float)(now_ms() - t0);
        return j;
    }
};

// ------------------------------------------------------------------ facade
DeviceProver::DeviceProver(const std::string& zkey_path, int device, int shard_rank, int shard_world)
    : impl_(new DeviceProverImpl(zkey_path, device, shard_rank, shard_world))
{
}
DeviceProver::~DeviceProver() {}
uint32_t DeviceProver::n_vars() const { return impl_->n_vars; }
uint32_t DeviceProver::n_public() const { return impl_->n_public; }
uint32_t DeviceProver::domain_size() const { return impl_->domain; }
uint64_t DeviceProver::n_coefs() const { return impl_->n_coefs; }
int      DeviceProver::device() const { return impl_->device; }
void     DeviceProver::upload_witness(const uint8_t* values, uint64_t n) { impl_->upload(values, n); }
void     DeviceProver::run_gpu() { impl_->run_gpu(); }
const ShardPartials& DeviceProver::partials() const { return impl_->parts; }
std::string DeviceProver::assemble(const ShardPartials* parts, int count, const uint8_t* r32,
                                   const uint8_t* s32)
{
    return impl_->assemble(parts, count, r32, s32);
}
std::string DeviceProver::prove(const uint8_t* values, uint64_t n, const uint8_t* r32, const uint8_t* s32)
{
    double t0 = now_ms();
    impl_->upload(values, n);
    impl_->run_gpu();
    std::string j           = impl_->assemble(&impl_->parts, 1, r32, s32);
    impl_->tm.total_host_ms = (float)(now_ms() - t0);
    return j;
}
const ProveTimings& DeviceProver::timings() const { return impl_->tm; }
void DeviceProver::msm_profile(int which, float* ms, uint64_t* entries) const
{
    impl_->set_device();
    switch (which)
    {
    case 0: msm_last_accumulate(impl_->sc_a, ms, entries); break;
    case 1: msm_last_accumulate(impl_->sc_b1, ms, entries); break;
    case 2: msm_last_accumulate(impl_->sc_b2, ms, entries); break;
    case 3: msm_last_accumulate(impl_->sc_c, ms, entries); break;
    case 4: msm_last_accumulate(impl_->sc_h, ms, entries); break;
    default: throw FormatError("msm index out of range");
    }
}
const MsmArtefacts& DeviceProver::msm_artefacts() const { return impl_->art; }
void DeviceProver::copy_h(uint8_t* out) const
{
    impl_->set_device();
    KZP_CUDA_CHECK(cudaMemcpy(out, impl_->d_h, (size_t)impl_->domain * 32, cudaMemcpyDeviceToHost));
}
void DeviceProver::set_keep_ab(bool on)
{
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
    if (!impl_->d_keep_a)
        throw FormatError("set_keep_ab(true) was not called before the proof");
    impl_->set_device();
    size_t vec = (size_t)impl_->domain * 32;
    KZP_CUDA_CHECK(cudaMemcpy(out, impl_->d_keep_a, vec, cudaMemcpyDeviceToHost));
    KZP_CUDA_CHECK(cudaMemcpy(out + vec, impl_->d_keep_b, vec, cudaMemcpyDeviceToHost));
}

} // namespace kzp
