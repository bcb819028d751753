// extern "C" surface declared in include/kzp_b200.h. Every entry point catches all exceptions: nothing
// may unwind across the FFI boundary (the reference lets non-invalid_argument/system_error exceptions
// escape, fullprover.cpp:80-101; SURVEY.md §8(b) asks the replacement to catch everything).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kzp_b200.h"
#include "binfile.hpp"
#include "device.hpp"
#include "hostff.hpp"
#include "prover.hpp"

using namespace kzp;

static thread_local std::string g_last_error;

const char* kzp_last_error(void) { return g_last_error.c_str(); }
const char* kzp_version(void) { return "kzp_b200 0.1 (sm_100a)"; }
void        kzp_free(void* p) { free(p); }

int kzp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

template <class Fn>
static int guarded(Fn&& fn)
{
    try
    {
        fn();
        return KZP_OK;
    }
    catch (const CudaError& e)
    {
        g_last_error = e.what();
        return KZP_ERR_CUDA;
    }
    catch (const FormatError& e)
    {
        g_last_error = e.what();
        return KZP_ERR_FORMAT;
    }
    catch (const LoadError& e)
    {
        g_last_error = e.what();
        return KZP_ERR_IO;
    }
    catch (const std::exception& e)
    {
        g_last_error = e.what();
        return KZP_ERR_FORMAT;
    }
    catch (...)
    {
        g_last_error = "unknown exception";
        return KZP_ERR_FORMAT;
    }
}

static int pick_device(int device)
{
    if (device >= 0)
        return device;
    const char* env = getenv("KZP_DEVICE");
    return env ? atoi(env) : 0;
}

static void use_device(int device)
{
    int n = kzp_device_count();
    if (n == 0)
        throw CudaError("no CUDA device available (this library has no CPU fallback)");
    int d = pick_device(device);
    if (d >= n)
        throw CudaError("CUDA device index out of range");
    KZP_CUDA_CHECK(cudaSetDevice(d));
}

// ------------------------------------------------------------------------------------------- prover
struct kzp_prover
{
    DeviceProver* prover      = nullptr;
    int           state       = KZP_STATE_OK;
    int           last_status = KZP_OK; // KZP_ERR_* class of the last failed call on this handle
    std::string   why;
};

// After a CUDA failure: is the context still usable? A sticky error (illegal address, launch failure, ECC, lost
// device) makes every later runtime call fail with the same code; out-of-memory and invalid-argument errors do not.
static bool device_is_dead()
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
        e = cudaGetLastError();
    else
        cudaGetLastError();
    return e != cudaSuccess && e != cudaErrorMemoryAllocation;
}

template <class Make>
static kzp_prover* new_prover_handle(Make&& make, int* state_out)
{
    kzp_prover* h = new (std::nothrow) kzp_prover();
    if (!h)
        return nullptr;
    try
    {
        if (kzp_device_count() == 0)
            throw CudaError("no CUDA device available (this library has no CPU fallback)");
        h->prover = make();
        h->state  = KZP_STATE_OK;
    }
    catch (const LoadError& e)
    {
        h->state = KZP_STATE_ZKEY_FILE_LOAD_ERROR; // std::system_error in the reference
        h->why   = e.what();
    }
    catch (const FormatError& e)
    {
        h->state = KZP_STATE_UNSUPPORTED_ZKEY_CURVE; // std::invalid_argument in the reference
        h->why   = e.what();
    }
    catch (const std::exception& e)
    {
        // CUDA failures have no counterpart in the reference's enum; report them as a load error so the
        // Rust side maps them to ProverInitError::ZKeyFileLoadError (rust-rapidsnark/src/lib.rs:53-60)
        h->state = KZP_STATE_ZKEY_FILE_LOAD_ERROR;
        h->why   = e.what();
    }
    catch (...)
    {
        h->state = KZP_STATE_ZKEY_FILE_LOAD_ERROR;
        h->why   = "unknown exception";
    }
    g_last_error = h->why;
    if (state_out)
        *state_out = h->state;
    return h;
}

kzp_prover* kzp_prover_new_sharded(const char* zkey_path, int device, int rank, int world, int* state_out)
{
    return new_prover_handle([&] { return new DeviceProver(zkey_path ? zkey_path : "", pick_device(device), rank, world); },
                             state_out);
}

// "0,1,2,3" -> {0, 1, 2, 3}; anything that is not a list of small non-negative integers -> empty
static std::vector<int> parse_device_list(const char* text)
{
    std::vector<int> out;
    if (!text)
        return out;
    const char* p = text;
    while (*p)
    {
        while (*p == ' ' || *p == ',')
            p++;
        if (!*p)
            break;
        if (*p < '0' || *p > '9')
            return {};
        int v = 0;
        while (*p >= '0' && *p <= '9' && v < 100000)
            v = v * 10 + (*p++ - '0');
        out.push_back(v);
    }
    return out;
}

kzp_prover* kzp_prover_new_group(const char* zkey_path, const int* devices, int n_devices, int* state_out)
{
    std::vector<int> devs;
    if (devices && n_devices > 0)
        devs.assign(devices, devices + n_devices);
    else
        devs = parse_device_list(getenv("KZP_SHARD_DEVICES"));
    return new_prover_handle(
        [&]() -> DeviceProver* {
            if (devs.empty())
                throw FormatError("no devices listed for the prover group (argument or $KZP_SHARD_DEVICES)");
            return new DeviceProver(zkey_path ? zkey_path : "", devs);
        },
        state_out);
}

kzp_prover* kzp_prover_new(const char* zkey_path, int device, int* state_out)
{
    // $KZP_SHARD_DEVICES="0,1,..." turns every prover made through the reference-shaped constructor (device < 0:
    // "you choose") into one proof sharded over those GPUs — FullProver::FullProver(zkeyPath) has no other knob
    if (device < 0)
    {
        const char* sd = getenv("KZP_SHARD_DEVICES");
        if (sd && *sd)
            return kzp_prover_new_group(zkey_path, nullptr, 0, state_out);
    }
    return kzp_prover_new_sharded(zkey_path, device, 0, 1, state_out);
}

int kzp_prover_group_info(kzp_prover* p, int* shards, int* fused_exchange, int* distributed_ntt)
{
    if (!p || !p->prover)
        return KZP_ERR_STATE;
    if (shards)
        *shards = p->prover->group_size();
    if (fused_exchange)
        *fused_exchange = p->prover->group_fused_exchange() ? 1 : 0;
    if (distributed_ntt)
        *distributed_ntt = p->prover->group_distributed_ntt() ? 1 : 0;
    return KZP_OK;
}

void kzp_prover_free(kzp_prover* p)
{
    if (!p)
        return;
    try
    {
        delete p->prover;
    }
    catch (...)
    {
    }
    delete p;
}

static char* dup_string(const std::string& s)
{
    char* out = (char*)malloc(s.size() + 1);
    if (out)
        memcpy(out, s.c_str(), s.size() + 1);
    return out;
}

// shared tail of the two prove entry points: `run` produces the proof JSON (throws on failure)
template <class Run>
static int prove_common(kzp_prover* p, Run&& run, char** json_out, int* error_out, int* prover_time_ms)
{
    auto fail = [&](int err) {
        if (error_out)
            *error_out = err;
        return KZP_RESPONSE_ERROR;
    };
    std::string json;
    auto        t0 = std::chrono::steady_clock::now();
    int         rc = guarded([&] { json = run(); });
    p->last_status = rc;
    if (rc == KZP_ERR_CUDA)
    {
        // Not the client's fault: answer PROVER_NOT_READY, and when the device is gone for good take the handle out
        // of service so that health checks and the pool see it (every later call answers NOT_READY at once).
        std::string what = g_last_error;
        if (device_is_dead())
        {
            p->state = KZP_STATE_DEVICE_FAULT;
            p->why   = "CUDA device fault: " + what;
        }
        g_last_error = what;
        return fail(KZP_PROVER_ERROR_NOT_READY);
    }
    if (rc != KZP_OK)
        return fail(KZP_PROVER_ERROR_INVALID_INPUT); // KZP_ERR_FORMAT / KZP_ERR_IO: the witness is at fault
    auto t1 = std::chrono::steady_clock::now();
    if (prover_time_ms)
        *prover_time_ms = (int)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
    *json_out = dup_string(json);
    if (error_out)
        *error_out = KZP_PROVER_ERROR_NONE;
    return KZP_RESPONSE_SUCCESS;
}

int kzp_prover_prove_mem(kzp_prover* p, const uint8_t* witness, uint64_t n, const uint8_t* r32,
                         const uint8_t* s32, char** json_out, int* error_out, int* prover_time_ms)
{
    if (json_out)
        *json_out = nullptr;
    if (prover_time_ms)
        *prover_time_ms = 0;
    auto fail = [&](int err) {
        if (error_out)
            *error_out = err;
        return KZP_RESPONSE_ERROR;
    };
    if (!p || p->state != KZP_STATE_OK || !p->prover)
    {
        g_last_error = p ? p->why : "null prover";
        return fail(KZP_PROVER_ERROR_NOT_READY);
    }
    if (!witness || !json_out)
    {
        g_last_error = "null argument";
        return fail(KZP_PROVER_ERROR_INVALID_INPUT);
    }
    return prove_common(p, [&] { return p->prover->prove(witness, n, r32, s32); }, json_out, error_out,
                        prover_time_ms);
}

int kzp_prover_prove_resident(kzp_prover* p, const uint8_t* r32, const uint8_t* s32, char** json_out, int* error_out,
                              int* prover_time_ms)
{
    if (json_out)
        *json_out = nullptr;
    if (prover_time_ms)
        *prover_time_ms = 0;
    auto fail = [&](int err) {
        if (error_out)
            *error_out = err;
        return KZP_RESPONSE_ERROR;
    };
    if (!p || p->state != KZP_STATE_OK || !p->prover)
    {
        g_last_error = p ? p->why : "null prover";
        return fail(KZP_PROVER_ERROR_NOT_READY);
    }
    if (!json_out)
    {
        g_last_error = "null argument";
        return fail(KZP_PROVER_ERROR_INVALID_INPUT);
    }
    return prove_common(p, [&] { return p->prover->prove_resident(r32, s32); }, json_out, error_out, prover_time_ms);
}

int kzp_prover_prove(kzp_prover* p, const char* wtns_path, const uint8_t* r32, const uint8_t* s32,
                     char** json_out, int* error_out, int* prover_time_ms)
{
    if (json_out)
        *json_out = nullptr;
    if (prover_time_ms)
        *prover_time_ms = 0;
    auto fail = [&](int err) {
        if (error_out)
            *error_out = err;
        return KZP_RESPONSE_ERROR;
    };
    if (!p || p->state != KZP_STATE_OK || !p->prover)
    {
        g_last_error = p ? p->why : "null prover";
        return fail(KZP_PROVER_ERROR_NOT_READY);
    }
    if (!json_out)
    {
        g_last_error = "null argument";
        return fail(KZP_PROVER_ERROR_INVALID_INPUT);
    }
    int ret = KZP_RESPONSE_ERROR;
    int rc  = guarded([&] {
        // The reference throws out of prove() on an unreadable witness (fullprover.cpp:212, a bug noted in
        // SURVEY.md §8(b)); here it is reported as INVALID_INPUT. Only the headers are read through the mapping;
        // the values go file -> pinned staging buffer by pread().
        MappedFile file(wtns_path ? wtns_path : "");
        BinView    bin(file.data(), file.size(), "wtns", 2);
        WtnsHeader wh = parse_wtns(bin);
        if (!wh.prime_is_bn254_r)
        {
            g_last_error = "witness file uses a different curve than bn128";
            ret          = fail(KZP_PROVER_ERROR_WITNESS_GENERATION_INVALID_CURVE);
            return;
        }
        uint64_t n   = wh.values_bytes / 32;
        uint64_t off = (uint64_t)(wh.values - file.data());
        // (classifying straight out of the mapping instead of pread() into a cache-resident bounce buffer was measured
        //  again this round: 11.3 ms end to end against 10.8 ms — the page faults of a fresh mapping cost more than the copy)
        ret = prove_common(p, [&] { return p->prover->prove_fd(file.fd(), off, n, r32, s32); }, json_out, error_out,
                           prover_time_ms);
    });
    if (rc != KZP_OK)
    {
        p->last_status = rc;
        return fail(KZP_PROVER_ERROR_INVALID_INPUT);
    }
    return ret;
}

int kzp_prover_state(const kzp_prover* p) { return p ? p->state : KZP_STATE_ZKEY_FILE_LOAD_ERROR; }
int kzp_prover_last_status(const kzp_prover* p) { return p ? p->last_status : KZP_ERR_STATE; }

#define KZP_REQUIRE_READY(p)                                                                     \
    if (!(p) || (p)->state != KZP_STATE_OK || !(p)->prover)                                        \
    {                                                                                            \
        g_last_error = (p) ? (p)->why : "null prover";                                            \
        return KZP_ERR_STATE;                                                                    \
    }

int kzp_prover_upload_witness(kzp_prover* p, const uint8_t* witness, uint64_t n)
{
    KZP_REQUIRE_READY(p);
    return guarded([&] { p->prover->upload_witness(witness, n); });
}

int kzp_prover_upload_witness_file(kzp_prover* p, const char* wtns_path)
{
    KZP_REQUIRE_READY(p);
    return guarded([&] {
        MappedFile file(wtns_path ? wtns_path : "");
        BinView    bin(file.data(), file.size(), "wtns", 2);
        WtnsHeader wh = parse_wtns(bin);
        if (!wh.prime_is_bn254_r)
            throw FormatError("witness file uses a different curve than bn128");
        p->prover->upload_witness_fd(file.fd(), (uint64_t)(wh.values - file.data()), wh.values_bytes / 32);
    });
}

int kzp_prover_run_gpu(kzp_prover* p)
{
    KZP_REQUIRE_READY(p);
    return guarded([&] { p->prover->run_gpu(); });
}

int kzp_prover_get_partials(kzp_prover* p, uint8_t* out768)
{
    KZP_REQUIRE_READY(p);
    memcpy(out768, p->prover->partials().bytes, KZP_PARTIALS_BYTES);
    return KZP_OK;
}

int kzp_prover_assemble(kzp_prover* p, const uint8_t* partials, int count, const uint8_t* r32,
                        const uint8_t* s32, char** json_out)
{
    KZP_REQUIRE_READY(p);
    if (!partials || count < 1 || !json_out)
    {
        g_last_error = "bad arguments";
        return KZP_ERR_FORMAT;
    }
    return guarded([&] {
        std::vector<ShardPartials> ps(count);
        for (int k = 0; k < count; k++)
            memcpy(ps[k].bytes, partials + (size_t)k * KZP_PARTIALS_BYTES, KZP_PARTIALS_BYTES);
        std::string j = p->prover->assemble(ps.data(), count, r32, s32);
        *json_out     = dup_string(j);
    });
}

int kzp_prover_info(kzp_prover* p, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size,
                    uint64_t* n_coefs, int* device)
{
    KZP_REQUIRE_READY(p);
    if (n_vars)
        *n_vars = p->prover->n_vars();
    if (n_public)
        *n_public = p->prover->n_public();
    if (domain_size)
        *domain_size = p->prover->domain_size();
    if (n_coefs)
        *n_coefs = p->prover->n_coefs();
    if (device)
        *device = p->prover->device();
    return KZP_OK;
}

static int timings_out(const ProveTimings& t, float* out, int cap);

int kzp_prover_group_shard_timings(kzp_prover* p, int shard, float* out, int cap)
{
    if (!p || !p->prover || !out || shard < 0 || shard >= p->prover->group_size())
        return 0;
    return timings_out(p->prover->shard_timings(shard), out, cap);
}

int kzp_prover_timings(kzp_prover* p, float* out, int cap)
{
    if (!p || !p->prover || !out)
        return 0;
    return timings_out(p->prover->timings(), out, cap);
}

static int timings_out(const ProveTimings& t, float* out, int cap)
{
    float v[12] = {t.h2d_ms,     t.spmv_ms,     t.ntt_ms, t.msm_h_ms,         t.msm_wsort_ms,  t.msm_wg1_ms,
                   t.msm_wg2_ms, t.h2d_mbytes, t.gpu_ms, t.assemble_host_ms, t.total_host_ms,
                   (float)t.kernel_launches};
    int   n     = cap < 12 ? cap : 12;
    for (int i = 0; i < n; i++)
        out[i] = v[i];
    return n;
}

int kzp_prover_msm_profile(kzp_prover* p, int which, float* accumulate_ms, uint64_t* entries)
{
    KZP_REQUIRE_READY(p);
    return guarded([&] { p->prover->msm_profile(which, accumulate_ms, entries); });
}

int kzp_prover_get_h(kzp_prover* p, uint8_t* out, uint64_t out_bytes)
{
    KZP_REQUIRE_READY(p);
    if (out_bytes < (uint64_t)p->prover->domain_size() * 32)
    {
        g_last_error = "buffer too small";
        return KZP_ERR_FORMAT;
    }
    return guarded([&] { p->prover->copy_h(out); });
}

int kzp_prover_keep_ab(kzp_prover* p, int on)
{
    KZP_REQUIRE_READY(p);
    return guarded([&] { p->prover->set_keep_ab(on != 0); });
}

int kzp_prover_get_ab(kzp_prover* p, uint8_t* out, uint64_t out_bytes)
{
    KZP_REQUIRE_READY(p);
    if (out_bytes < (uint64_t)p->prover->domain_size() * 64)
    {
        g_last_error = "buffer too small";
        return KZP_ERR_FORMAT;
    }
    return guarded([&] { p->prover->copy_ab(out); });
}

int kzp_prover_get_msm_results(kzp_prover* p, uint8_t* out384)
{
    KZP_REQUIRE_READY(p);
    memcpy(out384, p->prover->msm_artefacts().bytes, 384);
    return KZP_OK;
}

// ------------------------------------------------------------------------------------------- NTT
static uint32_t log2_exact(uint64_t n)
{
    if (n == 0 || (n & (n - 1)))
        throw FormatError("size is not a power of two");
    uint32_t l = 0;
    while ((1ull << l) < n)
        l++;
    return l;
}

struct DevBuf
{
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { KZP_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 16)); }
    ~DevBuf() { cudaFree(p); }
};

int kzp_fr_ntt(uint8_t* data, uint64_t n, int inverse, int device)
{
    return guarded([&] {
        use_device(device);
        uint32_t  log_n = log2_exact(n);
        NttDomain d;
        ntt_domain_create(d, log_n);
        try
        {
            DevBuf buf(n * 32);
            Fr*    x = (Fr*)buf.p;
            KZP_CUDA_CHECK(cudaMemcpy(x, data, n * 32, cudaMemcpyHostToDevice));
            if (inverse)
            {
                ntt_inverse_dif(d, x, nullptr, 0);
                ntt_bitrev_permute(x, log_n, 0);
                fr_scale(x, n, d.n_inv, 0);
            }
            else
            {
                ntt_bitrev_permute(x, log_n, 0);
                ntt_forward_dit(d, x, 0);
            }
            KZP_CUDA_CHECK(cudaDeviceSynchronize());
            KZP_CUDA_CHECK(cudaMemcpy(data, x, n * 32, cudaMemcpyDeviceToHost));
        }
        catch (...)
        {
            ntt_domain_destroy(d);
            throw;
        }
        ntt_domain_destroy(d);
    });
}

int kzp_fr_coset_chain(uint8_t* data, uint64_t n, int device)
{
    return guarded([&] {
        use_device(device);
        uint32_t  log_n = log2_exact(n);
        NttDomain d;
        ntt_domain_create(d, log_n);
        try
        {
            DevBuf buf(n * 32);
            Fr*    x = (Fr*)buf.p;
            KZP_CUDA_CHECK(cudaMemcpy(x, data, n * 32, cudaMemcpyHostToDevice));
            Fr* xs[1] = {x};
            ntt_coset_chain(d, xs, 1, 0);
            KZP_CUDA_CHECK(cudaDeviceSynchronize());
            KZP_CUDA_CHECK(cudaMemcpy(data, x, n * 32, cudaMemcpyDeviceToHost));
        }
        catch (...)
        {
            ntt_domain_destroy(d);
            throw;
        }
        ntt_domain_destroy(d);
    });
}

int kzp_fr_ntt_bench(uint32_t log_n, int iters, int device, float* ms_per_chain)
{
    return guarded([&] {
        use_device(device);
        NttDomain d;
        ntt_domain_create(d, log_n);
        try
        {
            uint64_t n = 1ull << log_n;
            DevBuf   buf(n * 32);
            Fr*      x = (Fr*)buf.p;
            // any canonical data will do: reuse the coset table as input
            KZP_CUDA_CHECK(cudaMemcpy(x, d.coset_br, n * 32, cudaMemcpyDeviceToDevice));
            cudaEvent_t e0, e1;
            KZP_CUDA_CHECK(cudaEventCreate(&e0));
            KZP_CUDA_CHECK(cudaEventCreate(&e1));
            Fr* xs[1] = {x};
            for (int w = 0; w < 2; w++)
                ntt_coset_chain(d, xs, 1, 0);
            KZP_CUDA_CHECK(cudaDeviceSynchronize());
            KZP_CUDA_CHECK(cudaEventRecord(e0, 0));
            for (int it = 0; it < iters; it++)
                ntt_coset_chain(d, xs, 1, 0);
            KZP_CUDA_CHECK(cudaEventRecord(e1, 0));
            KZP_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            KZP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            *ms_per_chain = ms / (float)(iters > 0 ? iters : 1);
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        catch (...)
        {
            ntt_domain_destroy(d);
            throw;
        }
        ntt_domain_destroy(d);
    });
}

// ------------------------------------------------------------------------------------------- MSM
struct kzp_msm
{
    int                group  = 0;
    int                device = 0;
    uint64_t           n      = 0;
    MsmSort            sort;
    MsmBases<G1Xyzz>   b1;
    MsmScratch<G1Xyzz> s1;
    MsmBases<G2Xyzz>   b2;
    MsmScratch<G2Xyzz> s2;
};

static void msm_launch(kzp_msm* m, const uint32_t* scalars)
{
    msm_sort_run(m->sort, scalars, 0);
    if (m->group == 0)
    {
        const MsmBases<G1Xyzz>* b[1] = {&m->b1};
        MsmScratch<G1Xyzz>*     s[1] = {&m->s1};
        msm_reduce_batch<G1Xyzz>(m->sort, b, s, 1, 0);
    }
    else
    {
        const MsmBases<G2Xyzz>* b[1] = {&m->b2};
        MsmScratch<G2Xyzz>*     s[1] = {&m->s2};
        msm_reduce_batch<G2Xyzz>(m->sort, b, s, 1, 0);
    }
}

kzp_msm* kzp_msm_new(int group, const uint8_t* bases, uint64_t n, int device)
{
    const char* env = getenv("KZP_MSM_WINDOW");
    return kzp_msm_new_ex(group, bases, n, device, env ? atoi(env) : 0, 0);
}

kzp_msm* kzp_msm_new_ex(int group, const uint8_t* bases, uint64_t n, int device, int window_bits, int two_level)
{
    kzp_msm* m  = nullptr;
    int      rc = guarded([&] {
        if (group != 0 && group != 1)
            throw FormatError("group must be 0 (G1) or 1 (G2)");
        if (window_bits == 0)
            window_bits = 16;
        if (window_bits < (int)kMsmMinWindowBits || window_bits > (int)kMsmMaxWindowBits)
            throw FormatError("window_bits must be 0 (default) or 16..22");
        const uint32_t c = (uint32_t)window_bits;
        use_device(device);
        m         = new kzp_msm();
        m->group  = group;
        m->device = pick_device(device);
        m->n      = n;
        if (group == 0)
        {
            msm_bases_create<G1Xyzz>(m->b1, bases, n, true, 0, c);
            msm_sort_create(m->sort, m->b1.n, m->b1.scalar_idx, 0, c, two_level != 0);
            const char* ce = getenv("KZP_MSM_CHUNK"); // entries per accumulate thread (0 / unset: msm_default_chunk)
            msm_scratch_create<G1Xyzz>(m->s1, m->sort, ce ? (uint32_t)atoi(ce) : 0);
        }
        else
        {
            msm_bases_create<G2Xyzz>(m->b2, bases, n, true, 0, c);
            msm_sort_create(m->sort, m->b2.n, m->b2.scalar_idx, 0, c, two_level != 0);
            msm_scratch_create<G2Xyzz>(m->s2, m->sort, 0);
        }
    });
    if (rc != KZP_OK)
    {
        kzp_msm_free(m);
        return nullptr;
    }
    return m;
}

void kzp_msm_free(kzp_msm* m)
{
    if (!m)
        return;
    cudaSetDevice(m->device);
    msm_sort_destroy(m->sort);
    msm_bases_destroy(m->b1);
    msm_scratch_destroy(m->s1);
    msm_bases_destroy(m->b2);
    msm_scratch_destroy(m->s2);
    delete m;
}

typedef Fp2T<HFq>    HFq2c;
typedef XyzzT<HFq>   HG1c;
typedef XyzzT<HFq2c> HG2c;

static void msm_result_to_canonical(kzp_msm* m, uint8_t* out)
{
    if (m->group == 0)
    {
        HG1c p;
        KZP_CUDA_CHECK(cudaMemcpy(&p, m->s1.result, 128, cudaMemcpyDeviceToHost));
        AffineT<HFq> a;
        HG1c::to_affine(a, p);
        HFq t;
        HFq::from_mont(t, a.x);
        memcpy(out, &t, 32);
        HFq::from_mont(t, a.y);
        memcpy(out + 32, &t, 32);
    }
    else
    {
        HG2c p;
        KZP_CUDA_CHECK(cudaMemcpy(&p, m->s2.result, 256, cudaMemcpyDeviceToHost));
        AffineT<HFq2c> a;
        HG2c::to_affine(a, p);
        HFq t;
        HFq::from_mont(t, a.x.a);
        memcpy(out, &t, 32);
        HFq::from_mont(t, a.x.b);
        memcpy(out + 32, &t, 32);
        HFq::from_mont(t, a.y.a);
        memcpy(out + 64, &t, 32);
        HFq::from_mont(t, a.y.b);
        memcpy(out + 96, &t, 32);
    }
}

int kzp_msm_run(kzp_msm* m, const uint8_t* scalars, uint8_t* out)
{
    if (!m)
    {
        g_last_error = "null msm handle";
        return KZP_ERR_STATE;
    }
    return guarded([&] {
        KZP_CUDA_CHECK(cudaSetDevice(m->device));
        DevBuf sc(m->n * 32);
        KZP_CUDA_CHECK(cudaMemcpy(sc.p, scalars, m->n * 32, cudaMemcpyHostToDevice));
        msm_launch(m, (const uint32_t*)sc.p);
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
        msm_result_to_canonical(m, out);
    });
}

int kzp_msm_bench(kzp_msm* m, const uint8_t* scalars, int iters, float* ms_per_msm, uint64_t* entries)
{
    if (!m)
    {
        g_last_error = "null msm handle";
        return KZP_ERR_STATE;
    }
    return guarded([&] {
        KZP_CUDA_CHECK(cudaSetDevice(m->device));
        DevBuf sc(m->n * 32);
        KZP_CUDA_CHECK(cudaMemcpy(sc.p, scalars, m->n * 32, cudaMemcpyHostToDevice));
        auto run = [&] { msm_launch(m, (const uint32_t*)sc.p); };
        for (int w = 0; w < 2; w++)
            run();
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        KZP_CUDA_CHECK(cudaEventCreate(&e0));
        KZP_CUDA_CHECK(cudaEventCreate(&e1));
        KZP_CUDA_CHECK(cudaEventRecord(e0, 0));
        for (int it = 0; it < iters; it++)
            run();
        KZP_CUDA_CHECK(cudaEventRecord(e1, 0));
        KZP_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        KZP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (ms_per_msm)
            *ms_per_msm = ms / (float)(iters > 0 ? iters : 1);
        if (entries)
        {
            uint32_t total = 0;
            KZP_CUDA_CHECK(cudaMemcpy(&total, m->sort.offsets + m->sort.shape.buckets + 1, 4, cudaMemcpyDeviceToHost));
            *entries = total;
        }
    });
}

// ------------------------------------------------------------------------------------------- field / point ops
int kzp_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t count,
                 int device)
{
    return guarded([&] {
        if (field < 0 || field > 2 || op < 0 || op > 8)
            throw FormatError("bad field/op");
        use_device(device);
        size_t esz = field == 2 ? 64 : 32;
        DevBuf da(count * esz), db(count * esz), dout(count * esz);
        KZP_CUDA_CHECK(cudaMemcpy(da.p, a, count * esz, cudaMemcpyHostToDevice));
        if (b)
            KZP_CUDA_CHECK(cudaMemcpy(db.p, b, count * esz, cudaMemcpyHostToDevice));
        field_op(field, op, da.p, b ? db.p : nullptr, dout.p, count, 0);
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
        KZP_CUDA_CHECK(cudaMemcpy(out, dout.p, count * esz, cudaMemcpyDeviceToHost));
    });
}

int kzp_point_op(int group, int op, const uint8_t* p, const uint8_t* q, uint8_t* out, uint64_t count,
                 int device)
{
    return guarded([&] {
        if (group < 0 || group > 1 || op < 0 || op > 4)
            throw FormatError("bad group/op");
        use_device(device);
        size_t psz = group == 0 ? 128 : 256;
        size_t qsz = op == 0 ? psz / 2 : psz;
        DevBuf dp(count * psz), dq(count * qsz), dout(count * psz);
        KZP_CUDA_CHECK(cudaMemcpy(dp.p, p, count * psz, cudaMemcpyHostToDevice));
        if (q && op != 2 && op != 4)
            KZP_CUDA_CHECK(cudaMemcpy(dq.p, q, count * qsz, cudaMemcpyHostToDevice));
        point_op(group, op, dp.p, dq.p, dout.p, count, 0);
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
        KZP_CUDA_CHECK(cudaMemcpy(out, dout.p, count * psz, cudaMemcpyDeviceToHost));
    });
}

int kzp_imad_peak(int iters, int device, float* ms, uint64_t* multiply_adds)
{
    return guarded([&] {
        use_device(device);
        DevBuf      sink(16);
        cudaEvent_t e0, e1;
        KZP_CUDA_CHECK(cudaEventCreate(&e0));
        KZP_CUDA_CHECK(cudaEventCreate(&e1));
        imad_probe((uint32_t*)sink.p, 64, 0); // warm-up
        KZP_CUDA_CHECK(cudaDeviceSynchronize());
        KZP_CUDA_CHECK(cudaEventRecord(e0, 0));
        uint64_t n = imad_probe((uint32_t*)sink.p, iters, 0);
        KZP_CUDA_CHECK(cudaEventRecord(e1, 0));
        KZP_CUDA_CHECK(cudaEventSynchronize(e1));
        float t = 0;
        KZP_CUDA_CHECK(cudaEventElapsedTime(&t, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (ms)
            *ms = t;
        if (multiply_adds)
            *multiply_adds = n;
    });
}

// ------------------------------------------------------------------------------------------- host-only helpers
int kzp_host_parse_zkey(const char* path, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size,
                        uint64_t* n_coefs, int* state_out)
{
    int state = KZP_STATE_OK;
    int rc    = guarded([&] {
        MappedFile file(path ? path : "");
        BinView    bin(file.data(), file.size(), "zkey", 1);
        ZkeyHeader zh = parse_zkey(bin);
        if (n_vars)
            *n_vars = zh.n_vars;
        if (n_public)
            *n_public = zh.n_public;
        if (domain_size)
            *domain_size = zh.domain_size;
        if (n_coefs)
            *n_coefs = zh.n_coefs;
    });
    if (rc == KZP_ERR_IO)
        state = KZP_STATE_ZKEY_FILE_LOAD_ERROR;
    else if (rc != KZP_OK)
        state = KZP_STATE_UNSUPPORTED_ZKEY_CURVE;
    if (state_out)
        *state_out = state;
    return rc;
}

int kzp_host_assemble(const char* zkey_path, const uint8_t* partials, int count, const uint8_t* r32,
                      const uint8_t* s32, char** json_out, uint8_t* msm_out384)
{
    if (!partials || count < 1 || !json_out)
    {
        g_last_error = "bad arguments";
        return KZP_ERR_FORMAT;
    }
    return guarded([&] {
        MappedFile file(zkey_path ? zkey_path : "");
        BinView    bin(file.data(), file.size(), "zkey", 1);
        ZkeyHeader zh = parse_zkey(bin);
        HostVk     vk;
        memcpy(vk.alpha1, zh.alpha1, 64);
        memcpy(vk.beta1, zh.beta1, 64);
        memcpy(vk.delta1, zh.delta1, 64);
        memcpy(vk.beta2, zh.beta2, 128);
        memcpy(vk.delta2, zh.delta2, 128);
        std::vector<ShardPartials> ps(count);
        for (int k = 0; k < count; k++)
            memcpy(ps[k].bytes, partials + (size_t)k * KZP_PARTIALS_BYTES, KZP_PARTIALS_BYTES);
        MsmArtefacts art;
        std::string  j = assemble_proof(vk, ps.data(), count, r32, s32, &art);
        if (msm_out384)
            memcpy(msm_out384, art.bytes, 384);
        *json_out = dup_string(j);
    });
}

int kzp_host_pack_witness_slice(const uint8_t* values, uint32_t count, uint8_t* out, uint64_t out_cap, uint64_t* packed_bytes,
                                uint32_t* n_full)
{
    return guarded([&] {
        if (!values || !out || out_cap < pack_slice_capacity() || ((uintptr_t)out & 15u))
            throw FormatError("output buffer too small or not 16-byte aligned");
        uint32_t nf = 0;
        size_t   nb = pack_witness_slice(values, count, out, &nf);
        if (packed_bytes)
            *packed_bytes = nb;
        if (n_full)
            *n_full = nf;
    });
}

int kzp_host_fq_decimal(const uint8_t* mont32, char* out, size_t cap)
{
    HFq v;
    memcpy(&v, mont32, 32);
    std::string s = HFq::to_decimal(v);
    if (s.size() + 1 > cap)
        return KZP_ERR_FORMAT;
    memcpy(out, s.c_str(), s.size() + 1);
    return KZP_OK;
}

template <class F>
static void host_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out)
{
    F x, y, r;
    memcpy(&x, a, sizeof(F));
    if (b)
        memcpy(&y, b, sizeof(F));
    else
        y = x;
    switch (op)
    {
    case 0: F::mul(r, x, y); break;
    case 1: F::add(r, x, y); break;
    case 2: F::sub(r, x, y); break;
    case 3: F::neg(r, x); break;
    case 6: F::sqr(r, x); break;
    case 7: F::inv(r, x); break;
    case 8:
        if constexpr (F::kFusedMulAdd2)
            F::mul_add2(r, x, y, y, y);
        else
        {
            F t, u;
            F::mul(t, x, y);
            F::mul(u, y, y);
            F::add(r, t, u);
        }
        break;
    default: r = x;
    }
    memcpy(out, &r, sizeof(F));
}

// fields 0..2: host 64-bit-limb implementation used by the proof assembly; fields 10..12: the portable
// 32-bit-limb path of the device templates (ff.cuh compiled for the host), so the CPU test-suite can
// check the exact formulas the kernels instantiate.
int kzp_host_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out)
{
    return guarded([&] {
        auto mont = [&](auto tag) {
            typedef decltype(tag) F;
            F x, r;
            memcpy(&x, a, sizeof(F));
            if (op == 4)
                F::to_mont(r, x);
            else
                F::from_mont(r, x);
            memcpy(out, &r, sizeof(F));
        };
        bool is_mont = (op == 4 || op == 5);
        switch (field)
        {
        case 0: is_mont ? mont(HFr()) : host_op<HFr>(op, a, b, out); break;
        case 1: is_mont ? mont(HFq()) : host_op<HFq>(op, a, b, out); break;
        case 2: host_op<Fp2T<HFq>>(op, a, b, out); break;
        case 10: is_mont ? mont(Fr()) : host_op<Fr>(op, a, b, out); break;
        case 11: is_mont ? mont(Fq()) : host_op<Fq>(op, a, b, out); break;
        case 12: host_op<Fq2>(op, a, b, out); break;
        default: throw FormatError("bad field");
        }
    });
}
