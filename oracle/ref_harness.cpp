/* TEST INFRASTRUCTURE — NOT part of the product. Only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline leg may load what this file builds.
 *
 * Thin C-ABI harness around the UNMODIFIED reference prover sources that live under
 * /root/reference/rust-rapidsnark/rapidsnark/src (compiled in place by oracle/Makefile
 * into oracle/_ref/libkzp_ref.so; nothing is copied into this repository). It exposes
 * the reference's own entry points so parity tests can compare artefact by artefact:
 *
 *   kzp_ref_prove        -> FullProver::FullProver + FullProver::prove (fullprover.hpp:54-64)
 *   kzp_ref_dump         -> re-drives Prover::prove's stages (groth16.cpp:88-283) through the
 *                           public FFT / Curve::multiMulByScalar APIs and returns a,b after
 *                           SpMV, the H coefficients and the five MSM results (affine, canonical)
 *   kzp_ref_fr_ntt       -> FFT<RawFr>::fft / ifft (fft.cpp:192-246)
 *   kzp_ref_msm_g1/g2    -> Curve::multiMulByScalar (curve.hpp:209-215, multiexp.cpp:183-245)
 *   kzp_ref_f{r,q}_*     -> RawFr/RawFq raw Montgomery ops (fr.hpp:206-281)
 *   kzp_ref_g{1,2}_mul   -> Curve::mulByScalar (exp.hpp:9-31)
 *
 * The blinding scalars (r,s) are injected through the sodium.h shim (oracle/shim_sodium),
 * which is what groth16.cpp's randombytes_buf() resolves to when built with -DUSE_SODIUM.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <string>

#include <fcntl.h>
#include <omp.h>
#include <unistd.h>

#include "alt_bn128.hpp"
#include "binfile_utils.hpp"
#include "fft.hpp"
#include "fullprover.hpp"
#include "groth16.hpp"
#include "wtns_utils.hpp"
#include "zkey_utils.hpp"

// ---------------------------------------------------------------- r,s injection
static std::mutex g_rs_mutex;
static bool       g_rs_fixed = false;
static int        g_rs_next  = 0;
static uint8_t    g_rs[2][32];

extern "C" void kzp_oracle_fixed_rs_set(const uint8_t* r32, const uint8_t* s32)
{
    std::lock_guard<std::mutex> lock(g_rs_mutex);
    memcpy(g_rs[0], r32, 32);
    memcpy(g_rs[1], s32, 32);
    g_rs_fixed = true;
    g_rs_next  = 0;
}

extern "C" void kzp_oracle_fixed_rs_clear(void)
{
    std::lock_guard<std::mutex> lock(g_rs_mutex);
    g_rs_fixed = false;
    g_rs_next  = 0;
}

extern "C" void randombytes_buf(void* const buf, const size_t size)
{
    std::lock_guard<std::mutex> lock(g_rs_mutex);
    if (g_rs_fixed && size == 32)
    {
        memcpy(buf, g_rs[g_rs_next & 1], 32);
        g_rs_next++;
        return;
    }
    std::random_device                          engine;
    std::uniform_int_distribution<unsigned int> distr(0, 255);
    uint8_t*                                    p = static_cast<uint8_t*>(buf);
    for (size_t i = 0; i < size; i++)
        p[i] = (uint8_t)distr(engine);
}

// ---------------------------------------------------------------- helpers
namespace
{

using Engine = AltBn128::Engine;

// The reference logs to stdout (fullprover.cpp:67-78,138,237). Callers that need a
// clean stdout (bench.py prints exactly one JSON line) ask for it to be silenced.
struct StdoutSilencer
{
    int saved = -1;
    explicit StdoutSilencer(bool on)
    {
        if (!on)
            return;
        fflush(stdout);
        saved  = dup(1);
        int dn = open("/dev/null", O_WRONLY);
        dup2(dn, 1);
        close(dn);
    }
    ~StdoutSilencer()
    {
        if (saved < 0)
            return;
        fflush(stdout);
        std::cout.flush();
        dup2(saved, 1);
        close(saved);
    }
};

void g1_to_canonical(uint8_t* out64, Engine::G1Point& p)
{
    auto&                 E = Engine::engine;
    Engine::G1PointAffine a;
    E.g1.copy(a, p);
    Engine::F1Element t;
    E.f1.fromMontgomery(t, a.x);
    memcpy(out64, &t, 32);
    E.f1.fromMontgomery(t, a.y);
    memcpy(out64 + 32, &t, 32);
}

void g2_to_canonical(uint8_t* out128, Engine::G2Point& p)
{
    auto&                 E = Engine::engine;
    Engine::G2PointAffine a;
    E.g2.copy(a, p);
    Engine::F1Element t;
    E.f1.fromMontgomery(t, a.x.a);
    memcpy(out128, &t, 32);
    E.f1.fromMontgomery(t, a.x.b);
    memcpy(out128 + 32, &t, 32);
    E.f1.fromMontgomery(t, a.y.a);
    memcpy(out128 + 64, &t, 32);
    E.f1.fromMontgomery(t, a.y.b);
    memcpy(out128 + 96, &t, 32);
}

} // namespace

extern "C"
{

    int  kzp_ref_num_threads(void) { return omp_get_max_threads(); }
    void kzp_ref_set_threads(int n) { omp_set_num_threads(n); }

    // ---- persistent FullProver handle (so the zkey is loaded once, like prover-service)
    void* kzp_ref_prover_new(const char* zkey_path, int quiet, int* state_out)
    {
        StdoutSilencer sil(quiet != 0);
        FullProver*    p = new FullProver(zkey_path);
        // FullProver = { impl*, state } (fullprover.hpp:56-57); state is private, read by offset
        // exactly like the Rust binding does (rust-rapidsnark/src/lib.rs:53).
        int st = *reinterpret_cast<int*>(reinterpret_cast<char*>(p) + sizeof(void*));
        if (state_out)
            *state_out = st;
        return p;
    }

    void kzp_ref_prover_free(void* h) { delete static_cast<FullProver*>(h); }

    // returns: 0 ok (json copied), >0 ProverError, -1 buffer too small
    int kzp_ref_prover_prove(void* h, const char* wtns_path, const uint8_t* r32,
                             const uint8_t* s32, int quiet, char* json_out, size_t json_cap,
                             int* prover_time_ms)
    {
        StdoutSilencer sil(quiet != 0);
        if (r32 && s32)
            kzp_oracle_fixed_rs_set(r32, s32);
        else
            kzp_oracle_fixed_rs_clear();
        ProverResponse resp = static_cast<FullProver*>(h)->prove(wtns_path);
        kzp_oracle_fixed_rs_clear();
        if (prover_time_ms)
            *prover_time_ms = resp.metrics.prover_time;
        if (resp.type != ProverResponseType::SUCCESS)
            return (int)resp.error;
        size_t len = strlen(resp.raw_json);
        if (len + 1 > json_cap)
            return -1;
        memcpy(json_out, resp.raw_json, len + 1);
        return 0;
    }

    int kzp_ref_prove(const char* zkey_path, const char* wtns_path, const uint8_t* r32,
                      const uint8_t* s32, int quiet, char* json_out, size_t json_cap,
                      int* prover_time_ms)
    {
        int   st = 0;
        void* h  = kzp_ref_prover_new(zkey_path, quiet, &st);
        if (st != 0)
        {
            kzp_ref_prover_free(h);
            return 100 + st;
        }
        int rc = kzp_ref_prover_prove(h, wtns_path, r32, s32, quiet, json_out, json_cap,
                                      prover_time_ms);
        kzp_ref_prover_free(h);
        return rc;
    }

    /* Re-drive the stages of Prover::prove (groth16.cpp:88-283) and hand back every parity
     * artefact of SURVEY.md Appendix C. Any output pointer may be NULL.
     *   ab_out : 2*domainSize*32 B  a then b after the SpMV (Montgomery)
     *   h_out  : domainSize*32 B    H coefficients (canonical, natural index)
     *   msm_out: 384 B              A(64) B1(64) B2(128) C(64) H(64), affine canonical LE
     */
    int kzp_ref_dump(const char* zkey_path, const char* wtns_path, uint8_t* ab_out,
                     uint8_t* h_out, uint8_t* msm_out)
    {
        try
        {
            auto& E    = Engine::engine;
            auto  zkey = BinFileUtils::BinFile::make_from_file(zkey_path, "zkey", 1);
            auto  zh   = ZKeyUtils::Header::make_from_bin_file(*zkey);
            auto  wtns = BinFileUtils::BinFile::make_from_file(wtns_path, "wtns", 2);
            auto  wh   = WtnsUtils::Header::make_from_bin_file(*wtns);
            (void)wh;

            uint32_t nVars = zh->nVars, nPublic = zh->nPublic, N = zh->domainSize;
            uint64_t nCoefs = zh->nCoefs;
            auto*    w      = (Engine::FrElement*)wtns->getSectionData(2);
            auto*    coefs =
                (Groth16::Coef<Engine>*)((uint64_t)zkey->getSectionData(4) + 4);
            auto* pointsA  = (Engine::G1PointAffine*)zkey->getSectionData(5);
            auto* pointsB1 = (Engine::G1PointAffine*)zkey->getSectionData(6);
            auto* pointsB2 = (Engine::G2PointAffine*)zkey->getSectionData(7);
            auto* pointsC  = (Engine::G1PointAffine*)zkey->getSectionData(8);
            auto* pointsH  = (Engine::G1PointAffine*)zkey->getSectionData(9);

            Engine::G1Point pA, pB1, pC, pH;
            Engine::G2Point pB2;
            uint32_t        sW = sizeof(w[0]);
            if (msm_out)
            {
                E.g1.multiMulByScalar(pA, pointsA, (uint8_t*)w, sW, nVars);
                E.g1.multiMulByScalar(pB1, pointsB1, (uint8_t*)w, sW, nVars);
                E.g2.multiMulByScalar(pB2, pointsB2, (uint8_t*)w, sW, nVars);
                E.g1.multiMulByScalar(pC, pointsC, (uint8_t*)(w + nPublic + 1), sW,
                                      nVars - nPublic - 1);
            }

            std::unique_ptr<Engine::FrElement[]> a(new Engine::FrElement[N]);
            std::unique_ptr<Engine::FrElement[]> b(new Engine::FrElement[N]);
            std::unique_ptr<Engine::FrElement[]> c(new Engine::FrElement[N]);
            for (uint32_t i = 0; i < N; i++)
            {
                E.fr.copy(a[i], E.fr.zero());
                E.fr.copy(b[i], E.fr.zero());
            }
            // groth16.cpp:141-156 (sequential here: same field sums, order-independent)
            for (uint64_t i = 0; i < nCoefs; i++)
            {
                Engine::FrElement* ab = (coefs[i].m == 0) ? a.get() : b.get();
                Engine::FrElement  aux;
                E.fr.mul(aux, w[coefs[i].s], coefs[i].coef);
                E.fr.add(ab[coefs[i].c], ab[coefs[i].c], aux);
            }
            if (ab_out)
            {
                memcpy(ab_out, a.get(), (size_t)N * 32);
                memcpy(ab_out + (size_t)N * 32, b.get(), (size_t)N * 32);
            }
            for (uint32_t i = 0; i < N; i++)
                E.fr.mul(c[i], a[i], b[i]);

            FFT<Engine::Fr>    fft(2 * (uint64_t)N);
            uint32_t           domainPower = fft.log2(N);
            Engine::FrElement* vecs[3]     = {a.get(), b.get(), c.get()};
            for (auto* x : vecs)
            {
                fft.ifft(x, N);
                for (uint32_t i = 0; i < N; i++)
                    E.fr.mul(x[i], x[i], fft.root(domainPower + 1, i));
                fft.fft(x, N);
            }
            for (uint32_t i = 0; i < N; i++)
            {
                E.fr.mul(a[i], a[i], b[i]);
                E.fr.sub(a[i], a[i], c[i]);
                E.fr.fromMontgomery(a[i], a[i]);
            }
            if (h_out)
                memcpy(h_out, a.get(), (size_t)N * 32);
            if (msm_out)
            {
                E.g1.multiMulByScalar(pH, pointsH, (uint8_t*)a.get(), sizeof(a[0]), N);
                g1_to_canonical(msm_out + 0, pA);
                g1_to_canonical(msm_out + 64, pB1);
                g2_to_canonical(msm_out + 128, pB2);
                g1_to_canonical(msm_out + 256, pC);
                g1_to_canonical(msm_out + 320, pH);
            }
            return 0;
        }
        catch (std::exception& e)
        {
            fprintf(stderr, "kzp_ref_dump: %s\n", e.what());
            return 1;
        }
    }

    // ---- component entry points -------------------------------------------------
    // data: n Fr elements, Montgomery form, natural order, in place.
    int kzp_ref_fr_ntt(uint8_t* data, uint64_t n, int inverse)
    {
        try
        {
            FFT<Engine::Fr> fft(n < 2 ? 2 : n);
            if (inverse)
                fft.ifft((Engine::FrElement*)data, n);
            else
                fft.fft((Engine::FrElement*)data, n);
            return 0;
        }
        catch (std::exception& e)
        {
            return 1;
        }
    }

    // root(domainPow, idx) of an FFT built for maxDomain (fft.hpp:40-43), Montgomery.
    int kzp_ref_fr_root(uint64_t maxDomain, uint32_t domainPow, uint64_t idx, uint8_t* out32)
    {
        FFT<Engine::Fr> fft(maxDomain);
        memcpy(out32, &fft.root(domainPow, idx), 32);
        return 0;
    }

    // bases: n affine Montgomery points (64 B / 128 B); scalars: n x 32 B canonical LE.
    void kzp_ref_msm_g1(const uint8_t* bases, const uint8_t* scalars, uint64_t n,
                        uint8_t* out64)
    {
        auto&           E = Engine::engine;
        Engine::G1Point r;
        E.g1.multiMulByScalar(r, (Engine::G1PointAffine*)bases, (uint8_t*)scalars, 32,
                              (unsigned int)n);
        g1_to_canonical(out64, r);
    }

    void kzp_ref_msm_g2(const uint8_t* bases, const uint8_t* scalars, uint64_t n,
                        uint8_t* out128)
    {
        auto&           E = Engine::engine;
        Engine::G2Point r;
        E.g2.multiMulByScalar(r, (Engine::G2PointAffine*)bases, (uint8_t*)scalars, 32,
                              (unsigned int)n);
        g2_to_canonical(out128, r);
    }

    // base: affine Montgomery; scalar 32 B canonical LE; out: affine Montgomery.
    void kzp_ref_g1_mul(const uint8_t* base64, const uint8_t* scalar32, uint8_t* out64)
    {
        auto&                 E = Engine::engine;
        Engine::G1Point       r;
        Engine::G1PointAffine b, ra;
        memcpy(&b, base64, 64);
        E.g1.mulByScalar(r, b, (uint8_t*)scalar32, 32);
        E.g1.copy(ra, r);
        memcpy(out64, &ra, 64);
    }

    void kzp_ref_g2_mul(const uint8_t* base128, const uint8_t* scalar32, uint8_t* out128)
    {
        auto&                 E = Engine::engine;
        Engine::G2Point       r;
        Engine::G2PointAffine b, ra;
        memcpy(&b, base128, 128);
        E.g2.mulByScalar(r, b, (uint8_t*)scalar32, 32);
        E.g2.copy(ra, r);
        memcpy(out128, &ra, 128);
    }

    // affine Montgomery + affine Montgomery -> affine Montgomery
    void kzp_ref_g1_add(const uint8_t* p64, const uint8_t* q64, uint8_t* out64)
    {
        auto&                 E = Engine::engine;
        Engine::G1PointAffine p, q, ra;
        Engine::G1Point       r;
        memcpy(&p, p64, 64);
        memcpy(&q, q64, 64);
        E.g1.add(r, p, q);
        E.g1.copy(ra, r);
        memcpy(out64, &ra, 64);
    }

    void kzp_ref_g2_add(const uint8_t* p128, const uint8_t* q128, uint8_t* out128)
    {
        auto&                 E = Engine::engine;
        Engine::G2PointAffine p, q, ra;
        Engine::G2Point       r;
        memcpy(&p, p128, 128);
        memcpy(&q, q128, 128);
        E.g2.add(r, p, q);
        E.g2.copy(ra, r);
        memcpy(out128, &ra, 128);
    }

    // op: 0 mont-mul, 1 add, 2 sub, 3 neg(a), 4 toMontgomery(a), 5 fromMontgomery(a),
    //     6 mont-square(a), 7 inverse(a) (Montgomery in/out)
    void kzp_ref_fr_op(int op, const uint8_t* a32, const uint8_t* b32, uint8_t* out32,
                       uint64_t count)
    {
        auto& F = Engine::engine.fr;
        for (uint64_t i = 0; i < count; i++)
        {
            Engine::FrElement a, b, r;
            memcpy(&a, a32 + 32 * i, 32);
            if (b32)
                memcpy(&b, b32 + 32 * i, 32);
            switch (op)
            {
            case 0: F.mul(r, a, b); break;
            case 1: F.add(r, a, b); break;
            case 2: F.sub(r, a, b); break;
            case 3: F.neg(r, a); break;
            case 4: F.toMontgomery(r, a); break;
            case 5: F.fromMontgomery(r, a); break;
            case 6: F.square(r, a); break;
            case 7: F.inv(r, a); break;
            default: memset(&r, 0, 32);
            }
            memcpy(out32 + 32 * i, &r, 32);
        }
    }

    void kzp_ref_fq_op(int op, const uint8_t* a32, const uint8_t* b32, uint8_t* out32,
                       uint64_t count)
    {
        auto& F = Engine::engine.f1;
        for (uint64_t i = 0; i < count; i++)
        {
            Engine::F1Element a, b, r;
            memcpy(&a, a32 + 32 * i, 32);
            if (b32)
                memcpy(&b, b32 + 32 * i, 32);
            switch (op)
            {
            case 0: F.mul(r, a, b); break;
            case 1: F.add(r, a, b); break;
            case 2: F.sub(r, a, b); break;
            case 3: F.neg(r, a); break;
            case 4: F.toMontgomery(r, a); break;
            case 5: F.fromMontgomery(r, a); break;
            case 6: F.square(r, a); break;
            case 7: F.inv(r, a); break;
            default: memset(&r, 0, 32);
            }
            memcpy(out32 + 32 * i, &r, 32);
        }
    }

    // Fq2 ops on 64-byte elements (a then b of a+bu): 0 mul, 1 add, 2 sub, 3 neg, 6 square, 7 inv
    void kzp_ref_fq2_op(int op, const uint8_t* a64, const uint8_t* b64, uint8_t* out64,
                        uint64_t count)
    {
        auto& F = Engine::engine.f2;
        for (uint64_t i = 0; i < count; i++)
        {
            Engine::F2Element a, b, r;
            memcpy(&a, a64 + 64 * i, 64);
            if (b64)
                memcpy(&b, b64 + 64 * i, 64);
            switch (op)
            {
            case 0: F.mul(r, a, b); break;
            case 1: F.add(r, a, b); break;
            case 2: F.sub(r, a, b); break;
            case 3: F.neg(r, a); break;
            case 6: F.square(r, a); break;
            case 7: F.inv(r, a); break;
            default: memset(&r, 0, 64);
            }
            memcpy(out64 + 64 * i, &r, 64);
        }
    }

} // extern "C"
