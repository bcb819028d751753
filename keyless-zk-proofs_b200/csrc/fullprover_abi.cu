// Itanium-ABI shim: defines the reference's FullProver / ProverResponse symbols on top of the C ABI, so
// rust-rapidsnark's bindgen binding (src/lib.rs:41-106) links against libkzp_b200 unchanged.
// Behaviour per SURVEY.md §8(b):
//   * the constructor never throws; failures become FullProverState values (fullprover.cpp:80-101)
//   * prove() on a failed prover returns PROVER_NOT_READY (fullprover.cpp:114-125)
//   * raw_json is malloc'd and stays valid until ~ProverResponse (which Rust never runs, lib.rs:68-77):
//     no buffer is ever reused
//   * prove may be called from any thread, one call at a time per object (prover_state.rs:21):
//     every call selects its CUDA device itself
// Configuration comes from the environment because the header is frozen:
//   KZP_DEVICE   CUDA device index (default 0)
//   KZP_FIXED_RS 128 hex digits = r then s, 32 bytes each little-endian — ONLY in a translation unit compiled with
//                -DKZP_TEST_HOOKS (the ABI parity test builds its own copy of this file that way and links it in front
//                of the library). The release library is built without it: fixing r and s removes zero-knowledge.
//   KZP_LOG      when set, print the reference's stdout log lines (fullprover.cpp:67-78,237-238)
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/fullprover_b200.hpp"
#include "../../include/kzp_b200.h"

class FullProverImpl
{
public:
    kzp_prover* handle = nullptr;
};

// layout checks; the members are private, so measure through a standard-layout mirror
namespace
{
struct FullProverMirror
{
    void* impl;
    int   state;
};
static_assert(sizeof(FullProver) == 16 && sizeof(FullProverMirror) == 16, "FullProver layout");
static_assert(offsetof(FullProverMirror, state) == 8, "FullProver::state offset");
static_assert(sizeof(ProverResponse) == 24, "ProverResponse layout");
static_assert(offsetof(ProverResponse, raw_json) == 8 && offsetof(ProverResponse, error) == 16 &&
                  offsetof(ProverResponse, metrics) == 20,
              "ProverResponse member offsets");

#if defined(KZP_TEST_HOOKS)
bool parse_fixed_rs(unsigned char* r, unsigned char* s)
{
    const char* env = getenv("KZP_FIXED_RS");
    if (!env || strlen(env) != 128)
        return false;
    auto hex = [](char c) -> int {
        if (c >= '0' && c <= '9')
            return c - '0';
        if (c >= 'a' && c <= 'f')
            return c - 'a' + 10;
        if (c >= 'A' && c <= 'F')
            return c - 'A' + 10;
        return -1;
    };
    for (int i = 0; i < 64; i++)
    {
        int hi = hex(env[2 * i]), lo = hex(env[2 * i + 1]);
        if (hi < 0 || lo < 0)
            return false;
        unsigned char b = (unsigned char)(hi * 16 + lo);
        if (i < 32)
            r[i] = b;
        else
            s[i - 32] = b;
    }
    return true;
}
#else
bool parse_fixed_rs(unsigned char*, unsigned char*) { return false; } // release build: r, s always come from the CSPRNG
#endif

void log_line(const char* level, const char* msg)
{
    if (!getenv("KZP_LOG"))
        return;
    // the message may quote a path or a CUDA error string: keep the line valid JSON
    char   esc[512];
    size_t k = 0;
    for (const char* p = msg ? msg : ""; *p && k + 7 < sizeof esc; p++)
    {
        unsigned char c = (unsigned char)*p;
        if (c == '"' || c == '\\')
        {
            esc[k++] = '\\';
            esc[k++] = (char)c;
        }
        else if (c < 0x20)
            k += (size_t)snprintf(esc + k, 7, "\\u%04x", c);
        else
            esc[k++] = (char)c;
    }
    esc[k] = 0;
    printf("{\"level\":\"%s\",\"message\":\"%s\",\"native_code\":\"1\",\"target\":\"prover_service::rapidsnark\"}\n",
           level, esc);
    fflush(stdout);
}
} // namespace

char const* const ProverResponse::empty_string = "";

ProverResponse::ProverResponse(ProverError _error)
    : type(ProverResponseType::ERROR)
    , raw_json(ProverResponse::empty_string)
    , error(_error)
    , metrics(ProverResponseMetrics())
{
}

ProverResponse::ProverResponse(const char* _raw_json, ProverResponseMetrics _metrics)
    : type(ProverResponseType::SUCCESS)
    , raw_json(_raw_json)
    , error(ProverError::NONE)
    , metrics(_metrics)
{
}

ProverResponse::~ProverResponse()
{
    if (raw_json != empty_string)
        free(const_cast<char*>(raw_json));
}

FullProver::FullProver(const char* _zkeyFileName)
    : impl(nullptr)
    , state(FullProverState::ZKEY_FILE_LOAD_ERROR)
{
    int st       = KZP_STATE_ZKEY_FILE_LOAD_ERROR;
    kzp_prover* h = kzp_prover_new(_zkeyFileName, -1, &st);
    if (h)
    {
        impl = new (std::nothrow) FullProverImpl();
        if (impl)
            impl->handle = h;
        else
        {
            kzp_prover_free(h);
            st = KZP_STATE_ZKEY_FILE_LOAD_ERROR;
        }
    }
    state = st == KZP_STATE_OK ? FullProverState::OK
                               : (st == KZP_STATE_UNSUPPORTED_ZKEY_CURVE
                                      ? FullProverState::UNSUPPORTED_ZKEY_CURVE
                                      : FullProverState::ZKEY_FILE_LOAD_ERROR);
}

FullProver::~FullProver()
{
    if (impl)
    {
        kzp_prover_free(impl->handle);
        delete impl;
    }
}

ProverResponse FullProver::prove(const char* input) const
{
    if (state != FullProverState::OK || !impl)
        return ProverResponse(ProverError::PROVER_NOT_READY);
    log_line("INFO", "FullProverImpl::prove begin");
    unsigned char r[32], s[32];
    bool          fixed = parse_fixed_rs(r, s);
    char*         json  = nullptr;
    int           err = 0, ms = 0;
    int rc = kzp_prover_prove(impl->handle, input, fixed ? r : nullptr, fixed ? s : nullptr, &json, &err, &ms);
    if (rc != KZP_RESPONSE_SUCCESS || !json)
    {
        log_line("ERROR", kzp_last_error());
        return ProverResponse((ProverError)err);
    }
    if (getenv("KZP_LOG"))
        printf("Time taken for Groth16 prover: %d milliseconds\n", ms);
    log_line("INFO", "FullProverImpl::prove end");
    ProverResponseMetrics m;
    m.prover_time = ms;
    return ProverResponse(json, m);
}
