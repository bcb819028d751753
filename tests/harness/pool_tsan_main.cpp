// Test harness (tests/test_host_library.py::test_pool_under_thread_sanitizer): csrc/pool.cpp — the real checkout queue,
// fault retirement, fused-verify path and the drain in kzp_pool_free — rebuilt with -fsanitize=thread on top of STUB
// provers (no GPU, no CUDA): a stub proof holds its prover for a few hundred microseconds, refuses to be entered by
// two callers at once, and can be told to "fault" (its state turns KZP_STATE_DEVICE_FAULT, which retires the slot).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/kzp_b200.h"

struct kzp_prover
{
    int              device = 0;
    std::atomic<int> state{KZP_STATE_OK};
    std::atomic<int> inside{0};
    std::atomic<int> served{0};
};

static std::atomic<int> g_overlap{0};     // two callers inside one prover at once
static std::atomic<int> g_live_provers{0};
static std::atomic<int> g_use_after_free{0};
static std::atomic<int> g_fault_every{0}; // the k-th proof of every prover on device 1 faults
static std::atomic<int> g_hold_us{200};   // how long a stub proof holds its prover
static std::atomic<int> g_verify_calls{0};

extern "C" {
int         kzp_device_count(void) { return 4; }
const char* kzp_last_error(void) { return ""; }
void        kzp_free(void* p) { free(p); }
kzp_prover* kzp_prover_new(const char*, int device, int* state_out)
{
    kzp_prover* p = new kzp_prover();
    p->device     = device;
    g_live_provers++;
    if (state_out)
        *state_out = KZP_STATE_OK;
    return p;
}
void kzp_prover_free(kzp_prover* p)
{
    if (!p)
        return;
    if (p->inside.load() != 0)
        g_use_after_free++;
    g_live_provers--;
    delete p;
}
int kzp_prover_state(const kzp_prover* p) { return p ? p->state.load() : KZP_STATE_ZKEY_FILE_LOAD_ERROR; }

static int stub_prove(kzp_prover* p, char** json_out, int* error_out, int* ms)
{
    if (p->inside.fetch_add(1) != 0)
        g_overlap++;
    std::this_thread::sleep_for(std::chrono::microseconds(g_hold_us.load()));
    int k  = p->served.fetch_add(1) + 1;
    int rc = KZP_RESPONSE_SUCCESS;
    if (p->device == 1 && g_fault_every.load() > 0 && k == g_fault_every.load())
    {
        p->state = KZP_STATE_DEVICE_FAULT;
        if (error_out)
            *error_out = KZP_PROVER_ERROR_NOT_READY;
        rc = KZP_RESPONSE_ERROR;
    }
    else
    {
        *json_out = strdup("{\"stub\":1}");
        if (error_out)
            *error_out = KZP_PROVER_ERROR_NONE;
        if (ms)
            *ms = 1;
    }
    p->inside.fetch_sub(1);
    return rc;
}
int kzp_prover_prove(kzp_prover* p, const char*, const uint8_t*, const uint8_t*, char** json_out, int* error_out, int* ms)
{
    return stub_prove(p, json_out, error_out, ms);
}
int kzp_prover_prove_mem(kzp_prover* p, const uint8_t*, uint64_t, const uint8_t*, const uint8_t*, char** json_out,
                         int* error_out, int* ms)
{
    return stub_prove(p, json_out, error_out, ms);
}
int kzp_host_parse_zkey(const char*, uint32_t*, uint32_t* n_public, uint32_t*, uint64_t*, int* state_out)
{
    if (n_public)
        *n_public = 1;
    if (state_out)
        *state_out = KZP_STATE_OK;
    return KZP_OK;
}
// "verifies" iff the public signal is even
int kzp_host_verify(const char*, const char*, const uint8_t* public32, uint32_t, int* valid_out)
{
    g_verify_calls++;
    *valid_out = (public32[0] & 1) ? 0 : 1;
    return KZP_OK;
}
}

int main()
{
    int devices[6] = {0, 1, 2, 3, 0, 1};
    int failures   = 0;
    // phase 1: 16 clients, 6 provers, of which the two on device 1 fault on their 5th proof; verify on; odd public
    // signals are rejected by the (stub) verifier
    {
        int       st   = -1;
        kzp_pool* pool = kzp_pool_new("stub.zkey", devices, 6, &st);
        if (!pool || st != KZP_STATE_OK || kzp_pool_size(pool) != 6)
            return 10;
        kzp_pool_set_verify(pool, 1);
        g_fault_every = 5;
        std::atomic<int>         ok{0}, rejected{0}, not_ready{0};
        std::vector<std::thread> th;
        for (int t = 0; t < 16; t++)
            th.emplace_back([&, t] {
                for (int j = 0; j < 40; j++)
                {
                    uint8_t w[64] = {0};
                    w[0]          = 1;
                    w[32]         = (uint8_t)((t + j) & 1); // public signal: odd ones do not "verify"
                    char* js      = nullptr;
                    int   err = -1, ms = 0, slot = -1;
                    int   rc = kzp_pool_prove_mem(pool, w, 2, nullptr, nullptr, &js, &err, &ms, &slot);
                    if (rc == KZP_RESPONSE_SUCCESS && js && !(w[32] & 1))
                        ok++;
                    else if (rc == KZP_RESPONSE_ERROR && !js && err == KZP_PROVER_ERROR_INVALID_INPUT && (w[32] & 1))
                        rejected++;
                    else if (rc == KZP_RESPONSE_ERROR && !js && err == KZP_PROVER_ERROR_NOT_READY)
                        not_ready++;
                    else
                        failures++;
                    kzp_free(js);
                }
            });
        for (auto& t : th)
            t.join();
        uint64_t per[8] = {0}, maxw = 0;
        int      n      = kzp_pool_stats(pool, per, 8, &maxw);
        uint64_t total  = 0;
        for (int i = 0; i < n; i++)
            total += per[i];
        printf("phase1 ok %d rejected %d not_ready %d healthy %d total %llu maxw %llu\n", ok.load(), rejected.load(),
               not_ready.load(), kzp_pool_healthy(pool), (unsigned long long)total, (unsigned long long)maxw);
        // one proof per faulting prover was lost, both slots retired, every request was answered; a proof that fails
        // the check AFTER its prover has faulted under the next caller is NOT_READY as well (documented: the device
        // is suspected before the witness), so not_ready may exceed 2 — by at most the 4 proofs each faulting prover served before
        if (not_ready < 2 || not_ready > 10 || g_verify_calls != 638 || kzp_pool_healthy(pool) != 4 || total != 640 ||
            ok + rejected + not_ready != 640 || ok < 318)
            failures++;
        kzp_pool_free(pool);
    }
    // phase 2: free the pool while 10 callers are queued behind 2 busy provers; every call must return (the queued
    // ones with NOT_READY), no prover may be freed while a caller is inside it, nothing may touch the pool afterwards.
    // (kzp_pool_free's contract: no NEW call may start once it is called; the calls already inside are the pool's to
    // drain — so wait until all 12 have entered the queue, which the deepest-queue statistic shows.)
    g_fault_every = 0;
    g_hold_us     = 100000;
    for (int round = 0; round < 5; round++)
    {
        int       st   = -1;
        kzp_pool* pool = kzp_pool_new("stub.zkey", devices, 2, &st);
        std::atomic<int>         answered{0}, proved{0}, turned_away{0};
        std::vector<std::thread> th;
        for (int t = 0; t < 12; t++)
            th.emplace_back([&] {
                uint8_t w[64] = {0};
                char*   js    = nullptr;
                int     err = -1, ms = 0, slot = -1;
                int     rc = kzp_pool_prove_mem(pool, w, 2, nullptr, nullptr, &js, &err, &ms, &slot);
                if (rc == KZP_RESPONSE_SUCCESS && js)
                    proved++;
                else if (rc == KZP_RESPONSE_ERROR && !js && err == KZP_PROVER_ERROR_NOT_READY)
                    turned_away++;
                else
                    failures++;
                kzp_free(js);
                answered++;
            });
        uint64_t per[2], maxw = 0;
        while (maxw < 10 && answered.load() < 12) // all 12 are inside the pool (or, on a very slow box, already done)
        {
            kzp_pool_stats(pool, per, 2, &maxw);
            std::this_thread::sleep_for(std::chrono::microseconds(100));
        }
        kzp_pool_free(pool);
        for (auto& t : th)
            t.join();
        if (answered != 12 || proved < 2 || proved + turned_away != 12)
            failures++;
    }
    printf("overlap %d use_after_free %d live_provers %d verify_calls %d failures %d\n", g_overlap.load(),
           g_use_after_free.load(), g_live_provers.load(), g_verify_calls.load(), failures);
    return (g_overlap || g_use_after_free || g_live_provers || failures) ? 1 : 0;
}
