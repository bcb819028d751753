// kzp_pool_*: a set of resident provers with GPU-per-request checkout (include/kzp_b200.h, pool.hpp).
// Replaces, on the service side, the single mutex-guarded FullProver of prover-service/src/prover_state.rs:21-47;
// INTEGRATION.md shows the Rust patch that binds it.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kzp_b200.h"
#include "binfile.hpp"
#include "pool.hpp"

using namespace kzp;

struct kzp_pool
{
    std::vector<kzp_prover*> provers;
    std::vector<int>         devices;
    SlotScheduler*           sched = nullptr;
    int                      state = KZP_STATE_OK;
    std::string              zkey_path;
    uint32_t                 n_public = 0;
    std::atomic<int>         verify{0};    // verify every proof under the zkey's VK before returning it
    std::atomic<int>         in_flight{0}; // callers inside pool_run (queueing, proving or verifying)
};

namespace
{
struct InFlight
{
    std::atomic<int>& n;
    explicit InFlight(std::atomic<int>& c) : n(c) { n.fetch_add(1); }
    ~InFlight() { n.fetch_sub(1); }
};
} // namespace

static std::vector<int> pool_devices(const int* devices, int n)
{
    std::vector<int> out;
    if (devices && n > 0)
    {
        out.assign(devices, devices + n);
        return out;
    }
    // $KZP_POOL_DEVICES = "0,1,2,3" (a device may be listed more than once: several provers on one GPU)
    if (const char* env = getenv("KZP_POOL_DEVICES"))
    {
        const char* p = env;
        while (*p)
        {
            char* end = nullptr;
            long  v   = strtol(p, &end, 10);
            if (end == p)
                break;
            out.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
        if (!out.empty())
            return out;
    }
    int cnt = kzp_device_count();
    for (int i = 0; i < cnt; i++)
        out.push_back(i);
    return out;
}

kzp_pool* kzp_pool_new(const char* zkey_path, const int* devices, int n_devices, int* state_out)
{
    kzp_pool* pool = new (std::nothrow) kzp_pool();
    if (!pool)
        return nullptr;
    try
    {
        pool->zkey_path = zkey_path ? zkey_path : "";
        pool->devices   = pool_devices(devices, n_devices);
        size_t n      = pool->devices.size();
        pool->provers.assign(n, nullptr);
        std::vector<int>         states(n, KZP_STATE_ZKEY_FILE_LOAD_ERROR);
        std::vector<std::thread> th;
        // keys are parsed and uploaded concurrently, one loader thread per prover
        for (size_t i = 0; i < n; i++)
            th.emplace_back([&, i] { pool->provers[i] = kzp_prover_new(zkey_path, pool->devices[i], &states[i]); });
        for (auto& t : th)
            t.join();
        pool->state = n == 0 ? KZP_STATE_ZKEY_FILE_LOAD_ERROR : KZP_STATE_OK;
        for (size_t i = 0; i < n; i++)
            if (!pool->provers[i] || states[i] != KZP_STATE_OK)
                pool->state = pool->provers[i] ? states[i] : KZP_STATE_ZKEY_FILE_LOAD_ERROR;
        pool->sched = new SlotScheduler((int)n);
        if (pool->state == KZP_STATE_OK)
        {
            int st = 0;
            kzp_host_parse_zkey(pool->zkey_path.c_str(), nullptr, &pool->n_public, nullptr, nullptr, &st);
        }
    }
    catch (...)
    {
        pool->state = KZP_STATE_ZKEY_FILE_LOAD_ERROR;
    }
    if (state_out)
        *state_out = pool->state;
    return pool;
}

void kzp_pool_free(kzp_pool* pool)
{
    if (!pool)
        return;
    if (pool->sched)
    {
        pool->sched->close();
    }
    // Every caller that entered pool_run — still queueing (woken by close()), proving, or in the verify step, which
    // reads the pool after its slot went back — must have left before anything is destroyed.
    while (pool->in_flight.load() > 0)
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    for (kzp_prover* p : pool->provers)
        kzp_prover_free(p);
    delete pool->sched;
    delete pool;
}

int kzp_pool_size(const kzp_pool* pool) { return pool ? (int)pool->provers.size() : 0; }

int kzp_pool_healthy(const kzp_pool* pool)
{
    if (!pool || pool->state != KZP_STATE_OK)
        return 0;
    int n = 0;
    for (kzp_prover* p : pool->provers)
        n += kzp_prover_state(p) == KZP_STATE_OK ? 1 : 0;
    return n;
}

int kzp_pool_device(const kzp_pool* pool, int slot)
{
    if (!pool || slot < 0 || slot >= (int)pool->devices.size())
        return -1;
    return pool->devices[slot];
}

// Fused verify-before-return: the service re-verifies every proof it hands out
// (prover_handler.rs:329-336); with kzp_pool_set_verify(pool, 1) that check runs here, on the calling thread and
// AFTER the prover has been released, so the GPU already works on the next request while the pairing is checked.
int kzp_pool_set_verify(kzp_pool* pool, int on)
{
    if (!pool)
        return KZP_ERR_STATE;
    pool->verify.store(on ? 1 : 0);
    return KZP_OK;
}

template <class Run, class Publics>
static int pool_run(kzp_pool* pool, char** json_out, int* error_out, int* prover_time_ms, int* slot_out, Run&& run,
                    Publics&& publics)
{
    if (json_out)
        *json_out = nullptr;
    if (prover_time_ms)
        *prover_time_ms = 0;
    if (slot_out)
        *slot_out = -1;
    if (!pool || pool->state != KZP_STATE_OK || !pool->sched)
    {
        if (error_out)
            *error_out = KZP_PROVER_ERROR_NOT_READY;
        return KZP_RESPONSE_ERROR;
    }
    InFlight guard(pool->in_flight);
    int      slot = pool->sched->acquire();
    if (slot < 0)
    {
        if (error_out)
            *error_out = KZP_PROVER_ERROR_NOT_READY;
        return KZP_RESPONSE_ERROR;
    }
    if (slot_out)
        *slot_out = slot;
    kzp_prover* prover = pool->provers[slot];
    int         rc     = run(prover);
    // a prover whose device faulted is taken out of rotation instead of answering every later request with an error
    if (kzp_prover_state(prover) != KZP_STATE_OK)
        pool->sched->retire(slot);
    else
        pool->sched->release(slot);
    if (rc == KZP_RESPONSE_SUCCESS && pool->verify.load() && json_out && *json_out)
    {
        auto drop = [&](int err) {
            kzp_free(*json_out);
            *json_out = nullptr;
            if (error_out)
                *error_out = err;
            return KZP_RESPONSE_ERROR;
        };
        std::vector<uint8_t> pub((size_t)pool->n_public * 32);
        int                  valid = 0;
        if (!publics(pub.data(), pool->n_public))
            return drop(KZP_PROVER_ERROR_INVALID_INPUT); // the witness has no public signals to check against
        if (kzp_host_verify(pool->zkey_path.c_str(), *json_out, pub.data(), pool->n_public, &valid) != KZP_OK)
            return drop(KZP_PROVER_ERROR_NOT_READY); // the verifier could not run (zkey unreadable, malformed proof)
        if (!valid)
        {
            // The prover returned a well-formed proof that does not verify. With a healthy device that is a witness
            // which does not satisfy the circuit (the client's fault); a device that has faulted since then is ours.
            return drop(kzp_prover_state(prover) == KZP_STATE_OK ? KZP_PROVER_ERROR_INVALID_INPUT
                                                                 : KZP_PROVER_ERROR_NOT_READY);
        }
    }
    return rc;
}

int kzp_pool_prove(kzp_pool* pool, const char* wtns_path, const uint8_t* r32, const uint8_t* s32, char** json_out,
                   int* error_out, int* prover_time_ms, int* slot_out)
{
    return pool_run(
        pool, json_out, error_out, prover_time_ms, slot_out,
        [&](kzp_prover* p) { return kzp_prover_prove(p, wtns_path, r32, s32, json_out, error_out, prover_time_ms); },
        [&](uint8_t* out, uint32_t n_public) {
            // public signals = witness values 1 .. nPublic (wtns_utils.hpp:28-43)
            try
            {
                MappedFile file(wtns_path ? wtns_path : "");
                BinView    bin(file.data(), file.size(), "wtns", 2);
                WtnsHeader wh = parse_wtns(bin);
                if (wh.values_bytes < 32ull * (n_public + 1))
                    return false;
                memcpy(out, wh.values + 32, 32ull * n_public);
                return true;
            }
            catch (...)
            {
                return false;
            }
        });
}

int kzp_pool_prove_mem(kzp_pool* pool, const uint8_t* witness, uint64_t n, const uint8_t* r32, const uint8_t* s32,
                       char** json_out, int* error_out, int* prover_time_ms, int* slot_out)
{
    return pool_run(
        pool, json_out, error_out, prover_time_ms, slot_out,
        [&](kzp_prover* p) { return kzp_prover_prove_mem(p, witness, n, r32, s32, json_out, error_out, prover_time_ms); },
        [&](uint8_t* out, uint32_t n_public) {
            if (!witness || n < (uint64_t)n_public + 1)
                return false;
            memcpy(out, witness + 32, 32ull * n_public);
            return true;
        });
}

int kzp_pool_stats(kzp_pool* pool, uint64_t* proofs_per_slot, int cap, uint64_t* max_waiting)
{
    if (!pool || !pool->sched)
        return 0;
    int n = pool->sched->slots();
    for (int i = 0; i < n && i < cap; i++)
        proofs_per_slot[i] = pool->sched->jobs(i);
    if (max_waiting)
        *max_waiting = pool->sched->max_waiting();
    return n < cap ? n : cap;
}

// Host-only exercise of the scheduler (no GPU): `threads` callers each run `jobs_per_thread` jobs that hold a slot
// for hold_us microseconds. Reports per-slot job counts, the largest number of simultaneous holders seen on any one
// slot (must be 1) and the deepest queue (callers waiting or being served) seen.
int kzp_pool_sched_selftest(int slots, int threads, int jobs_per_thread, int hold_us, uint64_t* per_slot_out,
                            int* max_concurrent_per_slot, uint64_t* max_waiting)
{
    if (slots <= 0 || threads <= 0 || jobs_per_thread < 0 || !per_slot_out)
        return KZP_ERR_FORMAT;
    SlotScheduler                 sched(slots);
    std::vector<std::atomic<int>> holders(slots);
    for (auto& h : holders)
        h = 0;
    std::atomic<int>         worst(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&] {
            for (int j = 0; j < jobs_per_thread; j++)
            {
                int slot = sched.acquire();
                if (slot < 0)
                    return;
                int now = ++holders[slot];
                int w   = worst.load();
                while (now > w && !worst.compare_exchange_weak(w, now))
                {
                }
                if (hold_us > 0)
                    std::this_thread::sleep_for(std::chrono::microseconds(hold_us));
                --holders[slot];
                sched.release(slot);
            }
        });
    for (auto& t : th)
        t.join();
    for (int i = 0; i < slots; i++)
        per_slot_out[i] = sched.jobs(i);
    if (max_concurrent_per_slot)
        *max_concurrent_per_slot = worst.load();
    if (max_waiting)
        *max_waiting = sched.max_waiting();
    return KZP_OK;
}
