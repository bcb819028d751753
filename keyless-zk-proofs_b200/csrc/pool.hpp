// GPU-per-request scheduling for a set of resident provers.
//
// The reference service holds ONE FullProver behind Arc<tokio::Mutex<Option<_>>> (prover-service/src/
// prover_state.rs:21,38-47,101) and calls prove() on the async worker itself (prover_handler.rs:266-283), so proofs
// are strictly serialised. With one resident prover per GPU (or several per GPU, so that one proof's host-side
// staging and assembly overlap the other's kernels) a request only needs *a* free prover: this header is the
// checkout — a FIFO ticket queue over a small set of slots, least-recently-released slot first so that work spreads
// over the GPUs. Host-only C++ (no CUDA types): unit-tested without a device through kzp_pool_sched_selftest.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <vector>

namespace kzp
{

class SlotScheduler
{
  public:
    explicit SlotScheduler(int slots) : jobs_(slots > 0 ? slots : 0, 0), alive_(slots > 0 ? slots : 0)
    {
        for (int i = 0; i < slots; i++)
            free_.push_back(i);
    }

    // Blocks until this caller is at the head of the queue and a slot is free. Returns the slot, or -1 after close().
    int acquire()
    {
        std::unique_lock<std::mutex> lk(m_);
        const uint64_t               ticket = next_ticket_++;
        uint64_t                     depth  = next_ticket_ - serving_;
        if (depth > max_waiting_)
            max_waiting_ = depth;
        cv_.wait(lk, [&] { return closed_ || alive_ == 0 || (ticket == serving_ && !free_.empty()); });
        if (closed_ || alive_ == 0)
        {
            // keep the queue moving for the waiters behind this one
            if (ticket == serving_)
                serving_++;
            cv_.notify_all();
            return -1;
        }
        serving_++;
        int slot = free_.front();
        free_.pop_front();
        jobs_[slot]++;
        cv_.notify_all(); // the next ticket may proceed if another slot is free
        return slot;
    }

    void release(int slot)
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            free_.push_back(slot);
        }
        cv_.notify_all();
    }

    // The slot's prover is out of service (device fault): it is never handed out again. When the last slot is retired
    // every waiter and every later acquire() gets -1.
    void retire(int slot)
    {
        (void)slot;
        {
            std::lock_guard<std::mutex> lk(m_);
            alive_--;
        }
        cv_.notify_all();
    }

    // Wakes every waiter with -1 and makes later acquire() calls fail; slots already handed out stay valid.
    void close()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            closed_ = true;
        }
        cv_.notify_all();
    }

    int      slots() const { return (int)jobs_.size(); }
    uint64_t jobs(int slot)
    {
        std::lock_guard<std::mutex> lk(m_);
        return jobs_[slot];
    }
    uint64_t max_waiting()
    {
        std::lock_guard<std::mutex> lk(m_);
        return max_waiting_;
    }
    // number of slots currently handed out (retired slots are neither free nor busy)
    int busy()
    {
        std::lock_guard<std::mutex> lk(m_);
        return alive_ - (int)free_.size();
    }
    int alive()
    {
        std::lock_guard<std::mutex> lk(m_);
        return alive_;
    }

  private:
    std::mutex              m_;
    std::condition_variable cv_;
    std::deque<int>         free_;
    std::vector<uint64_t>   jobs_;
    int                     alive_       = 0; // slots not retired
    uint64_t                next_ticket_ = 0, serving_ = 0, max_waiting_ = 0;
    bool                    closed_      = false;
};

} // namespace kzp
