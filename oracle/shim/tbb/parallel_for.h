/* TEST INFRASTRUCTURE — oracle build shim, never part of the product library.
 *
 * Minimal stand-in for the slice of oneTBB the reference prover uses
 * (rust-rapidsnark/rapidsnark/src/{groth16,fft,multiexp}.cpp): blocked_range,
 * the two parallel_for overloads, and this_task_arena::{max_concurrency,
 * current_thread_index}. Work is split into sub-ranges executed by an OpenMP
 * team; the thread index is the OpenMP thread number inside that team, which is
 * what the reference needs for its per-thread MSM accumulators
 * (multiexp.cpp:61, :150, :188).
 */
#ifndef KZP_ORACLE_TBB_SHIM_H
#define KZP_ORACLE_TBB_SHIM_H

#include <omp.h>

#include <cstddef>
#include <cstdint>

namespace tbb
{

template <typename T>
class blocked_range
{
    T b_;
    T e_;

public:
    blocked_range(T b, T e)
        : b_(b)
        , e_(e)
    {
    }
    T begin() const { return b_; }
    T end() const { return e_; }
    bool empty() const { return !(b_ < e_); }
};

namespace this_task_arena
{
inline int max_concurrency() { return omp_get_max_threads(); }
inline int current_thread_index() { return omp_get_thread_num(); }
} // namespace this_task_arena

template <typename T, typename F>
void parallel_for(blocked_range<T> const& r, F const& f)
{
    if (r.empty())
        return;
    std::uint64_t const total = (std::uint64_t)(r.end() - r.begin());
    std::uint64_t       parts = (std::uint64_t)omp_get_max_threads() * 8;
    if (parts > total)
        parts = total;
    std::uint64_t const step = (total + parts - 1) / parts;
#pragma omp parallel for schedule(dynamic, 1)
    for (std::uint64_t p = 0; p < parts; p++)
    {
        std::uint64_t lo = p * step;
        std::uint64_t hi = lo + step;
        if (lo >= total)
            continue;
        if (hi > total)
            hi = total;
        f(blocked_range<T>((T)(r.begin() + lo), (T)(r.begin() + hi)));
    }
}

template <typename F>
void parallel_for(int first, int last, F const& f)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = first; i < last; i++)
    {
        f(i);
    }
}

} // namespace tbb

#endif
