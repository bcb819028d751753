// Digit sort of a scalar vector for the Pippenger MSMs (sm_100a): signed base-2^16 digits -> histogram -> exclusive
// scan -> scatter of (window, base) entries into per-bucket ranges. The histogram pass uses no global atomics: one
// CTA per SM owns a contiguous slice of the scalars and keeps all 2^15 bucket counters in its shared memory (128 KiB
// of the 227 KiB); the per-CTA histograms are summed per bucket afterwards (H MSM, 33.5 M entries: 0.39 ms with
// global atomics -> 0.03 + 0.03 ms). Group independent, so the four MSMs over the
// witness (A, B1, C in G1 and B2 in G2; groth16.cpp:88-112) sort once. Replaces the per-chunk digit extraction of
// ParallelMultiexp::processChunk / getChunk (rust-rapidsnark/rapidsnark/src/multiexp.cpp:26-71).
#include <algorithm>

#include "device.hpp"

namespace kzp
{

static inline unsigned int sort_div_up(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

// Loads a 32-byte little-endian integer and brings it below r by repeated subtraction (at most 5 times
// for any 256-bit value). For points of order r this matches the reference, which uses the raw bits.
__device__ __forceinline__ void load_scalar(const uint32_t* __restrict__ scalars, uint32_t idx, Fr& s)
{
    const uint4* p  = reinterpret_cast<const uint4*>(scalars + (size_t)idx * 8);
    uint4        lo = p[0], hi = p[1];
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
    s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
    while (s.v[7] >= FrParams::P7 && Fr::geq_p(s))
    {
        uint32_t bw = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            uint32_t pi = modulus_limb<FrParams>(i);
            uint64_t d  = (uint64_t)s.v[i] - pi - bw;
            s.v[i]      = (uint32_t)d;
            bw          = (uint32_t)(d >> 63);
        }
    }
}

// signed base-2^C digits d_j in [-2^(C-1), 2^(C-1)], sum d_j 2^(C j) = s, for s < 2^255 (j must be a compile-time
// constant after unrolling: the limb index is then static and the scalar stays in registers)
template <int C>
__device__ __forceinline__ int32_t next_digit(const Fr& s, int j, uint32_t& carry)
{
    constexpr uint32_t mask = (1u << C) - 1u, half = 1u << (C - 1);
    const int          o = C * j, limb = o >> 5, sh = o & 31;
    uint32_t           lo = limb < 8 ? s.v[limb] : 0u;
    uint32_t           hi = limb + 1 < 8 ? s.v[limb + 1] : 0u;
    uint32_t           raw = (sh ? __funnelshift_r(lo, hi, sh) : lo) & mask;
    uint32_t           v   = raw + carry;
    if (v > half)
    {
        carry = 1;
        return (int32_t)v - (int32_t)(1u << C);
    }
    carry = 0;
    return (int32_t)v;
}

template <int C>
struct SortCfg
{
    static constexpr int      W = (255 + C - 1) / C;
    static constexpr uint32_t B = 1u << (C - 1);
};

// ---- one-level sort (c = 16) ------------------------------------------------------------------------------
constexpr int      kMsmWindows = SortCfg<16>::W;
constexpr uint32_t kMsmBuckets = SortCfg<16>::B;

__device__ __forceinline__ bool scalar_is_small(const Fr& s)
{
    return (s.v[1] | s.v[2] | s.v[3] | s.v[4] | s.v[5] | s.v[6] | s.v[7]) == 0 && s.v[0] <= 0x8000u;
}

constexpr int    kSortThreads = 1024;
constexpr size_t kSortSmem    = (size_t)(kMsmBuckets + 1) * sizeof(uint32_t);

// Pass 1: per-CTA histogram of bucket ids in shared memory. Small scalars (one digit; the bulk of a circom witness:
// bits, bytes) are warp-aggregated so that a million equal digits do not serialise on one counter.
static __global__ void __launch_bounds__(kSortThreads, 1)
    k_msm_hist(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t scalar_offset,
               uint32_t n, uint32_t per_cta, uint32_t* __restrict__ cta_hist)
{
    extern __shared__ uint32_t sm_cnt[];
    for (uint32_t b = threadIdx.x; b <= kMsmBuckets; b += kSortThreads)
        sm_cnt[b] = 0;
    __syncthreads();
    const uint32_t lo   = blockIdx.x * per_cta;
    const uint32_t hi   = min(n, lo + per_cta);
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = lo; base < hi; base += kSortThreads) // trip count is uniform over the CTA
    {
        uint32_t i     = base + threadIdx.x;
        bool     valid = i < hi;
        Fr       s     = Fr::zero();
        if (valid)
            load_scalar(scalars, scalar_idx ? scalar_idx[i] : scalar_offset + i, s);
        bool     small = valid && scalar_is_small(s);
        uint32_t key   = (small && s.v[0] != 0) ? s.v[0] : (0xffff0000u | lane);
        uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (small)
        {
            if (s.v[0] != 0 && lane == (uint32_t)(__ffs(peers) - 1))
                atomicAdd(&sm_cnt[s.v[0]], (uint32_t)__popc(peers));
        }
        else if (valid)
        {
            uint32_t carry = 0;
#pragma unroll
            for (int j = 0; j < kMsmWindows; j++)
            {
                int32_t d = next_digit<16>(s, j, carry);
                if (d != 0)
                    atomicAdd(&sm_cnt[d < 0 ? -d : d], 1u);
            }
        }
    }
    __syncthreads();
    uint32_t* out = cta_hist + (size_t)blockIdx.x * (kMsmBuckets + 1);
    for (uint32_t b = threadIdx.x; b <= kMsmBuckets; b += kSortThreads)
        out[b] = sm_cnt[b];
}

// Per bucket: the total over the CTA histograms.
static __global__ void __launch_bounds__(256)
    k_msm_colsum(const uint32_t* __restrict__ cta_hist, uint32_t ctas, uint32_t* __restrict__ counts)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > kMsmBuckets)
        return;
    uint32_t run = 0;
#pragma unroll 4
    for (uint32_t c = 0; c < ctas; c++)
        run += cta_hist[(size_t)c * (kMsmBuckets + 1) + b];
    counts[b] = run;
}

// Exclusive scan of counts[0..B] -> offsets[b] (start of bucket b; counts[0] == 0), offsets[B+1] = total; cursor =
// offsets. One CTA; every thread owns 32 consecutive buckets starting at a 128-byte boundary, read with 128-bit loads.
static __global__ void __launch_bounds__(1024)
    k_msm_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor)
{
    static_assert(kMsmBuckets % 1024 == 0 && (kMsmBuckets / 1024) % 4 == 0, "bucket count must split into uint4 runs");
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    constexpr uint32_t  per = kMsmBuckets / 1024;
    uint32_t            tid = threadIdx.x;
    uint32_t            b0  = tid * per;
    uint32_t            loc[per];
    uint32_t            sum = 0;
    const uint4*        src = reinterpret_cast<const uint4*>(counts + b0);
#pragma unroll
    for (uint32_t k = 0; k < per / 4; k++)
    {
        uint4 c        = src[k];
        loc[4 * k]     = sum; sum += c.x;
        loc[4 * k + 1] = sum; sum += c.y;
        loc[4 * k + 2] = sum; sum += c.z;
        loc[4 * k + 3] = sum; sum += c.w;
    }
    uint32_t lane = tid & 31, wid = tid >> 5;
    uint32_t inc  = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o)
            inc += t;
    }
    if (lane == 31)
        warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0)
    {
        uint32_t ws = warp_sums[lane];
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o)
                wi += t;
        }
        warp_sums[lane] = wi - ws; // exclusive
        if (lane == 31)
            carry_s = wi;
    }
    __syncthreads();
    uint32_t base = warp_sums[wid] + inc - sum;
    uint4*   o4   = reinterpret_cast<uint4*>(offsets + b0);
    uint4*   c4   = reinterpret_cast<uint4*>(cursor + b0);
#pragma unroll
    for (uint32_t k = 0; k < per / 4; k++)
    {
        uint4 v = make_uint4(base + loc[4 * k], base + loc[4 * k + 1], base + loc[4 * k + 2], base + loc[4 * k + 3]);
        o4[k]   = v;
        c4[k]   = v;
    }
    if (tid == 0)
    {
        // the last bucket (index B) and the grand total
        uint32_t before_last     = carry_s;
        offsets[kMsmBuckets]     = before_last;
        cursor[kMsmBuckets]      = before_last;
        offsets[kMsmBuckets + 1] = before_last + counts[kMsmBuckets];
    }
}

// Pass 2: scatter entry = base | window << 27 | sign << 31 into its bucket's range. Positions come from global
// cursors on purpose: all CTAs then fill every bucket range front to back, so the write frontier is one sector per
// bucket (1 MB, L2 resident) and finished sectors leave L2 fully written. Handing out positions from per-CTA
// shared-memory cursors instead (tried: 0.71 ms vs 0.47 ms) gives 148 x 2^15 frontiers = 154 MB of partially written
// sectors, more than the L2 holds.
static __global__ void __launch_bounds__(256)
    k_msm_scatter(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t scalar_offset,
                  uint32_t n, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted)
{
    uint32_t i     = blockIdx.x * blockDim.x + threadIdx.x;
    bool     valid = i < n;
    Fr       s     = Fr::zero();
    if (valid)
        load_scalar(scalars, scalar_idx ? scalar_idx[i] : scalar_offset + i, s);
    bool     small  = valid && scalar_is_small(s);
    uint32_t lane   = threadIdx.x & 31;
    uint32_t key    = (small && s.v[0] != 0) ? s.v[0] : (0xffff0000u | lane);
    uint32_t peers  = __match_any_sync(0xffffffffu, key);
    uint32_t leader = (uint32_t)(__ffs(peers) - 1);
    uint32_t base   = 0;
    if (small && s.v[0] != 0 && lane == leader)
        base = atomicAdd(&cursor[s.v[0]], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (small)
    {
        if (s.v[0] != 0)
            sorted[base + __popc(peers & ((1u << lane) - 1))] = i; // window 0, positive
        return;
    }
    if (!valid)
        return;
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < kMsmWindows; j++)
    {
        int32_t d = next_digit<16>(s, j, carry);
        if (d != 0)
        {
            uint32_t bkt = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            uint32_t pos = atomicAdd(&cursor[bkt], 1u);
            sorted[pos]  = i | ((uint32_t)j << kMsmEntryBaseBits) | (d < 0 ? 0x80000000u : 0u);
        }
    }
}


// ---- two-level sort (any window size) ---------------------------------------------------------------------
// With 2^19 buckets (c = 20) neither the per-CTA histogram (2 MB of counters) nor one global cursor per bucket
// (every entry a lone atomic + a lone 4-byte store: the one-level scatter is L2-request bound) works. Two passes,
// every global write part of a contiguous run:
//   count     per-partition totals (partition = high bits of the bucket index), shared-memory counters per tile
//   scan      partition offsets
//   partition one CTA per tile of scalars: digits -> counting sort by partition in shared memory -> one global
//             atomic per (tile, partition) reserves a run -> 8-byte records (entry, bucket index) copied out in
//             partition order, consecutive threads on consecutive addresses
//   local     one CTA per partition: counting sort by the low bucket bits, staged in shared memory and written
//             as one contiguous block; the same CTA writes the bucket offsets of its partition
constexpr int    kS2Threads   = 1024;
constexpr size_t kS2SmemMax   = 227 * 1024;
constexpr int    kS2MaxParts  = 1024;
constexpr int    kS2MaxSubBits = 12;
constexpr uint32_t kS2LocalCap = 48 * 1024; // entries one partition may have to take the staged path

template <int C, class Fn>
__device__ __forceinline__ void for_each_digit(const Fr& s, uint32_t i, Fn&& fn)
{
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < SortCfg<C>::W; j++)
    {
        int32_t d = next_digit<C>(s, j, carry);
        if (d != 0)
        {
            uint32_t bkt = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            fn(bkt - 1u, i | ((uint32_t)j << kMsmEntryBaseBits) | (d < 0 ? 0x80000000u : 0u));
        }
    }
}

template <int C>
static __global__ void __launch_bounds__(kS2Threads, 1)
    k_s2_count(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t scalar_offset,
               uint32_t n, uint32_t tile, uint32_t sub_bits, uint32_t parts, uint32_t* __restrict__ part_cnt)
{
    __shared__ uint32_t cnt[kS2MaxParts];
    for (uint32_t p = threadIdx.x; p < parts; p += kS2Threads)
        cnt[p] = 0;
    __syncthreads();
    const uint32_t lo = blockIdx.x * tile, hi = min(n, lo + tile);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += kS2Threads)
    {
        Fr s;
        load_scalar(scalars, scalar_idx ? scalar_idx[i] : scalar_offset + i, s);
        for_each_digit<C>(s, i, [&](uint32_t b, uint32_t) { atomicAdd(&cnt[b >> sub_bits], 1u); });
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < parts; p += kS2Threads)
        if (cnt[p])
            atomicAdd(&part_cnt[p], cnt[p]);
}

// part_cnt[0..P) -> exclusive offsets in place, part_cnt[P] = total; cursors = offsets; offsets[0] = 0 and
// offsets[B + 1] = total (the per-bucket offsets in between are written by the local pass)
static __global__ void __launch_bounds__(kS2MaxParts)
    k_s2_scan(uint32_t* __restrict__ part_cnt, uint32_t* __restrict__ part_cursor, uint32_t parts,
              uint32_t* __restrict__ offsets, uint32_t buckets)
{
    __shared__ uint32_t warp_sums[32];
    uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t v   = tid < parts ? part_cnt[tid] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o)
            inc += t;
    }
    if (lane == 31)
        warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0)
    {
        uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o)
                wi += t;
        }
        warp_sums[lane] = wi - ws;
    }
    __syncthreads();
    uint32_t excl = warp_sums[wid] + inc - v;
    if (tid < parts)
    {
        part_cnt[tid]    = excl;
        part_cursor[tid] = excl;
    }
    if (tid == parts - 1)
    {
        part_cnt[parts]      = excl + v;
        offsets[0]           = 0;
        offsets[buckets + 1] = excl + v;
    }
}

template <int C>
static __global__ void __launch_bounds__(kS2Threads, 1)
    k_s2_partition(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t scalar_offset,
                   uint32_t n, uint32_t tile, uint32_t sub_bits, uint32_t parts, uint32_t* __restrict__ part_cursor,
                   uint2* __restrict__ inter)
{
    extern __shared__ uint4 s2_smem[];
    uint32_t* cnt   = reinterpret_cast<uint32_t*>(s2_smem); // counts, then write cursors
    uint32_t* start = cnt + parts;                          // first staging slot of each partition
    uint32_t* gbase = start + parts;                        // first global slot of this tile's run
    uint2*    stage = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(s2_smem) + ((3 * parts * 4 + 15u) & ~15u));
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t p = tid; p < parts; p += kS2Threads)
        cnt[p] = 0;
    __syncthreads();
    const uint32_t lo = blockIdx.x * tile, hi = min(n, lo + tile);
    for (uint32_t i = lo + tid; i < hi; i += kS2Threads)
    {
        Fr s;
        load_scalar(scalars, scalar_idx ? scalar_idx[i] : scalar_offset + i, s);
        for_each_digit<C>(s, i, [&](uint32_t b, uint32_t) { atomicAdd(&cnt[b >> sub_bits], 1u); });
    }
    __syncthreads();
    // exclusive scan of the (<= 1024) partition counts, one per thread; reserve the global runs
    {
        uint32_t v = tid < parts ? cnt[tid] : 0u, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o)
                inc += t;
        }
        if (lane == 31)
            warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0)
        {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o)
                    wi += t;
            }
            warp_sums[lane] = wi - ws;
            if (lane == 31)
                s_total = wi;
        }
        __syncthreads();
        if (tid < parts)
        {
            uint32_t excl = warp_sums[wid] + inc - v;
            start[tid]    = excl;
            cnt[tid]      = excl; // becomes the write cursor
            gbase[tid]    = v ? atomicAdd(&part_cursor[tid], v) : 0u;
        }
    }
    __syncthreads();
    for (uint32_t i = lo + tid; i < hi; i += kS2Threads)
    {
        Fr s;
        load_scalar(scalars, scalar_idx ? scalar_idx[i] : scalar_offset + i, s);
        for_each_digit<C>(s, i, [&](uint32_t b, uint32_t e) {
            uint32_t pos = atomicAdd(&cnt[b >> sub_bits], 1u);
            stage[pos]   = make_uint2(e, b);
        });
    }
    __syncthreads();
    const uint32_t total = s_total;
    for (uint32_t k = tid; k < total; k += kS2Threads)
    {
        uint2    rec = stage[k];
        uint32_t p   = rec.y >> sub_bits;
        inter[gbase[p] + (k - start[p])] = rec;
    }
}

static __global__ void __launch_bounds__(kS2Threads, 1)
    k_s2_local(const uint2* __restrict__ inter, const uint32_t* __restrict__ part_off, uint32_t sub_bits,
               uint32_t* __restrict__ offsets, uint32_t* __restrict__ sorted)
{
    extern __shared__ uint4 s2_smem[];
    const uint32_t subs = 1u << sub_bits;
    uint32_t*      cnt  = reinterpret_cast<uint32_t*>(s2_smem); // counts, then cursors
    uint32_t*      stage = cnt + subs;
    __shared__ uint32_t warp_sums[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t p = blockIdx.x, base = part_off[p], n_p = part_off[p + 1] - base;
    const uint32_t sub_mask = subs - 1u;
    const uint2*   in = inter + base;
    for (uint32_t k = tid; k < subs; k += kS2Threads)
        cnt[k] = 0;
    __syncthreads();
    for (uint32_t k = tid; k < n_p; k += kS2Threads)
        atomicAdd(&cnt[in[k].y & sub_mask], 1u);
    __syncthreads();
    // exclusive scan over the sub-buckets: every thread owns `per` consecutive counters
    {
        const uint32_t per = (subs + kS2Threads - 1) / kS2Threads; // 1..4
        uint32_t       loc[1u << (kS2MaxSubBits - 10)];
        uint32_t       sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < (1u << (kS2MaxSubBits - 10)); q++)
            if (q < per)
            {
                uint32_t idx = tid * per + q;
                loc[q]       = sum;
                sum += idx < subs ? cnt[idx] : 0u;
            }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o)
                inc += t;
        }
        if (lane == 31)
            warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0)
        {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o)
                    wi += t;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        uint32_t excl = warp_sums[wid] + inc - sum;
#pragma unroll
        for (uint32_t q = 0; q < (1u << (kS2MaxSubBits - 10)); q++)
            if (q < per)
            {
                uint32_t idx = tid * per + q;
                if (idx < subs)
                {
                    uint32_t o = excl + loc[q];
                    cnt[idx]   = o;
                    offsets[1u + (p << sub_bits) + idx] = base + o; // bucket id = index + 1
                }
            }
    }
    __syncthreads();
    if (n_p <= kS2LocalCap)
    {
        for (uint32_t k = tid; k < n_p; k += kS2Threads)
        {
            uint2    rec = in[k];
            uint32_t pos = atomicAdd(&cnt[rec.y & sub_mask], 1u);
            stage[pos]   = rec.x;
        }
        __syncthreads();
        for (uint32_t k = tid; k < n_p; k += kS2Threads)
            sorted[base + k] = stage[k];
    }
    else
    {
        // oversized partition (skewed digits): positions still come from the shared-memory cursors, the stores go
        // straight to global memory
        for (uint32_t k = tid; k < n_p; k += kS2Threads)
        {
            uint2    rec = in[k];
            uint32_t pos = atomicAdd(&cnt[rec.y & sub_mask], 1u);
            sorted[base + pos] = rec.x;
        }
    }
}

template <int C>
static void s2_run(MsmSort& s, const uint32_t* scalars, cudaStream_t st)
{
    const uint32_t parts = 1u << s.part_bits, sub_bits = (uint32_t)(C - 1) - s.part_bits;
    const uint32_t W = SortCfg<C>::W;
    // tile of scalars per CTA: the staging area takes 8 bytes per possible entry
    size_t   head = ((size_t)3 * parts * 4 + 15) & ~(size_t)15;
    uint32_t tile = (uint32_t)((kS2SmemMax - 1024 - head) / ((size_t)W * 8));
    tile          = std::min<uint32_t>(tile / kS2Threads * kS2Threads, 4 * kS2Threads);
    size_t   smem_a = head + (size_t)tile * W * 8;
    size_t   smem_b = ((size_t)1 << sub_bits) * 4 + (size_t)kS2LocalCap * 4;
    uint32_t tiles  = std::max<uint32_t>(1, sort_div_up(s.n, tile));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_s2_partition<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_s2_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    KZP_CUDA_CHECK(cudaMemsetAsync(s.part_cnt, 0, ((size_t)parts + 1) * 4, st));
    k_s2_count<C><<<tiles, kS2Threads, 0, st>>>(scalars, s.scalar_idx, s.scalar_offset, s.n, tile, sub_bits, parts, s.part_cnt);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_s2_scan<<<1, kS2MaxParts, 0, st>>>(s.part_cnt, s.part_cursor, parts, s.offsets, SortCfg<C>::B);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_s2_partition<C><<<tiles, kS2Threads, smem_a, st>>>(scalars, s.scalar_idx, s.scalar_offset, s.n, tile, sub_bits, parts,
                                                          s.part_cursor, s.inter);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_s2_local<<<parts, kS2Threads, smem_b, st>>>(s.inter, s.part_cnt, sub_bits, s.offsets, s.sorted);
    KZP_CUDA_CHECK(cudaGetLastError());
}

uint32_t msm_default_chunk(uint64_t n)
{
    // Sorted entries per accumulate thread when the caller does not say (the component-level MSM entry points; the
    // prover sets its own). Measured with the cooperative bucket reduction (2^22 points, chunk 32 / 64 / 128 / 256):
    // uniform scalars 11.6 / 11.5 / 11.4 / 11.6 ms, keyless-like scalars 1.6 / 1.6 / 2.2 / 3.8 ms — long chunks starve
    // a bit-heavy vector of threads, short ones cost a uniform vector nothing.
    return n > (1u << 19) ? 64 : 32;
}

void msm_sort_create(MsmSort& s, uint32_t n, const uint32_t* scalar_idx, uint32_t scalar_offset, uint32_t window_bits,
                     bool force_two_level)
{
    if (n > kMsmEntryBaseMask)
        throw CudaError("MSM too large for 27-bit base ids");
    if (window_bits < kMsmMinWindowBits || window_bits > kMsmMaxWindowBits)
        throw CudaError("MSM window size out of range");
    s.shape         = msm_shape(window_bits);
    s.n             = n;
    s.scalar_idx    = scalar_idx;
    s.scalar_offset = scalar_offset;
    uint64_t cap    = (uint64_t)n * s.shape.windows;
    if (cap > 0xffffffffull)
        throw CudaError("MSM too large: more than 2^32 digit entries");
    s.cap_entries   = (uint32_t)cap;
    s.two_level     = force_two_level || window_bits != 16;
    size_t nb       = (size_t)s.shape.buckets + 2;
    KZP_CUDA_CHECK(cudaMalloc(&s.offsets, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.sorted, std::max<uint64_t>(cap, 1) * 4));
    if (s.two_level)
    {
        // partitions: enough that an average partition fits the local pass's staging area with room to spare, at
        // most 1024 (one scan thread each), at least so many that the sub-bucket counters fit shared memory
        uint32_t want = 0;
        while (want < 10 && (cap >> want) > 32 * 1024)
            want++;
        uint32_t min_bits = window_bits - 1 > (uint32_t)kS2MaxSubBits ? window_bits - 1 - kS2MaxSubBits : 0;
        s.part_bits       = std::min<uint32_t>(std::max(want, min_bits), window_bits - 1);
        if (s.part_bits > 10)
            throw CudaError("MSM window too large for the two-level sort");
        size_t parts = (size_t)1 << s.part_bits;
        KZP_CUDA_CHECK(cudaMalloc(&s.part_cnt, (parts + 1) * 4));
        KZP_CUDA_CHECK(cudaMalloc(&s.part_cursor, parts * 4));
        KZP_CUDA_CHECK(cudaMalloc(&s.inter, std::max<uint64_t>(cap, 1) * 8));
        return;
    }
    KZP_CUDA_CHECK(cudaMalloc(&s.counts, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.cursor, nb * 4));
    // one CTA per SM, each owning a contiguous slice of the scalars (a multiple of the block size)
    int dev = 0, sms = 1;
    KZP_CUDA_CHECK(cudaGetDevice(&dev));
    KZP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t blocks = std::max<uint32_t>(1, sort_div_up(n, kSortThreads));
    s.sort_ctas     = std::min<uint32_t>((uint32_t)sms, blocks);
    s.per_cta       = sort_div_up(blocks, s.sort_ctas) * kSortThreads;
    s.sort_ctas     = std::max<uint32_t>(1, sort_div_up(n, s.per_cta));
    KZP_CUDA_CHECK(cudaMalloc(&s.cta_hist, (size_t)s.sort_ctas * (kMsmBuckets + 1) * 4));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem));
}

void msm_sort_destroy(MsmSort& s)
{
    cudaFree(s.counts);
    cudaFree(s.offsets);
    cudaFree(s.cursor);
    cudaFree(s.sorted);
    cudaFree(s.cta_hist);
    cudaFree(s.part_cnt);
    cudaFree(s.part_cursor);
    cudaFree(s.inter);
    s = MsmSort();
}

void msm_sort_run(MsmSort& s, const uint32_t* scalars, cudaStream_t st)
{
    if (s.two_level)
    {
        switch (s.shape.c)
        {
        case 16: s2_run<16>(s, scalars, st); break;
        case 17: s2_run<17>(s, scalars, st); break;
        case 18: s2_run<18>(s, scalars, st); break;
        case 19: s2_run<19>(s, scalars, st); break;
        case 20: s2_run<20>(s, scalars, st); break;
        case 21: s2_run<21>(s, scalars, st); break;
        case 22: s2_run<22>(s, scalars, st); break;
        default: throw CudaError("MSM window size out of range");
        }
        return;
    }
    k_msm_hist<<<s.sort_ctas, kSortThreads, kSortSmem, st>>>(scalars, s.scalar_idx, s.scalar_offset, s.n, s.per_cta,
                                                             s.cta_hist);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_msm_colsum<<<sort_div_up(kMsmBuckets + 1, 256), 256, 0, st>>>(s.cta_hist, s.sort_ctas, s.counts);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_msm_scan<<<1, 1024, 0, st>>>(s.counts, s.offsets, s.cursor);
    KZP_CUDA_CHECK(cudaGetLastError());
    if (s.n > 0)
    {
        k_msm_scatter<<<sort_div_up(s.n, 256), 256, 0, st>>>(scalars, s.scalar_idx, s.scalar_offset, s.n, s.cursor,
                                                              s.sorted);
        KZP_CUDA_CHECK(cudaGetLastError());
    }
}

} // namespace kzp
