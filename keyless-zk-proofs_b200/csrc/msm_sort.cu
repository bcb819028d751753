// Digit sort of a scalar vector for the Pippenger MSMs (sm_100a): signed base-2^16 digits -> histogram -> exclusive
// scan -> scatter of (window, base) entries into per-bucket ranges. Group independent, so the four MSMs over the
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

// signed base-2^16 digits d_j in [-2^15, 2^15], sum d_j 2^(16 j) = s, for s < 2^255
__device__ __forceinline__ int32_t next_digit(const Fr& s, int j, uint32_t& carry)
{
    uint32_t raw = (s.v[j >> 1] >> (16 * (j & 1))) & 0xffffu;
    uint32_t v   = raw + carry;
    if (v > 0x8000u)
    {
        carry = 1;
        return (int32_t)v - 0x10000;
    }
    carry = 0;
    return (int32_t)v;
}

__device__ __forceinline__ bool scalar_is_small(const Fr& s)
{
    return (s.v[1] | s.v[2] | s.v[3] | s.v[4] | s.v[5] | s.v[6] | s.v[7]) == 0 && s.v[0] <= 0x8000u;
}

// Pass 1: histogram of bucket ids. Small scalars (one digit; the bulk of a circom witness: bits, bytes)
// are warp-aggregated so that a million equal digits do not serialise on one counter.
static __global__ void __launch_bounds__(256)
    k_msm_count(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t scalar_offset,
                uint32_t n, uint32_t* __restrict__ counts)
{
    uint32_t i     = blockIdx.x * blockDim.x + threadIdx.x;
    bool     valid = i < n;
    Fr       s     = Fr::zero();
    if (valid)
        load_scalar(scalars, scalar_idx ? scalar_idx[i] : scalar_offset + i, s);
    bool     small = valid && scalar_is_small(s);
    uint32_t lane  = threadIdx.x & 31;
    uint32_t key   = (small && s.v[0] != 0) ? s.v[0] : (0xffff0000u | lane);
    uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (small)
    {
        if (s.v[0] != 0 && lane == (uint32_t)(__ffs(peers) - 1))
            atomicAdd(&counts[s.v[0]], (uint32_t)__popc(peers));
        return;
    }
    if (!valid)
        return;
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < kMsmWindows; j++)
    {
        int32_t d = next_digit(s, j, carry);
        if (d != 0)
            atomicAdd(&counts[d < 0 ? -d : d], 1u);
    }
}

// Exclusive scan of counts[1..B] -> offsets[b] (start of bucket b), offsets[B+1] = total; cursor = offsets.
static __global__ void __launch_bounds__(1024)
    k_msm_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    constexpr uint32_t  per = (kMsmBuckets + 1023) / 1024;
    uint32_t            tid = threadIdx.x;
    uint32_t            b0  = 1 + tid * per;
    uint32_t            loc[per];
    uint32_t            sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < per; k++)
    {
        uint32_t b = b0 + k;
        uint32_t c = (b <= kMsmBuckets) ? counts[b] : 0;
        loc[k]     = sum;
        sum += c;
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
#pragma unroll
    for (uint32_t k = 0; k < per; k++)
    {
        uint32_t b = b0 + k;
        if (b <= kMsmBuckets)
        {
            offsets[b] = base + loc[k];
            cursor[b]  = base + loc[k];
        }
    }
    if (tid == 0)
    {
        offsets[0]               = 0;
        offsets[kMsmBuckets + 1] = carry_s;
    }
}

// Pass 2: scatter entry = base | window << 27 | sign << 31 into its bucket's range.
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
        int32_t d = next_digit(s, j, carry);
        if (d != 0)
        {
            uint32_t bkt = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            uint32_t pos = atomicAdd(&cursor[bkt], 1u);
            sorted[pos]  = i | ((uint32_t)j << kMsmEntryBaseBits) | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

uint32_t msm_default_chunk(uint64_t n)
{
    // keep the average bucket at <= ~16 partial sums: chunk >= 16 n / (2^15 * 16)
    uint32_t c = 32;
    while ((uint64_t)c * 16384 < n && c < 4096)
        c <<= 1;
    return c;
}

void msm_sort_create(MsmSort& s, uint32_t n, const uint32_t* scalar_idx, uint32_t scalar_offset)
{
    if (n > kMsmEntryBaseMask)
        throw CudaError("MSM too large for 27-bit base ids");
    s.n             = n;
    s.scalar_idx    = scalar_idx;
    s.scalar_offset = scalar_offset;
    uint64_t cap    = (uint64_t)n * kMsmWindows;
    s.cap_entries   = (uint32_t)cap;
    size_t nb       = kMsmBuckets + 2;
    KZP_CUDA_CHECK(cudaMalloc(&s.counts, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.offsets, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.cursor, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.sorted, std::max<uint64_t>(cap, 1) * 4));
}

void msm_sort_destroy(MsmSort& s)
{
    cudaFree(s.counts);
    cudaFree(s.offsets);
    cudaFree(s.cursor);
    cudaFree(s.sorted);
    s = MsmSort();
}

void msm_sort_run(MsmSort& s, const uint32_t* scalars, cudaStream_t st)
{
    size_t nb = kMsmBuckets + 2;
    KZP_CUDA_CHECK(cudaMemsetAsync(s.counts, 0, nb * 4, st));
    if (s.n > 0)
    {
        k_msm_count<<<sort_div_up(s.n, 256), 256, 0, st>>>(scalars, s.scalar_idx, s.scalar_offset, s.n, s.counts);
        KZP_CUDA_CHECK(cudaGetLastError());
    }
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
