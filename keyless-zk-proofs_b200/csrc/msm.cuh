// Pippenger MSM pipeline templates (signed digits -> counting sort -> chunked bucket accumulation -> bucket
// reduction) for sm_100a; instantiated for G1 in msm_g1.cu and for G2 in msm_g2.cu so the two build in parallel.
// Replaces ParallelMultiexp<Curve>::multiexp (rust-rapidsnark/rapidsnark/src/multiexp.cpp:183-245) and the
// group law it calls (curve.cpp). Integer pipes only.
#pragma once

#include <algorithm>
#include <cstring>
#include <vector>

#include "device.hpp"

namespace kzp
{

static inline unsigned int msm_div_up(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

// =====================================================================================================
// MSM
// =====================================================================================================
// ---- scalar handling --------------------------------------------------------------------------------
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
    k_msm_count(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t n,
                uint32_t* __restrict__ counts)
{
    uint32_t i     = blockIdx.x * blockDim.x + threadIdx.x;
    bool     valid = i < n;
    Fr       s     = Fr::zero();
    if (valid)
        load_scalar(scalars, scalar_idx[i], s);
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
    k_msm_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
               uint32_t* __restrict__ cursor)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t      per = (kMsmBuckets + 1023) / 1024; // 32
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
    // block exclusive scan of `sum`
    uint32_t lane = tid & 31, wid = tid >> 5;
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

// Pass 2: scatter entry = (window * n + base) | sign << 31 into its bucket's range.
static __global__ void __launch_bounds__(256)
    k_msm_scatter(const uint32_t* __restrict__ scalars, const uint32_t* __restrict__ scalar_idx, uint32_t n,
                  uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted)
{
    uint32_t i     = blockIdx.x * blockDim.x + threadIdx.x;
    bool     valid = i < n;
    Fr       s     = Fr::zero();
    if (valid)
        load_scalar(scalars, scalar_idx[i], s);
    bool     small = valid && scalar_is_small(s);
    uint32_t lane  = threadIdx.x & 31;
    uint32_t key   = (small && s.v[0] != 0) ? s.v[0] : (0xffff0000u | lane);
    uint32_t peers = __match_any_sync(0xffffffffu, key);
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
            sorted[pos]  = ((uint32_t)j * n + i) | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

// ---- bucket accumulation ----------------------------------------------------------------------------
// Thread t owns sorted[t*L, (t+1)*L). It emits one partial sum ("record") per bucket it touches at slot
// t + bucket: the map (t, bucket) -> t + bucket is injective and monotone over the pairs that occur, and
// the records of bucket b are exactly slots [lo/L + b, (hi-1)/L + b] for its range [lo, hi). Work per
// thread is therefore independent of the digit distribution (a bit-heavy witness puts ~half of all
// entries in bucket 1).
template <class XY>
__global__ void __launch_bounds__(128)
    k_msm_accumulate(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ sorted,
                     const typename XY::Affine* __restrict__ table, XY* __restrict__ records, uint32_t chunk)
{
    typedef typename XY::Affine Affine;
    typedef typename XY::Field  F;
    uint32_t t     = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total = offsets[kMsmBuckets + 1];
    uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= total)
        return;
    uint32_t start = (uint32_t)start64;
    uint32_t end   = min(start + chunk, total);
    // largest b in [1, B] with offsets[b] <= start
    uint32_t lo = 1, hi = kMsmBuckets;
    while (lo < hi)
    {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (offsets[mid] <= start)
            lo = mid;
        else
            hi = mid - 1;
    }
    uint32_t cur  = lo;
    uint32_t next = offsets[cur + 1];
    XY       acc;
    XY::set_inf(acc);
    for (uint32_t pos = start; pos < end; pos++)
    {
        if (pos >= next)
        {
            records[t + cur] = acc;
            XY::set_inf(acc);
            do
            {
                cur++;
                next = offsets[cur + 1];
            } while (pos >= next);
        }
        uint32_t e = sorted[pos];
        Affine   p = table[e & 0x7fffffffu];
        if (e >> 31)
            F::neg(p.y, p.y);
        XY::madd(acc, p);
    }
    records[t + cur] = acc;
}

// Out-of-line group operations for the cold kernels (bucket finalisation, reduction, table construction):
// keeps their code size and compile time down; the hot accumulate kernel inlines everything.
template <class XY>
__device__ __noinline__ void cold_add(XY& acc, const XY& q)
{
    XY::add(acc, q);
}
template <class XY>
__device__ __noinline__ void cold_dbl(XY& p)
{
    XY t = p;
    XY::dbl(p, t);
}

template <class XY>
__device__ __forceinline__ void block_tree_sum(XY* sm, uint32_t active, uint32_t tid)
{
    // sm[0..active) hold points; result in sm[0]
    for (uint32_t stride = active >> 1; stride > 0; stride >>= 1)
    {
        __syncthreads();
        if (tid < stride)
        {
            XY a = sm[tid], b = sm[tid + stride];
            cold_add(a, b);
            sm[tid] = a;
        }
    }
    __syncthreads();
}

// Bucket finalisation: sum the records of each bucket. Light buckets (the common case: a few dozen records) take
// one thread each; buckets with more than kHeavyRecords records (a bit-heavy witness puts ~half of all entries in
// bucket 1) are queued and summed by whole blocks in a second kernel, so no thread serialises a long run.
constexpr uint32_t kHeavyRecords = 96;

template <class XY>
__global__ void __launch_bounds__(128)
    k_msm_bucket_finalize(const uint32_t* __restrict__ offsets, const XY* __restrict__ records,
                          XY* __restrict__ buckets, uint32_t chunk, uint32_t* __restrict__ heavy_count,
                          uint32_t* __restrict__ heavy_ids)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (b > kMsmBuckets)
        return;
    uint32_t lo = offsets[b], hi = offsets[b + 1];
    XY       acc;
    XY::set_inf(acc);
    if (lo != hi)
    {
        uint32_t t0  = lo / chunk;
        uint32_t t1  = (hi - 1) / chunk;
        uint32_t cnt = t1 - t0 + 1;
        if (cnt > kHeavyRecords)
        {
            heavy_ids[atomicAdd(heavy_count, 1u)] = b;
            return;
        }
        const XY* rec = records + (size_t)t0 + b;
        acc           = rec[0];
        for (uint32_t k = 1; k < cnt; k++)
        {
            XY r = rec[k];
            if (sizeof(XY) > 128)
                cold_add(acc, r);
            else
                XY::add(acc, r);
        }
    }
    buckets[b] = acc;
}

template <class XY>
__global__ void __launch_bounds__(256)
    k_msm_bucket_finalize_heavy(const uint32_t* __restrict__ offsets, const XY* __restrict__ records,
                                XY* __restrict__ buckets, uint32_t chunk, const uint32_t* __restrict__ heavy_count,
                                const uint32_t* __restrict__ heavy_ids)
{
    extern __shared__ uint4 smem_raw[];
    XY*                     sm  = reinterpret_cast<XY*>(smem_raw);
    uint32_t                tid = threadIdx.x;
    uint32_t                n   = *heavy_count;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x)
    {
        uint32_t  b   = heavy_ids[i];
        uint32_t  lo  = offsets[b], hi = offsets[b + 1];
        uint32_t  t0  = lo / chunk;
        uint32_t  cnt = (hi - 1) / chunk - t0 + 1;
        const XY* rec = records + (size_t)t0 + b;
        XY        acc;
        XY::set_inf(acc);
        for (uint32_t k = tid; k < cnt; k += blockDim.x)
        {
            XY r = rec[k];
            cold_add(acc, r);
        }
        sm[tid] = acc;
        block_tree_sum(sm, blockDim.x, tid);
        if (tid == 0)
            buckets[b] = sm[0];
        __syncthreads();
    }
}

// ---- bucket reduction: sum_b b * bucket[b] ------------------------------------------------------------
// Weighted tree over a power-of-two run held in shared memory: each node keeps (S, W) = (sum P_i,
// sum i * P_i) of its segment; merging two segments of length len: S = Sl + Sr, W = Wl + Wr + len * Sr.
template <class XY>
__device__ __forceinline__ void weighted_tree(XY* S, XY* W, uint32_t count, uint32_t tid)
{
    uint32_t log_len = 0;
    for (uint32_t cnt = count; cnt > 1; cnt >>= 1, log_len++)
    {
        XY   s, w;
        bool on = tid < (cnt >> 1);
        __syncthreads();
        if (on)
        {
            XY sl = S[2 * tid], sr = S[2 * tid + 1];
            XY wl = W[2 * tid], wr = W[2 * tid + 1];
            s     = sl;
            cold_add(s, sr);
            w = wl;
            cold_add(w, wr);
            for (uint32_t k = 0; k < log_len; k++)
                cold_dbl(sr);
            cold_add(w, sr);
        }
        __syncthreads();
        if (on)
        {
            S[tid] = s;
            W[tid] = w;
        }
    }
    __syncthreads();
}

// stage 1: block k reduces buckets [256k+1, 256k+256] to (S_k, W_k) with local weights 0..255
template <class XY>
__global__ void __launch_bounds__(256)
    k_msm_reduce1(const XY* __restrict__ buckets, XY* __restrict__ partial)
{
    extern __shared__ uint4 smem_raw[];
    XY*                     S   = reinterpret_cast<XY*>(smem_raw);
    XY*                     W   = S + 256;
    uint32_t                tid = threadIdx.x;
    S[tid]                      = buckets[(size_t)blockIdx.x * 256 + 1 + tid];
    XY z;
    XY::set_inf(z);
    W[tid] = z;
    weighted_tree(S, W, 256, tid);
    if (tid == 0)
    {
        partial[2 * blockIdx.x]     = S[0];
        partial[2 * blockIdx.x + 1] = W[0];
    }
}

// stage 2: result = sum_k W_k + sum_k S_k + 256 * sum_k k * S_k      (K = B/256 blocks of stage 1)
template <class XY>
__global__ void __launch_bounds__(128)
    k_msm_reduce2(const XY* __restrict__ partial, XY* __restrict__ result)
{
    extern __shared__ uint4 smem_raw[];
    constexpr uint32_t      K   = kMsmBuckets / 256;
    XY*                     S   = reinterpret_cast<XY*>(smem_raw);
    XY*                     W   = S + K;
    XY*                     T   = W + K;
    uint32_t                tid = threadIdx.x;
    XY                      z;
    XY::set_inf(z);
    if (tid < K)
    {
        S[tid] = partial[2 * tid];
        T[tid] = partial[2 * tid + 1];
        W[tid] = z;
    }
    __syncthreads();
    block_tree_sum(T, K, tid);
    weighted_tree(S, W, K, tid);
    if (tid == 0)
    {
        XY r = W[0];
        for (int k = 0; k < 8; k++)
            cold_dbl(r);
        XY s0 = S[0], t0 = T[0];
        cold_add(r, s0);
        cold_add(r, t0);
        result[0] = r;
    }
}

// ---- table construction (once per proving key) ---------------------------------------------------------
template <class XY>
__global__ void __launch_bounds__(128)
    k_tbl_double(XY* __restrict__ cur, uint32_t n, int doublings, const typename XY::Affine* __restrict__ src)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    XY p;
    if (src)
        XY::from_affine(p, src[i]);
    else
        p = cur[i];
    for (int k = 0; k < doublings; k++)
        cold_dbl(p);
    cur[i] = p;
}

// Affine normalisation with Montgomery's batch inversion over kBatch consecutive points per thread.
template <class XY>
__global__ void __launch_bounds__(128)
    k_tbl_normalise(const XY* __restrict__ cur, uint32_t n, typename XY::Affine* __restrict__ out)
{
    typedef typename XY::Field F;
    constexpr int              kBatch = 8;
    uint32_t                   t      = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t                   first  = t * kBatch;
    if (first >= n)
        return;
    uint32_t cnt = min((uint32_t)kBatch, n - first);
    F        prefix[kBatch];
    F        run = F::one();
    for (uint32_t k = 0; k < cnt; k++)
    {
        F z = cur[first + k].zzz;
        if (F::is_zero(z))
            z = F::one();
        F::mul(run, run, z);
        prefix[k] = run;
    }
    F inv;
    F::inv(inv, run);
    for (int k = (int)cnt - 1; k >= 0; k--)
    {
        XY p = cur[first + k];
        typename XY::Affine a;
        if (F::is_zero(p.zzz))
        {
            a.x = F::zero();
            a.y = F::zero();
        }
        else
        {
            F izzz;
            if (k > 0)
                F::mul(izzz, inv, prefix[k - 1]);
            else
                izzz = inv;
            F::mul(inv, inv, p.zzz);
            F tt, izz;
            F::mul(tt, p.zz, izzz);
            F::sqr(izz, tt);
            F::mul(a.x, p.x, izz);
            F::mul(a.y, p.y, izzz);
        }
        out[first + k] = a;
    }
}

template <class XY>
void msm_bases_create(MsmBases<XY>& out, const uint8_t* bases_host, uint64_t first, uint64_t count,
                      uint32_t scalar_offset, cudaStream_t st)
{
    typedef typename XY::Affine Affine;
    const size_t                psz = sizeof(Affine);
    std::vector<uint32_t>       idx;
    idx.reserve(count);
    std::vector<uint8_t> packed;
    packed.reserve(count * psz);
    for (uint64_t k = 0; k < count; k++)
    {
        const uint8_t* p  = bases_host + (first + k) * psz;
        bool           nz = false;
        for (size_t q = 0; q < psz / 8; q++)
        {
            uint64_t v;
            memcpy(&v, p + 8 * q, 8);
            if (v)
            {
                nz = true;
                break;
            }
        }
        if (!nz)
            continue; // infinity base: skipped exactly like multiexp.cpp:57
        idx.push_back(scalar_offset + (uint32_t)k);
        packed.insert(packed.end(), p, p + psz);
    }
    out.n = (uint32_t)idx.size();
    if ((uint64_t)out.n * kMsmWindows >= 0x80000000ull)
        throw CudaError("MSM too large for 31-bit entry ids");
    if (out.n == 0)
        return;
    size_t n = out.n;
    KZP_CUDA_CHECK(cudaMalloc(&out.scalar_idx, n * 4));
    KZP_CUDA_CHECK(cudaMalloc(&out.table, n * kMsmWindows * psz));
    KZP_CUDA_CHECK(cudaMemcpyAsync(out.scalar_idx, idx.data(), n * 4, cudaMemcpyHostToDevice, st));
    KZP_CUDA_CHECK(cudaMemcpyAsync(out.table, packed.data(), n * psz, cudaMemcpyHostToDevice, st));
    XY* cur = nullptr;
    KZP_CUDA_CHECK(cudaMalloc(&cur, n * sizeof(XY)));
    unsigned int grid = msm_div_up(n, 128);
    for (int j = 1; j < kMsmWindows; j++)
    {
        k_tbl_double<XY><<<grid, 128, 0, st>>>(cur, out.n, kMsmWindowBits, j == 1 ? out.table : nullptr);
        KZP_CUDA_CHECK(cudaGetLastError());
        k_tbl_normalise<XY><<<msm_div_up(n, 128 * 8), 128, 0, st>>>(cur, out.n, out.table + (size_t)j * n);
        KZP_CUDA_CHECK(cudaGetLastError());
    }
    KZP_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(cur);
}

template <class XY>
void msm_bases_destroy(MsmBases<XY>& b)
{
    cudaFree(b.scalar_idx);
    cudaFree(b.table);
    b.scalar_idx = nullptr;
    b.table      = nullptr;
    b.n          = 0;
}

template <class XY>
static void msm_set_smem_attrs()
{
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_bucket_finalize_heavy<XY>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(256 * sizeof(XY))));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_reduce1<XY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(512 * sizeof(XY))));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_reduce2<XY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(3 * (kMsmBuckets / 256) * sizeof(XY))));
}

template <class XY>
void msm_scratch_create(MsmScratch<XY>& s, uint32_t n_active)
{
    msm_set_smem_attrs<XY>(); // function attributes are per device: set on the device that owns the scratch
    uint64_t cap  = (uint64_t)n_active * kMsmWindows;
    s.cap_entries = (uint32_t)cap;
    size_t nb     = kMsmBuckets + 2;
    KZP_CUDA_CHECK(cudaMalloc(&s.counts, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.offsets, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.cursor, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.sorted, std::max<uint64_t>(cap, 1) * 4));
    s.chunk = cap >= (1u << 24) ? 2 * kMsmChunk : kMsmChunk; // fewer, longer runs for the big MSMs
    KZP_CUDA_CHECK(cudaMalloc(&s.records, (cap / s.chunk + nb + 1) * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy, (nb + 1) * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.buckets, (kMsmBuckets + 1) * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.partial, 2 * (kMsmBuckets / 256) * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.result, sizeof(XY)));
    KZP_CUDA_CHECK(cudaEventCreate(&s.ev_acc0));
    KZP_CUDA_CHECK(cudaEventCreate(&s.ev_acc1));
}

template <class XY>
void msm_scratch_destroy(MsmScratch<XY>& s)
{
    cudaFree(s.counts);
    cudaFree(s.offsets);
    cudaFree(s.cursor);
    cudaFree(s.sorted);
    cudaFree(s.records);
    cudaFree(s.heavy);
    cudaFree(s.buckets);
    cudaFree(s.partial);
    cudaFree(s.result);
    if (s.ev_acc0)
        cudaEventDestroy(s.ev_acc0);
    if (s.ev_acc1)
        cudaEventDestroy(s.ev_acc1);
    s = MsmScratch<XY>();
}

template <class XY>
void msm_run(const MsmBases<XY>& b, MsmScratch<XY>& s, const uint32_t* scalars, cudaStream_t st)
{
    size_t nb = kMsmBuckets + 2;
    KZP_CUDA_CHECK(cudaMemsetAsync(s.counts, 0, nb * 4, st));
    if (b.n > 0)
    {
        if ((uint64_t)b.n * kMsmWindows > s.cap_entries)
            throw CudaError("MSM scratch too small");
        k_msm_count<<<msm_div_up(b.n, 256), 256, 0, st>>>(scalars, b.scalar_idx, b.n, s.counts);
        KZP_CUDA_CHECK(cudaGetLastError());
    }
    k_msm_scan<<<1, 1024, 0, st>>>(s.counts, s.offsets, s.cursor);
    KZP_CUDA_CHECK(cudaGetLastError());
    if (b.n > 0)
    {
        k_msm_scatter<<<msm_div_up(b.n, 256), 256, 0, st>>>(scalars, b.scalar_idx, b.n, s.cursor, s.sorted);
        KZP_CUDA_CHECK(cudaGetLastError());
        uint64_t threads = ((uint64_t)b.n * kMsmWindows + s.chunk - 1) / s.chunk;
        KZP_CUDA_CHECK(cudaEventRecord(s.ev_acc0, st));
        k_msm_accumulate<XY><<<msm_div_up(threads, 128), 128, 0, st>>>(s.offsets, s.sorted, b.table, s.records, s.chunk);
        KZP_CUDA_CHECK(cudaGetLastError());
        KZP_CUDA_CHECK(cudaEventRecord(s.ev_acc1, st));
    }
    KZP_CUDA_CHECK(cudaMemsetAsync(s.heavy, 0, 4, st));
    k_msm_bucket_finalize<XY><<<msm_div_up(kMsmBuckets, 128), 128, 0, st>>>(s.offsets, s.records, s.buckets, s.chunk, s.heavy, s.heavy + 1);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_msm_bucket_finalize_heavy<XY><<<128, 256, 256 * sizeof(XY), st>>>(s.offsets, s.records, s.buckets, s.chunk, s.heavy, s.heavy + 1);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_msm_reduce1<XY><<<kMsmBuckets / 256, 256, 512 * sizeof(XY), st>>>(s.buckets, s.partial);
    KZP_CUDA_CHECK(cudaGetLastError());
    k_msm_reduce2<XY><<<1, 128, 3 * (kMsmBuckets / 256) * sizeof(XY), st>>>(s.partial, s.result);
    KZP_CUDA_CHECK(cudaGetLastError());
}

template <class XY>
void msm_last_accumulate(const MsmScratch<XY>& s, float* ms, uint64_t* entries)
{
    float t = 0;
    if (s.ev_acc0 && cudaEventElapsedTime(&t, s.ev_acc0, s.ev_acc1) != cudaSuccess)
    {
        cudaGetLastError();
        t = 0;
    }
    uint32_t total = 0;
    KZP_CUDA_CHECK(cudaMemcpy(&total, s.offsets + kMsmBuckets + 1, 4, cudaMemcpyDeviceToHost));
    if (ms)
        *ms = t;
    if (entries)
        *entries = total;
}

template <class XY>
__global__ void __launch_bounds__(128)
    k_point_op(int op, const XY* __restrict__ p, const void* __restrict__ q, XY* __restrict__ out, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    XY acc = p[i];
    if (op == 0)
        XY::madd(acc, reinterpret_cast<const typename XY::Affine*>(q)[i]);
    else if (op == 1)
        XY::add(acc, reinterpret_cast<const XY*>(q)[i]);
    else
    {
        XY t = acc;
        XY::dbl(acc, t);
    }
    out[i] = acc;
}

template <class XY>
void point_op_t(int op, const void* p, const void* q, void* out, uint64_t count, cudaStream_t st)
{
    if (count == 0)
        return;
    k_point_op<XY><<<msm_div_up(count, 128), 128, 0, st>>>(op, (const XY*)p, q, (XY*)out, count);
    KZP_CUDA_CHECK(cudaGetLastError());
}

} // namespace kzp
