// Group-specific part of the Pippenger MSM for sm_100a: chunked bucket accumulation over a shared digit sort
// (msm_sort.cu), heavy-bucket pre-reduction, bucket finalisation fused with a base-32 digit fold, and the final
// combination. Instantiated for G1 in msm_g1.cu and for G2 in msm_g2.cu so the two build in parallel.
// Replaces ParallelMultiexp<Curve>::multiexp (rust-rapidsnark/rapidsnark/src/multiexp.cpp:183-245) and the
// group law it calls (curve.cpp). Integer pipes only.
//
// Reduction sum_b b * B_b over bucket ids b = idx + 1, idx = ... + i2 * 1024 + i1 * 32 + i0 (base-32 digits):
//     result = sum_b B_b + sum_l 32^l * sum_v v * S_l[v],      S_l[v] = sum of the buckets whose digit l equals v
// so the 2^(c-1) buckets collapse into levels x 32 plain class sums (3 levels for c = 16, 4 for c = 20) (perfectly parallel trees, no weights) and one
// 32-term weighted sum that a single block finishes by bit decomposition. Every stage is sized by sequential
// point additions, which is what bounds these latency-limited kernels.
#pragma once

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "device.hpp"

namespace kzp
{

static inline unsigned int msm_div_up(uint64_t a, uint64_t b) { return (unsigned int)((a + b - 1) / b); }

template <class XY>
struct MsmBatchArgs
{
    const typename XY::Affine* table[kMsmMaxBatch];
    const uint8_t*             skip[kMsmMaxBatch]; // per base column: 1 = infinity (nullptr = no flags)
    XY*                        records[kMsmMaxBatch];
    uint32_t*                  heavy_count; // classification is shared by the batch (same sort, same chunk)
    uint32_t*                  work_counter; // kMsmMaxBatch chunk tickets of the accumulate kernel (heavy_count + 1)
    uint32_t*                  heavy_ids;
    uint32_t*                  heavy_slot;
    XY*                        heavy_partial[kMsmMaxBatch];
    XY*                        heavy_sum[kMsmMaxBatch];
    uint32_t*                  heavy_done[kMsmMaxBatch];
    XY*                        bsum[kMsmMaxBatch];
    XY*                        s0part[kMsmMaxBatch];
    XY*                        s1part[kMsmMaxBatch];
    XY*                        classes[kMsmMaxBatch];
    XY*                        result[kMsmMaxBatch];
};

// ---- bucket accumulation ----------------------------------------------------------------------------
// Thread t owns sorted[t*L, (t+1)*L). It emits one partial sum ("record") per bucket it touches at slot
// t + bucket: the map (t, bucket) -> t + bucket is injective and monotone over the pairs that occur, and
// the records of bucket b are exactly slots [lo/L + b, (hi-1)/L + b] for its range [lo, hi). Work per
// thread is therefore independent of the digit distribution (a bit-heavy witness puts ~half of all
// entries in bucket 1). The MSMs of a batch (same sort, different base tables) share ONE ticket counter: ticket k is
// the (k / nb)-th run of 32 chunks of MSM k % nb. (With one counter and one grid slice per MSM, the warps of the
// first slice raced for a third of the tickets, the losers exited, and every CTA stayed resident with one or two
// live warps until those finished — three under-filled phases, 1.2 ms for the A/B1/C batch instead of 0.75.)
template <class XY, int MINB>
__global__ void __launch_bounds__(128, MINB)
    k_msm_accumulate(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ sorted,
                     MsmBatchArgs<XY> args, uint32_t chunk, uint32_t n, uint32_t nbuckets, uint32_t nb)
{
    typedef typename XY::Affine Affine;
    typedef typename XY::Field  F;
    uint32_t*                   ticket  = args.work_counter;
    const uint32_t              total   = offsets[nbuckets + 1];
    const uint32_t              lane    = threadIdx.x & 31;
    const uint32_t              runs    = ((total + chunk - 1) / chunk + 31) / 32; // runs of 32 chunks per MSM
    // The grid is sized to what is resident at once; warps take 32 consecutive chunks at a time from a ticket
    // counter until the sorted lists are used up, so no SM idles through the tail of a last partial wave.
    for (;;)
    {
        uint32_t k = 0;
        if (lane == 0)
            k = atomicAdd(ticket, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= runs * nb)
            return;
        const uint32_t y     = k % nb;
        const uint32_t first = (k / nb) * 32;
        // (selects, not an indexed read of the parameter arrays: those would be copied to local memory)
        static_assert(kMsmMaxBatch == 3, "the selects below name the batch members");
        const Affine* __restrict__  table   = y == 0 ? args.table[0] : (y == 1 ? args.table[1] : args.table[2]);
        const uint8_t* __restrict__ skip    = y == 0 ? args.skip[0] : (y == 1 ? args.skip[1] : args.skip[2]);
        XY* __restrict__            records = y == 0 ? args.records[0] : (y == 1 ? args.records[1] : args.records[2]);
        uint32_t t       = first + lane;
        uint64_t start64 = (uint64_t)t * chunk;
        if (start64 >= total)
            continue;
        uint32_t start = (uint32_t)start64;
        uint32_t end   = min(start + chunk, total);
        // largest b in [1, B] with offsets[b] <= start
        uint32_t lo = 1, hi = nbuckets;
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
            uint32_t i = e & kMsmEntryBaseMask;
            if (skip && skip[i])
                continue; // infinity column: costs one byte, not a point load
            Affine p = table[(size_t)((e >> kMsmEntryBaseBits) & 15u) * n + i];
            if (e >> 31)
                F::neg(p.y, p.y);
            XY::madd(acc, p);
        }
        records[t + cur] = acc;
    }
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

// ---- cooperative group operations for the latency-bound kernels ---------------------------------------------------
// A lone point addition is a chain of 14 dependent field products (~5 us for a warp that has an SM to itself), and the
// bucket reduction is nothing but sequential additions. kCoop = 4 consecutive lanes therefore act as ONE logical
// thread: all four hold the same operands (loaded from the same addresses), each computes one of the up to four
// independent products of a dependency level, and the results travel by warp shuffle inside the group. The chain
// shrinks from 14 products to 4 (9 to 3 for a doubling). Exceptional cases are decided on replicated data, so the
// four lanes always branch together; the shuffles name only the group's lanes in their mask.
constexpr uint32_t kCoop = 4;

__device__ __forceinline__ uint32_t coop_mask() { return 0xfu << (threadIdx.x & 28u); }

// value of lane `src` of the group, on every lane of the group
template <class F>
__device__ __forceinline__ F coop_bcast(const F& v, uint32_t src, uint32_t mask)
{
    F               r;
    const uint32_t* in  = reinterpret_cast<const uint32_t*>(&v);
    uint32_t*       out = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(F) / 4); i++)
        out[i] = __shfl_sync(mask, in[i], (int)src, (int)kCoop);
    return r;
}

template <class F>
__device__ __forceinline__ F coop_pick(uint32_t sub, const F& a0, const F& a1, const F& a2, const F& a3)
{
    return sub == 0 ? a0 : (sub == 1 ? a1 : (sub == 2 ? a2 : a3));
}

// acc += q (add-2008-s, the same formulas as XyzzT::add) by the 4 lanes of a group; sub = lane index in the group
template <class XY>
__device__ __noinline__ void coop_add(XY& acc, const XY& q, uint32_t sub)
{
    typedef typename XY::Field F;
    if (XY::is_inf(q))
        return;
    if (XY::is_inf(acc))
    {
        acc = q;
        return;
    }
    const uint32_t mask = coop_mask();
    F              t, u;
    // level 1: U1 = x1 zz2 | U2 = x2 zz1 | S1 = y1 zzz2 | S2 = y2 zzz1
    F::mul(t, coop_pick(sub, acc.x, q.x, acc.y, q.y), coop_pick(sub, q.zz, acc.zz, q.zzz, acc.zzz));
    F U1 = coop_bcast(t, 0, mask), S1 = coop_bcast(t, 2, mask), P, R;
    u = coop_bcast(t, 1, mask);
    F::sub(P, u, U1);
    u = coop_bcast(t, 3, mask);
    F::sub(R, u, S1);
    if (F::is_zero(P))
    {
        if (F::is_zero(R))
        {
            XY c = acc;
            XY::dbl(acc, c); // same point twice: rare, every lane doubles for itself
            return;
        }
        XY::set_inf(acc);
        return;
    }
    // level 2: PP = P^2 | RR = R^2 | ZZ = zz1 zz2 | ZZZ = zzz1 zzz2
    F::mul(t, coop_pick(sub, P, R, acc.zz, acc.zzz), coop_pick(sub, P, R, q.zz, q.zzz));
    F PP = coop_bcast(t, 0, mask), RR = coop_bcast(t, 1, mask);
    // level 3: PPP = P PP | Q = U1 PP | zz3 = ZZ PP | (lane 3 keeps ZZZ, computes PPP as well)
    F keep = t;
    F::mul(t, coop_pick(sub, P, U1, keep, P), PP);
    F PPP = coop_bcast(t, 0, mask), Q = coop_bcast(t, 1, mask);
    F zz3 = coop_bcast(t, 2, mask);
    F x3;
    F::sub(x3, RR, PPP);
    F::sub(x3, x3, Q);
    F::sub(x3, x3, Q);
    // level 4: B = S1 PPP | A = R (Q - x3) | - | zzz3 = ZZZ PPP
    F::sub(u, Q, x3);
    F::mul(t, coop_pick(sub, S1, R, S1, keep), coop_pick(sub, PPP, u, PPP, PPP));
    F A = coop_bcast(t, 1, mask);
    u   = coop_bcast(t, 0, mask);
    acc.x = x3;
    F::sub(acc.y, A, u);
    acc.zz  = zz3;
    acc.zzz = coop_bcast(t, 3, mask);
}

// p = 2 p (dbl-2008-s with a = 0, the same formulas as XyzzT::dbl) by the 4 lanes of a group
template <class XY>
__device__ __noinline__ void coop_dbl(XY& p, uint32_t sub)
{
    typedef typename XY::Field F;
    if (XY::is_inf(p))
        return;
    const uint32_t mask = coop_mask();
    F              U, t, u;
    F::add(U, p.y, p.y);
    // level 1: V = U^2 | XX = x^2
    F::mul(t, coop_pick(sub, U, p.x, U, p.x), coop_pick(sub, U, p.x, U, p.x));
    F V = coop_bcast(t, 0, mask), M = coop_bcast(t, 1, mask);
    F::add(u, M, M);
    F::add(M, M, u); // M = 3 x^2
    // level 2: W = U V | S = x V | MM = M^2 | zz3 = V zz
    F::mul(t, coop_pick(sub, U, p.x, M, V), coop_pick(sub, V, V, M, p.zz));
    F W = coop_bcast(t, 0, mask), S = coop_bcast(t, 1, mask), x3 = coop_bcast(t, 2, mask), zz3 = coop_bcast(t, 3, mask);
    F::sub(x3, x3, S);
    F::sub(x3, x3, S);
    // level 3: A = M (S - x3) | B = W y | zzz3 = W zzz
    F::sub(u, S, x3);
    F::mul(t, coop_pick(sub, M, W, W, W), coop_pick(sub, u, p.y, p.zzz, p.zzz));
    F A = coop_bcast(t, 0, mask);
    u   = coop_bcast(t, 1, mask);
    p.x = x3;
    F::sub(p.y, A, u);
    p.zz  = zz3;
    p.zzz = coop_bcast(t, 2, mask);
}

// one addition by a logical thread of LANES lanes: 4 = cooperative (latency), 1 = a lane on its own (throughput)
template <class XY, int LANES>
__device__ __forceinline__ void group_add(XY& acc, const XY& q, uint32_t sub)
{
    if constexpr (LANES == 1)
        cold_add(acc, q);
    else
        coop_add(acc, q, sub);
}

// sm[0..active) hold points (one per LOGICAL thread lt = threadIdx.x / kCoop, written by all four of its lanes);
// result in sm[0]. Every thread of the block must call this.
template <class XY>
__device__ __forceinline__ void block_tree_sum(XY* sm, uint32_t active, uint32_t lt, uint32_t sub)
{
    for (uint32_t stride = active >> 1; stride > 0; stride >>= 1)
    {
        __syncthreads();
        XY a;
        if (lt < stride)
        {
            a    = sm[lt];
            XY b = sm[lt + stride];
            coop_add(a, b, sub);
        }
        __syncthreads(); // all lanes of a group have read before any of them writes the slot back
        if (lt < stride && sub == 0)
            sm[lt] = a;
    }
    __syncthreads();
}


// L2-coherent load of a point written by another block of the same launch
template <class XY>
__device__ __forceinline__ XY load_cg(const XY* p)
{
    XY           r;
    const uint4* src = reinterpret_cast<const uint4*>(p);
    uint4*       dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XY) / 16); i++)
        dst[i] = __ldcg(src + i);
    return r;
}

// ---- bucket classification ------------------------------------------------------------------------------
// A bucket whose range spans more accumulate threads than max(kMsmHeavyRecords, 2 x average) is "heavy": it gets a
// slot in the heavy list and is pre-reduced by whole blocks (k_msm_heavy) instead of one finalise thread.
static __global__ void __launch_bounds__(256)
    k_msm_classify(const uint32_t* __restrict__ offsets, uint32_t chunk, uint32_t* __restrict__ heavy_count,
                   uint32_t* __restrict__ heavy_ids, uint32_t* __restrict__ heavy_slot, uint32_t nbuckets)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (b > nbuckets)
        return;
    uint32_t total = offsets[nbuckets + 1];
    uint32_t thr   = max(kMsmHeavyRecords, 2u * (total / chunk / nbuckets + 1u));
    uint32_t lo = offsets[b], hi = offsets[b + 1];
    uint32_t slot = 0;
    if (hi > lo && (hi - 1) / chunk - lo / chunk + 1 > thr)
    {
        uint32_t pos = atomicAdd(heavy_count, 1u);
        if (pos < kMsmMaxHeavy)
        {
            heavy_ids[pos] = b;
            slot           = pos + 1;
        }
    }
    heavy_slot[b] = slot;
}

// ---- heavy buckets -------------------------------------------------------------------------------------
// A bucket whose range spans more accumulate threads than the sort's heavy threshold (bucket 1 of a bit-heavy
// witness holds ~half of all entries; byte-valued wires fill buckets 2..255) is summed by whole blocks: its records
// are cut into slices of ~kMsmHeavySlice, one block reduces one slice, and for multi-slice buckets the last block
// to arrive (arrival counter) sums the slice results. Work items (bucket, slice) are spread over the grid.
// grid = (kMsmHeavyGrid, batch).
template <class XY>
__global__ void __launch_bounds__(256)
    k_msm_heavy(const uint32_t* __restrict__ offsets, MsmBatchArgs<XY> args, uint32_t chunk)
{
    const uint32_t* __restrict__ heavy_ids = args.heavy_ids;
    extern __shared__ uint4 smem_raw[];
    __shared__ uint32_t     s_last;
    __shared__ uint32_t     pre[kMsmMaxHeavy + 1];
    XY*                     sm      = reinterpret_cast<XY*>(smem_raw);
    const XY* __restrict__  records = args.records[blockIdx.y];
    XY*                     partial = args.heavy_partial[blockIdx.y];
    XY*                     sums    = args.heavy_sum[blockIdx.y];
    uint32_t*               done    = args.heavy_done[blockIdx.y];
    uint32_t                tid     = threadIdx.x;
    const uint32_t          lt = tid / kCoop, sub = tid % kCoop, lthreads = blockDim.x / kCoop; // logical threads (coop_add)
    uint32_t                n       = min(*args.heavy_count, kMsmMaxHeavy);
    if (n == 0)
        return;
    // slices per heavy bucket, then an exclusive prefix (n <= 1024: one thread scans)
    for (uint32_t h = tid; h < n; h += blockDim.x)
    {
        uint32_t b   = heavy_ids[h];
        uint32_t lo  = offsets[b], hi = offsets[b + 1];
        uint32_t cnt = (hi - 1) / chunk - lo / chunk + 1;
        pre[h + 1]   = min(kMsmHeavyBlocks, (cnt + kMsmHeavySlice - 1) / kMsmHeavySlice);
    }
    __syncthreads();
    if (tid == 0)
    {
        pre[0] = 0;
        for (uint32_t h = 0; h < n; h++)
            pre[h + 1] += pre[h];
    }
    __syncthreads();
    uint32_t items = pre[n];
    for (uint32_t item = blockIdx.x; item < items; item += gridDim.x)
    {
        // h = largest index with pre[h] <= item
        uint32_t l = 0, r = n - 1;
        while (l < r)
        {
            uint32_t mid = (l + r + 1) >> 1;
            if (pre[mid] <= item)
                l = mid;
            else
                r = mid - 1;
        }
        uint32_t  h   = l;
        uint32_t  j   = item - pre[h];
        uint32_t  nsl = pre[h + 1] - pre[h];
        uint32_t  b   = heavy_ids[h];
        uint32_t  lo  = offsets[b], hi = offsets[b + 1];
        uint32_t  t0  = lo / chunk;
        uint32_t  cnt = (hi - 1) / chunk - t0 + 1;
        const XY* rec = records + (size_t)t0 + b;
        uint32_t  per = (cnt + nsl - 1) / nsl;
        uint32_t  s   = j * per;
        uint32_t  e   = min(s + per, cnt);
        XY        acc;
        XY::set_inf(acc);
        for (uint32_t k = s + lt; k < e; k += lthreads)
        {
            XY q = rec[k];
            coop_add(acc, q, sub);
        }
        if (sub == 0)
            sm[lt] = acc;
        block_tree_sum(sm, lthreads, lt, sub);
        if (nsl == 1)
        {
            if (tid == 0)
                sums[h] = sm[0];
            __syncthreads();
            continue;
        }
        if (tid == 0)
        {
            partial[(size_t)h * kMsmHeavyBlocks + j] = sm[0];
            __threadfence();
            uint32_t ticket = atomicAdd(&done[h], 1u);
            s_last          = (ticket == nsl - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (s_last)
        {
            __threadfence();
            // (a block has at least 16 logical threads; kMsmHeavyBlocks slots are filled in strides of that)
            for (uint32_t k = lt; k < kMsmHeavyBlocks; k += lthreads)
                if (sub == 0)
                {
                    if (k < nsl)
                        sm[k] = load_cg(partial + (size_t)h * kMsmHeavyBlocks + k);
                    else
                        XY::set_inf(sm[k]);
                }
            block_tree_sum(sm, kMsmHeavyBlocks, lt, sub);
            if (tid == 0)
            {
                sums[h] = sm[0];
                done[h] = 0;
            }
        }
        __syncthreads();
    }
}

// ---- bucket sums ----------------------------------------------------------------------------------------
// Logical thread = bucket: sum its records (or take the heavy sum) and store the bucket's point. Every lane does the
// same kind of work (a bucket of a uniform scalar vector spans one or two chunks). LANES lanes per bucket: four
// (cooperative additions) when the MSM is small and the kernel is a latency chain, one when there are enough buckets
// to fill the chip (the replicated additions, subtractions and shuffles of the cooperative form then cost throughput:
// 2^19 buckets take 0.26 ms with four lanes, 0.16 ms with one). grid = (buckets * LANES / 128, batch).
template <class XY, int LANES>
__global__ void __launch_bounds__(kMsmFoldBlock)
    k_msm_bucket_sums(const uint32_t* __restrict__ offsets, MsmBatchArgs<XY> args, uint32_t chunk)
{
    const uint32_t* __restrict__ heavy_slot = args.heavy_slot;
    const XY* __restrict__       records    = args.records[blockIdx.y];
    const uint32_t               sub        = threadIdx.x % LANES;
    uint32_t                     b          = blockIdx.x * (kMsmFoldBlock / LANES) + threadIdx.x / LANES + 1;
    uint32_t                     lo = offsets[b], hi = offsets[b + 1];
    XY                           acc;
    XY::set_inf(acc);
    if (lo != hi)
    {
        uint32_t slot = heavy_slot[b];
        if (slot)
            acc = args.heavy_sum[blockIdx.y][slot - 1];
        else
        {
            uint32_t  t0  = lo / chunk;
            uint32_t  cnt = (hi - 1) / chunk - t0 + 1;
            const XY* rec = records + (size_t)t0 + b;
            acc           = rec[0];
            for (uint32_t k = 1; k < cnt; k++)
            {
                XY r = rec[k];
                group_add<XY, LANES>(acc, r, sub);
            }
        }
    }
    if (sub == 0)
        args.bsum[blockIdx.y][b - 1] = acc;
}

// lane exchange of a whole point (for the shuffle trees below)
template <class XY>
__device__ __forceinline__ XY shfl_xor_point(const XY& p, int mask)
{
    XY              r;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&p);
    uint32_t*       dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XY) / 4); i++)
        dst[i] = __shfl_xor_sync(0xffffffffu, src[i], mask);
    return r;
}

// ---- first fold: one plane of 1024 buckets (base-32 digits d1, d0 of the bucket index) per block -----------
// Threads 0..127 = (row d1, quarter): 8 consecutive buckets each, then a shuffle tree over the 4 quarters -> the
// row sum s1part[plane][d1] (sum over d0; feeds digit classes 1 and up). Threads 128..255 = (quarter of d1, d0):
// 8 buckets 32 apart -> s0q[plane][quarter][d0] (partial sums over d1; feed the digit-0 classes). All lanes run the
// same 7 sequential additions: no idle lanes as in a per-warp tree over 32 buckets. grid = (buckets / 1024, batch).
template <class XY, int LANES>
__global__ void __launch_bounds__(256) k_msm_plane_fold(MsmBatchArgs<XY> args)
{
    // LANES lanes per logical thread (see k_msm_bucket_sums): a block is 256 / LANES of the plane's 256 logical
    // threads, grid.x = LANES x planes
    const uint32_t         plane = blockIdx.x / LANES;
    const XY* __restrict__ bs    = args.bsum[blockIdx.y] + (size_t)plane * 1024;
    const uint32_t         sub   = threadIdx.x % LANES;
    const uint32_t         tid   = (blockIdx.x % LANES) * (256u / LANES) + threadIdx.x / LANES; // logical thread of the plane
    if (tid < 128)
    {
        const uint32_t row = tid >> 2, part = tid & 3u;
        const XY*      src = bs + row * 32 + part * 8;
        XY             acc = src[0];
#pragma unroll 1
        for (uint32_t k = 1; k < 8; k++)
        {
            XY r = src[k];
            group_add<XY, LANES>(acc, r, sub);
        }
#pragma unroll 1
        for (int m = 1; m <= 2; m <<= 1)
        {
            XY o = shfl_xor_point(acc, m * LANES); // the logical neighbour is LANES lanes away
            group_add<XY, LANES>(acc, o, sub);
        }
        if (part == 0 && sub == 0)
            args.s1part[blockIdx.y][(size_t)plane * 32 + row] = acc;
    }
    else
    {
        const uint32_t t = tid - 128, q4 = t >> 5, d0 = t & 31u;
        const XY*      src = bs + (q4 * 8) * 32 + d0;
        XY             acc = src[0];
#pragma unroll 1
        for (uint32_t k = 1; k < 8; k++)
        {
            XY r = src[k * 32];
            group_add<XY, LANES>(acc, r, sub);
        }
        if (sub == 0)
            args.s0part[blockIdx.y][((size_t)plane * 4 + q4) * 32 + d0] = acc;
    }
}

// ---- second fold: the levels x 32 class sums ---------------------------------------------------------------
// Class (l, v) is the sum of every partial of the buckets whose base-32 digit l equals v, scaled by 32^l:
//   l = 0: s0part[plane][quarter][v] over all planes and quarters
//   l = 1: s1part[plane][v] over all planes
//   l >= 2: digit l of the index is digit l - 2 of the plane number: all 32 rows of the matching planes
// The kernel is a latency chain (sequential point additions on a handful of SMs), so with many buckets every class is
// cut into `slices` <= kMsmFoldSlices slices, one block each: a block adds up its share of the terms, and the last slice
// of a class to arrive (arrival counter) adds the slice sums and applies the 5 l doublings. The grid must stay one wave
// (measured: 512 blocks of 256 threads are two waves and gain nothing over 128 unsliced blocks; 512 x 128 threads fit).
// grid = (32 * levels * slices, batch).
constexpr uint32_t kMsmFoldSlices = 4;
// layout of the `classes` scratch: [0, 32 L) class sums, then 32 L x slices slice sums (reused by k_msm_final for its
// six partial results); of `heavy_done` beyond the heavy buckets: 32 L class counters + 1 for k_msm_final
constexpr uint32_t kMsmFoldPartOffset = 32 * kMsmMaxLevels;
constexpr uint32_t kMsmFoldDoneOffset = kMsmMaxHeavy;

template <class XY>
__global__ void __launch_bounds__(256) k_msm_fold2(MsmBatchArgs<XY> args, uint32_t nbuckets, uint32_t slices)
{
    extern __shared__ uint4 smem_raw[];
    __shared__ uint32_t     s_last;
    XY*                     sm     = reinterpret_cast<XY*>(smem_raw);
    const uint32_t          planes = nbuckets >> 10;
    const XY* __restrict__  s0     = args.s0part[blockIdx.y];
    const XY* __restrict__  s1     = args.s1part[blockIdx.y];
    const uint32_t          tid = threadIdx.x / kCoop, sub = threadIdx.x % kCoop, nthr = blockDim.x / kCoop; // logical threads
    const uint32_t          cls = blockIdx.x / slices, slice = blockIdx.x % slices;
    const uint32_t          l = cls >> 5, v = cls & 31;
    const uint32_t          first = slice * nthr + tid, step = nthr * slices;
    XY                      acc;
    XY::set_inf(acc);
    if (l == 0)
    {
        for (uint32_t t = first; t < planes * 4; t += step)
        {
            XY r = s0[(size_t)t * 32 + v];
            coop_add(acc, r, sub);
        }
    }
    else if (l == 1)
    {
        for (uint32_t t = first; t < planes; t += step)
        {
            XY r = s1[(size_t)t * 32 + v];
            coop_add(acc, r, sub);
        }
    }
    else
    {
        // planes p with ((p >> shift) & 31) == v, shift = 5 (l - 2); term t = (matching plane index, row)
        uint32_t shift = 5 * (l - 2);
        uint32_t span  = 1u << shift; // consecutive planes sharing the digit
        uint32_t reps  = (planes + (span << 5) - 1) / (span << 5);
        uint32_t nterm = reps * span * 32;
        for (uint32_t t = first; t < nterm; t += step)
        {
            uint32_t row = t & 31u, q = t >> 5;
            uint32_t pl  = ((q >> shift) << (shift + 5)) | (v << shift) | (q & (span - 1));
            if (pl < planes)
            {
                XY r = s1[(size_t)pl * 32 + row];
                coop_add(acc, r, sub);
            }
        }
    }
    if (sub == 0)
        sm[tid] = acc;
    block_tree_sum(sm, nthr, tid, sub);
    XY*       part = args.classes[blockIdx.y] + kMsmFoldPartOffset + (size_t)cls * kMsmFoldSlices;
    uint32_t* done = args.heavy_done[blockIdx.y] + kMsmFoldDoneOffset + cls;
    if (threadIdx.x == 0)
    {
        part[slice] = sm[0];
        __threadfence();
        uint32_t ticket = atomicAdd(done, 1u);
        s_last          = (ticket == slices - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    if (tid == 0) // the first four lanes, as one logical thread
    {
        XY r = load_cg(part);
        for (uint32_t k = 1; k < slices; k++)
        {
            XY q = load_cg(part + k);
            coop_add(r, q, sub);
        }
        for (uint32_t k = 0; k < 5 * l; k++)
            coop_dbl(r, sub);
        if (sub == 0)
        {
            args.classes[blockIdx.y][cls] = r;
            *done                         = 0;
        }
    }
}

// ---- final combination -----------------------------------------------------------------------------------
// result = T + sum_v v * E[v], E[v] = sum_l classes[l][v] (already scaled), T = sum_v classes[0][v].
// Block k < 5 builds Z_k = sum of E[v] over v with bit k set, block 5 builds T: 32 logical threads of four lanes
// (cooperative additions), a tree over the 32 values; the last block to arrive runs Horner over k on its first four
// lanes. 3 + 5 + 9 sequential cooperative operations. grid = (6, batch), 128 threads.
template <class XY>
__global__ void __launch_bounds__(128) k_msm_final(MsmBatchArgs<XY> args, int levels)
{
    extern __shared__ uint4 smem_raw[];
    __shared__ uint32_t     s_last;
    XY*                     sm  = reinterpret_cast<XY*>(smem_raw);
    const XY* __restrict__  cls = args.classes[blockIdx.y];
    const uint32_t          v = threadIdx.x / kCoop, sub = threadIdx.x % kCoop, w = blockIdx.x;
    XY                      e   = cls[v];
    if (w < 5)
    {
        if ((v >> w) & 1u)
        {
#pragma unroll 1
            for (int l = 1; l < levels; l++)
            {
                XY u = cls[32 * l + v];
                coop_add(e, u, sub);
            }
        }
        else
            XY::set_inf(e);
    }
    if (sub == 0)
        sm[v] = e;
    block_tree_sum(sm, 32, v, sub);
    XY*       part = args.classes[blockIdx.y] + kMsmFoldPartOffset; // (the slice sums of k_msm_fold2 are spent)
    uint32_t* done = args.heavy_done[blockIdx.y] + kMsmFoldDoneOffset + 32 * kMsmMaxLevels;
    if (threadIdx.x == 0)
    {
        part[w] = sm[0];
        __threadfence();
        uint32_t ticket = atomicAdd(done, 1u);
        s_last          = (ticket == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last)
        return;
    __threadfence();
    if (v == 0)
    {
        XY r = load_cg(part + 4);
        for (int k = 3; k >= 0; k--)
        {
            coop_dbl(r, sub);
            XY z = load_cg(part + k);
            coop_add(r, z, sub);
        }
        XY t = load_cg(part + 5);
        coop_add(r, t, sub);
        if (sub == 0)
        {
            args.result[blockIdx.y][0] = r;
            *done                      = 0;
        }
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
void msm_bases_create(MsmBases<XY>& out, const uint8_t* points_host, uint64_t count, bool filter_inf,
                      cudaStream_t st, uint32_t window_bits)
{
    if (window_bits < kMsmMinWindowBits || window_bits > kMsmMaxWindowBits)
        throw CudaError("MSM window size out of range");
    out.shape = msm_shape(window_bits);
    typedef typename XY::Affine Affine;
    const size_t                psz = sizeof(Affine);
    std::vector<uint32_t>       idx;
    std::vector<uint8_t>        packed;
    const uint8_t*              src = points_host;
    if (filter_inf)
    {
        idx.reserve(count);
        packed.reserve(count * psz);
        for (uint64_t k = 0; k < count; k++)
        {
            const uint8_t* p  = points_host + k * psz;
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
            idx.push_back((uint32_t)k);
            packed.insert(packed.end(), p, p + psz);
        }
        out.n = (uint32_t)idx.size();
        src   = packed.data();
    }
    else
        out.n = (uint32_t)count;
    if (out.n > kMsmEntryBaseMask)
        throw CudaError("MSM too large for 27-bit base ids");
    if (out.n == 0)
        return;
    size_t n = out.n;
    if (!filter_inf)
    {
        std::vector<uint8_t> flags(n, 0);
        bool                 any = false;
        for (uint64_t k = 0; k < count; k++)
        {
            const uint8_t* p  = points_host + k * psz;
            bool           nz = false;
            for (size_t q = 0; q < psz / 8 && !nz; q++)
            {
                uint64_t v;
                memcpy(&v, p + 8 * q, 8);
                nz = v != 0;
            }
            flags[k] = nz ? 0 : 1;
            any |= !nz;
        }
        if (any)
        {
            KZP_CUDA_CHECK(cudaMalloc(&out.skip, n));
            KZP_CUDA_CHECK(cudaMemcpyAsync(out.skip, flags.data(), n, cudaMemcpyHostToDevice, st));
            KZP_CUDA_CHECK(cudaStreamSynchronize(st)); // flags is a local
        }
    }
    if (filter_inf)
    {
        KZP_CUDA_CHECK(cudaMalloc(&out.scalar_idx, n * 4));
        KZP_CUDA_CHECK(cudaMemcpyAsync(out.scalar_idx, idx.data(), n * 4, cudaMemcpyHostToDevice, st));
    }
    KZP_CUDA_CHECK(cudaMalloc(&out.table, n * out.shape.windows * psz));
    KZP_CUDA_CHECK(cudaMemcpyAsync(out.table, src, n * psz, cudaMemcpyHostToDevice, st));
    XY* cur = nullptr;
    KZP_CUDA_CHECK(cudaMalloc(&cur, n * sizeof(XY)));
    unsigned int grid = msm_div_up(n, 128);
    for (int j = 1; j < (int)out.shape.windows; j++)
    {
        k_tbl_double<XY><<<grid, 128, 0, st>>>(cur, out.n, (int)out.shape.c, j == 1 ? out.table : nullptr);
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
    cudaFree(b.skip);
    b.skip       = nullptr;
    b.scalar_idx = nullptr;
    b.table      = nullptr;
    b.n          = 0;
}

template <class XY>
static void msm_set_smem_attrs()
{
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_heavy<XY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(256 * sizeof(XY))));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_fold2<XY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(256 * sizeof(XY))));
    KZP_CUDA_CHECK(cudaFuncSetAttribute(k_msm_final<XY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(32 * sizeof(XY))));
}

template <class XY>
void msm_scratch_create(MsmScratch<XY>& s, const MsmSort& sort, uint32_t chunk)
{
    msm_set_smem_attrs<XY>(); // function attributes are per device: set on the device that owns the scratch
    s.shape     = sort.shape;
    size_t nb   = (size_t)s.shape.buckets + 2;
    s.chunk     = chunk ? chunk : msm_default_chunk(sort.n);
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy_count, 4 * (1 + kMsmMaxBatch))); // [0] heavy buckets, [1..] accumulate tickets
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy_ids, kMsmMaxHeavy * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy_slot, nb * 4));
    KZP_CUDA_CHECK(cudaMalloc(&s.records, ((size_t)sort.cap_entries / s.chunk + nb + 1) * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy_partial, (size_t)kMsmMaxHeavy * kMsmHeavyBlocks * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy_sum, (size_t)kMsmMaxHeavy * sizeof(XY)));
    // (arrival counters, all self-resetting: heavy buckets, then the classes of k_msm_fold2, then k_msm_final's)
    const size_t n_done = (size_t)kMsmFoldDoneOffset + 32 * kMsmMaxLevels + 1;
    KZP_CUDA_CHECK(cudaMalloc(&s.heavy_done, n_done * 4));
    KZP_CUDA_CHECK(cudaMemset(s.heavy_done, 0, n_done * 4));
    size_t planes = s.shape.buckets >> 10;
    KZP_CUDA_CHECK(cudaMalloc(&s.bsum, (size_t)s.shape.buckets * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.s0part, planes * 4 * 32 * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.s1part, planes * 32 * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.classes, (size_t)kMsmMaxLevels * 32 * (1 + kMsmFoldSlices) * sizeof(XY)));
    KZP_CUDA_CHECK(cudaMalloc(&s.result, sizeof(XY)));
    KZP_CUDA_CHECK(cudaEventCreate(&s.ev_acc0));
    KZP_CUDA_CHECK(cudaEventCreate(&s.ev_acc1));
}

template <class XY>
void msm_scratch_destroy(MsmScratch<XY>& s)
{
    cudaFree(s.heavy_count);
    cudaFree(s.heavy_ids);
    cudaFree(s.heavy_slot);
    cudaFree(s.records);
    cudaFree(s.heavy_partial);
    cudaFree(s.heavy_sum);
    cudaFree(s.heavy_done);
    cudaFree(s.bsum);
    cudaFree(s.s0part);
    cudaFree(s.s1part);
    cudaFree(s.classes);
    cudaFree(s.result);
    if (s.ev_acc0)
        cudaEventDestroy(s.ev_acc0);
    if (s.ev_acc1)
        cudaEventDestroy(s.ev_acc1);
    s = MsmScratch<XY>();
}

// CTAs of the accumulate kernel that are resident at once on the current device (occupancy x SMs)
template <class XY, int MINB>
static unsigned int msm_resident_blocks()
{
    static thread_local int cached_dev = -1;
    static thread_local unsigned int cached = 0;
    int dev = 0;
    KZP_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev != cached_dev)
    {
        int per_sm = 0, sms = 0;
        KZP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_msm_accumulate<XY, MINB>, 128, 0));
        KZP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cached     = (unsigned int)std::max(1, per_sm * sms);
        cached_dev = dev;
    }
    return cached;
}

template <class XY>
void msm_reduce_batch(const MsmSort& sort, const MsmBases<XY>* const* bases, MsmScratch<XY>* const* scr, int nb,
                      cudaStream_t st)
{
    if (nb < 1 || nb > kMsmMaxBatch)
        throw CudaError("MSM batch size out of range");
    MsmBatchArgs<XY> a;
    for (int k = 0; k < kMsmMaxBatch; k++)
    {
        int j = k < nb ? k : 0;
        if (bases[j]->n != sort.n || bases[j]->shape.c != sort.shape.c || scr[j]->shape.c != sort.shape.c)
            throw CudaError("MSM bases do not match the digit sort");
        a.table[k]         = bases[j]->table;
        a.skip[k]          = bases[j]->skip;
        a.records[k]       = scr[j]->records;
        a.heavy_partial[k] = scr[j]->heavy_partial;
        a.heavy_sum[k]     = scr[j]->heavy_sum;
        a.heavy_done[k]    = scr[j]->heavy_done;
        a.bsum[k]          = scr[j]->bsum;
        a.s0part[k]        = scr[j]->s0part;
        a.s1part[k]        = scr[j]->s1part;
        a.classes[k]       = scr[j]->classes;
        a.result[k]        = scr[j]->result;
    }
    const uint32_t chunk = scr[0]->chunk;
    for (int k = 1; k < nb; k++)
        if (scr[k]->chunk != chunk)
            throw CudaError("MSM batch members must share one chunk size");
    a.heavy_count  = scr[0]->heavy_count;
    a.work_counter = scr[0]->heavy_count + 1;
    a.heavy_ids   = scr[0]->heavy_ids;
    a.heavy_slot  = scr[0]->heavy_slot;
    const uint32_t nbuckets = sort.shape.buckets;
    dim3 by(1, (unsigned int)nb, 1);
    KZP_CUDA_CHECK(cudaMemsetAsync(a.heavy_count, 0, 4 * (1 + kMsmMaxBatch), st));
    k_msm_classify<<<msm_div_up(nbuckets, 256), 256, 0, st>>>(sort.offsets, chunk, a.heavy_count, a.heavy_ids,
                                                               a.heavy_slot, nbuckets);
    KZP_CUDA_CHECK(cudaGetLastError());
    KZP_CUDA_CHECK(cudaEventRecord(scr[0]->ev_acc0, st));
    if (sort.n > 0)
    {
        // one grid for the whole batch (the MSMs share the ticket counter), sized to what is resident at once
        uint64_t threads = (((uint64_t)sort.cap_entries + chunk - 1) / chunk) * (uint64_t)nb;
        dim3     ga(1, 1, 1);
        // resident CTAs per SM the kernel is compiled for (register cap): KZP_ACC_OCC = 4 | 5 | 6 (G1 only)
        static const int occ = getenv("KZP_ACC_OCC") ? atoi(getenv("KZP_ACC_OCC")) : 4;
        if (sizeof(XY) == 128 && occ == 5)
        {
            ga.x = std::min(msm_div_up(threads, 128), msm_resident_blocks<XY, 5>());
            k_msm_accumulate<XY, 5><<<ga, 128, 0, st>>>(sort.offsets, sort.sorted, a, chunk, sort.n, nbuckets, (uint32_t)nb);
        }
        else if (sizeof(XY) == 128 && occ == 6)
        {
            ga.x = std::min(msm_div_up(threads, 128), msm_resident_blocks<XY, 6>());
            k_msm_accumulate<XY, 6><<<ga, 128, 0, st>>>(sort.offsets, sort.sorted, a, chunk, sort.n, nbuckets, (uint32_t)nb);
        }
        else if constexpr (sizeof(XY) == 256)
        {
            // G2: capped at 168 registers so that three CTAs stay resident per SM (uncapped, the lazy-reduction Fq2
            // products raise the kernel to 197 registers: two CTAs)
            ga.x = std::min(msm_div_up(threads, 128), msm_resident_blocks<XY, 3>());
            k_msm_accumulate<XY, 3><<<ga, 128, 0, st>>>(sort.offsets, sort.sorted, a, chunk, sort.n, nbuckets, (uint32_t)nb);
        }
        else
        {
            ga.x = std::min(msm_div_up(threads, 128), msm_resident_blocks<XY, 1>());
            k_msm_accumulate<XY, 1><<<ga, 128, 0, st>>>(sort.offsets, sort.sorted, a, chunk, sort.n, nbuckets, (uint32_t)nb);
        }
        KZP_CUDA_CHECK(cudaGetLastError());
    }
    KZP_CUDA_CHECK(cudaEventRecord(scr[0]->ev_acc1, st));
    // The cold kernels below are latency-bound chains of point additions; what they cost the proof is the register
    // file they hold while the NTT / H-MSM CTAs of the other stream wait for room (a 256-thread G2 CTA holds all
    // 64 K registers of its SM). G2 therefore runs them as many narrow CTAs: measured on the keyless proof,
    // 256-thread x 128-CTA -> 64-thread x 256-CTA for k_msm_heavy and 256 -> 64 threads for k_msm_fold2 moves the
    // whole proof from 12.72 ms to 12.33 ms (B2 finishes later, in the shadow of the H MSM).
    constexpr bool is_g2   = sizeof(XY) == 256;
    static const bool wide = getenv("KZP_COLD_WIDE") && atoi(getenv("KZP_COLD_WIDE")) != 0; // A/B switch: old shapes
    // (four lanes per logical thread since the cooperative additions: KZP_COLD_G2_T threads per G2 CTA, default 128)
    static const uint32_t g2_t = getenv("KZP_COLD_G2_T") ? (uint32_t)atoi(getenv("KZP_COLD_G2_T")) : 128u;
    const uint32_t heavy_t = (is_g2 && !wide) ? g2_t : 256, heavy_g = (is_g2 && !wide) ? 2 * kMsmHeavyGrid : kMsmHeavyGrid;
    const uint32_t fold2_t = (is_g2 && !wide) ? g2_t : 256;
    by.x = heavy_g;
    k_msm_heavy<XY><<<by, heavy_t, heavy_t * sizeof(XY), st>>>(sort.offsets, a, chunk);
    KZP_CUDA_CHECK(cudaGetLastError());
    // bucket sums and the plane fold: cooperative additions (four lanes per logical thread) while the MSM is too
    // small to fill the chip with one lane per bucket, plain ones from KZP_FOLD_COOP_MAX buckets per batch on
    static const uint32_t coop_max = getenv("KZP_FOLD_COOP_MAX") ? (uint32_t)atoi(getenv("KZP_FOLD_COOP_MAX")) : (1u << 17);
    if ((uint64_t)nbuckets * (unsigned int)nb < coop_max)
    {
        by.x = nbuckets / (kMsmFoldBlock / kCoop);
        k_msm_bucket_sums<XY, (int)kCoop><<<by, kMsmFoldBlock, 0, st>>>(sort.offsets, a, chunk);
        KZP_CUDA_CHECK(cudaGetLastError());
        by.x = (nbuckets >> 10) * kCoop;
        k_msm_plane_fold<XY, (int)kCoop><<<by, 256, 0, st>>>(a);
    }
    else
    {
        by.x = nbuckets / kMsmFoldBlock;
        k_msm_bucket_sums<XY, 1><<<by, kMsmFoldBlock, 0, st>>>(sort.offsets, a, chunk);
        KZP_CUDA_CHECK(cudaGetLastError());
        by.x = nbuckets >> 10;
        k_msm_plane_fold<XY, 1><<<by, 256, 0, st>>>(a);
    }
    KZP_CUDA_CHECK(cudaGetLastError());
    // (sliced classes from 2^18 buckets on: the H MSM; 128-thread blocks then, so that the grid stays one wave)
    static const uint32_t slices_env = getenv("KZP_FOLD_SLICES") ? (uint32_t)atoi(getenv("KZP_FOLD_SLICES")) : 0u;
    const uint32_t slices  = slices_env ? std::min(slices_env, kMsmFoldSlices) : (nbuckets >= (1u << 18) ? kMsmFoldSlices : 1u);
    const uint32_t fold2_b = slices > 1 ? std::min(fold2_t, 128u) : fold2_t;
    by.x = 32 * sort.shape.levels * slices;
    k_msm_fold2<XY><<<by, fold2_b, fold2_b * sizeof(XY), st>>>(a, nbuckets, slices);
    KZP_CUDA_CHECK(cudaGetLastError());
    by.x = 6;
    k_msm_final<XY><<<by, 128, 32 * sizeof(XY), st>>>(a, (int)sort.shape.levels);
    KZP_CUDA_CHECK(cudaGetLastError());
}

template <class XY>
void msm_last_accumulate(const MsmSort& sort, const MsmScratch<XY>& s, float* ms, uint64_t* entries)
{
    float t = 0;
    if (s.ev_acc0 && cudaEventElapsedTime(&t, s.ev_acc0, s.ev_acc1) != cudaSuccess)
    {
        cudaGetLastError();
        t = 0;
    }
    uint32_t total = 0;
    KZP_CUDA_CHECK(cudaMemcpy(&total, sort.offsets + sort.shape.buckets + 1, 4, cudaMemcpyDeviceToHost));
    if (ms)
        *ms = t;
    if (entries)
        *entries = total;
}

// ops 3 / 4: the cooperative addition / doubling of the cold kernels, 4 lanes per point
template <class XY>
__global__ void __launch_bounds__(128)
    k_point_op_coop(int op, const XY* __restrict__ p, const XY* __restrict__ q, XY* __restrict__ out, uint64_t n)
{
    uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, i = gid / kCoop;
    uint32_t sub = (uint32_t)(gid % kCoop);
    if (i >= n)
        return;
    XY acc = p[i];
    if (op == 3)
    {
        XY b = q[i];
        coop_add(acc, b, sub);
    }
    else
        coop_dbl(acc, sub);
    if (sub == 0)
        out[i] = acc;
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
    if (op == 3 || op == 4)
    {
        k_point_op_coop<XY><<<msm_div_up(count * kCoop, 128), 128, 0, st>>>(op, (const XY*)p, (const XY*)q, (XY*)out, count);
        KZP_CUDA_CHECK(cudaGetLastError());
        return;
    }
    k_point_op<XY><<<msm_div_up(count, 128), 128, 0, st>>>(op, (const XY*)p, q, (XY*)out, count);
    KZP_CUDA_CHECK(cudaGetLastError());
}

} // namespace kzp
