// Shared-memory radix-128 NTT levels for sm_100a, one warp per pair of columns.
//
// A transform of size 2^k is split into levels of 7 stages. One level applied to index bits [lo, lo+7) is a plain
// 128-point DFT along those bits (twiddles are powers of w_128 only) plus one "boundary" twiddle per element that
// connects it to the remaining levels (the classic four-step factorisation, applied recursively):
//   DIF (inverse direction here, natural in -> bit-reversed out): levels from the top bits down, boundary twiddle
//        w^( (pos mod 2^lo) * bitrev7(t) * 2^(k-hi) ) applied on the way out;
//   DIT (forward, bit-reversed in -> natural out): levels from the low bits up, boundary twiddle
//        w^( bitrev_{k-lo}(pos >> lo) * ((pos >> plo) mod 2^(lo-plo)) * 2^plo ) applied on the way in.
// Stages left over when k is not a multiple of 7 run as plain radix-2 stages (kernels.cu), which compose with the
// levels because both are exact factorisations of the same DFT (validated stage by stage in tests).
//
// Work split: a "column" is the 128 elements t = 0..127 of one level-DFT. A WARP owns two adjacent columns
// (2 x 128 elements, 8 KiB of shared memory); lane = (column e, row group g), 8 elements per thread. The 7 stages
// run as three register rounds over 8 rows each:
//   R1  rows g + 16 q        stages with half-distance 64, 32     (8 products)
//   R2  rows 32 (g>>2) + (g&3) + 4 q   stages 16, 8               (8 products)
//   R3  rows 8 g + q         stages 4, 2, 1                       (5 products: the twiddles of these stages depend
//                                                                  on q only, so w^0 = 1 is skipped at compile time)
// i.e. 21 products per thread and level plus the 8 boundary twiddles, every warp the same amount. The first round
// loads its rows straight from global memory and the last one stores straight to it (each warp instruction touches
// 16 rows x 64 contiguous bytes), so an element crosses shared memory twice per level, between rounds, and only
// __syncwarp is needed: warps of a CTA drift apart and their memory phases overlap the others' arithmetic.
// Shared-memory slot of row t: the low three bits of t are XORed with bits 3..5, which makes the 8 lanes of every
// quarter-warp hit 8 different 16-byte bank groups in all three rounds; elements are split in two 16-byte planes.
// Replaces FFT<Fr>::fft / ifft (rust-rapidsnark/rapidsnark/src/fft.cpp:192-246) together with kernels.cu.
#pragma once

#include <cuda.h> // CUtensorMap (the type only: the encoder is looked up in the driver at run time)

#include "device.hpp"

namespace kzp
{

constexpr int kNttTileBits  = 7;
constexpr int kNttColBits   = 4;                  // columns per CTA: 16 = 8 warps x 2 columns
constexpr int kNttTileCols  = 1 << kNttColBits;
constexpr int kNttThreads   = 16 * kNttTileCols;  // every thread owns 8 elements
constexpr int kNttTileElems = kNttTileCols << kNttTileBits;
constexpr int kNttMinCtas   = 512 / kNttThreads;  // 128 registers per thread
constexpr int kNttMaxBatch = 3;

// vectors transformed by one launch (blockIdx.y selects): the prover runs a, b and c through every level together
struct NttBatch
{
    Fr* x[kNttMaxBatch];
};

// 16-byte slot of (row t, column e of the warp) inside the warp's plane of 256 slots
__device__ __forceinline__ uint32_t ntt_slot(uint32_t t, uint32_t e) { return (e << 7) | (t & ~7u) | ((t ^ (t >> 3)) & 7u); }

__device__ __forceinline__ void ntt_sm_store(uint4* wsm, uint32_t t, uint32_t e, const Fr& v)
{
    uint32_t p   = ntt_slot(t, e);
    wsm[p]       = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    wsm[256 + p] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

__device__ __forceinline__ void ntt_sm_load(const uint4* wsm, uint32_t t, uint32_t e, Fr& v)
{
    uint32_t p  = ntt_slot(t, e);
    uint4    lo = wsm[p], hi = wsm[256 + p];
    v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
    v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
}

// twiddles w_128^j, j < 64, as two planes of 16 bytes (consecutive j 16 bytes apart: no bank conflicts)
__device__ __forceinline__ void ntt_tw_load(const uint4* twp, uint32_t j, Fr& v)
{
    uint4 lo = twp[j], hi = twp[64 + j];
    v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
    v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
}

__device__ __forceinline__ void ntt_tw_fill(uint4* twp, const Fr* __restrict__ tw, uint32_t j, uint32_t k)
{
    const uint4* src = reinterpret_cast<const uint4*>(tw + ((size_t)j << (k - kNttTileBits)));
    twp[j]           = src[0];
    twp[64 + j]      = src[1];
}

// w^e for e < 2^k from the half table tw[i] = w^i, i < 2^(k-1)   (w^(2^(k-1)) = -1)
__device__ __forceinline__ void ntt_root(Fr& r, const Fr* __restrict__ tw, uint32_t e, uint32_t k)
{
    uint32_t half = 1u << (k - 1);
    if (e < half)
        r = tw[e];
    else
    {
        Fr t = tw[e - half];
        Fr::neg(r, t);
    }
}

// ---- bulk-async staging (sm_90+ async proxy: cp.async.bulk / TMA completing on an mbarrier) ------------------------
// A warp's next tile is fetched into its own shared-memory buffer by the copy engine while the warp is still in the
// arithmetic of the current one: one elected lane arms the warp's mbarrier with the byte count and issues the copy,
// all lanes wait on the barrier's phase before the first round reads. No registers, no LSU issue slots, no address
// arithmetic are spent on the loads, and their latency sits under the last round's products and stores.
__device__ __forceinline__ uint32_t ntt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ntt_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ntt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ntt_mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void ntt_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ntt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ntt_mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done, addr = ntt_smem_u32(bar), spins = 0;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (!done && ++spins > (1u << 24)) // a copy that never lands (bad descriptor) must fail the launch, not hang the GPU
            asm volatile("trap;");
    } while (!done);
}
// generic-proxy accesses of this thread to shared memory are ordered before later async-proxy (copy engine) accesses
__device__ __forceinline__ void ntt_fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// contiguous global -> shared copy by the copy engine (SASS: UBLKCP); bytes and both addresses multiples of 16
__device__ __forceinline__ void ntt_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ntt_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(ntt_smem_u32(bar))
                 : "memory");
}

// 3-D tiled tensor copy global -> shared through a tensor map (SASS: UTMALDG); coordinates innermost first
__device__ __forceinline__ void ntt_tma_load_3d(void* dst_smem, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     ntt_smem_u32(dst_smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(ntt_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void ntt_tma_load_4d(void* dst_smem, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     ntt_smem_u32(dst_smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(ntt_smem_u32(bar))
                 : "memory");
}

// One butterfly stage on v[0..7]: pairs (q, q | 1 << B) for the q with bit B clear. The twiddle of pair q is
// w_128^idx with idx = (j0 + ((q & mask) << qsh)) << ish where mask = (1 << B) - 1, i.e. j = row mod half-distance.
// DIF: (a, b) -> (a + b, (a - b) w);  DIT: (a, b) -> (a + w b, a - w b).  KNOWN0: j0 == 0 at compile time, so
// pairs with (q & mask) == 0 have twiddle 1 and skip the product.
// Lazy reduction: inside a level every value lives in [0, 2p) (ff.cuh: BN254's r leaves two spare bits). Products skip
// their final conditional subtraction, and the DIF difference feeds its product as a + 2p - b without any comparison
// (below 4p, times a canonical twiddle: the product is below 2p again). A level's inputs are canonical and its outputs
// are made canonical before they are stored (reduce_2p), so nothing outside the level sees the wider range.
template <bool DIT, int B, bool KNOWN0>
__device__ __forceinline__ void ntt_stage(Fr (&v)[8], const uint4* twp, uint32_t j0, uint32_t qsh, uint32_t ish)
{
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        if (q & (1 << B))
            continue;
        const int q2 = q | (1 << B);
        const int qm = q & ((1 << B) - 1);
        if (KNOWN0 && qm == 0)
        {
            Fr a = v[q], b = v[q2];
            Fr::add_lazy(v[q], a, b);
            Fr::sub_lazy(v[q2], a, b);
            continue;
        }
        Fr w;
        ntt_tw_load(twp, (j0 + ((uint32_t)qm << qsh)) << ish, w);
        Fr a = v[q], b = v[q2];
        if (DIT)
        {
            Fr::mul_lazy(b, b, w);
            Fr::add_lazy(v[q], a, b);
            Fr::sub_lazy(v[q2], a, b);
        }
        else
        {
            Fr d;
            Fr::add_lazy(v[q], a, b);
            Fr::sub_plus_2p(d, a, b);
            Fr::mul_lazy(v[q2], d, w);
        }
    }
}

// rows of the three rounds for row group g, element q
__device__ __forceinline__ uint32_t ntt_row1(uint32_t g, uint32_t q) { return g + 16u * q; }
__device__ __forceinline__ uint32_t ntt_row2(uint32_t g, uint32_t q) { return 32u * (g >> 2) + (g & 3u) + 4u * q; }
__device__ __forceinline__ uint32_t ntt_row3(uint32_t g, uint32_t q) { return 8u * g + q; }

struct NttNoHook
{
    __device__ __forceinline__ void operator()() const {}
};

// The three rounds of a 128-point DIF (R1, R2, R3) between registers: in: rows of R1, out: rows of R3. smem_free() runs
// once the warp's shared-memory buffer is no longer needed (before the last round's products): the place to start the
// next prefetch.
template <class Hook = NttNoHook>
__device__ __forceinline__ void ntt_dif_rounds(Fr (&v)[8], uint4* wsm, const uint4* twp, uint32_t g, uint32_t e, Hook smem_free = Hook())
{
    // R1: half-distance 64 (q bit 2; j = g + 16 (q & 3)), 32 (q bit 1; j = g + 16 (q & 1), index j * 2)
    ntt_stage<false, 2, false>(v, twp, g, 4, 0);
    ntt_stage<false, 1, false>(v, twp, g, 4, 1);
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_store(wsm, ntt_row1(g, q), e, v[q]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_load(wsm, ntt_row2(g, q), e, v[q]);
    // R2: half-distance 16 (j = (g & 3) + 4 (q & 3), index j * 4), 8 (j = (g & 3) + 4 (q & 1), index j * 8)
    ntt_stage<false, 2, false>(v, twp, g & 3u, 2, 2);
    ntt_stage<false, 1, false>(v, twp, g & 3u, 2, 3);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_store(wsm, ntt_row2(g, q), e, v[q]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_load(wsm, ntt_row3(g, q), e, v[q]);
    __syncwarp();
    smem_free();
    // R3: half-distance 4 (j = q & 3, index j * 16), 2 (j = q & 1, index j * 32), 1 (no twiddle)
    ntt_stage<false, 2, true>(v, twp, 0, 0, 4);
    ntt_stage<false, 1, true>(v, twp, 0, 0, 5);
    ntt_stage<false, 0, true>(v, twp, 0, 0, 6);
}

// The three rounds of a 128-point DIT (R3, R2, R1): in: rows of R3, out: rows of R1. smem_free() runs once the warp's
// shared-memory buffer is no longer needed (before the last round's products): the place to start the next prefetch.
template <class Hook = NttNoHook>
__device__ __forceinline__ void ntt_dit_rounds(Fr (&v)[8], uint4* wsm, const uint4* twp, uint32_t g, uint32_t e, Hook smem_free = Hook())
{
    ntt_stage<true, 0, true>(v, twp, 0, 0, 6);
    ntt_stage<true, 1, true>(v, twp, 0, 0, 5);
    ntt_stage<true, 2, true>(v, twp, 0, 0, 4);
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_store(wsm, ntt_row3(g, q), e, v[q]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_load(wsm, ntt_row2(g, q), e, v[q]);
    ntt_stage<true, 1, false>(v, twp, g & 3u, 2, 3);
    ntt_stage<true, 2, false>(v, twp, g & 3u, 2, 2);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_store(wsm, ntt_row2(g, q), e, v[q]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++)
        ntt_sm_load(wsm, ntt_row1(g, q), e, v[q]);
    __syncwarp();
    smem_free();
    ntt_stage<true, 1, false>(v, twp, g, 4, 1);
    ntt_stage<true, 2, false>(v, twp, g, 4, 0);
}

constexpr size_t kNttWarpSmem  = 2 * 256 * sizeof(uint4);                                   // two planes of 256 slots
constexpr size_t kNttLevelSmem = (kNttThreads / 32) * kNttWarpSmem + 2 * 64 * sizeof(uint4); // + one twiddle table
constexpr size_t kNttMidSmem   = (kNttThreads / 32) * kNttWarpSmem + 4 * 64 * sizeof(uint4) + (kNttThreads / 32) * sizeof(uint64_t); // + two twiddle tables + one mbarrier per warp

// Tile index of this CTA under the route's block partition (device.hpp NttRoute).
__device__ __forceinline__ uint32_t ntt_route_block(const NttRoute& rt, uint32_t b, uint32_t n_tiles)
{
    if (rt.blocks == kNttBlocksLow)
    {
        // column index bits [7-k, 7) = tile index bits [3-k, 3) (a tile is 16 columns): insert g there
        const uint32_t low = (uint32_t)(7 - kNttColBits) - rt.k;
        return ((b >> low) << (7 - kNttColBits)) | (rt.g << low) | (b & ((1u << low) - 1u));
    }
    if (rt.blocks == kNttBlocksTop)
        return rt.g * n_tiles + b;
    return b;
}

// Destination of position pos of vector blockIdx.y under the route's store mode
template <bool ROUTED>
__device__ __forceinline__ Fr* ntt_route_dst(const NttRoute& rt, Fr* local, uint32_t pos, uint32_t vec)
{
    if (!ROUTED)
        return local + pos;
    int r = 0;
    if (rt.store == kNttStoreBits)
        r = (int)((pos >> rt.shift) & rt.mask);
    else
        while (r + 1 < rt.world && pos >= rt.bound[r + 1])
            r++;
    return rt.dst[vec][r] + pos;
}

// One level on bits [lo, lo+7) of a size-2^k transform. tw: w^i (forward) or w^-i (inverse), i < 2^(k-1).
// post (DIF only, may be null): element at position pos is multiplied by post[pos] on the way out.
// ROUTED: the outputs are stored where rt says (peer stores); the tile selection of rt applies either way.
template <bool DIT, bool ROUTED = false>
__global__ void __launch_bounds__(kNttThreads, kNttMinCtas)
    k_ntt_level(NttBatch batch, const Fr* __restrict__ tw, uint32_t k, uint32_t lo, uint32_t plo,
                const Fr* __restrict__ post, const NttRoute rt)
{
    Fr* __restrict__        x = batch.x[blockIdx.y];
    extern __shared__ uint4 ntt_smem[];
    const uint32_t          tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4*                  wsm = ntt_smem + warp * (kNttWarpSmem / sizeof(uint4));
    uint4*                  twp = ntt_smem + (kNttThreads / 32) * (kNttWarpSmem / sizeof(uint4));
    const uint32_t          hi  = lo + kNttTileBits;
    if (tid < 64)
        ntt_tw_fill(twp, tw, tid, k);
    __syncthreads();

    const uint32_t e = lane >> 4, g = lane & 15u;
    const uint32_t rest     = ntt_route_block(rt, blockIdx.x, gridDim.x) * kNttTileCols + 2u * warp + e;
    const uint32_t low_mask = (1u << lo) - 1u;
    const uint32_t col_base = ((rest >> lo) << hi) | (rest & low_mask); // position of row 0 of this column
    Fr             v[8];
    if (!DIT)
    {
#pragma unroll
        for (int q = 0; q < 8; q++)
            v[q] = x[col_base | (ntt_row1(g, q) << lo)];
        ntt_dif_rounds(v, wsm, twp, g, e);
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            uint32_t t   = ntt_row3(g, q);
            uint32_t pos = col_base | (t << lo);
            if (lo > 0)
            {
                uint32_t m  = pos & low_mask;
                uint32_t ex = (m * (__brev(t) >> 25)) << (k - hi);
                if (ex != 0)
                {
                    Fr w;
                    ntt_root(w, tw, ex, k);
                    Fr::mul_lazy(v[q], v[q], w);
                }
            }
            if (post)
                Fr::mul_lazy(v[q], v[q], post[pos]);
            Fr::reduce_2p(v[q], v[q]);
            *ntt_route_dst<ROUTED>(rt, x, pos, blockIdx.y) = v[q];
        }
    }
    else
    {
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            uint32_t pos = col_base | (ntt_row3(g, q) << lo);
            v[q]         = x[pos];
            if (lo > 0)
            {
                uint32_t upper = pos >> lo;
                uint32_t brv   = __brev(upper) >> (32 - (k - lo));
                uint32_t kt    = (pos >> plo) & ((1u << (lo - plo)) - 1u);
                uint32_t ex    = (brv * kt) << plo;
                if (ex != 0)
                {
                    Fr w;
                    ntt_root(w, tw, ex, k);
                    Fr::mul_lazy(v[q], v[q], w);
                }
            }
        }
        ntt_dit_rounds(v, wsm, twp, g, e);
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            Fr::reduce_2p(v[q], v[q]);
            *ntt_route_dst<ROUTED>(rt, x, col_base | (ntt_row1(g, q) << lo), blockIdx.y) = v[q];
        }
    }
}

// ---- levels with lo >= 1: persistent warps, tiles staged by TMA -----------------------------------------------------
// The vector is described to the copy engine as a 3-D tensor of 8-byte words (NttMaps, kernels.cu):
//     dim 0 = 2^lo * 4 words (the contiguous low part of the position), dim 1 = 128 rows t (stride 2^lo elements),
//     dim 2 = the upper part of the position (stride 2^hi elements), 64-byte swizzle.
// A warp's two columns are the box {8 words, 128 rows, 1}: 128 rows of 64 bytes, 8 KiB, one tensor copy (DIF). The
// first round wants lane g to hold rows g + 16 q (DIF) or 8 g + q (DIT); so that the lanes of a quarter-warp always
// read CONSECUTIVE 64-byte rows of the landed image (with the 64-byte swizzle: eight different bank groups), the DIT
// kernel sees the rows as two dimensions {q: 8, g: 16} (a 4-D tensor) and fetches the tile as eight copies of the box
// {8 words, 1, 16, 1}, copy q (rows q, q + 8, ..., q + 120) landing at q * 1 KiB: in both cases virtual row g + 16 q of
// the image is the element the lane needs as v[q].
struct alignas(64) NttMaps
{
    CUtensorMap m[kNttMaxBatch];
};

// v[q] <- element (virtual row g + 16 q, column e) of a landed tile (64-byte rows, 64-byte swizzle: the 16-byte chunk
// index is XORed with bits 1..2 of the row)
__device__ __forceinline__ void ntt_landed_load(const uint4* wsm, uint32_t g, uint32_t e, Fr (&v)[8])
{
    const uint32_t sw = (g >> 1) & 3u; // (row >> 1) & 3 with row = g + 16 q
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        const uint4* row = wsm + (g + 16u * q) * 4u;
        uint4        lo = row[(2u * e) ^ sw], hi = row[(2u * e + 1u) ^ sw];
        v[q].v[0] = lo.x; v[q].v[1] = lo.y; v[q].v[2] = lo.z; v[q].v[3] = lo.w;
        v[q].v[4] = hi.x; v[q].v[5] = hi.y; v[q].v[6] = hi.z; v[q].v[7] = hi.w;
    }
}

constexpr size_t kNttTmaSmem = kNttLevelSmem + (kNttThreads / 32) * sizeof(uint64_t); // + one mbarrier per warp

// One level on bits [lo, lo+7), lo >= 1. Work unit u < n_tiles * count: vector u % count, tile u / count (16 columns).
template <bool DIT, bool ROUTED = false>
__global__ void __launch_bounds__(kNttThreads, kNttMinCtas)
    k_ntt_level_tma(const __grid_constant__ NttMaps maps, NttBatch batch, const Fr* __restrict__ tw, uint32_t k, uint32_t lo, uint32_t plo,
                    const NttRoute rt, uint32_t n_tiles, uint32_t count)
{
    extern __shared__ __align__(1024) uint4 ntt_smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4*         wsm = ntt_smem + warp * (kNttWarpSmem / sizeof(uint4));
    uint4*         twp = ntt_smem + (kNttThreads / 32) * (kNttWarpSmem / sizeof(uint4));
    uint64_t*      bar = reinterpret_cast<uint64_t*>(twp + 128) + warp;
    const uint32_t hi  = lo + kNttTileBits;
    if (tid < 64)
        ntt_tw_fill(twp, tw, tid, k);
    if (lane == 0)
    {
        ntt_mbar_init(bar, 1);
        ntt_mbar_init_fence();
    }
    __syncthreads();

    const uint32_t e = lane >> 4, g = lane & 15u;
    const uint32_t low_mask = (1u << lo) - 1u;
    const uint32_t units    = n_tiles * count;
    // column index of the warp's first column in work unit u
    auto pair_rest = [&](uint32_t u) { return ntt_route_block(rt, u / count, n_tiles) * kNttTileCols + 2u * warp; };
    auto prefetch  = [&](uint32_t u) {
        const uint32_t     rest = pair_rest(u);
        const CUtensorMap* map  = &maps.m[u % count];
        const uint32_t     c0 = (rest & low_mask) * 4u, c2 = rest >> lo;
        if (lane == 0)
            ntt_mbar_expect_tx(bar, (uint32_t)kNttWarpSmem);
        if (!DIT)
        {
            if (lane == 0)
                ntt_tma_load_3d(wsm, map, c0, 0, c2, bar);
        }
        else
        {
            __syncwarp();
            if (lane < 8)
                ntt_tma_load_4d(wsm + lane * 64u, map, c0, lane, 0, c2, bar); // rows lane, lane + 8, ..., lane + 120
        }
    };
    uint32_t u = blockIdx.x, parity = 0;
    if (u < units)
        prefetch(u);
    for (; u < units; u += gridDim.x)
    {
        Fr* __restrict__ x        = batch.x[u % count];
        const uint32_t   rest     = pair_rest(u) + e;
        const uint32_t   col_base = ((rest >> lo) << hi) | (rest & low_mask); // position of row 0 of this column
        const uint32_t   next     = u + gridDim.x;
        auto             smem_free = [&] {
            if (next < units)
            {
                ntt_fence_async_proxy(); // this lane's generic-proxy accesses to the buffer, before the copy engine's
                __syncwarp();
                prefetch(next);
            }
        };
        __syncthreads(); // keeps the CTA's warps on the same stretch of the (large, unrolled) code: see k_ntt_mid
        ntt_mbar_wait(bar, parity);
        parity ^= 1u;
        Fr v[8];
        ntt_landed_load(wsm, g, e, v);
        __syncwarp(); // every lane has its rows before the rounds start to exchange through the same buffer
        if (!DIT)
        {
            ntt_dif_rounds(v, wsm, twp, g, e, smem_free);
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                uint32_t t   = ntt_row3(g, q);
                uint32_t pos = col_base | (t << lo);
                uint32_t m   = pos & low_mask;
                uint32_t ex  = (m * (__brev(t) >> 25)) << (k - hi);
                if (ex != 0)
                {
                    Fr w;
                    ntt_root(w, tw, ex, k);
                    Fr::mul_lazy(v[q], v[q], w);
                }
                Fr::reduce_2p(v[q], v[q]);
                *ntt_route_dst<ROUTED>(rt, x, pos, u % count) = v[q];
            }
        }
        else
        {
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                uint32_t pos   = col_base | (ntt_row3(g, q) << lo);
                uint32_t upper = pos >> lo;
                uint32_t brv   = __brev(upper) >> (32 - (k - lo));
                uint32_t kt    = (pos >> plo) & ((1u << (lo - plo)) - 1u);
                uint32_t ex    = (brv * kt) << plo;
                if (ex != 0)
                {
                    Fr w;
                    ntt_root(w, tw, ex, k);
                    Fr::mul_lazy(v[q], v[q], w);
                }
            }
            ntt_dit_rounds(v, wsm, twp, g, e, smem_free);
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                Fr::reduce_2p(v[q], v[q]);
                *ntt_route_dst<ROUTED>(rt, x, col_base | (ntt_row1(g, q) << lo), u % count) = v[q];
            }
        }
    }
}

// Fused middle of the prover's ifft -> coset shift -> fft chain when k is a multiple of 7: the inverse transform's
// last level and the forward transform's first level both act on the same 128 contiguous elements (lo = 0), so one
// warp runs the DIF rounds, multiplies by post[pos] (= w_2n^bitrev(pos) / n: ifft scaling + coset shift) and
// continues with the DIT rounds. R3 is the last DIF round and the first DIT round, so the hand-over happens in
// registers. Saves one full read + write of the vector and a shared-memory pass per chain.
// Persistent warps with bulk-async staging: a warp's two columns are 8 KiB of CONTIGUOUS global memory, fetched by one
// cp.async.bulk into the warp's shared-memory buffer (the same buffer the rounds exchange through: it is free again
// once the last round's rows are in registers, which is when the copy for the warp's next tile is issued).
// Work unit u < n_tiles * count: vector u % count, tile u / count; CTA c takes u = c, c + gridDim.x, ...
template <bool ROUTED = false>
__global__ void __launch_bounds__(kNttThreads, kNttMinCtas)
    k_ntt_mid(NttBatch batch, const Fr* __restrict__ tw_inv, const Fr* __restrict__ tw_fwd, uint32_t k,
              const Fr* __restrict__ post, const NttRoute rt, uint32_t n_tiles, uint32_t count)
{
    extern __shared__ uint4 ntt_smem[];
    const uint32_t          tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4*                  wsm = ntt_smem + warp * (kNttWarpSmem / sizeof(uint4));
    uint4*                  twI = ntt_smem + (kNttThreads / 32) * (kNttWarpSmem / sizeof(uint4));
    uint4*                  twF = twI + 128;
    uint64_t*               bar = reinterpret_cast<uint64_t*>(twF + 128) + warp;
    if (tid < 64)
        ntt_tw_fill(twI, tw_inv, tid, k);
    else if (tid < 128)
        ntt_tw_fill(twF, tw_fwd, tid - 64, k);
    if (lane == 0)
    {
        ntt_mbar_init(bar, 1);
        ntt_mbar_init_fence();
    }
    __syncthreads();
    const uint32_t e = lane >> 4, g = lane & 15u;
    const uint32_t units = n_tiles * count;
    // first element of this warp's column pair in work unit u
    auto pair_base = [&](uint32_t u) { return (ntt_route_block(rt, u / count, n_tiles) * kNttTileCols + 2u * warp) << kNttTileBits; };
    auto prefetch  = [&](uint32_t u) {
        if (lane == 0)
        {
            ntt_mbar_expect_tx(bar, (uint32_t)kNttWarpSmem);
            ntt_bulk_load(wsm, batch.x[u % count] + pair_base(u), (uint32_t)kNttWarpSmem, bar);
        }
    };
    uint32_t u = blockIdx.x, parity = 0;
    if (u < units)
        prefetch(u);
    for (; u < units; u += gridDim.x)
    {
        Fr* __restrict__ x        = batch.x[u % count];
        const uint32_t   col_base = pair_base(u) + (e << kNttTileBits); // 128 contiguous elements
        // The unrolled rounds are ~370 KB of code: warps that drift apart over the tiles of a persistent CTA evict one
        // another's instructions (ncu: 3.7 no-instruction stall cycles per issue without this barrier). One CTA barrier
        // per tile keeps the eight warps on the same stretch of code.
        __syncthreads();
        ntt_mbar_wait(bar, parity);
        parity ^= 1u;
        Fr v[8];
        {
            // landed layout: element (column e, row t) at 16-byte slot e * 256 + 2 t (+1). Quarter-warps read rows
            // t, t + 1, .. t + 7: lanes 4..7 take the upper half first, so that the eight 16-byte accesses of one
            // instruction fall into eight different bank groups.
            const uint32_t h = (g >> 2) & 1u;
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                const uint4* src = wsm + e * 256u + 2u * ntt_row1(g, q);
                uint4        a = src[h], b = src[h ^ 1u];
                uint4        lo = h ? b : a, hi = h ? a : b;
                v[q].v[0] = lo.x; v[q].v[1] = lo.y; v[q].v[2] = lo.z; v[q].v[3] = lo.w;
                v[q].v[4] = hi.x; v[q].v[5] = hi.y; v[q].v[6] = hi.z; v[q].v[7] = hi.w;
            }
        }
        __syncwarp(); // every lane has its rows before the rounds start to exchange through the same buffer
        ntt_dif_rounds(v, wsm, twI, g, e);
        const Fr* pp = post + col_base + ntt_row3(g, 0);
#pragma unroll
        for (int q = 0; q < 8; q++)
            Fr::mul_lazy(v[q], v[q], pp[q]);
        __syncwarp();
        const uint32_t next = u + gridDim.x;
        ntt_dit_rounds(v, wsm, twF, g, e, [&] {
            if (next < units)
            {
                ntt_fence_async_proxy(); // this lane's generic-proxy accesses to the buffer, before the copy engine's
                __syncwarp();
                prefetch(next);
            }
        });
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            Fr::reduce_2p(v[q], v[q]);
            *ntt_route_dst<ROUTED>(rt, x, col_base + ntt_row1(g, q), u % count) = v[q];
        }
    }
}

} // namespace kzp
