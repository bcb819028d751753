// Shared-memory radix-128 NTT levels for sm_100a.
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
// One CTA = 256 threads = 16 tiles x 128 points = 2048 elements = 64 KiB of shared memory, stored as two planes of
// uint4 with the column index rotated by the row (phys = 16 t + ((c + t) & 15)) so that both the column-fastest
// accesses of the butterfly rounds and the row-fastest accesses of the contiguous level are bank-conflict free.
// Global traffic is coalesced 128-bit accesses in runs of >= 512 B; every element is read and written once per
// level. Each thread runs radix-8 butterflies on 8 elements held in registers (rounds of 3, 3 and 1 stages).
// Replaces FFT<Fr>::fft / ifft (rust-rapidsnark/rapidsnark/src/fft.cpp:192-246) together with kernels.cu.
#pragma once

#include "device.hpp"

namespace kzp
{

constexpr int kNttTileBits  = 7;
constexpr int kNttColBits   = 4;                  // columns per CTA: 16 -> 256 threads, 64 KiB, half the register file.
                                                  // Measured with 8 (128 threads): the transform alone is 4 % faster
                                                  // (0.95 vs 0.99 ms at 2^21) but the whole proof is 0.25 ms slower
constexpr int kNttTileCols  = 1 << kNttColBits;
constexpr int kNttThreads   = 16 * kNttTileCols;  // every thread owns 8 elements
constexpr int kNttTileElems = kNttTileCols << kNttTileBits;
constexpr int kNttMinCtas   = 512 / kNttThreads;  // 128 registers per thread either way
constexpr int kNttMaxBatch = 3;

// vectors transformed by one launch (blockIdx.y selects): the prover runs a, b and c through every level together
struct NttBatch
{
    Fr* x[kNttMaxBatch];
};

__device__ __forceinline__ uint32_t ntt_phys(uint32_t t, uint32_t c) { return (uint32_t)kNttTileCols * t + ((c + t) & (uint32_t)(kNttTileCols - 1)); }

__device__ __forceinline__ void ntt_sm_store(uint4* sm, uint32_t t, uint32_t c, const Fr& v)
{
    uint32_t p   = ntt_phys(t, c);
    sm[p]        = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    sm[kNttTileElems + p] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

__device__ __forceinline__ void ntt_sm_load(const uint4* sm, uint32_t t, uint32_t c, Fr& v)
{
    uint32_t p  = ntt_phys(t, c);
    uint4    lo = sm[p], hi = sm[kNttTileElems + p];
    v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
    v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
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

// Radix-8 butterflies on v[0..7] = elements t = t_rest + (q << sh), q = 0..7, covering the stages with half-distance
// 2^(sh+2), 2^(sh+1), 2^sh (DIF order) or the reverse (DIT order). nst = number of stages (3, or 1 for the last round,
// where only the half-distance-2^sh stage runs on pairs (q, q+1)). twT[j] = w_128^j, j < 64.
template <bool DIT>
__device__ __forceinline__ void ntt_round(Fr (&v)[8], const Fr* twT, uint32_t t_low, uint32_t sh, int nst)
{
    if (nst == 1)
    {
        // single stage, half-distance 2^sh with sh == 0: twiddle index j = t mod 1 = 0 -> no multiplication
#pragma unroll
        for (int q = 0; q < 8; q += 2)
        {
            Fr u = v[q], w = v[q + 1];
            Fr::add(v[q], u, w);
            Fr::sub(v[q + 1], u, w);
        }
        return;
    }
    if (!DIT)
    {
#pragma unroll
        for (int u = 2; u >= 0; u--)
        {
            // stage h = 2^(sh+u): pairs differ in bit u of q; j = t mod h = t_low + ((q & (2^u - 1)) << sh)
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                if (q & (1 << u))
                    continue;
                int      q2  = q | (1 << u);
                uint32_t j   = t_low + ((uint32_t)(q & ((1 << u) - 1)) << sh);
                uint32_t idx = j << (6 - sh - u);
                Fr       a = v[q], b = v[q2], d;
                Fr::add(v[q], a, b);
                Fr::sub(d, a, b);
                if (idx != 0)
                    Fr::mul(v[q2], d, twT[idx]);
                else
                    v[q2] = d;
            }
        }
    }
    else
    {
#pragma unroll
        for (int u = 0; u <= 2; u++)
        {
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                if (q & (1 << u))
                    continue;
                int      q2  = q | (1 << u);
                uint32_t j   = t_low + ((uint32_t)(q & ((1 << u) - 1)) << sh);
                uint32_t idx = j << (6 - sh - u);
                Fr       a = v[q], b = v[q2];
                if (idx != 0)
                    Fr::mul(b, b, twT[idx]);
                Fr::add(v[q], a, b);
                Fr::sub(v[q2], a, b);
            }
        }
    }
}

// One level on bits [lo, lo+7) of a size-2^k transform. tw: w^i (forward) or w^-i (inverse), i < 2^(k-1).
// post (DIF only, may be null): element at position pos is multiplied by post[pos] on the way out.
template <bool DIT>
__global__ void __launch_bounds__(kNttThreads, kNttMinCtas)
    k_ntt_level(NttBatch batch, const Fr* __restrict__ tw, uint32_t k, uint32_t lo, uint32_t plo,
                const Fr* __restrict__ post)
{
    Fr* __restrict__        x = batch.x[blockIdx.y];
    extern __shared__ uint4 ntt_smem[];
    uint4*                  sm  = ntt_smem;                                   // 2 planes x 2048 uint4
    Fr*                     twT = reinterpret_cast<Fr*>(ntt_smem + 2 * kNttTileElems); // 64 roots of order 128
    const uint32_t          tid = threadIdx.x;
    const uint32_t          hi  = lo + kNttTileBits;
    if (tid < 64)
        twT[tid] = tw[(size_t)tid << (k - kNttTileBits)];

    const uint32_t rest0    = blockIdx.x * kNttTileCols;
    const uint32_t low_mask = (1u << lo) - 1u;
    // element (t, c) of this CTA lives at global position pos(t, c)
    auto pos_of = [&](uint32_t t, uint32_t c) -> uint32_t {
        uint32_t rest = rest0 + c;
        return ((rest >> lo) << hi) | (t << lo) | (rest & low_mask);
    };

    // ---- load (coalesced), boundary twiddle for DIT
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        uint32_t e = tid + (uint32_t)kNttThreads * q;
        uint32_t t, c;
        if (lo == 0)
        {
            c = e >> 7;
            t = e & 127u;
        }
        else
        {
            t = e >> kNttColBits;
            c = e & (uint32_t)(kNttTileCols - 1);
        }
        uint32_t pos = pos_of(t, c);
        Fr       val = x[pos];
        if (DIT && lo > 0)
        {
            uint32_t upper = pos >> lo;
            uint32_t brv   = __brev(upper) >> (32 - (k - lo));
            uint32_t kt    = (pos >> plo) & ((1u << (lo - plo)) - 1u);
            uint32_t ex    = (brv * kt) << plo;
            if (ex != 0)
            {
                Fr w;
                ntt_root(w, tw, ex, k);
                Fr::mul(val, val, w);
            }
        }
        ntt_sm_store(sm, t, c, val);
    }
    __syncthreads();

    // ---- three rounds of radix-8 butterflies in registers
    const uint32_t c = tid & (uint32_t)(kNttTileCols - 1);
    const uint32_t g = tid >> kNttColBits; // 0..15
    Fr             v[8];
#pragma unroll 1
    for (int r = 0; r < 3; r++)
    {
        int      round = DIT ? 2 - r : r; // DIF: A, B, C ; DIT: C, B, A
        uint32_t t_rest, sh;
        int      nst;
        if (round == 0)
        {
            t_rest = g;
            sh     = 4;
            nst    = 3;
        }
        else if (round == 1)
        {
            t_rest = (g >> 1) * 16u + (g & 1u);
            sh     = 1;
            nst    = 3;
        }
        else
        {
            t_rest = g * 8u;
            sh     = 0;
            nst    = 1;
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_load(sm, t_rest + ((uint32_t)q << sh), c, v[q]);
        ntt_round<DIT>(v, twT, t_rest & ((1u << sh) - 1u), sh, nst);
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_store(sm, t_rest + ((uint32_t)q << sh), c, v[q]);
        __syncthreads();
    }

    // ---- store (coalesced), boundary twiddle for DIF, optional pointwise post-multiplier
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        uint32_t e = tid + (uint32_t)kNttThreads * q;
        uint32_t t, cc;
        if (lo == 0)
        {
            cc = e >> 7;
            t  = e & 127u;
        }
        else
        {
            t  = e >> kNttColBits;
            cc = e & (uint32_t)(kNttTileCols - 1);
        }
        uint32_t pos = pos_of(t, cc);
        Fr       val;
        ntt_sm_load(sm, t, cc, val);
        if (!DIT && lo > 0)
        {
            uint32_t m  = pos & low_mask;
            uint32_t ex = (m * (__brev(t) >> 25)) << (k - hi);
            if (ex != 0)
            {
                Fr w;
                ntt_root(w, tw, ex, k);
                Fr::mul(val, val, w);
            }
        }
        if (!DIT && post)
            Fr::mul(val, val, post[pos]);
        x[pos] = val;
    }
}


// Fused middle of the prover's ifft -> coset shift -> fft chain when k is a multiple of 7: the inverse transform's
// last level and the forward transform's first level both act on the same 128 contiguous elements (lo = 0), so one
// CTA runs DIF rounds A, B, C, multiplies by post[pos] (= w_2n^bitrev(pos) / n: ifft scaling + coset shift), and
// continues with DIT rounds C, B, A. Round C of both directions uses the same thread -> element map, so the hand-over
// happens in registers. Saves one full read + write of the vector and four shared-memory passes per chain.
__global__ void __launch_bounds__(kNttThreads, kNttMinCtas)
    k_ntt_mid(NttBatch batch, const Fr* __restrict__ tw_inv, const Fr* __restrict__ tw_fwd, uint32_t k,
              const Fr* __restrict__ post)
{
    Fr* __restrict__        x = batch.x[blockIdx.y];
    extern __shared__ uint4 ntt_smem[];
    uint4*                  sm   = ntt_smem;
    Fr*                     twI  = reinterpret_cast<Fr*>(ntt_smem + 2 * kNttTileElems);
    Fr*                     twF  = twI + 64;
    const uint32_t          tid  = threadIdx.x;
    if (tid < 64)
        twI[tid] = tw_inv[(size_t)tid << (k - kNttTileBits)];
    else if (tid < 128)
        twF[tid - 64] = tw_fwd[(size_t)(tid - 64) << (k - kNttTileBits)];
    const uint32_t base = blockIdx.x * (kNttTileCols << kNttTileBits); // 2048 contiguous elements
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        uint32_t e = tid + (uint32_t)kNttThreads * q;
        Fr       val = x[base + e];
        ntt_sm_store(sm, e & 127u, e >> 7, val);
    }
    __syncthreads();
    const uint32_t c = tid & (uint32_t)(kNttTileCols - 1);
    const uint32_t g = tid >> kNttColBits;
    Fr             v[8];
    // DIF rounds A and B through shared memory
#pragma unroll 1
    for (int round = 0; round < 2; round++)
    {
        uint32_t t_rest = round == 0 ? g : (g >> 1) * 16u + (g & 1u);
        uint32_t sh     = round == 0 ? 4u : 1u;
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_load(sm, t_rest + ((uint32_t)q << sh), c, v[q]);
        ntt_round<false>(v, twI, t_rest & ((1u << sh) - 1u), sh, 3);
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_store(sm, t_rest + ((uint32_t)q << sh), c, v[q]);
        __syncthreads();
    }
    // round C of both directions in registers, with the pointwise multiplier in between
    {
        uint32_t t_rest = g * 8u;
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_load(sm, t_rest + (uint32_t)q, c, v[q]);
        ntt_round<false>(v, twI, 0, 0, 1);
        const Fr* pp = post + base + c * 128u + t_rest;
#pragma unroll
        for (int q = 0; q < 8; q++)
            Fr::mul(v[q], v[q], pp[q]);
        ntt_round<true>(v, twF, 0, 0, 1);
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_store(sm, t_rest + (uint32_t)q, c, v[q]);
        __syncthreads();
    }
    // DIT rounds B and A
#pragma unroll 1
    for (int round = 1; round >= 0; round--)
    {
        uint32_t t_rest = round == 0 ? g : (g >> 1) * 16u + (g & 1u);
        uint32_t sh     = round == 0 ? 4u : 1u;
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_load(sm, t_rest + ((uint32_t)q << sh), c, v[q]);
        ntt_round<true>(v, twF, t_rest & ((1u << sh) - 1u), sh, 3);
#pragma unroll
        for (int q = 0; q < 8; q++)
            ntt_sm_store(sm, t_rest + ((uint32_t)q << sh), c, v[q]);
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        uint32_t e = tid + (uint32_t)kNttThreads * q;
        Fr       val;
        ntt_sm_load(sm, e & 127u, e >> 7, val);
        x[base + e] = val;
    }
}

constexpr size_t kNttMidSmem = 2 * kNttTileElems * sizeof(uint4) + 128 * sizeof(Fr);
constexpr size_t kNttLevelSmem = 2 * kNttTileElems * sizeof(uint4) + 64 * sizeof(Fr);

} // namespace kzp
