// Host-visible interface of the CUDA side (kernels.cu). Plain C++ types and raw device pointers;
// the callers are prover.cu (proof pipeline) and capi.cu (component-level C-ABI entry points).
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

#include <cuda_runtime.h>

#include "ec.cuh"

namespace kzp
{

struct CudaError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

#define KZP_CUDA_CHECK(expr)                                                                      \
    do                                                                                            \
    {                                                                                             \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw ::kzp::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
    } while (0)

// ------------------------------------------------------------------ NTT over Fr
// Domain of size n = 2^log_n. Inverse transforms run decimation-in-frequency (natural in, bit-reversed
// out), forward transforms decimation-in-time (bit-reversed in, natural out), so the reference's
// ifft -> coset shift -> fft chain (groth16.cpp:172-262) needs no permutation pass at all.
struct NttDomain
{
    uint32_t log_n    = 0;
    Fr*      tw_fwd   = nullptr; // w_n^i,  i < n/2   (Montgomery)
    Fr*      tw_inv   = nullptr; // w_n^-i, i < n/2
    Fr*      coset_br = nullptr; // w_2n^bitrev(p) / n at position p  (ifft scale + coset shift fused)
    Fr*      ninv_br  = nullptr; // unused placeholder for plain ifft scaling (1/n is a constant)
    Fr       n_inv;              // 1/n (Montgomery)
};

void ntt_domain_create(NttDomain& d, uint32_t log_n);
void ntt_domain_destroy(NttDomain& d);

// x: n elements, Montgomery, in place.
//   ntt_inverse_dif : natural order in -> bit-reversed out, UNSCALED unless post != nullptr, in which case
//                     position p is multiplied by post[p] in the last stage.
//   ntt_forward_dit : bit-reversed in -> natural out.
void ntt_inverse_dif(const NttDomain& d, Fr* x, const Fr* post, cudaStream_t st);
void ntt_forward_dit(const NttDomain& d, Fr* x, cudaStream_t st);
// natural <-> bit-reversed permutation (only used by the component-level entry points that expose
// the reference's natural-in/natural-out FFT::fft / FFT::ifft contract, fft.cpp:192-246)
void ntt_bitrev_permute(Fr* x, uint32_t log_n, cudaStream_t st);
void fr_scale(Fr* x, uint64_t n, const Fr& k, cudaStream_t st);

// ------------------------------------------------------------------ SpMV (groth16.cpp:125-167)
struct CoefCsr
{
    uint32_t  n_rows  = 0;       // = domain size
    uint64_t  nnz     = 0;
    uint32_t* row_ptr = nullptr; // 2*n_rows + 1 : [2*row + m] .. entries of matrix m (0 = A, 1 = B) of that row
    uint32_t* wire    = nullptr; // nnz
    Fr*       coef    = nullptr; // nnz, value * R^2 mod r exactly as stored in zkey section 4
};
// a = A.w, b = B.w (Montgomery), c = a o b. w: raw canonical witness values.
void spmv_abc(const CoefCsr& m, const Fr* w, Fr* a, Fr* b, Fr* c, cudaStream_t st);
// h[i] = fromMontgomery(a[i]*b[i] - c[i])  (groth16.cpp:266-275)
void h_pointwise(const Fr* a, const Fr* b, const Fr* c, Fr* h, uint64_t n, cudaStream_t st);

// ------------------------------------------------------------------ MSM
// Signed 16-bit windows, 16 windows, and a per-key table of 2^(16 j) * P_i so that all windows share
// ONE bucket set (2^15 buckets): bucket reduction and window combination happen once per MSM.
constexpr int      kMsmWindowBits = 16;
constexpr int      kMsmWindows    = 16;
constexpr uint32_t kMsmBuckets    = 1u << (kMsmWindowBits - 1); // bucket ids 1..kMsmBuckets
constexpr uint32_t kMsmChunk      = 32;                          // sorted entries per accumulate thread

template <class XY>
struct MsmBases
{
    typedef typename XY::Affine Affine;
    uint32_t  n          = 0;       // active (non-infinity) bases
    uint32_t* scalar_idx = nullptr; // n : index of each active base's scalar in the scalar vector
    Affine*   table      = nullptr; // kMsmWindows x n affine points: table[j*n + i] = 2^(16 j) * P_i
};

template <class XY>
struct MsmScratch
{
    uint32_t  cap_entries = 0;
    uint32_t  chunk       = kMsmChunk; // sorted entries per accumulate thread for this MSM
    uint32_t* heavy       = nullptr;   // [0] = number of heavy buckets, [1..] their ids
    uint32_t* counts      = nullptr; // kMsmBuckets + 2
    uint32_t* offsets     = nullptr; // kMsmBuckets + 2 (offsets[b] = first entry of bucket b; [B+1] = total)
    uint32_t* cursor      = nullptr; // kMsmBuckets + 2
    uint32_t* sorted      = nullptr; // cap_entries
    XY*       records     = nullptr; // cap_entries / kMsmChunk + kMsmBuckets + 2
    XY*       buckets     = nullptr; // kMsmBuckets + 1
    XY*       partial     = nullptr; // 2 * (kMsmBuckets / 256)
    XY*       result      = nullptr; // 1 (device)
    cudaEvent_t ev_acc0 = nullptr, ev_acc1 = nullptr; // bracket the bucket-accumulation kernel of the last run
};

// bases_host: n_total affine Montgomery points exactly as in the zkey (64 B G1 / 128 B G2), (0,0) = infinity.
// Active bases are the non-infinity ones in [first, first+count) — base k takes scalar index scalar_offset + k.
template <class XY>
void msm_bases_create(MsmBases<XY>& out, const uint8_t* bases_host, uint64_t first, uint64_t count,
                      uint32_t scalar_offset, cudaStream_t st);
template <class XY>
void msm_bases_destroy(MsmBases<XY>& b);
template <class XY>
void msm_scratch_create(MsmScratch<XY>& s, uint32_t n_active);
template <class XY>
void msm_scratch_destroy(MsmScratch<XY>& s);
// scalars: device array of 32-byte little-endian integers (canonical or not; reduced mod r on the fly).
// Leaves the XYZZ result in s.result (device).
template <class XY>
void msm_run(const MsmBases<XY>& b, MsmScratch<XY>& s, const uint32_t* scalars, cudaStream_t st);

// duration of the last bucket-accumulation launch (ms, CUDA events on its stream) and its number of sorted entries
template <class XY>
void msm_last_accumulate(const MsmScratch<XY>& s, float* ms, uint64_t* entries);

extern template struct MsmBases<G1Xyzz>;
extern template struct MsmBases<G2Xyzz>;

// ------------------------------------------------------------------ diagnostics / tests
// elementwise field kernels: field 0 = Fr, 1 = Fq, 2 = Fq2; op: 0 mul 1 add 2 sub 3 neg 4 toMont 5 fromMont
// 6 sqr 7 inv. Device pointers.
void field_op(int field, int op, const void* a, const void* b, void* out, uint64_t count,
              cudaStream_t st);
// out[i] = op(p[i], q[i]) on points; group 0 = G1, 1 = G2; op: 0 madd (xyzz += affine), 1 add (xyzz += xyzz),
// 2 dbl. p, out XYZZ arrays; q affine (op 0) or XYZZ (op 1).
void point_op(int group, int op, const void* p, const void* q, void* out, uint64_t count,
              cudaStream_t st);
void point_op_g1(int op, const void* p, const void* q, void* out, uint64_t count, cudaStream_t st);
void point_op_g2(int op, const void* p, const void* q, void* out, uint64_t count, cudaStream_t st);
// dependent-free IMAD throughput probe: returns multiply-adds executed; time it with events.
uint64_t imad_probe(uint32_t* sink, int iters, cudaStream_t st);

} // namespace kzp
