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
uint32_t ntt_launches(uint32_t log_n); // kernels per transform
// How an NTT kernel of a multi-GPU proof is routed. Every field zero = an ordinary local launch.
//  * blocks: which tiles this GPU runs. kNttBlocksLow: the tiles whose columns have index bits [7-k, 7) equal to g
//    (a partition by LOW position bits: valid for every level with lo >= 7); kNttBlocksTop: the g-th contiguous
//    1/2^k of the tiles (a partition by the TOP position bits: valid for the fused middle level).
//  * stores: where outputs go. kNttStoreBounds: position pos of vector i goes to dst[i][r] + pos for the shard r
//    with bound[r] <= pos < bound[r+1]; kNttStoreBits: r = (pos >> shift) & mask. dst[i][r] is shard r's buffer of
//    vector i (peer-mapped over NVLink, or the launching device's own): the transposes between the two partitions
//    and the final delivery of every shard's H slice are peer stores fused into the butterflies' last round.
constexpr int kNttMaxShards = 8;
enum : int { kNttBlocksAll = 0, kNttBlocksLow = 1, kNttBlocksTop = 2 };
enum : int { kNttStoreLocal = 0, kNttStoreBounds = 1, kNttStoreBits = 2 };
struct NttRoute
{
    Fr*      dst[3][kNttMaxShards];
    uint32_t bound[kNttMaxShards + 1];
    int      world;
    int      store;
    uint32_t shift, mask;
    int      blocks;
    uint32_t k, g; // log2(shards), this shard
};
// the batched path of ntt_coset_chain (the only one that can be routed) applies to this size
bool ntt_chain_is_batched(uint32_t log_n);
// ifft -> coset shift -> fft (groth16.cpp:172-262) on `count` <= 3 vectors, all of them through each launch together;
// returns the number of kernels launched. last_store (optional, batched path only): routing of the last level's
// stores (kNttStoreBounds: one whole chain per GPU, every shard's slice delivered by the last level).
uint32_t ntt_coset_chain(const NttDomain& d, Fr* const* xs, int count, cudaStream_t st, const NttRoute* last_store = nullptr);
// The same chain spread over 2^k GPUs, in three phases with a cross-GPU hand-over after each (the caller orders
// them with events): phase 0 = inverse levels down to lo = 7 under the low-bit partition, the last one storing by top
// bits; phase 1 = fused middle level under the top-bit partition, storing by low bits; phase 2 = forward levels
// under the low-bit partition, the last one storing every position to the shard that owns it (top bits again).
// dst[i][r] = vector i on shard r for all r < 2^k. Returns the number of kernels launched.
uint32_t ntt_coset_chain_phase(const NttDomain& d, Fr* const* xs, int count, cudaStream_t st, int phase, uint32_t k,
                               uint32_t g, Fr* const (*dst)[kNttMaxShards]);
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
// a = A.w, b = B.w (Montgomery), c = a o b. w: raw canonical witness values. which: bit 0 / 1 / 2 = write a / b / c.
// part_k / part_g: only the rows whose index bits [7 - part_k, 7) equal part_g (the low-bit partition of NttRoute).
void spmv_abc(const CoefCsr& m, const Fr* w, Fr* a, Fr* b, Fr* c, cudaStream_t st, uint32_t which = 7, uint32_t part_k = 0,
              uint32_t part_g = 0);
// h[i] = fromMontgomery(a[i]*b[i] - c[i])  (groth16.cpp:266-275)
void h_pointwise(const Fr* a, const Fr* b, const Fr* c, Fr* h, uint64_t n, cudaStream_t st);

// ------------------------------------------------------------------ MSM
// Signed c-bit windows and a per-key table of 2^(c j) * P_i, j < W = ceil(255 / c), so that all windows share ONE
// bucket set (2^(c-1) buckets): bucket reduction and window combination happen once per MSM, and with the table
// resident in HBM the number of mixed additions per scalar is W, whatever c is. c = 16 (16 windows, 2^15 buckets)
// suits the witness MSMs, whose scalars are mostly single digits; the H MSM, whose scalars are uniform 254-bit
// values, runs c = 20 (13 windows, 2^19 buckets: 19 % fewer additions).
// The digit sort of a scalar vector (MsmSort) is separate from the group-specific part so that the
// MSMs that share scalars (A, B1, C in G1 and B2 in G2 all run over the witness, groth16.cpp:88-112)
// sort once and run their bucket work as one batched launch per stage.
struct MsmShape
{
    uint32_t c       = 16;       // window bits
    uint32_t windows = 16;       // ceil(255 / c): a scalar below r < 2^254 plus the signed-digit carry
    uint32_t buckets = 1u << 15; // 2^(c-1), bucket ids 1..buckets
    uint32_t levels  = 3;        // base-32 digits of a bucket index: ceil((c - 1) / 5)
};
inline MsmShape msm_shape(uint32_t c)
{
    MsmShape s;
    s.c       = c;
    s.windows = (255 + c - 1) / c;
    s.buckets = 1u << (c - 1);
    s.levels  = (c - 1 + 4) / 5;
    return s;
}
constexpr uint32_t kMsmMinWindowBits = 16;   // W = ceil(255 / c) <= 16: the window index is a 4-bit field of a sorted entry
constexpr uint32_t kMsmMaxWindowBits = 22;
constexpr uint32_t kMsmMaxWindows    = 16;
constexpr int      kMsmMaxLevels     = 5;
constexpr int      kMsmMaxBatch     = 3;                           // MSMs per batched launch
constexpr uint32_t kMsmHeavyRecords = 8;    // a bucket with more partial sums than max(this, 2 x average) is "heavy":
                                            // pre-reduced by whole blocks instead of one finalise thread
constexpr uint32_t kMsmHeavyBlocks  = 32;   // most slices (blocks) one heavy bucket is cut into
constexpr uint32_t kMsmHeavySlice   = 512;  // target records per slice
constexpr uint32_t kMsmHeavyGrid    = 128;  // blocks of the heavy kernel (work items are spread over them)
constexpr uint32_t kMsmMaxHeavy     = 1024; // heavy buckets handled that way (the rest stay thread-serial)
constexpr uint32_t kMsmEntryBaseBits = 27;  // sorted entry = base | window << 27 | sign << 31
constexpr uint32_t kMsmEntryBaseMask = (1u << kMsmEntryBaseBits) - 1u;
constexpr uint32_t kMsmFoldBlock    = 128;  // buckets per block of the bucket-sum kernel

struct MsmSort
{
    MsmShape  shape;
    uint32_t  n             = 0;       // scalars covered
    uint32_t  scalar_offset = 0;       // scalar index of element 0 when scalar_idx == nullptr
    const uint32_t* scalar_idx = nullptr; // optional gather list (n entries, not owned)
    uint32_t  cap_entries   = 0;
    uint32_t* counts        = nullptr; // buckets + 2
    uint32_t* offsets       = nullptr; // buckets + 2 (offsets[b] = first entry of bucket b; [B+1] = total)
    uint32_t* cursor        = nullptr; // buckets + 2
    uint32_t* sorted        = nullptr; // cap_entries
    // one-level path (c == 16): per-CTA shared-memory histograms over all 2^15 buckets, global cursors
    uint32_t* cta_hist      = nullptr; // sort_ctas x (buckets + 1): per-CTA bucket histograms
    uint32_t  sort_ctas     = 0;       // CTAs of the histogram pass (one per SM at most)
    uint32_t  per_cta       = 0;       // scalars per CTA (multiple of the block size)
    // two-level path (any c): partition by the high bucket bits through shared-memory staging, then one CTA sorts
    // each partition by the low bits in shared memory; every global write is part of a contiguous run
    bool      two_level     = false;
    uint32_t  part_bits     = 0;       // partitions = 2^part_bits, sub-buckets per partition = buckets >> part_bits
    uint32_t* part_cnt      = nullptr; // partitions + 1 counts, then offsets (exclusive scan; [P] = total)
    uint32_t* part_cursor   = nullptr; // partitions
    uint2*    inter         = nullptr; // cap_entries records (entry, bucket index - 1) grouped by partition
};

template <class XY>
struct MsmBases
{
    typedef typename XY::Affine Affine;
    MsmShape  shape;
    uint32_t  n          = 0;       // table columns (bases kept)
    uint32_t* scalar_idx = nullptr; // n, only when infinity bases were filtered out: scalar index of each kept base
    uint8_t*  skip       = nullptr; // n, only when infinity columns were kept: 1 = column is infinity
    Affine*   table      = nullptr; // windows x n affine points: table[j*n + i] = 2^(c j) * P_i
};

template <class XY>
struct MsmScratch
{
    MsmShape  shape;
    uint32_t  chunk         = 32;      // sorted entries per accumulate thread (per MSM: G2 wants smaller chunks)
    uint32_t* heavy_count   = nullptr; // 1: number of heavy buckets of the current sort at this chunk size
    uint32_t* heavy_ids     = nullptr; // kMsmMaxHeavy
    uint32_t* heavy_slot    = nullptr; // buckets + 2: 0 = light bucket, k + 1 = k-th heavy bucket
    XY*       records       = nullptr; // cap_entries / chunk + buckets + 2 partial sums, grouped by bucket
    XY*       heavy_partial = nullptr; // kMsmMaxHeavy x kMsmHeavyBlocks
    XY*       heavy_sum     = nullptr; // kMsmMaxHeavy
    uint32_t* heavy_done    = nullptr; // arrival counters (self-resetting): kMsmMaxHeavy heavy buckets, then the fold classes + 1 (msm.cuh)
    XY*       bsum          = nullptr; // buckets : the sum of every bucket
    XY*       s0part        = nullptr; // (buckets / 1024) x 4 x 32 : per plane and quarter of digit 1, per digit 0
    XY*       s1part        = nullptr; // (buckets / 1024) x 32     : per plane, per digit 1 (sum over digit 0)
    XY*       classes       = nullptr; // levels x 32 weighted class sums, then their slice sums (msm.cuh: kMsmFoldPartOffset)
    XY*       result        = nullptr; // 1 (device)
    cudaEvent_t ev_acc0 = nullptr, ev_acc1 = nullptr; // bracket the bucket-accumulation kernel of the last run
};

// window_bits: 16 = the one-level sort; anything else (kMsmMinWindowBits..kMsmMaxWindowBits) the two-level sort
void msm_sort_create(MsmSort& s, uint32_t n, const uint32_t* scalar_idx, uint32_t scalar_offset, uint32_t window_bits = 16,
                     bool force_two_level = false);
void msm_sort_destroy(MsmSort& s);
// scalars: device array of 32-byte little-endian integers (canonical or not; reduced mod r on the fly).
void msm_sort_run(MsmSort& s, const uint32_t* scalars, cudaStream_t st);
uint32_t msm_default_chunk(uint64_t n);

// points_host: `count` affine Montgomery points exactly as in the zkey (64 B G1 / 128 B G2), (0,0) = infinity.
// filter_inf: drop infinity bases (like multiexp.cpp:57) and remember the scalar index of each kept base;
// otherwise all `count` columns are kept (infinity columns are all-zero and cost one load per digit).
template <class XY>
void msm_bases_create(MsmBases<XY>& out, const uint8_t* points_host, uint64_t count, bool filter_inf,
                      cudaStream_t st, uint32_t window_bits = 16);
template <class XY>
void msm_bases_destroy(MsmBases<XY>& b);
// chunk = sorted entries per accumulate thread (0 = msm_default_chunk(sort.n))
template <class XY>
void msm_scratch_create(MsmScratch<XY>& s, const MsmSort& sort, uint32_t chunk);
template <class XY>
void msm_scratch_destroy(MsmScratch<XY>& s);
// Bucket accumulation + reduction of `nb` <= kMsmMaxBatch MSMs that share one digit sort (bases[k]->n == sort.n)
// and one chunk size. Leaves each XYZZ result in scr[k]->result (device).
template <class XY>
void msm_reduce_batch(const MsmSort& sort, const MsmBases<XY>* const* bases, MsmScratch<XY>* const* scr, int nb,
                      cudaStream_t st);

// duration of the last bucket-accumulation launch (ms, CUDA events on its stream) and its number of sorted entries
template <class XY>
void msm_last_accumulate(const MsmSort& sort, const MsmScratch<XY>& s, float* ms, uint64_t* entries);
// kernels launched by one msm_sort_run / one msm_reduce_batch
constexpr uint32_t kMsmSortLaunches   = 4; // hist, column sums, bucket scan, scatter  |  count, scan, partition, local sort
constexpr uint32_t kMsmReduceLaunches = 7; // classify, accumulate, heavy, bucket sums, plane fold, class fold, final

extern template struct MsmBases<G1Xyzz>;
extern template struct MsmBases<G2Xyzz>;

// ------------------------------------------------------------------ diagnostics / tests
// elementwise field kernels: field 0 = Fr, 1 = Fq, 2 = Fq2; op: 0 mul 1 add 2 sub 3 neg 4 toMont 5 fromMont
// 6 sqr 7 inv 8 a*b + b*b (single-reduction dual product, prime fields only). Device pointers.
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
