/* kzp_b200 — C ABI of the B200-native Groth16/BN254 prover (libkzp_b200.so).
 *
 * Plain pointers and sizes only; no C++ or torch types. Two layers:
 *
 *  (1) The prover object. Same life cycle and error behaviour as the reference's C++ boundary class
 *      FullProver (rust-rapidsnark/rapidsnark/src/fullprover.hpp:54-64, fullprover.cpp:80-125,204-250):
 *      construct from a zkey path (never fails hard: the state says why it is unusable), prove from a
 *      .wtns path, get a malloc'd compact proof JSON plus prover_time in ms. The reference binds that
 *      class through bindgen (rust-rapidsnark/build.rs:183-206, src/lib.rs:41-106); the same mangled C++
 *      symbols are ALSO exported by this library (include/fullprover_b200.hpp), so either binding works.
 *      INTEGRATION.md shows both.
 *
 *  (2) Component entry points used by the parity tests and the micro-benchmarks: Fr NTT
 *      (FFT<Fr>::fft/ifft, fft.cpp:192-246), G1/G2 MSM (Curve::multiMulByScalar, curve.hpp:209-215 ->
 *      multiexp.cpp:183-245), raw field and group operations (fr.hpp:206-281, curve.cpp), all with host
 *      buffers in the reference's own byte layouts.
 *
 * Every function returns 0 on success unless stated otherwise; on failure kzp_last_error() (thread-local)
 * describes the problem. Nothing here falls back to the CPU: without a CUDA device every compute entry
 * point fails with KZP_ERR_CUDA.
 */
#ifndef KZP_B200_H
#define KZP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------------- */
#define KZP_OK 0
#define KZP_ERR_CUDA 1     /* no device, launch failure, out of memory */
#define KZP_ERR_FORMAT 2   /* malformed zkey/wtns, wrong curve, bad arguments */
#define KZP_ERR_IO 3       /* open/fstat/mmap failure */
#define KZP_ERR_STATE 4    /* prover not ready / call order */

/* FullProverState (fullprover.hpp:11-16) */
#define KZP_STATE_OK 0
#define KZP_STATE_ZKEY_FILE_LOAD_ERROR 1
#define KZP_STATE_UNSUPPORTED_ZKEY_CURVE 2
/* not in the reference's enum: the CUDA context died under this prover (sticky error). The handle answers every
 * later call with PROVER_NOT_READY; the Itanium shim reports it the same way (prove() is const there). */
#define KZP_STATE_DEVICE_FAULT 3
/* ProverResponseType (fullprover.hpp:5-9) */
#define KZP_RESPONSE_SUCCESS 0
#define KZP_RESPONSE_ERROR 1
/* ProverError (fullprover.hpp:18-24) */
#define KZP_PROVER_ERROR_NONE 0
#define KZP_PROVER_ERROR_NOT_READY 1
#define KZP_PROVER_ERROR_INVALID_INPUT 2
#define KZP_PROVER_ERROR_WITNESS_GENERATION_INVALID_CURVE 3

const char* kzp_last_error(void);
int         kzp_device_count(void); /* number of CUDA devices visible, 0 if none / no driver */
const char* kzp_version(void);
void        kzp_free(void* p);      /* frees buffers returned by this library (malloc family) */

/* ---- (1) prover --------------------------------------------------------------------------------- */
typedef struct kzp_prover kzp_prover;

/* FullProver::FullProver(const char*) — fullprover.cpp:80-101. Always returns a handle (NULL only when
 * out of host memory); *state_out receives a KZP_STATE_* value. device < 0 selects $KZP_DEVICE or 0. */
kzp_prover* kzp_prover_new(const char* zkey_path, int device, int* state_out);
/* MSM base ranges of shard `rank` of `world` only; used one process per GPU. */
kzp_prover* kzp_prover_new_sharded(const char* zkey_path, int device, int rank, int world, int* state_out);
/* ONE proof over n_devices GPUs of this process (1..8): shard r on devices[r] holds base range r of
 * every MSM section. With 2, 4 or 8 shards every coset-NTT chain is spread over all of them: each GPU runs 1/N of every
 * level of a, b and c, and the two transposes between the level partitions plus the delivery of the H slices are peer
 * stores over NVLink fused into the level kernels ($KZP_GROUP_NTT=chain turns this off). Otherwise each chain is
 * computed by one shard (the reference runs them as three std::async tasks, groth16.cpp:172-262) and its output crosses
 * NVLink as peer stores of the chain's last level (peer copies when peer access is unavailable, the domain is too small
 * for the batched chain, or $KZP_GROUP_SCATTER=0). The 768-byte partial results are summed on the host.
 * The handle behaves like any other: kzp_prover_prove / _prove_mem / _prove_resident / _upload_witness* / _run_gpu /
 * _timings (per-stage maximum over the shards) / _get_h / _get_msm_results; kzp_prover_keep_ab is not available.
 * devices == NULL: $KZP_SHARD_DEVICES ("0,1,2,3"). kzp_prover_new(path, -1, ..) — and therefore the reference-shaped
 * FullProver::FullProver(zkeyPath) — builds a group whenever $KZP_SHARD_DEVICES is set. */
kzp_prover* kzp_prover_new_group(const char* zkey_path, const int* devices, int n_devices, int* state_out);
int         kzp_prover_group_info(kzp_prover* p, int* shards, int* fused_exchange, int* distributed_ntt); /* 1 shard = single-GPU prover */
void        kzp_prover_free(kzp_prover* p); /* FullProver::~FullProver */

/* FullProver::prove(const char* wtnsPath) — fullprover.cpp:114-125,204-250.
 * Returns KZP_RESPONSE_*; *json_out is malloc'd (kzp_free) on success, NULL otherwise; *error_out is a
 * KZP_PROVER_ERROR_*; *prover_time_ms covers witness upload + GPU work + proof JSON, like the reference's
 * metrics.prover_time (file mapping excluded). r32/s32: optional fixed blinding scalars (32-byte LE,
 * canonical); pass NULL for fresh randomness (groth16.cpp:296-316). */
int kzp_prover_prove(kzp_prover* p, const char* wtns_path, const uint8_t* r32, const uint8_t* s32,
                     char** json_out, int* error_out, int* prover_time_ms);
/* Error classes: a malformed / unreadable / wrong-curve witness -> INVALID_INPUT (or WITNESS_GENERATION_INVALID_CURVE);
 * a CUDA failure (launch error, out of memory, lost device) -> PROVER_NOT_READY, and when the error is sticky the
 * handle's state becomes KZP_STATE_DEVICE_FAULT. kzp_prover_last_status gives the KZP_ERR_* class of the last failure. */
int kzp_prover_state(const kzp_prover* p);       /* KZP_STATE_* now (health check) */
int kzp_prover_last_status(const kzp_prover* p); /* KZP_OK or the KZP_ERR_* of the last failed call */
/* Additive entry: witness values already in memory, n x 32-byte LE canonical. */
int kzp_prover_prove_mem(kzp_prover* p, const uint8_t* witness, uint64_t n, const uint8_t* r32,
                         const uint8_t* s32, char** json_out, int* error_out, int* prover_time_ms);

/* Split life cycle, used by bench.py and by the sharded (one process per GPU) mode. */
int kzp_prover_upload_witness(kzp_prover* p, const uint8_t* witness, uint64_t n);
int kzp_prover_upload_witness_file(kzp_prover* p, const char* wtns_path);
int kzp_prover_run_gpu(kzp_prover* p);                         /* all kernels, result = this shard's partials */
/* One whole proof on the witness already resident in HBM (kzp_prover_upload_witness*): same code path as
 * kzp_prover_prove minus the upload, i.e. the host-side proof assembly overlaps the GPU work. Result like
 * kzp_prover_prove. */
int kzp_prover_prove_resident(kzp_prover* p, const uint8_t* r32, const uint8_t* s32, char** json_out, int* error_out,
                              int* prover_time_ms);
#define KZP_PARTIALS_BYTES 768                                 /* A,B1,C,H as XYZZ (4x128) + B2 XYZZ (256) */
int kzp_prover_get_partials(kzp_prover* p, uint8_t* out768);
/* sum `count` shards' partials (count*768 bytes), blind and print */
int kzp_prover_assemble(kzp_prover* p, const uint8_t* partials, int count, const uint8_t* r32,
                        const uint8_t* s32, char** json_out);

/* introspection */
int kzp_prover_info(kzp_prover* p, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size,
                    uint64_t* n_coefs, int* device);
/* last proof: floats {h2d, spmv, ntt, msm_h, msm_witness_sort, msm_witness_g1 (A,B1,C batched), msm_witness_g2 (B2),
 * h2d_megabytes (what the upload moved over PCIe), gpu, assemble_host, total_host, kernel_launches}; returns the number of values written (<= cap) */
int kzp_prover_timings(kzp_prover* p, float* out, int cap);
/* the same values for one shard of a group (kzp_prover_new_group): where the time of a sharded proof goes */
int kzp_prover_group_shard_timings(kzp_prover* p, int shard, float* out, int cap);
/* bucket-accumulation kernel (the dominant kernel) of MSM `which` (0 A, 1 B1, 2 B2, 3 C, 4 H) in the last proof:
 * its duration from CUDA events on its stream and the number of (point, bucket) entries it summed. A, B1 and C
 * share one batched launch over one digit sort: 0, 1 and 3 return that launch and the shared entry count. */
int kzp_prover_msm_profile(kzp_prover* p, int which, float* accumulate_ms, uint64_t* entries);
/* parity artefacts of the last proof */
int kzp_prover_get_h(kzp_prover* p, uint8_t* out, uint64_t out_bytes);       /* domain x 32, canonical */
int kzp_prover_keep_ab(kzp_prover* p, int on);
int kzp_prover_get_ab(kzp_prover* p, uint8_t* out, uint64_t out_bytes);      /* 2 x domain x 32, Montgomery */
int kzp_prover_get_msm_results(kzp_prover* p, uint8_t* out384);              /* A,B1,B2,C,H affine canonical */

/* ---- (1b) prover pool: GPU-per-request scheduling --------------------------------
 * The reference service owns ONE FullProver behind Arc<tokio::Mutex<Option<_>>> and proves on the async worker
 * (prover-service/src/prover_state.rs:21,38-47,101; prover_handler.rs:266-283): proofs are serialised. A pool holds
 * one resident prover per listed device (list a device twice for two provers on that GPU: one proof's witness
 * staging and host assembly then overlap the other's kernels) and hands each request the least recently used free
 * prover, FIFO. Every call is thread-safe and blocks until a prover is free; call it from spawn_blocking. */
typedef struct kzp_pool kzp_pool;
/* devices == NULL: $KZP_POOL_DEVICES ("0,1,2,...") or every visible device. Keys load concurrently. Always returns
 * a handle; *state_out is KZP_STATE_OK only when every prover is ready. */
kzp_pool* kzp_pool_new(const char* zkey_path, const int* devices, int n_devices, int* state_out);
void      kzp_pool_free(kzp_pool* pool); /* waits for proofs in flight */
int       kzp_pool_size(const kzp_pool* pool);
int       kzp_pool_device(const kzp_pool* pool, int slot);
int       kzp_pool_healthy(const kzp_pool* pool); /* provers whose state is still KZP_STATE_OK (health check) */
/* kzp_prover_prove / kzp_prover_prove_mem on the next free prover; *slot_out (optional) says which one ran it */
int kzp_pool_prove(kzp_pool* pool, const char* wtns_path, const uint8_t* r32, const uint8_t* s32, char** json_out,
                   int* error_out, int* prover_time_ms, int* slot_out);
int kzp_pool_prove_mem(kzp_pool* pool, const uint8_t* witness, uint64_t n, const uint8_t* r32, const uint8_t* s32,
                       char** json_out, int* error_out, int* prover_time_ms, int* slot_out);
/* Fused verify-before-return: with on != 0 every proof is checked under the zkey's verifying key
 * (kzp_host_verify, host only) before it is returned — after the prover has been released, so the GPU is already on the
 * next request. A proof that does not verify is dropped: KZP_RESPONSE_ERROR / KZP_PROVER_ERROR_INVALID_INPUT when the
 * witness is at fault (unreadable public signals, or the device that produced it is still healthy: the witness does
 * not satisfy the circuit), KZP_PROVER_ERROR_NOT_READY when the verifier itself failed or the producing device has
 * faulted since (the slot is then retired). Off by default. */
int kzp_pool_set_verify(kzp_pool* pool, int on);
/* proofs served per slot (returns the number of slots written) and the deepest queue seen */
int kzp_pool_stats(kzp_pool* pool, uint64_t* proofs_per_slot, int cap, uint64_t* max_waiting);
/* host-only exercise of the checkout queue (no GPU needed; used by the CPU test suite) */
int kzp_pool_sched_selftest(int slots, int threads, int jobs_per_thread, int hold_us, uint64_t* per_slot_out,
                            int* max_concurrent_per_slot, uint64_t* max_waiting);

/* ---- (2) components ------------------------------------------------------------------------------ */
/* FFT<Fr>::fft / ifft contract: n = 2^k Montgomery elements, natural order in and out, in place. */
int kzp_fr_ntt(uint8_t* data, uint64_t n, int inverse, int device);
/* The prover's H chain on one vector: ifft, multiply by w_2n^i, fft (groth16.cpp:172-203). */
int kzp_fr_coset_chain(uint8_t* data, uint64_t n, int device);
/* times `iters` coset chains on a device-resident vector of size 2^log_n; *ms_per_chain averaged */
int kzp_fr_ntt_bench(uint32_t log_n, int iters, int device, float* ms_per_chain);

typedef struct kzp_msm kzp_msm;
/* group: 0 = G1 (64-byte bases), 1 = G2 (128-byte bases); bases affine Montgomery, (0,0) = infinity */
kzp_msm* kzp_msm_new(int group, const uint8_t* bases, uint64_t n, int device);
/* window_bits: signed-digit window size c, 16..22 (0 = default 16, or $KZP_MSM_WINDOW): W = ceil(255 / c) table
 * windows share one set of 2^(c-1) buckets. two_level != 0 forces the two-pass digit sort (always used for c != 16). */
kzp_msm* kzp_msm_new_ex(int group, const uint8_t* bases, uint64_t n, int device, int window_bits, int two_level);
void     kzp_msm_free(kzp_msm* m);
/* scalars: n x 32-byte LE integers; out: affine canonical LE (64 / 128 bytes), zeros for infinity */
int kzp_msm_run(kzp_msm* m, const uint8_t* scalars, uint8_t* out);
/* device-resident timing: uploads scalars once, runs `iters` MSMs, average ms; entries = non-zero digits */
int kzp_msm_bench(kzp_msm* m, const uint8_t* scalars, int iters, float* ms_per_msm, uint64_t* entries);

/* field: 0 Fr, 1 Fq, 2 Fq2 (64-byte elements); op: 0 mul 1 add 2 sub 3 neg 4 toMontgomery 5 fromMontgomery
 * 6 square 7 inverse 8 a*b + b*b (dual product with one reduction per component). Host buffers, `count` elements; b may be NULL for unary ops. */
int kzp_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t count,
                 int device);
/* group: 0 G1, 1 G2; op: 0 xyzz += affine, 1 xyzz += xyzz, 2 double, 3 / 4 the same addition / doubling as the
 * bucket-reduction kernels run it (four lanes per point). p/out XYZZ (128/256 B per point). */
int kzp_point_op(int group, int op, const uint8_t* p, const uint8_t* q, uint8_t* out, uint64_t count,
                 int device);
/* integer-pipe roofline probe: carry-chained 32x32+64 multiply-adds (IMAD.WIDE.U32[.X]) on every SM */
int kzp_imad_peak(int iters, int device, float* ms, uint64_t* multiply_adds);

/* host-only helpers (no GPU needed): decimal printing and file parsing, for the CPU test suite */
int kzp_host_parse_zkey(const char* path, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size,
                        uint64_t* n_coefs, int* state_out);
/* proof assembly without a GPU: sums `count` 768-byte shard partials, blinds and prints (host arithmetic only;
 * this is the step rank 0 performs after the all-gather in sharded mode). msm_out384 may be NULL. */
int kzp_host_assemble(const char* zkey_path, const uint8_t* partials, int count, const uint8_t* r32,
                      const uint8_t* s32, char** json_out, uint8_t* msm_out384);
/* Host-only pairing product check: *result_out = 1 iff prod_i e(P_i, Q_i) == 1. g1: n x 64-byte
 * affine Montgomery points (zkey layout, zeros = infinity), g2: n x 128 bytes. Points off the curve -> KZP_ERR_FORMAT. */
int kzp_host_pairing_check(const uint8_t* g1, const uint8_t* g2, int n, int* result_out);
/* Groth16 verification of a proof JSON (the string the prove calls return) under the verifying key stored in the
 * zkey (section 2 + IC section 3) — what prover-service does after every proof through ark-groth16
 * (prover-service/src/request_handler/prover_handler.rs:329-336). public32: n_public x 32-byte LE canonical signals.
 * Returns 0 and *valid_out = 1 / 0; non-zero for malformed inputs (kzp_verify_last_error()). No GPU involved. */
int kzp_host_verify(const char* zkey_path, const char* proof_json, const uint8_t* public32, uint32_t n_public,
                    int* valid_out);
const char* kzp_verify_last_error(void);
/* Host-only: how one slice (count <= 32768 values of 32 bytes) of a witness is packed for the trip over PCIe
 * (csrc/prover.cu): out (>= 1,085,440 bytes, 16-byte aligned) = [count small bytes, padded to 32768][4096 flag bytes]
 * [32 bytes per value >= 256]; *packed_bytes = 36,864 + 32 * *n_full is the prefix that is actually copied. */
int kzp_host_pack_witness_slice(const uint8_t* values, uint32_t count, uint8_t* out, uint64_t out_cap, uint64_t* packed_bytes,
                                uint32_t* n_full);
int kzp_host_fq_decimal(const uint8_t* mont32, char* out, size_t cap);
int kzp_host_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* KZP_B200_H */
