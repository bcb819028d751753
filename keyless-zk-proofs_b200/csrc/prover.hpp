// Groth16 prover resident on one GPU: proving key uploaded once, one proof per prove() call.
// Replaces Groth16::Prover<Engine> + FullProverImpl of the reference
// (rust-rapidsnark/rapidsnark/src/groth16.hpp:44-106, fullprover.cpp:136-250).
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace kzp
{

// stage timings of the last proof, milliseconds (CUDA events on the launching streams, except *_host)
struct ProveTimings
{
    float h2d_ms      = 0; // witness host -> device (packed copy + expansion kernel)
    float spmv_ms     = 0;
    float ntt_ms      = 0; // 3 x (iNTT + coset + NTT) + pointwise
    float msm_h_ms    = 0;
    float msm_wsort_ms = 0; // digit sort of the witness (shared by A, B1, B2, C)
    float msm_wg1_ms   = 0; // A, B1, C: one batched G1 launch per stage
    float msm_wg2_ms   = 0; // B2
    float h2d_mbytes   = 0; // MB the last witness upload moved over PCIe (packed: bits/bytes travel as one byte)
    float gpu_ms      = 0; // first kernel to last kernel (both streams)
    float assemble_host_ms = 0;
    float total_host_ms    = 0; // wall clock of prove_*()
    uint32_t kernel_launches = 0;
};

// The five MSM results of the last proof before blinding, affine canonical
// little-endian: A(64) B1(64) B2(128) C(64) H(64)
struct MsmArtefacts
{
    uint8_t bytes[384];
};

// Jacobian-free partial results of one shard (XYZZ, Montgomery): A, B1, C, H (128 B each) then B2 (256 B)
struct ShardPartials
{
    uint8_t bytes[4 * 128 + 256];
};

// the five points of the proving key the proof assembly needs (zkey section 2; affine Montgomery bytes)
struct HostVk
{
    uint8_t alpha1[64], beta1[64], delta1[64];
    uint8_t beta2[128], delta2[128];
};

// Host-only: sum `count` shards' partials, blind with (r,s) (nullptr = sample like groth16.cpp:296-316),
// print the proof JSON. art_out (optional) receives the five unblinded MSM results.
std::string assemble_proof(const HostVk& vk, const ShardPartials* parts, int count, const uint8_t* r32,
                           const uint8_t* s32, MsmArtefacts* art_out);

// Host-only view of the packed witness transfer (prover.cu): packs `count` <= 32768 plain 32-byte values exactly as a
// staging worker packs one slice — `out` (pack_slice_capacity() bytes, 16-byte aligned) receives
// [1 byte per wire][1 flag bit per wire][32 bytes per value >= 256]; returns the number of bytes that would cross
// PCIe and stores the number of full-width values in *n_full. Used by the CPU test suite.
size_t pack_slice_capacity();
size_t pack_witness_slice(const uint8_t* values, uint32_t count, uint8_t* out, uint32_t* n_full);

class DeviceProverImpl;
class ProverGroup;

class DeviceProver
{
    std::unique_ptr<DeviceProverImpl> impl_;  // one GPU ...
    std::unique_ptr<ProverGroup>      group_; // ... or one proof over several (exactly one of the two is set)
    DeviceProverImpl&                 first() const;

public:
    // Throws kzp::LoadError / kzp::FormatError / kzp::CudaError.
    // shard_rank/shard_world: this instance holds only base range [rank*n/world, (rank+1)*n/world) of every
    // MSM section; world == 1 is the ordinary single-GPU prover.
    DeviceProver(const std::string& zkey_path, int device, int shard_rank = 0, int shard_world = 1);
    // One proof over devices.size() GPUs of this process (1..8; a device may be listed more than once): shard r on
    // devices[r] holds base range r of every MSM section and computes one of the three coset-NTT chains (when there
    // are fewer shards than chains, shard 1 computes two); the chain outputs cross NVLink slice by slice, the
    // 768-byte partial results are summed on the host. Every prove*/upload*/run_gpu call below then drives all of them.
    DeviceProver(const std::string& zkey_path, const std::vector<int>& devices);
    ~DeviceProver();
    int  group_size() const;           // 1 for a single-GPU prover
    bool group_fused_exchange() const; // the slices travel as peer stores of the chain's last level (else peer copies)
    bool group_distributed_ntt() const; // every coset-NTT chain is spread over all shards (else one chain per shard)

    uint32_t n_vars() const;
    uint32_t n_public() const;
    uint32_t domain_size() const;
    uint64_t n_coefs() const;
    int      device() const;

    // Upload a witness (n values x 32 B canonical LE) from host memory; n must be >= n_vars.
    void upload_witness(const uint8_t* values, uint64_t n);
    // Same, reading the values with pread() from an open witness file at `file_offset`.
    void upload_witness_fd(int fd, uint64_t file_offset, uint64_t n);
    // Run the GPU part on the resident witness; leaves partial MSM results on the host.
    void run_gpu();
    const ShardPartials& partials() const;
    // Sum `count` shard partials, blind with (r,s) (32 B canonical LE each; nullptr = sample like
    // groth16.cpp:296-316) and print the proof JSON (groth16.cpp:379-410 + nlohmann dump()).
    std::string assemble(const ShardPartials* parts, int count, const uint8_t* r32, const uint8_t* s32);

    // one proof on the witness that is already resident (upload_witness*): GPU part + assembly, with the host work
    // overlapped with the GPU exactly as in prove()/prove_fd()
    std::string prove_resident(const uint8_t* r32, const uint8_t* s32);

    // convenience: upload + run + assemble for world == 1
    std::string prove(const uint8_t* values, uint64_t n, const uint8_t* r32, const uint8_t* s32);
    std::string prove_fd(int fd, uint64_t file_offset, uint64_t n, const uint8_t* r32, const uint8_t* s32);

    const ProveTimings& timings() const;                 // a group: per-stage maximum over the shards
    const ProveTimings& shard_timings(int shard) const;  // one shard of a group (shard 0 of a single prover)
    // bucket-accumulation kernel of MSM `which` (0 A, 1 B1, 2 B2, 3 C, 4 H) in the last proof: duration and entries
    void msm_profile(int which, float* accumulate_ms, uint64_t* entries) const;
    const MsmArtefacts& msm_artefacts() const; // filled by assemble()
    // H coefficients of the last proof (domain_size x 32 B canonical, natural order) -> host buffer
    void copy_h(uint8_t* out) const;
    // when enabled, a and b after the SpMV are kept (device copy) and can be fetched with copy_ab
    void set_keep_ab(bool on);
    void copy_ab(uint8_t* out) const; // 2 * domain_size * 32 B, Montgomery
};

} // namespace kzp
