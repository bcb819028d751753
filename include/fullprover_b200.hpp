// Drop-in declaration of the reference's C++ boundary for the GPU prover.
//
// prover-service reaches the prover through bindgen over rust-rapidsnark/rapidsnark/src/fullprover.hpp
// (rust-rapidsnark/wrapper.hpp:1-5, build.rs:183-206). That interface is an Itanium-ABI C++ class, not
// extern "C", so a replacement library has to export the very same mangled symbols with the very same
// object layouts. This header re-declares that interface (type names, enumerator order, member order and
// signatures are dictated by the ABI) and libkzp_b200.so defines it:
//
//   _ZN10FullProverC1EPKc / C2            FullProver::FullProver(const char*)      fullprover.hpp:61
//   _ZN10FullProverD1Ev   / D2            FullProver::~FullProver()                fullprover.hpp:62
//   _ZNK10FullProver5proveEPKc            FullProver::prove(const char*) const     fullprover.hpp:63
//   _ZN14ProverResponseC1E11ProverError                                            fullprover.hpp:45
//   _ZN14ProverResponseC1EPKc21ProverResponseMetrics                               fullprover.hpp:46
//   _ZN14ProverResponseD1Ev                                                        fullprover.hpp:52
//   _ZN14ProverResponse12empty_stringE                                             fullprover.hpp:42
//
// Layouts (checked by static_asserts in csrc/fullprover_abi.cu and by tests/test_host_library.py::test_cxx_abi_against_reference_header):
//   FullProver      16 bytes : impl pointer @0, state @8 (Rust reads .state directly, lib.rs:53)
//   ProverResponse  24 bytes : type @0, raw_json @8, error @16, metrics.prover_time @20
//
// The existing Rust crate keeps compiling against the reference's own header; only its link line changes
// (INTEGRATION.md). C and ctypes users should prefer the extern "C" twin in kzp_b200.h.
#pragma once

class FullProverImpl; // here: owns a kzp::DeviceProver (GPU-resident proving key)

enum ProverResponseType
{
    SUCCESS, // raw_json holds the proof
    ERROR    // error says why there is none
};

enum FullProverState
{
    OK,
    ZKEY_FILE_LOAD_ERROR,  // open/fstat/mmap failed, or the GPU could not take the key
    UNSUPPORTED_ZKEY_CURVE // not a BN254 groth16 zkey (bad magic/version/protocol/prime)
};

enum ProverError
{
    NONE,
    PROVER_NOT_READY,                 // constructor did not reach state OK
    INVALID_INPUT,                    // unreadable/short witness (the reference never produces this value)
    WITNESS_GENERATION_INVALID_CURVE  // witness prime is not the BN254 scalar field
};

struct ProverResponseMetrics
{
    int prover_time; // milliseconds: witness upload + GPU proof + JSON (file mapping excluded)
};

struct ProverResponse
{
    ProverResponseType    type;
    char const*           raw_json; // malloc-family allocation unless it is empty_string
    ProverError           error;
    ProverResponseMetrics metrics;

private:
    static char const* const empty_string;

public:
    ProverResponse(ProverError _error);
    ProverResponse(const char* _raw_json, ProverResponseMetrics _metrics);

    ProverResponse()                                 = delete;
    ProverResponse(ProverResponse const&)            = delete;
    ProverResponse& operator=(ProverResponse const&) = delete;

    ~ProverResponse();
};

class FullProver
{
    FullProverImpl* impl;
    FullProverState state;

public:
    FullProver() = delete;
    FullProver(const char* _zkeyFileName);
    ~FullProver();
    ProverResponse prove(const char* input) const;
};
