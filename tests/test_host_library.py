"""CPU-side tests of the product library: it loads, exports every symbol the headers declare, its host-only helpers
agree with the oracle, and every compute entry point refuses to run without a GPU (no CPU fallback exists)."""
import ctypes
import json
import os
import random
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT


def _declared_c_symbols():
    text = open(os.path.join(ROOT, "include", "kzp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kzp_[a-z0-9_]+)\s*\(", text)))


MANGLED = [
    "_ZN10FullProverC1EPKc", "_ZN10FullProverD1Ev", "_ZNK10FullProver5proveEPKc",
    "_ZN14ProverResponseC1E11ProverError", "_ZN14ProverResponseC1EPKc21ProverResponseMetrics",
    "_ZN14ProverResponseD1Ev", "_ZN14ProverResponse12empty_stringE",
]


def test_library_exports_every_declared_symbol(kzp):
    out = subprocess.check_output(["nm", "-D", "--defined-only", kzp.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    declared = _declared_c_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/kzp_b200.h but not exported: %s" % missing
    # the Itanium-ABI symbols bindgen binds for rust-rapidsnark
    assert not [s for s in MANGLED if s not in exported]
    L = kzp.lib()
    for s in declared:
        getattr(L, s)


def test_no_oracle_in_product(kzp):
    """The product library must not link or embed the CPU checkers."""
    out = subprocess.check_output(["ldd", kzp.LIB_PATH], text=True)
    assert "kzp_ref" not in out and "kzp_port" not in out and "gmp" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "keyless-zk-proofs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import bn254" not in src and "libkzp_port" not in src and "libkzp_ref" not in src, f


def test_host_field_ops_match_oracle(kzp, oracle):
    rnd = random.Random(7)
    L = kzp.lib()
    # 0/1/2: 64-bit-limb host fields used by the proof assembly; 10/11/12: the portable path of the device templates
    for field, mod in ((0, oracle.R_MOD), (1, oracle.Q_MOD), (10, oracle.R_MOD), (11, oracle.Q_MOD)):
        for i in range(300):
            a, b = rnd.randrange(mod), rnd.randrange(mod)
            if i % 11 == 0:
                a = mod - 1
            if i % 13 == 0:
                b = 0
            for op, want in ((0, oracle.mont_mul(a, b, mod)), (1, (a + b) % mod), (2, (a - b) % mod), (3, (-a) % mod),
                             (4, oracle.to_mont(a, mod)), (5, oracle.from_mont(a, mod)), (6, oracle.mont_mul(a, a, mod)),
                             (8, (oracle.mont_mul(a, b, mod) + oracle.mont_mul(b, b, mod)) % mod)):
                out = ctypes.create_string_buffer(32)
                assert L.kzp_host_field_op(field, op, oracle.le32(a), oracle.le32(b), out) == 0
                assert oracle.from_le(out.raw) == want, (field, op, hex(a), hex(b))
        a = rnd.randrange(1, mod)
        out = ctypes.create_string_buffer(32)
        L.kzp_host_field_op(field, 7, oracle.le32(oracle.to_mont(a, mod)), None, out)
        assert oracle.from_mont(oracle.from_le(out.raw), mod) == pow(a, -1, mod)
    for field in (2, 12):
        enc = lambda x: oracle.le32(oracle.to_mont(x[0], oracle.Q_MOD)) + oracle.le32(oracle.to_mont(x[1], oracle.Q_MOD))
        dec = lambda r: (oracle.from_mont(oracle.from_le(r[:32]), oracle.Q_MOD), oracle.from_mont(oracle.from_le(r[32:]), oracle.Q_MOD))
        for _ in range(40):
            a = (rnd.randrange(oracle.Q_MOD), rnd.randrange(oracle.Q_MOD))
            b = (rnd.randrange(oracle.Q_MOD), rnd.randrange(oracle.Q_MOD))
            out = ctypes.create_string_buffer(64)
            L.kzp_host_field_op(field, 0, enc(a), enc(b), out)
            assert dec(out.raw) == oracle.f2_mul(a, b)
            L.kzp_host_field_op(field, 6, enc(a), None, out)
            assert dec(out.raw) == oracle.f2_sqr(a)
            L.kzp_host_field_op(field, 7, enc(a), None, out)
            assert dec(out.raw) == oracle.f2_inv(a)


def test_field_kats_on_host_paths(kzp, oracle):
    """The reference's limb-level KATs (canonical inputs) through both host instantiations of the field code."""
    L = kzp.lib()
    opcode = {"mul": 0, "add": 1, "sub": 2, "square": 6}
    n = 0
    for rec in json.load(open(os.path.join(GOLDEN, "field_kats.json"))):
        mod = oracle.R_MOD if rec["field"] == "Fr" else oracle.Q_MOD
        a, b = int(rec["a"], 16), int(rec.get("b", "0x0"), 16)
        if a >= mod or b >= mod:
            continue
        for field in ((0, 10) if rec["field"] == "Fr" else (1, 11)):
            out = ctypes.create_string_buffer(32)
            L.kzp_host_field_op(field, opcode[rec["op"]], oracle.le32(a), oracle.le32(b), out)
            assert oracle.from_le(out.raw) == int(rec["expected"], 16), rec
            n += 1
    assert n >= 40


def test_decimal_printing(kzp, oracle):
    L = kzp.lib()
    buf = ctypes.create_string_buffer(100)
    rnd = random.Random(3)
    vals = [0, 1, 9, 10, 999999999, 1000000000, 10 ** 18, oracle.Q_MOD - 1] + [rnd.randrange(oracle.Q_MOD) for _ in range(50)]
    for v in vals:
        assert L.kzp_host_fq_decimal(oracle.le32(oracle.to_mont(v, oracle.Q_MOD)), buf, 100) == 0
        assert buf.value.decode() == str(v)
    assert L.kzp_host_fq_decimal(oracle.le32(5), buf, 1) != 0  # buffer too small


def test_zkey_parsing_and_error_classes(kzp, oracle, workdir):
    L = kzp.lib()
    nv, npub, dom, nc, st = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_int()
    args = (ctypes.byref(nv), ctypes.byref(npub), ctypes.byref(dom), ctypes.byref(nc), ctypes.byref(st))
    toy = os.path.join(GOLDEN, "toy", "toy_1.zkey")
    assert L.kzp_host_parse_zkey(toy.encode(), *args) == 0
    assert (nv.value, npub.value, dom.value, nc.value, st.value) == (3, 1, 4, 4, 0)
    syn = os.path.join(GOLDEN, "syn256", "syn256.zkey")
    assert L.kzp_host_parse_zkey(syn.encode(), *args) == 0
    zk = oracle.read_zkey(syn)
    assert (nv.value, npub.value, dom.value, nc.value) == (zk.n_vars, zk.n_public, zk.domain_size, len(zk.coefs))
    # missing file -> ZKEY_FILE_LOAD_ERROR (std::system_error in the reference, fullprover.cpp:96-100)
    assert L.kzp_host_parse_zkey(b"/nonexistent/x.zkey", *args) == 3 and st.value == 1
    # wrong magic / version / protocol / prime -> UNSUPPORTED_ZKEY_CURVE (std::invalid_argument, fullprover.cpp:91-95)
    data = bytearray(open(toy, "rb").read())
    cases = {}
    cases["magic"] = bytes(b"wtns") + bytes(data[4:])
    v = bytearray(data); v[4] = 2; cases["version"] = bytes(v)
    p = bytearray(data); i = data.index(bytes.fromhex("010000f093f5e143")); p[i] ^= 0x02; cases["prime"] = bytes(p)
    cases["truncated"] = bytes(data[:200])
    for name, blob in cases.items():
        path = os.path.join(workdir, "bad_%s.zkey" % name)
        open(path, "wb").write(blob)
        assert L.kzp_host_parse_zkey(path.encode(), *args) == 2, name
        assert st.value == 2, name


def test_everything_refuses_without_gpu(kzp, oracle):
    if kzp.device_count() > 0:
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU-only box")
    with pytest.raises(kzp.ZKeyFileLoadError) as e:
        kzp.FullProver(os.path.join(GOLDEN, "toy", "toy_1.zkey"))
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(kzp.KzpError):
        kzp.fr_ntt(bytes(64))
    with pytest.raises(kzp.KzpError):
        kzp.Msm(0, bytes(64))
    with pytest.raises(kzp.KzpError):
        kzp.field_op(0, 0, bytes(32), bytes(32))
    with pytest.raises(kzp.KzpError):
        kzp.imad_peak(16)
    # the sharded prover (one proof over several GPUs) refuses the same way, through the explicit entry point and
    # through $KZP_SHARD_DEVICES behind the reference-shaped constructor
    with pytest.raises(kzp.ZKeyFileLoadError) as e:
        kzp.FullProver(os.path.join(GOLDEN, "toy", "toy_1.zkey"), devices=[0, 1])
    assert "no CPU fallback" in str(e.value)
    os.environ["KZP_SHARD_DEVICES"] = "0,1"
    try:
        with pytest.raises(kzp.ZKeyFileLoadError):
            kzp.FullProver(os.path.join(GOLDEN, "toy", "toy_1.zkey"))
    finally:
        del os.environ["KZP_SHARD_DEVICES"]


def _partials_from_oracle(oracle, zk, w, lo_hi):
    """What one shard must produce, computed by the oracle: MSM sums over the shard's base ranges, XYZZ-encoded."""
    (alo, ahi), (clo, chi), (hlo, hhi), h = lo_hi
    A = oracle.msm_naive_g1(zk.points_a[alo:ahi], w[alo:ahi])
    B1 = oracle.msm_naive_g1(zk.points_b1[alo:ahi], w[alo:ahi])
    B2 = oracle.msm_naive_g2(zk.points_b2[alo:ahi], w[alo:ahi])
    cw = w[zk.n_public + 1:]
    C = oracle.msm_naive_g1(zk.points_c[clo:chi], cw[clo:chi])
    H = oracle.msm_naive_g1(zk.points_h[hlo:hhi], h[hlo:hhi])
    one = oracle.le32(oracle.to_mont(1, oracle.Q_MOD))

    def g1(p):
        return (oracle.g1_to_zkey_bytes(p) + one + one) if p is not None else (one + one + bytes(64))

    def g2(p):
        return (oracle.g2_to_zkey_bytes(p) + one + bytes(32) + one + bytes(32)) if p is not None else \
            (one + bytes(32) + one + bytes(32) + bytes(128))

    return g1(A) + g1(B1) + g1(C) + g1(H) + g2(B2)


def test_host_assemble_of_sharded_partials(kzp, oracle):
    """Rank-0 step of the sharded mode with partials supplied by the oracle: the product's host
    assembly must reproduce the reference's proof bytes for 1, 2 and 3 shards."""
    d = os.path.join(GOLDEN, "syn256")
    exp = json.load(open(os.path.join(d, "expected.json")))
    zk = oracle.read_zkey(os.path.join(d, "syn256.zkey"))
    w = oracle.read_wtns(os.path.join(d, "syn256.wtns"))
    h = [oracle.from_le(bytes.fromhex(exp["h"])[i * 32:(i + 1) * 32]) for i in range(zk.domain_size)]
    r, s = bytes.fromhex(exp["r"]), bytes.fromhex(exp["s"])
    for world in (1, 2, 3):
        parts = []
        for k in range(world):
            rng = lambda n: (k * n // world, (k + 1) * n // world)
            parts.append(_partials_from_oracle(oracle, zk, w, (rng(zk.n_vars), rng(zk.n_vars - zk.n_public - 1),
                                                                  rng(zk.domain_size), h)))
        js, msm = kzp.host_assemble(os.path.join(d, "syn256.zkey"), parts, r, s)
        assert js == exp["proof"], world
        assert msm.hex() == exp["msm"], world
    # fresh randomness: still a valid proof
    js, _ = kzp.host_assemble(os.path.join(d, "syn256.zkey"), parts, None, None)
    assert js != exp["proof"]
    pa, pb, pc = oracle.proof_from_json(js)
    assert oracle.groth16_verify(oracle.vk_from_zkey(zk), exp["public"], pa, pb, pc)


def test_cxx_abi_header_compiles_and_links(kzp, workdir):
    """A C++ translation unit written against include/fullprover_b200.hpp links against libkzp_b200.so and sees the
    reference's object layouts; without a GPU the constructor reports a state instead of throwing."""
    src = os.path.join(workdir, "abi_check.cpp")
    exe = os.path.join(workdir, "abi_check")
    open(src, "w").write(r'''
#include <cstdio>
#include <cstring>
#include "fullprover_b200.hpp"
int main(int argc, char** argv) {
    static_assert(sizeof(FullProver) == 16, "FullProver size");
    static_assert(sizeof(ProverResponse) == 24, "ProverResponse size");
    FullProver p(argv[1]);
    int state; memcpy(&state, reinterpret_cast<char*>(&p) + 8, 4);
    ProverResponse r = p.prove(argc > 2 ? argv[2] : "/nonexistent.wtns");
    printf("%d %d %d %d %s\n", state, (int)r.type, (int)r.error, r.metrics.prover_time, r.raw_json);
    return 0;
}
''')
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           kzp.LIB_PATH, "-Wl,-rpath," + os.path.dirname(kzp.LIB_PATH)])
    out = subprocess.check_output([exe, "/nonexistent/k.zkey"], text=True).split()
    # state ZKEY_FILE_LOAD_ERROR(1), response ERROR(1), error PROVER_NOT_READY(1)
    assert out[:3] == ["1", "1", "1"]


REF_HEADER_DIR = "/root/reference/rust-rapidsnark/rapidsnark/src"


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_HEADER_DIR, "fullprover.hpp")),
                    reason="the reference tree is only mounted in the build container")
def test_cxx_abi_against_reference_header(kzp, workdir):
    """A translation unit compiled against the REFERENCE'S OWN rust-rapidsnark/rapidsnark/src/fullprover.hpp (what
    bindgen reads, build.rs:183-206) links against libkzp_b200.so: every constructor, destructor and prove() symbol the
    reference header declares resolves in the library, the object layouts it implies are the ones the library was
    built for, and the no-GPU behaviour is the reference's (state instead of exception, PROVER_NOT_READY)."""
    src = os.path.join(workdir, "abi_ref.cpp")
    exe = os.path.join(workdir, "abi_ref")
    open(src, "w").write(r'''
#include <cstddef>
#include <cstdio>
#include <cstring>
#include "fullprover.hpp"   // the reference's header, not the repo's re-declaration
static_assert(sizeof(FullProver) == 16, "FullProver size");
static_assert(sizeof(ProverResponse) == 24, "ProverResponse size");
static_assert(offsetof(ProverResponse, raw_json) == 8 && offsetof(ProverResponse, error) == 16 &&
              offsetof(ProverResponse, metrics) == 20, "ProverResponse member offsets");
static_assert(sizeof(ProverResponseMetrics) == 4, "metrics");
static_assert((int)SUCCESS == 0 && (int)ERROR == 1, "ProverResponseType");
static_assert((int)OK == 0 && (int)ZKEY_FILE_LOAD_ERROR == 1 && (int)UNSUPPORTED_ZKEY_CURVE == 2, "FullProverState");
static_assert((int)NONE == 0 && (int)PROVER_NOT_READY == 1 && (int)INVALID_INPUT == 2 &&
              (int)WITNESS_GENERATION_INVALID_CURVE == 3, "ProverError");
int main(int argc, char** argv) {
    FullProver p(argv[1]);
    int state; memcpy(&state, reinterpret_cast<char*>(&p) + 8, 4);
    ProverResponse r = p.prove("/nonexistent.wtns");
    ProverResponse e(INVALID_INPUT);
    printf("%d %d %d %d [%s] %d\n", state, (int)r.type, (int)r.error, r.metrics.prover_time, r.raw_json, (int)e.error);
    return 0;
}
''')
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", REF_HEADER_DIR, src, "-o", exe,
                           kzp.LIB_PATH, "-Wl,-rpath," + os.path.dirname(kzp.LIB_PATH)])
    # the executable's undefined FullProver / ProverResponse symbols are exactly the library's exported ones
    und = {l.split()[-1] for l in subprocess.check_output(["nm", "-u", exe], text=True).splitlines()
           if "FullProver" in l or "ProverResponse" in l}
    dyn = {l.split()[-1] for l in subprocess.check_output(["nm", "-D", "--defined-only", kzp.LIB_PATH], text=True).splitlines()}
    assert und and und <= dyn, und - dyn
    assert {"_ZN10FullProverC1EPKc", "_ZN10FullProverD1Ev", "_ZNK10FullProver5proveEPKc"} <= und
    if kzp.device_count() == 0:
        out = subprocess.check_output([exe, "/nonexistent/k.zkey"], text=True).split()
        assert out[:3] == ["1", "1", "1"] and out[-1] == "2"


def test_pool_checkout_queue(kzp):
    """The prover pool's checkout (csrc/pool.hpp, host-only): a slot never has two holders, every job is served,
    work spreads over the slots, and callers beyond the slot count queue up."""
    per_slot, worst, deepest = kzp.pool_sched_selftest(4, 16, 40, 100)
    assert worst == 1
    assert sum(per_slot) == 16 * 40
    assert min(per_slot) >= 0.8 * (16 * 40 / 4)  # least-recently-released first: even spread
    assert 4 < deepest <= 16
    per_slot, worst, _ = kzp.pool_sched_selftest(1, 8, 25, 20)
    assert per_slot == [200] and worst == 1


def test_pool_without_gpu_reports_load_error(kzp):
    if kzp.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(kzp.ZKeyFileLoadError):
        kzp.ProverPool(os.path.join(GOLDEN, "toy", "toy_1.zkey"))
    with pytest.raises(kzp.ZKeyFileLoadError):
        kzp.ProverPool(os.path.join(GOLDEN, "toy", "toy_1.zkey"), devices=[0, 0])


def test_cli_exit_codes_without_proving(kzp, workdir):
    """kzp_prove (csrc/cli_main.cpp): usage and unusable-zkey paths need no GPU; without a device the prover refuses
    to run (exit 2) instead of falling back to anything."""
    assert os.path.exists(kzp.CLI_PATH), "run keyless-zk-proofs_b200/build.py"
    r = subprocess.run([kzp.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
    out = [os.path.join(workdir, "cli_proof.json"), os.path.join(workdir, "cli_public.json")]
    r = subprocess.run([kzp.CLI_PATH, os.path.join(workdir, "nope.zkey"), "x.wtns"] + out, capture_output=True, text=True)
    assert r.returncode == 2 and "zkey not usable" in r.stderr
    if kzp.device_count() == 0:
        toy = os.path.join(GOLDEN, "toy")
        r = subprocess.run([kzp.CLI_PATH, os.path.join(toy, "toy_1.zkey"), os.path.join(toy, "toy.wtns")] + out,
                           capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr
        assert not os.path.exists(out[0])


# ---------------------------------------------------------------- native verifier (csrc/pairing.hpp, host only)
def test_pairing_check_bilinearity(kzp, oracle):
    """prod e(P_i, Q_i) == 1 on the host: bilinear in both arguments, non-degenerate, infinity-neutral; inputs off the
    curve are refused. Points come from the Python oracle's scalar multiplication."""
    o, rnd = oracle, random.Random(21)
    g1, g2 = o.g1_to_zkey_bytes, o.g2_to_zkey_bytes
    G1, G2 = o.G1_GEN, o.G2_GEN
    for _ in range(2):
        a, b = rnd.randrange(1, o.R_MOD), rnd.randrange(1, o.R_MOD)
        aP, bQ = o.g1_mul(G1, a), o.g2_mul(G2, b)
        abP = o.g1_mul(G1, a * b % o.R_MOD)
        assert kzp.host_pairing_check(g1(aP) + g1(o.g1_neg(abP)), g2(bQ) + g2(G2))          # e(aP,bQ) = e(abP,Q)
        assert kzp.host_pairing_check(g1(aP) + g1(o.g1_neg(G1)), g2(bQ) + g2(o.g2_mul(G2, a * b % o.R_MOD)))
        assert not kzp.host_pairing_check(g1(aP) + g1(o.g1_neg(o.g1_mul(abP, 2))), g2(bQ) + g2(G2))
    assert not kzp.host_pairing_check(g1(G1), g2(G2))                                           # non-degenerate
    assert kzp.host_pairing_check(g1(G1) + g1(o.g1_neg(G1)), g2(G2) + g2(G2))
    assert kzp.host_pairing_check(bytes(64) + g1(G1), g2(G2) + bytes(128))                       # infinity pairs
    assert kzp.host_pairing_check(b"", b"")
    bad = bytearray(g1(G1))
    bad[0] ^= 1
    with pytest.raises(kzp.KzpError):
        kzp.host_pairing_check(bytes(bad), g2(G2))


@pytest.mark.parametrize("name,zkey", [("toy", "toy_1.zkey"), ("syn256", "syn256.zkey")])
def test_host_verify_agrees_with_oracle(kzp, oracle, name, zkey):
    """kzp_host_verify on the reference's recorded proofs: accepts them with the right public input, rejects a wrong
    public input, a swapped pi_a/pi_c and a tampered coordinate — the same verdicts as the oracle's verifier."""
    d = os.path.join(GOLDEN, name)
    exp = json.load(open(os.path.join(d, "expected.json")))
    z = os.path.join(d, zkey)
    vk = oracle.vk_from_zkey(oracle.read_zkey(z))
    proof = exp["proof"]
    cases = [(proof, exp["public"]), (proof, [exp["public"][0] + 1])]
    js = json.loads(proof)
    swapped = dict(js, pi_a=js["pi_c"], pi_c=js["pi_a"])
    cases.append((json.dumps(swapped, separators=(",", ":")), exp["public"]))
    neg = dict(js, pi_a=[js["pi_a"][0], str(oracle.Q_MOD - int(js["pi_a"][1])), "1"])  # -A: on the curve, wrong proof
    cases.append((json.dumps(neg, separators=(",", ":")), exp["public"]))
    for pj, pub in cases:
        want = oracle.groth16_verify(vk, pub, *oracle.proof_from_json(pj))
        assert kzp.host_verify(z, pj, pub) == want
    assert kzp.host_verify(z, proof, exp["public"])
    off = dict(js, pi_c=[js["pi_c"][0], str((int(js["pi_c"][1]) + 1) % oracle.Q_MOD), "1"])     # off the curve
    assert not kzp.host_verify(z, json.dumps(off), exp["public"])
    with pytest.raises(kzp.KzpError):
        kzp.host_verify(z, proof, exp["public"] + [1])
    with pytest.raises(kzp.KzpError):
        kzp.host_verify(z, "{}", exp["public"])
    with pytest.raises(kzp.KzpError):
        kzp.host_verify(z, proof.replace(js["pi_a"][0], str(oracle.Q_MOD)), exp["public"])


def test_witness_packing_round_trip(kzp, oracle):
    """The host half of the packed witness transfer (pack_values in csrc/prover.cu): values below 256 travel as one
    byte, the rest as 32 bytes behind a flag bit. Re-expanding the packed slice in Python gives the input back, for
    every classification boundary and for counts that are not multiples of the 128-value inner step."""
    rnd = random.Random(33)
    edge = [0, 1, 255, 256, 257, 0xFFFF, 1 << 8, 1 << 63, 1 << 64, 1 << 128, 1 << 248, (1 << 248) + 1, oracle.R_MOD - 1,
            (1 << 256) - 1, 255 << 248, 1 << 255]
    for count in (1, 3, 127, 128, 129, 4095, 4096, 4097, 10000, 32768):
        vals = [rnd.choice(edge) if rnd.random() < 0.5 else (rnd.randrange(256) if rnd.random() < 0.7 else rnd.randrange(1 << 256))
                for _ in range(count)]
        small, flags, full, moved = kzp.host_pack_witness_slice(b"".join(v.to_bytes(32, "little") for v in vals))
        n_full = sum(1 for v in vals if v >= 256)
        assert len(full) == 32 * n_full and moved == 32768 + 4096 + 32 * n_full
        k = 0
        for i, v in enumerate(vals):
            is_full = (flags[i >> 3] >> (i & 7)) & 1
            assert is_full == (1 if v >= 256 else 0), (count, i)
            if is_full:
                assert int.from_bytes(full[32 * k:32 * k + 32], "little") == v
                assert small[i] == 0
                k += 1
            else:
                assert small[i] == v
        # padding of the last 128-value group is classified as zeros
        pad_end = (count + 127) // 128 * 128
        assert all(b == 0 for b in small[count:pad_end])
        assert all(((flags[i >> 3] >> (i & 7)) & 1) == 0 for i in range(count, pad_end))
    with pytest.raises(kzp.KzpError):
        kzp.host_pack_witness_slice(bytes(32 * 32769))


def _require_sanitizer(flag, workdir):
    """Skips when this machine cannot build or run a trivial program under the sanitizer (missing runtime, or a kernel
    whose address-space layout the runtime refuses): the tests below are about OUR code, not about the toolchain."""
    src = os.path.join(workdir, "san_probe.cpp")
    exe = os.path.join(workdir, "san_probe")
    open(src, "w").write("#include <thread>\nint x; int main() { std::thread t([] { x = 1; }); t.join(); return x - 1; }\n")
    ok = subprocess.run(["g++", "-std=c++17", flag, src, "-o", exe, "-lpthread"], capture_output=True).returncode == 0
    ok = ok and subprocess.run([exe], capture_output=True, timeout=60).returncode == 0
    if not ok:
        pytest.skip("g++ %s does not work on this machine" % flag)


def _mutated_zkeys(base: bytes, rnd, count):
    """Seeded corruptions of a zkey image: bit flips, extreme header fields, truncations, extensions."""
    head = min(len(base), 1200)  # container header, section 1 and 2, start of the IC section
    for it in range(count):
        blob = bytearray(base)
        kind = it % 6
        if kind == 0:    # a few bit flips in the headers
            for _ in range(rnd.randrange(1, 4)):
                blob[rnd.randrange(head)] ^= 1 << rnd.randrange(8)
        elif kind == 1:  # an extreme 32-bit value at a 4-byte aligned header position
            pos = 4 * rnd.randrange(head // 4 - 1)
            blob[pos:pos + 4] = rnd.choice([0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE]).to_bytes(4, "little")
        elif kind == 2:  # an extreme 64-bit section size somewhere in the section table walk
            pos = rnd.randrange(8, head - 8)
            blob[pos:pos + 8] = rnd.choice([0, 0xFFFFFFFFFFFFFFFF, 0x8000000000000000, len(base), len(base) + 1]).to_bytes(8, "little")
        elif kind == 3:  # truncation
            blob = blob[:rnd.randrange(0, len(base))]
        elif kind == 4:  # random bytes anywhere
            for _ in range(rnd.randrange(1, 16)):
                blob[rnd.randrange(len(blob))] = rnd.randrange(256)
        else:            # trailing garbage
            blob += bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 64)))
        yield bytes(blob)


def test_file_parsers_survive_mutated_inputs(kzp, workdir):
    """Seeded fuzz of the bounds-checked loaders (csrc/binfile.hpp; the reference asserts instead,
    binfile_utils.cpp:21,143): corrupted copies of the reference's toy zkey and of the generated 256-wire zkey go
    through kzp_host_parse_zkey and kzp_host_verify (which also walks the IC section). Every outcome must be an error
    code or a verdict — never a crash — and an accepted header must describe sections that really fit in the file."""
    L = kzp.lib()
    nv, npub, dom, nc, st = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_int()
    args = (ctypes.byref(nv), ctypes.byref(npub), ctypes.byref(dom), ctypes.byref(nc), ctypes.byref(st))
    rnd = random.Random(20261018)
    path = os.path.join(workdir, "fuzz.zkey")
    n_ok = n_err = 0
    for name, zkey in (("toy", "toy_1.zkey"), ("syn256", "syn256.zkey")):
        d = os.path.join(GOLDEN, name)
        base = open(os.path.join(d, zkey), "rb").read()
        exp = json.load(open(os.path.join(d, "expected.json")))
        pub = b"".join(int(v).to_bytes(32, "little") for v in exp["public"])
        for it, blob in enumerate(_mutated_zkeys(base, rnd, 300)):
            open(path, "wb").write(blob)
            rc = L.kzp_host_parse_zkey(path.encode(), *args)
            assert rc in (0, 2, 3), (name, it, rc)
            if rc == 0:
                n_ok += 1
                assert st.value == 0 and dom.value & (dom.value - 1) == 0 and npub.value + 1 <= nv.value
                # the sections an accepted header promises fit in the file
                assert 64 * 2 * nv.value + 128 * nv.value + 64 * dom.value + 44 * nc.value <= len(blob)
            else:
                n_err += 1
                assert st.value in (1, 2)
            res = ctypes.c_int(-1)
            rc = L.kzp_host_verify(path.encode(), exp["proof"].encode(), pub, len(exp["public"]), ctypes.byref(res))
            assert rc in (0, 2, 3) and (rc != 0 or res.value in (0, 1)), (name, it, rc, res.value)
    assert n_ok > 50 and n_err > 50, (n_ok, n_err)  # both sides of the parser were exercised


def test_loaders_and_verifier_clean_under_sanitizers(kzp, workdir):
    """The same corrupted files through csrc/verify.cpp (binfile.hpp loaders + IC walk + pairing.hpp) rebuilt with
    -fsanitize=address,undefined: no out-of-bounds read, no undefined shift or overflow in the 4x64-bit host field
    code, and the instrumented build gives the same return codes and verdicts as the shipped library."""
    _require_sanitizer("-fsanitize=address,undefined", workdir)
    exe = os.path.join(workdir, "verify_fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined",
                           "-fno-sanitize-recover=undefined",
                           os.path.join(ROOT, "keyless-zk-proofs_b200", "csrc", "verify.cpp"),
                           os.path.join(ROOT, "tests", "harness", "verify_fuzz_main.cpp"), "-o", exe])
    L = kzp.lib()
    rnd = random.Random(7)
    for name, zkey in (("toy", "toy_1.zkey"), ("syn256", "syn256.zkey")):
        d = os.path.join(GOLDEN, name)
        exp = json.load(open(os.path.join(d, "expected.json")))
        pub = b"".join(int(v).to_bytes(32, "little") for v in exp["public"])
        proof_path, pub_path = os.path.join(workdir, "san_proof.json"), os.path.join(workdir, "san_pub.bin")
        open(proof_path, "w").write(exp["proof"])
        open(pub_path, "wb").write(pub)
        files = [os.path.join(d, zkey)]
        for it, blob in enumerate(_mutated_zkeys(open(files[0], "rb").read(), rnd, 120)):
            files.append(os.path.join(workdir, "san_%s_%d.zkey" % (name, it)))
            open(files[-1], "wb").write(blob)
        r = subprocess.run([exe, proof_path, pub_path] + files, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
        lines = r.stdout.split("\n")
        assert lines[0] == "0 1"  # the intact key accepts the reference's recorded proof
        for f, line in zip(files, lines):
            res = ctypes.c_int(-1)
            rc = L.kzp_host_verify(f.encode(), exp["proof"].encode(), pub, len(exp["public"]), ctypes.byref(res))
            assert line == "%d %d" % (rc, res.value if rc == 0 else 0) or (rc != 0 and line.split()[0] == str(rc)), (f, line, rc)
        assert "pairing 0 1" in lines and "pairing 2 -1" in lines
        for f in files[1:]:
            os.unlink(f)


def test_pool_under_thread_sanitizer(workdir):
    """csrc/pool.cpp itself (checkout queue, retirement of a faulted prover, fused verify-before-return, the drain in
    kzp_pool_free) rebuilt with -fsanitize=thread over stub provers (tests/harness/pool_tsan_main.cpp): 16 client
    threads x 40 requests over 6 provers of which two fault, then five rounds of freeing a pool with 10 callers queued
    behind 2 busy provers. No data race, no two callers inside one prover, no prover freed under a caller, every
    request answered with the documented code."""
    _require_sanitizer("-fsanitize=thread", workdir)
    exe = os.path.join(workdir, "pool_tsan")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread",
                           os.path.join(ROOT, "keyless-zk-proofs_b200", "csrc", "pool.cpp"),
                           os.path.join(ROOT, "tests", "harness", "pool_tsan_main.cpp"), "-o", exe, "-lpthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (r.stdout, r.stderr[-3000:])
    assert "healthy 4 total 640" in r.stdout and "failures 0" in r.stdout


def test_shipped_binary_is_sm100a_with_copy_engine_staged_ntt(kzp):
    """Static properties of the shipped libkzp_b200.so, read with cuobjdump (no GPU): every cubin with code is sm_100a;
    the NTT level kernels fetch their tiles with tensor copies (UTMALDG) and the fused middle level with a bulk copy
    (UBLKCP), both completed through an mbarrier (SYNCS.ARRIVE / SYNCS.PHASECHK) — north_star's "TMA-staged" NTT;
    the field arithmetic is on the wide integer multiply-add (IMAD.WIDE); the accumulate kernels keep their register
    budgets (G1 <= 128: four CTAs of 128 threads per SM; G2 <= 168: three)."""
    import shutil

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run(["cuobjdump", "--list-elf", kzp.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elf))
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", kzp.LIB_PATH], capture_output=True, text=True).stdout
    regs, arch_of, arch = {}, {}, None
    fn = None
    for line in res.split("\n"):
        m = re.match(r"\s*arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            arch_of[fn] = arch
        m = re.match(r"\s*REG:(\d+)", line)
        if m and fn:
            regs[fn] = int(m.group(1))
            fn = None
    assert regs and set(arch_of.values()) == {"sm_100a"}, (archs, set(arch_of.values()))
    sass = subprocess.run(["cuobjdump", "-sass", kzp.LIB_PATH], capture_output=True, text=True).stdout
    ops, cur = {}, None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            ops[cur] = {}
            continue
        if cur:
            for k in ("UTMALDG", "UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "IMAD.WIDE"):
                if k in line:
                    ops[cur][k] = ops[cur].get(k, 0) + 1
    tma = [f for f in ops if "k_ntt_level_tma" in f]
    mid = [f for f in ops if "k_ntt_mid" in f]
    assert len(tma) == 4 and len(mid) == 2
    for f in tma:
        assert ops[f].get("UTMALDG", 0) >= 1 and ops[f].get("SYNCS.ARRIVE", 0) >= 1 and ops[f].get("SYNCS.PHASECHK", 0) >= 1, (f, ops[f])
    for f in mid:
        assert ops[f].get("UBLKCP", 0) >= 1 and ops[f].get("SYNCS.PHASECHK", 0) >= 1, (f, ops[f])
    for f in tma + mid:
        assert ops[f]["IMAD.WIDE"] > 3000  # the butterflies' Montgomery products
    acc = {f: r for f, r in regs.items() if "k_msm_accumulate" in f}
    assert acc
    for f, r in acc.items():
        g2 = "Fp2T" in f
        assert r <= (168 if g2 else 128), (f, r)
