"""Pins the oracles (oracle/bn254.py, oracle/kzp_port.c, oracle/_ref) to the reference's own golden vectors.
CPU only. Reference vectors restated here cite rust-rapidsnark/rapidsnark/src (RS/)."""
import json
import os
import random

import pytest

from conftest import GOLDEN


# ---------------------------------------------------------------- field KATs (RS/test_prover.cpp, extracted)
def _kats():
    return json.load(open(os.path.join(GOLDEN, "field_kats.json")))


def _math(o, rec):
    mod = o.R_MOD if rec["field"] == "Fr" else o.Q_MOD
    a, b = int(rec["a"], 16), int(rec.get("b", "0x0"), 16)
    op = rec["op"]
    want = {"mul": lambda: o.mont_mul(a, b, mod), "square": lambda: o.mont_mul(a, a, mod),
            "add": lambda: (a + b) % mod, "sub": lambda: (a - b) % mod}[op]()
    return mod, a, b, want


def test_field_kats_python_oracle(oracle):
    """Every reference KAT whose inputs are canonical (< p, the only inputs the prover path produces) equals the
    exact modular definition the Python oracle uses. The reference's results on out-of-contract inputs (>= p) are
    artefacts of its reduction strategy and are checked against oracle/_ref only (test below)."""
    n = 0
    for rec in _kats():
        mod, a, b, want = _math(oracle, rec)
        if a < mod and b < mod:
            assert want == int(rec["expected"], 16), rec
            n += 1
    assert n >= 20


def test_field_kats_port_and_ref(oracle, port, ref):
    opcode = {"mul": 0, "add": 1, "sub": 2, "square": 6}
    for rec in _kats():
        mod, a, b, _ = _math(oracle, rec)
        f = 0 if rec["field"] == "Fr" else 1
        got_ref = ref.field_op(f, opcode[rec["op"]], oracle.le32(a), oracle.le32(b))
        assert oracle.from_le(got_ref) == int(rec["expected"], 16), ("reference build disagrees with its own KAT", rec)
        if a < mod and b < mod and rec["op"] != "square":
            got_port = port.field_op(f, opcode[rec["op"]], oracle.le32(a), oracle.le32(b))
            assert oracle.from_le(got_port) == int(rec["expected"], 16), rec


def test_field_constants(oracle):
    # RS/fr_raw_generic.cpp:5-7, RS/fq_raw_generic.cpp:6-8
    assert oracle.R_MOD == 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    assert oracle.Q_MOD == 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    assert pow(oracle.MONT_R, 2, oracle.R_MOD) == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert pow(oracle.MONT_R, 2, oracle.Q_MOD) == 0x06D89F71CAB8351F47AB1EFF0A417FF6B5E71911D44501FBF32CFC5B538AFA89
    assert (-pow(oracle.R_MOD, -1, 1 << 64)) % (1 << 64) == 0xC2E1F593EFFFFFFF
    assert (-pow(oracle.Q_MOD, -1, 1 << 64)) % (1 << 64) == 0x87D20782E4866389
    assert oracle.fr_nqr() == 5  # RS/fft.cpp:60-66
    assert oracle.g1_is_on_curve(oracle.G1_GEN) and oracle.g2_is_on_curve(oracle.G2_GEN)


# ---------------------------------------------------------------- RS/alt_bn128_test.cpp restated
def test_f2_simple_mul(oracle):
    # f2_simpleMul :12-29 : (2+2u)(3+3u) = 12u
    assert oracle.f2_mul((2, 2), (3, 3)) == (0, 12)


def test_g1_small_multiples(oracle):
    # g1_times_3 / g1_times_5 / g1_times_65_exp :60-136 : repeated addition == scalar multiplication
    g = oracle.G1_GEN
    acc = None
    for k in range(1, 66):
        acc = oracle.g1_add(acc, g)
        if k in (3, 5, 65):
            assert acc == oracle.g1_mul(g, k)


def test_exp_to_order(oracle, port):
    # g1_expToOrder / g2_expToOrder :138-170 : r * G = infinity
    assert oracle.g1_mul(oracle.G1_GEN, oracle.R_MOD - 1) == oracle.g1_neg(oracle.G1_GEN)
    assert oracle.g1_add(oracle.g1_mul(oracle.G1_GEN, oracle.R_MOD - 1), oracle.G1_GEN) is None
    assert oracle.g2_add(oracle.g2_mul(oracle.G2_GEN, oracle.R_MOD - 1), oracle.G2_GEN) is None
    assert port.g1_gen_mul(oracle.le32(oracle.R_MOD)) == bytes(64)
    assert port.g2_gen_mul(oracle.le32(oracle.R_MOD)) == bytes(128)


def test_multiexp_sum_of_squares(oracle, port, ref):
    # multiExp :172-212 : bases (i+1)G, scalars i+1  =>  (sum (i+1)^2) G ; reference size n = 40000
    n = 40000
    pts = []
    acc = None
    for i in range(n):
        acc = oracle.g1_add(acc, oracle.G1_GEN)
        pts.append(acc)
    bases = b"".join(oracle.g1_to_zkey_bytes(p) for p in pts)
    scalars = b"".join(oracle.le32(i + 1) for i in range(n))
    want = oracle.g1_to_canonical_bytes(oracle.g1_mul(oracle.G1_GEN, sum((i + 1) ** 2 for i in range(n))))
    assert port.msm(0, bases, scalars) == want
    assert ref.msm(0, bases, scalars) == want
    m = 300  # the pure-Python restatement of the Pippenger loop, small
    assert oracle.g1_to_canonical_bytes(oracle.msm_g1(pts[:m], list(range(1, m + 1)))) == \
        oracle.g1_to_canonical_bytes(oracle.g1_mul(oracle.G1_GEN, sum((i + 1) ** 2 for i in range(m))))


def test_multiexp2_kat(oracle, port, ref):
    # multiExp2 :215-248 : 2-point MSM with decimal coordinates
    b0 = (1626275109576878988287730541908027724405348106427831594181487487855202143055,
          18706364085805828895917702468512381358405767972162700276238017959231481018884)
    b1 = (17245156998235704504461341147511350131061011207199931581281143511105381019978,
          3858908536032228066651712470282632925312300188207189106507111128103204506804)
    s = [1, 20187316456970436521602619671088988952475789765726813868033071292105413408473]
    want = (9163953212624378696742080269971059027061360176019470242548968584908855004282,
            20922060990592511838374895951081914567856345629513259026540392951012456141360)
    assert oracle.msm_g1([b0, b1], s) == want
    assert oracle.msm_naive_g1([b0, b1], s) == want
    bases = oracle.g1_to_zkey_bytes(b0) + oracle.g1_to_zkey_bytes(b1)
    scal = oracle.le32(s[0]) + oracle.le32(s[1])
    assert port.msm(0, bases, scal) == oracle.g1_to_canonical_bytes(want)
    assert ref.msm(0, bases, scal) == oracle.g1_to_canonical_bytes(want)


def test_fft_round_trip(oracle, port, ref):
    # fft :250-271 : a[i] = i+1, n = 2^10, ifft(fft(a)) == a
    n = 1 << 10
    a = list(range(1, n + 1))
    f = oracle.fr_fft(a)
    assert oracle.fr_ifft(f) == a
    data = b"".join(oracle.le32(oracle.to_mont(v, oracle.R_MOD)) for v in a)
    fm = b"".join(oracle.le32(oracle.to_mont(v, oracle.R_MOD)) for v in f)
    assert port.ntt(data) == fm and ref.ntt(data) == fm
    assert port.ntt(fm, inverse=True) == data and ref.ntt(fm, inverse=True) == data


def test_msm_chunk_extraction(oracle):
    # getChunk RS/multiexp.cpp:26-41 incl. the clipped last window
    s = oracle.le32((1 << 256) - 1)
    assert oracle.msm_get_chunk(s, 0, 16) == 0xFFFF
    assert oracle.msm_get_chunk(s, 15, 16) == 0xFFFF
    assert oracle.msm_get_chunk(s, 19, 13) == 0x1FF  # bits 247..255 : 9 effective bits
    assert oracle.msm_window_bits(2) == 2 and oracle.msm_window_bits(40000) == 14 and oracle.msm_window_bits(1 << 21) == 16


# ---------------------------------------------------------------- fixtures: toy triple + syn256
@pytest.mark.parametrize("name,zkey,wtns", [("toy", "toy_1.zkey", "toy.wtns"), ("syn256", "syn256.zkey", "syn256.wtns")])
def test_python_oracle_matches_reference_outputs(oracle, name, zkey, wtns):
    d = os.path.join(GOLDEN, name)
    exp = json.load(open(os.path.join(d, "expected.json")))
    zk = oracle.read_zkey(os.path.join(d, zkey))
    w = oracle.read_wtns(os.path.join(d, wtns))
    r, s = oracle.from_le(bytes.fromhex(exp["r"])), oracle.from_le(bytes.fromhex(exp["s"]))
    art = oracle.groth16_prove(zk, w, r, s)
    assert art.json == exp["proof"]
    assert b"".join(oracle.le32(v) for v in art.h).hex() == exp["h"]
    msm = (oracle.g1_to_canonical_bytes(art.msm_a) + oracle.g1_to_canonical_bytes(art.msm_b1) +
           oracle.g2_to_canonical_bytes(art.msm_b2) + oracle.g1_to_canonical_bytes(art.msm_c) +
           oracle.g1_to_canonical_bytes(art.msm_h))
    assert msm.hex() == exp["msm"]
    ab = b"".join(oracle.le32(oracle.to_mont(v, oracle.R_MOD)) for v in art.a + art.b)
    assert ab.hex() == exp["ab"]
    # and it verifies under the circuit's VK (what prover-service asserts, tests/prover_handler.rs:288)
    assert oracle.groth16_verify(oracle.vk_from_zkey(zk), exp["public"], art.pi_a, art.pi_b, art.pi_c)
    assert not oracle.groth16_verify(oracle.vk_from_zkey(zk), [exp["public"][0] + 1], art.pi_a, art.pi_b, art.pi_c)


def test_toy_vk_json_matches_zkey(oracle):
    d = os.path.join(GOLDEN, "toy")
    zk = oracle.read_zkey(os.path.join(d, "toy_1.zkey"))
    vk = oracle.vk_from_snarkjs_json(json.load(open(os.path.join(d, "toy_vk.json"))))
    assert (vk.alpha1, vk.beta2, vk.gamma2, vk.delta2, vk.ic) == (zk.alpha1, zk.beta2, zk.gamma2, zk.delta2, zk.ic)
    assert oracle.read_wtns(os.path.join(d, "toy.wtns")) == [1, 2, 3]


@pytest.mark.parametrize("name,zkey,wtns,domain", [("toy", "toy_1.zkey", "toy.wtns", 4), ("syn256", "syn256.zkey", "syn256.wtns", 512)])
def test_ref_and_port_reproduce_golden(oracle, ref, port, name, zkey, wtns, domain):
    d = os.path.join(GOLDEN, name)
    exp = json.load(open(os.path.join(d, "expected.json")))
    r, s = bytes.fromhex(exp["r"]), bytes.fromhex(exp["s"])
    z, w = os.path.join(d, zkey), os.path.join(d, wtns)
    js, _ = ref.prove(z, w, r, s)
    assert js == exp["proof"]
    pj, _, ph, pm = port.prove(z, w, r, s, domain, want_artefacts=True)
    assert pj == exp["proof"] and ph.hex() == exp["h"] and pm.hex() == exp["msm"]


def test_port_generator_equals_python_generator(oracle, port, workdir):
    """oracle/kzp_port.c and oracle/bn254.py build byte-identical zkey + wtns from the same seed."""
    zc, wc = os.path.join(workdir, "c.zkey"), os.path.join(workdir, "c.wtns")
    info = port.make_setup(120, 100, 5, zc, wc)
    r1cs, w = oracle.synth_circuit(120, 100, seed=5)
    assert oracle.check_r1cs(r1cs, w)
    zk, trap = oracle.trapdoor_setup(r1cs, seed=5)
    zp, wp = os.path.join(workdir, "p.zkey"), os.path.join(workdir, "p.wtns")
    oracle.write_zkey(zp, zk)
    oracle.write_wtns(wp, w)
    assert open(zc, "rb").read() == open(zp, "rb").read()
    assert open(wc, "rb").read() == open(wp, "rb").read()
    assert info["domain"] == zk.domain_size and info["n_coefs"] == len(zk.coefs)
    # the trapdoor predicts the proof and the port prover reproduces it
    r, s = 987654321, 123456789
    pa, pb, pc = oracle.trapdoor_expected_proof(zk, trap, w, r, s)
    pj, _, _, _ = port.prove(zc, wc, oracle.le32(r), oracle.le32(s))
    assert pj == oracle.proof_json(pa, pb, pc)


def test_port_matches_reference_medium(oracle, port, ref, workdir):
    """2^12-domain keyless-shaped circuit: C port == reference on proof bytes, H and all five MSM results."""
    z, w = os.path.join(workdir, "m12.zkey"), os.path.join(workdir, "m12.wtns")
    info = port.make_setup(4000, 3800, 3, z, w)
    r, s = oracle.le32(random.Random(1).randrange(oracle.R_MOD >> 2)), oracle.le32(random.Random(2).randrange(oracle.R_MOD >> 2))
    rj, _ = ref.prove(z, w, r, s)
    _, rh, rm = ref.dump(z, w, info["domain"])
    pj, _, ph, pm = port.prove(z, w, r, s, info["domain"], want_artefacts=True)
    assert pj == rj and ph == rh and pm == rm
    zk = oracle.read_zkey(z)
    pa, pb, pc = oracle.proof_from_json(rj)
    assert oracle.groth16_verify(oracle.vk_from_zkey(zk), [info["public_input"]], pa, pb, pc)


def test_microbench_generator_matches_oracle(oracle):
    """tools/setupgen.c's micro-benchmark inputs (bases (s0+i)G and the closed-form MSM answer) against the
    Python oracle's scalar multiplication — this is what pins the full-size MSM checks of the GPU suite."""
    import ctypes

    import bench

    lib = bench.ensure_setupgen()
    lib.kzp_gen_consecutive_points.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p]
    lib.kzp_msm_closed_form.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    o, rnd = oracle, random.Random(3)
    cases = [(0, o.G1_GEN, o.g1_mul, o.g1_to_zkey_bytes, o.g1_to_canonical_bytes, 64),
             (1, o.G2_GEN, o.g2_mul, o.g2_to_zkey_bytes, o.g2_to_canonical_bytes, 128)]
    for group, gen, mul, to_zkey, to_canon, sz in cases:
        n, s0 = 4200, rnd.randrange(1 << 250)
        out = ctypes.create_string_buffer(n * sz)
        assert lib.kzp_gen_consecutive_points(group, n, o.le32(s0), out) == 0
        for i in (0, 1, 4095, 4096, 4097, n - 1):  # 4096 = the generator's block size
            assert out.raw[i * sz:(i + 1) * sz] == to_zkey(mul(gen, s0 + i)), (group, i)
        ks = [rnd.randrange(o.R_MOD) for _ in range(n)]
        res = ctypes.create_string_buffer(sz)
        assert lib.kzp_msm_closed_form(group, n, o.le32(s0), b"".join(o.le32(k) for k in ks), res) == 0
        tot = sum(k * (s0 + i) for i, k in enumerate(ks)) % o.R_MOD
        assert res.raw == to_canon(mul(gen, tot)), group


def test_pairing_constants(oracle):
    """The integers csrc/pairing.hpp embeds or relies on: the ate loop count T = t - 1 = 6x^2 (congruent to p mod r),
    the BN parameter x, and the base-p decomposition of the hard part of the final exponentiation."""
    import re

    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "keyless-zk-proofs_b200", "csrc", "pairing.hpp")).read()
    p, r, x = oracle.Q_MOD, oracle.R_MOD, 4965661367192848881
    assert p == 36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1 and r == 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1
    lo, hi = re.search(r"kAteLoop\[2\] = \{0x([0-9a-f]+)ull, 0x([0-9a-f]+)ull\}", src).groups()
    T = int(hi, 16) << 64 | int(lo, 16)
    assert T == 6 * x * x and (T - p) % r == 0 and T.bit_length() == 127
    assert int(re.search(r"kBnX = 0x([0-9a-f]+)ull", src).group(1), 16) == x
    # hard part of the final exponentiation in base p (Devegili-Scott-Dahab), as used by final_exponentiation()
    l3, l2, l1, l0 = 1, 6 * x * x + 1, -36 * x**3 - 18 * x * x - 12 * x + 1, -36 * x**3 - 30 * x * x - 18 * x - 2
    assert (p**4 - p**2 + 1) % r == 0 and l0 + l1 * p + l2 * p**2 + l3 * p**3 == (p**4 - p**2 + 1) // r
    assert (p - 1) % 6 == 0


def test_reference_asm_build_equals_portable_build(oracle, ref, workdir):
    """oracle/_ref/libkzp_ref_asm.so = the reference compiled with its own x86-64 assembly field arithmetic
    (fr.asm / fq.asm rewritten for GNU as by oracle/nasm2gas.py at build time). It is the CPU baseline of bench.py,
    so it must be the same prover: recorded proofs, raw field operations on random and edge inputs, MSM results."""
    import refutil

    asm = refutil.load_ref_asm()
    if asm is None:
        pytest.skip("asm build absent or CPU without ADX/BMI2")
    for name, zkey, wtns in (("toy", "toy_1.zkey", "toy.wtns"), ("syn256", "syn256.zkey", "syn256.wtns")):
        d = os.path.join(GOLDEN, name)
        exp = json.load(open(os.path.join(d, "expected.json")))
        js, _ = asm.prove(os.path.join(d, zkey), os.path.join(d, wtns), bytes.fromhex(exp["r"]), bytes.fromhex(exp["s"]))
        assert js == exp["proof"], name
        _, h, m = asm.dump(os.path.join(d, zkey), os.path.join(d, wtns), len(bytes.fromhex(exp["h"])) // 32)
        assert h.hex() == exp["h"] and m.hex() == exp["msm"], name
    rnd = random.Random(17)
    for field, mod in ((0, oracle.R_MOD), (1, oracle.Q_MOD)):
        vals = [0, 1, 2, mod - 1, mod - 2, (1 << 64) - 1, 1 << 64, 1 << 128, 1 << 253] + [rnd.randrange(mod) for _ in range(500)]
        a = b"".join(oracle.le32(v) for v in vals)
        b = b"".join(oracle.le32(v) for v in reversed(vals))
        for op in (0, 1, 2, 3, 4, 5, 6):
            assert asm.field_op(field, op, a, b) == ref.field_op(field, op, a, b), (field, op)
