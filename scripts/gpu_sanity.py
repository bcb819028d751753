"""Scratch GPU sanity run (superseded by tests/ -m gpu): field ops, NTT, MSM, toy + small synthetic proof."""
import ctypes, os, random, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bn254 as o
import keyless_zk_proofs_b200 as kzp

random.seed(5)
print("devices:", kzp.device_count(), flush=True)
ops, t = kzp.imad_peak(2048); print("IMAD peak: %.2f T mad/s (%.3f ms)" % (ops / 1e12, t), flush=True)

def fld(field, mod):
    n = 512
    a = [random.randrange(mod) for _ in range(n)]; b = [random.randrange(mod) for _ in range(n)]
    a[0] = 0; b[1] = 0; a[2] = mod - 1; b[2] = mod - 1; a[3] = 1; b[4] = a[4]
    A = b"".join(map(o.le32, a)); B = b"".join(map(o.le32, b))
    exp = {0: lambda x, y: o.mont_mul(x, y, mod), 1: lambda x, y: (x + y) % mod, 2: lambda x, y: (x - y) % mod,
           3: lambda x, y: (-x) % mod, 4: lambda x, y: o.to_mont(x, mod), 5: lambda x, y: o.from_mont(x, mod),
           6: lambda x, y: o.mont_mul(x, x, mod)}
    for op, fn in exp.items():
        out = kzp.field_op(field, op, A, B)
        got = [o.from_le(out[i * 32:(i + 1) * 32]) for i in range(n)]
        want = [fn(x, y) for x, y in zip(a, b)]
        bad = sum(g != w for g, w in zip(got, want))
        print("field", field, "op", op, "mismatches", bad, flush=True)
        assert bad == 0
    out = kzp.field_op(field, 7, A[:32 * 16], None)
    for i in range(16):
        g = o.from_le(out[i * 32:(i + 1) * 32])
        w = o.to_mont(pow(o.from_mont(a[i], mod), -1, mod), mod) if a[i] else 0
        assert g == w, ("inv", i)
    print("field", field, "inv ok", flush=True)

fld(0, o.R_MOD); fld(1, o.Q_MOD)

# NTT
for logn in (0, 1, 2, 3, 6, 10):
    n = 1 << logn
    x = [random.randrange(o.R_MOD) for _ in range(n)]
    X = b"".join(o.le32(o.to_mont(v, o.R_MOD)) for v in x)
    f = kzp.fr_ntt(X, False); want = o.fr_fft(x) if n > 1 else x
    got = [o.from_mont(o.from_le(f[i * 32:(i + 1) * 32]), o.R_MOD) for i in range(n)]
    assert got == want, ("fft", logn)
    g = kzp.fr_ntt(f, True)
    assert g == X, ("ifft", logn)
    print("ntt ok", logn, flush=True)

# MSM vs reference
ref = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libkzp_ref.so"))
def ref_msm(group, bases, scalars, n):
    out = ctypes.create_string_buffer(64 if group == 0 else 128)
    (ref.kzp_ref_msm_g1 if group == 0 else ref.kzp_ref_msm_g2)(bases, scalars, ctypes.c_uint64(n), out)
    return out.raw
t0 = time.time()
npts = 3000
ks = [random.randrange(1, o.R_MOD) for _ in range(npts)]
pts = [o.g1_mul(o.G1_GEN, k) for k in ks]
pts[5] = None; pts[7] = pts[6]; pts[9] = o.g1_neg(pts[8])
bases = b"".join(o.g1_to_zkey_bytes(p) for p in pts)
print("gen g1 points", time.time() - t0, flush=True)
for name, sc in (("uniform", [random.randrange(o.R_MOD) for _ in range(npts)]),
                 ("bits", [random.randrange(2) for _ in range(npts)]),
                 ("mixed", [random.choice([0, 1, random.randrange(256), random.randrange(o.R_MOD), 0x8000, 0x8001, 0xffff, 0x10000]) for _ in range(npts)]),
                 ("zeros", [0] * npts)):
    if name == "mixed":
        sc[6] = sc[7] = 12345; sc[8] = sc[9] = 77
    S = b"".join(map(o.le32, sc))
    for n in (npts, 1, 2, 37):
        m = kzp.Msm(0, bases[:64 * n]); got = m.run(S[:32 * n]); m.close()
        want = ref_msm(0, bases[:64 * n], S[:32 * n], n)
        print("msm g1", name, n, got == want, flush=True)
        assert got == want
npts2 = 400
pts2 = [o.g2_mul(o.G2_GEN, random.randrange(1, o.R_MOD)) for _ in range(npts2)]
pts2[3] = None
bases2 = b"".join(o.g2_to_zkey_bytes(p) for p in pts2)
for name, sc in (("uniform", [random.randrange(o.R_MOD) for _ in range(npts2)]), ("bits", [random.randrange(2) for _ in range(npts2)])):
    S = b"".join(map(o.le32, sc))
    m = kzp.Msm(1, bases2); got = m.run(S); m.close()
    want = ref_msm(1, bases2, S, npts2)
    print("msm g2", name, got == want, flush=True)
    assert got == want

# proofs
ref.kzp_ref_prove.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
def ref_prove(zk, wt, r, s):
    buf = ctypes.create_string_buffer(4096); tm = ctypes.c_int()
    rc = ref.kzp_ref_prove(zk.encode(), wt.encode(), r, s, 1, buf, 4096, ctypes.byref(tm)); assert rc == 0
    return buf.value.decode()
r = o.le32(random.randrange(o.R_MOD >> 2)); s = o.le32(random.randrange(o.R_MOD >> 2))
toy = os.path.join(ROOT, "tests/golden/toy/")
p = kzp.FullProver(toy + "toy_1.zkey")
js, met = p.prove(toy + "toy.wtns", r, s)
want = ref_prove(toy + "toy_1.zkey", toy + "toy.wtns", r, s)
print("toy proof match:", js == want, met, p.timings(), flush=True)
assert js == want
p.close()
os.makedirs("/tmp/kzp", exist_ok=True)
r1, w = o.synth_circuit(300, 256, seed=11)
zk, trap = o.trapdoor_setup(r1, seed=11)
o.write_zkey("/tmp/kzp/s.zkey", zk); o.write_wtns("/tmp/kzp/s.wtns", w)
p = kzp.FullProver("/tmp/kzp/s.zkey")
js, met = p.prove("/tmp/kzp/s.wtns", r, s)
want = ref_prove("/tmp/kzp/s.zkey", "/tmp/kzp/s.wtns", r, s)
print("synthetic proof match:", js == want, met, p.timings(), flush=True)
assert js == want
pa, pb, pc = o.proof_from_json(js)
print("verifies:", o.groth16_verify(o.vk_from_zkey(zk), [w[1]], pa, pb, pc), flush=True)
print("ALL OK")
