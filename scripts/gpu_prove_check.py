import ctypes, os, random, sys, time, json, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bn254 as o
import keyless_zk_proofs_b200 as kzp
random.seed(5)
ref = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libkzp_ref.so"))
ref.kzp_ref_prove.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
def ref_prove(zk, wt, r, s):
    buf = ctypes.create_string_buffer(4096); tm = ctypes.c_int()
    rc = ref.kzp_ref_prove(zk.encode(), wt.encode(), r, s, 1, buf, 4096, ctypes.byref(tm)); assert rc == 0
    return buf.value.decode()
r = o.le32(random.randrange(o.R_MOD >> 2)); s = o.le32(random.randrange(o.R_MOD >> 2))
toy = os.path.join(ROOT, "tests/golden/toy/")
print("creating prover", flush=True)
p = kzp.FullProver(toy + "toy_1.zkey")
print("created", p.n_vars, p.domain_size, flush=True)
js, met = p.prove(toy + "toy.wtns", r, s)
print(js, flush=True)
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
