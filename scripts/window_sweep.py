"""G1 MSM time over (size, window bits, chunk) for uniform scalars: picks the H-MSM shape of a sharded proof's slices.
    python scripts/window_sweep.py [--sizes 17,18,19,20] [--windows 16,17,18,19,20] [--chunks 32,64]"""
import argparse, ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench, microbench
import keyless_zk_proofs_b200 as kzp
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="17,18,19,20")
ap.add_argument("--windows", default="16,17,18,19,20")
ap.add_argument("--chunks", default="0,32,64")
a = ap.parse_args()
gen = bench.ensure_setupgen()
gen.kzp_gen_consecutive_points.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p]
rng = np.random.default_rng(5)
s0b = ((0x1234567 << 128) + 0xDEADBEEF).to_bytes(32, "little")
for lg in [int(x) for x in a.sizes.split(",")]:
    n = 1 << lg
    bases = ctypes.create_string_buffer(n * 64)
    assert gen.kzp_gen_consecutive_points(0, n, s0b, bases) == 0
    sc = microbench.uniform_scalars(n, rng).tobytes()
    ref = None
    for c in [int(x) for x in a.windows.split(",")]:
        for ch in [int(x) for x in a.chunks.split(",")]:
            os.environ["KZP_MSM_CHUNK"] = str(ch)
            m = kzp.Msm(0, bases, window_bits=c, two_level=True)
            got = m.run(sc)
            ref = ref or got
            ms, ent = m.bench(sc, 10)
            print(json.dumps({"log_n": lg, "c": c, "chunk": ch, "ms": round(ms, 4), "entries": ent, "same": got == ref}), flush=True)
            m.close()
