#!/usr/bin/env python
"""CPU points beside the GPU micro-benchmark sweep ("CPU points beside the sweep"): the reference's own
Curve::multiMulByScalar and FFT<Fr>::ifft/fft (oracle/_ref, all host cores) timed at a few sizes of the sweep.
Lives under tests/ because it executes oracle/ (checker code): it is a measurement of the REFERENCE, never a product path.

    python tests/ref_cpu_sweep.py [--sizes 16,18,20] [--out gpurun_out/cpu_sweep.json]
"""
import argparse, ctypes, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
import bench, refutil  # noqa: E402
import microbench  # noqa: E402  (scalar generators only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,18,20")
    ap.add_argument("--ntt", default="16,18,20,21")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cpu_sweep.json"))
    a = ap.parse_args()
    ref = refutil.load_ref_asm() or refutil.load_ref()
    assert ref is not None, "oracle/_ref is missing"
    gen = bench.ensure_setupgen()
    gen.kzp_gen_consecutive_points.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p]
    rng = np.random.default_rng(5)
    s0b = ((0x1234567 << 128) + 0xDEADBEEF).to_bytes(32, "little")
    res = {"cores": ref.lib.kzp_ref_num_threads(), "impl": "reference (oracle/_ref), multiexp.cpp / fft.cpp", "msm": [], "ntt": []}
    for group, psz in ((0, 64), (1, 128)):
        for lg in [int(x) for x in a.sizes.split(",")]:
            n = 1 << lg
            bases = ctypes.create_string_buffer(n * psz)
            assert gen.kzp_gen_consecutive_points(group, n, s0b, bases) == 0
            for mix, make in (("uniform", microbench.uniform_scalars), ("keyless-mix", microbench.keyless_mix_scalars)):
                sc = make(n, rng).tobytes()
                ref.msm(group, bases.raw, sc)
                t0 = time.perf_counter()
                ref.msm(group, bases.raw, sc)
                ms = 1e3 * (time.perf_counter() - t0)
                row = {"group": "G1" if group == 0 else "G2", "log_n": lg, "scalars": mix, "ms": ms, "pairs_per_s": n / (ms * 1e-3)}
                res["msm"].append(row)
                print(json.dumps(row), flush=True)
    for lg in [int(x) for x in a.ntt.split(",")]:
        n = 1 << lg
        data = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64).tobytes()
        ref.ntt(data, True)
        t0 = time.perf_counter()
        x = ref.ntt(data, True)
        ref.ntt(x, False)
        ms = 1e3 * (time.perf_counter() - t0)
        row = {"log_n": lg, "what": "ifft + fft (the coset shift between them is n multiplications more)", "ms": ms}
        res["ntt"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
