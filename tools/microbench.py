#!/usr/bin/env python
"""Synthetic BN254 micro-benchmark sweep (BASELINE.json configs[4]):
G1 / G2 MSM and the Fr coset-NTT chain at 2^16 ... 2^24, each against its roofline, through the C ABI.

    python tools/microbench.py [--out gpurun_out/microbench.json] [--g1 16,18,20,21,22,24] [--g2 16,18,20,21,22]
                               [--ntt 16,17,...,24] [--iters 5]

MSM bases are P_i = (s0 + i) * G (distinct multiples of the generator, tools/setupgen.c kzp_gen_consecutive_points);
scalars are (a) uniform below r and (b) the keyless-like mix of the bench workload (84 % bits, 12 % bytes, 4 % full
width). Every uniform-scalar MSM result is CHECKED against the closed form (sum k_i (s0+i) mod r) * G computed on the
host with one fixed-base multiplication, so the sweep is also a full-size parity test (bit-exact affine coordinates).
Rooflines: MSM vs the carry-chained IMAD.WIDE peak measured in the same process (kzp_imad_peak); NTT vs both the HBM
peak of MEASURED_PEAKS.json (algorithmic bytes 2 transforms x 2 x n x 32 B per chain) and the same integer peak.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (ensure_setupgen)
import keyless_zk_proofs_b200 as kzp  # noqa: E402

R_TOP = 0x30644E72E131A029  # top limb of r: limbs below it free => value < r
FQ_MUL_PER_MADD = {0: 10, 1: 28}  # XYZZ mixed add: 8M + 2S in Fq (G1) / in Fq2 = 8*3 + 2*2 (G2)
WIDE_PER_FQ_MUL = 128
# wide multiply-adds the kernels actually issue per mixed addition (csrc/ff.cuh, csrc/ec.cuh): G1 = 6 products x 128 +
# 2 squarings x 100 + 1 dual product x 192; G2 (lazy reduction over Fq2) = 6 products x (3 x 64 + 2 x 64) + 2 squarings
# x 2 x 128 + 1 dual product x (6 x 64 + 2 x 64). "int_pipe_frac" keeps the schoolbook count (comparable across
# rounds); "int_pipe_frac_issued" divides the issued work by the same peak.
WIDE_ISSUED_PER_MADD = {0: 6 * 128 + 2 * 100 + 192, 1: 6 * 320 + 2 * 256 + 512}


def uniform_scalars(n, rng):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] = rng.integers(0, R_TOP, size=n, dtype=np.uint64)
    return a


def keyless_mix_scalars(n, rng):
    a = uniform_scalars(n, rng)
    kind = rng.random(n)
    bits = kind < 0.84
    byt = (kind >= 0.84) & (kind < 0.96)
    a[bits, 1:] = 0
    a[bits, 0] = rng.integers(0, 2, size=int(bits.sum()), dtype=np.uint64)
    a[byt, 1:] = 0
    a[byt, 0] = rng.integers(0, 256, size=int(byt.sum()), dtype=np.uint64)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    ap.add_argument("--g1", default="16,18,20,21,22,24")
    ap.add_argument("--g2", default="16,18,20,21,22")
    ap.add_argument("--ntt", default="16,17,18,19,20,21,22,23,24")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    gen = bench.ensure_setupgen()
    gen.kzp_gen_consecutive_points.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p]
    gen.kzp_msm_closed_form.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p]
    kzp.lib()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    imad_peak, _ = kzp.imad_peak(8192, device=args.device)
    res = {"device": args.device, "imad_wide_peak_per_s": imad_peak, "hbm_peak_gbs": hbm_peak,
           "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)", "msm": [], "ntt": []}
    rng = np.random.default_rng(5)
    s0 = (0x1234567 << 128) + 0xDEADBEEF
    s0b = s0.to_bytes(32, "little")

    for group, sizes in ((0, args.g1), (1, args.g2)):
        psz = 64 if group == 0 else 128
        for lg in [int(x) for x in sizes.split(",") if x]:
            n = 1 << lg
            t0 = time.time()
            bases = ctypes.create_string_buffer(n * psz)
            assert gen.kzp_gen_consecutive_points(group, n, s0b, bases) == 0
            t_gen = time.time() - t0
            t0 = time.time()
            m = kzp.Msm(group, bases, device=args.device)
            t_tbl = time.time() - t0
            del bases
            for mix, make in (("uniform", uniform_scalars), ("keyless-mix", keyless_mix_scalars)):
                sc = make(n, rng)
                scb = sc.tobytes()
                got = m.run(scb)
                want = ctypes.create_string_buffer(psz)
                assert gen.kzp_msm_closed_form(group, n, s0b, sc.ctypes.data, want) == 0
                ok = got == want.raw
                ms, entries = m.bench(scb, args.iters)
                wide = entries * FQ_MUL_PER_MADD[group] * WIDE_PER_FQ_MUL
                row = {"group": "G1" if group == 0 else "G2", "log_n": lg, "scalars": mix, "ms": ms,
                       "pairs_per_s": n / (ms * 1e-3), "entries": entries,
                       "fq_mul_per_s": entries * FQ_MUL_PER_MADD[group] / (ms * 1e-3),
                       "int_pipe_frac": wide / (ms * 1e-3) / imad_peak,
                       "int_pipe_frac_issued": entries * WIDE_ISSUED_PER_MADD[group] / (ms * 1e-3) / imad_peak, "matches_closed_form": ok,
                       "gen_s": round(t_gen, 2), "table_s": round(t_tbl, 2)}
                res["msm"].append(row)
                print(json.dumps(row), flush=True)
                assert ok, "MSM result differs from the closed form"
            m.close()

    for lg in [int(x) for x in args.ntt.split(",") if x]:
        n = 1 << lg
        ms = kzp.fr_ntt_bench(lg, max(args.iters, 10), device=args.device)
        alg_bytes = 2 * 2 * n * 32
        muls = 2 * (n // 2) * lg + n  # butterflies of both transforms + the fused scale/shift
        row = {"log_n": lg, "chain": "ifft + coset shift + fft", "ms": ms, "algorithmic_gbs": alg_bytes / (ms * 1e-3) / 1e9,
               "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "fr_mul_per_s": muls / (ms * 1e-3),
               "int_pipe_frac": muls * WIDE_PER_FQ_MUL / (ms * 1e-3) / imad_peak}
        res["ntt"].append(row)
        print(json.dumps(row), flush=True)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
