#!/usr/bin/env python
"""One keyless-shaped proof sharded over N GPUs INSIDE the prove call (kzp_prover_new_group): latency of
FullProver.prove_resident (witness in HBM) and FullProver.prove(wtns_path) (file -> pack -> H2D on every GPU -> kernels),
per-shard stage timings, proof checked against the single-GPU prover's bytes for the same (r, s).

    python tools/group_bench.py --devices 0,1,2,3 [--steps 20] [--warmup 3] [--scatter 0|1] [--out file.json]
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(q * (len(xs) - 1))))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scatter", type=int, default=None)
    ap.add_argument("--workload", default="keyless")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import bench

    zkey, wtns, info = bench.ensure_inputs(args.workload)
    import keyless_zk_proofs_b200 as kzp

    devices = [int(x) for x in args.devices.split(",")]
    if args.scatter is not None:
        os.environ["KZP_GROUP_SCATTER"] = str(args.scatter)
    r, s = (12345).to_bytes(32, "little"), (67890).to_bytes(32, "little")
    expected = None
    if not args.no_check:
        with kzp.FullProver(zkey, device=devices[0]) as single:
            expected, _ = single.prove(wtns, r, s)
    t0 = time.time()
    p = kzp.FullProver(zkey, devices=devices)
    load_s = time.time() - t0
    shards, fused, dist = p.group_info()
    js, _ = p.prove(wtns, r, s)
    ok = expected is None or js == expected
    values = open(wtns, "rb").read()[-p.n_vars * 32:]
    p.upload_witness(values)
    for _ in range(args.warmup):
        p.prove_resident()
    res, shard_tm = [], []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        p.prove_resident()
        res.append(1e3 * (time.perf_counter() - t1))
        shard_tm.append([p.shard_timings(k) for k in range(shards)])
    for _ in range(min(2, args.warmup)):
        p.prove(wtns)
    e2e = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        p.prove(wtns)
        e2e.append(1e3 * (time.perf_counter() - t1))
    tm = p.timings()
    keys = ("h2d_ms", "spmv_ms", "ntt_ms", "msm_h_ms", "msm_wsort_ms", "msm_wg1_ms", "msm_wg2_ms", "gpu_ms")
    line = {
        "what": "one proof sharded over %d GPUs inside kzp_prover_prove (prover group)" % shards,
        "devices": devices, "fused_peer_store_exchange": fused, "distributed_ntt": dist, "proof_equals_single_gpu": ok,
        "resident_ms_p50": statistics.median(res), "resident_ms_p95": pct(res, 0.95), "resident_ms_min": min(res),
        "e2e_ms_p50": statistics.median(e2e), "e2e_ms_p95": pct(e2e, 0.95),
        "steps": args.steps, "load_seconds": load_s, "h2d_mbytes_all_gpus": tm["h2d_mbytes"],
        "kernel_launches_all_gpus": int(tm["kernel_launches"]),
        "shard_stage_ms_median": [{k: round(statistics.median(st[i][k] for st in shard_tm), 3) for k in keys}
                                  for i in range(shards)],
        "workload": bench.workload_config(argparse.Namespace(workload=args.workload, gpus=shards), info)["workload"],
    }
    p.close()
    out = json.dumps(line)
    print(out)
    if args.out:
        open(args.out, "w").write(out + "\n")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
