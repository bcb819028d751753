"""Times the reference-facing call (FullProver.prove(wtns_path)) N times: python scripts/e2e_once.py [workload] [n]"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import keyless_zk_proofs_b200 as kzp
workload = sys.argv[1] if len(sys.argv) > 1 else "keyless"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
zkey, wtns, info = bench.ensure_inputs(workload)
p = kzp.FullProver(zkey)
ts = []
for i in range(n + 3):
    t0 = time.perf_counter()
    p.prove(wtns)
    ts.append(1e3 * (time.perf_counter() - t0))
tm = p.timings()
print("e2e p50 %.3f ms  min %.3f  h2d %.3f gpu %.3f assemble %.3f" % (statistics.median(ts[3:]), min(ts[3:]), tm["h2d_ms"], tm["gpu_ms"], tm["assemble_host_ms"]), flush=True)
p.close()
