import os, sys, statistics, time
sys.path.insert(0, os.getcwd())
import bench, keyless_zk_proofs_b200 as kzp
zkey, wtns, info = bench.ensure_inputs("keyless")
p = kzp.FullProver(zkey)
for _ in range(4): p.prove(wtns)
e=[]; g=[]
for _ in range(30):
    t=time.perf_counter(); p.prove(wtns); e.append(1e3*(time.perf_counter()-t)); g.append(p.timings()["gpu_ms"])
r=[]
for _ in range(30):
    t=time.perf_counter(); p.prove_resident(); r.append(1e3*(time.perf_counter()-t))
print("e2e p50 %.2f  resident p50 %.2f  gpu %.2f" % (statistics.median(e), statistics.median(r), statistics.median(g)), flush=True)
p.close()
