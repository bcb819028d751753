"""Runs N proofs of the bench workload (for ncu captures): python scripts/prove_once.py [workload] [n_proofs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import keyless_zk_proofs_b200 as kzp
workload = sys.argv[1] if len(sys.argv) > 1 else "keyless"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
zkey, wtns, info = bench.ensure_inputs(workload)
p = kzp.FullProver(zkey)
p.upload_witness_file(wtns)
for i in range(n):
    p.run_gpu()
    p.assemble([p.partials()])
    print({k: round(v, 3) for k, v in p.timings().items()}, flush=True)
p.close()
