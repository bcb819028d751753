"""world_size-2 gloo test of the N>1 path's host logic (no GPU): every rank contributes its shard's partial MSM
results, one all-gather moves the 768-byte partials, rank 0 assembles with the product's host code and must
reproduce the reference's proof bytes. The partials themselves come from the oracle here (on the GPU box the
same plumbing is fed by kzp_prover_run_gpu; tests/test_gpu_parity.py checks that side)."""
import json
import os
import socket
import sys

import pytest

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bn254 as oracle
    import keyless_zk_proofs_b200 as kzp
    from test_host_library import _partials_from_oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = os.path.join(GOLDEN, "syn256")
    exp = json.load(open(os.path.join(d, "expected.json")))
    zk = oracle.read_zkey(os.path.join(d, "syn256.zkey"))
    w = oracle.read_wtns(os.path.join(d, "syn256.wtns"))
    h = [oracle.from_le(bytes.fromhex(exp["h"])[i * 32:(i + 1) * 32]) for i in range(zk.domain_size)]
    rng = lambda n: (rank * n // world, (rank + 1) * n // world)
    mine = _partials_from_oracle(oracle, zk, w, (rng(zk.n_vars), rng(zk.n_vars - zk.n_public - 1), rng(zk.domain_size), h))
    assert len(mine) == kzp.PARTIALS_BYTES
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    if rank == 0:
        parts = [g.numpy().tobytes() for g in gathered]
        js, msm = kzp.host_assemble(os.path.join(d, "syn256.zkey"), parts, bytes.fromhex(exp["r"]), bytes.fromhex(exp["s"]))
        json.dump({"ok": js == exp["proof"] and msm.hex() == exp["msm"]}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_assemble(kzp, workdir):
    import torch.multiprocessing as mp

    out = os.path.join(workdir, "dist_result.json")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert json.load(open(out))["ok"]
