#!/usr/bin/env python
"""bench.py — keyless Groth16 prove on B200 (contract in the task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload keyless|small]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a keyless-SHAPED synthetic circuit — nVars 1,343,588, nPublic 1, 1,376,867
constraint rows (+2 public rows) -> domain 2^21 — with a bit/byte-heavy witness, built from a known trapdoor by
tools/setupgen.c (the real keyless zkey cannot be downloaded offline). One "step" = one Groth16 proof per GPU.

  value    proofs/s over all GPUs with the witness already resident in HBM (all kernels + result D2H + host assembly)
  e2e      the same through the reference-facing call FullProver.prove(wtns_path): file mapping, pinned staging,
           H2D of the witness, kernels, D2H, proof JSON — what prover-service would see
  N > 1    replicas: independent proofs one per GPU, no data-path collective (scaling "weak"); additionally rank 0
           times ONE proof sharded over all N GPUs inside the library call (kzp_prover_new_group: MSM base ranges
           split, every coset-NTT chain spread over the GPUs with peer-store transposes over NVLink, partial sums on
           the host; no Python and no collective library in the loop) while the other ranks idle:
           sharded_latency_ms_p50 (witness resident) and sharded_e2e_ms_p50.

Only the cpu_baseline leg and --impl reference touch oracle/ (they time the reference's own CPU prover,
oracle/_ref, or the C port when that library is absent).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_constraints, n_vars, seed)  — SURVEY.md §8(d) configs 2 and 1
    "keyless": (1376867, 1343588, 2),
    "small": (60000, 58000, 1),
    # the keyless shape with a pessimistic witness: ~30 % full-width field elements instead of ~4 % (bounds the
    # distribution assumption of the headline workload; the witness-side MSMs and the packed upload both grow)
    "keyless-wide": (1376867, 1343588, 2),
}
# gate mix per workload (tools/setupgen.c kzp_setupgen_set_mix) and what the resulting witness looks like
MIXES = {"keyless-wide": ((35, 55, 70), "~56% bits / ~14% bytes / ~30% full-width")}
DEFAULT_MIX = ((50, 80, 95), "~84% bits / ~12% bytes / ~4% full-width")
# XYZZ madd-2008-s in G1 (curve.cpp:203-249) is 8M + 2S. As issued by csrc/ec.cuh + csrc/ff.cuh: 6 products of 128 wide
# (32x32+64) multiply-adds, 2 dedicated squarings of 100 (28 off-diagonal + 8 diagonal + 64 reduction) and one dual
# product (Q - x3) R + (-y1) PPP with a single reduction, 192. The 8 low IMADs per reduction (m) are not counted.
IMAD_PER_FQ_MUL = 128
IMAD_PER_MIXED_ADD = 6 * 128 + 2 * 100 + 192


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries print there too (NCCL's version banner goes to stdout even with
    NCCL_DEBUG_FILE set), so file descriptor 1 is pointed at stderr for the whole run and the result line is written to
    a private duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def data_dir():
    for d in ("/dev/shm", "/tmp"):
        if os.path.isdir(d) and os.access(d, os.W_OK):
            p = os.path.join(d, "kzp_bench")
            os.makedirs(p, exist_ok=True)
            return p
    raise RuntimeError("no writable scratch directory")


def ensure_setupgen():
    so = os.path.join(ROOT, "tools", "libkzp_setupgen.so")
    src = os.path.join(ROOT, "tools", "setupgen.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-fopenmp", "-shared", src, "-o", so, "-lm"])
    lib = ctypes.CDLL(so)
    lib.kzp_setupgen.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p,
                                 ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p]
    return lib


def ensure_inputs(workload):
    """Synthesises zkey + wtns for the workload once per box (cached in /dev/shm)."""
    nc, nv, seed = WORKLOADS[workload]
    d = data_dir()
    zkey, wtns, meta = (os.path.join(d, "%s.%s" % (workload, e)) for e in ("zkey", "wtns", "json"))
    if os.path.exists(meta) and os.path.exists(zkey) and os.path.exists(wtns):
        info = json.load(open(meta))
        if info.get("n_vars") == nv and os.path.getsize(zkey) == info.get("zkey_bytes"):
            return zkey, wtns, info
    t0 = time.time()
    lib = ensure_setupgen()
    arr = (ctypes.c_uint64 * 8)()
    assert lib.kzp_setupgen_set_mix(*MIXES.get(workload, DEFAULT_MIX)[0]) == 0
    rc = lib.kzp_setupgen(nc, nv, seed, (zkey + ".tmp").encode(), (wtns + ".tmp").encode(), arr, None)
    lib.kzp_setupgen_set_mix(*DEFAULT_MIX[0])
    if rc != 0:
        raise RuntimeError("setup generation failed")
    os.replace(zkey + ".tmp", zkey)
    os.replace(wtns + ".tmp", wtns)
    info = {"n_vars": arr[0], "n_public": arr[1], "domain": arr[2], "n_coefs": arr[3], "n_constraints": arr[4],
            "public_input": arr[5], "zkey_bytes": os.path.getsize(zkey), "wtns_bytes": os.path.getsize(wtns),
            "seed": seed, "gen_seconds": round(time.time() - t0, 1)}
    json.dump(info, open(meta, "w"))
    log("[bench] generated %s inputs in %.1fs" % (workload, time.time() - t0))
    return zkey, wtns, info


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


REF_FLAVOUR = ""


def cpu_reference_prover(zkey):
    """The reference's own CPU prover (oracle/_ref) or, when that library did not travel, the C port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refutil
    ref, flavour = refutil.load_ref_asm(), "x86-64 assembly field path (the reference's fr.asm / fq.asm, MULX/ADCX/ADOX, assembled with GNU as)"
    if ref is None or os.environ.get("KZP_REF_GENERIC"):
        ref, flavour = refutil.load_ref(), "portable GMP-mpn field path (the asm build is absent or this CPU lacks ADX/BMI2)"
    global REF_FLAVOUR
    if ref is not None:
        REF_FLAVOUR = flavour + " + OpenMP stand-in for oneTBB"
        h = ref.lib.kzp_ref_prover_new(zkey.encode(), 1, None)
        buf = ctypes.create_string_buffer(8192)
        ms = ctypes.c_int()
        cores = ref.lib.kzp_ref_num_threads()

        def prove(wtns):
            t0 = time.perf_counter()
            rc = ref.lib.kzp_ref_prover_prove(h, wtns.encode(), None, None, 1, buf, 8192, ctypes.byref(ms))
            assert rc == 0
            return time.perf_counter() - t0
        return "reference", cores, prove
    port = refutil.load_port()
    REF_FLAVOUR = "C restatement oracle/kzp_port.c (oracle/_ref did not travel)"
    cores = port.lib.kzp_port_num_threads()
    r, s = (12345).to_bytes(32, "little"), (67890).to_bytes(32, "little")

    def prove(wtns):
        t0 = time.perf_counter()
        port.prove(zkey, wtns, r, s)
        return time.perf_counter() - t0
    return "port", cores, prove


def run_reference_arm(args, zkey, wtns, info, rank):
    if rank != 0:
        return
    kind, cores, prove = cpu_reference_prover(zkey)
    budget = float(os.environ.get("KZP_REF_BUDGET_S", "420"))
    t_start = time.perf_counter()
    times, warm = [], 0
    for _ in range(max(1, args.warmup)):
        prove(wtns)
        warm += 1
        if time.perf_counter() - t_start > budget / 3:
            break
    for _ in range(args.steps):
        times.append(prove(wtns))
        if time.perf_counter() - t_start > budget and len(times) >= 1:
            break
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * total / len(times),
        "latency_ms_p50": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (254-bit modular integers)", "data": "synthetic",
        "config": workload_config(args, info),
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": kind,
                         "sample": "%d full proofs of the same zkey/witness, one at a time (prover-service serialises "
                                   "proofs behind one mutex); %s" % (len(times), REF_FLAVOUR)},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


METRIC = "keyless Groth16 prove throughput (proofs/s); p50 latency (ms) in latency_ms_p50"


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(q * (len(xs) - 1))))]


def measured_traffic():
    """DRAM bytes per launch from the committed ncu capture (scripts/ncu_traffic.sh -> profiles/r02_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return None


def workload_config(args, info):
    return {"workload": "%s-shaped synthetic circuit (trapdoor setup, tools/setupgen.c seed %d): nVars %d, nPublic %d, "
                        "%d constraints, domain 2^%d, nCoefs %d; witness %s"
                        % (args.workload, info["seed"], info["n_vars"], info["n_public"], info["n_constraints"],
                           int(info["domain"]).bit_length() - 1, info["n_coefs"], MIXES.get(args.workload, DEFAULT_MIX)[1]),
            "proofs_per_step_per_gpu": 1, "parallelism": "replicas x%d (one proof per GPU)" % args.gpus,
            "l2": "working set (multi-GB resident key tables, 4 x %d MiB vectors) exceeds the 126 MB L2; no flush needed"
                  % (info["domain"] * 32 >> 20)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="keyless", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    if rank == 0:
        # torchrun pins OMP_NUM_THREADS=1; the input generator and the CPU reference arm are OpenMP programs that
        # should use the box's cores (libgomp reads this when it is first loaded, i.e. before `import torch`)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

    if args.impl == "reference":
        # rank 0 alone runs and prints; the others exit without work (no process group needed)
        if rank == 0:
            zkey, wtns, info = ensure_inputs(args.workload)
            run_reference_arm(args, zkey, wtns, info, rank)
        return 0

    if rank == 0:
        zkey, wtns, info = ensure_inputs(args.workload)  # before torch (and its libgomp) is loaded; other ranks wait at init

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the prover has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idle_group = dist.new_group(backend="gloo")  # a barrier that leaves the GPUs alone (NCCL's spins a kernel)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    if rank != 0:
        zkey, wtns, info = ensure_inputs(args.workload)

    import keyless_zk_proofs_b200 as kzp
    kzp.lib()  # raises if the CUDA library is missing: there is no other implementation
    t0 = time.time()
    prover = kzp.FullProver(zkey, device=local_rank)
    load_s = time.time() - t0
    witness = open(wtns, "rb").read()
    # wtns: 12-byte header, section 1 (12 + 40 bytes), section 2 header 12 bytes, then the values
    values = witness[len(witness) - prover.n_vars * 32:]
    imad_peak, _ = kzp.imad_peak(8192, device=local_rank)

    # ---- device-resident arm: witness already in HBM ------------------------------------------------------
    prover.upload_witness(values)
    for _ in range(args.warmup):
        prover.prove_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.6)  # let nvidia-smi start sampling before the timed region (outside the timing)
    barrier()
    step_ms, gpu_ms, acc_ms, ntt_ms, stage = [], [], [], [], []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        prover.prove_resident()
        step_ms.append(1e3 * (time.perf_counter() - t1))
        tm = prover.timings()
        gpu_ms.append(tm["gpu_ms"])
        ntt_ms.append(tm["ntt_ms"])
        stage.append(tm)
        acc_ms.append(prover.msm_profile(4))
    barrier()
    elapsed = time.perf_counter() - t_begin
    clocks = sampler.stop() if rank == 0 else None
    launches_per_proof = int(stage[-1]["kernel_launches"])

    # ---- end-to-end arm: the reference-facing call with host buffers --------------------------------------
    for _ in range(min(2, args.warmup)):
        prover.prove(wtns)
    barrier()
    e2e_ms = []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        last_proof, _ = prover.prove(wtns)
        e2e_ms.append(1e3 * (time.perf_counter() - t1))
    barrier()
    e2e_elapsed = time.perf_counter() - t_begin
    e2e_stage = prover.timings()
    # a proof out of the timed loop is checked under the circuit's verifying key (host pairing check, no oracle)
    proof_ok = kzp.host_verify(zkey, last_proof, [info["public_input"]])
    if not proof_ok:
        raise SystemExit("bench.py: a proof produced in the timed region does not verify")

    # ---- sharded single proof (N > 1): one library call drives all N GPUs; the other ranks keep theirs idle ----
    sharded = None
    if distributed:
        barrier()
        if rank == 0:
            with kzp.FullProver(zkey, devices=list(range(world))) as sp:
                shards, fused, dist_ntt = sp.group_info()
                r32, s32 = (12345).to_bytes(32, "little"), (67890).to_bytes(32, "little")
                want, _ = prover.prove(wtns, r32, s32)
                got, _ = sp.prove(wtns, r32, s32)
                sp.upload_witness(values)
                for _ in range(args.warmup):
                    sp.prove_resident()
                lat, lat_e2e, shard_tm = [], [], []
                for _ in range(max(args.steps, 10)):
                    t1 = time.perf_counter()
                    sp.prove_resident()
                    lat.append(1e3 * (time.perf_counter() - t1))
                    shard_tm.append([sp.shard_timings(k) for k in range(shards)])
                for _ in range(max(args.steps, 10)):
                    t1 = time.perf_counter()
                    sp.prove(wtns)
                    lat_e2e.append(1e3 * (time.perf_counter() - t1))
                keys = ("spmv_ms", "ntt_ms", "msm_h_ms", "msm_wg1_ms", "msm_wg2_ms", "gpu_ms")
                sharded = {"latency_ms_p50": statistics.median(lat), "latency_ms_p95": pct(lat, 0.95),
                           "e2e_ms_p50": statistics.median(lat_e2e), "e2e_ms_p95": pct(lat_e2e, 0.95),
                           "proof_equals_single_gpu": got == want, "shards": shards, "distributed_ntt": dist_ntt,
                           "fused_peer_store_exchange": fused,
                           "slowest_shard_stage_ms_median": {k: max(statistics.median(st[i][k] for st in shard_tm)
                                                                    for i in range(shards)) for k in keys}}
        dist.barrier(group=idle_group)

    # ---- reduce over ranks: max time ----------------------------------------------------------------------
    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    elapsed = max_over_ranks(elapsed)
    e2e_elapsed = max_over_ranks(e2e_elapsed)
    p50 = max_over_ranks(statistics.median(step_ms))
    e2e_p50 = max_over_ranks(statistics.median(e2e_ms))

    if rank == 0:
        n = world if distributed else 1
        value = n * args.steps / elapsed
        acc_t = statistics.median(a for a, _ in acc_ms)
        entries = acc_ms[-1][1]
        imads = entries * IMAD_PER_MIXED_ADD
        achieved = imads / (acc_t * 1e-3) / 1e12
        ntt_t = statistics.median(ntt_ms)
        ntt_bytes = 6 * 2 * info["domain"] * 32
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tr = measured_traffic() if args.workload == "keyless" else None
        acc_traffic = tr["h_accumulate"]["bytes_per_launch"] if tr else None
        ntt_traffic = tr["ntt_chain"]["bytes_per_proof"] if tr else None
        log_n = int(info["domain"]).bit_length() - 1
        ntt_muls = 3 * (2 * (info["domain"] // 2) * log_n + info["domain"]) + 2 * info["domain"]  # butterflies + coset scale + pointwise
        line = {
            "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "latency_ms_p50": p50,
            "latency_ms_p95": pct(step_ms, 0.95),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (254-bit modular integers, 8 x 32-bit Montgomery)", "data": "synthetic",
            "config": workload_config(args, info),
            "clocks": clocks,
            "e2e": {"value": n * args.steps / e2e_elapsed, "unit": "proofs/s", "latency_ms_p50": e2e_p50,
                    "latency_ms_p95": pct(e2e_ms, 0.95), "proof_verifies": proof_ok,
                    "h2d_bytes_per_step": int(e2e_stage["h2d_mbytes"] * 1e6), "witness_bytes": prover.n_vars * 32, "d2h_bytes_per_step": kzp.PARTIALS_BYTES,
                    "api": "FullProver.prove(wtns_path) via kzp_prover_prove (C ABI): witness file -> classify + pack into pinned "
                           "memory -> H2D of the packed slices -> expansion kernel -> proof kernels -> D2H -> proof JSON", "h2d_ms": e2e_stage["h2d_ms"]},
            "gpu_launches": launches_per_proof * args.steps * n,
            "gpu_launches_per_proof": launches_per_proof,
            "stage_ms_median": {k: statistics.median(s[k] for s in stage) for k in
                                ("spmv_ms", "ntt_ms", "msm_h_ms", "msm_wsort_ms", "msm_wg1_ms", "msm_wg2_ms",
                                 "gpu_ms", "assemble_host_ms")},
            "roofline": {
                "kernel": "k_msm_accumulate<G1> of the H MSM (bucket accumulation, XYZZ mixed adds)",
                "bound": "integer-pipe (IMAD.WIDE; BASELINE.json: 'MSM vs integer-multiply peak') — neither hbm nor tensor: "
                         "ncu shows sm__pipe_fmaheavy_cycles_active 89% and dram throughput 11% for this kernel",
                "achieved": achieved, "peak": imad_peak / 1e12, "unit": "T wide-multiply-add/s", "frac": achieved / (imad_peak / 1e12),
                "peak_source": "kzp_imad_peak: carry-chained mad.lo.cc/madc.hi.cc (IMAD.WIDE.U32[.X]) on all SMs, measured in "
                               "this run; IMAD.WIDE issues at half the 32-bit IMAD rate on sm_100a (profiles/r01_ubench_int_fp64_pipes.txt)",
                "algorithmic_work": "%d sorted (point,bucket) entries x %d wide multiply-adds per mixed addition (6 products x 128 + 2 squarings x 100 + 1 dual product x 192)" % (entries, IMAD_PER_MIXED_ADD),
                "launch_ms": acc_t,
                "launch_note": "timed inside the proof with CUDA events on its stream, while the A/B1/C witness batch and B2 share "
                               "the SMs (they are scheduled into the H digit sort and MSM on purpose); alone (ncu, serialised) the same "
                               "launch is faster: profiles/r02_traffic.json ms_under_ncu",
                "traffic": acc_traffic,
                "traffic_unit": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu "
                                "capture profiles/r02_traffic.json (scripts/ncu_traffic.sh; null when that file is absent or the "
                                "workload differs); algorithmic bytes = entries x (64 B point + 4 B entry) = %.2e: every random "
                                "64-byte point gather costs a 128-byte DRAM->L2 fill (DESIGN.md section 4)" % (entries * 68.0),
                "hbm_view": None if not acc_traffic else {"achieved": acc_traffic / (acc_t * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                          "frac": acc_traffic / (acc_t * 1e-3) / 1e9 / hbm_peak}},
            "roofline_ntt": {
                "kernel": "NTT level kernels, 3 x (iNTT + coset + NTT) + pointwise, per proof",
                "bound": "hbm", "achieved": ntt_bytes / (ntt_t * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": ntt_bytes / (ntt_t * 1e-3) / 1e9 / hbm_peak,
                "algorithmic_bytes": ntt_bytes, "launch_ms": ntt_t, "traffic": ntt_traffic,
                "traffic_unit": "bytes per proof, summed over the NTT launches of one proof (profiles/r02_traffic.json): "
                                "a 2^21 transform is 3 radix-128 passes (2.5 with the fused middle), each reading and writing a, b, c once",
                "int_view": {"achieved": ntt_muls * IMAD_PER_FQ_MUL / (ntt_t * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                             "unit": "T wide-multiply-add/s", "frac": ntt_muls * IMAD_PER_FQ_MUL / (ntt_t * 1e-3) / imad_peak,
                             "algorithmic_work": "%d Fr products (n/2 log n butterflies per transform x 6, coset scale, pointwise) x 128" % ntt_muls},
                "note": "algorithmic bytes = 6 transforms x 2 x n x 32 B (SURVEY.md §8(d)); the binding roof for a 254-bit NTT "
                        "on B200 is the integer pipe (int_view); launch_ms is the in-proof time, with the B2 MSM sharing the SMs"},
            "load_seconds": load_s,
        }
        if sharded is not None:
            line["sharded_latency_ms_p50"] = sharded["latency_ms_p50"]
            line["sharded_e2e_ms_p50"] = sharded["e2e_ms_p50"]
            line["sharded"] = dict(sharded, what="ONE proof over %d GPUs inside kzp_prover_prove (prover group in libkzp_b200.so, "
                                                 "driven by rank 0 alone while the other ranks idle): MSM base ranges split, coset-NTT "
                                                 "chains spread over the GPUs with peer-store transposes over NVLink, 768-byte partials "
                                                 "summed on the host" % n)
        if not args.no_cpu_baseline and n == 1:
            kind, cores, prove = cpu_reference_prover(zkey)
            prove(wtns)  # first call pages the key in
            t = prove(wtns)
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": "proofs/s", "latency_ms": 1e3 * t, "cores": cores, "kind": kind,
                                    "sample": "1 full proof of the same zkey/witness after 1 warm-up; " + REF_FLAVOUR}
        emit(line)
    prover.close()
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
