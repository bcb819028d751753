"""keyless-zk-proofs_b200 — B200-native Groth16/BN254 prover behind rust-rapidsnark's FullProver boundary.

This package is only the Python-side mirror of the reference's binding (rust-rapidsnark/src/lib.rs:41-106):
a ctypes view of ``libkzp_b200.so`` (include/kzp_b200.h). All computation happens in the CUDA library; there
is no CPU fallback and nothing here imports ``oracle/``. Import through ``keyless_zk_proofs_b200`` (the
underscore alias at the repo root) or ``importlib.import_module("keyless-zk-proofs_b200")``.
"""
from __future__ import annotations

import ctypes
import enum
import os
from typing import Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkzp_b200.so")
CLI_PATH = os.path.join(_HERE, "kzp_prove")  # csrc/cli_main.cpp: zkey + wtns -> proof.json + public.json

PARTIALS_BYTES = 768


class ProverInitError(Exception):
    """Mirror of rust-rapidsnark's ProverInitError (src/lib.rs:17-22)."""


class ZKeyFileLoadError(ProverInitError):
    pass


class UnsupportedZKeyCurve(ProverInitError):
    pass


class ProverError(Exception):
    """Mirror of rust-rapidsnark's ProverError (src/lib.rs:24-33)."""


class ProverNotReady(ProverError):
    pass


class InvalidInput(ProverError):
    pass


class WitnessGenerationInvalidCurve(ProverError):
    pass


class KzpError(RuntimeError):
    """Component-level failure (status code + kzp_last_error())."""


class Field(enum.IntEnum):
    FR = 0
    FQ = 1
    FQ2 = 2


class FieldOp(enum.IntEnum):
    MUL = 0
    ADD = 1
    SUB = 2
    NEG = 3
    TO_MONTGOMERY = 4
    FROM_MONTGOMERY = 5
    SQUARE = 6
    INVERSE = 7


_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    from importlib import util

    spec = util.spec_from_file_location("_kzp_build", os.path.join(_HERE, "build.py"))
    mod = util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib() -> ctypes.CDLL:
    """Loads the CUDA library. Raises if it has not been built — there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libkzp_b200.so is missing: run `python keyless-zk-proofs_b200/build.py` (needs nvcc). "
            "This package has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    u8p, vp, i32p = c.c_char_p, c.c_void_p, c.POINTER(c.c_int)
    sig = {
        "kzp_last_error": (c.c_char_p, []),
        "kzp_version": (c.c_char_p, []),
        "kzp_device_count": (c.c_int, []),
        "kzp_free": (None, [vp]),
        "kzp_prover_new": (vp, [c.c_char_p, c.c_int, i32p]),
        "kzp_prover_new_sharded": (vp, [c.c_char_p, c.c_int, c.c_int, c.c_int, i32p]),
        "kzp_prover_new_group": (vp, [c.c_char_p, c.POINTER(c.c_int), c.c_int, i32p]),
        "kzp_prover_group_info": (c.c_int, [vp, i32p, i32p, i32p]),
        "kzp_prover_free": (None, [vp]),
        "kzp_prover_prove": (c.c_int, [vp, c.c_char_p, u8p, u8p, c.POINTER(vp), i32p, i32p]),
        "kzp_prover_prove_mem": (c.c_int, [vp, u8p, c.c_uint64, u8p, u8p, c.POINTER(vp), i32p, i32p]),
        "kzp_prover_prove_resident": (c.c_int, [vp, u8p, u8p, c.POINTER(vp), i32p, i32p]),
        "kzp_prover_upload_witness": (c.c_int, [vp, u8p, c.c_uint64]),
        "kzp_prover_upload_witness_file": (c.c_int, [vp, c.c_char_p]),
        "kzp_prover_run_gpu": (c.c_int, [vp]),
        "kzp_prover_get_partials": (c.c_int, [vp, u8p]),
        "kzp_prover_assemble": (c.c_int, [vp, u8p, c.c_int, u8p, u8p, c.POINTER(vp)]),
        "kzp_prover_info": (c.c_int, [vp, c.POINTER(c.c_uint32), c.POINTER(c.c_uint32), c.POINTER(c.c_uint32),
                                      c.POINTER(c.c_uint64), i32p]),
        "kzp_prover_timings": (c.c_int, [vp, c.POINTER(c.c_float), c.c_int]),
        "kzp_prover_group_shard_timings": (c.c_int, [vp, c.c_int, c.POINTER(c.c_float), c.c_int]),
        "kzp_prover_msm_profile": (c.c_int, [vp, c.c_int, c.POINTER(c.c_float), c.POINTER(c.c_uint64)]),
        "kzp_prover_get_h": (c.c_int, [vp, u8p, c.c_uint64]),
        "kzp_prover_keep_ab": (c.c_int, [vp, c.c_int]),
        "kzp_prover_get_ab": (c.c_int, [vp, u8p, c.c_uint64]),
        "kzp_prover_get_msm_results": (c.c_int, [vp, u8p]),
        "kzp_pool_new": (vp, [c.c_char_p, c.POINTER(c.c_int), c.c_int, i32p]),
        "kzp_pool_free": (None, [vp]),
        "kzp_pool_size": (c.c_int, [vp]),
        "kzp_pool_device": (c.c_int, [vp, c.c_int]),
        "kzp_pool_prove": (c.c_int, [vp, c.c_char_p, u8p, u8p, c.POINTER(vp), i32p, i32p, i32p]),
        "kzp_pool_prove_mem": (c.c_int, [vp, u8p, c.c_uint64, u8p, u8p, c.POINTER(vp), i32p, i32p, i32p]),
        "kzp_pool_set_verify": (c.c_int, [vp, c.c_int]),
        "kzp_pool_stats": (c.c_int, [vp, c.POINTER(c.c_uint64), c.c_int, c.POINTER(c.c_uint64)]),
        "kzp_pool_sched_selftest": (c.c_int, [c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_uint64), i32p,
                                              c.POINTER(c.c_uint64)]),
        "kzp_fr_ntt": (c.c_int, [u8p, c.c_uint64, c.c_int, c.c_int]),
        "kzp_fr_coset_chain": (c.c_int, [u8p, c.c_uint64, c.c_int]),
        "kzp_fr_ntt_bench": (c.c_int, [c.c_uint32, c.c_int, c.c_int, c.POINTER(c.c_float)]),
        "kzp_msm_new": (vp, [c.c_int, u8p, c.c_uint64, c.c_int]),
        "kzp_msm_new_ex": (vp, [c.c_int, u8p, c.c_uint64, c.c_int, c.c_int, c.c_int]),
        "kzp_prover_state": (c.c_int, [vp]),
        "kzp_prover_last_status": (c.c_int, [vp]),
        "kzp_pool_healthy": (c.c_int, [vp]),
        "kzp_msm_free": (None, [vp]),
        "kzp_msm_run": (c.c_int, [vp, u8p, u8p]),
        "kzp_msm_bench": (c.c_int, [vp, u8p, c.c_int, c.POINTER(c.c_float), c.POINTER(c.c_uint64)]),
        "kzp_field_op": (c.c_int, [c.c_int, c.c_int, u8p, u8p, u8p, c.c_uint64, c.c_int]),
        "kzp_point_op": (c.c_int, [c.c_int, c.c_int, u8p, u8p, u8p, c.c_uint64, c.c_int]),
        "kzp_imad_peak": (c.c_int, [c.c_int, c.c_int, c.POINTER(c.c_float), c.POINTER(c.c_uint64)]),
        "kzp_host_parse_zkey": (c.c_int, [c.c_char_p, c.POINTER(c.c_uint32), c.POINTER(c.c_uint32),
                                          c.POINTER(c.c_uint32), c.POINTER(c.c_uint64), i32p]),
        "kzp_host_assemble": (c.c_int, [c.c_char_p, u8p, c.c_int, u8p, u8p, c.POINTER(vp), u8p]),
        "kzp_host_pairing_check": (c.c_int, [u8p, u8p, c.c_int, i32p]),
        "kzp_host_verify": (c.c_int, [c.c_char_p, c.c_char_p, u8p, c.c_uint32, i32p]),
        "kzp_verify_last_error": (c.c_char_p, []),
        "kzp_host_pack_witness_slice": (c.c_int, [u8p, c.c_uint32, vp, c.c_uint64, c.POINTER(c.c_uint64),
                                                   c.POINTER(c.c_uint32)]),
        "kzp_host_fq_decimal": (c.c_int, [u8p, c.c_char_p, c.c_size_t]),
        "kzp_host_field_op": (c.c_int, [c.c_int, c.c_int, u8p, u8p, u8p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here means the library and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def last_error() -> str:
    return (lib().kzp_last_error() or b"").decode()


def device_count() -> int:
    return lib().kzp_device_count()


def _check(rc: int):
    if rc != 0:
        raise KzpError("kzp status %d: %s" % (rc, last_error()))


def _take_string(ptr: ctypes.c_void_p) -> str:
    s = ctypes.string_at(ptr).decode()
    lib().kzp_free(ptr)
    return s


TIMING_KEYS = ("h2d_ms", "spmv_ms", "ntt_ms", "msm_h_ms", "msm_wsort_ms", "msm_wg1_ms", "msm_wg2_ms", "h2d_mbytes",
               "gpu_ms", "assemble_host_ms", "total_host_ms", "kernel_launches")


class FullProver:
    """Python mirror of ``rust_rapidsnark::FullProver`` (rust-rapidsnark/src/lib.rs:41-106).

    ``FullProver(zkey_path)`` raises ZKeyFileLoadError / UnsupportedZKeyCurve exactly where the Rust
    ``FullProver::new`` returns those errors; ``prove(wtns_path)`` returns ``(proof_json, metrics)`` where
    ``metrics["prover_time"]`` is in milliseconds, and raises the ProverError variants of the Rust binding.
    """

    def __init__(self, zkey_path: str, device: int = -1, shard: Optional[Tuple[int, int]] = None,
                 devices: Optional[Sequence[int]] = None):
        """``devices=[0, 1, ...]``: one proof sharded over those GPUs inside every prove call (kzp_prover_new_group);
        ``shard=(rank, world)``: one shard of the one-process-per-GPU mode; neither: one GPU (or $KZP_SHARD_DEVICES)."""
        L = lib()
        st = ctypes.c_int(0)
        if devices is not None:
            arr = (ctypes.c_int * len(devices))(*devices)
            self._h = L.kzp_prover_new_group(os.fsencode(zkey_path), arr, len(devices), ctypes.byref(st))
        elif shard is None:
            self._h = L.kzp_prover_new(os.fsencode(zkey_path), device, ctypes.byref(st))
        else:
            self._h = L.kzp_prover_new_sharded(os.fsencode(zkey_path), device, shard[0], shard[1], ctypes.byref(st))
        self.state = st.value
        if self.state != 0:
            msg = last_error()
            h, self._h = self._h, None
            L.kzp_prover_free(h)
            if self.state == 1:
                raise ZKeyFileLoadError(msg)
            raise UnsupportedZKeyCurve(msg)
        nv, npub, dom = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        nc, dev = ctypes.c_uint64(), ctypes.c_int()
        _check(L.kzp_prover_info(self._h, ctypes.byref(nv), ctypes.byref(npub), ctypes.byref(dom), ctypes.byref(nc),
                                 ctypes.byref(dev)))
        self.n_vars, self.n_public, self.domain_size = nv.value, npub.value, dom.value
        self.n_coefs, self.device = nc.value, dev.value

    def group_info(self) -> Tuple[int, bool, bool]:
        """(number of shards, whether the slices travel as fused peer stores, whether every NTT chain is spread over
        all shards)"""
        n, f, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib().kzp_prover_group_info(self._h, ctypes.byref(n), ctypes.byref(f), ctypes.byref(d)))
        return n.value, bool(f.value), bool(d.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().kzp_prover_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def _raise_prover_error(err: int):
        msg = last_error()
        if err == 1:
            raise ProverNotReady(msg)
        if err == 3:
            raise WitnessGenerationInvalidCurve(msg)
        raise InvalidInput(msg)

    def prove(self, wtns_path: str, r: Optional[bytes] = None, s: Optional[bytes] = None):
        out, err, ms = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        rc = lib().kzp_prover_prove(self._h, os.fsencode(wtns_path), r, s, ctypes.byref(out), ctypes.byref(err),
                                    ctypes.byref(ms))
        if rc != 0:
            self._raise_prover_error(err.value)
        return _take_string(out), {"prover_time": ms.value}

    def prove_mem(self, witness: bytes, r: Optional[bytes] = None, s: Optional[bytes] = None):
        out, err, ms = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        rc = lib().kzp_prover_prove_mem(self._h, witness, len(witness) // 32, r, s, ctypes.byref(out),
                                        ctypes.byref(err), ctypes.byref(ms))
        if rc != 0:
            self._raise_prover_error(err.value)
        return _take_string(out), {"prover_time": ms.value}

    def prove_resident(self, r: Optional[bytes] = None, s: Optional[bytes] = None):
        """One proof on the witness already uploaded with upload_witness*(): GPU work + overlapped host assembly."""
        out, err, ms = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int()
        rc = lib().kzp_prover_prove_resident(self._h, r, s, ctypes.byref(out), ctypes.byref(err), ctypes.byref(ms))
        if rc != 0:
            self._raise_prover_error(err.value)
        return _take_string(out), {"prover_time": ms.value}

    # ---- split life cycle (bench, sharded mode)
    def upload_witness(self, witness: bytes):
        _check(lib().kzp_prover_upload_witness(self._h, witness, len(witness) // 32))

    def upload_witness_file(self, path: str):
        _check(lib().kzp_prover_upload_witness_file(self._h, os.fsencode(path)))

    def run_gpu(self):
        _check(lib().kzp_prover_run_gpu(self._h))

    def partials(self) -> bytes:
        buf = ctypes.create_string_buffer(PARTIALS_BYTES)
        _check(lib().kzp_prover_get_partials(self._h, buf))
        return buf.raw

    def assemble(self, partials: Sequence[bytes], r: Optional[bytes] = None, s: Optional[bytes] = None) -> str:
        blob = b"".join(partials)
        out = ctypes.c_void_p()
        _check(lib().kzp_prover_assemble(self._h, blob, len(partials), r, s, ctypes.byref(out)))
        return _take_string(out)

    # ---- parity artefacts / diagnostics
    def timings(self) -> dict:
        arr = (ctypes.c_float * 12)()
        n = lib().kzp_prover_timings(self._h, arr, 12)
        return {k: float(arr[i]) for i, k in enumerate(TIMING_KEYS[:n])}

    def shard_timings(self, shard: int) -> dict:
        arr = (ctypes.c_float * 12)()
        n = lib().kzp_prover_group_shard_timings(self._h, shard, arr, 12)
        return {k: float(arr[i]) for i, k in enumerate(TIMING_KEYS[:n])}

    def msm_profile(self, which: int):
        """(accumulate kernel ms, sorted entries) of MSM 0=A 1=B1 2=B2 3=C 4=H in the last proof."""
        ms, ent = ctypes.c_float(), ctypes.c_uint64()
        _check(lib().kzp_prover_msm_profile(self._h, which, ctypes.byref(ms), ctypes.byref(ent)))
        return ms.value, ent.value

    def h_coefficients(self) -> bytes:
        buf = ctypes.create_string_buffer(self.domain_size * 32)
        _check(lib().kzp_prover_get_h(self._h, buf, len(buf)))
        return buf.raw

    def keep_ab(self, on: bool = True):
        _check(lib().kzp_prover_keep_ab(self._h, 1 if on else 0))

    def ab(self) -> bytes:
        buf = ctypes.create_string_buffer(self.domain_size * 64)
        _check(lib().kzp_prover_get_ab(self._h, buf, len(buf)))
        return buf.raw

    def msm_results(self) -> bytes:
        buf = ctypes.create_string_buffer(384)
        _check(lib().kzp_prover_get_msm_results(self._h, buf))
        return buf.raw


class ProverPool:
    """GPU-per-request prover pool (include/kzp_b200.h §1b): the replacement for prover-service's single
    Arc<Mutex<Option<FullProver>>> (prover-service/src/prover_state.rs:21,38-47). ``prove`` may be called from any
    number of threads; each call blocks until a prover is free (ctypes releases the GIL during the call)."""

    def __init__(self, zkey_path: str, devices: Optional[Sequence[int]] = None):
        state = ctypes.c_int(-1)
        if devices is None:
            arr, n = None, 0
        else:
            arr, n = (ctypes.c_int * len(devices))(*devices), len(devices)
        self._h = lib().kzp_pool_new(os.fsencode(zkey_path), arr, n, ctypes.byref(state))
        if not self._h:
            raise MemoryError("kzp_pool_new returned NULL")
        if state.value != 0:
            why = last_error()
            lib().kzp_pool_free(self._h)
            self._h = None
            if state.value == 2:
                raise UnsupportedZKeyCurve(why)
            raise ZKeyFileLoadError(why)
        self.size = lib().kzp_pool_size(self._h)
        self.devices = [lib().kzp_pool_device(self._h, i) for i in range(self.size)]

    def close(self):
        if getattr(self, "_h", None):
            lib().kzp_pool_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _finish(self, rc, out, err, ms, slot):
        if rc != 0:
            FullProver._raise_prover_error(err.value)
        return _take_string(out), {"prover_time": ms.value, "slot": slot.value}

    def prove(self, wtns_path: str, r: Optional[bytes] = None, s: Optional[bytes] = None):
        out, err, ms, slot = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = lib().kzp_pool_prove(self._h, os.fsencode(wtns_path), r, s, ctypes.byref(out), ctypes.byref(err),
                                  ctypes.byref(ms), ctypes.byref(slot))
        return self._finish(rc, out, err, ms, slot)

    def prove_mem(self, witness: bytes, r: Optional[bytes] = None, s: Optional[bytes] = None):
        out, err, ms, slot = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = lib().kzp_pool_prove_mem(self._h, witness, len(witness) // 32, r, s, ctypes.byref(out),
                                      ctypes.byref(err), ctypes.byref(ms), ctypes.byref(slot))
        return self._finish(rc, out, err, ms, slot)

    def set_verify(self, on: bool = True):
        """Verify every proof under the zkey's VK before returning it (host pairing check, after the prover is released)."""
        _check(lib().kzp_pool_set_verify(self._h, 1 if on else 0))

    def stats(self):
        arr, mw = (ctypes.c_uint64 * max(1, self.size))(), ctypes.c_uint64()
        n = lib().kzp_pool_stats(self._h, arr, self.size, ctypes.byref(mw))
        return {"proofs_per_slot": list(arr[:n]), "max_waiting": mw.value}


def pool_sched_selftest(slots: int, threads: int, jobs_per_thread: int, hold_us: int = 200):
    """Host-only run of the pool's checkout queue: (jobs per slot, max simultaneous holders of one slot, deepest queue)."""
    arr, worst, mw = (ctypes.c_uint64 * slots)(), ctypes.c_int(), ctypes.c_uint64()
    _check(lib().kzp_pool_sched_selftest(slots, threads, jobs_per_thread, hold_us, arr, ctypes.byref(worst),
                                         ctypes.byref(mw)))
    return list(arr), worst.value, mw.value


# ---- component wrappers ------------------------------------------------------------------------------------
def fr_ntt(data: bytes, inverse: bool = False, device: int = -1) -> bytes:
    """FFT<Fr>::fft / ifft (fft.cpp:192-246): Montgomery elements, natural order in and out."""
    buf = ctypes.create_string_buffer(data, len(data))
    _check(lib().kzp_fr_ntt(buf, len(data) // 32, 1 if inverse else 0, device))
    return buf.raw


def fr_coset_chain(data: bytes, device: int = -1) -> bytes:
    buf = ctypes.create_string_buffer(data, len(data))
    _check(lib().kzp_fr_coset_chain(buf, len(data) // 32, device))
    return buf.raw


def fr_ntt_bench(log_n: int, iters: int = 10, device: int = -1) -> float:
    ms = ctypes.c_float()
    _check(lib().kzp_fr_ntt_bench(log_n, iters, device, ctypes.byref(ms)))
    return ms.value


class Msm:
    """Curve::multiMulByScalar (curve.hpp:209-215) with the bases resident on the GPU."""

    def __init__(self, group: int, bases: bytes, device: int = -1, window_bits: int = 0, two_level: bool = False):
        """window_bits: signed-digit window size (16..22; 0 = the default, 16); two_level forces the two-pass digit
        sort that every window size other than 16 uses."""
        self.group = group
        self.point_bytes = 64 if group == 0 else 128
        self.n = len(bases) // self.point_bytes
        if window_bits or two_level:
            self._h = lib().kzp_msm_new_ex(group, bases, self.n, device, window_bits, 1 if two_level else 0)
        else:
            self._h = lib().kzp_msm_new(group, bases, self.n, device)
        if not self._h:
            raise KzpError("kzp_msm_new failed: " + last_error())

    def run(self, scalars: bytes) -> bytes:
        assert len(scalars) == self.n * 32
        out = ctypes.create_string_buffer(self.point_bytes)
        _check(lib().kzp_msm_run(self._h, scalars, out))
        return out.raw

    def bench(self, scalars: bytes, iters: int = 5):
        ms, ent = ctypes.c_float(), ctypes.c_uint64()
        _check(lib().kzp_msm_bench(self._h, scalars, iters, ctypes.byref(ms), ctypes.byref(ent)))
        return ms.value, ent.value

    def close(self):
        if getattr(self, "_h", None):
            lib().kzp_msm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def field_op(field: int, op: int, a: bytes, b: Optional[bytes] = None, device: int = -1) -> bytes:
    esz = 64 if field == 2 else 32
    out = ctypes.create_string_buffer(len(a))
    _check(lib().kzp_field_op(int(field), int(op), a, b, out, len(a) // esz, device))
    return out.raw


def point_op(group: int, op: int, p: bytes, q: Optional[bytes] = None, device: int = -1) -> bytes:
    psz = 128 if group == 0 else 256
    out = ctypes.create_string_buffer(len(p))
    _check(lib().kzp_point_op(group, op, p, q, out, len(p) // psz, device))
    return out.raw


def host_assemble(zkey_path: str, partials: Sequence[bytes], r: Optional[bytes] = None, s: Optional[bytes] = None):
    """Rank-0 step of the sharded mode, host only: sum the shards' partial MSM results, blind, print.
    Returns (proof_json, msm_results_384_bytes)."""
    out = ctypes.c_void_p()
    art = ctypes.create_string_buffer(384)
    _check(lib().kzp_host_assemble(os.fsencode(zkey_path), b"".join(partials), len(partials), r, s,
                                   ctypes.byref(out), art))
    return _take_string(out), art.raw


def host_pairing_check(g1: bytes, g2: bytes) -> bool:
    """prod_i e(P_i, Q_i) == 1 on the host (csrc/pairing.hpp); points in zkey byte layout."""
    res = ctypes.c_int()
    rc = lib().kzp_host_pairing_check(g1, g2, len(g1) // 64, ctypes.byref(res))
    if rc != 0:
        raise KzpError("kzp_host_pairing_check: " + lib().kzp_verify_last_error().decode())
    return bool(res.value)


def host_verify(zkey_path: str, proof_json: str, public: Sequence[int]) -> bool:
    """Groth16 verification under the VK inside the zkey — the check prover-service performs after every proof
    (prover_handler.rs:329-336), done natively on the host next to the prover."""
    res = ctypes.c_int()
    pub = b"".join(int(v).to_bytes(32, "little") for v in public)
    rc = lib().kzp_host_verify(os.fsencode(zkey_path), proof_json.encode(), pub, len(public), ctypes.byref(res))
    if rc != 0:
        raise KzpError("kzp_host_verify: " + lib().kzp_verify_last_error().decode())
    return bool(res.value)


def host_pack_witness_slice(values: bytes):
    """Packs up to 32768 witness values the way the staging workers do; returns (small bytes, flag bytes, full-width
    values, bytes that cross PCIe)."""
    n = len(values) // 32
    cap = 32768 + 4096 + 32768 * 32
    raw = ctypes.create_string_buffer(cap + 64)
    base = ctypes.addressof(raw)
    off = (-base) % 64
    nb, nf = ctypes.c_uint64(), ctypes.c_uint32()
    _check(lib().kzp_host_pack_witness_slice(values, n, ctypes.c_void_p(base + off), cap, ctypes.byref(nb), ctypes.byref(nf)))
    blob = raw.raw[off:off + cap]
    return blob[:32768], blob[32768:32768 + 4096], blob[36864:36864 + 32 * nf.value], nb.value


def imad_peak(iters: int = 4096, device: int = -1):
    """Returns (multiply-adds per second, ms) of the dependent-free IMAD.WIDE probe."""
    ms, n = ctypes.c_float(), ctypes.c_uint64()
    _check(lib().kzp_imad_peak(iters, device, ctypes.byref(ms), ctypes.byref(n)))
    return n.value / (ms.value * 1e-3), ms.value
