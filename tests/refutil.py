"""ctypes views of the two CPU checkers: oracle/_ref (the reference itself) and oracle/libkzp_port.so (C port)."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libkzp_ref.so")
REF_ASM_PATH = os.path.join(ROOT, "oracle", "_ref", "libkzp_ref_asm.so")  # the reference with its x86-64 asm field path
PORT_PATH = os.path.join(ROOT, "oracle", "libkzp_port.so")


class Ref:
    def __init__(self, lib):
        self.lib = lib
        c = ctypes
        lib.kzp_ref_prove.argtypes = [c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p, c.c_int, c.c_char_p, c.c_size_t,
                                      c.POINTER(c.c_int)]
        lib.kzp_ref_dump.argtypes = [c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p]
        lib.kzp_ref_fr_ntt.argtypes = [c.c_char_p, c.c_uint64, c.c_int]
        lib.kzp_ref_msm_g1.argtypes = [c.c_char_p, c.c_char_p, c.c_uint64, c.c_char_p]
        lib.kzp_ref_msm_g2.argtypes = [c.c_char_p, c.c_char_p, c.c_uint64, c.c_char_p]
        for f in (lib.kzp_ref_fr_op, lib.kzp_ref_fq_op, lib.kzp_ref_fq2_op):
            f.argtypes = [c.c_int, c.c_char_p, c.c_char_p, c.c_char_p, c.c_uint64]
        lib.kzp_ref_prover_new.restype = c.c_void_p
        lib.kzp_ref_prover_new.argtypes = [c.c_char_p, c.c_int, c.POINTER(c.c_int)]
        lib.kzp_ref_prover_free.argtypes = [c.c_void_p]
        lib.kzp_ref_prover_prove.argtypes = [c.c_void_p, c.c_char_p, c.c_char_p, c.c_char_p, c.c_int, c.c_char_p,
                                             c.c_size_t, c.POINTER(c.c_int)]

    def prove(self, zkey, wtns, r=None, s=None):
        buf = ctypes.create_string_buffer(8192)
        ms = ctypes.c_int()
        rc = self.lib.kzp_ref_prove(zkey.encode(), wtns.encode(), r, s, 1, buf, 8192, ctypes.byref(ms))
        assert rc == 0, "reference prover failed rc=%d" % rc
        return buf.value.decode(), ms.value

    def dump(self, zkey, wtns, domain, want_ab=False):
        """-> (ab or None, h bytes, msm 384 bytes) by re-driving the reference's stages (oracle/ref_harness.cpp)."""
        h = ctypes.create_string_buffer(domain * 32)
        m = ctypes.create_string_buffer(384)
        ab = ctypes.create_string_buffer(domain * 64) if want_ab else None
        rc = self.lib.kzp_ref_dump(zkey.encode(), wtns.encode(), ab, h, m)
        assert rc == 0
        return (ab.raw if ab else None), h.raw, m.raw

    def ntt(self, data, inverse=False):
        buf = ctypes.create_string_buffer(data, len(data))
        assert self.lib.kzp_ref_fr_ntt(buf, len(data) // 32, 1 if inverse else 0) == 0
        return buf.raw

    def msm(self, group, bases, scalars):
        n = len(scalars) // 32
        out = ctypes.create_string_buffer(64 if group == 0 else 128)
        (self.lib.kzp_ref_msm_g1 if group == 0 else self.lib.kzp_ref_msm_g2)(bases, scalars, n, out)
        return out.raw

    def field_op(self, field, op, a, b=None):
        esz = 64 if field == 2 else 32
        out = ctypes.create_string_buffer(len(a))
        fn = (self.lib.kzp_ref_fr_op, self.lib.kzp_ref_fq_op, self.lib.kzp_ref_fq2_op)[field]
        fn(op, a, b, out, len(a) // esz)
        return out.raw


class Port:
    def __init__(self, lib):
        self.lib = lib
        c = ctypes
        lib.kzp_port_make_setup.argtypes = [c.c_uint32, c.c_uint32, c.c_uint64, c.c_char_p, c.c_char_p,
                                            c.POINTER(c.c_uint64), c.c_char_p]
        lib.kzp_port_prove.argtypes = [c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p, c.c_char_p, c.c_size_t,
                                       c.c_char_p, c.c_char_p, c.POINTER(c.c_double)]
        lib.kzp_port_msm_g1.argtypes = [c.c_char_p, c.c_char_p, c.c_uint64, c.c_char_p]
        lib.kzp_port_msm_g2.argtypes = [c.c_char_p, c.c_char_p, c.c_uint64, c.c_char_p]
        lib.kzp_port_fr_ntt.argtypes = [c.c_char_p, c.c_uint64, c.c_int]
        lib.kzp_port_field_op.argtypes = [c.c_int, c.c_int, c.c_char_p, c.c_char_p, c.c_char_p, c.c_uint64]
        lib.kzp_port_g1_gen_mul.argtypes = [c.c_char_p, c.c_char_p]
        lib.kzp_port_g2_gen_mul.argtypes = [c.c_char_p, c.c_char_p]

    def make_setup(self, n_constraints, n_vars, seed, zkey, wtns):
        info = (ctypes.c_uint64 * 8)()
        trap = ctypes.create_string_buffer(160)
        rc = self.lib.kzp_port_make_setup(n_constraints, n_vars, seed, zkey.encode(), wtns.encode(), info, trap)
        assert rc == 0
        keys = ("n_vars", "n_public", "domain", "n_coefs", "n_rows", "public_input")
        d = dict(zip(keys, list(info)[:6]))
        d["trapdoor"] = trap.raw
        return d

    def prove(self, zkey, wtns, r, s, domain=None, want_artefacts=False):
        buf = ctypes.create_string_buffer(8192)
        sec = ctypes.c_double()
        h = ctypes.create_string_buffer(domain * 32) if want_artefacts else None
        m = ctypes.create_string_buffer(384) if want_artefacts else None
        rc = self.lib.kzp_port_prove(zkey.encode(), wtns.encode(), r, s, buf, 8192, h, m, ctypes.byref(sec))
        assert rc == 0, "port prover failed rc=%d" % rc
        return buf.value.decode(), sec.value, (h.raw if h else None), (m.raw if m else None)

    def msm(self, group, bases, scalars):
        n = len(scalars) // 32
        out = ctypes.create_string_buffer(64 if group == 0 else 128)
        (self.lib.kzp_port_msm_g1 if group == 0 else self.lib.kzp_port_msm_g2)(bases, scalars, n, out)
        return out.raw

    def ntt(self, data, inverse=False):
        buf = ctypes.create_string_buffer(data, len(data))
        assert self.lib.kzp_port_fr_ntt(buf, len(data) // 32, 1 if inverse else 0) == 0
        return buf.raw

    def field_op(self, field, op, a, b=None):
        out = ctypes.create_string_buffer(len(a))
        self.lib.kzp_port_field_op(field, op, a, b, out, len(a) // 32)
        return out.raw

    def g1_gen_mul(self, k):
        out = ctypes.create_string_buffer(64)
        self.lib.kzp_port_g1_gen_mul(k, out)
        return out.raw

    def g2_gen_mul(self, k):
        out = ctypes.create_string_buffer(128)
        self.lib.kzp_port_g2_gen_mul(k, out)
        return out.raw


def load_ref():
    if not os.path.exists(REF_PATH):
        return None
    return Ref(ctypes.CDLL(REF_PATH))


def cpu_has_adx_bmi2():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " adx" in flags and " bmi2" in flags


def load_ref_asm():
    """The reference built with its own x86-64 assembly field arithmetic (oracle/Makefile ref_asm); None when it was
    not built or this CPU lacks MULX/ADCX/ADOX."""
    if not os.path.exists(REF_ASM_PATH) or not cpu_has_adx_bmi2():
        return None
    return Ref(ctypes.CDLL(REF_ASM_PATH))


def load_port():
    if not os.path.exists(PORT_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    return Port(ctypes.CDLL(PORT_PATH))
