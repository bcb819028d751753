"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import
this module (task rule ③). It is a plain-Python (arbitrary precision ``int``) restatement of
the arithmetic the reference prover performs, small-case only; big cases go through
``oracle/_ref`` (the reference's own sources compiled in place) or ``oracle/kzp_port.c``.

Citations are relative to /root/reference/rust-rapidsnark/rapidsnark/src (``RS/``).

Pinned against (tests/test_oracle_*.py):
  * field constants          RS/fr_raw_generic.cpp:5-7, RS/fq_raw_generic.cpp:6-8
  * f2_simpleMul             RS/alt_bn128_test.cpp:12-29
  * g1/g2_expToOrder         RS/alt_bn128_test.cpp:138-170
  * multiExp (sum i^2)       RS/alt_bn128_test.cpp:172-212
  * multiExp2 2-point KAT    RS/alt_bn128_test.cpp:215-248
  * fft round trip           RS/alt_bn128_test.cpp:250-271
  * toy zkey/wtns/vk triple  prover-service/resources/toy_circuit/*
  * outputs of oracle/_ref (the reference itself) on the committed golden fixtures
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------- constants
# RS/fullprover.cpp:140-143 (r) ; RS/fq_raw_generic.cpp:6 (q)
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
MONT_R = 1 << 256  # both fields use R = 2^256 (4 x 64-bit limbs), RS/fr_raw_generic.cpp:107-148
G1_B = 3  # RS/alt_bn128.hpp:45
# RS/alt_bn128.hpp:46-50 : G2 curve coefficient b' = 3/(9+u) and the G2 generator
G2_B = (
    19485874751759354771024239261021720505790618469301721065564631296452457478373,
    266929791119991161246907387137283842545076965332900288569378510910307636690,
)
G1_GEN = (1, 2)
G2_GEN = (
    (
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ),
    (
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ),
)
FR_TWO_ADICITY = 28


def to_mont(x: int, mod: int) -> int:
    return (x * MONT_R) % mod


def from_mont(x: int, mod: int) -> int:
    return (x * pow(MONT_R, -1, mod)) % mod


def mont_mul(a: int, b: int, mod: int) -> int:
    """Fr_rawMMul / Fq_rawMMul (RS/fr_raw_generic.cpp:107-148): a*b*R^-1 mod p, canonical."""
    return (a * b * pow(MONT_R, -1, mod)) % mod


_RINV = {R_MOD: pow(MONT_R, -1, R_MOD), Q_MOD: pow(MONT_R, -1, Q_MOD)}


def le32(x: int) -> bytes:
    return x.to_bytes(32, "little")


def from_le(b: bytes) -> int:
    return int.from_bytes(b, "little")


# --------------------------------------------------------------------------- Fq2
# F2Field with non-residue -1 (RS/alt_bn128.hpp:43, RS/f2field.cpp:94-189)
Fq2 = Tuple[int, int]


def f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)


def f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)


def f2_neg(a: Fq2) -> Fq2:
    return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)


def f2_mul(a: Fq2, b: Fq2) -> Fq2:
    # (a0 + a1 u)(b0 + b1 u), u^2 = -1  (RS/f2field.cpp:122-141)
    return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)


def f2_sqr(a: Fq2) -> Fq2:
    return f2_mul(a, a)


def f2_inv(a: Fq2) -> Fq2:
    # RS/f2field.cpp:178-189 : 1/(a0^2 + a1^2) * (a0 - a1 u)
    t = pow((a[0] * a[0] + a[1] * a[1]) % Q_MOD, -1, Q_MOD)
    return ((a[0] * t) % Q_MOD, (-a[1] * t) % Q_MOD)


def f2_muls(a: Fq2, k: int) -> Fq2:
    return ((a[0] * k) % Q_MOD, (a[1] * k) % Q_MOD)


# --------------------------------------------------------------------------- curves (affine, None = infinity)
G1Point = Optional[Tuple[int, int]]
G2Point = Optional[Tuple[Fq2, Fq2]]


def g1_is_on_curve(p: G1Point) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - G1_B) % Q_MOD == 0


def g1_neg(p: G1Point) -> G1Point:
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_add(p: G1Point, q: G1Point) -> G1Point:
    """Group law; matches Curve::add incl. P==Q -> dbl and P==-Q -> infinity (RS/curve.cpp:132,219,291)."""
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def g1_mul(p: G1Point, k: int) -> G1Point:
    k %= R_MOD if k >= 0 else R_MOD
    # Jacobian double-and-add for speed; result is the unique group element.
    if p is None or k == 0:
        return None
    X, Y, Z = p[0], p[1], 1
    RX, RY, RZ = 0, 1, 0
    for bit in bin(k)[2:]:
        RX, RY, RZ = _jac_dbl(RX, RY, RZ)
        if bit == "1":
            RX, RY, RZ = _jac_add(RX, RY, RZ, X, Y, Z)
    return _jac_to_affine(RX, RY, RZ)


def _jac_dbl(X, Y, Z):
    if Z == 0:
        return X, Y, Z
    A = X * X % Q_MOD
    B = Y * Y % Q_MOD
    C = B * B % Q_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % Q_MOD
    E = 3 * A % Q_MOD
    F = E * E % Q_MOD
    X3 = (F - 2 * D) % Q_MOD
    Y3 = (E * (D - X3) - 8 * C) % Q_MOD
    Z3 = 2 * Y * Z % Q_MOD
    return X3, Y3, Z3


def _jac_add(X1, Y1, Z1, X2, Y2, Z2):
    if Z1 == 0:
        return X2, Y2, Z2
    if Z2 == 0:
        return X1, Y1, Z1
    Z1Z1 = Z1 * Z1 % Q_MOD
    Z2Z2 = Z2 * Z2 % Q_MOD
    U1 = X1 * Z2Z2 % Q_MOD
    U2 = X2 * Z1Z1 % Q_MOD
    S1 = Y1 * Z2 * Z2Z2 % Q_MOD
    S2 = Y2 * Z1 * Z1Z1 % Q_MOD
    if U1 == U2:
        if S1 == S2:
            return _jac_dbl(X1, Y1, Z1)
        return 0, 1, 0
    H = (U2 - U1) % Q_MOD
    I = 4 * H * H % Q_MOD
    J = H * I % Q_MOD
    r = 2 * (S2 - S1) % Q_MOD
    V = U1 * I % Q_MOD
    X3 = (r * r - J - 2 * V) % Q_MOD
    Y3 = (r * (V - X3) - 2 * S1 * J) % Q_MOD
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % Q_MOD
    return X3, Y3, Z3


def _jac_to_affine(X, Y, Z) -> G1Point:
    if Z == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 * zi % Q_MOD)


def g2_is_on_curve(p: G2Point) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), G2_B)) == (0, 0)


def g2_neg(p: G2Point) -> G2Point:
    return None if p is None else (p[0], f2_neg(p[1]))


def g2_add(p: G2Point, q: G2Point) -> G2Point:
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_muls(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(p: G2Point, k: int) -> G2Point:
    k %= R_MOD
    res: G2Point = None
    if p is None:
        return None
    # projective-free but inversion heavy; G2 muls are rare in the oracle. Use a Jacobian loop.
    X, Y, Z = p[0], p[1], (1, 0)
    RX, RY, RZ = (0, 0), (1, 0), (0, 0)
    for bit in bin(k)[2:] if k else "":
        RX, RY, RZ = _jac2_dbl(RX, RY, RZ)
        if bit == "1":
            RX, RY, RZ = _jac2_add(RX, RY, RZ, X, Y, Z)
    if RZ == (0, 0):
        return None
    zi = f2_inv(RZ)
    zi2 = f2_sqr(zi)
    return (f2_mul(RX, zi2), f2_mul(RY, f2_mul(zi2, zi)))


def _jac2_dbl(X, Y, Z):
    if Z == (0, 0):
        return X, Y, Z
    A = f2_sqr(X)
    B = f2_sqr(Y)
    C = f2_sqr(B)
    t = f2_add(X, B)
    D = f2_muls(f2_sub(f2_sub(f2_sqr(t), A), C), 2)
    E = f2_muls(A, 3)
    F = f2_sqr(E)
    X3 = f2_sub(F, f2_muls(D, 2))
    Y3 = f2_sub(f2_mul(E, f2_sub(D, X3)), f2_muls(C, 8))
    Z3 = f2_muls(f2_mul(Y, Z), 2)
    return X3, Y3, Z3


def _jac2_add(X1, Y1, Z1, X2, Y2, Z2):
    if Z1 == (0, 0):
        return X2, Y2, Z2
    if Z2 == (0, 0):
        return X1, Y1, Z1
    Z1Z1 = f2_sqr(Z1)
    Z2Z2 = f2_sqr(Z2)
    U1 = f2_mul(X1, Z2Z2)
    U2 = f2_mul(X2, Z1Z1)
    S1 = f2_mul(f2_mul(Y1, Z2), Z2Z2)
    S2 = f2_mul(f2_mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return _jac2_dbl(X1, Y1, Z1)
        return (0, 0), (1, 0), (0, 0)
    H = f2_sub(U2, U1)
    I = f2_muls(f2_sqr(H), 4)
    J = f2_mul(H, I)
    r = f2_muls(f2_sub(S2, S1), 2)
    V = f2_mul(U1, I)
    X3 = f2_sub(f2_sub(f2_sqr(r), J), f2_muls(V, 2))
    Y3 = f2_sub(f2_mul(r, f2_sub(V, X3)), f2_muls(f2_mul(S1, J), 2))
    Z3 = f2_mul(f2_sub(f2_sub(f2_sqr(f2_add(Z1, Z2)), Z1Z1), Z2Z2), H)
    return X3, Y3, Z3


# --------------------------------------------------------------------------- byte encodings
def g1_to_zkey_bytes(p: G1Point) -> bytes:
    """Affine, Montgomery, LE; infinity = 64 zero bytes."""
    if p is None:
        return bytes(64)
    return le32(to_mont(p[0], Q_MOD)) + le32(to_mont(p[1], Q_MOD))


def g1_from_zkey_bytes(b: bytes) -> G1Point:
    x, y = from_le(b[:32]), from_le(b[32:64])
    if x == 0 and y == 0:
        return None
    return (from_mont(x, Q_MOD), from_mont(y, Q_MOD))


def g2_to_zkey_bytes(p: G2Point) -> bytes:
    if p is None:
        return bytes(128)
    (xa, xb), (ya, yb) = p
    return b"".join(le32(to_mont(v, Q_MOD)) for v in (xa, xb, ya, yb))


def g2_from_zkey_bytes(b: bytes) -> G2Point:
    v = [from_le(b[i * 32 : (i + 1) * 32]) for i in range(4)]
    if not any(v):
        return None
    v = [from_mont(t, Q_MOD) for t in v]
    return ((v[0], v[1]), (v[2], v[3]))


def g1_to_canonical_bytes(p: G1Point) -> bytes:
    if p is None:
        return bytes(64)
    return le32(p[0]) + le32(p[1])


def g2_to_canonical_bytes(p: G2Point) -> bytes:
    if p is None:
        return bytes(128)
    (xa, xb), (ya, yb) = p
    return le32(xa) + le32(xb) + le32(ya) + le32(yb)


# --------------------------------------------------------------------------- NTT (RS/fft.cpp)
def fr_nqr() -> int:
    """Smallest quadratic non-residue >= 2 (RS/fft.cpp:60-66). 5 for BN254 Fr."""
    n = 2
    while pow(n, (R_MOD - 1) // 2, R_MOD) == 1:
        n += 1
    return n


def fr_root_of_unity(log_n: int) -> int:
    """roots[1] of an FFT table with s = log_n (RS/fft.cpp:72-96): nqr^((r-1)/2^s)."""
    assert log_n <= FR_TWO_ADICITY
    return pow(fr_nqr(), (R_MOD - 1) >> log_n, R_MOD)


def _bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def fr_fft(a: Sequence[int], table_log: Optional[int] = None) -> List[int]:
    """FFT::fft (RS/fft.cpp:192-219): bit-reversal then log n DIT stages. Plain (non-Montgomery) ints.

    ``table_log`` is the log2 size of the roots table the reference instance was built with
    (the prover builds FFT(2*domainSize), RS/groth16.hpp:96); root(s, j) = w_table^(j << (S - s)).
    The transform itself does not depend on it (same primitive 2^s-th roots).
    """
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    S = table_log if table_log is not None else max(log_n, 1)
    w_table = fr_root_of_unity(S)
    a = list(a)
    for i in range(n):
        r = _bitrev(i, log_n)
        if i > r:
            a[i], a[r] = a[r], a[i]
    for s in range(1, log_n + 1):
        m = 1 << s
        mdiv2 = m >> 1
        wm = pow(w_table, 1 << (S - s), R_MOD)
        tw = [1] * mdiv2
        for j in range(1, mdiv2):
            tw[j] = tw[j - 1] * wm % R_MOD
        for k in range(0, n, m):
            for j in range(mdiv2):
                t = tw[j] * a[k + j + mdiv2] % R_MOD
                u = a[k + j]
                a[k + j] = (u + t) % R_MOD
                a[k + j + mdiv2] = (u - t) % R_MOD
    return a


def fr_ifft(a: Sequence[int], table_log: Optional[int] = None) -> List[int]:
    """FFT::ifft (RS/fft.cpp:222-246): fft, then a[i] <-> a[n-i] and scale by 2^-log n."""
    n = len(a)
    f = fr_fft(a, table_log)
    ninv = pow(n, -1, R_MOD)
    out = [0] * n
    for i in range(n):
        out[i] = f[(n - i) % n] * ninv % R_MOD
    return out


# --------------------------------------------------------------------------- MSM (RS/multiexp.cpp)
def msm_window_bits(n: int) -> int:
    """bitsPerChunk = clamp(floor(log2(n/2)), 2, 16) (RS/multiexp.cpp:206-211, misc.hpp:12-20)."""
    v = n // 2
    c = v.bit_length() - 1 if v > 0 else 0
    return max(2, min(16, c))


def msm_get_chunk(scalar_le: bytes, chunk_idx: int, bits_per_chunk: int) -> int:
    """ParallelMultiexp::getChunk (RS/multiexp.cpp:26-41), 32-byte scalars."""
    scalar_size = len(scalar_le)
    bit_start = chunk_idx * bits_per_chunk
    byte_start = bit_start // 8
    eff = bits_per_chunk
    if byte_start > scalar_size - 8:
        byte_start = scalar_size - 8
    if bit_start + bits_per_chunk > scalar_size * 8:
        eff = scalar_size * 8 - bit_start
    shift = bit_start - byte_start * 8
    v = int.from_bytes(scalar_le[byte_start : byte_start + 8], "little")
    v >>= shift
    v &= (1 << eff) - 1
    return v


def msm_pippenger(add, neg, mul, bases: Sequence, scalars: Sequence[int]):
    """Restatement of ParallelMultiexp::multiexp (RS/multiexp.cpp:183-245) with one thread:
    unsigned c-bit windows, bucket fill skipping infinity bases and zero digits, recursive-halving
    reduce (:133-180) restated as its closed form sum_d d*acc[d], Horner window combine (:236-243)."""
    n = len(bases)
    if n == 0:
        return None
    if n == 1:
        return mul(bases[0], scalars[0])
    c = msm_window_bits(n)
    n_chunks = (32 * 8 - 1) // c + 1
    chunk_results = []
    sb = [le32(s) for s in scalars]
    for ch in range(n_chunks):
        accs = [None] * (1 << c)
        for i in range(n):
            if bases[i] is None:
                continue
            d = msm_get_chunk(sb[i], ch, c)
            if d:
                accs[d] = add(accs[d], bases[i])
        running = None
        total = None
        for d in range((1 << c) - 1, 0, -1):
            running = add(running, accs[d])
            total = add(total, running)
        chunk_results.append(total)
    r = chunk_results[-1]
    for j in range(n_chunks - 2, -1, -1):
        for _ in range(c):
            r = add(r, r)
        r = add(r, chunk_results[j])
    return r


def msm_g1(bases: Sequence[G1Point], scalars: Sequence[int]) -> G1Point:
    return msm_pippenger(g1_add, g1_neg, g1_mul, bases, scalars)


def msm_g2(bases: Sequence[G2Point], scalars: Sequence[int]) -> G2Point:
    return msm_pippenger(g2_add, g2_neg, g2_mul, bases, scalars)


def msm_naive_g1(bases, scalars) -> G1Point:
    acc = None
    for b, s in zip(bases, scalars):
        if b is not None and s:
            acc = g1_add(acc, g1_mul(b, s))
    return acc


def msm_naive_g2(bases, scalars) -> G2Point:
    acc = None
    for b, s in zip(bases, scalars):
        if b is not None and s:
            acc = g2_add(acc, g2_mul(b, s))
    return acc


# --------------------------------------------------------------------------- iden3 binfile container (RS/binfile_utils.cpp:13-58)
def read_binfile(path: str, expected_type: bytes, max_version: int):
    data = open(path, "rb").read()
    if data[:4] != expected_type:
        raise ValueError("Invalid file type")
    version, n_sections = struct.unpack_from("<II", data, 4)
    if version > max_version:
        raise ValueError("Invalid version")
    pos = 12
    sections = {}
    for _ in range(n_sections):
        sid, ssize = struct.unpack_from("<IQ", data, pos)
        pos += 12
        sections.setdefault(sid, []).append(data[pos : pos + ssize])
        pos += ssize
    return sections


def write_binfile(path: str, ftype: bytes, version: int, sections: Sequence[Tuple[int, bytes]]):
    with open(path, "wb") as f:
        f.write(ftype)
        f.write(struct.pack("<II", version, len(sections)))
        for sid, payload in sections:
            f.write(struct.pack("<IQ", sid, len(payload)))
            f.write(payload)


@dataclass
class ZKey:
    """Decoded groth16 zkey (RS/zkey_utils.hpp:48-87, RS/fullprover.cpp:164-174)."""

    n_vars: int
    n_public: int
    domain_size: int
    alpha1: G1Point
    beta1: G1Point
    beta2: G2Point
    gamma2: G2Point
    delta1: G1Point
    delta2: G2Point
    ic: List[G1Point]
    coefs: List[Tuple[int, int, int, int]]  # (m, c, s, value) value = plain int (file holds value*R^2)
    points_a: List[G1Point]
    points_b1: List[G1Point]
    points_b2: List[G2Point]
    points_c: List[G1Point]
    points_h: List[G1Point]
    q: int = Q_MOD
    r: int = R_MOD


def read_zkey(path: str) -> ZKey:
    sec = read_binfile(path, b"zkey", 1)
    (protocol,) = struct.unpack_from("<I", sec[1][0], 0)
    if protocol != 1:
        raise ValueError("zkey file is not groth16")
    h = sec[2][0]
    pos = 0
    (n8q,) = struct.unpack_from("<I", h, pos)
    pos += 4
    q = from_le(h[pos : pos + n8q])
    pos += n8q
    (n8r,) = struct.unpack_from("<I", h, pos)
    pos += 4
    r = from_le(h[pos : pos + n8r])
    pos += n8r
    n_vars, n_public, domain_size = struct.unpack_from("<III", h, pos)
    pos += 12
    alpha1 = g1_from_zkey_bytes(h[pos : pos + 64])
    pos += 64
    beta1 = g1_from_zkey_bytes(h[pos : pos + 64])
    pos += 64
    beta2 = g2_from_zkey_bytes(h[pos : pos + 128])
    pos += 128
    gamma2 = g2_from_zkey_bytes(h[pos : pos + 128])
    pos += 128
    delta1 = g1_from_zkey_bytes(h[pos : pos + 64])
    pos += 64
    delta2 = g2_from_zkey_bytes(h[pos : pos + 128])
    pos += 128

    def g1s(b):
        return [g1_from_zkey_bytes(b[i : i + 64]) for i in range(0, len(b), 64)]

    def g2s(b):
        return [g2_from_zkey_bytes(b[i : i + 128]) for i in range(0, len(b), 128)]

    c4 = sec[4][0]
    n_coefs = len(c4) // (12 + n8r)  # RS/zkey_utils.hpp:84 (the 4-byte count is absorbed by the division)
    coefs = []
    r2inv = pow(MONT_R * MONT_R, -1, R_MOD)
    for i in range(n_coefs):
        off = 4 + i * 44
        m, c, s = struct.unpack_from("<III", c4, off)
        v = from_le(c4[off + 12 : off + 44])
        coefs.append((m, c, s, v * r2inv % R_MOD))
    return ZKey(
        n_vars, n_public, domain_size, alpha1, beta1, beta2, gamma2, delta1, delta2,
        g1s(sec[3][0]), coefs, g1s(sec[5][0]), g1s(sec[6][0]), g2s(sec[7][0]), g1s(sec[8][0]),
        g1s(sec[9][0]), q, r,
    )


def write_zkey(path: str, zk: ZKey):
    hdr = struct.pack("<I", 32) + le32(Q_MOD) + struct.pack("<I", 32) + le32(R_MOD)
    hdr += struct.pack("<III", zk.n_vars, zk.n_public, zk.domain_size)
    hdr += g1_to_zkey_bytes(zk.alpha1) + g1_to_zkey_bytes(zk.beta1) + g2_to_zkey_bytes(zk.beta2)
    hdr += g2_to_zkey_bytes(zk.gamma2) + g1_to_zkey_bytes(zk.delta1) + g2_to_zkey_bytes(zk.delta2)
    c4 = struct.pack("<I", len(zk.coefs))
    r2 = MONT_R * MONT_R % R_MOD
    for m, c, s, v in zk.coefs:
        c4 += struct.pack("<III", m, c, s) + le32(v * r2 % R_MOD)
    secs = [
        (1, struct.pack("<I", 1)),
        (2, hdr),
        (3, b"".join(g1_to_zkey_bytes(p) for p in zk.ic)),
        (4, c4),
        (5, b"".join(g1_to_zkey_bytes(p) for p in zk.points_a)),
        (6, b"".join(g1_to_zkey_bytes(p) for p in zk.points_b1)),
        (7, b"".join(g2_to_zkey_bytes(p) for p in zk.points_b2)),
        (8, b"".join(g1_to_zkey_bytes(p) for p in zk.points_c)),
        (9, b"".join(g1_to_zkey_bytes(p) for p in zk.points_h)),
        (10, struct.pack("<I", 0)),
    ]
    write_binfile(path, b"zkey", 1, secs)


def read_wtns(path: str) -> List[int]:
    """RS/wtns_utils.hpp:28-43 + RS/fullprover.cpp:223-224 (canonical, non-Montgomery values)."""
    sec = read_binfile(path, b"wtns", 2)
    h = sec[1][0]
    (n8,) = struct.unpack_from("<I", h, 0)
    prime = from_le(h[4 : 4 + n8])
    if prime != R_MOD:
        raise ValueError("witness uses a different curve")
    (n,) = struct.unpack_from("<I", h, 4 + n8)
    d = sec[2][0]
    return [from_le(d[i * 32 : (i + 1) * 32]) for i in range(n)]


def write_wtns(path: str, w: Sequence[int]):
    h = struct.pack("<I", 32) + le32(R_MOD) + struct.pack("<I", len(w))
    write_binfile(path, b"wtns", 2, [(1, h), (2, b"".join(le32(v) for v in w))])


# --------------------------------------------------------------------------- Groth16 prove (RS/groth16.cpp:43-360)
def h_coefficients(zk: ZKey, w: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """Returns (a, b, h): a,b after SpMV (plain ints; the reference holds them in Montgomery form) and
    the H coefficients as canonical ints in natural order (RS/groth16.cpp:116-275)."""
    n = zk.domain_size
    a = [0] * n
    b = [0] * n
    for m, c, s, v in zk.coefs:  # :141-156
        if m == 0:
            a[c] = (a[c] + w[s] * v) % R_MOD
        else:
            b[c] = (b[c] + w[s] * v) % R_MOD
    cc = [x * y % R_MOD for x, y in zip(a, b)]  # :160-167
    log_n = n.bit_length() - 1
    w2n = fr_root_of_unity(log_n + 1)  # fft_.root(domainPower+1, 1)
    shift = [1] * n
    for i in range(1, n):
        shift[i] = shift[i - 1] * w2n % R_MOD
    out = []
    for vec in (a, b, cc):  # :172-262
        t = fr_ifft(vec, log_n + 1)
        t = [x * s % R_MOD for x, s in zip(t, shift)]
        out.append(fr_fft(t, log_n + 1))
    ea, eb, ec = out
    h = [(x * y - z) % R_MOD for x, y, z in zip(ea, eb, ec)]  # :266-275
    return a, b, h


@dataclass
class ProofArtefacts:
    a: List[int]
    b: List[int]
    h: List[int]
    msm_a: G1Point
    msm_b1: G1Point
    msm_b2: G2Point
    msm_c: G1Point
    msm_h: G1Point
    pi_a: G1Point = None
    pi_b: G2Point = None
    pi_c: G1Point = None
    json: str = ""


def proof_json(pi_a: G1Point, pi_b: G2Point, pi_c: G1Point) -> str:
    """Proof::toJson + nlohmann dump() (RS/groth16.cpp:379-410, RS/fullprover.cpp:246): compact, keys sorted.
    Infinity prints as ("0","0") because copy(affine, infinity) gives (0,0) (RS/curve.cpp:565-576)."""
    a = pi_a or (0, 0)
    b = pi_b or ((0, 0), (0, 0))
    c = pi_c or (0, 0)
    obj = {
        "pi_a": [str(a[0]), str(a[1]), "1"],
        "pi_b": [[str(b[0][0]), str(b[0][1])], [str(b[1][0]), str(b[1][1])], ["1", "0"]],
        "pi_c": [str(c[0]), str(c[1]), "1"],
        "protocol": "groth16",
    }
    return json.dumps(obj, separators=(",", ":"), sort_keys=True)


def groth16_prove(zk: ZKey, w: Sequence[int], r: int, s: int, naive_msm: bool = True) -> ProofArtefacts:
    mg1 = msm_naive_g1 if naive_msm else msm_g1
    mg2 = msm_naive_g2 if naive_msm else msm_g2
    a, b, h = h_coefficients(zk, w)
    pa = mg1(zk.points_a, w[: zk.n_vars])
    pb1 = mg1(zk.points_b1, w[: zk.n_vars])
    pb2 = mg2(zk.points_b2, w[: zk.n_vars])
    pc = mg1(zk.points_c, w[zk.n_public + 1 : zk.n_vars])
    ph = mg1(zk.points_h, h)
    art = ProofArtefacts(a, b, h, pa, pb1, pb2, pc, ph)
    # RS/groth16.cpp:328-352
    pi_a = g1_add(g1_add(pa, zk.alpha1), g1_mul(zk.delta1, r))
    pi_b = g2_add(g2_add(pb2, zk.beta2), g2_mul(zk.delta2, s))
    pib1 = g1_add(g1_add(pb1, zk.beta1), g1_mul(zk.delta1, s))
    pi_c = g1_add(pc, ph)
    pi_c = g1_add(pi_c, g1_mul(pi_a, s))
    pi_c = g1_add(pi_c, g1_mul(pib1, r))
    pi_c = g1_add(pi_c, g1_neg(g1_mul(zk.delta1, r * s % R_MOD)))
    art.pi_a, art.pi_b, art.pi_c = pi_a, pi_b, pi_c
    art.json = proof_json(pi_a, pi_b, pi_c)
    return art


# --------------------------------------------------------------------------- pairing (verification only)
# Fq12 = Fq[w]/(w^12 - 18 w^6 + 82); Fq2 embeds via u = w^6 - 9 (last bullet).
_FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63


def _p12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for i in range(22, 11, -1):
        top = t[i]
        if top:
            t[i - 6] += 18 * top
            t[i - 12] -= 82 * top
    return [v % Q_MOD for v in t[:12]]


def _p12_one():
    return [1] + [0] * 11


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_rounded_div(a, b):
    dega, degb = _poly_deg(a), _poly_deg(b)
    temp = list(a)
    o = [0] * len(a)
    binv = pow(b[degb], -1, Q_MOD)
    for i in range(dega - degb, -1, -1):
        q = temp[degb + i] * binv % Q_MOD
        o[i] = q
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - q * b[c]) % Q_MOD
    return o[: _poly_deg(o) + 1]


def _p12_inv(a):
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [v % Q_MOD for v in _FQ12_MOD] + [1]
    while _poly_deg(low):
        r = _poly_rounded_div(high, low)
        r += [0] * (13 - len(r))
        nm = list(hm)
        new = list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q_MOD
                new[i + j] = (new[i + j] - low[i] * r[j]) % Q_MOD
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, Q_MOD)
    return [v * inv0 % Q_MOD for v in lm[:12]]


def _p12_pow(a, e):
    res = _p12_one()
    base = a
    while e:
        if e & 1:
            res = _p12_mul(res, base)
        base = _p12_mul(base, base)
        e >>= 1
    return res


def _p12_scalar(x):
    return [x % Q_MOD] + [0] * 11


def _p12_add(a, b):
    return [(x + y) % Q_MOD for x, y in zip(a, b)]


def _p12_sub(a, b):
    return [(x - y) % Q_MOD for x, y in zip(a, b)]


def _twist(pt: G2Point):
    (x0, x1), (y0, y1) = pt
    xc = [(x0 - 9 * x1) % Q_MOD, x1]
    yc = [(y0 - 9 * y1) % Q_MOD, y1]
    nx = [0] * 12
    ny = [0] * 12
    # nx * w^2, ny * w^3
    nx[2], nx[8] = xc[0], xc[1]
    ny[3], ny[9] = yc[0], yc[1]
    return (nx, ny)


def _p12_pt_double(pt):
    x, y = pt
    lam = _p12_mul(_p12_mul(_p12_scalar(3), _p12_mul(x, x)), _p12_inv(_p12_mul(_p12_scalar(2), y)))
    nx = _p12_sub(_p12_sub(_p12_mul(lam, lam), x), x)
    ny = _p12_sub(_p12_mul(lam, _p12_sub(x, nx)), y)
    return (nx, ny)


def _p12_pt_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2 and y1 == y2:
        return _p12_pt_double(p1)
    if x1 == x2:
        return None
    lam = _p12_mul(_p12_sub(y2, y1), _p12_inv(_p12_sub(x2, x1)))
    nx = _p12_sub(_p12_sub(_p12_mul(lam, lam), x1), x2)
    ny = _p12_sub(_p12_mul(lam, _p12_sub(x1, nx)), y1)
    return (nx, ny)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = _p12_mul(_p12_sub(y2, y1), _p12_inv(_p12_sub(x2, x1)))
        return _p12_sub(_p12_mul(m, _p12_sub(xt, x1)), _p12_sub(yt, y1))
    if y1 == y2:
        m = _p12_mul(_p12_mul(_p12_scalar(3), _p12_mul(x1, x1)), _p12_inv(_p12_mul(_p12_scalar(2), y1)))
        return _p12_sub(_p12_mul(m, _p12_sub(xt, x1)), _p12_sub(yt, y1))
    return _p12_sub(xt, x1)


def miller_loop(q: G2Point, p: G1Point):
    """Optimal-ate Miller loop (no final exponentiation). Returns an Fq12 element (list of 12 ints)."""
    if q is None or p is None:
        return _p12_one()
    Q = _twist(q)
    P = (_p12_scalar(p[0]), _p12_scalar(p[1]))
    R = Q
    f = _p12_one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = _p12_mul(_p12_mul(f, f), _linefunc(R, R, P))
        R = _p12_pt_double(R)
        if ATE_LOOP_COUNT & (1 << i):
            f = _p12_mul(f, _linefunc(R, Q, P))
            R = _p12_pt_add(R, Q)
    Q1 = (_p12_pow(Q[0], Q_MOD), _p12_pow(Q[1], Q_MOD))
    nQ2 = (_p12_pow(Q1[0], Q_MOD), [(-v) % Q_MOD for v in _p12_pow(Q1[1], Q_MOD)])
    f = _p12_mul(f, _linefunc(R, Q1, P))
    R = _p12_pt_add(R, Q1)
    f = _p12_mul(f, _linefunc(R, nQ2, P))
    return f


def final_exponentiate(f):
    return _p12_pow(f, (Q_MOD**12 - 1) // R_MOD)


def pairing_product_is_one(pairs: Sequence[Tuple[G1Point, G2Point]]) -> bool:
    f = _p12_one()
    for p, q in pairs:
        f = _p12_mul(f, miller_loop(q, p))
    return final_exponentiate(f) == _p12_one()


@dataclass
class VerifyingKey:
    alpha1: G1Point
    beta2: G2Point
    gamma2: G2Point
    delta2: G2Point
    ic: List[G1Point]


def vk_from_snarkjs_json(obj: dict) -> VerifyingKey:
    def g1(v):
        return (int(v[0]), int(v[1]))

    def g2(v):
        return ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))

    return VerifyingKey(g1(obj["vk_alpha_1"]), g2(obj["vk_beta_2"]), g2(obj["vk_gamma_2"]),
                        g2(obj["vk_delta_2"]), [g1(v) for v in obj["IC"]])


def vk_from_zkey(zk: ZKey) -> VerifyingKey:
    return VerifyingKey(zk.alpha1, zk.beta2, zk.gamma2, zk.delta2, list(zk.ic))


def proof_from_json(s: str):
    o = json.loads(s)

    def g1(v):
        p = (int(v[0]), int(v[1]))
        return None if p == (0, 0) else p

    b = o["pi_b"]
    pb = ((int(b[0][0]), int(b[0][1])), (int(b[1][0]), int(b[1][1])))
    return g1(o["pi_a"]), (None if pb == ((0, 0), (0, 0)) else pb), g1(o["pi_c"])


def groth16_verify(vk: VerifyingKey, public_inputs: Sequence[int], pi_a: G1Point, pi_b: G2Point,
                   pi_c: G1Point) -> bool:
    """e(-A,B) e(alpha,beta) e(vk_x,gamma) e(C,delta) == 1 — what prover-service checks through ark-groth16
    (prover-service/src/request_handler/prover_handler.rs:329-336)."""
    if not (g1_is_on_curve(pi_a) and g2_is_on_curve(pi_b) and g1_is_on_curve(pi_c)):
        return False
    vk_x = vk.ic[0]
    for i, x in enumerate(public_inputs):
        vk_x = g1_add(vk_x, g1_mul(vk.ic[i + 1], x))
    return pairing_product_is_one(
        [(g1_neg(pi_a), pi_b), (vk.alpha1, vk.beta2), (vk_x, vk.gamma2), (pi_c, vk.delta2)]
    )


# --------------------------------------------------------------------------- synthetic circuits + trapdoor setup
@dataclass
class R1CS:
    n_vars: int
    n_public: int
    # rows: list of (A: {wire: coef}, B: {wire: coef}, C: {wire: coef})
    rows: List[Tuple[dict, dict, dict]] = field(default_factory=list)


class _Lcg:
    """Tiny deterministic PRNG shared (bit for bit) with oracle/kzp_port.c so both generate the same circuits."""

    def __init__(self, seed: int):
        self.s = (seed * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        # splitmix64
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def fr(self) -> int:
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        return v % R_MOD


def synth_circuit(n_constraints: int, n_vars: int, seed: int, n_public: int = 1) -> Tuple[R1CS, List[int]]:
    """Keyless-shaped synthetic R1CS with a satisfying witness: ~80 % of the wires
    are bits (XOR / AND gates), ~15 % bytes composed from 8 bits, ~5 % full-width field products; surplus
    constraints are booleanity checks b*(b-1)=0 (C-less rows). Wire 0 = 1, wire 1 = public.
    Bit-for-bit the same generator as circuit_synth() in oracle/kzp_port.c (same PRNG call order)."""
    assert n_public == 1 and n_vars >= 16 and n_constraints >= n_vars - 9
    rng = _Lcg(seed)
    w = [0] * n_vars
    w[0] = 1
    rows: List[Tuple[dict, dict, dict]] = []
    bits: List[int] = []
    fields: List[int] = []
    for i in range(2, 10):
        w[i] = rng.below(2)
        bits.append(i)

    def and_gate(out):
        i = bits[rng.below(len(bits))]
        j = bits[rng.below(len(bits))]
        w[out] = w[i] & w[j]
        rows.append(({i: 1}, {j: 1}, {out: 1}))

    and_gate(1)
    for out in range(10, n_vars):
        t = rng.below(100)
        if t < 50 or len(bits) < 8:
            i = bits[rng.below(len(bits))]
            j = bits[rng.below(len(bits))]
            w[out] = w[i] ^ w[j]
            c = {i: 2} if i == j else {i: 1, j: 1}
            c[out] = R_MOD - 1
            rows.append(({i: 2}, {j: 1}, c))  # 2ij = i + j - out
            bits.append(out)
        elif t < 80:
            and_gate(out)
            bits.append(out)
        elif t < 95:
            a = {}
            val = 0
            for k in range(8):
                i = bits[rng.below(len(bits))]
                a[i] = a.get(i, 0) + (1 << k)
                val += w[i] << k
            w[out] = val
            rows.append((a, {0: 1}, {out: 1}))
        else:
            if len(fields) < 2:
                x = rng.fr()
                w[out] = x
                rows.append(({0: x}, {0: 1}, {out: 1}))
            else:
                i = fields[rng.below(len(fields))]
                j = fields[rng.below(len(fields))]
                k = bits[rng.below(len(bits))]
                w[out] = (w[i] + w[k]) * (w[j] + 3) % R_MOD
                rows.append(({i: 1, k: 1}, {j: 1, 0: 3}, {out: 1}))
            fields.append(out)
    while len(rows) < n_constraints:
        i = bits[rng.below(len(bits))]
        rows.append(({i: 1}, {i: 1, 0: R_MOD - 1}, {}))
    return R1CS(n_vars, n_public, rows), w


def check_r1cs(r1cs: R1CS, w: Sequence[int]) -> bool:
    for A, B, C in r1cs.rows:
        a = sum(w[s] * v for s, v in A.items()) % R_MOD
        b = sum(w[s] * v for s, v in B.items()) % R_MOD
        c = sum(w[s] * v for s, v in C.items()) % R_MOD
        if a * b % R_MOD != c:
            return False
    return True


def trapdoor_setup(r1cs: R1CS, seed: int) -> Tuple[ZKey, dict]:
    """snarkjs-compatible groth16 setup from a known trapdoor."""
    rng = _Lcg(seed ^ 0x5EED)
    tau, alpha, beta, gamma, delta = (rng.fr() or 1 for _ in range(5))
    m = len(r1cs.rows)
    n_pub = r1cs.n_public
    n = 1
    while n < m + n_pub + 1:
        n <<= 1
    log_n = n.bit_length() - 1
    omega = fr_root_of_unity(log_n)
    # L_j(tau) = (tau^n - 1) w^j / (n (tau - w^j))
    tn1 = (pow(tau, n, R_MOD) - 1) % R_MOD
    ninv = pow(n, -1, R_MOD)
    L = []
    wj = 1
    for j in range(m + n_pub + 1):
        L.append(tn1 * wj % R_MOD * ninv % R_MOD * pow((tau - wj) % R_MOD, -1, R_MOD) % R_MOD)
        wj = wj * omega % R_MOD
    nv = r1cs.n_vars
    At = [0] * nv
    Bt = [0] * nv
    Ct = [0] * nv
    coefs = []
    for j, (A, B, C) in enumerate(r1cs.rows):
        for s, v in A.items():
            At[s] = (At[s] + v * L[j]) % R_MOD
            coefs.append((0, j, s, v % R_MOD))
        for s, v in B.items():
            Bt[s] = (Bt[s] + v * L[j]) % R_MOD
            coefs.append((1, j, s, v % R_MOD))
        for s, v in C.items():
            Ct[s] = (Ct[s] + v * L[j]) % R_MOD
    for s in range(n_pub + 1):
        At[s] = (At[s] + L[m + s]) % R_MOD
        coefs.append((0, m + s, s, 1))
    ginv = pow(gamma, -1, R_MOD)
    dinv = pow(delta, -1, R_MOD)
    ic = [g1_mul(G1_GEN, (beta * At[s] + alpha * Bt[s] + Ct[s]) * ginv % R_MOD) for s in range(n_pub + 1)]
    pa = [g1_mul(G1_GEN, At[s]) if At[s] else None for s in range(nv)]
    pb1 = [g1_mul(G1_GEN, Bt[s]) if Bt[s] else None for s in range(nv)]
    pb2 = [g2_mul(G2_GEN, Bt[s]) if Bt[s] else None for s in range(nv)]
    pc = []
    for s in range(n_pub + 1, nv):
        k = (beta * At[s] + alpha * Bt[s] + Ct[s]) * dinv % R_MOD
        pc.append(g1_mul(G1_GEN, k) if k else None)
    # H[i] = L^{(2n)}_{2i+1}(tau) / delta
    w2n = fr_root_of_unity(log_n + 1)
    t2n1 = (pow(tau, 2 * n, R_MOD) - 1) % R_MOD
    inv2n = pow(2 * n, -1, R_MOD)
    ph = []
    wk = w2n
    w2n_sq = w2n * w2n % R_MOD
    for i in range(n):
        k = t2n1 * wk % R_MOD * inv2n % R_MOD * pow((tau - wk) % R_MOD, -1, R_MOD) % R_MOD * dinv % R_MOD
        ph.append(g1_mul(G1_GEN, k) if k else None)
        wk = wk * w2n_sq % R_MOD
    zk = ZKey(
        nv, n_pub, n, g1_mul(G1_GEN, alpha), g1_mul(G1_GEN, beta), g2_mul(G2_GEN, beta),
        g2_mul(G2_GEN, gamma), g1_mul(G1_GEN, delta), g2_mul(G2_GEN, delta), ic, coefs, pa, pb1, pb2, pc, ph,
    )
    trap = dict(tau=tau, alpha=alpha, beta=beta, gamma=gamma, delta=delta, At=At, Bt=Bt, Ct=Ct)
    return zk, trap


def trapdoor_expected_proof(zk: ZKey, trap: dict, w: Sequence[int], r: int, s: int):
    """Expected (pi_a, pi_b, pi_c) computed in Fr scalars from the trapdoor (last bullet)."""
    al, be, de = trap["alpha"], trap["beta"], trap["delta"]
    At, Bt, Ct = trap["At"], trap["Bt"], trap["Ct"]
    nv = zk.n_vars
    a = sum(w[i] * At[i] for i in range(nv)) % R_MOD
    b = sum(w[i] * Bt[i] for i in range(nv)) % R_MOD
    c = sum(w[i] * Ct[i] for i in range(nv)) % R_MOD
    dinv = pow(de, -1, R_MOD)
    A = (al + a + r * de) % R_MOD
    B = (be + b + s * de) % R_MOD
    priv = sum(w[i] * ((be * At[i] + al * Bt[i] + Ct[i]) % R_MOD) for i in range(zk.n_public + 1, nv)) % R_MOD
    C = (priv * dinv + (a * b - c) * dinv + s * A + r * B - r * s % R_MOD * de) % R_MOD
    return g1_mul(G1_GEN, A), g2_mul(G2_GEN, B), g1_mul(G1_GEN, C)
