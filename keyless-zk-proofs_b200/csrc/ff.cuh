// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Byte layout of an element is identical to the reference's FrRawElement / FqRawElement
// (4 x u64 little-endian limbs, rust-rapidsnark/rapidsnark/src/fr_element.hpp:6-13), so zkey
// sections and witness values are consumed without conversion. Semantics follow the reference
// raw API (fr_raw_generic.cpp:11-39,68-80,107-148,192-232): every result is canonical (< p).
//
// Device path: the Montgomery product is a word-serial (CIOS) loop whose multiply-accumulates are
// written as mad.lo.cc / madc.hi.cc carry chains over an even-aligned and an odd-aligned
// accumulator, so ptxas can pair each lo/hi couple into one IMAD.WIDE.U32 with carry; 16 wide
// multiplies for a*b_i, 16 for m*p and one IMAD for m per limb of b (136 per product).
// Host path (same templates, used by the host-side proof assembly and by tests): portable
// 64-bit C++.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define KZP_HD __host__ __device__ __forceinline__
#define KZP_D __device__ __forceinline__
#else
#define KZP_HD inline
#define KZP_D inline
#endif

namespace kzp
{

// ------------------------------------------------------------------ parameters
// Constants: fr_raw_generic.cpp:5-7 / fq_raw_generic.cpp:6-8 (q, R^2, -p^-1), re-split in 32-bit limbs.
struct FrParams
{
    static constexpr uint32_t P0 = 0xf0000001u, P1 = 0x43e1f593u, P2 = 0x79b97091u, P3 = 0x2833e848u,
                              P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
    static constexpr uint32_t NP0 = 0xefffffffu; // -p^-1 mod 2^32
    static constexpr uint32_t R1[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                       0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    static constexpr uint32_t R2[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                       0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
};

struct FqParams
{
    static constexpr uint32_t P0 = 0xd87cfd47u, P1 = 0x3c208c16u, P2 = 0x6871ca8du, P3 = 0x97816a91u,
                              P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
    static constexpr uint32_t NP0 = 0xe4866389u;
    static constexpr uint32_t R1[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                       0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    static constexpr uint32_t R2[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                       0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
};

template <class P>
KZP_HD uint32_t modulus_limb(int i)
{
    switch (i)
    {
    case 0: return P::P0;
    case 1: return P::P1;
    case 2: return P::P2;
    case 3: return P::P3;
    case 4: return P::P4;
    case 5: return P::P5;
    case 6: return P::P6;
    default: return P::P7;
    }
}

// ------------------------------------------------------------------ element
template <class P>
struct alignas(16) Fp
{
    uint32_t v[8];

    typedef P Params;
    static constexpr bool kFusedMulAdd2 = true; // mul_add2 is a single-reduction dual product (used by the group law)

    static KZP_HD Fp zero()
    {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++)
            r.v[i] = 0;
        return r;
    }

    // Montgomery representation of 1 (R mod p)
    static KZP_HD Fp one()
    {
        Fp r;
        r.v[0] = P::R1[0]; r.v[1] = P::R1[1]; r.v[2] = P::R1[2]; r.v[3] = P::R1[3];
        r.v[4] = P::R1[4]; r.v[5] = P::R1[5]; r.v[6] = P::R1[6]; r.v[7] = P::R1[7];
        return r;
    }

    static KZP_HD Fp r2()
    {
        Fp r;
        r.v[0] = P::R2[0]; r.v[1] = P::R2[1]; r.v[2] = P::R2[2]; r.v[3] = P::R2[3];
        r.v[4] = P::R2[4]; r.v[5] = P::R2[5]; r.v[6] = P::R2[6]; r.v[7] = P::R2[7];
        return r;
    }

    static KZP_HD bool is_zero(const Fp& a)
    {
        uint32_t t = a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7];
        return t == 0;
    }

    static KZP_HD bool eq(const Fp& a, const Fp& b)
    {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
            t |= a.v[i] ^ b.v[i];
        return t == 0;
    }

    // a >= p ?
    static KZP_HD bool geq_p(const Fp& a)
    {
#pragma unroll
        for (int i = 7; i >= 0; i--)
        {
            uint32_t pi = modulus_limb<P>(i);
            if (a.v[i] > pi)
                return true;
            if (a.v[i] < pi)
                return false;
        }
        return true;
    }

    // ---------------------------------------------------------------- add / sub / neg
    // r = a + b mod p (Fr_rawAdd: fr_raw_generic.cpp:11-22). Inputs canonical -> no 2^256 overflow.
    static KZP_HD void add(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7, t0, t1, t2, t3, t4, t5, t6, t7, bw;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7)
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
              "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
              "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7),
              "=r"(bw)
            : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(s7), "r"(P::P0),
              "r"(P::P1), "r"(P::P2), "r"(P::P3), "r"(P::P4), "r"(P::P5), "r"(P::P6), "r"(P::P7));
        bool keep = (bw != 0); // borrow -> sum < p -> keep the plain sum
        r.v[0] = keep ? s0 : t0; r.v[1] = keep ? s1 : t1; r.v[2] = keep ? s2 : t2;
        r.v[3] = keep ? s3 : t3; r.v[4] = keep ? s4 : t4; r.v[5] = keep ? s5 : t5;
        r.v[6] = keep ? s6 : t6; r.v[7] = keep ? s7 : t7;
#else
        uint32_t s[8], t[8];
        uint64_t c = 0;
        for (int i = 0; i < 8; i++)
        {
            c += (uint64_t)a.v[i] + b.v[i];
            s[i] = (uint32_t)c;
            c >>= 32;
        }
        int64_t bw = 0;
        for (int i = 0; i < 8; i++)
        {
            bw += (int64_t)s[i] - (int64_t)modulus_limb<P>(i);
            t[i] = (uint32_t)bw;
            bw >>= 32;
        }
        for (int i = 0; i < 8; i++)
            r.v[i] = bw ? s[i] : t[i];
#endif
    }

    // r = a - b mod p (Fr_rawSub: fr_raw_generic.cpp:24-32)
    static KZP_HD void sub(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7, bw;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7),
              "=r"(bw)
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
              "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
              "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        // bw = 0xffffffff when a < b: add p back
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7)
            : "r"(P::P0 & bw), "r"(P::P1 & bw), "r"(P::P2 & bw), "r"(P::P3 & bw), "r"(P::P4 & bw),
              "r"(P::P5 & bw), "r"(P::P6 & bw), "r"(P::P7 & bw));
        r.v[0] = s0; r.v[1] = s1; r.v[2] = s2; r.v[3] = s3;
        r.v[4] = s4; r.v[5] = s5; r.v[6] = s6; r.v[7] = s7;
#else
        uint32_t s[8];
        int64_t  bw = 0;
        for (int i = 0; i < 8; i++)
        {
            bw += (int64_t)a.v[i] - (int64_t)b.v[i];
            s[i] = (uint32_t)bw;
            bw >>= 32;
        }
        uint32_t mask = bw ? 0xffffffffu : 0u;
        uint64_t c    = 0;
        for (int i = 0; i < 8; i++)
        {
            c += (uint64_t)s[i] + (modulus_limb<P>(i) & mask);
            r.v[i] = (uint32_t)c;
            c >>= 32;
        }
#endif
    }

    // r = -a mod p, with -0 = 0 (Fr_rawNeg: fr_raw_generic.cpp:34-39)
    static KZP_HD void neg(Fp& r, const Fp& a)
    {
        if (is_zero(a))
        {
            r = a;
            return;
        }
        Fp z = zero();
        sub(r, z, a);
    }

    static KZP_HD void dbl(Fp& r, const Fp& a) { add(r, a, a); }

    // ---------------------------------------------------------------- Montgomery product
#if defined(__CUDA_ARCH__)
    // x0 += pend (carry into limb 1); y[k], k=0..7 (limb offsets 1..8) += {m1,m3,m5,m7} * b
    static KZP_D void chain_odd_pend(uint32_t& x0, uint32_t pend, uint32_t (&y)[8], uint32_t m1,
                                     uint32_t m3, uint32_t m5, uint32_t m7, uint32_t b)
    {
        asm("add.cc.u32 %0, %0, %9;\n\t"
            "madc.lo.cc.u32 %1, %10, %14, %1;\n\t"
            "madc.hi.cc.u32 %2, %10, %14, %2;\n\t"
            "madc.lo.cc.u32 %3, %11, %14, %3;\n\t"
            "madc.hi.cc.u32 %4, %11, %14, %4;\n\t"
            "madc.lo.cc.u32 %5, %12, %14, %5;\n\t"
            "madc.hi.cc.u32 %6, %12, %14, %6;\n\t"
            "madc.lo.cc.u32 %7, %13, %14, %7;\n\t"
            "madc.hi.u32 %8, %13, %14, %8;"
            : "+r"(x0), "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]),
              "+r"(y[6]), "+r"(y[7])
            : "r"(pend), "r"(m1), "r"(m3), "r"(m5), "r"(m7), "r"(b));
    }

    // y[k] += {m1,m3,m5,m7} * b (no carry-in)
    static KZP_D void chain_odd(uint32_t (&y)[8], uint32_t m1, uint32_t m3, uint32_t m5,
                                uint32_t m7, uint32_t b)
    {
        asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
            "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
            "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
            "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
            "madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]),
              "+r"(y[7])
            : "r"(m1), "r"(m3), "r"(m5), "r"(m7), "r"(b));
    }

    // x[k], k=0..7 (limb offsets 0..7) += {m0,m2,m4,m6} * b ; carry out of limb 7 goes to ytop (limb 8)
    static KZP_D void chain_even(uint32_t (&x)[8], uint32_t& ytop, uint32_t m0, uint32_t m2,
                                 uint32_t m4, uint32_t m6, uint32_t b)
    {
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]),
              "+r"(x[7]), "+r"(ytop)
            : "r"(m0), "r"(m2), "r"(m4), "r"(m6), "r"(b));
    }
#endif

#if defined(__CUDACC__)
    // (bodies exist in the device pass only: nvcc's host pass just needs the declarations)
    // s = a * b * R^-1 + (a multiple of p), NOT reduced: s < a b / R + p. For canonical a, b that is below 2p; more
    // generally below 2p whenever a b < 4.2 p^2 (p / R = 0.19 for both BN254 fields), e.g. a < 4p and b < p.
    static KZP_D void mul_core(uint32_t (&s)[8], const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        // T = X + (Y << 32) + pend, X at limb offsets 0..7, Y at 1..8; T < 2^288 throughout.
        uint32_t x[8], y[8], pend = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            x[i] = 0;
            y[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            uint32_t bi = b.v[i];
            chain_odd_pend(x[0], pend, y, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            chain_even(x, y[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
            uint32_t m = x[0] * P::NP0;
            chain_even(x, y[7], P::P0, P::P2, P::P4, P::P6, m);
            chain_odd(y, P::P1, P::P3, P::P5, P::P7, m);
            pend = x[1];
            uint32_t t0 = y[0], t1 = y[1], t2 = y[2], t3 = y[3], t4 = y[4], t5 = y[5], t6 = y[6],
                     t7 = y[7];
            y[0] = x[2]; y[1] = x[3]; y[2] = x[4]; y[3] = x[5]; y[4] = x[6]; y[5] = x[7];
            y[6] = 0; y[7] = 0;
            x[0] = t0; x[1] = t1; x[2] = t2; x[3] = t3; x[4] = t4; x[5] = t5; x[6] = t6; x[7] = t7;
        }
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
            : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]),
              "r"(x[7]), "r"(pend), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]),
              "r"(y[5]), "r"(y[6]));
#endif
    }

    // ---- lazy-reduction variants for chains of butterflies (values kept in [0, 2p): BN254's moduli leave two spare
    //      bits in 256, so sums of two such values and a - b + 2p still fit). Used by the NTT levels only; everything
    //      that leaves a kernel is canonical again (reduce_2p).
    static constexpr uint32_t twice_p(int i)
    {
        const uint32_t m[8] = {P::P0, P::P1, P::P2, P::P3, P::P4, P::P5, P::P6, P::P7};
        return (m[i] << 1) | (i ? (m[i - 1] >> 31) : 0u);
    }
    // r = a b R^-1 mod p as a value below 2p; a b < 4.2 p^2 (see mul_core)
    static KZP_D void mul_lazy(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s[8];
        mul_core(s, a, b);
#pragma unroll
        for (int i = 0; i < 8; i++)
            r.v[i] = s[i];
#endif
    }
    // r = a - k (256-bit) if that does not borrow, else a: one conditional subtraction of the constant k
    template <bool TWICE>
    static KZP_D void cond_sub_const(Fp& r, const uint32_t (&s)[8])
    {
#if defined(__CUDA_ARCH__)
        constexpr uint32_t k0 = TWICE ? twice_p(0) : P::P0, k1 = TWICE ? twice_p(1) : P::P1, k2 = TWICE ? twice_p(2) : P::P2,
                           k3 = TWICE ? twice_p(3) : P::P3, k4 = TWICE ? twice_p(4) : P::P4, k5 = TWICE ? twice_p(5) : P::P5,
                           k6 = TWICE ? twice_p(6) : P::P6, k7 = TWICE ? twice_p(7) : P::P7;
        uint32_t t0, t1, t2, t3, t4, t5, t6, t7, bw;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7), "=r"(bw)
            : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(k0), "r"(k1), "r"(k2),
              "r"(k3), "r"(k4), "r"(k5), "r"(k6), "r"(k7));
        bool keep = (bw != 0);
        r.v[0] = keep ? s[0] : t0; r.v[1] = keep ? s[1] : t1; r.v[2] = keep ? s[2] : t2; r.v[3] = keep ? s[3] : t3;
        r.v[4] = keep ? s[4] : t4; r.v[5] = keep ? s[5] : t5; r.v[6] = keep ? s[6] : t6; r.v[7] = keep ? s[7] : t7;
#endif
    }
    // [0, 2p) -> canonical
    static KZP_D void reduce_2p(Fp& r, const Fp& a)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            s[i] = a.v[i];
        cond_sub_const<false>(r, s);
#endif
    }
    // a, b in [0, 2p) -> a + b mod 2p-window: a value in [0, 2p) congruent to a + b
    static KZP_D void add_lazy(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s[8];
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
              "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        cond_sub_const<true>(r, s);
#endif
    }
    // a, b in [0, 2p) -> a value in [0, 2p) congruent to a - b (2p added back on borrow)
    static KZP_D void sub_lazy(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7, bw;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7), "=r"(bw)
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
              "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7)
            : "r"(twice_p(0) & bw), "r"(twice_p(1) & bw), "r"(twice_p(2) & bw), "r"(twice_p(3) & bw), "r"(twice_p(4) & bw),
              "r"(twice_p(5) & bw), "r"(twice_p(6) & bw), "r"(twice_p(7) & bw));
        r.v[0] = s0; r.v[1] = s1; r.v[2] = s2; r.v[3] = s3; r.v[4] = s4; r.v[5] = s5; r.v[6] = s6; r.v[7] = s7;
#endif
    }
    // a, b in [0, 2p) -> a + 2p - b, a value in (0, 4p): no comparison at all; feed it to mul_lazy with a canonical
    // second factor (4p * p < 4.2 p^2)
    static KZP_D void sub_plus_2p(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7)
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
              "r"(twice_p(0)), "r"(twice_p(1)), "r"(twice_p(2)), "r"(twice_p(3)), "r"(twice_p(4)), "r"(twice_p(5)),
              "r"(twice_p(6)), "r"(twice_p(7)));
        asm("sub.cc.u32 %0, %0, %8;\n\t"
            "subc.cc.u32 %1, %1, %9;\n\t"
            "subc.cc.u32 %2, %2, %10;\n\t"
            "subc.cc.u32 %3, %3, %11;\n\t"
            "subc.cc.u32 %4, %4, %12;\n\t"
            "subc.cc.u32 %5, %5, %13;\n\t"
            "subc.cc.u32 %6, %6, %14;\n\t"
            "subc.u32 %7, %7, %15;"
            : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3), "+r"(s4), "+r"(s5), "+r"(s6), "+r"(s7)
            : "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        r.v[0] = s0; r.v[1] = s1; r.v[2] = s2; r.v[3] = s3; r.v[4] = s4; r.v[5] = s5; r.v[6] = s6; r.v[7] = s7;
#endif
    }
#endif

    // r = a * b * R^-1 mod p, canonical (Fr_rawMMul: fr_raw_generic.cpp:107-148)
    static KZP_HD void mul(Fp& r, const Fp& a, const Fp& b)
    {
#if defined(__CUDA_ARCH__)
        // T = X + (Y << 32) + pend, X at limb offsets 0..7, Y at 1..8; T < 2^288 throughout.
        uint32_t x[8], y[8], pend = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            x[i] = 0;
            y[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            uint32_t bi = b.v[i];
            chain_odd_pend(x[0], pend, y, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            chain_even(x, y[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
            uint32_t m = x[0] * P::NP0;
            chain_even(x, y[7], P::P0, P::P2, P::P4, P::P6, m);
            chain_odd(y, P::P1, P::P3, P::P5, P::P7, m);
            // divide by 2^32: x[0] == 0 now; x[1] lands on limb 0 (kept pending), the odd-aligned
            // array becomes the even-aligned one and vice versa.
            pend = x[1];
            uint32_t t0 = y[0], t1 = y[1], t2 = y[2], t3 = y[3], t4 = y[4], t5 = y[5], t6 = y[6],
                     t7 = y[7];
            y[0] = x[2]; y[1] = x[3]; y[2] = x[4]; y[3] = x[5]; y[4] = x[6]; y[5] = x[7];
            y[6] = 0; y[7] = 0;
            x[0] = t0; x[1] = t1; x[2] = t2; x[3] = t3; x[4] = t4; x[5] = t5; x[6] = t6; x[7] = t7;
        }
        // merge: T = pend + X + (Y << 32)  (T < 2p < 2^255, so limb 8 is zero)
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7)
            : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]),
              "r"(x[7]), "r"(pend), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]),
              "r"(y[5]), "r"(y[6]));
        uint32_t t0, t1, t2, t3, t4, t5, t6, t7, bw;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7),
              "=r"(bw)
            : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(s7), "r"(P::P0),
              "r"(P::P1), "r"(P::P2), "r"(P::P3), "r"(P::P4), "r"(P::P5), "r"(P::P6), "r"(P::P7));
        bool keep = (bw != 0);
        r.v[0] = keep ? s0 : t0; r.v[1] = keep ? s1 : t1; r.v[2] = keep ? s2 : t2;
        r.v[3] = keep ? s3 : t3; r.v[4] = keep ? s4 : t4; r.v[5] = keep ? s5 : t5;
        r.v[6] = keep ? s6 : t6; r.v[7] = keep ? s7 : t7;
#else
        // portable CIOS
        uint32_t t[10];
        for (int i = 0; i < 10; i++)
            t[i] = 0;
        for (int i = 0; i < 8; i++)
        {
            uint64_t c = 0;
            for (int j = 0; j < 8; j++)
            {
                c += (uint64_t)a.v[j] * b.v[i] + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[8] = (uint32_t)c;
            t[9] = (uint32_t)(c >> 32);
            uint32_t m = t[0] * P::NP0;
            c          = ((uint64_t)m * modulus_limb<P>(0) + t[0]) >> 32;
            for (int j = 1; j < 8; j++)
            {
                c += (uint64_t)m * modulus_limb<P>(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[7] = (uint32_t)c;
            t[8] = t[9] + (uint32_t)(c >> 32);
        }
        Fp s;
        for (int i = 0; i < 8; i++)
            s.v[i] = t[i];
        if (t[8] || geq_p(s))
        {
            int64_t bw = 0;
            for (int i = 0; i < 8; i++)
            {
                bw += (int64_t)s.v[i] - (int64_t)modulus_limb<P>(i);
                s.v[i] = (uint32_t)bw;
                bw >>= 32;
            }
        }
        r = s;
#endif
    }

#if defined(__CUDA_ARCH__)
    // acc[0..2N) += {x0..x(N-1)} * b as one carry chain of N wide multiply-adds; returns the carry out of acc[2N-1]
    static KZP_D uint32_t row1(uint32_t& a0, uint32_t& a1, uint32_t x0, uint32_t b)
    {
        uint32_t c;
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, 0, 0;"
            : "+r"(a0), "+r"(a1), "=r"(c)
            : "r"(x0), "r"(b));
        return c;
    }
    static KZP_D uint32_t row2(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3, uint32_t x0, uint32_t x1,
                               uint32_t b)
    {
        uint32_t c;
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, 0, 0;"
            : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "=r"(c)
            : "r"(x0), "r"(x1), "r"(b));
        return c;
    }
    static KZP_D uint32_t row3(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3, uint32_t& a4, uint32_t& a5,
                               uint32_t x0, uint32_t x1, uint32_t x2, uint32_t b)
    {
        uint32_t c;
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, 0, 0;"
            : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "=r"(c)
            : "r"(x0), "r"(x1), "r"(x2), "r"(b));
        return c;
    }
    static KZP_D void wide(uint32_t& lo, uint32_t& hi, uint32_t x, uint32_t y)
    {
        uint64_t p = (uint64_t)x * y;
        lo         = (uint32_t)p;
        hi         = (uint32_t)(p >> 32);
    }

    // One Montgomery reduction round on T = X + (Y << 32) + pend: adds m * p with m chosen so that limb 0 cancels,
    // divides by 2^32 and shifts `inject` in at limb 7 of the quotient (the next limb of a 512-bit operand, or 0).
    static KZP_D void redc_round(uint32_t (&x)[8], uint32_t (&y)[8], uint32_t& pend, uint32_t inject)
    {
        uint32_t m = (x[0] + pend) * P::NP0;
        chain_odd_pend(x[0], pend, y, P::P1, P::P3, P::P5, P::P7, m);
        chain_even(x, y[7], P::P0, P::P2, P::P4, P::P6, m);
        pend = x[1];
        uint32_t t0 = y[0], t1 = y[1], t2 = y[2], t3 = y[3], t4 = y[4], t5 = y[5], t6 = y[6], t7 = y[7];
        y[0] = x[2]; y[1] = x[3]; y[2] = x[4]; y[3] = x[5]; y[4] = x[6]; y[5] = x[7];
        y[6] = inject; y[7] = 0;
        x[0] = t0; x[1] = t1; x[2] = t2; x[3] = t3; x[4] = t4; x[5] = t5; x[6] = t6; x[7] = t7;
    }

    // r = pend + X + (Y << 32), brought below p by one conditional subtraction (the value is < 2p)
    static KZP_D void merge_reduce(Fp& r, const uint32_t (&x)[8], const uint32_t (&y)[8], uint32_t pend)
    {
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7;
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7)
            : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]),
              "r"(x[7]), "r"(pend), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]),
              "r"(y[5]), "r"(y[6]));
        uint32_t t0, t1, t2, t3, t4, t5, t6, t7, bw;
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7),
              "=r"(bw)
            : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(s7), "r"(P::P0),
              "r"(P::P1), "r"(P::P2), "r"(P::P3), "r"(P::P4), "r"(P::P5), "r"(P::P6), "r"(P::P7));
        bool keep = (bw != 0);
        r.v[0] = keep ? s0 : t0; r.v[1] = keep ? s1 : t1; r.v[2] = keep ? s2 : t2;
        r.v[3] = keep ? s3 : t3; r.v[4] = keep ? s4 : t4; r.v[5] = keep ? s5 : t5;
        r.v[6] = keep ? s6 : t6; r.v[7] = keep ? s7 : t7;
    }
#endif

    // r = a^2 * R^-1 mod p, canonical (Fr_rawMSquare: fr_raw_generic.cpp:150-190 computes the same value).
    // Device: the 28 off-diagonal products once (even- and odd-aligned accumulators as in mul), doubled, plus the
    // 8 diagonal squares, then 8 reduction rounds: 28 + 8 + 64 = 100 wide multiply-adds instead of 128.
    static KZP_HD void sqr(Fp& r, const Fp& a)
    {
#if defined(__CUDA_ARCH__)
        const uint32_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4], a5 = a.v[5], a6 = a.v[6],
                       a7 = a.v[7];
        // O[k], E[k]: limb k of the odd- / even-aligned sums of a_i a_j, i < j (product i,j starts at limb i + j)
        uint32_t O1, O2, O3, O4, O5, O6, O7, O8, O9, O10 = 0, O11, O12 = 0, O13, O14 = 0;
        uint32_t E2, E3, E4, E5, E6, E7, E8 = 0, E9 = 0, E10, E11 = 0, E12, E13 = 0;
        wide(O1, O2, a1, a0); wide(O3, O4, a3, a0); wide(O5, O6, a5, a0); wide(O7, O8, a7, a0);
        wide(E2, E3, a2, a0); wide(E4, E5, a4, a0); wide(E6, E7, a6, a0);
        O9 = row3(O3, O4, O5, O6, O7, O8, a2, a4, a6, a1);
        (void)row3(E4, E5, E6, E7, E8, E9, a3, a5, a7, a1);
        (void)row3(O5, O6, O7, O8, O9, O10, a3, a5, a7, a2);
        E10 = row2(E6, E7, E8, E9, a4, a6, a2);
        O11 = row2(O7, O8, O9, O10, a4, a6, a3);
        (void)row2(E8, E9, E10, E11, a5, a7, a3);
        (void)row2(O9, O10, O11, O12, a5, a7, a4);
        E12 = row1(E10, E11, a6, a4);
        O13 = row1(O11, O12, a6, a5);
        (void)row1(E12, E13, a7, a5);
        (void)row1(O13, O14, a7, a6);
        // M = E + O (limbs 1..14), D = 2 M (limbs 1..15)
        uint32_t M2, M3, M4, M5, M6, M7, M8, M9, M10, M11, M12, M13, M14;
        asm("add.cc.u32 %0, %13, %26;\n\t"
            "addc.cc.u32 %1, %14, %27;\n\t"
            "addc.cc.u32 %2, %15, %28;\n\t"
            "addc.cc.u32 %3, %16, %29;\n\t"
            "addc.cc.u32 %4, %17, %30;\n\t"
            "addc.cc.u32 %5, %18, %31;\n\t"
            "addc.cc.u32 %6, %19, %32;\n\t"
            "addc.cc.u32 %7, %20, %33;\n\t"
            "addc.cc.u32 %8, %21, %34;\n\t"
            "addc.cc.u32 %9, %22, %35;\n\t"
            "addc.cc.u32 %10, %23, %36;\n\t"
            "addc.cc.u32 %11, %24, %37;\n\t"
            "addc.u32 %12, %25, 0;"
            : "=r"(M2), "=r"(M3), "=r"(M4), "=r"(M5), "=r"(M6), "=r"(M7), "=r"(M8), "=r"(M9), "=r"(M10), "=r"(M11),
              "=r"(M12), "=r"(M13), "=r"(M14)
            : "r"(O2), "r"(O3), "r"(O4), "r"(O5), "r"(O6), "r"(O7), "r"(O8), "r"(O9), "r"(O10), "r"(O11), "r"(O12),
              "r"(O13), "r"(O14), "r"(E2), "r"(E3), "r"(E4), "r"(E5), "r"(E6), "r"(E7), "r"(E8), "r"(E9), "r"(E10),
              "r"(E11), "r"(E12), "r"(E13));
        uint32_t T[16];
        T[0]  = 0;
        T[1]  = O1 << 1;
        T[2]  = __funnelshift_l(O1, M2, 1);
        T[3]  = __funnelshift_l(M2, M3, 1);
        T[4]  = __funnelshift_l(M3, M4, 1);
        T[5]  = __funnelshift_l(M4, M5, 1);
        T[6]  = __funnelshift_l(M5, M6, 1);
        T[7]  = __funnelshift_l(M6, M7, 1);
        T[8]  = __funnelshift_l(M7, M8, 1);
        T[9]  = __funnelshift_l(M8, M9, 1);
        T[10] = __funnelshift_l(M9, M10, 1);
        T[11] = __funnelshift_l(M10, M11, 1);
        T[12] = __funnelshift_l(M11, M12, 1);
        T[13] = __funnelshift_l(M12, M13, 1);
        T[14] = __funnelshift_l(M13, M14, 1);
        T[15] = M14 >> 31;
        // T += sum a_i^2 2^(64 i): one chain of 8 wide multiply-adds over all 16 limbs
        asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
            "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
            "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
            "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
            "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
            "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
            "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
            "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
            "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
            "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
            "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
            "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
            "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
            "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
            "madc.hi.u32 %15, %23, %23, %15;"
            : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]),
              "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7));
        // Montgomery reduction of the 512-bit square, the upper limbs entering one per round
        uint32_t x[8], y[8], pend = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            x[i] = T[i];
            y[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
            redc_round(x, y, pend, T[8 + i]);
        // after the last round the injected limb sits at y[6] (limb 7 of the quotient): that is T[15], in place
        merge_reduce(r, x, y, pend);
#else
        mul(r, a, a);
#endif
    }

    // r = (a*b + c*d) * R^-1 mod p, canonical, with ONE reduction: 128 product + 64 reduction multiply-adds
    // instead of 256 for two products and an addition. T stays < 3p*2^32 < 2^288 through the rounds and the
    // result is < (2p^2 + Rp)/R < 2p.
    static KZP_HD void mul_add2(Fp& r, const Fp& a, const Fp& b, const Fp& c, const Fp& d)
    {
#if defined(__CUDA_ARCH__)
        uint32_t x[8], y[8], pend = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            x[i] = 0;
            y[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            uint32_t bi = b.v[i], di = d.v[i];
            chain_odd_pend(x[0], pend, y, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            chain_even(x, y[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
            chain_odd(y, c.v[1], c.v[3], c.v[5], c.v[7], di);
            chain_even(x, y[7], c.v[0], c.v[2], c.v[4], c.v[6], di);
            pend = 0; // folded into x[0] by chain_odd_pend
            redc_round(x, y, pend, 0);
        }
        merge_reduce(r, x, y, pend);
#else
        Fp t, u;
        mul(t, a, b);
        mul(u, c, d);
        add(r, t, u);
#endif
    }

#if defined(__CUDA_ARCH__)
    // ---- unreduced 512-bit products for the quadratic extension (lazy reduction: an Fq2 product is three wide
    //      products and TWO reductions, 3 x 64 + 2 x 72 multiply-adds instead of 3 x 136)
    // limb k (0..15) of p^2
    static constexpr uint32_t psq_limb(int k)
    {
        const uint32_t m[8]    = {P::P0, P::P1, P::P2, P::P3, P::P4, P::P5, P::P6, P::P7};
        uint32_t       out[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 8; i++)
        {
            uint64_t carry = 0;
            for (int j = 0; j < 8; j++)
            {
                uint64_t t = (uint64_t)m[i] * m[j] + out[i + j] + carry;
                out[i + j] = (uint32_t)t;
                carry      = t >> 32;
            }
            out[i + 8] = (uint32_t)carry;
        }
        return out[k];
    }
    // r = a + b as a plain 256-bit integer (operands below 2^255: no carry out)
    static KZP_D void add_raw(Fp& r, const Fp& a, const Fp& b)
    {
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
              "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    }
    // T (16 limbs) = a * b, the plain integer product (a, b < 2^256). Same even/odd accumulator walk as mul(), without
    // the m * p rows: limb 0 of the running sum is final after each row and leaves through T[i]. 64 multiply-adds.
    static KZP_D void mul_wide(uint32_t (&T)[16], const Fp& a, const Fp& b)
    {
        uint32_t x[8], y[8], pend = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            x[i] = 0;
            y[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            uint32_t bi = b.v[i];
            chain_odd_pend(x[0], pend, y, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            chain_even(x, y[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
            T[i] = x[0];
            pend = x[1];
            uint32_t t0 = y[0], t1 = y[1], t2 = y[2], t3 = y[3], t4 = y[4], t5 = y[5], t6 = y[6], t7 = y[7];
            y[0] = x[2]; y[1] = x[3]; y[2] = x[4]; y[3] = x[5]; y[4] = x[6]; y[5] = x[7];
            y[6] = 0; y[7] = 0;
            x[0] = t0; x[1] = t1; x[2] = t2; x[3] = t3; x[4] = t4; x[5] = t5; x[6] = t6; x[7] = t7;
        }
        // upper half = pend + X + (Y << 32) (the product is below 2^512: nothing beyond limb 7 of this sum)
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
            : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(pend), "r"(y[0]),
              "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]));
    }
    // T += U, T -= U (512-bit, no carry / borrow out by the callers' bounds)
    static KZP_D void wide_add(uint32_t (&T)[16], const uint32_t (&U)[16])
    {
        asm("add.cc.u32 %0, %0, %16;\n\t"
            "addc.cc.u32 %1, %1, %17;\n\t"
            "addc.cc.u32 %2, %2, %18;\n\t"
            "addc.cc.u32 %3, %3, %19;\n\t"
            "addc.cc.u32 %4, %4, %20;\n\t"
            "addc.cc.u32 %5, %5, %21;\n\t"
            "addc.cc.u32 %6, %6, %22;\n\t"
            "addc.cc.u32 %7, %7, %23;\n\t"
            "addc.cc.u32 %8, %8, %24;\n\t"
            "addc.cc.u32 %9, %9, %25;\n\t"
            "addc.cc.u32 %10, %10, %26;\n\t"
            "addc.cc.u32 %11, %11, %27;\n\t"
            "addc.cc.u32 %12, %12, %28;\n\t"
            "addc.cc.u32 %13, %13, %29;\n\t"
            "addc.cc.u32 %14, %14, %30;\n\t"
            "addc.u32 %15, %15, %31;"
            : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]),
              "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
            : "r"(U[0]), "r"(U[1]), "r"(U[2]), "r"(U[3]), "r"(U[4]), "r"(U[5]), "r"(U[6]), "r"(U[7]), "r"(U[8]), "r"(U[9]),
              "r"(U[10]), "r"(U[11]), "r"(U[12]), "r"(U[13]), "r"(U[14]), "r"(U[15]));
    }
    static KZP_D void wide_sub(uint32_t (&T)[16], const uint32_t (&U)[16])
    {
        asm("sub.cc.u32 %0, %0, %16;\n\t"
            "subc.cc.u32 %1, %1, %17;\n\t"
            "subc.cc.u32 %2, %2, %18;\n\t"
            "subc.cc.u32 %3, %3, %19;\n\t"
            "subc.cc.u32 %4, %4, %20;\n\t"
            "subc.cc.u32 %5, %5, %21;\n\t"
            "subc.cc.u32 %6, %6, %22;\n\t"
            "subc.cc.u32 %7, %7, %23;\n\t"
            "subc.cc.u32 %8, %8, %24;\n\t"
            "subc.cc.u32 %9, %9, %25;\n\t"
            "subc.cc.u32 %10, %10, %26;\n\t"
            "subc.cc.u32 %11, %11, %27;\n\t"
            "subc.cc.u32 %12, %12, %28;\n\t"
            "subc.cc.u32 %13, %13, %29;\n\t"
            "subc.cc.u32 %14, %14, %30;\n\t"
            "subc.u32 %15, %15, %31;"
            : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(T[8]),
              "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
            : "r"(U[0]), "r"(U[1]), "r"(U[2]), "r"(U[3]), "r"(U[4]), "r"(U[5]), "r"(U[6]), "r"(U[7]), "r"(U[8]), "r"(U[9]),
              "r"(U[10]), "r"(U[11]), "r"(U[12]), "r"(U[13]), "r"(U[14]), "r"(U[15]));
    }
    // T += K p^2 (K = 1 or 2): keeps a difference of products non-negative without changing its residue
    template <int K>
    static KZP_D void wide_add_psq(uint32_t (&T)[16])
    {
        uint32_t U[16];
#pragma unroll
        for (int i = 0; i < 16; i++)
            U[i] = K == 1 ? psq_limb(i) : ((psq_limb(i) << 1) | (i ? (psq_limb(i - 1) >> 31) : 0u));
        wide_add(T, U);
    }
    // r = T R^-1 mod p, canonical, for T < 4 p^2 (< p R for both BN254 fields: the quotient is below 2p)
    static KZP_D void redc_wide(Fp& r, const uint32_t (&T)[16])
    {
        uint32_t x[8], y[8], pend = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            x[i] = T[i];
            y[i] = 0;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
            redc_round(x, y, pend, T[8 + i]);
        merge_reduce(r, x, y, pend);
    }
#endif

    // canonical integer -> Montgomery (Fr_rawToMontgomery: fr_raw_generic.cpp:192-196)
    static KZP_HD void to_mont(Fp& r, const Fp& a)
    {
        Fp k = r2();
        mul(r, a, k);
    }

    // Montgomery -> canonical integer (Fr_rawFromMontgomery: fr_raw_generic.cpp:198-232)
    static KZP_HD void from_mont(Fp& r, const Fp& a)
    {
        Fp k = zero();
        k.v[0] = 1;
        mul(r, a, k);
    }

    // r = a^e (e given as 8 x 32-bit limbs, plain integer); Montgomery in/out.
    static KZP_HD void pow(Fp& r, const Fp& a, const uint32_t (&e)[8])
    {
        Fp acc = one();
        for (int i = 255; i >= 0; i--)
        {
            sqr(acc, acc);
            if ((e[i >> 5] >> (i & 31)) & 1)
                mul(acc, acc, a);
        }
        r = acc;
    }

    // r = a^-1 via Fermat (a^(p-2)); inverse of 0 is 0. Montgomery in/out.
    static KZP_HD void inv(Fp& r, const Fp& a)
    {
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            e[i] = modulus_limb<P>(i);
        e[0] -= 2; // p is odd and p0 >= 2 for both fields: no borrow
        pow(r, a, e);
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

} // namespace kzp
