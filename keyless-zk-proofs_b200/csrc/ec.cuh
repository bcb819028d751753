// Fq2 and the short-Weierstrass group law in XYZZ coordinates for BN254 G1 (over Fq) and G2 (over Fq2).
//
// Mirrors the reference's representation so results can be compared artefact by artefact:
//   Fq2 = Fq[u]/(u^2+1), element = a + b*u, 3-multiplication product and 2-multiplication square
//         (rust-rapidsnark/rapidsnark/src/f2field.cpp:122-175, alt_bn128.hpp:43)
//   Point = (x, y, zz, zzz) with affine = (x/zz, y/zzz); infinity <=> zz == 0 (curve.cpp:532-534);
//   affine infinity <=> (0,0) (curve.cpp:537-539)
//   formulas: EFD xyzz madd-2008-s / add-2008-s / dbl-2008-s with the reference's exceptional
//   cases (P==Q -> dbl, P==-Q -> zz=0) (curve.cpp:91-166, 185-250, 340-458). a = 0 for both curves.
// Everything is templated on the field type so the same code serves G1/G2 on the device and the
// host-side proof assembly (hostff.hpp supplies a 64-bit-limb field with the same interface).
#pragma once

#include "ff.cuh"

namespace kzp
{

// ------------------------------------------------------------------ Fq2
template <class F>
struct Fp2T;
#if defined(__CUDACC__)
// On the device an Fq2 product is a CALL: a G2 mixed addition is 28 base-field products, 5-6 K instructions when
// everything is inlined — more than the instruction cache holds for warps that are not in lockstep (ncu on the G2
// accumulate kernel: 1.9 no-instruction stall cycles per issue, 49 % of the multiplier pipe). Out of line the whole
// group law is a few dozen calls of two small bodies.
template <class F>
__device__ __noinline__ void fp2_mul_ool(Fp2T<F>& r, const Fp2T<F>& x, const Fp2T<F>& y);
template <class F>
__device__ __noinline__ void fp2_sqr_ool(Fp2T<F>& r, const Fp2T<F>& x);
template <class F>
__device__ __noinline__ void fp2_mul_add2_ool(Fp2T<F>& r, const Fp2T<F>& x, const Fp2T<F>& y, const Fp2T<F>& z, const Fp2T<F>& w);
#endif

template <class F>
struct alignas(16) Fp2T
{
    F a; // real part
    F b; // coefficient of u

    // y3 = (Q - x3) R - y1 PPP as one dual product (mul_add2) over the device's base field; the host's 4 x 64-bit
    // base field (hostff.hpp) keeps the two-product form
    static constexpr bool kFusedMulAdd2 = F::kFusedMulAdd2;

    static KZP_HD Fp2T zero()
    {
        Fp2T r;
        r.a = F::zero();
        r.b = F::zero();
        return r;
    }
    static KZP_HD Fp2T one()
    {
        Fp2T r;
        r.a = F::one();
        r.b = F::zero();
        return r;
    }
    static KZP_HD bool is_zero(const Fp2T& x) { return F::is_zero(x.a) && F::is_zero(x.b); }
    static KZP_HD bool eq(const Fp2T& x, const Fp2T& y) { return F::eq(x.a, y.a) && F::eq(x.b, y.b); }
    static KZP_HD void add(Fp2T& r, const Fp2T& x, const Fp2T& y)
    {
        F::add(r.a, x.a, y.a);
        F::add(r.b, x.b, y.b);
    }
    static KZP_HD void sub(Fp2T& r, const Fp2T& x, const Fp2T& y)
    {
        F::sub(r.a, x.a, y.a);
        F::sub(r.b, x.b, y.b);
    }
    static KZP_HD void dbl(Fp2T& r, const Fp2T& x)
    {
        F::add(r.a, x.a, x.a);
        F::add(r.b, x.b, x.b);
    }
    static KZP_HD void neg(Fp2T& r, const Fp2T& x)
    {
        F::neg(r.a, x.a);
        F::neg(r.b, x.b);
    }
    // (a+bu)(c+du) = (ac - bd) + ((a+b)(c+d) - ac - bd) u     [f2field.cpp:122-141]
    static KZP_HD void mul(Fp2T& r, const Fp2T& x, const Fp2T& y)
    {
#if defined(__CUDA_ARCH__)
        fp2_mul_ool<F>(r, x, y);
#else
        mul_body(r, x, y);
#endif
    }
    static KZP_HD void sqr(Fp2T& r, const Fp2T& x)
    {
#if defined(__CUDA_ARCH__)
        fp2_sqr_ool<F>(r, x);
#else
        sqr_body(r, x);
#endif
    }
    // (r may be x or y: every operand is read before the first write)
    // Device: lazy reduction — the three Karatsuba products stay unreduced 512-bit integers, the additions and
    // subtractions happen there (p^2 keeps the real part non-negative), and only the two results are reduced:
    // 3 x 64 + 2 x 72 wide multiply-adds instead of 3 x 136. Results are canonical, i.e. the same bits.
    static KZP_HD void mul_body(Fp2T& r, const Fp2T& x, const Fp2T& y)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (F::kFusedMulAdd2) // (the device's own base field; host-field instantiations take the plain form)
        {
        uint32_t Ta[16], Tb[16], Tc[16];
        F        s1, s2;
        F::add_raw(s1, x.a, x.b);
        F::add_raw(s2, y.a, y.b);
        F::mul_wide(Tc, s1, s2);
        F::mul_wide(Ta, x.a, y.a);
        F::mul_wide(Tb, x.b, y.b);
        F::wide_sub(Tc, Ta);
        F::wide_sub(Tc, Tb); // x.a y.b + x.b y.a < 2 p^2
        F::template wide_add_psq<1>(Ta);
        F::wide_sub(Ta, Tb); // x.a y.a - x.b y.b + p^2 in (0, 2 p^2)
        F::redc_wide(r.a, Ta);
        F::redc_wide(r.b, Tc);
        return;
        }
#endif
        F aa, bb, s1, s2, t;
        F::mul(aa, x.a, y.a);
        F::mul(bb, x.b, y.b);
        F::add(s1, x.a, x.b);
        F::add(s2, y.a, y.b);
        F::mul(t, s1, s2);
        F::sub(t, t, aa);
        F::sub(r.b, t, bb);
        F::sub(r.a, aa, bb);
    }
    // r = x y + z w with one reduction per component (device: six wide products, two reductions)
    static KZP_HD void mul_add2(Fp2T& r, const Fp2T& x, const Fp2T& y, const Fp2T& z, const Fp2T& w)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (F::kFusedMulAdd2)
        {
            fp2_mul_add2_ool<F>(r, x, y, z, w);
            return;
        }
#endif
        Fp2T t, u;
        mul_body(t, x, y);
        mul_body(u, z, w);
        add(r, t, u);
    }
#if defined(__CUDA_ARCH__)
    static KZP_D void mul_add2_body(Fp2T& r, const Fp2T& x, const Fp2T& y, const Fp2T& z, const Fp2T& w)
    {
        uint32_t Ta[16], Tc[16], U[16];
        F        s1, s2;
        F::add_raw(s1, x.a, x.b);
        F::add_raw(s2, y.a, y.b);
        F::mul_wide(Tc, s1, s2);
        F::mul_wide(Ta, x.a, y.a);
        F::mul_wide(U, x.b, y.b);
        F::wide_sub(Tc, Ta);
        F::wide_sub(Tc, U);
        F::template wide_add_psq<2>(Ta);
        F::wide_sub(Ta, U); // x.a y.a - x.b y.b + 2 p^2
        F::add_raw(s1, z.a, z.b);
        F::add_raw(s2, w.a, w.b);
        F::mul_wide(U, s1, s2);
        F::wide_add(Tc, U);
        F::mul_wide(U, z.a, w.a);
        F::wide_add(Ta, U);
        F::wide_sub(Tc, U);
        F::mul_wide(U, z.b, w.b);
        F::wide_sub(Ta, U); // real part + 2 p^2, in (0, 4 p^2)
        F::wide_sub(Tc, U); // imaginary part, in [0, 4 p^2)
        F::redc_wide(r.a, Ta);
        F::redc_wide(r.b, Tc);
    }
#endif
    // (a+bu)^2 = (a+b)(a-b) + 2ab u                            [f2field.cpp:144-175]
    static KZP_HD void sqr_body(Fp2T& r, const Fp2T& x)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (F::kFusedMulAdd2)
        {
            // a + b and 2a stay unreduced (below 2p: the products stay inside the multiplier's input bound)
            F s, d, a2, ra;
            F::add_raw(s, x.a, x.b);
            F::sub(d, x.a, x.b);
            F::add_raw(a2, x.a, x.a);
            F::mul(ra, s, d);
            F::mul(r.b, a2, x.b);
            r.a = ra;
            return;
        }
#endif
        F s, d, ab;
        F::add(s, x.a, x.b);
        F::sub(d, x.a, x.b);
        F::mul(ab, x.a, x.b);
        F::mul(r.a, s, d);
        F::add(r.b, ab, ab);
    }
    // 1/(a+bu) = (a - bu)/(a^2+b^2)                             [f2field.cpp:178-189]
    static KZP_HD void inv(Fp2T& r, const Fp2T& x)
    {
        F t0, t1;
        F::sqr(t0, x.a);
        F::sqr(t1, x.b);
        F::add(t0, t0, t1);
        F::inv(t1, t0);
        F::mul(r.a, x.a, t1);
        F::mul(t0, x.b, t1);
        F::neg(r.b, t0);
    }
};

typedef Fp2T<Fq> Fq2;

#if defined(__CUDACC__)
template <class F>
__device__ __noinline__ void fp2_mul_ool(Fp2T<F>& r, const Fp2T<F>& x, const Fp2T<F>& y)
{
    Fp2T<F>::mul_body(r, x, y);
}
template <class F>
__device__ __noinline__ void fp2_sqr_ool(Fp2T<F>& r, const Fp2T<F>& x)
{
    Fp2T<F>::sqr_body(r, x);
}
template <class F>
__device__ __noinline__ void fp2_mul_add2_ool(Fp2T<F>& r, const Fp2T<F>& x, const Fp2T<F>& y, const Fp2T<F>& z, const Fp2T<F>& w)
{
    Fp2T<F>::mul_add2_body(r, x, y, z, w);
}
#endif

// ------------------------------------------------------------------ points
template <class F>
struct alignas(16) AffineT
{
    F x, y;
    static KZP_HD bool is_inf(const AffineT& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
    static KZP_HD void neg(AffineT& r, const AffineT& p)
    {
        r.x = p.x;
        F::neg(r.y, p.y);
    }
};

template <class F>
struct alignas(16) XyzzT
{
    F x, y, zz, zzz;

    typedef F         Field;
    typedef AffineT<F> Affine;

    static KZP_HD bool is_inf(const XyzzT& p) { return F::is_zero(p.zz); }

    static KZP_HD void set_inf(XyzzT& p)
    {
        // same encoding the reference uses for copy(Point, affine infinity) (curve.cpp:547-553)
        p.x   = F::one();
        p.y   = F::one();
        p.zz  = F::zero();
        p.zzz = F::zero();
    }

    static KZP_HD void from_affine(XyzzT& r, const Affine& p)
    {
        if (Affine::is_inf(p))
        {
            set_inf(r);
            return;
        }
        r.x   = p.x;
        r.y   = p.y;
        r.zz  = F::one();
        r.zzz = F::one();
    }

    static KZP_HD void neg(XyzzT& r, const XyzzT& p)
    {
        r.x = p.x;
        F::neg(r.y, p.y);
        r.zz  = p.zz;
        r.zzz = p.zzz;
    }

    // dbl-2008-s with a = 0 (curve.cpp:340-401)
    static KZP_HD void dbl(XyzzT& r, const XyzzT& p)
    {
        if (is_inf(p))
        {
            r = p;
            return;
        }
        F U, V, W, S, M, t, x3, y3;
        F::add(U, p.y, p.y);
        F::sqr(V, U);
        F::mul(W, U, V);
        F::mul(S, p.x, V);
        F::sqr(M, p.x);
        F::add(t, M, M);
        F::add(M, M, t);
        F::sqr(x3, M);
        F::sub(x3, x3, S);
        F::sub(x3, x3, S);
        F::mul(t, W, p.y);
        F::sub(y3, S, x3);
        F::mul(y3, M, y3);
        F::sub(y3, y3, t);
        F::mul(r.zz, V, p.zz);
        F::mul(r.zzz, W, p.zzz);
        r.x = x3;
        r.y = y3;
    }

    // mdbl-2008-s with a = 0 (curve.cpp:411-458)
    static KZP_HD void dbl_affine(XyzzT& r, const Affine& p)
    {
        if (Affine::is_inf(p))
        {
            set_inf(r);
            return;
        }
        F U, S, M, t, x3, y3, V, W;
        F::add(U, p.y, p.y);
        F::sqr(V, U);
        F::mul(W, U, V);
        F::mul(S, p.x, V);
        F::sqr(M, p.x);
        F::add(t, M, M);
        F::add(M, M, t);
        F::sqr(x3, M);
        F::sub(x3, x3, S);
        F::sub(x3, x3, S);
        F::mul(t, W, p.y);
        F::sub(y3, S, x3);
        F::mul(y3, M, y3);
        F::sub(y3, y3, t);
        r.x   = x3;
        r.y   = y3;
        r.zz  = V;
        r.zzz = W;
    }

    // acc += q (mixed, madd-2008-s; curve.cpp:185-250). All exceptional cases handled.
    static KZP_HD void madd(XyzzT& acc, const Affine& q)
    {
        if (Affine::is_inf(q))
            return;
        if (is_inf(acc))
        {
            acc.x   = q.x;
            acc.y   = q.y;
            acc.zz  = F::one();
            acc.zzz = F::one();
            return;
        }
        F U2, S2, P, R;
        F::mul(U2, q.x, acc.zz);
        F::mul(S2, q.y, acc.zzz);
        F::sub(P, U2, acc.x);
        F::sub(R, S2, acc.y);
        if (F::is_zero(P))
        {
            if (F::is_zero(R))
            {
                dbl_affine(acc, q);
                return;
            }
            set_inf(acc); // P == -Q
            return;
        }
        F PP, PPP, Q, t;
        F::sqr(PP, P);
        F::mul(PPP, P, PP);
        F::mul(Q, acc.x, PP);
        F::sqr(acc.x, R);
        F::sub(acc.x, acc.x, PPP);
        F::sub(acc.x, acc.x, Q);
        F::sub(acc.x, acc.x, Q);
        if constexpr (F::kFusedMulAdd2)
        {
            // y3 = (Q - x3) R - y1 PPP as ONE dual product with a single reduction (same residue, canonical)
            F ny;
            F::neg(ny, acc.y);
            F::sub(t, Q, acc.x);
            F::mul_add2(acc.y, t, R, ny, PPP);
        }
        else
        {
            F::mul(t, acc.y, PPP);
            F::sub(acc.y, Q, acc.x);
            F::mul(acc.y, acc.y, R);
            F::sub(acc.y, acc.y, t);
        }
        F::mul(acc.zz, acc.zz, PP);
        F::mul(acc.zzz, acc.zzz, PPP);
    }

    // acc += q (add-2008-s; curve.cpp:91-166)
    static KZP_HD void add(XyzzT& acc, const XyzzT& q)
    {
        if (is_inf(q))
            return;
        if (is_inf(acc))
        {
            acc = q;
            return;
        }
        F U1, U2, S1, S2, P, R;
        F::mul(U1, acc.x, q.zz);
        F::mul(U2, q.x, acc.zz);
        F::mul(S1, acc.y, q.zzz);
        F::mul(S2, q.y, acc.zzz);
        F::sub(P, U2, U1);
        F::sub(R, S2, S1);
        if (F::is_zero(P))
        {
            if (F::is_zero(R))
            {
                XyzzT t = acc;
                dbl(acc, t);
                return;
            }
            set_inf(acc);
            return;
        }
        F PP, PPP, Q, t;
        F::sqr(PP, P);
        F::mul(PPP, P, PP);
        F::mul(Q, U1, PP);
        F::sqr(acc.x, R);
        F::sub(acc.x, acc.x, PPP);
        F::sub(acc.x, acc.x, Q);
        F::sub(acc.x, acc.x, Q);
        if constexpr (F::kFusedMulAdd2)
        {
            F ns;
            F::neg(ns, S1);
            F::sub(t, Q, acc.x);
            F::mul_add2(acc.y, t, R, ns, PPP);
        }
        else
        {
            F::mul(t, S1, PPP);
            F::sub(acc.y, Q, acc.x);
            F::mul(acc.y, acc.y, R);
            F::sub(acc.y, acc.y, t);
        }
        F::mul(acc.zz, acc.zz, q.zz);
        F::mul(acc.zz, acc.zz, PP);
        F::mul(acc.zzz, acc.zzz, q.zzz);
        F::mul(acc.zzz, acc.zzz, PPP);
    }

    // affine = (x/zz, y/zzz) (curve.cpp:565-576); one inversion: 1/zz = (zz/zzz)^2 since zz^3 = zzz^2.
    static KZP_HD void to_affine(Affine& r, const XyzzT& p)
    {
        if (is_inf(p))
        {
            r.x = F::zero();
            r.y = F::zero();
            return;
        }
        F izzz, t, izz;
        F::inv(izzz, p.zzz);
        F::mul(t, p.zz, izzz);
        F::sqr(izz, t);
        F::mul(r.x, p.x, izz);
        F::mul(r.y, p.y, izzz);
    }
};

typedef AffineT<Fq>  G1Affine;
typedef AffineT<Fq2> G2Affine;
typedef XyzzT<Fq>    G1Xyzz;
typedef XyzzT<Fq2>   G2Xyzz;

} // namespace kzp
