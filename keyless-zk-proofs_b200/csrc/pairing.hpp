// Host-side BN254 pairing check and Groth16 verification.
//
// The reference service re-verifies every proof it produced before answering (prover-service/src/request_handler/
// prover_handler.rs:329-336, through aptos-types -> ark-groth16 0.4.0 / ark-bn254 0.4.0, Cargo.lock:501-596). Those
// crates are not in /root/reference, so this is a restatement of the published algorithm, not of reference code:
//   Groth16 verify:  e(A, B) = e(alpha1, beta2) * e(sum_i pub_i IC_i, gamma2) * e(C, delta2)
//   written as a product check   e(-A, B) e(alpha1, beta2) e(L, gamma2) e(C, delta2) == 1.
// Any non-degenerate bilinear pairing on (G1, G2) accepts exactly the same proofs, so the plain ate pairing
// a(Q, P) = f_{T,Q}(P)^((p^12-1)/r) with T = t - 1 = 6x^2 is used (the optimal ate loop 6x+2 would save half the
// Miller iterations at the price of two Frobenius-twisted line steps).
//   tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), xi = 9 + u, Fq12 = Fq6[w]/(w^2 - v)   (alt_bn128 / EIP-197)
//   twist: E'(Fq2): y^2 = x^3 + 3/xi (D-type), untwist (x', y') -> (x' w^2, y' w^3)
//   final exponentiation: easy part (p^6-1)(p^2+1), hard part in base p with three powers by the BN parameter x;
//   the Frobenius constants xi^(i(p-1)/6) are computed at start-up from the modulus, none are typed in.
// Host only (64-bit-limb field of hostff.hpp); a few milliseconds per proof. Pinned by tests against the Python
// oracle's independent verifier (oracle/bn254.py groth16_verify) and by bilinearity checks.
#pragma once

#include <cstdint>
#include <vector>

#include "ec.cuh"
#include "hostff.hpp"

namespace kzp
{
namespace pairing
{

typedef Fp2T<HFq> F2;

static inline void mul_xi(F2& r, const F2& x)
{
    // (a + bu)(9 + u) = (9a - b) + (9b + a) u
    HFq a8, b8, a9, b9;
    HFq::add(a8, x.a, x.a); HFq::add(a8, a8, a8); HFq::add(a8, a8, a8); HFq::add(a9, a8, x.a);
    HFq::add(b8, x.b, x.b); HFq::add(b8, b8, b8); HFq::add(b8, b8, b8); HFq::add(b9, b8, x.b);
    F2 t;
    HFq::sub(t.a, a9, x.b);
    HFq::add(t.b, b9, x.a);
    r = t;
}

struct F6
{
    F2 c0, c1, c2;
    static F6 zero() { return F6{F2::zero(), F2::zero(), F2::zero()}; }
    static F6 one() { return F6{F2::one(), F2::zero(), F2::zero()}; }
    static bool is_zero(const F6& x) { return F2::is_zero(x.c0) && F2::is_zero(x.c1) && F2::is_zero(x.c2); }
    static bool eq(const F6& x, const F6& y) { return F2::eq(x.c0, y.c0) && F2::eq(x.c1, y.c1) && F2::eq(x.c2, y.c2); }
    static void add(F6& r, const F6& x, const F6& y) { F2::add(r.c0, x.c0, y.c0); F2::add(r.c1, x.c1, y.c1); F2::add(r.c2, x.c2, y.c2); }
    static void sub(F6& r, const F6& x, const F6& y) { F2::sub(r.c0, x.c0, y.c0); F2::sub(r.c1, x.c1, y.c1); F2::sub(r.c2, x.c2, y.c2); }
    static void neg(F6& r, const F6& x) { F2::neg(r.c0, x.c0); F2::neg(r.c1, x.c1); F2::neg(r.c2, x.c2); }
    // Karatsuba over v^3 = xi: 6 Fq2 products
    static void mul(F6& r, const F6& x, const F6& y)
    {
        F2 v0, v1, v2, t0, t1, t2, s;
        F2::mul(v0, x.c0, y.c0); F2::mul(v1, x.c1, y.c1); F2::mul(v2, x.c2, y.c2);
        // c0 = v0 + xi((x1+x2)(y1+y2) - v1 - v2)
        F2::add(t0, x.c1, x.c2); F2::add(s, y.c1, y.c2); F2::mul(t0, t0, s); F2::sub(t0, t0, v1); F2::sub(t0, t0, v2);
        mul_xi(t0, t0); F2::add(t0, t0, v0);
        // c1 = (x0+x1)(y0+y1) - v0 - v1 + xi v2
        F2::add(t1, x.c0, x.c1); F2::add(s, y.c0, y.c1); F2::mul(t1, t1, s); F2::sub(t1, t1, v0); F2::sub(t1, t1, v1);
        mul_xi(s, v2); F2::add(t1, t1, s);
        // c2 = (x0+x2)(y0+y2) - v0 - v2 + v1
        F2::add(t2, x.c0, x.c2); F2::add(s, y.c0, y.c2); F2::mul(t2, t2, s); F2::sub(t2, t2, v0); F2::sub(t2, t2, v2);
        F2::add(t2, t2, v1);
        r.c0 = t0; r.c1 = t1; r.c2 = t2;
    }
    static void mul_v(F6& r, const F6& x)
    {
        F2 t;
        mul_xi(t, x.c2);
        F6 o{t, x.c0, x.c1};
        r = o;
    }
    static void inv(F6& r, const F6& x)
    {
        F2 t0, t1, t2, s, d;
        F2::sqr(t0, x.c0); F2::mul(s, x.c1, x.c2); mul_xi(s, s); F2::sub(t0, t0, s);      // c0^2 - xi c1 c2
        F2::sqr(t1, x.c2); mul_xi(t1, t1); F2::mul(s, x.c0, x.c1); F2::sub(t1, t1, s);      // xi c2^2 - c0 c1
        F2::sqr(t2, x.c1); F2::mul(s, x.c0, x.c2); F2::sub(t2, t2, s);                      // c1^2 - c0 c2
        F2 a, b;
        F2::mul(a, x.c2, t1); F2::mul(b, x.c1, t2); F2::add(a, a, b); mul_xi(a, a);
        F2::mul(d, x.c0, t0); F2::add(d, d, a);
        F2::inv(d, d);
        F2::mul(r.c0, t0, d); F2::mul(r.c1, t1, d); F2::mul(r.c2, t2, d);
    }
};

struct F12
{
    F6 a, b; // a + b w, w^2 = v
    static F12 one() { return F12{F6::one(), F6::zero()}; }
    static bool is_one(const F12& x) { return F6::eq(x.a, F6::one()) && F6::is_zero(x.b); }
    static void mul(F12& r, const F12& x, const F12& y)
    {
        F6 aa, bb, s, t, m;
        F6::mul(aa, x.a, y.a); F6::mul(bb, x.b, y.b);
        F6::add(s, x.a, x.b); F6::add(t, y.a, y.b); F6::mul(m, s, t); F6::sub(m, m, aa); F6::sub(m, m, bb);
        F6::mul_v(bb, bb); F6::add(aa, aa, bb);
        r.a = aa; r.b = m;
    }
    static void sqr(F12& r, const F12& x) { mul(r, x, x); }
    static void conj(F12& r, const F12& x) { r.a = x.a; F6::neg(r.b, x.b); }
    static void inv(F12& r, const F12& x)
    {
        F6 t0, t1;
        F6::mul(t0, x.a, x.a); F6::mul(t1, x.b, x.b); F6::mul_v(t1, t1); F6::sub(t0, t0, t1);
        F6::inv(t0, t0);
        F6::mul(r.a, x.a, t0);
        F6::mul(t1, x.b, t0); F6::neg(r.b, t1);
    }
};

struct G1Aff { HFq x, y; bool inf; };
struct G2Aff { F2 x, y; bool inf; };

// f *= line through the twist point (x1, y1) with twist slope lam, evaluated at P:
//   l(P) = yP + (-lam xP) w + (lam x1 - y1) w^3,   w^3 = v w
static inline void mul_line(F12& f, const F2& lam, const F2& x1, const F2& y1, const G1Aff& P)
{
    F12 l;
    l.a = F6::zero();
    l.a.c0.a = P.y;
    F2 t;
    t.a = P.x; t.b = HFq::zero();
    F2::mul(t, lam, t); F2::neg(l.b.c0, t);
    F2::mul(t, lam, x1); F2::sub(l.b.c1, t, y1);
    l.b.c2 = F2::zero();
    F12::mul(f, f, l);
}

// T = t - 1 = 6 x^2, x = 4965661367192848881 (127 bits)
static const uint64_t kAteLoop[2] = {0xf83e9682e87cfd46ull, 0x6f4d8248eeb859fbull};
// in-place inversion of every element with ONE field inversion (Montgomery's trick); zero entries are not allowed
static inline void batch_inv(std::vector<F2>& x)
{
    if (x.empty())
        return;
    std::vector<F2> pre(x.size());
    F2              run = F2::one();
    for (size_t i = 0; i < x.size(); i++)
    {
        pre[i] = run;
        F2::mul(run, run, x[i]);
    }
    F2 inv;
    F2::inv(inv, run);
    for (size_t i = x.size(); i-- > 0;)
    {
        F2 t;
        F2::mul(t, inv, pre[i]);
        F2::mul(inv, inv, x[i]);
        x[i] = t;
    }
}

// One Miller step for every live pair: slope numerators/denominators are collected first so that the step costs one
// field inversion in total. add == false: tangent at R; add == true: chord through R and Q.
static inline void miller_step(F12& f, std::vector<G2Aff>& R, const std::vector<G2Aff>& Qs, const std::vector<G1Aff>& Ps,
                               const std::vector<size_t>& live, bool add)
{
    std::vector<F2>     num, den;
    std::vector<size_t> who;
    for (size_t k = 0; k < live.size(); k++)
    {
        G2Aff&       r = R[k];
        const G2Aff& q = Qs[live[k]];
        if (r.inf)
        {
            if (add)
                r = q; // O + Q
            continue;
        }
        bool tangent = !add;
        F2   n, d;
        if (add)
        {
            F2::sub(d, q.x, r.x);
            F2::sub(n, q.y, r.y);
            if (F2::is_zero(d))
            {
                if (!F2::is_zero(n))
                {
                    r.inf = true; // R == -Q: vertical line, eliminated by the final exponentiation
                    continue;
                }
                tangent = true; // R == Q
            }
        }
        if (tangent)
        {
            if (F2::is_zero(r.y))
            {
                r.inf = true; // 2-torsion: vertical tangent (not reachable for points of order r)
                continue;
            }
            F2 t;
            F2::sqr(t, r.x); F2::dbl(n, t); F2::add(n, n, t); // 3 x^2
            F2::dbl(d, r.y);                                  // 2 y
        }
        num.push_back(n);
        den.push_back(d);
        who.push_back(k);
    }
    batch_inv(den);
    for (size_t j = 0; j < who.size(); j++)
    {
        size_t       k = who[j];
        G2Aff&       r = R[k];
        const G2Aff& q = Qs[live[k]];
        F2           lam, t, x3, y3;
        F2::mul(lam, num[j], den[j]);
        mul_line(f, lam, r.x, r.y, Ps[live[k]]);
        F2::sqr(x3, lam); F2::sub(x3, x3, r.x); F2::sub(x3, x3, add ? q.x : r.x);
        F2::sub(t, r.x, x3); F2::mul(y3, lam, t); F2::sub(y3, y3, r.y);
        r.x = x3; r.y = y3;
    }
}

// prod_i f_{T,Q_i}(P_i); pairs with an infinity member contribute 1
static inline F12 miller_loop(const std::vector<G1Aff>& Ps, const std::vector<G2Aff>& Qs)
{
    std::vector<size_t> live;
    for (size_t i = 0; i < Ps.size(); i++)
        if (!Ps[i].inf && !Qs[i].inf)
            live.push_back(i);
    std::vector<G2Aff> R;
    for (size_t i : live)
        R.push_back(Qs[i]);
    F12 f = F12::one();
    for (int bit = 125; bit >= 0; bit--)
    {
        F12::sqr(f, f);
        miller_step(f, R, Qs, Ps, live, false);
        if ((kAteLoop[bit >> 6] >> (bit & 63)) & 1)
            miller_step(f, R, Qs, Ps, live, true);
    }
    return f;
}

// gamma_i = xi^(i (p-1)/6), i < 6: w^p = gamma_1 w, so the p-power Frobenius acts on f = sum_i c_i w^i (c_i in Fq2)
// as c_i -> conj(c_i) gamma_i. Computed once from the field modulus (no typed-in constants).
struct FrobeniusTable
{
    F2 g[6];
    FrobeniusTable()
    {
        // e = (p - 1) / 6 by long division of the 4 x 64-bit modulus
        uint64_t e[4];
        for (int i = 0; i < 4; i++)
            e[i] = HFq::p(i);
        e[0] -= 1;
        unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--)
        {
            unsigned __int128 cur = (rem << 64) | e[i];
            e[i]                  = (uint64_t)(cur / 6);
            rem                   = cur % 6;
        }
        F2 xi;
        mul_xi(xi, F2::one());
        F2 acc = F2::one();
        for (int i = 255; i >= 0; i--)
        {
            F2::sqr(acc, acc);
            if ((e[i >> 6] >> (i & 63)) & 1)
                F2::mul(acc, acc, xi);
        }
        g[0] = F2::one();
        for (int i = 1; i < 6; i++)
            F2::mul(g[i], g[i - 1], acc);
    }
};

static inline void frobenius(F12& r, const F12& x)
{
    static const FrobeniusTable T;
    auto                        cj = [](const F2& c, const F2& g) {
        F2 t = c;
        HFq::neg(t.b, c.b);
        F2::mul(t, t, g);
        return t;
    };
    F12 o;
    o.a.c0 = cj(x.a.c0, T.g[0]); // w^0
    o.b.c0 = cj(x.b.c0, T.g[1]); // w^1
    o.a.c1 = cj(x.a.c1, T.g[2]); // w^2 = v
    o.b.c1 = cj(x.b.c1, T.g[3]); // w^3 = v w
    o.a.c2 = cj(x.a.c2, T.g[4]); // w^4 = v^2
    o.b.c2 = cj(x.b.c2, T.g[5]); // w^5 = v^2 w
    r = o;
}

static inline void pow_u64(F12& r, const F12& f, uint64_t e)
{
    F12 acc = F12::one();
    for (int i = 63; i >= 0; i--)
    {
        F12::sqr(acc, acc);
        if ((e >> i) & 1)
            F12::mul(acc, acc, f);
    }
    r = acc;
}

// f^((p^12 - 1)/r) = easy part (p^6 - 1)(p^2 + 1), then the hard part (p^4 - p^2 + 1)/r written in base p
// (Devegili-Scott-Dahab): lambda3 = 1, lambda2 = 6x^2 + 1, lambda1 = -36x^3 - 18x^2 - 12x + 1,
// lambda0 = -36x^3 - 30x^2 - 18x - 2, x = 4965661367192848881 — an exact integer identity, re-checked in
// tests/test_oracle_golden.py::test_pairing_constants. After the easy part the element is unitary: inverse = conjugate.
static const uint64_t kBnX = 0x44e992b44a6909f1ull;

static inline F12 final_exponentiation(const F12& f0)
{
    F12 c, i, f, t;
    F12::conj(c, f0);
    F12::inv(i, f0);
    F12::mul(f, c, i);      // f0^(p^6 - 1)
    frobenius(t, f);
    frobenius(t, t);
    F12::mul(f, t, f);      // ^(p^2 + 1)

    F12 fx, fx2, fx3;
    pow_u64(fx, f, kBnX);
    pow_u64(fx2, fx, kBnX);
    pow_u64(fx3, fx2, kBnX);
    F12 a6, a12, a18, b6, b18, b30, c36, f2, y, r0, r1, r2, r3;
    pow_u64(a6, fx, 6);   F12::sqr(a12, a6);  F12::mul(a18, a12, a6);           // fx^6, ^12, ^18
    pow_u64(b6, fx2, 6);  pow_u64(b18, b6, 3); pow_u64(b30, b6, 5);              // fx2^6, ^18, ^30
    pow_u64(c36, fx3, 36);                                                       // fx3^36
    F12::sqr(f2, f);
    F12::mul(r2, b6, f);                                                         // f^lambda2
    F12::mul(y, c36, b18); F12::mul(y, y, a12); F12::conj(y, y); F12::mul(r1, y, f);      // f^lambda1
    F12::mul(y, c36, b30); F12::mul(y, y, a18); F12::mul(y, y, f2); F12::conj(r0, y);     // f^lambda0
    frobenius(r3, f); frobenius(r3, r3); frobenius(r3, r3);                      // f^(p^3)
    frobenius(r2, r2); frobenius(r2, r2);
    frobenius(r1, r1);
    F12 out;
    F12::mul(out, r3, r2); F12::mul(out, out, r1); F12::mul(out, out, r0);
    return out;
}

// prod_i e(P_i, Q_i) == 1 ?
static inline bool pairing_product_is_one(const std::vector<G1Aff>& Ps, const std::vector<G2Aff>& Qs)
{
    F12 f = miller_loop(Ps, Qs);
    return F12::is_one(final_exponentiation(f));
}

// curve membership (y^2 = x^3 + 3 on G1; y^2 = x^3 + 3/xi on the twist); infinity is accepted.
// Subgroup membership of G2 points is NOT checked (the proving key's and the prover's own points are trusted to be
// in the r-torsion; ark-groth16's verifier does not check it either, deserialisation does).
static inline bool g1_on_curve(const G1Aff& p)
{
    if (p.inf)
        return true;
    HFq y2, x3, three, t;
    HFq::sqr(y2, p.y); HFq::sqr(x3, p.x); HFq::mul(x3, x3, p.x);
    t = HFq::one(); HFq::add(three, t, t); HFq::add(three, three, t);
    HFq::add(x3, x3, three);
    return HFq::eq(y2, x3);
}
static inline bool g2_on_curve(const G2Aff& p)
{
    if (p.inf)
        return true;
    F2 y2, x3, b, xi, three;
    F2::sqr(y2, p.y); F2::sqr(x3, p.x); F2::mul(x3, x3, p.x);
    three = F2::one(); F2 o = F2::one(); F2::add(three, three, o); F2::add(three, three, o);
    xi = F2::zero(); mul_xi(xi, F2::one());
    F2::inv(xi, xi); F2::mul(b, three, xi);
    F2::add(x3, x3, b);
    return F2::eq(y2, x3);
}

} // namespace pairing
} // namespace kzp
