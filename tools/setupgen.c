/* Synthetic input generator for benchmarks and tests — NOT the product and NOT the oracle.
 *
 * Builds a keyless-shaped R1CS with a satisfying witness and a snarkjs-compatible Groth16 proving key from a known
 * trapdoor, and writes them in the reference's on-disk formats (zkey / wtns, SURVEY.md
 * Appendix A). bench.py uses it to synthesise the workload BASELINE.json names (the real keyless zkey cannot be
 * downloaded offline). Self-contained CPU code (its own copy of the little BN254 arithmetic it needs), so that the
 * benchmark's measured arm never touches oracle/. tests/test_oracle_golden.py checks that this generator and the
 * oracle's (oracle/kzp_port.c, oracle/bn254.py) emit byte-identical files.
 *
 *   cc -O3 -fopenmp -shared -fPIC tools/setupgen.c -o tools/libkzp_setupgen.so
 *   int kzp_setupgen(n_constraints, n_vars, seed, zkey_path, wtns_path, info[8])
 */
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;

/* ------------------------------------------------------------------ field parameters */
typedef struct
{
    uint64_t p[4];
    uint64_t np;  /* -p^-1 mod 2^64 */
    fe       r2;  /* R^2 mod p */
    fe       one; /* R mod p */
} fparams;

/* RS/fr_raw_generic.cpp:5-7 */
static const fparams FR = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                           0xc2e1f593efffffffull,
                           {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}},
                           {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}}};
/* RS/fq_raw_generic.cpp:6-8 */
static const fparams FQ = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                           0x87d20782e4866389ull,
                           {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}},
                           {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}}};

static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b)
{
    return ((a->v[0] ^ b->v[0]) | (a->v[1] ^ b->v[1]) | (a->v[2] ^ b->v[2]) | (a->v[3] ^ b->v[3])) == 0;
}
static inline int fe_geq(const fe* a, const uint64_t p[4])
{
    for (int i = 3; i >= 0; i--)
    {
        if (a->v[i] > p[i]) return 1;
        if (a->v[i] < p[i]) return 0;
    }
    return 1;
}
static inline void fe_sub_p(fe* a, const uint64_t p[4])
{
    uint64_t bw = 0;
    for (int i = 0; i < 4; i++)
    {
        u128 d  = (u128)a->v[i] - p[i] - bw;
        a->v[i] = (uint64_t)d;
        bw      = (uint64_t)(d >> 64) & 1;
    }
}
static inline void f_add(fe* r, const fe* a, const fe* b, const fparams* F)
{
    fe   s;
    u128 c = 0;
    for (int i = 0; i < 4; i++)
    {
        c += (u128)a->v[i] + b->v[i];
        s.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || fe_geq(&s, F->p)) fe_sub_p(&s, F->p);
    *r = s;
}
static inline void f_sub(fe* r, const fe* a, const fe* b, const fparams* F)
{
    fe       s;
    uint64_t bw = 0;
    for (int i = 0; i < 4; i++)
    {
        u128 d = (u128)a->v[i] - b->v[i] - bw;
        s.v[i] = (uint64_t)d;
        bw     = (uint64_t)(d >> 64) & 1;
    }
    if (bw)
    {
        u128 c = 0;
        for (int i = 0; i < 4; i++)
        {
            c += (u128)s.v[i] + F->p[i];
            s.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    *r = s;
}
static inline void f_neg(fe* r, const fe* a, const fparams* F)
{
    if (fe_is_zero(a)) { *r = *a; return; }
    fe z = {{0, 0, 0, 0}};
    f_sub(r, &z, a, F);
}
/* Fr_rawMMul: word-serial Montgomery product, canonical output */
static inline void f_mul(fe* r, const fe* a, const fe* b, const fparams* F)
{
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
    {
        u128 c = 0;
        for (int j = 0; j < 4; j++)
        {
            c += (u128)a->v[j] * b->v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4]       = (uint64_t)c;
        t[5]       = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->np;
        c          = ((u128)m * F->p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++)
        {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe s = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq(&s, F->p)) fe_sub_p(&s, F->p);
    *r = s;
}
static inline void f_to_mont(fe* r, const fe* a, const fparams* F) { f_mul(r, a, &F->r2, F); }
static inline void f_from_mont(fe* r, const fe* a, const fparams* F)
{
    fe k = {{1, 0, 0, 0}};
    f_mul(r, a, &k, F);
}
static void f_pow(fe* r, const fe* a, const uint64_t e[4], const fparams* F)
{
    fe acc = F->one;
    for (int i = 255; i >= 0; i--)
    {
        f_mul(&acc, &acc, &acc, F);
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(&acc, &acc, a, F);
    }
    *r = acc;
}
static void f_inv(fe* r, const fe* a, const fparams* F)
{
    uint64_t e[4] = {F->p[0] - 2, F->p[1], F->p[2], F->p[3]};
    f_pow(r, a, e, F);
}
static inline void f_from_u64(fe* r, uint64_t x, const fparams* F)
{
    fe t = {{x, 0, 0, 0}};
    f_to_mont(r, &t, F);
}

#define fq_add(r, a, b) f_add(r, a, b, &FQ)
#define fq_sub(r, a, b) f_sub(r, a, b, &FQ)
#define fq_mul(r, a, b) f_mul(r, a, b, &FQ)
#define fq_neg(r, a) f_neg(r, a, &FQ)
#define fr_add(r, a, b) f_add(r, a, b, &FR)
#define fr_sub(r, a, b) f_sub(r, a, b, &FR)
#define fr_mul(r, a, b) f_mul(r, a, b, &FR)

/* ------------------------------------------------------------------ Fq2 (RS/f2field.cpp) */
typedef struct { fe a, b; } fe2;
static inline int  f2_is_zero(const fe2* x) { return fe_is_zero(&x->a) && fe_is_zero(&x->b); }
static inline void f2_add(fe2* r, const fe2* x, const fe2* y) { fq_add(&r->a, &x->a, &y->a); fq_add(&r->b, &x->b, &y->b); }
static inline void f2_sub(fe2* r, const fe2* x, const fe2* y) { fq_sub(&r->a, &x->a, &y->a); fq_sub(&r->b, &x->b, &y->b); }
static inline void f2_neg(fe2* r, const fe2* x) { fq_neg(&r->a, &x->a); fq_neg(&r->b, &x->b); }
static inline void f2_mul(fe2* r, const fe2* x, const fe2* y)
{
    fe aa, bb, s1, s2, t;
    fq_mul(&aa, &x->a, &y->a);
    fq_mul(&bb, &x->b, &y->b);
    fq_add(&s1, &x->a, &x->b);
    fq_add(&s2, &y->a, &y->b);
    fq_mul(&t, &s1, &s2);
    fq_sub(&t, &t, &aa);
    fq_sub(&r->b, &t, &bb);
    fq_sub(&r->a, &aa, &bb);
}
static inline void f2_sqr(fe2* r, const fe2* x)
{
    fe s, d, ab;
    fq_add(&s, &x->a, &x->b);
    fq_sub(&d, &x->a, &x->b);
    fq_mul(&ab, &x->a, &x->b);
    fq_mul(&r->a, &s, &d);
    fq_add(&r->b, &ab, &ab);
}
static void f2_inv(fe2* r, const fe2* x)
{
    fe t0, t1;
    fq_mul(&t0, &x->a, &x->a);
    fq_mul(&t1, &x->b, &x->b);
    fq_add(&t0, &t0, &t1);
    f_inv(&t1, &t0, &FQ);
    fq_mul(&r->a, &x->a, &t1);
    fq_mul(&t0, &x->b, &t1);
    fq_neg(&r->b, &t0);
}

/* ------------------------------------------------------------------ group law, generated for G1 (fe) and G2 (fe2) */
#define DEFINE_CURVE(PFX, FT, ADD, SUB, MUL, SQR, ISZ, ONE_INIT, ZERO_INIT)                                 \
    typedef struct { FT x, y; } PFX##_aff;                                                                  \
    typedef struct { FT x, y, zz, zzz; } PFX##_pt;                                                          \
    static inline int  PFX##_aff_is_inf(const PFX##_aff* p) { return ISZ(&p->x) && ISZ(&p->y); }            \
    static inline int  PFX##_is_inf(const PFX##_pt* p) { return ISZ(&p->zz); }                              \
    static inline void PFX##_set_inf(PFX##_pt* p)                                                           \
    {                                                                                                       \
        FT one = ONE_INIT, zero = ZERO_INIT;                                                                \
        p->x = one; p->y = one; p->zz = zero; p->zzz = zero;                                                \
    }                                                                                                       \
    static inline void PFX##_from_aff(PFX##_pt* r, const PFX##_aff* p)                                      \
    {                                                                                                       \
        if (PFX##_aff_is_inf(p)) { PFX##_set_inf(r); return; }                                              \
        FT one = ONE_INIT;                                                                                  \
        r->x = p->x; r->y = p->y; r->zz = one; r->zzz = one;                                                \
    }                                                                                                       \
    /* RS/curve.cpp:340-401, a = 0 */                                                                       \
    static void PFX##_dbl(PFX##_pt* r, const PFX##_pt* p)                                                   \
    {                                                                                                       \
        if (PFX##_is_inf(p)) { *r = *p; return; }                                                           \
        FT U, V, W, S, M, t, x3, y3;                                                                        \
        ADD(&U, &p->y, &p->y); SQR(&V, &U); MUL(&W, &U, &V); MUL(&S, &p->x, &V); SQR(&M, &p->x);            \
        ADD(&t, &M, &M); ADD(&M, &M, &t); SQR(&x3, &M); SUB(&x3, &x3, &S); SUB(&x3, &x3, &S);               \
        MUL(&t, &W, &p->y); SUB(&y3, &S, &x3); MUL(&y3, &M, &y3); SUB(&y3, &y3, &t);                        \
        MUL(&r->zz, &V, &p->zz); MUL(&r->zzz, &W, &p->zzz); r->x = x3; r->y = y3;                           \
    }                                                                                                       \
    /* RS/curve.cpp:411-458 */                                                                              \
    static void PFX##_dbl_aff(PFX##_pt* r, const PFX##_aff* p)                                              \
    {                                                                                                       \
        PFX##_pt t; PFX##_from_aff(&t, p); PFX##_dbl(r, &t);                                                \
    }                                                                                                       \
    /* RS/curve.cpp:185-250 */                                                                              \
    static void PFX##_madd(PFX##_pt* acc, const PFX##_aff* q)                                               \
    {                                                                                                       \
        if (PFX##_aff_is_inf(q)) return;                                                                    \
        if (PFX##_is_inf(acc)) { PFX##_from_aff(acc, q); return; }                                          \
        FT U2, S2, P, R, PP, PPP, Q, t;                                                                     \
        MUL(&U2, &q->x, &acc->zz); MUL(&S2, &q->y, &acc->zzz); SUB(&P, &U2, &acc->x); SUB(&R, &S2, &acc->y);\
        if (ISZ(&P)) { if (ISZ(&R)) { PFX##_dbl_aff(acc, q); return; } PFX##_set_inf(acc); return; }        \
        SQR(&PP, &P); MUL(&PPP, &P, &PP); MUL(&Q, &acc->x, &PP);                                            \
        SQR(&acc->x, &R); SUB(&acc->x, &acc->x, &PPP); SUB(&acc->x, &acc->x, &Q); SUB(&acc->x, &acc->x, &Q);\
        MUL(&t, &acc->y, &PPP); SUB(&acc->y, &Q, &acc->x); MUL(&acc->y, &acc->y, &R); SUB(&acc->y, &acc->y, &t); \
        MUL(&acc->zz, &acc->zz, &PP); MUL(&acc->zzz, &acc->zzz, &PPP);                                      \
    }                                                                                                       \
    /* RS/curve.cpp:91-166 */                                                                               \
    static void PFX##_add(PFX##_pt* acc, const PFX##_pt* q)                                                 \
    {                                                                                                       \
        if (PFX##_is_inf(q)) return;                                                                        \
        if (PFX##_is_inf(acc)) { *acc = *q; return; }                                                       \
        FT U1, U2, S1, S2, P, R, PP, PPP, Q, t;                                                             \
        MUL(&U1, &acc->x, &q->zz); MUL(&U2, &q->x, &acc->zz); MUL(&S1, &acc->y, &q->zzz);                   \
        MUL(&S2, &q->y, &acc->zzz); SUB(&P, &U2, &U1); SUB(&R, &S2, &S1);                                   \
        if (ISZ(&P)) { if (ISZ(&R)) { PFX##_pt c = *acc; PFX##_dbl(acc, &c); return; } PFX##_set_inf(acc); return; } \
        SQR(&PP, &P); MUL(&PPP, &P, &PP); MUL(&Q, &U1, &PP);                                                \
        SQR(&acc->x, &R); SUB(&acc->x, &acc->x, &PPP); SUB(&acc->x, &acc->x, &Q); SUB(&acc->x, &acc->x, &Q);\
        MUL(&t, &S1, &PPP); SUB(&acc->y, &Q, &acc->x); MUL(&acc->y, &acc->y, &R); SUB(&acc->y, &acc->y, &t);\
        MUL(&acc->zz, &acc->zz, &q->zz); MUL(&acc->zz, &acc->zz, &PP);                                      \
        MUL(&acc->zzz, &acc->zzz, &q->zzz); MUL(&acc->zzz, &acc->zzz, &PPP);                                \
    }

static inline void fq_sqr_(fe* r, const fe* a) { fq_mul(r, a, a); }
static inline void fq_add_(fe* r, const fe* a, const fe* b) { fq_add(r, a, b); }
static inline void fq_sub_(fe* r, const fe* a, const fe* b) { fq_sub(r, a, b); }
static inline void fq_mul_(fe* r, const fe* a, const fe* b) { fq_mul(r, a, b); }
#define FQ_ONE_INIT {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}}
#define FQ_ZERO_INIT {{0, 0, 0, 0}}
#define FQ2_ONE_INIT {FQ_ONE_INIT, FQ_ZERO_INIT}
#define FQ2_ZERO_INIT {FQ_ZERO_INIT, FQ_ZERO_INIT}
DEFINE_CURVE(g1, fe, fq_add_, fq_sub_, fq_mul_, fq_sqr_, fe_is_zero, FQ_ONE_INIT, FQ_ZERO_INIT)
DEFINE_CURVE(g2, fe2, f2_add, f2_sub, f2_mul, f2_sqr, f2_is_zero, FQ2_ONE_INIT, FQ2_ZERO_INIT)


/* affine = (x/zz, y/zzz) (RS/curve.cpp:565-576) */
static void g1_to_aff(g1_aff* r, const g1_pt* p)
{
    if (g1_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
    fe izzz, t, izz;
    f_inv(&izzz, &p->zzz, &FQ);
    fq_mul(&t, &p->zz, &izzz);
    fq_mul(&izz, &t, &t);
    fq_mul(&r->x, &p->x, &izz);
    fq_mul(&r->y, &p->y, &izzz);
}
static void g2_to_aff(g2_aff* r, const g2_pt* p)
{
    if (g2_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
    fe2 izzz, t, izz;
    f2_inv(&izzz, &p->zzz);
    f2_mul(&t, &p->zz, &izzz);
    f2_sqr(&izz, &t);
    f2_mul(&r->x, &p->x, &izz);
    f2_mul(&r->y, &p->y, &izzz);
}

/* batch affine conversion (Montgomery's trick) — generator side only */
static void g1_batch_to_aff(g1_aff* out, const g1_pt* in, size_t n)
{
    fe* pre = (fe*)malloc(sizeof(fe) * (n ? n : 1));
    fe  run = FQ.one;
    for (size_t i = 0; i < n; i++)
    {
        if (!g1_is_inf(&in[i])) fq_mul(&run, &run, &in[i].zzz);
        pre[i] = run;
    }
    fe inv;
    f_inv(&inv, &run, &FQ);
    for (size_t k = n; k-- > 0;)
    {
        if (g1_is_inf(&in[k])) { memset(&out[k], 0, sizeof(g1_aff)); continue; }
        fe izzz, t, izz;
        if (k > 0) fq_mul(&izzz, &inv, &pre[k - 1]); else izzz = inv;
        fq_mul(&inv, &inv, &in[k].zzz);
        fq_mul(&t, &in[k].zz, &izzz);
        fq_mul(&izz, &t, &t);
        fq_mul(&out[k].x, &in[k].x, &izz);
        fq_mul(&out[k].y, &in[k].y, &izzz);
    }
    free(pre);
}
static void g2_batch_to_aff(g2_aff* out, const g2_pt* in, size_t n)
{
    fe2* pre = (fe2*)malloc(sizeof(fe2) * (n ? n : 1));
    fe2  run = FQ2_ONE_INIT;
    for (size_t i = 0; i < n; i++)
    {
        if (!g2_is_inf(&in[i])) f2_mul(&run, &run, &in[i].zzz);
        pre[i] = run;
    }
    fe2 inv;
    f2_inv(&inv, &run);
    for (size_t k = n; k-- > 0;)
    {
        if (g2_is_inf(&in[k])) { memset(&out[k], 0, sizeof(g2_aff)); continue; }
        fe2 izzz, t, izz;
        if (k > 0) f2_mul(&izzz, &inv, &pre[k - 1]); else izzz = inv;
        f2_mul(&inv, &inv, &in[k].zzz);
        f2_mul(&t, &in[k].zz, &izzz);
        f2_sqr(&izz, &t);
        f2_mul(&out[k].x, &in[k].x, &izz);
        f2_mul(&out[k].y, &in[k].y, &izzz);
    }
    free(pre);
}

static void fr_root_of_unity(fe* w, uint32_t log_n)
{
    /* nqr = 5 is the smallest non-residue of Fr (RS/fft.cpp:60-66); w = 5^((r-1)/2^log_n) */
    uint64_t e[4] = {FR.p[0] - 1, FR.p[1], FR.p[2], FR.p[3]};
    for (uint32_t k = 0; k < log_n; k++)
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? (e[i + 1] << 63) : 0);
    fe g; f_from_u64(&g, 5, &FR);
    f_pow(w, &g, e, &FR);
}


/* ------------------------------------------------------------------ deterministic PRNG (splitmix64) */
typedef struct { uint64_t s; } rng_t;
static void     rng_init(rng_t* r, uint64_t seed) { r->s = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull; }
static uint64_t rng_next(rng_t* r)
{
    r->s += 0x9E3779B97F4A7C15ull;
    uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint64_t rng_below(rng_t* r, uint64_t n) { return rng_next(r) % n; }
/* 256 random bits reduced mod r (canonical) */
static void rng_fr(rng_t* r, fe* out)
{
    fe v;
    for (int i = 0; i < 4; i++) v.v[i] = rng_next(r);
    while (fe_geq(&v, FR.p)) fe_sub_p(&v, FR.p);
    *out = v;
}

/* ------------------------------------------------------------------ synthetic circuit (keyless-shaped: ~80 % bit wires, ~15 % bytes, ~5 % field products) */
typedef struct { uint32_t wire; fe coef; /* canonical */ } term_t;
typedef struct
{
    uint32_t n_vars, n_public, n_rows;
    uint64_t *a_ptr, *b_ptr, *c_ptr; /* n_rows + 1 */
    term_t   *a, *b, *c;
    uint64_t  na, nb, nc, cap_a, cap_b, cap_c;
    fe*       w;                      /* canonical witness */
} circuit_t;

static void push_term(term_t** arr, uint64_t* n, uint64_t* cap, uint32_t wire, const fe* coef)
{
    /* merge duplicates inside the current row is handled by the caller where needed */
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *arr = (term_t*)realloc(*arr, *cap * sizeof(term_t)); }
    (*arr)[*n].wire = wire; (*arr)[*n].coef = *coef; (*n)++;
}
static void fe_small(fe* r, uint64_t x) { r->v[0] = x; r->v[1] = r->v[2] = r->v[3] = 0; }
static void fe_neg_small(fe* r, uint64_t x) { fe t; fe_small(&t, x); fe z = {{0,0,0,0}}; f_sub(r, &z, &t, &FR); }

static void circuit_free(circuit_t* c)
{
    if (!c) return;
    free(c->a_ptr); free(c->b_ptr); free(c->c_ptr); free(c->a); free(c->b); free(c->c); free(c->w); free(c);
}

/* Gate mix of the synthetic circuit, as cumulative percentages: XOR below g_mix[0], AND below g_mix[1], byte
 * recomposition below g_mix[2], field product above. The default (50 / 80 / 95) gives the keyless-like witness of the
 * bench (~84 % bits, ~12 % bytes, ~4 % full-width values); kzp_setupgen_set_mix(35, 55, 70) makes ~30 % of the wires
 * full-width field elements — the pessimistic end for the witness-side MSMs. */
static int g_mix[3] = {50, 80, 95};
int kzp_setupgen_set_mix(int xor_below, int and_below, int byte_below)
{
    if (xor_below < 1 || xor_below > and_below || and_below > byte_below || byte_below > 100)
        return -1;
    g_mix[0] = xor_below; g_mix[1] = and_below; g_mix[2] = byte_below;
    return 0;
}

static circuit_t* circuit_synth(uint32_t n_constraints, uint32_t n_vars, uint64_t seed)
{
    circuit_t* C = (circuit_t*)calloc(1, sizeof(circuit_t));
    C->n_vars = n_vars; C->n_public = 1; C->n_rows = 0;
    C->a_ptr = (uint64_t*)calloc((size_t)n_constraints + 1, 8);
    C->b_ptr = (uint64_t*)calloc((size_t)n_constraints + 1, 8);
    C->c_ptr = (uint64_t*)calloc((size_t)n_constraints + 1, 8);
    C->w     = (fe*)calloc(n_vars, sizeof(fe));
    uint32_t* bits   = (uint32_t*)malloc(sizeof(uint32_t) * n_vars); uint32_t n_bits = 0;
    uint32_t* fields = (uint32_t*)malloc(sizeof(uint32_t) * n_vars); uint32_t n_fields = 0;
    rng_t rng; rng_init(&rng, seed);
    fe one, two, three, negone; fe_small(&one, 1); fe_small(&two, 2); fe_small(&three, 3); fe_neg_small(&negone, 1);
    fe_small(&C->w[0], 1);
    for (uint32_t i = 2; i < 10; i++) { fe_small(&C->w[i], rng_below(&rng, 2)); bits[n_bits++] = i; }
#define END_ROW() do { C->n_rows++; C->a_ptr[C->n_rows] = C->na; C->b_ptr[C->n_rows] = C->nb; C->c_ptr[C->n_rows] = C->nc; } while (0)
#define AND_GATE(out) do { uint32_t i_ = bits[rng_below(&rng, n_bits)], j_ = bits[rng_below(&rng, n_bits)];          \
        fe_small(&C->w[out], C->w[i_].v[0] & C->w[j_].v[0]);                                                         \
        push_term(&C->a, &C->na, &C->cap_a, i_, &one); push_term(&C->b, &C->nb, &C->cap_b, j_, &one);                 \
        push_term(&C->c, &C->nc, &C->cap_c, out, &one); END_ROW(); } while (0)
    AND_GATE(1);
    for (uint32_t out = 10; out < n_vars; out++)
    {
        uint64_t t = rng_below(&rng, 100);
        if (t < (uint64_t)g_mix[0] || n_bits < 8)
        {
            /* XOR: 2ij = i + j - out */
            uint32_t i = bits[rng_below(&rng, n_bits)], j = bits[rng_below(&rng, n_bits)];
            fe_small(&C->w[out], C->w[i].v[0] ^ C->w[j].v[0]);
            push_term(&C->a, &C->na, &C->cap_a, i, &two);
            push_term(&C->b, &C->nb, &C->cap_b, j, &one);
            if (i == j) push_term(&C->c, &C->nc, &C->cap_c, i, &two);
            else { push_term(&C->c, &C->nc, &C->cap_c, i, &one); push_term(&C->c, &C->nc, &C->cap_c, j, &one); }
            push_term(&C->c, &C->nc, &C->cap_c, out, &negone);
            END_ROW();
            bits[n_bits++] = out;
        }
        else if (t < (uint64_t)g_mix[1]) { AND_GATE(out); bits[n_bits++] = out; }
        else if (t < (uint64_t)g_mix[2])
        {
            /* byte = sum 2^k bit_k ; duplicates merged like a Python dict keeps insertion order of first use */
            uint32_t ws[8]; uint64_t cs[8]; int cnt = 0; uint64_t val = 0;
            for (int k = 0; k < 8; k++)
            {
                uint32_t i = bits[rng_below(&rng, n_bits)];
                val += C->w[i].v[0] << k;
                int f = -1;
                for (int q = 0; q < cnt; q++) if (ws[q] == i) f = q;
                if (f >= 0) cs[f] += 1ull << k; else { ws[cnt] = i; cs[cnt] = 1ull << k; cnt++; }
            }
            fe_small(&C->w[out], val);
            for (int q = 0; q < cnt; q++) { fe cf; fe_small(&cf, cs[q]); push_term(&C->a, &C->na, &C->cap_a, ws[q], &cf); }
            push_term(&C->b, &C->nb, &C->cap_b, 0, &one);
            push_term(&C->c, &C->nc, &C->cap_c, out, &one);
            END_ROW();
        }
        else
        {
            if (n_fields < 2)
            {
                fe x; rng_fr(&rng, &x);
                C->w[out] = x;
                push_term(&C->a, &C->na, &C->cap_a, 0, &x);
                push_term(&C->b, &C->nb, &C->cap_b, 0, &one);
                push_term(&C->c, &C->nc, &C->cap_c, out, &one);
                END_ROW();
            }
            else
            {
                uint32_t i = fields[rng_below(&rng, n_fields)], j = fields[rng_below(&rng, n_fields)];
                uint32_t k = bits[rng_below(&rng, n_bits)];
                fe l, rr, lm, rm, pm;
                fr_add(&l, &C->w[i], &C->w[k]);
                fr_add(&rr, &C->w[j], &three);
                f_to_mont(&lm, &l, &FR); f_to_mont(&rm, &rr, &FR); fr_mul(&pm, &lm, &rm); f_from_mont(&C->w[out], &pm, &FR);
                push_term(&C->a, &C->na, &C->cap_a, i, &one); push_term(&C->a, &C->na, &C->cap_a, k, &one);
                if (j == 0) { fe four; fe_small(&four, 4); push_term(&C->b, &C->nb, &C->cap_b, 0, &four); }
                else { push_term(&C->b, &C->nb, &C->cap_b, j, &one); push_term(&C->b, &C->nb, &C->cap_b, 0, &three); }
                push_term(&C->c, &C->nc, &C->cap_c, out, &one);
                END_ROW();
            }
            fields[n_fields++] = out;
        }
    }
    while (C->n_rows < n_constraints)
    {
        /* booleanity: b * (b - 1) = 0 */
        uint32_t i = bits[rng_below(&rng, n_bits)];
        push_term(&C->a, &C->na, &C->cap_a, i, &one);
        push_term(&C->b, &C->nb, &C->cap_b, i, &one);
        push_term(&C->b, &C->nb, &C->cap_b, 0, &negone);
        END_ROW();
    }
    free(bits); free(fields);
    return C;
}

/* ------------------------------------------------------------------ fixed-base tables for the generator */
#define FB_BITS 16
#define FB_WINDOWS 16
typedef struct { g1_aff* t; } fb1_t; /* [FB_WINDOWS][1<<FB_BITS] entry d = d * 2^(16 j) * G */
typedef struct { g2_aff* t; } fb2_t;

static void g1_gen(g1_aff* g) { f_from_u64(&g->x, 1, &FQ); f_from_u64(&g->y, 2, &FQ); }
static void fq_from_dec(fe* out, const char* s)
{
    fe acc = {{0,0,0,0}}, ten; f_from_u64(&ten, 10, &FQ);
    for (; *s; s++) { fe d; f_from_u64(&d, (uint64_t)(*s - '0'), &FQ); fq_mul(&acc, &acc, &ten); fq_add(&acc, &acc, &d); }
    *out = acc;
}
static void g2_gen(g2_aff* g)
{
    /* RS/alt_bn128.hpp:47-50 */
    fq_from_dec(&g->x.a, "10857046999023057135944570762232829481370756359578518086990519993285655852781");
    fq_from_dec(&g->x.b, "11559732032986387107991004021392285783925812861821192530917403151452391805634");
    fq_from_dec(&g->y.a, "8495653923123431417604973247489272438418190587263600148770280649306958101930");
    fq_from_dec(&g->y.b, "4082367875863433681332203403145435568316851327593401208105741076214120093531");
}

static fb1_t* fb1_new(void)
{
    fb1_t* T = (fb1_t*)malloc(sizeof(fb1_t));
    size_t per = (size_t)1 << FB_BITS;
    T->t = (g1_aff*)malloc(sizeof(g1_aff) * per * FB_WINDOWS);
    g1_aff g; g1_gen(&g);
    g1_pt bases[FB_WINDOWS]; g1_from_aff(&bases[0], &g);
    for (int j = 1; j < FB_WINDOWS; j++) { bases[j] = bases[j - 1]; for (int k = 0; k < FB_BITS; k++) { g1_pt t = bases[j]; g1_dbl(&bases[j], &t); } }
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < FB_WINDOWS; j++)
    {
        g1_pt* row = (g1_pt*)malloc(sizeof(g1_pt) * per);
        g1_set_inf(&row[0]);
        for (size_t d = 1; d < per; d++) { row[d] = row[d - 1]; g1_add(&row[d], &bases[j]); }
        g1_batch_to_aff(T->t + (size_t)j * per, row, per);
        free(row);
    }
    return T;
}
static fb2_t* fb2_new(void)
{
    fb2_t* T = (fb2_t*)malloc(sizeof(fb2_t));
    size_t per = (size_t)1 << FB_BITS;
    T->t = (g2_aff*)malloc(sizeof(g2_aff) * per * FB_WINDOWS);
    g2_aff g; g2_gen(&g);
    g2_pt bases[FB_WINDOWS]; g2_from_aff(&bases[0], &g);
    for (int j = 1; j < FB_WINDOWS; j++) { bases[j] = bases[j - 1]; for (int k = 0; k < FB_BITS; k++) { g2_pt t = bases[j]; g2_dbl(&bases[j], &t); } }
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < FB_WINDOWS; j++)
    {
        g2_pt* row = (g2_pt*)malloc(sizeof(g2_pt) * per);
        g2_set_inf(&row[0]);
        for (size_t d = 1; d < per; d++) { row[d] = row[d - 1]; g2_add(&row[d], &bases[j]); }
        g2_batch_to_aff(T->t + (size_t)j * per, row, per);
        free(row);
    }
    return T;
}
/* k canonical (non-Montgomery) scalar */
static void fb1_mul(const fb1_t* T, g1_pt* out, const fe* k)
{
    g1_pt acc; g1_set_inf(&acc);
    for (int j = 0; j < FB_WINDOWS; j++)
    {
        uint32_t d = (uint32_t)(k->v[j >> 2] >> (16 * (j & 3))) & 0xffff;
        if (d) g1_madd(&acc, &T->t[((size_t)j << FB_BITS) + d]);
    }
    *out = acc;
}
static void fb2_mul(const fb2_t* T, g2_pt* out, const fe* k)
{
    g2_pt acc; g2_set_inf(&acc);
    for (int j = 0; j < FB_WINDOWS; j++)
    {
        uint32_t d = (uint32_t)(k->v[j >> 2] >> (16 * (j & 3))) & 0xffff;
        if (d) g2_madd(&acc, &T->t[((size_t)j << FB_BITS) + d]);
    }
    *out = acc;
}

/* scalars (Montgomery Fr) -> affine G1 points in zkey byte layout; zero scalar -> 64 zero bytes (infinity) */
static void gen_g1_points(const fb1_t* T, const fe* k_mont, size_t n, uint8_t* out)
{
    const size_t blk = 4096;
#pragma omp parallel
    {
        g1_pt* tmp = (g1_pt*)malloc(sizeof(g1_pt) * blk);
#pragma omp for schedule(dynamic, 1)
        for (size_t b0 = 0; b0 < n; b0 += blk)
        {
            size_t cnt = n - b0 < blk ? n - b0 : blk;
            for (size_t i = 0; i < cnt; i++) { fe k; f_from_mont(&k, &k_mont[b0 + i], &FR); fb1_mul(T, &tmp[i], &k); }
            g1_batch_to_aff((g1_aff*)(out + b0 * 64), tmp, cnt);
        }
        free(tmp);
    }
}
static void gen_g2_points(const fb2_t* T, const fe* k_mont, size_t n, uint8_t* out)
{
    const size_t blk = 2048;
#pragma omp parallel
    {
        g2_pt* tmp = (g2_pt*)malloc(sizeof(g2_pt) * blk);
#pragma omp for schedule(dynamic, 1)
        for (size_t b0 = 0; b0 < n; b0 += blk)
        {
            size_t cnt = n - b0 < blk ? n - b0 : blk;
            for (size_t i = 0; i < cnt; i++) { fe k; f_from_mont(&k, &k_mont[b0 + i], &FR); fb2_mul(T, &tmp[i], &k); }
            g2_batch_to_aff((g2_aff*)(out + b0 * 128), tmp, cnt);
        }
        free(tmp);
    }
}

/* batch inversion in Fr (Montgomery) */
static void fr_batch_inv(fe* x, size_t n)
{
    int nt = omp_get_max_threads();
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nt; b++)
    {
        size_t lo = n * (size_t)b / nt, hi = n * (size_t)(b + 1) / nt;
        if (lo >= hi) continue;
        fe* pre = (fe*)malloc(sizeof(fe) * (hi - lo));
        fe  run = FR.one;
        for (size_t i = lo; i < hi; i++) { fr_mul(&run, &run, &x[i]); pre[i - lo] = run; }
        fe inv; f_inv(&inv, &run, &FR);
        for (size_t i = hi; i-- > lo;)
        {
            fe t;
            if (i > lo) fr_mul(&t, &inv, &pre[i - lo - 1]); else t = inv;
            fr_mul(&inv, &inv, &x[i]);
            x[i] = t;
        }
        free(pre);
    }
}

/* ------------------------------------------------------------------ file writers */
static void wr_u32(FILE* f, uint32_t v) { fwrite(&v, 4, 1, f); }
static void wr_u64(FILE* f, uint64_t v) { fwrite(&v, 8, 1, f); }
static void wr_section(FILE* f, uint32_t id, const void* data, uint64_t size)
{
    wr_u32(f, id); wr_u64(f, size);
    if (size) fwrite(data, 1, size, f);
}

/* Generates circuit + witness + trapdoor setup; writes <zkey_path>, <wtns_path>. Returns 0 on success.
 * out_info (optional, 8 x u64): n_vars, n_public, domain, n_coefs, n_rows, public input (low limb), 0, 0.
 * trapdoor_out (optional): tau, alpha, beta, gamma, delta canonical (5 x 32 bytes). */
int kzp_setupgen(uint32_t n_constraints, uint32_t n_vars, uint64_t seed, const char* zkey_path,
                        const char* wtns_path, uint64_t* out_info, uint8_t* trapdoor_out)
{
    circuit_t* C = circuit_synth(n_constraints, n_vars, seed);
    uint32_t m = C->n_rows, np = C->n_public;
    uint64_t n = 1; uint32_t lg = 0;
    while (n < (uint64_t)m + np + 1) { n <<= 1; lg++; }
    rng_t rng; rng_init(&rng, seed ^ 0x5EEDull);
    fe trap[5];
    for (int i = 0; i < 5; i++) { rng_fr(&rng, &trap[i]); if (fe_is_zero(&trap[i])) fe_small(&trap[i], 1); }
    if (trapdoor_out) memcpy(trapdoor_out, trap, 160);
    fe tau, alpha, beta, gamma, delta;
    f_to_mont(&tau, &trap[0], &FR); f_to_mont(&alpha, &trap[1], &FR); f_to_mont(&beta, &trap[2], &FR);
    f_to_mont(&gamma, &trap[3], &FR); f_to_mont(&delta, &trap[4], &FR);

    /* L_j(tau) = (tau^n - 1) w^j / (n (tau - w^j)), j < m + np + 1 */
    size_t nl = (size_t)m + np + 1;
    fe omega; fr_root_of_unity(&omega, lg);
    fe* wj  = (fe*)malloc(sizeof(fe) * nl);
    fe* den = (fe*)malloc(sizeof(fe) * nl);
    {
        int nt = omp_get_max_threads();
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nt; b++)
        {
            size_t lo = nl * (size_t)b / nt, hi = nl * (size_t)(b + 1) / nt;
            if (lo >= hi) continue;
            uint64_t e[4] = {lo, 0, 0, 0};
            fe cur; f_pow(&cur, &omega, e, &FR);
            for (size_t j = lo; j < hi; j++) { wj[j] = cur; fr_sub(&den[j], &tau, &cur); fr_mul(&cur, &cur, &omega); }
        }
    }
    fr_batch_inv(den, nl);
    fe tn1, ninv, nn;
    { uint64_t e[4] = {n, 0, 0, 0}; f_pow(&tn1, &tau, e, &FR); fr_sub(&tn1, &tn1, &FR.one); }
    f_from_u64(&nn, n, &FR); f_inv(&ninv, &nn, &FR);
    fe* L = (fe*)malloc(sizeof(fe) * nl);
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < nl; j++) { fe t; fr_mul(&t, &tn1, &wj[j]); fr_mul(&t, &t, &ninv); fr_mul(&L[j], &t, &den[j]); }
    free(wj); free(den);

    uint32_t nv = C->n_vars;
    fe* At = (fe*)calloc(nv, sizeof(fe)); fe* Bt = (fe*)calloc(nv, sizeof(fe)); fe* Ct = (fe*)calloc(nv, sizeof(fe));
    uint64_t n_coefs = C->na + C->nb + np + 1;
    uint8_t* sec4 = (uint8_t*)malloc(4 + n_coefs * 44);
    { uint32_t nc32 = (uint32_t)n_coefs; memcpy(sec4, &nc32, 4); }
    uint8_t* cp = sec4 + 4;
    fe r2 = FR.r2;
    for (uint32_t j = 0; j < m; j++)
    {
        for (uint64_t e = C->a_ptr[j]; e < C->a_ptr[j + 1]; e++)
        {
            fe cm, t; f_to_mont(&cm, &C->a[e].coef, &FR); fr_mul(&t, &cm, &L[j]); fr_add(&At[C->a[e].wire], &At[C->a[e].wire], &t);
            uint32_t mm = 0; fe cr2; fr_mul(&cr2, &cm, &r2); /* value * R^2 */
            memcpy(cp, &mm, 4); memcpy(cp + 4, &j, 4); memcpy(cp + 8, &C->a[e].wire, 4); memcpy(cp + 12, &cr2, 32); cp += 44;
        }
        for (uint64_t e = C->b_ptr[j]; e < C->b_ptr[j + 1]; e++)
        {
            fe cm, t; f_to_mont(&cm, &C->b[e].coef, &FR); fr_mul(&t, &cm, &L[j]); fr_add(&Bt[C->b[e].wire], &Bt[C->b[e].wire], &t);
            uint32_t mm = 1; fe cr2; fr_mul(&cr2, &cm, &r2);
            memcpy(cp, &mm, 4); memcpy(cp + 4, &j, 4); memcpy(cp + 8, &C->b[e].wire, 4); memcpy(cp + 12, &cr2, 32); cp += 44;
        }
        for (uint64_t e = C->c_ptr[j]; e < C->c_ptr[j + 1]; e++)
        {
            fe cm, t; f_to_mont(&cm, &C->c[e].coef, &FR); fr_mul(&t, &cm, &L[j]); fr_add(&Ct[C->c[e].wire], &Ct[C->c[e].wire], &t);
        }
    }
    for (uint32_t s = 0; s <= np; s++)
    {
        fr_add(&At[s], &At[s], &L[m + s]);
        uint32_t mm = 0, row = m + s; fe cr2; fr_mul(&cr2, &FR.one, &r2);
        memcpy(cp, &mm, 4); memcpy(cp + 4, &row, 4); memcpy(cp + 8, &s, 4); memcpy(cp + 12, &cr2, 32); cp += 44;
    }
    free(L);

    fb1_t* T1 = fb1_new();
    fb2_t* T2 = fb2_new();
    fe ginv, dinv; f_inv(&ginv, &gamma, &FR); f_inv(&dinv, &delta, &FR);

    /* header section 2 */
    uint8_t hdr[4 + 32 + 4 + 32 + 12 + 64 * 3 + 128 * 3]; size_t hp = 0;
    { uint32_t n8 = 32; memcpy(hdr + hp, &n8, 4); hp += 4; memcpy(hdr + hp, FQ.p, 32); hp += 32;
      memcpy(hdr + hp, &n8, 4); hp += 4; memcpy(hdr + hp, FR.p, 32); hp += 32;
      uint32_t dom = (uint32_t)n; memcpy(hdr + hp, &nv, 4); memcpy(hdr + hp + 4, &np, 4); memcpy(hdr + hp + 8, &dom, 4); hp += 12; }
    gen_g1_points(T1, &alpha, 1, hdr + hp); hp += 64;
    gen_g1_points(T1, &beta, 1, hdr + hp); hp += 64;
    gen_g2_points(T2, &beta, 1, hdr + hp); hp += 128;
    gen_g2_points(T2, &gamma, 1, hdr + hp); hp += 128;
    gen_g1_points(T1, &delta, 1, hdr + hp); hp += 64;
    gen_g2_points(T2, &delta, 1, hdr + hp); hp += 128;

    /* IC and C scalars */
    fe* kc = (fe*)malloc(sizeof(fe) * nv);
#pragma omp parallel for schedule(static)
    for (uint32_t s = 0; s < nv; s++)
    {
        fe t, u;
        fr_mul(&t, &beta, &At[s]); fr_mul(&u, &alpha, &Bt[s]); fr_add(&t, &t, &u); fr_add(&t, &t, &Ct[s]);
        fr_mul(&kc[s], &t, s <= np ? &ginv : &dinv);
    }
    uint8_t* sec3 = (uint8_t*)malloc((size_t)(np + 1) * 64);
    gen_g1_points(T1, kc, np + 1, sec3);
    uint8_t* sec5 = (uint8_t*)malloc((size_t)nv * 64);
    uint8_t* sec6 = (uint8_t*)malloc((size_t)nv * 64);
    uint8_t* sec7 = (uint8_t*)malloc((size_t)nv * 128);
    uint8_t* sec8 = (uint8_t*)malloc((size_t)(nv - np - 1) * 64 + 64);
    gen_g1_points(T1, At, nv, sec5);
    gen_g1_points(T1, Bt, nv, sec6);
    gen_g2_points(T2, Bt, nv, sec7);
    gen_g1_points(T1, kc + np + 1, nv - np - 1, sec8);
    free(kc);

    /* H[i] = L^(2n)_{2i+1}(tau) / delta */
    fe w2n; fr_root_of_unity(&w2n, lg + 1);
    fe* kh = (fe*)malloc(sizeof(fe) * n);
    fe* wk = (fe*)malloc(sizeof(fe) * n);
    {
        fe w2n_sq; fr_mul(&w2n_sq, &w2n, &w2n);
        int nt = omp_get_max_threads();
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nt; b++)
        {
            size_t lo = n * (size_t)b / nt, hi = n * (size_t)(b + 1) / nt;
            if (lo >= hi) continue;
            uint64_t e[4] = {2 * lo + 1, 0, 0, 0};
            fe cur; f_pow(&cur, &w2n, e, &FR);
            for (size_t i = lo; i < hi; i++) { wk[i] = cur; fr_sub(&kh[i], &tau, &cur); fr_mul(&cur, &cur, &w2n_sq); }
        }
    }
    fr_batch_inv(kh, n);
    fe t2n1, inv2n, n2;
    { uint64_t e[4] = {2 * n, 0, 0, 0}; f_pow(&t2n1, &tau, e, &FR); fr_sub(&t2n1, &t2n1, &FR.one); }
    f_from_u64(&n2, 2 * n, &FR); f_inv(&inv2n, &n2, &FR);
    fe cst; fr_mul(&cst, &t2n1, &inv2n); fr_mul(&cst, &cst, &dinv);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { fe t; fr_mul(&t, &cst, &wk[i]); fr_mul(&kh[i], &t, &kh[i]); }
    free(wk);
    uint8_t* sec9 = (uint8_t*)malloc((size_t)n * 64);
    gen_g1_points(T1, kh, n, sec9);
    free(kh);

    FILE* f = fopen(zkey_path, "wb");
    if (!f) return 1;
    fwrite("zkey", 1, 4, f); wr_u32(f, 1); wr_u32(f, 10);
    { uint32_t proto = 1; wr_section(f, 1, &proto, 4); }
    wr_section(f, 2, hdr, hp);
    wr_section(f, 3, sec3, (uint64_t)(np + 1) * 64);
    wr_section(f, 4, sec4, 4 + n_coefs * 44);
    wr_section(f, 5, sec5, (uint64_t)nv * 64);
    wr_section(f, 6, sec6, (uint64_t)nv * 64);
    wr_section(f, 7, sec7, (uint64_t)nv * 128);
    wr_section(f, 8, sec8, (uint64_t)(nv - np - 1) * 64);
    wr_section(f, 9, sec9, (uint64_t)n * 64);
    { uint32_t zero = 0; wr_section(f, 10, &zero, 4); }
    fclose(f);

    f = fopen(wtns_path, "wb");
    if (!f) return 1;
    fwrite("wtns", 1, 4, f); wr_u32(f, 2); wr_u32(f, 2);
    { uint8_t h[40]; uint32_t n8 = 32; memcpy(h, &n8, 4); memcpy(h + 4, FR.p, 32); memcpy(h + 36, &nv, 4); wr_section(f, 1, h, 40); }
    wr_section(f, 2, C->w, (uint64_t)nv * 32);
    fclose(f);

    if (out_info)
    {
        out_info[0] = nv; out_info[1] = np; out_info[2] = n; out_info[3] = n_coefs; out_info[4] = m;
        out_info[5] = C->w[1].v[0]; out_info[6] = out_info[7] = 0;
    }
    free(sec3); free(sec4); free(sec5); free(sec6); free(sec7); free(sec8); free(sec9);
    free(At); free(Bt); free(Ct); free(T1->t); free(T1); free(T2->t); free(T2);
    circuit_free(C);
    return 0;
}


/* ------------------------------------------------------------------ micro-benchmark inputs
 * Bases P_i = (s0 + i) * G for i < n ("k*G distinct"), written in zkey byte layout (affine Montgomery), and the
 * closed-form answer of an MSM over them: sum_i k_i P_i = (sum_i k_i (s0 + i) mod r) * G, which pins a 2^24-point MSM
 * with ONE fixed-base multiplication instead of a CPU MSM. */
static fb1_t* g_fb1 = NULL;
static fb2_t* g_fb2 = NULL;
static void   fb_tables(int group)
{
#pragma omp critical(kzp_fb_tables)
    {
        if (group == 0 && !g_fb1) g_fb1 = fb1_new();
        if (group == 1 && !g_fb2) g_fb2 = fb2_new();
    }
}

/* group 0 = G1 (64 B/point), 1 = G2 (128 B/point); s0: canonical 32-byte scalar, s0 + n < r assumed */
int kzp_gen_consecutive_points(int group, uint64_t n, const uint8_t* s0_32, uint8_t* out)
{
    if (group < 0 || group > 1) return -1;
    fb_tables(group);
    const size_t blk = 4096;
    fe s0; memcpy(&s0, s0_32, 32);
#pragma omp parallel
    {
        void* tmp = malloc((group == 0 ? sizeof(g1_pt) : sizeof(g2_pt)) * blk);
#pragma omp for schedule(dynamic, 1)
        for (size_t b0 = 0; b0 < n; b0 += blk)
        {
            size_t cnt = n - b0 < blk ? n - b0 : blk;
            fe k = s0; /* k = s0 + b0 (plain integers, no reduction needed below r) */
            unsigned __int128 c = (unsigned __int128)k.v[0] + b0; k.v[0] = (uint64_t)c; c >>= 64;
            for (int j = 1; j < 4; j++) { c += k.v[j]; k.v[j] = (uint64_t)c; c >>= 64; }
            if (group == 0)
            {
                g1_pt* t = (g1_pt*)tmp; g1_aff g; g1_gen(&g);
                fb1_mul(g_fb1, &t[0], &k);
                for (size_t i = 1; i < cnt; i++) { t[i] = t[i - 1]; g1_madd(&t[i], &g); }
                g1_batch_to_aff((g1_aff*)(out + b0 * 64), t, cnt);
            }
            else
            {
                g2_pt* t = (g2_pt*)tmp; g2_aff g; g2_gen(&g);
                fb2_mul(g_fb2, &t[0], &k);
                for (size_t i = 1; i < cnt; i++) { t[i] = t[i - 1]; g2_madd(&t[i], &g); }
                g2_batch_to_aff((g2_aff*)(out + b0 * 128), t, cnt);
            }
        }
        free(tmp);
    }
    return 0;
}

/* out: affine CANONICAL coordinates (64 / 128 bytes, zeros for infinity) of (sum_i k_i (s0 + i) mod r) * G;
 * scalars: n x 32-byte canonical integers < r */
int kzp_msm_closed_form(int group, uint64_t n, const uint8_t* s0_32, const uint8_t* scalars, uint8_t* out)
{
    if (group < 0 || group > 1) return -1;
    fb_tables(group);
    fe s0m; { fe s0; memcpy(&s0, s0_32, 32); f_to_mont(&s0m, &s0, &FR); }
    int nt = omp_get_max_threads();
    fe* part = (fe*)calloc((size_t)nt, sizeof(fe));
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nt; b++)
    {
        size_t lo = n * (size_t)b / nt, hi = n * (size_t)(b + 1) / nt;
        fe acc = {{0, 0, 0, 0}}, idx, one = FR.one;
        f_from_u64(&idx, lo, &FR); fr_add(&idx, &idx, &s0m);
        for (size_t i = lo; i < hi; i++)
        {
            fe k, km, t; memcpy(&k, scalars + i * 32, 32);
            f_to_mont(&km, &k, &FR); fr_mul(&t, &km, &idx); fr_add(&acc, &acc, &t); fr_add(&idx, &idx, &one);
        }
        part[b] = acc;
    }
    fe tot = {{0, 0, 0, 0}};
    for (int b = 0; b < nt; b++) fr_add(&tot, &tot, &part[b]);
    free(part);
    fe k; f_from_mont(&k, &tot, &FR);
    if (group == 0)
    {
        g1_pt p; g1_aff a; fb1_mul(g_fb1, &p, &k); g1_to_aff(&a, &p);
        fe x, y; f_from_mont(&x, &a.x, &FQ); f_from_mont(&y, &a.y, &FQ);
        memcpy(out, &x, 32); memcpy(out + 32, &y, 32);
    }
    else
    {
        g2_pt p; g2_aff a; fb2_mul(g_fb2, &p, &k); g2_to_aff(&a, &p);
        fe c[4]; f_from_mont(&c[0], &a.x.a, &FQ); f_from_mont(&c[1], &a.x.b, &FQ);
        f_from_mont(&c[2], &a.y.a, &FQ); f_from_mont(&c[3], &a.y.b, &FQ);
        memcpy(out, c, 128);
    }
    return 0;
}
