// kzp_host_pairing_check / kzp_host_verify: host-only Groth16 verification next to the prover (SURVEY.md §8(f).3;
// pairing.hpp explains the algorithm and what it stands in for in the reference service).
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kzp_b200.h"
#include "binfile.hpp"
#include "pairing.hpp"

using namespace kzp;
using namespace kzp::pairing;

static thread_local std::string g_verify_error;
const char*                     kzp_verify_last_error(void) { return g_verify_error.c_str(); }

typedef XyzzT<HFq> HG1;

static G1Aff load_g1(const uint8_t* p)
{
    G1Aff a;
    memcpy(a.x.v, p, 32);
    memcpy(a.y.v, p + 32, 32);
    a.inf = HFq::is_zero(a.x) && HFq::is_zero(a.y);
    return a;
}
static G2Aff load_g2(const uint8_t* p)
{
    G2Aff a;
    memcpy(a.x.a.v, p, 32);
    memcpy(a.x.b.v, p + 32, 32);
    memcpy(a.y.a.v, p + 64, 32);
    memcpy(a.y.b.v, p + 96, 32);
    a.inf = F2::is_zero(a.x) && F2::is_zero(a.y);
    return a;
}

// decimal string -> Montgomery Fq; false when not a number or >= q
static bool parse_fq(const std::string& s, HFq& out)
{
    if (s.empty() || s.size() > 78)
        return false;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (char c : s)
    {
        if (c < '0' || c > '9')
            return false;
        uint64_t carry = (uint64_t)(c - '0');
        for (int i = 0; i < 8; i++)
        {
            uint64_t cur = (uint64_t)w[i] * 10u + carry;
            w[i]         = (uint32_t)cur;
            carry        = cur >> 32;
        }
        if (carry)
            return false;
    }
    HFq c;
    memcpy(c.v, w, 32);
    if (HFq::geq_p(c))
        return false;
    HFq::to_mont(out, c);
    return true;
}

// the next `count` quoted strings after `key`
static bool strings_after(const std::string& js, const char* key, int count, std::vector<std::string>& out)
{
    size_t pos = js.find(std::string("\"") + key + "\"");
    if (pos == std::string::npos)
        return false;
    pos += strlen(key) + 2;
    for (int i = 0; i < count; i++)
    {
        size_t a = js.find('"', pos);
        if (a == std::string::npos)
            return false;
        size_t b = js.find('"', a + 1);
        if (b == std::string::npos)
            return false;
        out.push_back(js.substr(a + 1, b - a - 1));
        pos = b + 1;
    }
    return true;
}

int kzp_host_pairing_check(const uint8_t* g1, const uint8_t* g2, int n, int* result_out)
{
    if (!g1 || !g2 || n < 0 || !result_out)
        return KZP_ERR_FORMAT;
    std::vector<G1Aff> Ps;
    std::vector<G2Aff> Qs;
    for (int i = 0; i < n; i++)
    {
        Ps.push_back(load_g1(g1 + 64 * (size_t)i));
        Qs.push_back(load_g2(g2 + 128 * (size_t)i));
        if (!g1_on_curve(Ps.back()) || !g2_on_curve(Qs.back()))
        {
            g_verify_error = "point " + std::to_string(i) + " is not on the curve";
            return KZP_ERR_FORMAT;
        }
    }
    *result_out = pairing_product_is_one(Ps, Qs) ? 1 : 0;
    return KZP_OK;
}

int kzp_host_verify(const char* zkey_path, const char* proof_json, const uint8_t* public32, uint32_t n_public,
                    int* valid_out)
{
    if (!zkey_path || !proof_json || !valid_out || (n_public && !public32))
        return KZP_ERR_FORMAT;
    *valid_out = 0;
    try
    {
        MappedFile file(zkey_path);
        BinView    bin(file.data(), file.size(), "zkey", 1);
        ZkeyHeader zh = parse_zkey(bin);
        if (zh.n_public != n_public)
        {
            g_verify_error = "expected " + std::to_string(zh.n_public) + " public signals";
            return KZP_ERR_FORMAT;
        }
        const Section& ic = bin.section(3);
        if (ic.size < 64ull * ((uint64_t)n_public + 1))
            throw FormatError("zkey IC section too short");

        // proof points (Proof::toJson layout, groth16.cpp:379-410): pi_a [x, y, 1], pi_b [[x.a, x.b], [y.a, y.b], [1, 0]]
        std::string              js(proof_json);
        std::vector<std::string> sa, sb, sc;
        if (!strings_after(js, "pi_a", 2, sa) || !strings_after(js, "pi_b", 4, sb) || !strings_after(js, "pi_c", 2, sc))
        {
            g_verify_error = "proof JSON lacks pi_a / pi_b / pi_c";
            return KZP_ERR_FORMAT;
        }
        G1Aff A, C;
        G2Aff B;
        if (!parse_fq(sa[0], A.x) || !parse_fq(sa[1], A.y) || !parse_fq(sc[0], C.x) || !parse_fq(sc[1], C.y) ||
            !parse_fq(sb[0], B.x.a) || !parse_fq(sb[1], B.x.b) || !parse_fq(sb[2], B.y.a) || !parse_fq(sb[3], B.y.b))
        {
            g_verify_error = "proof coordinate is not a canonical field element";
            return KZP_ERR_FORMAT;
        }
        A.inf = HFq::is_zero(A.x) && HFq::is_zero(A.y);
        C.inf = HFq::is_zero(C.x) && HFq::is_zero(C.y);
        B.inf = F2::is_zero(B.x) && F2::is_zero(B.y);
        if (!g1_on_curve(A) || !g1_on_curve(C) || !g2_on_curve(B))
            return KZP_OK; // a well-formed proof with points off the curve is simply invalid

        // L = IC_0 + sum_i pub_i IC_{i+1}
        HG1 L;
        HG1::set_inf(L);
        {
            HG1::Affine p0;
            memcpy(&p0, ic.data, 64);
            HG1::madd(L, p0);
        }
        for (uint32_t i = 0; i < n_public; i++)
        {
            HFr k;
            memcpy(k.v, public32 + 32ull * i, 32);
            if (HFr::geq_p(k))
            {
                g_verify_error = "public signal is not below r";
                return KZP_ERR_FORMAT;
            }
            HG1::Affine base;
            memcpy(&base, ic.data + 64ull * (i + 1), 64);
            HG1 acc;
            HG1::set_inf(acc);
            for (int b = 253; b >= 0; b--)
            {
                HG1 t = acc;
                HG1::dbl(acc, t);
                if ((k.v[b >> 6] >> (b & 63)) & 1)
                    HG1::madd(acc, base);
            }
            HG1::add(L, acc);
        }
        HG1::Affine la;
        HG1::to_affine(la, L);
        G1Aff Lp;
        Lp.x   = la.x;
        Lp.y   = la.y;
        Lp.inf = HG1::is_inf(L);

        G1Aff nA = A;
        HFq::neg(nA.y, A.y);
        std::vector<G1Aff> Ps = {nA, load_g1(zh.alpha1), Lp, C};
        std::vector<G2Aff> Qs = {B, load_g2(zh.beta2), load_g2(zh.gamma2), load_g2(zh.delta2)};
        *valid_out            = pairing_product_is_one(Ps, Qs) ? 1 : 0;
        return KZP_OK;
    }
    catch (const LoadError& e)
    {
        g_verify_error = e.what();
        return KZP_ERR_IO;
    }
    catch (const std::exception& e)
    {
        g_verify_error = e.what();
        return KZP_ERR_FORMAT;
    }
    catch (...)
    {
        g_verify_error = "unknown exception";
        return KZP_ERR_FORMAT;
    }
}
