// Host-side BN254 field with 4 x 64-bit limbs (unsigned __int128 products), same static interface as
// kzp::Fp so the group-law templates in ec.cuh instantiate over it. Used only by the thin host layer:
// proof assembly (6 scalar multiplications + a handful of additions, groth16.cpp:328-357 in the
// reference), affine conversion and decimal printing (fq.cpp toString), never for MSM/NTT work.
// Same byte layout as the device type (little-endian limbs, Montgomery R = 2^256).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>

#include "ff.cuh"

namespace kzp
{

template <class P>
struct alignas(16) HostFp
{
    uint64_t v[4];

    static constexpr bool kFusedMulAdd2 = false;

    static uint64_t p(int i)
    {
        return (uint64_t)modulus_limb<P>(2 * i) | ((uint64_t)modulus_limb<P>(2 * i + 1) << 32);
    }
    static uint64_t np64()
    {
        // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
        uint64_t p0  = p(0);
        uint64_t inv = 1;
        for (int i = 0; i < 6; i++)
            inv *= 2 - p0 * inv;
        return (uint64_t)0 - inv;
    }

    static HostFp zero()
    {
        HostFp r;
        r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0;
        return r;
    }
    static HostFp one()
    {
        HostFp r;
        for (int i = 0; i < 4; i++)
            r.v[i] = (uint64_t)P::R1[2 * i] | ((uint64_t)P::R1[2 * i + 1] << 32);
        return r;
    }
    static HostFp r2()
    {
        HostFp r;
        for (int i = 0; i < 4; i++)
            r.v[i] = (uint64_t)P::R2[2 * i] | ((uint64_t)P::R2[2 * i + 1] << 32);
        return r;
    }
    static bool is_zero(const HostFp& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
    static bool eq(const HostFp& a, const HostFp& b)
    {
        return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0;
    }
    static bool geq_p(const HostFp& a)
    {
        for (int i = 3; i >= 0; i--)
        {
            if (a.v[i] > p(i))
                return true;
            if (a.v[i] < p(i))
                return false;
        }
        return true;
    }
    static void sub_p(HostFp& a)
    {
        unsigned __int128 bw = 0;
        for (int i = 0; i < 4; i++)
        {
            unsigned __int128 d = (unsigned __int128)a.v[i] - p(i) - (uint64_t)bw;
            a.v[i]              = (uint64_t)d;
            bw                  = (d >> 64) & 1;
        }
    }
    // branch-free: the reduction decision is data dependent and mispredicts half of the time otherwise. (The
    // _addcarry_u64 / _subborrow_u64 intrinsics are 3x faster in an isolated loop but make the pairing 40 % slower
    // inside the library build, so the plain unsigned __int128 form stays.)
    static void add(HostFp& r, const HostFp& a, const HostFp& b)
    {
        uint64_t          s[4], t[4];
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++)
        {
            c += (unsigned __int128)a.v[i] + b.v[i];
            s[i] = (uint64_t)c;
            c >>= 64;
        }
        uint64_t bw = 0;
        for (int i = 0; i < 4; i++)
        {
            unsigned __int128 d = (unsigned __int128)s[i] - p(i) - bw;
            t[i]                = (uint64_t)d;
            bw                  = (uint64_t)(d >> 64) & 1;
        }
        uint64_t take = (uint64_t)0 - ((uint64_t)c | (bw ^ 1));
        for (int i = 0; i < 4; i++)
            r.v[i] = (t[i] & take) | (s[i] & ~take);
    }
    static void sub(HostFp& r, const HostFp& a, const HostFp& b)
    {
        uint64_t s[4];
        uint64_t bw = 0;
        for (int i = 0; i < 4; i++)
        {
            unsigned __int128 d = (unsigned __int128)a.v[i] - b.v[i] - bw;
            s[i]                = (uint64_t)d;
            bw                  = (uint64_t)(d >> 64) & 1;
        }
        uint64_t          mask = (uint64_t)0 - bw;
        unsigned __int128 c    = 0;
        for (int i = 0; i < 4; i++)
        {
            c += (unsigned __int128)s[i] + (p(i) & mask);
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    static void neg(HostFp& r, const HostFp& a)
    {
        if (is_zero(a))
        {
            r = a;
            return;
        }
        HostFp z = zero();
        sub(r, z, a);
    }
    static void dbl(HostFp& r, const HostFp& a) { add(r, a, a); }

    static void mul(HostFp& r, const HostFp& a, const HostFp& b)
    {
        static const uint64_t np = np64();
        uint64_t              t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++)
        {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; j++)
            {
                c += (unsigned __int128)a.v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4]       = (uint64_t)c;
            t[5]       = (uint64_t)(c >> 64);
            uint64_t m = t[0] * np;
            c          = ((unsigned __int128)m * p(0) + t[0]) >> 64;
            for (int j = 1; j < 4; j++)
            {
                c += (unsigned __int128)m * p(j) + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        HostFp s;
        for (int i = 0; i < 4; i++)
            s.v[i] = t[i];
        if (t[4] || geq_p(s))
            sub_p(s);
        r = s;
    }
    static void sqr(HostFp& r, const HostFp& a) { mul(r, a, a); }
    static void to_mont(HostFp& r, const HostFp& a)
    {
        HostFp k = r2();
        mul(r, a, k);
    }
    static void from_mont(HostFp& r, const HostFp& a)
    {
        HostFp k = zero();
        k.v[0]   = 1;
        mul(r, a, k);
    }
    static void inv(HostFp& r, const HostFp& a)
    {
        uint64_t e[4];
        for (int i = 0; i < 4; i++)
            e[i] = p(i);
        e[0] -= 2;
        HostFp acc = one();
        for (int i = 255; i >= 0; i--)
        {
            sqr(acc, acc);
            if ((e[i >> 6] >> (i & 63)) & 1)
                mul(acc, acc, a);
        }
        r = acc;
    }

    // decimal string of the canonical value of a Montgomery element (RawFq::toString, fr.cpp:225-236)
    static std::string to_decimal(const HostFp& mont)
    {
        HostFp c;
        from_mont(c, mont);
        uint32_t w[8];
        memcpy(w, c.v, 32);
        char buf[80];
        int  n = 0;
        bool nz;
        do
        {
            // divide the 256-bit number by 10^9
            uint64_t rem = 0;
            nz           = false;
            for (int i = 7; i >= 0; i--)
            {
                uint64_t cur = (rem << 32) | w[i];
                w[i]         = (uint32_t)(cur / 1000000000ull);
                rem          = cur % 1000000000ull;
                nz |= w[i] != 0;
            }
            for (int k = 0; k < 9; k++)
            {
                buf[n++] = (char)('0' + rem % 10);
                rem /= 10;
                if (!nz && rem == 0)
                    break;
            }
        } while (nz);
        std::string s(buf, buf + n);
        return std::string(s.rbegin(), s.rend());
    }
};

typedef HostFp<FqParams> HFq;
typedef HostFp<FrParams> HFr;

} // namespace kzp
