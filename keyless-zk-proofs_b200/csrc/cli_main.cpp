// kzp_prove — command-line prover over the C ABI.
//
//   kzp_prove <circuit.zkey> <witness.wtns> <proof.json> <public.json> [--device N] [--repeat K]
//
// Same contract as upstream rapidsnark's `prover` tool, which this fork of the reference removed (its build list
// rust-rapidsnark/rapidsnark/src/test.txt:17 still names main.cpp): proof.json is the compact proof the library
// returns (Proof::toJson, groth16.cpp:379-410) and public.json is the JSON array of the public signals
// witness[1 .. nPublic] as base-10 strings. Exit status 0 on success; 1 usage, 2 zkey not usable, 3 prove failed,
// 4 output not writable. Plain C++ host code: all proving happens in libkzp_b200.so on the GPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kzp_b200.h"

// 256-bit little-endian integer -> decimal
static std::string to_decimal(const uint8_t* le32)
{
    uint32_t w[8];
    for (int i = 0; i < 8; i++)
        w[i] = (uint32_t)le32[4 * i] | ((uint32_t)le32[4 * i + 1] << 8) | ((uint32_t)le32[4 * i + 2] << 16) |
               ((uint32_t)le32[4 * i + 3] << 24);
    std::string out;
    for (;;)
    {
        uint64_t rem  = 0;
        bool     zero = true;
        for (int i = 7; i >= 0; i--)
        {
            uint64_t cur = (rem << 32) | w[i];
            w[i]         = (uint32_t)(cur / 1000000000u);
            rem          = cur % 1000000000u;
            if (w[i])
                zero = false;
        }
        char buf[16];
        if (zero)
        {
            snprintf(buf, sizeof buf, "%u", (unsigned)rem);
            out.insert(0, buf);
            break;
        }
        snprintf(buf, sizeof buf, "%09u", (unsigned)rem);
        out.insert(0, buf);
    }
    return out;
}

static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// public signals from the witness file (format: SURVEY.md Appendix A; wtns_utils.hpp:28-43, binfile_utils.cpp:13-58)
static bool read_public(const char* path, uint32_t n_public, std::vector<std::string>& out, std::string& why)
{
    FILE* f = fopen(path, "rb");
    if (!f)
    {
        why = "cannot open witness file";
        return false;
    }
    uint8_t hdr[12];
    if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "wtns", 4) != 0)
    {
        fclose(f);
        why = "not a wtns file";
        return false;
    }
    uint32_t n_sections = rd32(hdr + 8);
    bool     found      = false;
    for (uint32_t s = 0; s < n_sections && !found; s++)
    {
        uint8_t sh[12];
        if (fread(sh, 1, 12, f) != 12)
            break;
        uint32_t id   = rd32(sh);
        uint64_t size = rd64(sh + 4);
        if (id == 2)
        {
            if (size < 32ull * (n_public + 1))
                break;
            std::vector<uint8_t> v(32ull * (n_public + 1));
            if (fread(v.data(), 1, v.size(), f) != v.size())
                break;
            for (uint32_t i = 1; i <= n_public; i++)
                out.push_back(to_decimal(v.data() + 32ull * i));
            found = true;
        }
        else if (fseek(f, (long)size, SEEK_CUR) != 0)
            break;
    }
    fclose(f);
    if (!found)
        why = "witness section missing or shorter than nPublic + 1 values";
    return found;
}

static bool write_file(const char* path, const std::string& s)
{
    FILE* f = fopen(path, "wb");
    if (!f)
        return false;
    bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    return fclose(f) == 0 && ok;
}

int main(int argc, char** argv)
{
    std::vector<const char*> pos;
    int                      device = -1, repeat = 1;
    for (int i = 1; i < argc; i++)
    {
        if (!strcmp(argv[i], "--device") && i + 1 < argc)
            device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc)
            repeat = atoi(argv[++i]);
        else
            pos.push_back(argv[i]);
    }
    if (pos.size() != 4 || repeat < 1)
    {
        fprintf(stderr, "usage: %s <circuit.zkey> <witness.wtns> <proof.json> <public.json> [--device N] [--repeat K]\n", argv[0]);
        return 1;
    }
    uint32_t n_vars = 0, n_public = 0, domain = 0;
    uint64_t n_coefs = 0;
    int      state   = 0;
    if (kzp_host_parse_zkey(pos[0], &n_vars, &n_public, &domain, &n_coefs, &state) != KZP_OK || state != KZP_STATE_OK)
    {
        fprintf(stderr, "kzp_prove: zkey not usable (state %d): %s\n", state, kzp_last_error());
        return 2;
    }
    kzp_prover* p = kzp_prover_new(pos[0], device, &state);
    if (!p || state != KZP_STATE_OK)
    {
        fprintf(stderr, "kzp_prove: prover not ready (state %d): %s\n", state, kzp_last_error());
        kzp_prover_free(p);
        return 2;
    }
    char* json = nullptr;
    int   err = 0, ms = 0;
    for (int it = 0; it < repeat; it++)
    {
        if (json)
            kzp_free(json);
        json = nullptr;
        if (kzp_prover_prove(p, pos[1], nullptr, nullptr, &json, &err, &ms) != KZP_RESPONSE_SUCCESS)
        {
            fprintf(stderr, "kzp_prove: prove failed (ProverError %d): %s\n", err, kzp_last_error());
            kzp_prover_free(p);
            return 3;
        }
        fprintf(stderr, "kzp_prove: proof %d in %d ms (nVars %u, nPublic %u, domain %u)\n", it + 1, ms, n_vars, n_public, domain);
    }
    std::vector<std::string> pub;
    std::string              why;
    if (!read_public(pos[1], n_public, pub, why))
    {
        fprintf(stderr, "kzp_prove: %s\n", why.c_str());
        kzp_free(json);
        kzp_prover_free(p);
        return 3;
    }
    std::string pj = "[";
    for (size_t i = 0; i < pub.size(); i++)
        pj += (i ? ",\"" : "\"") + pub[i] + "\"";
    pj += "]";
    bool ok = write_file(pos[2], json) && write_file(pos[3], pj);
    kzp_free(json);
    kzp_prover_free(p);
    if (!ok)
    {
        fprintf(stderr, "kzp_prove: cannot write outputs\n");
        return 4;
    }
    return 0;
}
