// Test harness (tests/test_host_library.py::test_loaders_and_verifier_clean_under_sanitizers): drives
// kzp_host_verify and kzp_host_pairing_check, built from csrc/verify.cpp with -fsanitize=address,undefined, over a
// list of (possibly corrupt) zkey files. Prints one line per file: "<rc> <verdict>".
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kzp_b200.h"

static std::string slurp(const char* path)
{
    std::string s;
    FILE*       f = fopen(path, "rb");
    if (!f)
        return s;
    char   buf[4096];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0)
        s.append(buf, k);
    fclose(f);
    return s;
}

// argv: proof.json public.bin zkey...
int main(int argc, char** argv)
{
    if (argc < 4)
        return 2;
    std::string proof = slurp(argv[1]), pub = slurp(argv[2]);
    for (int i = 3; i < argc; i++)
    {
        int verdict = -1;
        int rc      = kzp_host_verify(argv[i], proof.c_str(), (const uint8_t*)pub.data(), (uint32_t)(pub.size() / 32), &verdict);
        printf("%d %d\n", rc, verdict);
    }
    // off-curve and degenerate inputs of the pairing entry
    std::vector<uint8_t> g1(64 * 2, 0), g2(128 * 2, 0);
    int                  res = -1;
    int                  rc  = kzp_host_pairing_check(g1.data(), g2.data(), 2, &res); // all points at infinity
    printf("pairing %d %d\n", rc, res);
    g1[0] = 1;
    res   = -1;
    rc    = kzp_host_pairing_check(g1.data(), g2.data(), 2, &res); // (1, 0) is not on the curve
    printf("pairing %d %d\n", rc, res);
    return 0;
}
