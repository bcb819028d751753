// Batched-affine bucket accumulation, measured instead of estimated: how many affine point additions
// per second does B200 sustain when every thread advances B independent buckets by one addition per round and shares
// ONE field inversion among them (Montgomery's trick), with accumulators, incoming points and prefix products streaming
// through global memory — the only formulation in which the inversion amortises (B >= 256 per THREAD: a warp cannot
// share an inversion, its lanes would idle through the ~380 products of the Fermat exponentiation).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I keyless-zk-proofs_b200/csrc tools/ubench_affine.cu -o tools/ubench_affine
// Per addition: 6 products (1 prefix, 2 back-substitution, lambda, lambda^2, y3) + 380 / B for the inversion, 320 bytes
// of global traffic. Compare with the XYZZ mixed addition of the shipped accumulate kernel: 1160 wide multiply-adds
// (9.06 product equivalents), 68 bytes, 6.2-6.4 G additions/s. No exceptional cases are handled here (equal x, infinity):
// this is the upper bound of the approach.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ec.cuh"

using namespace kzp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// acc, pts: [b][thread] affine points; prefix: [b][thread] field elements
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
    k_affine_rounds(G1Affine* __restrict__ acc, const G1Affine* __restrict__ pts, Fq* __restrict__ prefix, uint32_t B, uint32_t rounds)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, T = (size_t)gridDim.x * blockDim.x;
    for (uint32_t r = 0; r < rounds; r++)
    {
        Fq run = Fq::one();
        for (uint32_t b = 0; b < B; b++)
        {
            Fq xa = acc[b * T + t].x, xp = pts[b * T + t].x, d;
            Fq::sub(d, xp, xa);
            prefix[b * T + t] = run;
            Fq::mul(run, run, d);
        }
        Fq inv;
        Fq::inv(inv, run);
        for (uint32_t b = B; b-- > 0;)
        {
            G1Affine a = acc[b * T + t], p = pts[b * T + t];
            Fq       d, dinv, lam, x3, y3, s;
            Fq::sub(d, p.x, a.x);
            Fq::mul(dinv, inv, prefix[b * T + t]);
            Fq::mul(inv, inv, d);
            Fq::sub(s, p.y, a.y);
            Fq::mul(lam, s, dinv);
            Fq::sqr(x3, lam);
            Fq::sub(x3, x3, a.x);
            Fq::sub(x3, x3, p.x);
            Fq::sub(s, a.x, x3);
            Fq::mul(y3, lam, s);
            Fq::sub(y3, y3, a.y);
            a.x = x3;
            a.y = y3;
            acc[b * T + t] = a;
        }
    }
}

__global__ void k_fill(uint32_t* p, size_t words, uint32_t seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words)
    {
        uint32_t x = (uint32_t)i * 2654435761u + seed;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        p[i] = (i % 8 == 7) ? (x & 0x0fffffffu) : x; // below the modulus
    }
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int bps : {2, 4})
        for (uint32_t B : {128u, 256u, 512u})
        {
            const int    grid = sms * bps;
            const size_t T = (size_t)grid * 128;
            G1Affine *   acc, *pts;
            Fq*          prefix;
            CK(cudaMalloc(&acc, T * B * sizeof(G1Affine)));
            CK(cudaMalloc(&pts, T * B * sizeof(G1Affine)));
            CK(cudaMalloc(&prefix, T * B * sizeof(Fq)));
            size_t words = T * B * 16;
            k_fill<<<(unsigned)((words + 255) / 256), 256>>>((uint32_t*)acc, words, 1u);
            k_fill<<<(unsigned)((words + 255) / 256), 256>>>((uint32_t*)pts, words, 77u);
            CK(cudaDeviceSynchronize());
            const uint32_t rounds = 4;
            for (int rep = 0; rep < 2; rep++)
            {
                CK(cudaEventRecord(e0));
                if (bps == 2)
                    k_affine_rounds<2><<<grid, 128>>>(acc, pts, prefix, B, rounds);
                else
                    k_affine_rounds<4><<<grid, 128>>>(acc, pts, prefix, B, rounds);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double adds = (double)T * B * rounds;
            printf("CTAs/SM %d, B %3u buckets per thread: %7.2f G affine additions/s  (%.2f ms for %.1f M additions, %.0f GB/s of global traffic)\n",
                   bps, B, adds / ms / 1e6, ms, adds / 1e6, adds * 320 / ms / 1e6);
            CK(cudaFree(acc));
            CK(cudaFree(pts));
            CK(cudaFree(prefix));
        }
    printf("reference: XYZZ mixed addition in the shipped accumulate kernel 6.2-6.4 G additions/s (DESIGN.md section 4)\n");
    return 0;
}
