// Integer-pipe micro-benchmarks for sm_100a: what a 254-bit Montgomery product can cost at best.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I keyless-zk-proofs_b200/csrc tools/ubench.cu -o tools/ubench
// Prints one line per probe: instruction (or field-op) throughput over all SMs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ec.cuh"

using namespace kzp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// ---- instruction probes: 8 independent chains per thread
__global__ void __launch_bounds__(256) p_wide(uint32_t* sink, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 3u;
    unsigned long long acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a + k;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)acc[(k + 1) & 7]), "r"(b));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567ull) sink[0] = (uint32_t)s;
}

// carry chains exactly as the Montgomery product issues them: 4 wide multiply-adds per chain (1 head + 3 .X)
__global__ void __launch_bounds__(256) p_wide_chain(uint32_t* sink, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 3u;
    uint32_t x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = a + k; y[k] = b + k; }
    for (int it = 0; it < iters; it++)
    {
        asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
            "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
            "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
            "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
            "madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
            : "r"(y[0]), "r"(y[2]), "r"(y[4]), "r"(y[6]), "r"(b));
        asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
            "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
            "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
            "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
            "madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7])
            : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k] ^ y[k];
    if (s == 0x1234567u) sink[0] = s;
}

// 32-bit IMAD (lo) and IMAD.HI, independent
__global__ void __launch_bounds__(256) p_lo(uint32_t* sink, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 3u;
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a + k;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 1) & 7]), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567u) sink[0] = s;
}
__global__ void __launch_bounds__(256) p_hi(uint32_t* sink, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 3u;
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a + k;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 1) & 7]), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567u) sink[0] = s;
}
// IADD3 with carry chains (alu pipe)
__global__ void __launch_bounds__(256) p_addc(uint32_t* sink, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + 1u, b = blockIdx.x * 40503u + 3u;
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a + k;
    for (int it = 0; it < iters; it++)
    {
        asm volatile("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %8;\n\t"
            "addc.cc.u32 %3, %3, %9;\n\t"
            "addc.cc.u32 %4, %4, %8;\n\t"
            "addc.cc.u32 %5, %5, %9;\n\t"
            "addc.cc.u32 %6, %6, %8;\n\t"
            "addc.u32 %7, %7, %9;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
            : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    if (s == 0x1234567u) sink[0] = s;
}
__global__ void __launch_bounds__(256) p_dfma(uint32_t* sink, int iters)
{
    double a = threadIdx.x * 1.0000001 + 1.0, b = 1.0 + 1e-9 * blockIdx.x;
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a + k;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k]) : "d"(acc[(k + 1) & 7]), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0.12345) sink[0] = 1;
}

// ---- field-op probes: dependent chain per thread, ILP from `W` independent elements per thread
template <class F, int W, int OP>
__global__ void __launch_bounds__(256) p_field(F* out, int iters)
{
    F x[W], y[W];
#pragma unroll
    for (int k = 0; k < W; k++)
    {
        x[k] = F::one();
        y[k] = F::r2();
        x[k].v[0] += threadIdx.x + k;
        y[k].v[1] += blockIdx.x;
    }
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int k = 0; k < W; k++)
        {
            if (OP == 0) F::mul(x[k], x[k], y[k]);
            else if (OP == 1) F::sqr(x[k], x[k]);
            else if (OP == 2) F::add(x[k], x[k], y[k]);
            else F::sub(x[k], x[k], y[k]);
        }
    }
    F s = x[0];
#pragma unroll
    for (int k = 1; k < W; k++) F::add(s, s, x[k]);
    if (s.v[0] == 0x12345u && s.v[7] == 77u) out[0] = s;
}

template <class XY, int OP>
__global__ void __launch_bounds__(128) p_point(XY* out, const typename XY::Affine* in, int iters)
{
    XY acc;
    XY::set_inf(acc);
    typename XY::Affine q = in[threadIdx.x & 7];
    XY::madd(acc, q);
    q = in[8 + (threadIdx.x & 7)];
    for (int it = 0; it < iters; it++)
    {
        if (OP == 0) XY::madd(acc, q);
        else { XY t = acc; XY::dbl(acc, t); }
    }
    if (XY::Field::is_zero(acc.x)) out[0] = acc;
}

template <class XY, int OP>
__global__ void __launch_bounds__(32) p_point_lat(XY* out, const typename XY::Affine* in, int iters)
{
    XY acc, q;
    XY::set_inf(acc);
    XY::set_inf(q);
    XY::madd(acc, in[threadIdx.x & 7]);
    XY::madd(q, in[8 + (threadIdx.x & 7)]);
    XY t = q;
    XY::dbl(q, t);
    for (int it = 0; it < iters; it++)
    {
        if (OP == 0) XY::add(acc, q);
        else { XY t2 = acc; XY::dbl(acc, t2); }
    }
    if (XY::Field::is_zero(acc.x)) out[0] = acc;
}

template <class K>
static float time_it(K&& launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, %d kHz\n", prop.name, sms, prop.clockRate);
    uint32_t* sink;
    CK(cudaMalloc(&sink, 1 << 20));
    const int iters = 4096;
    for (int bps = 1; bps <= 8; bps *= 2)
    {
        int    grid = sms * bps;
        double n    = (double)grid * 256 * 8 * iters;
        float  t;
        t = time_it([&] { p_wide<<<grid, 256>>>(sink, iters); });
        printf("blocks/SM %d  IMAD.WIDE plain      %7.2f T/s\n", bps, n / t / 1e9);
        t = time_it([&] { p_wide_chain<<<grid, 256>>>(sink, iters); });
        printf("blocks/SM %d  IMAD.WIDE.X chains   %7.2f T/s\n", bps, n / t / 1e9);
        t = time_it([&] { p_lo<<<grid, 256>>>(sink, iters); });
        printf("blocks/SM %d  IMAD lo              %7.2f T/s\n", bps, n / t / 1e9);
        t = time_it([&] { p_hi<<<grid, 256>>>(sink, iters); });
        printf("blocks/SM %d  IMAD.HI              %7.2f T/s\n", bps, n / t / 1e9);
        t = time_it([&] { p_addc<<<grid, 256>>>(sink, iters); });
        printf("blocks/SM %d  IADD3.X chains       %7.2f T/s\n", bps, n / t / 1e9);
        t = time_it([&] { p_dfma<<<grid, 256>>>(sink, iters); });
        printf("blocks/SM %d  DFMA                 %7.2f T/s\n", bps, n / t / 1e9);
    }
    const int fit = 2048;
    for (int bps = 1; bps <= 4; bps *= 2)
    {
        int   grid = sms * bps;
        float t;
        t = time_it([&] { p_field<Fq, 1, 0><<<grid, 256>>>((Fq*)sink, fit); });
        printf("blocks/SM %d  Fq mul W=1  %7.2f G mul/s\n", bps, (double)grid * 256 * fit / t / 1e6);
        t = time_it([&] { p_field<Fq, 2, 0><<<grid, 256>>>((Fq*)sink, fit); });
        printf("blocks/SM %d  Fq mul W=2  %7.2f G mul/s\n", bps, (double)grid * 256 * 2 * fit / t / 1e6);
        t = time_it([&] { p_field<Fq, 4, 0><<<grid, 256>>>((Fq*)sink, fit); });
        printf("blocks/SM %d  Fq mul W=4  %7.2f G mul/s\n", bps, (double)grid * 256 * 4 * fit / t / 1e6);
        t = time_it([&] { p_field<Fq, 2, 1><<<grid, 256>>>((Fq*)sink, fit); });
        printf("blocks/SM %d  Fq sqr W=2  %7.2f G sqr/s\n", bps, (double)grid * 256 * 2 * fit / t / 1e6);
        t = time_it([&] { p_field<Fq, 4, 2><<<grid, 256>>>((Fq*)sink, fit); });
        printf("blocks/SM %d  Fq add W=4  %7.2f G add/s\n", bps, (double)grid * 256 * 4 * fit / t / 1e6);
        t = time_it([&] { p_field<Fq, 4, 3><<<grid, 256>>>((Fq*)sink, fit); });
        printf("blocks/SM %d  Fq sub W=4  %7.2f G sub/s\n", bps, (double)grid * 256 * 4 * fit / t / 1e6);
    }
    // point ops: G1 generator multiples as inputs
    {
        G1Affine h[16];
        G1Xyzz   g, acc;
        G1Affine ga;
        ga.x = Fq::one();
        Fq two = Fq::one();
        Fq::add(two, two, two);
        ga.y = two;
        G1Xyzz::from_affine(g, ga);
        acc = g;
        for (int i = 0; i < 16; i++)
        {
            G1Xyzz t = acc;
            G1Xyzz::dbl(acc, t);
            G1Xyzz::add(acc, g);
            G1Xyzz::to_affine(h[i], acc);
        }
        G1Affine* d;
        CK(cudaMalloc(&d, sizeof(h)));
        CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
        for (int bps = 1; bps <= 8; bps *= 2)
        {
            int   grid = sms * bps;
            float t    = time_it([&] { p_point<G1Xyzz, 0><<<grid, 128>>>((G1Xyzz*)sink, d, 512); });
            printf("blocks/SM %d (128 thr)  G1 madd  %7.2f G madd/s  (%.2f G fq-mul/s)\n", bps, (double)grid * 128 * 512 / t / 1e6,
                   (double)grid * 128 * 512 * 10 / t / 1e6);
            t = time_it([&] { p_point<G1Xyzz, 1><<<grid, 128>>>((G1Xyzz*)sink, d, 512); });
            printf("blocks/SM %d (128 thr)  G1 dbl   %7.2f G dbl/s\n", bps, (double)grid * 128 * 512 / t / 1e6);
        }
    }
    // lone-warp latency (what bounds the MSM tail kernels): one warp per SM
    {
        G1Affine* d;
        CK(cudaMalloc(&d, 16 * sizeof(G1Affine)));
        G1Affine h[16];
        G1Xyzz   g, acc;
        G1Affine ga;
        ga.x = Fq::one();
        Fq two = Fq::one();
        Fq::add(two, two, two);
        ga.y = two;
        G1Xyzz::from_affine(g, ga);
        acc = g;
        for (int i = 0; i < 16; i++)
        {
            G1Xyzz t = acc;
            G1Xyzz::dbl(acc, t);
            G1Xyzz::add(acc, g);
            G1Xyzz::to_affine(h[i], acc);
        }
        CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
        for (int w = 1; w <= 4; w *= 2)
        {
            float t = time_it([&] { p_field<Fq, 1, 0><<<sms, 32 * w>>>((Fq*)sink, 2048); });
            printf("lone %d warp(s)/SM: Fq mul chain        %7.1f ns per mul\n", w, t * 1e6 / 2048);
            t = time_it([&] { p_field<Fq, 2, 0><<<sms, 32 * w>>>((Fq*)sink, 2048); });
            printf("lone %d warp(s)/SM: Fq mul 2 chains     %7.1f ns per mul\n", w, t * 1e6 / 4096);
            t = time_it([&] { p_field<Fq, 4, 0><<<sms, 32 * w>>>((Fq*)sink, 2048); });
            printf("lone %d warp(s)/SM: Fq mul 4 chains     %7.1f ns per mul\n", w, t * 1e6 / 8192);
        }
        float t = time_it([&] { p_point_lat<G1Xyzz, 0><<<sms, 32>>>((G1Xyzz*)sink, d, 256); });
        printf("lone warp: G1 add %7.2f us, ", t * 1e3 / 256);
        t = time_it([&] { p_point_lat<G1Xyzz, 1><<<sms, 32>>>((G1Xyzz*)sink, d, 256); });
        printf("G1 dbl %7.2f us\n", t * 1e3 / 256);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
