// ubench_pipes.cu -- issue-rate microbenchmark of the integer / FP32 instructions the codec kernels lean on
// (IMAD.HI for the encoder's fixed-point mul, IMAD, IMAD.WIDE, FFMA, LDS) on one B200.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) k(int32_t *out, int iters, int32_t a0, int32_t b0)
{
    __shared__ int32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = a0 + i;
    __syncthreads();
    int32_t a[8], b = b0 + threadIdx.x;
    float f[8], g = (float)b0;
    long long w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = a0 + i * 7 + threadIdx.x; f[i] = (float)a[i]; w[i] = a[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = __mulhi(a[i], b) + r;                    // IMAD.HI
                if (OP == 1) a[i] = a[i] * b + r;                            // IMAD
                if (OP == 2) w[i] = (long long)(int32_t)w[i] * b + w[i];     // IMAD.WIDE
                if (OP == 3) f[i] = fmaf(f[i], g, 1.0f);                     // FFMA
                if (OP == 4) a[i] = sm[(a[i] + r) & 1023];                   // LDS (dependent)
                if (OP == 5) a[i] += (uint32_t)__mulhi(a[i] ^ r, b);         // IMAD.HI + IADD (the kernels' pattern)
                if (OP == 6) a[i] = __umulhi((uint32_t)a[i], (uint32_t)b) + r; // IMAD.HI.U32
            }
        }
    }
    int32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (int32_t)f[i] + (int32_t)w[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, int32_t *d)
{
    const int iters = 2000, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(d, 10, 3, 5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d, iters, 3, 5);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * iters * 16 * 8;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.2f Gop/s  %6.2f lanes/clk/SM @ %d MHz nominal\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
}

int main()
{
    int32_t *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("IMAD.HI (mulhi)", d);
    run<6>("IMAD.HI.U32 (umulhi)", d);
    run<5>("IMAD.HI + IADD", d);
    run<1>("IMAD (32-bit)", d);
    run<2>("IMAD.WIDE", d);
    run<3>("FFMA", d);
    run<4>("LDS dependent", d);
    return 0;
}
