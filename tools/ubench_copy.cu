// ubench_copy.cu -- PCIe staging patterns used by the host-buffer pipelines: does a strided (2-D) pinned H2D copy
// return to the host immediately, how fast is it, and does it overlap a running kernel / a D2H copy?
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
using clk = std::chrono::steady_clock;
static double ms_since(clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); }
__global__ void spin(long long cycles, int *out)
{
    long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (out && threadIdx.x == 0 && blockIdx.x == 0) *out = 1;
}
int main()
{
    const size_t rows = 1000, pitch = 31749120, width = (524 * 1152 + 1056) * 4ull;   // bench shapes
    const size_t host_bytes = rows * pitch;
    char *hp = nullptr, *dp = nullptr, *hp2 = nullptr, *dp2 = nullptr;
    auto t0 = clk::now();
    if (cudaHostAlloc(&hp, host_bytes, cudaHostAllocDefault) != cudaSuccess) { printf("hostalloc failed\n"); return 1; }
    printf("cudaHostAlloc %.1f GB: %.0f ms\n", host_bytes / 1e9, ms_since(t0));
    cudaMalloc(&dp, rows * width);
    const size_t big = 4ull << 30;
    cudaHostAlloc(&hp2, big, cudaHostAllocDefault);
    cudaMalloc(&dp2, big);
    cudaStream_t s1, s2, s3;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        // (a) 2-D strided H2D
        t0 = clk::now();
        cudaEventRecord(e0, s1);
        cudaMemcpy2DAsync(dp, width, hp, pitch, width, rows, cudaMemcpyHostToDevice, s1);
        cudaEventRecord(e1, s1);
        double call = ms_since(t0);
        cudaStreamSynchronize(s1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("2D H2D  %.2f GB: call returned after %.2f ms, device %.1f ms = %.1f GB/s\n", rows * width / 1e9, call, ms, rows * width / ms / 1e6);
        // (b) per-row loop
        t0 = clk::now();
        cudaEventRecord(e0, s1);
        for (size_t r = 0; r < rows; r++) cudaMemcpyAsync(dp + r * width, hp + r * pitch, width, cudaMemcpyHostToDevice, s1);
        cudaEventRecord(e1, s1);
        call = ms_since(t0);
        cudaStreamSynchronize(s1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("row-loop H2D      : calls returned after %.2f ms, device %.1f ms = %.1f GB/s\n", call, ms, rows * width / ms / 1e6);
        // (c) contiguous D2H 4 GB, alone and with the 2-D H2D running
        cudaEventRecord(e0, s2);
        cudaMemcpyAsync(hp2, dp2, big, cudaMemcpyDeviceToHost, s2);
        cudaEventRecord(e1, s2);
        cudaStreamSynchronize(s2);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("D2H 4 GiB alone   : %.1f ms = %.1f GB/s\n", ms, big / ms / 1e6);
        t0 = clk::now();
        cudaMemcpy2DAsync(dp, width, hp, pitch, width, rows, cudaMemcpyHostToDevice, s1);
        cudaMemcpyAsync(hp2, dp2, big, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
        double both = ms_since(t0);
        printf("2D H2D + D2H together: %.1f ms wall (%.1f GB/s aggregate)\n", both, (rows * width + big) / both / 1e6);
        // (d) overlap with a kernel
        t0 = clk::now();
        spin<<<148, 128, 0, s3>>>(2000000000LL / 10, nullptr);   // ~100 ms at 1.965 GHz
        cudaMemcpy2DAsync(dp, width, hp, pitch, width, rows, cudaMemcpyHostToDevice, s1);
        cudaStreamSynchronize(s1); cudaStreamSynchronize(s3);
        printf("100 ms kernel + 2D H2D: %.1f ms wall\n", ms_since(t0));
    }
    return 0;
}
