// Micro-benchmark: does MUFU compute two exponentials per instruction in the packed-half form?
//   ex2.approx.ftz.f32 (one result per thread-instruction) against ex2.approx.ftz.f16x2 (two), 16 warps per SM, one CTA per SM.
// If the f16x2 form issues at the f32 rate, a softmax can halve its MUFU time for the share of exponentials that tolerates an
// fp16 mantissa (the probabilities are rounded to fp16 for the PV tensor-core product anyway).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu_f16x2 tools/ubench/mufu_f16x2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>   // 0: f32, 1: f16x2, 2: f16x2 fed by a cvt.rn.f16x2.f32 pack of two fp32 inputs (what a softmax would do)
__global__ void __launch_bounds__(512) bench_kernel(int iters, long long* cycles, float* sink) {
    float acc = 0.f;
    uint32_t hacc = 0;
    const float x = 0.001f * threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const float f = (x + (float)(it * 64 + i)) * 0.001f;
            if (MODE == 0) {
                float e;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f));
                acc += e;
            } else if (MODE == 1) {
                uint32_t h = __float_as_uint(f) & 0x3bff3bffu, e;       // two small halves
                asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(h));
                hacc ^= e;
            } else {
                uint32_t h, e;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f), "f"(f + 0.25f));
                asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(h));
                hacc ^= e;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(hacc);
}

template <int MODE>
static void run(const char* name, int warps, int iters) {
    long long* cyc;
    float* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 148 * 512 * sizeof(float));
    bench_kernel<MODE><<<148, warps * 32>>>(10, cyc, sink);
    bench_kernel<MODE><<<148, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: %s\n", name, cudaGetErrorString(e));
        return;
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double instr = (double)iters * 64 * 32 * warps;
    printf("%-34s %2d warps/SM: %9.0f cycles, %6.2f MUFU thread-instructions/clk/SM, %6.2f exponentials/clk/SM\n", name, warps, avg, instr / avg,
           instr * (MODE == 0 ? 1 : 2) / avg);
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {8, 16}) run<0>("ex2.approx.ftz.f32", w, 4000);
    for (int w : {8, 16}) run<1>("ex2.approx.ftz.f16x2", w, 4000);
    for (int w : {8, 16}) run<2>("cvt.rn.f16x2.f32 + ex2.f16x2", w, 4000);
    return 0;
}
