// Micro-benchmarks behind the attention kernel's ceiling analysis: per-SM throughput of
//   (1) tcgen05.ld (TMEM -> registers, 32x32b.x32: 4 KB per warp instruction),
//   (2) MUFU ex2.approx,
//   (3) both interleaved (the softmax inner loop does one TMEM column read and one ex2 per score),
// for 4 / 8 / 16 resident softmax warps per SM.  One CTA per SM (148 CTAs), clock64() around the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tmem_mufu tools/ubench/tmem_mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../../layoutllm_t2i_b200/csrc/ltt_ptx.cuh"

using namespace ltt;

template <int MODE>   // 0: tcgen05.ld only, 1: ex2 only, 2: ld + ex2 per element, 3: the same, next load issued before the exps
__global__ void __launch_bounds__(512) bench_kernel(int iters, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    float x = 0.001f * threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 3) {
        uint32_t va[32], vb[32];
        tmem_ld32(base, va);
        tmem_ld_wait();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 4; c += 2) {
                tmem_ld32(base + ((warp >> 2) * 128 + c * 32 + 32) % 512, vb);      // in flight while va is exponentiated
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float e;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(va[i]) * 0.001f));
                    acc += e;
                }
                tmem_ld_wait();
                tmem_ld32(base + ((warp >> 2) * 128 + c * 32 + 64) % 512, va);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float e;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(vb[i]) * 0.001f));
                    acc += e;
                }
                tmem_ld_wait();
            }
        }
    } else
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {             // 4 x 32 columns = one 128-column score tile row block per warp
            uint32_t v[32];
            if (MODE != 1) {
                tmem_ld32(base + ((warp >> 2) * 128 + c * 32) % 512, v);
                tmem_ld_wait();
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float f = MODE != 1 ? __uint_as_float(v[i]) : x + (float)(it * 128 + c * 32 + i);   // varies: nothing to hoist
                if (MODE != 0) {
                    float e;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f * 0.001f));
                    acc += e;
                } else {
                    acc += f;
                }
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(slot);
    }
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
    const double elems = (double)iters * 4 * 32 * 32 * warps;      // scores touched per SM
    printf("%-22s %2d warps/SM: %8.0f cycles, %.2f elements/clk/SM", name, warps, avg, elems / avg);
    if (MODE != 1) printf(", TMEM read %.1f B/clk/SM", elems * 4 / avg);
    if (MODE != 0) printf(", ex2 %.2f /clk/SM", elems / avg);
    printf("  -> a 128x128 score tile costs %.0f cycles\n", 16384.0 / (elems / avg));
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {4, 8, 16}) run<0>("tcgen05.ld only", w, 2000);
    for (int w : {4, 8, 16}) run<1>("ex2.approx only", w, 2000);
    for (int w : {4, 8, 16}) run<2>("tcgen05.ld + ex2", w, 2000);
    for (int w : {4, 8, 16}) run<3>("ld + ex2, pipelined", w, 2000);
    return 0;
}
