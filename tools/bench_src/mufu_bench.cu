// Dev microbenchmark: issue rate of MUFU.TANH / F2FP / FADD2 per SM sub-partition on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + i);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 1) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(x[i]), "f"(x[(i + 1) & 7])); acc ^= y; }
            if (MODE == 2) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i])); if (i & 1) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(x[i]), "f"(x[i - 1])); acc ^= y; } }
            if (MODE == 3) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    long long t1 = clock64();
    __syncthreads();
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int per_iter_xu) {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8 * 1024);
    for (int warps : {4, 8, 16}) {
        int iters = 2000;
        k<MODE><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double instr_per_quadrant = (double)iters * per_iter_xu * warps / 4.0;
        printf("%s warps/SM=%d: %lld cycles, %.2f cycles per warp-instruction per sub-partition\n", name, warps, h, h / instr_per_quadrant);
    }
}
int main() {
    run<0>("MUFU.TANH", 8);
    run<3>("MUFU.EX2", 8);
    run<1>("F2FP.BF16x2", 8);
    run<2>("TANH+F2FP(2:1)", 12);
    return 0;
}
