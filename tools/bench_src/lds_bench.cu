// lds_bench.cu -- what a 128-bit shared-memory load costs on sm_100a as a function of the address pattern of the warp
// (cycles per LDS.128 with 16 warps of one CTA loading back to back; 1 CTA per SM, so cycles ~ wavefronts of the LSU data pipe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench lds_bench.cu && ./lds_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) lds_kernel(float* out, long long* cycles) {
    __shared__ __align__(16) float s[8192];
    for (int i = threadIdx.x; i < 8192; i += 512) s[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx;                                                      // float4 index of this lane
    if (MODE == 0) idx = warp * 4;                                // the whole warp reads one address
    else if (MODE == 1) idx = warp * 4 + (lane >> 4) * 36;        // two addresses: one per half warp (36: another bank group)
    else if (MODE == 2) idx = warp * 4 + (lane >> 3) * 37;        // four addresses: one per quarter warp
    else if (MODE == 3) idx = (warp * 32 + lane) % 2048;          // 32 distinct, consecutive: 512 contiguous bytes
    else if (MODE == 4) idx = ((lane & 7) * 97 + (lane >> 3) * 0 + warp) % 2040;   // 8 distinct rows (pitch 16 B mod 128), same in every quarter
    else idx = ((lane & 15) * 97 + warp) % 2040;                  // 16 distinct rows, same in both halves
    const float4* p = reinterpret_cast<const float4*>(s);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const float4 v = p[(idx + u * 8 + it) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 512 + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE> void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 512 * sizeof(float)); cudaMalloc(&cyc, sizeof(long long));
    lds_kernel<MODE><<<1, 512>>>(out, cyc);
    lds_kernel<MODE><<<1, 512>>>(out, cyc);
    long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    // 16 warps x 256 x 16 loads
    printf("%-62s %6.2f cycles per warp-wide LDS.128 (SM-wide)\n", name, (double)c / (16.0 * 256 * 16));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("all 32 lanes one address");
    run<1>("two addresses (one per half warp)");
    run<2>("four addresses (one per quarter warp)");
    run<3>("32 distinct consecutive 16-byte words");
    run<4>("8 distinct rows per quarter warp, the same in all quarters");
    run<5>("16 distinct rows per half warp, the same in both halves");
    return 0;
}
