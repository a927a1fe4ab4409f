// Issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), 16 warps per SM, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int PACKED>
__global__ void __launch_bounds__(512, 1) k(float* out, int reps, float a, float b) {
    float2 v[8];
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) v[i] = __ffma2_rn(v[i], a2, b2);
            else { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(t1 - t0);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 512 * 4);
    const int reps = 4096;
    for (int p = 0; p < 2; ++p) {
        for (int w = 0; w < 2; ++w) { if (p) k<1><<<148, 512>>>(d, reps, 1.0001f, 0.5f); else k<0><<<148, 512>>>(d, reps, 1.0001f, 0.5f); }
        cudaDeviceSynchronize();
        float c; cudaMemcpy(&c, d, 4, cudaMemcpyDeviceToHost);
        // 16 warps x 16 fp32 FMAs per rep per thread
        printf("%s: %.0f cycles, %.2f fp32 FMA/clk/SM (%.2f warp-instr/clk)\n", p ? "FFMA2" : "FFMA ", c,
               512.0 * 16 * reps / c, 16.0 * (p ? 8 : 16) * reps / c);
    }
    return 0;
}
