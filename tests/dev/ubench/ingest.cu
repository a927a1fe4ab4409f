// Per-SM L2 -> SM ingest rate: LDG.128 into registers vs cp.async.bulk into shared memory, 148 CTAs x 512 threads,
// each CTA pulls `bytes` (L2-resident) per repetition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void __launch_bounds__(512, 1) k_ldg(const float4* src, float* out, int reps, int vec_per_thread) {
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const float4* p = src + (static_cast<size_t>(blockIdx.x) * reps + r) * 8192;
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) if (i < vec_per_thread) v[i] = __ldg(p + i * 512 + threadIdx.x);
#pragma unroll
        for (int i = 0; i < 16; ++i) if (i < vec_per_thread) acc += v[i].x + v[i].y + v[i].z + v[i].w;
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * 512 + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(t1 - t0);
}
__global__ void __launch_bounds__(512, 1) k_bulk(const float4* src, float* out, int reps, int bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const float4* p = src + (static_cast<size_t>(blockIdx.x) * reps + r) * 8192;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
            for (int o = 0; o < bytes; o += 32768)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + o)), "l"(reinterpret_cast<const uint8_t*>(p) + o), "r"(32768), "r"(smem_u32(&bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(r & 1) : "memory");
        acc += reinterpret_cast<float*>(smem)[threadIdx.x];
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * 512 + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(t1 - t0);
}
int main() {
    const int reps = 4;                       // 148 x 4 x 128 KB = 75.8 MB: L2-resident after the warm-up pass
    float4* src; float* out;
    cudaMalloc(&src, static_cast<size_t>(148) * reps * 131072);
    cudaMemset(src, 0, static_cast<size_t>(148) * reps * 131072);
    cudaMalloc(&out, 148 * 512 * 4);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    for (int vec = 4; vec <= 16; vec *= 2) {
        float c = 0;
        for (int w = 0; w < 3; ++w) k_ldg<<<148, 512>>>(src, out, reps, vec);
        cudaDeviceSynchronize(); cudaMemcpy(&c, out, 4, cudaMemcpyDeviceToHost);
        printf("LDG  %3d KB/rep: %.0f cycles/rep, %.1f B/clk/SM\n", vec * 8, c / reps, vec * 8192.0 * reps / c);
    }
    for (int kb = 32; kb <= 128; kb *= 2) {
        float c = 0;
        for (int w = 0; w < 3; ++w) k_bulk<<<148, 512, 131072>>>(src, out, reps, kb * 1024);
        cudaDeviceSynchronize(); cudaMemcpy(&c, out, 4, cudaMemcpyDeviceToHost);
        printf("BULK %3d KB/rep: %.0f cycles/rep, %.1f B/clk/SM\n", kb, c / reps, kb * 1024.0 * reps / c);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
