// One-CTA tcgen05 GEMM probe: pins the shared-memory descriptor conventions (K-major / MN-major canonical
// no-swizzle layouts, LBO/SBO meaning, padded strides, negate bit) that the production kernels rely on.
#pragma once
#include "umma.cuh"

namespace sedb {

// D[128,N] = A[128,K] * B[K,N]; a/b/d float32 row-major in global memory.  K <= 64, N <= 256.
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float* __restrict__ d, int N, int K, int a_major,
                                                            int b_major, int pad, int neg_b, int swap_fields) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // operand A: MN = 128 rows; operand B: MN = N
    const uint32_t sbo = 128 + pad;
    const uint32_t a_lbo = (128 / 8) * sbo;
    const uint32_t b_lbo = (N / 8) * sbo;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + (K / 8) * a_lbo;

    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i % K;
        const __nv_bfloat16 v = __float2bfloat16_rn(a[i]);
        const uint32_t off = (m / 8) * sbo + (k / 8) * a_lbo +
                             (a_major == 0 ? (m % 8) * 16 + (k % 8) * 2 : (k % 8) * 16 + (m % 8) * 2);
        *reinterpret_cast<__nv_bfloat16*>(a_s + off) = v;
    }
    for (int i = tid; i < K * N; i += 128) {
        const int k = i / N, n = i % N;
        const __nv_bfloat16 v = __float2bfloat16_rn(b[i]);
        const uint32_t off = (n / 8) * sbo + (k / 8) * b_lbo +
                             (b_major == 0 ? (n % 8) * 16 + (k % 8) * 2 : (k % 8) * 16 + (n % 8) * 2);
        *reinterpret_cast<__nv_bfloat16*>(b_s + off) = v;
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;

    if (a_major == 2) {
        // A through tensor memory: thread (row m) packs its K values two per column into TMEM columns 256.. and the
        // MMA reads them from there (K-major by construction)
        const int m = warp * 32 + lane;
        const uint32_t ta = tmem + 256 + (static_cast<uint32_t>(warp * 32) << 16);
        for (int k = 0; k < K; k += 4) {
            uint32_t w[2];
            for (int h = 0; h < 2; ++h) {
                const __nv_bfloat162 v = __floats2bfloat162_rn(a[m * K + k + 2 * h], a[m * K + k + 2 * h + 1]);
                w[h] = *reinterpret_cast<const uint32_t*>(&v);
            }
            tmem_st2(ta + k / 2, w[0], w[1]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (tid == 0) {
            const uint32_t idesc = make_idesc(kFmtBF16, kMajorK, b_major, 128, N, 0, neg_b);
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t b_addr = smem_u32(b_s) + ks * 2 * b_lbo;
                umma_f16_ts(tmem, tmem + 256 + ks * 8, make_smem_desc(b_addr, b_lbo, sbo), idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(&bar);
        }
    } else if (tid == 0) {
        const uint32_t idesc = make_idesc(kFmtBF16, a_major, b_major, 128, N, 0, neg_b);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = smem_u32(a_s) + ks * 2 * a_lbo;
            const uint32_t b_addr = smem_u32(b_s) + ks * 2 * b_lbo;
            const uint64_t da = swap_fields ? make_smem_desc(a_addr, sbo, a_lbo) : make_smem_desc(a_addr, a_lbo, sbo);
            const uint64_t db = swap_fields ? make_smem_desc(b_addr, sbo, b_lbo) : make_smem_desc(b_addr, b_lbo, sbo);
            umma_f16(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tl + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) d[(warp * 32 + lane) * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}


// tcgen05.mma throughput probe: every CTA issues `reps` MMAs of 128 x N x 16 (K-major or MN-major B, no swizzle)
// round-robin over `n_acc` accumulator tiles and reports cycles per MMA (block 0) in out[0].
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(unsigned long long* out, int N, int b_major, int n_acc,
                                                           int reps, int lbo_a, int lbo_b) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    const int mode = lbo_b >> 24;          // 0: lane-0 branch, 1: converged warp + elect
    lbo_b &= 0xffffff;
    const uint32_t a_off = static_cast<uint32_t>(lbo_a >> 24) * 16;   // A start offset (bytes, 16-byte units): tap shifts
    const int commit_log2 = (lbo_a >> 20) & 0xf;
    lbo_a &= 0xfffff;
    if (mode == 0) {
        if (tid == 0) {
            const uint32_t idesc = make_idesc(kFmtF16, kMajorK, b_major, 128, N);
            const uint64_t da = make_smem_desc(smem_u32(smem), lbo_a, 128);
            const uint64_t db = make_smem_desc(smem_u32(smem) + 32768, lbo_b, 128);
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) umma_f16(tmem + (r % n_acc) * (512 / 4), da + (r & 7) * 16, db, idesc, 1u);
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
        }
    } else if (mode == 2 && warp == 1) {
        // operand-switch probe: MMA r accumulates into tile r % n_acc, reads B block (r / b_run) % nb (blocks 4 KB apart)
        // and A view r % 8 (shifted by 16 bytes); lbo_b carries nb in bits 12..15 and b_run in bits 16..23
        if (tmem != 0) __trap();
        const int nb = (lbo_b >> 12) & 0xf, b_run = (lbo_b >> 16) & 0xff;
        const uint32_t idesc = make_idesc(kFmtF16, kMajorK, b_major, 128, N);
        const uint64_t da = make_smem_desc(smem_u32(smem) + a_off, lbo_a, 128);
        const uint64_t db = make_smem_desc(smem_u32(smem) + 16384, 2048, 128);
        const long long t0 = clock64();
        // n_acc, nb, b_run are powers of two (masks and shifts only on the issuing thread)
        // bits 20..23 of lbo_a: log2 of the commit interval (0 = only at the end); commits go to a barrier nobody waits on
        const uint32_t am = n_acc - 1, bm = nb - 1;
        const int bs = 31 - __clz(b_run);
        const int cint = commit_log2 ? (1 << commit_log2) : (1 << 30);
        for (int r = 0; r < reps; r += 4) {
            if (elect_one()) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    umma_f16(((r + q) & am) * 128, da + ((r + q) & 7), db + (((r + q) >> bs) & bm) * 256, idesc, 1u);
                if (((r + 4) & (cint - 1)) == 0) umma_commit(&bar2);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (tid & 31) == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
    } else if (mode == 1 && warp == 1) {
        // converged warp, uniform operands (the 512-column allocation always starts at TMEM address 0), one elected
        // lane issues
        if (tmem != 0) __trap();
        const uint32_t idesc = make_idesc(kFmtF16, kMajorK, b_major, 128, N);
        const uint64_t da = make_smem_desc(smem_u32(smem) + a_off, lbo_a, 128);
        const uint64_t db = make_smem_desc(smem_u32(smem) + 32768, lbo_b, 128);
        const uint32_t amask = (n_acc >= 4) ? 3u : (n_acc >= 2 ? 1u : 0u);
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
            if (elect_one()) {
                umma_f16(((r + 0) & amask) * 128, da, db, idesc, 1u);
                umma_f16(((r + 1) & amask) * 128, da + 16, db, idesc, 1u);
                umma_f16(((r + 2) & amask) * 128, da + 32, db, idesc, 1u);
                umma_f16(((r + 3) & amask) * 128, da + 48, db, idesc, 1u);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (tid & 31) == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// cp.async.bulk (1-D, TMA engine) global -> shared throughput probe: one converged warp keeps `depth` copies of
// `bytes` each in flight (ring of depth slots, one mbarrier per slot) for `reps` copies, all CTAs reading the same
// `span` bytes of `src` (span = bytes * nsrc); out[0] = block 0's cycles.  spin != 0 polls the barrier without the
// suspend-time hint.
__global__ void __launch_bounds__(256, 1) bulk_rate_kernel(unsigned long long* out, const uint8_t* __restrict__ src,
                                                           int bytes, int depth, int reps, int nsrc, int spin) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[8 * 16];
    const int tid = threadIdx.x, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) {
        for (int i = 0; i < 8 * 16; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // every warp runs its own ring of `depth` slots; the warps share the `reps` copies
    uint64_t* wb = bars + 16 * warp;
    uint8_t* ws = smem + static_cast<size_t>(warp) * depth * bytes;
    const int my_reps = reps / nwarps;
    const long long t0 = clock64();
    for (int r = 0; r < my_reps + depth; ++r) {
        const int s = r % depth, u = r / depth;
        if (u > 0) {                                        // wait for the previous copy into this slot
            if (spin) { while (!mbar_try_wait(&wb[s], (u - 1) & 1)) {} }
            else mbar_wait(&wb[s], (u - 1) & 1);
        }
        if (r < my_reps) {
            if (elect_one()) {
                mbar_arrive_expect_tx(&wb[s], bytes);
                bulk_g2s(ws + static_cast<size_t>(s) * bytes, src + static_cast<size_t>((r * nwarps + warp) % nsrc) * bytes,
                         bytes, &wb[s]);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
}

}  // namespace sedb
