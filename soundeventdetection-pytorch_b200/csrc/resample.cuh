// Sample-rate conversion in front of the log-mel path: band-limited (Kaiser-windowed sinc) polyphase FIR.
//
// Replaces the resampling branch of read_multichannel_audio (dataset/dataset_utils.py:77-84: librosa.resample, one
// channel at a time) for files that are not at the working rate.  For reduced rates orig / new = Lo / Ln there are Ln
// filters of `taps` = 2 width + Lo coefficients (host_tables.h: make_resample_filters, float64 on the host, resampy's
// kaiser_best design);   y[i Ln + p] = sum_k h[p][k] x[i Lo + k - width],   x = 0 outside the clip.
// CUDA cores: a CTA stages the input span of a tile of output blocks in shared memory; a thread owns one phase p and
// four consecutive blocks, so every coefficient (read coalesced across the warp from the [tap][phase] table, L2
// resident) feeds four FMAs against shared-memory samples.  Only the ~2 width taps of a phase that are not the clamped tail
// of the window are visited (136 of 283 for 44.1 -> 48 kHz).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "umma.cuh"

namespace sedb {

constexpr int kResampleThreads = 256;
constexpr int kResampleBlocksPerThread = 4;

struct ResampleParams {
    const float* x;        // [n_clips, in_stride]
    float* y;              // [n_clips, out_stride]
    const float* h;        // [span][Ln]: the non-negligible window of each phase (host_tables.h: compact_resample_filters)
    const int* first;      // [Ln]: first tap of that window
    long long in_stride, out_stride;
    int n_in, n_out;
    int lo, ln, width, taps, span;
    int nb;                // output blocks per tile (multiple of kResampleBlocksPerThread)
};

__global__ void __launch_bounds__(kResampleThreads) resample_fir_kernel(const ResampleParams p) {
    extern __shared__ float xs[];
    const int clip = blockIdx.y;
    const long long b0 = static_cast<long long>(blockIdx.x) * p.nb;              // first output block of the tile
    const float* __restrict__ x = p.x + static_cast<long long>(clip) * p.in_stride;
    float* __restrict__ y = p.y + static_cast<long long>(clip) * p.out_stride;
    const int span = (p.nb - 1) * p.lo + p.taps;
    const long long j0 = b0 * p.lo - p.width;                                    // clip sample of xs[0]
    for (int i = threadIdx.x; i < span; i += kResampleThreads) {
        const long long j = j0 + i;
        xs[i] = (j >= 0 && j < p.n_in) ? __ldg(x + j) : 0.f;
    }
    __syncthreads();
    const int groups = p.nb / kResampleBlocksPerThread;
    for (int item = threadIdx.x; item < groups * p.ln; item += kResampleThreads) {
        const int g = item / p.ln, ph = item - g * p.ln;
        const float* __restrict__ h = p.h + ph;
        const float* xg = xs + g * kResampleBlocksPerThread * p.lo + __ldg(p.first + ph);
        float acc[kResampleBlocksPerThread];
#pragma unroll
        for (int b = 0; b < kResampleBlocksPerThread; ++b) acc[b] = 0.f;
#pragma unroll 4
        for (int k = 0; k < p.span; ++k) {
            const float hv = __ldg(h + static_cast<long long>(k) * p.ln);
#pragma unroll
            for (int b = 0; b < kResampleBlocksPerThread; ++b) acc[b] = fmaf(hv, xg[b * p.lo + k], acc[b]);
        }
#pragma unroll
        for (int b = 0; b < kResampleBlocksPerThread; ++b) {
            const long long o = (b0 + g * kResampleBlocksPerThread + b) * p.ln + ph;
            if (o < p.n_out) y[o] = acc[b];
        }
    }
}


// -------------------------------------------------------------------------------------------------------
// The same conversion on the tensor cores.  y[i Ln + p] = sum_k x[i Lo + k - width] h[p][k] is a GEMM against a Toeplitz
// view of the input: M = 128 output blocks i per tile, N = Ln phases (padded to 16), K = taps (padded to 16).
//   * the input span of a tile (127 Lo + K samples, zero outside the clip) is staged in shared memory together with its
//     abs-max, which gives a power-of-two block scale (fp16 halves then cover the tile at ~2^-22 of its peak);
//   * builder thread i forms row i of the A operand one 16-tap K-step at a time (hi / lo fp16 halves, canonical K-major
//     tile, two A slots);  the filters come pre-packed per K-step in the same canonical layout (hi | lo, scaled by 2^8)
//     and are streamed by bulk copies through a ring that runs several K-steps ahead;
//   * one elected lane issues hi*hi + lo*hi + hi*lo per K-step into one TMEM accumulator (N columns); the builders drain
//     it (tcgen05.ld) and store the 128 x Ln outputs.
// Warps: 0-7 builders / epilogue (row = tid & 127; warps 4-7 take the second 8-tap K-group of a step and the odd 16-column
// groups of the accumulator), 8 MMA issuer, 9 filter producer.
constexpr int kRsThreads = 320;
constexpr int kRsBuilders = 256;
constexpr int kRsBSlots = 6;
constexpr int kRsASlots = 4;
constexpr int kRsFilterShift = 8;                   // filters are scaled by 2^8 before the fp16 split

struct ResampleUmmaParams {
    const float* x;
    float* y;
    const uint8_t* hpack;   // [K/16][hi | lo][Npad x 16] canonical K-major fp16 (host_tables.h: pack_resample_filters)
    long long in_stride, out_stride;
    int n_in, n_out;
    int lo, ln, width;
    int npad, nks;          // N padded to 16, K-steps
    int tiles_per_clip, n_tiles;
};

__device__ __forceinline__ void rs_builder_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kRsThreads, 1) resample_umma_kernel(const ResampleUmmaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int span = 127 * p.lo + 16 * p.nks;
    const int b_chunk = p.npad * 64;                                  // hi + lo block of one K-step
    float* xs = reinterpret_cast<float*>(smem);
    uint8_t* a_ring = smem + ((span * 4 + 127) / 128) * 128;          // kRsASlots x (hi 4 KB | lo 4 KB)
    uint8_t* b_ring = a_ring + kRsASlots * 8192;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + kRsBSlots * b_chunk);
    uint64_t* a_full = bars + 0;                 // [kRsASlots]  256 builders
    uint64_t* a_free = bars + kRsASlots;         // [kRsASlots]  tcgen05.commit
    uint64_t* b_full = bars + 2 * kRsASlots;     // [kRsBSlots] bulk copy
    uint64_t* b_free = b_full + kRsBSlots;       // [kRsBSlots] tcgen05.commit
    uint64_t* acc_full = b_free + kRsBSlots;
    uint64_t* epi_done = acc_full + 1;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(epi_done + 1);
    uint32_t* max_s = tmem_ptr_s + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kRsASlots; ++s) {
            mbar_init(&a_full[s], kRsBuilders);
            mbar_init(&a_free[s], 1);
        }
        for (int s = 0; s < kRsBSlots; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_free[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(epi_done, kRsBuilders);
        mbar_fence_init();
        *max_s = 0u;
    }
    if (warp == 8) tmem_alloc<256>(tmem_ptr_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;
    const int n_my = (static_cast<int>(blockIdx.x) < p.n_tiles)
                         ? (p.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                         : 0;

    if (warp == 9) {
        // ---------------------------------------------------------------- filter producer
        const int total = n_my * p.nks;
        for (int g = 0; g < total; ++g) {
            const int s = g % kRsBSlots, u = g / kRsBSlots;
            mbar_wait(&b_free[s], (u & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[s], b_chunk);
                bulk_g2s(b_ring + s * b_chunk, p.hpack + static_cast<size_t>(g % p.nks) * b_chunk, b_chunk, &b_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ---------------------------------------------------------------- MMA issuer
        const uint32_t idesc = make_idesc(kFmtF16, kMajorK, kMajorK, 128, p.npad);
        const uint32_t b_lbo = (p.npad / 8) * 128;
        int g = 0;
        for (int ti = 0; ti < n_my; ++ti) {
            if (ti > 0) {
                mbar_wait(epi_done, (ti - 1) & 1);                       // the accumulator has been drained
                tc_fence_after();
            }
            for (int ks = 0; ks < p.nks; ++ks, ++g) {
                const int sa = g % kRsASlots, ua = g / kRsASlots, sb = g % kRsBSlots, ub = g / kRsBSlots;
                mbar_wait(&a_full[sa], ua & 1);
                fence_proxy_async_smem();          // the builders' st.shared (released by their arrive) -> async proxy
                mbar_wait(&b_full[sb], ub & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t aH = make_smem_desc(smem_u32(a_ring + sa * 8192), 2048, 128);
                    const uint64_t aL = make_smem_desc(smem_u32(a_ring + sa * 8192 + 4096), 2048, 128);
                    const uint64_t bH = make_smem_desc(smem_u32(b_ring + sb * b_chunk), b_lbo, 128);
                    const uint64_t bL = make_smem_desc(smem_u32(b_ring + sb * b_chunk + p.npad * 32), b_lbo, 128);
                    umma_f16(tmem, aH, bH, idesc, ks > 0 ? 1u : 0u);
                    umma_f16(tmem, aL, bH, idesc, 1u);
                    umma_f16(tmem, aH, bL, idesc, 1u);
                    umma_commit(&a_free[sa]);
                    umma_commit(&b_free[sb]);
                    if (ks == p.nks - 1) umma_commit(acc_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- builders (row = tid & 127, K-group = tid >> 7) + epilogue
        const int row = tid & 127, kgrp = tid >> 7;
        int g = 0;
        for (int ti = 0; ti < n_my; ++ti) {
            const int tile = blockIdx.x + ti * gridDim.x;
            const int clip = tile / p.tiles_per_clip;
            const long long b0 = static_cast<long long>(tile - clip * p.tiles_per_clip) * 128;   // first output block
            const float* __restrict__ x = p.x + static_cast<long long>(clip) * p.in_stride;
            float* __restrict__ y = p.y + static_cast<long long>(clip) * p.out_stride;
            const long long j0 = b0 * p.lo - p.width;
            rs_builder_sync();                                           // everyone is done with the previous tile's span
            if (tid == 0) *max_s = 0u;
            float mx = 0.f;
            for (int i = tid; i < span; i += kRsBuilders) {
                const long long j = j0 + i;
                const float v = (j >= 0 && j < p.n_in) ? __ldg(x + j) : 0.f;
                xs[i] = v;
                mx = fmaxf(mx, fabsf(v));
            }
            rs_builder_sync();                                           // (max_s reset visible)
            const uint32_t mw = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
            if (lane == 0) atomicMax(max_s, mw);
            rs_builder_sync();
            // block scale 2^e: 2^e max|x| < 2^14; silence / non-finite input: 1
            int e = 0;
            {
                const float m = __uint_as_float(*max_s);
                if (m > 0.f && m < 3.0e38f) e = 13 - (static_cast<int>((__float_as_uint(m) >> 23) & 0xff) - 127);
                e = max(-100, min(100, e));
            }
            const float scale = __uint_as_float(static_cast<uint32_t>(127 + e) << 23);
            const float unscale = __uint_as_float(static_cast<uint32_t>(127 - e - kRsFilterShift) << 23);
            const float* xr = xs + row * p.lo + 8 * kgrp;
            const uint32_t a_off = (row >> 3) * 128 + (row & 7) * 16 + kgrp * 2048;
            for (int ks = 0; ks < p.nks; ++ks, ++g) {
                const int sa = g % kRsASlots, ua = g / kRsASlots;
                uint32_t hi[4], lw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float v0 = xr[16 * ks + 2 * j] * scale, v1 = xr[16 * ks + 2 * j + 1] * scale;
                    const __half2 hh = __floats2half2_rn(v0, v1);
                    const float2 hf = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    lw[j] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                mbar_wait(&a_free[sa], (ua & 1) ^ 1);
                uint8_t* a = a_ring + sa * 8192 + a_off;
                *reinterpret_cast<uint4*>(a) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(a + 4096) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                mbar_arrive(&a_full[sa]);          // (release; the proxy fence sits on the consumer side, see logmel.cuh)
            }
            // ------------------------------------------------------------ epilogue: lane = output block, columns = phases;
            // warps 0-3 take the even 16-column groups, warps 4-7 the odd ones
            mbar_wait(acc_full, ti & 1);
            tc_fence_after();
            const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
            const long long o0 = (b0 + row) * p.ln;
            for (int c0 = 16 * kgrp; c0 < p.npad; c0 += 32) {
                float v[16];
                tmem_ld16(tl + c0, v);
                tmem_ld_wait();
                float* dst = y + o0 + c0;
                if (c0 + 16 <= p.ln && o0 + c0 + 16 <= p.n_out && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(dst + i) =
                            make_float4(v[i] * unscale, v[i + 1] * unscale, v[i + 2] * unscale, v[i + 3] * unscale);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < p.ln && o0 + c0 + i < p.n_out) dst[i] = v[i] * unscale;
                }
            }
            tc_fence_before();
            mbar_arrive(epi_done);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}
}  // namespace sedb
