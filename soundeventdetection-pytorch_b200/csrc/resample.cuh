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

}  // namespace sedb
