// Fused framing + Hann window + real DFT (tcgen05 tensor cores) + |X|^2 + mel filterbank + log-dB.
//
// Replaces, for the reference's fixed configuration (dataset/common_config.py:2-8,
// dataset/spectogram/spectogram_configs.py:5-8), the composition
//     multichannel_complex_to_log_mel(multichannel_stft(x))      (dataset/spectogram/preprocess.py:21-45)
// without ever writing the (T x 16385) complex STFT to HBM.
//
// Algorithm (see DESIGN.md section "K1/K2"): N = 32768 = 256 (n1) x 128 (n2), n = 128 n1 + n2, k = k1 + 256 k2.
//   stage 1  GEMM  Dc/Ds[k1,n2] = sum_n1 {cos,-sin}(2 pi k1 n1/256) g[128 n1 + n2]        (constants = A, frame = B)
//   twiddle  CUDA  Z[k1,n2] = (Dc + i Ds) exp(-2 pi i k1 n2/32768)                         (TMEM -> regs -> smem)
//   stage 2  GEMM  X[k1 + 256 k2] = sum_n2 Z[k1,n2] exp(-2 pi i n2 k2/128)                 (Z = A, constants = B)
//   row 128  CUDA  X[128 + 256 k2] from Y[n2,128] (kept in row 0 of the sine block)
//   power, Hermitian bin map, banded mel dot products, 10 log10(max(1e-10, .)).
// All GEMM operands are split x = hi + lo (two 16-bit floats) and multiplied as hi*hi + lo*hi + hi*lo with
// fp32 accumulation in TMEM, which keeps the DFT at ~2^-17 (bf16) / 2^-22 (fp16) relative accuracy.
#pragma once
#include "umma.cuh"

namespace sedb {

constexpr int kSampleRate = 48000;
constexpr int kWin = 31680;          // frame_size  (dataset/common_config.py:4)
constexpr int kHop = 15840;          // hop_size    (dataset/common_config.py:5)
constexpr int kNfft = 32768;         // NFFT        (dataset/spectogram/spectogram_configs.py:5)
constexpr int kBins = kNfft / 2 + 1; // 16385
constexpr int kMel = 64;             // mel_bins
constexpr int kLpad = (kNfft - kWin) / 2;   // 544 zeros each side of the window (librosa pad_center)
constexpr int kPadRefl = kNfft / 2;         // 16384 reflect-padded samples each side (center=True)

// ---- shared memory map of the fused kernel -------------------------------------------------------------
constexpr int kB2ArrBytes = 128 * 128 * 2;                 // one resident stage-2 constant array (32 KB)
constexpr int kB2Bytes = 4 * kB2ArrBytes;                  // cH | cL | sH | sL
constexpr int kA1ArrBytes = 128 * 16 * 2;                  // one stage-1 constant array per K-chunk (4 KB)
constexpr int kA1ChunkBytes = 4 * kA1ArrBytes;             // cH | cL | sH | sL  (16 KB)
constexpr int kB1Sbo = 144;                                // padded stride between 8-sample groups (bank spread)
constexpr int kB1Lbo = 16 * kB1Sbo;                        // 2304: stride between groups of 8 rows (K)
constexpr int kB1ArrBytes = 2 * kB1Lbo;                    // 4608
constexpr int kSlotBytes = 26624;                          // >= 16384 + 2*4608 (stage 1) and >= 6*4096 (stage 2)
constexpr int kNumSlots = 3;
constexpr int kRingBytes = kSlotBytes * kNumSlots;         // 79872 >= 16385*4 (power spectrum aliases the ring)
constexpr int kA2ArrBytes = 128 * 16 * 2;                  // 4 KB

constexpr int kOffB2 = 0;
constexpr int kOffRing = kOffB2 + kB2Bytes;                // 131072
constexpr int kOffV = kOffRing + kRingBytes;               // Y[n2,128]   128 floats
constexpr int kOffCs = kOffV + 512;                        // exp(-2 pi i j/256) 256 float2
constexpr int kOffMelTab = kOffCs + 2048;                  // 64 x int4 {lo, cnt, woff, 0}
constexpr int kOffBars = kOffMelTab + 1024;                // mbarriers
constexpr int kOffTmem = kOffBars + 256;                   // tmem base address
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kRingBytes >= kBins * 4, "power spectrum must fit in the ring");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr int kWorkerWarps = 8;
constexpr int kWorkerThreads = kWorkerWarps * 32;
constexpr int kMmaWarp = 8;
constexpr int kCopyWarp = 9;
constexpr int kThreads = 320;

struct LogmelParams {
    const float* wave;        // [B, wave_stride] fp32
    long long wave_stride;    // elements between clips
    int n_samples;            // valid samples per clip
    int n_clips;
    int n_frames;             // T = 1 + n_samples / hop
    const uint8_t* a1;        // stage-1 constants, 16 chunks x 16 KB, canonical K-major, split hi/lo
    const uint8_t* b2;        // stage-2 constants, 4 x 32 KB
    const float* mel_w;       // compact mel weights (band of filter m starts at mel_tab[m].z)
    const int4* mel_tab;      // 64 x {first bin, count, weight offset, 0}
    const float* norm;        // nullable: mean[64] then std[64]  (SpectogramDataset.transform, logMel mode)
    float* out;               // MODE 0: [B, T, 64] fp32 log-mel
    float2* spec;             // MODE 1: [B, T, 16385] complex64 STFT
};

// named barrier among the 256 worker threads only
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float hann_padded(int n) {
    // np.hanning(31680) centre-padded to 32768: 0.5 - 0.5 cos(2 pi (n-544)/31679) on [544, 32224), else 0.
    // cos(2 pi f) = -cos(2 pi (f - 0.5)), argument kept in [-pi, pi] for the fast intrinsic.
    float f = static_cast<float>(n - kLpad) * (1.0f / static_cast<float>(kWin - 1));
    float c = __cosf((f - 0.5f) * 6.283185307179586f);
    return 0.5f + 0.5f * c;
}

__device__ __forceinline__ int reflect_index(int j, int n) {
    // np.pad(mode='reflect') for a pad shorter than the signal: -1 -> 1, n -> n-2
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    return j;
}

// Banded mel dot products over a power spectrum in shared memory, then dB (+ optional normalisation).
// Called by the 8 worker warps; warp w owns filters w, w+8, ...
__device__ __forceinline__ void mel_db_from_smem(const float* __restrict__ p_s, const int4* __restrict__ mel_tab_s,
                                                 const float* __restrict__ mel_w, const float* __restrict__ norm,
                                                 float inv_scale2, float* __restrict__ out_row, int warp, int lane) {
    float mine = 0.f;
#pragma unroll 1
    for (int i = 0; i < kMel / kWorkerWarps; ++i) {
        const int m = warp + i * kWorkerWarps;
        const int4 tab = mel_tab_s[m];
        const float* w = mel_w + tab.z;
        const float* p = p_s + tab.x;
        float acc = 0.f;
        for (int j = lane; j < tab.y; j += 32) acc = fmaf(p[j], __ldg(w + j), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == i) mine = acc;
    }
    if (lane < kMel / kWorkerWarps) {
        const int m = warp + lane * kWorkerWarps;
        float db = 10.0f * log10f(fmaxf(1e-10f, mine * inv_scale2));     // librosa.power_to_db(ref=1, amin=1e-10)
        if (norm != nullptr) db = (db - norm[m]) / norm[kMel + m];       // spectograms_dataset.py:105
        out_row[m] = db;
    }
}

template <int MODE>   // 0: log-mel output; 1: complex STFT output
__global__ void __launch_bounds__(kThreads, 1) logmel_fused_kernel(const LogmelParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* b2_s = smem + kOffB2;
    uint8_t* ring = smem + kOffRing;
    float* p_s = reinterpret_cast<float*>(ring);
    float* v_s = reinterpret_cast<float*>(smem + kOffV);
    float2* cs_s = reinterpret_cast<float2*>(smem + kOffCs);
    int4* mel_tab_s = reinterpret_cast<int4*>(smem + kOffMelTab);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    uint64_t* full1 = bars + 0;      // [3] stage-1 slot filled (8 worker warps + 1 bulk copy)
    uint64_t* empty1 = bars + 3;     // [3] stage-1 slot consumed (tcgen05.commit)
    uint64_t* full2 = bars + 6;      // [3] stage-2 slot filled (4 worker warps)
    uint64_t* empty2 = bars + 9;     // [3] stage-2 slot consumed
    uint64_t* d1_full = bars + 12;   // stage-1 accumulators complete
    uint64_t* d2_full = bars + 13;   // stage-2 accumulators complete
    uint64_t* ring_free = bars + 14; // workers finished the frame (power spectrum no longer aliases the ring)
    uint64_t* b2_full = bars + 15;   // resident stage-2 constants landed

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kNumSlots; ++s) {
            mbar_init(&full1[s], kWorkerWarps + 1);
            mbar_init(&empty1[s], 1);
            mbar_init(&full2[s], 4);
            mbar_init(&empty2[s], 1);
        }
        mbar_init(d1_full, 1);
        mbar_init(d2_full, 1);
        mbar_init(ring_free, kWorkerWarps);
        mbar_init(b2_full, 1);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc<512>(tmem_ptr_s);
    for (int i = tid; i < 256; i += kThreads) {
        float s, c;
        sincospif(static_cast<float>(i) * (1.0f / 128.0f), &s, &c);      // exp(-2 pi i j/256) = (c, -s)
        cs_s[i] = make_float2(c, -s);
    }
    for (int i = tid; i < kMel; i += kThreads) mel_tab_s[i] = prm.mel_tab[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    const long long total_frames = static_cast<long long>(prm.n_clips) * prm.n_frames;
    const int n_iter = (blockIdx.x < total_frames)
                           ? static_cast<int>((total_frames - blockIdx.x + gridDim.x - 1) / gridDim.x)
                           : 0;

    // ======================================================================== bulk-copy producer warp
    if (warp == kCopyWarp) {
        if (lane == 0) {
            mbar_arrive_expect_tx(b2_full, kB2Bytes);
            for (int a = 0; a < 4; ++a)
                bulk_g2s(b2_s + a * kB2ArrBytes, prm.b2 + a * kB2ArrBytes, kB2ArrBytes, b2_full);
            for (int it = 0; it < n_iter; ++it) {
                if (it > 0) mbar_wait(ring_free, (it - 1) & 1);
                for (int c = 0; c < 16; ++c) {
                    const int g = it * 16 + c;
                    const int s = g % kNumSlots;
                    const int u = g / kNumSlots;
                    mbar_wait(&empty1[s], (u & 1) ^ 1);
                    mbar_arrive_expect_tx(&full1[s], kA1ChunkBytes);
                    bulk_g2s(ring + s * kSlotBytes, prm.a1 + c * kA1ChunkBytes, kA1ChunkBytes, &full1[s]);
                }
            }
        }
    }
    // ======================================================================== MMA issuer warp
    else if (warp == kMmaWarp) {
        if (lane == 0) {
            constexpr uint32_t idesc1 = make_idesc(kSplitFmt, kMajorK, kMajorMN, 128, 128);
            constexpr uint32_t idesc2 = make_idesc(kSplitFmt, kMajorK, kMajorK, 128, 128);
            const uint32_t ring_a = smem_u32(ring);
            const uint32_t b2_a = smem_u32(b2_s);
            mbar_wait(b2_full, 0);
            for (int it = 0; it < n_iter; ++it) {
                // ---------------- stage 1: 16 K-chunks of 16 rows (n1)
                for (int c = 0; c < 16; ++c) {
                    const int g = it * 16 + c;
                    const int s = g % kNumSlots;
                    const int u = g / kNumSlots;
                    mbar_wait(&full1[s], u & 1);
                    tc_fence_after();
                    const uint32_t slot = ring_a + s * kSlotBytes;
                    const uint64_t cH = make_smem_desc(slot + 0 * kA1ArrBytes, 2048, 128);
                    const uint64_t cL = make_smem_desc(slot + 1 * kA1ArrBytes, 2048, 128);
                    const uint64_t sH = make_smem_desc(slot + 2 * kA1ArrBytes, 2048, 128);
                    const uint64_t sL = make_smem_desc(slot + 3 * kA1ArrBytes, 2048, 128);
                    const uint64_t xH = make_smem_desc(slot + kA1ChunkBytes, kB1Lbo, kB1Sbo);
                    const uint64_t xL = make_smem_desc(slot + kA1ChunkBytes + kB1ArrBytes, kB1Lbo, kB1Sbo);
                    const uint32_t acc = (c > 0) ? 1u : 0u;
                    umma_f16(tmem + 0, cH, xH, idesc1, acc);
                    umma_f16(tmem + 0, cL, xH, idesc1, 1u);
                    umma_f16(tmem + 0, cH, xL, idesc1, 1u);
                    umma_f16(tmem + 128, sH, xH, idesc1, acc);
                    umma_f16(tmem + 128, sL, xH, idesc1, 1u);
                    umma_f16(tmem + 128, sH, xL, idesc1, 1u);
                    umma_commit(&empty1[s]);
                }
                umma_commit(d1_full);
                // ---------------- stage 2: 8 K-chunks of 16 columns (n2), consumption order 0,4,1,5,...
                for (int j = 0; j < 8; ++j) {
                    const int g = it * 8 + j;
                    const int s = g % kNumSlots;
                    const int u = g / kNumSlots;
                    const int chunk = (j & 1) * 4 + (j >> 1);
                    mbar_wait(&full2[s], u & 1);
                    tc_fence_after();
                    const uint32_t slot = ring_a + s * kSlotBytes;
                    const uint64_t zrH = make_smem_desc(slot + 0 * kA2ArrBytes, 2048, 128);
                    const uint64_t zrL = make_smem_desc(slot + 1 * kA2ArrBytes, 2048, 128);
                    const uint64_t ziH = make_smem_desc(slot + 2 * kA2ArrBytes, 2048, 128);
                    const uint64_t ziL = make_smem_desc(slot + 3 * kA2ArrBytes, 2048, 128);
                    const uint64_t nrH = make_smem_desc(slot + 4 * kA2ArrBytes, 2048, 128);
                    const uint64_t nrL = make_smem_desc(slot + 5 * kA2ArrBytes, 2048, 128);
                    const uint32_t koff = static_cast<uint32_t>(chunk) * 2 * 2048;   // 16 n2 = 2 K-groups
                    const uint64_t cH = make_smem_desc(b2_a + 0 * kB2ArrBytes + koff, 2048, 128);
                    const uint64_t cL = make_smem_desc(b2_a + 1 * kB2ArrBytes + koff, 2048, 128);
                    const uint64_t sH = make_smem_desc(b2_a + 2 * kB2ArrBytes + koff, 2048, 128);
                    const uint64_t sL = make_smem_desc(b2_a + 3 * kB2ArrBytes + koff, 2048, 128);
                    const uint32_t acc = (j > 0) ? 1u : 0u;
                    // re = Zr c + Zi s
                    umma_f16(tmem + 256, zrH, cH, idesc2, acc);
                    umma_f16(tmem + 256, zrL, cH, idesc2, 1u);
                    umma_f16(tmem + 256, zrH, cL, idesc2, 1u);
                    umma_f16(tmem + 256, ziH, sH, idesc2, 1u);
                    umma_f16(tmem + 256, ziL, sH, idesc2, 1u);
                    umma_f16(tmem + 256, ziH, sL, idesc2, 1u);
                    // im = Zi c - Zr s
                    umma_f16(tmem + 384, ziH, cH, idesc2, acc);
                    umma_f16(tmem + 384, ziL, cH, idesc2, 1u);
                    umma_f16(tmem + 384, ziH, cL, idesc2, 1u);
                    umma_f16(tmem + 384, nrH, sH, idesc2, 1u);
                    umma_f16(tmem + 384, nrL, sH, idesc2, 1u);
                    umma_f16(tmem + 384, nrH, sL, idesc2, 1u);
                    umma_commit(&empty2[s]);
                }
                umma_commit(d2_full);
            }
        }
    }
    // ======================================================================== 8 worker warps
    else {
        const int q = warp & 3;                 // TMEM lane quarter
        const int h = warp >> 2;                // column half
        const int k1 = q * 32 + lane;           // this thread's stage-1 output row / stage-2 A row
        const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);

        // per-thread twiddle bases: w1^j (j=0..3), w1^(4i) (i=0..3), anchors w1^(16 cc + 64 h) (cc=0..3)
        float2 wj[4], w4[4], anc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s, c;
            sincospif(static_cast<float>(k1 * j) * (1.0f / 16384.0f), &s, &c);
            wj[j] = make_float2(c, -s);
            sincospif(static_cast<float>(k1 * 4 * j) * (1.0f / 16384.0f), &s, &c);
            w4[j] = make_float2(c, -s);
            sincospif(static_cast<float>(k1 * (16 * j + 64 * h)) * (1.0f / 16384.0f), &s, &c);
            anc[j] = make_float2(c, -s);
        }

        // stage-1 producer mapping: thread -> (row r of the 16-row chunk, group g of 8 samples)
        const int r = tid >> 4;                 // 0..15
        const int grp = tid & 15;               // 0..15
        const uint32_t b1_off = kA1ChunkBytes + grp * kB1Sbo + (r >> 3) * kB1Lbo + (r & 7) * 16;

        for (int it = 0; it < n_iter; ++it) {
            const long long f = blockIdx.x + static_cast<long long>(it) * gridDim.x;
            const int clip = static_cast<int>(f / prm.n_frames);
            const int t = static_cast<int>(f - static_cast<long long>(clip) * prm.n_frames);
            const float* __restrict__ y = prm.wave + static_cast<long long>(clip) * prm.wave_stride;
            const int L = prm.n_samples;
            const bool vec_ok = ((reinterpret_cast<uintptr_t>(y) & 15) == 0);

            // ---------------------------------------------------------------- stage 1: window + split
            float x[8];
            auto load_chunk = [&](int c) {
                const int n0 = 128 * (16 * c + r) + 8 * grp;          // position inside the padded frame
                const int j0 = t * kHop + n0 - kPadRefl;              // position inside the clip
                if (n0 < kLpad || n0 >= kLpad + kWin) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = 0.f;
                } else if (vec_ok && j0 >= 0 && j0 + 8 <= L) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(y + j0));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(y + j0 + 4));
                    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
                    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = __ldg(y + reflect_index(j0 + e, L));
                }
            };
            load_chunk(0);
#pragma unroll 1
            for (int c = 0; c < 16; ++c) {
                const int g = it * 16 + c;
                const int s = g % kNumSlots;
                const int u = g / kNumSlots;
                const int n0 = 128 * (16 * c + r) + 8 * grp;
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float a = x[2 * e], b = x[2 * e + 1];
                    if (n0 >= kLpad && n0 < kLpad + kWin) {
                        a *= hann_padded(n0 + 2 * e);
                        b *= hann_padded(n0 + 2 * e + 1);
                    }
                    split_pack2(a, b, hi[e], lo[e]);
                }
                if (c + 1 < 16) load_chunk(c + 1);                    // prefetch next chunk's samples
                mbar_wait(&empty1[s], (u & 1) ^ 1);
                uint8_t* dst = ring + s * kSlotBytes + b1_off;
                *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(dst + kB1ArrBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full1[s]);
            }

            // ---------------------------------------------------------------- twiddle + stage-2 A operand
            mbar_wait(d1_full, it & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int j = cc * 2 + h;                             // consumption order index
                const int g = it * 8 + j;
                const int s = g % kNumSlots;
                const int u = g / kNumSlots;
                const int n2_0 = 64 * h + 16 * cc;
                float yr[16], yi[16];
                tmem_ld16(tlane + n2_0, yr);
                tmem_ld16(tlane + 128 + n2_0, yi);
                tmem_ld_wait();
                if (k1 == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { v_s[n2_0 + i] = yi[i]; yi[i] = 0.f; }
                }
                uint32_t zrh[8], zrl[8], zih[8], zil[8];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    // base twiddle for columns n2_0 + 4*i4 .. +3
                    const float2 tb = make_float2(anc[cc].x * w4[i4].x - anc[cc].y * w4[i4].y,
                                                  anc[cc].x * w4[i4].y + anc[cc].y * w4[i4].x);
                    float zr[4], zi[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float twr = tb.x * wj[jj].x - tb.y * wj[jj].y;
                        const float twi = tb.x * wj[jj].y + tb.y * wj[jj].x;
                        const int i = i4 * 4 + jj;
                        zr[jj] = yr[i] * twr - yi[i] * twi;
                        zi[jj] = yr[i] * twi + yi[i] * twr;
                    }
                    split_pack2(zr[0], zr[1], zrh[i4 * 2], zrl[i4 * 2]);
                    split_pack2(zr[2], zr[3], zrh[i4 * 2 + 1], zrl[i4 * 2 + 1]);
                    split_pack2(zi[0], zi[1], zih[i4 * 2], zil[i4 * 2]);
                    split_pack2(zi[2], zi[3], zih[i4 * 2 + 1], zil[i4 * 2 + 1]);
                }
                mbar_wait(&empty2[s], (u & 1) ^ 1);
                uint8_t* dst = ring + s * kSlotBytes + k1 * 16;
#pragma unroll
                for (int kg = 0; kg < 2; ++kg) {
                    const uint4 vrh = make_uint4(zrh[kg * 4], zrh[kg * 4 + 1], zrh[kg * 4 + 2], zrh[kg * 4 + 3]);
                    const uint4 vrl = make_uint4(zrl[kg * 4], zrl[kg * 4 + 1], zrl[kg * 4 + 2], zrl[kg * 4 + 3]);
                    const uint4 vih = make_uint4(zih[kg * 4], zih[kg * 4 + 1], zih[kg * 4 + 2], zih[kg * 4 + 3]);
                    const uint4 vil = make_uint4(zil[kg * 4], zil[kg * 4 + 1], zil[kg * 4 + 2], zil[kg * 4 + 3]);
                    const uint32_t sg = 0x80008000u;                  // sign flip of both packed halves
                    const uint4 nrh = make_uint4(vrh.x ^ sg, vrh.y ^ sg, vrh.z ^ sg, vrh.w ^ sg);
                    const uint4 nrl = make_uint4(vrl.x ^ sg, vrl.y ^ sg, vrl.z ^ sg, vrl.w ^ sg);
                    uint8_t* d = dst + kg * 2048;
                    *reinterpret_cast<uint4*>(d + 0 * kA2ArrBytes) = vrh;
                    *reinterpret_cast<uint4*>(d + 1 * kA2ArrBytes) = vrl;
                    *reinterpret_cast<uint4*>(d + 2 * kA2ArrBytes) = vih;
                    *reinterpret_cast<uint4*>(d + 3 * kA2ArrBytes) = vil;
                    *reinterpret_cast<uint4*>(d + 4 * kA2ArrBytes) = nrh;
                    *reinterpret_cast<uint4*>(d + 5 * kA2ArrBytes) = nrl;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full2[s]);
            }

            // ---------------------------------------------------------------- power spectrum / complex output
            mbar_wait(d2_full, it & 1);
            tc_fence_after();
            float2* spec_row = nullptr;
            if (MODE == 1) spec_row = prm.spec + (static_cast<long long>(clip) * prm.n_frames + t) * kBins;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int k2_0 = 64 * h + 16 * cc;
                float re[16], im[16];
                tmem_ld16(tlane + 256 + k2_0, re);
                tmem_ld16(tlane + 384 + k2_0, im);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k2 = k2_0 + i;
                    const int k = k1 + 256 * k2;
                    int bin = -1;
                    float sgn = 1.f;
                    if (k <= kNfft / 2) bin = k;
                    else if (k1 >= 1) { bin = kNfft - k; sgn = -1.f; }       // Hermitian mirror (conjugate)
                    if (bin >= 0) {
                        if (MODE == 0) p_s[bin] = re[i] * re[i] + im[i] * im[i];
                        else spec_row[bin] = make_float2(re[i], sgn * im[i]);
                    }
                }
            }
            tc_fence_before();
            worker_sync();                                            // v_s complete, TMEM reads done
            // row k1 = 128: X[128 + 256 k2] = sum_n2 Y[n2,128] exp(-2 pi i n2 (2 k2 + 1)/256), k2 in [0,64)
            {
                const int k2 = tid >> 2;
                const int part = tid & 3;
                float ar = 0.f, ai = 0.f;
                const int m = 2 * k2 + 1;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const int n2 = part * 32 + i;
                    const float2 w = cs_s[(n2 * m) & 255];
                    const float v = v_s[n2];
                    ar = fmaf(v, w.x, ar);
                    ai = fmaf(v, w.y, ai);
                }
                ar += __shfl_xor_sync(0xffffffffu, ar, 1);
                ai += __shfl_xor_sync(0xffffffffu, ai, 1);
                ar += __shfl_xor_sync(0xffffffffu, ar, 2);
                ai += __shfl_xor_sync(0xffffffffu, ai, 2);
                if (part == 0) {
                    if (MODE == 0) p_s[128 + 256 * k2] = ar * ar + ai * ai;
                    else spec_row[128 + 256 * k2] = make_float2(ar, ai);
                }
            }
            if (MODE == 0) {
                worker_sync();                                        // power spectrum complete
                float* out_row = prm.out + (static_cast<long long>(clip) * prm.n_frames + t) * kMel;
                mel_db_from_smem(p_s, mel_tab_s, prm.mel_w, prm.norm, 1.0f, out_row, warp, lane);
            }
            worker_sync();                                            // ring (aliased by p_s) may be refilled
            if (lane == 0) mbar_arrive(ring_free);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// -------------------------------------------------------------------------------------------------------
// Un-fused drop-in for multichannel_complex_to_log_mel (preprocess.py:39-45): complex64 rows -> log-mel.
// One CTA per spectrogram row (frame); HBM-bound (131 KB read per row).
__global__ void __launch_bounds__(256) power_mel_db_kernel(const float2* __restrict__ spec, long long rows,
                                                           const float* __restrict__ mel_w,
                                                           const int4* __restrict__ mel_tab,
                                                           const float* __restrict__ norm, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* p_s = reinterpret_cast<float*>(smem);
    int4* mel_tab_s = reinterpret_cast<int4*>(smem + ((kBins * 4 + 15) / 16) * 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < kMel; i += 256) mel_tab_s[i] = mel_tab[i];
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float2* x = spec + row * kBins;
        for (int k = tid; k < kBins; k += 256) {
            const float2 v = x[k];
            p_s[k] = v.x * v.x + v.y * v.y;
        }
        __syncthreads();
        mel_db_from_smem(p_s, mel_tab_s, mel_w, norm, 1.0f, out + row * kMel, tid >> 5, tid & 31);
        __syncthreads();
    }
}

}  // namespace sedb
