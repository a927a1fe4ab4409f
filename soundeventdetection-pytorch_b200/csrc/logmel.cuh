// Fused framing + Hann window + real DFT (tcgen05 tensor cores) + |X|^2 + mel filterbank + log-dB.
//
// Replaces, for the reference's fixed configuration (dataset/common_config.py:2-8,
// dataset/spectogram/spectogram_configs.py:5-8), the composition
//     multichannel_complex_to_log_mel(multichannel_stft(x))      (dataset/spectogram/preprocess.py:21-45)
// without ever writing the (T x 16385) complex STFT to HBM.
//
// Algorithm (DESIGN.md "K1/K2"; numpy model in tests/dft_model.py::factored_power_spectrum_v2):
//   N = 32768 = 256 (n1) x 128 (n2), frame g[128 n1 + n2] = X[n1,n2], bin k = k1 + 256 k2, k1 in [0,128].
//   fold     CUDA  U[m] = X[m] + X[256-m], V[m] = X[m] - X[256-m]  (m = 1..127; U[0] = X[0], V[0] = 0)
//   stage 1  GEMM  Dc[k1,n2] = sum_m cos(2 pi k1 m/256) U[m,n2],  Ds[k1,n2] = sum_m -sin(2 pi k1 m/256) V[m,n2]
//                  (constants = A operand streamed by bulk copies, frame = B operand; K = 128)
//   twiddle  CUDA  Z = (Dc + (-1)^k1 X[128] + i Ds) exp(-2 pi i k1 n2/32768)
//   radix-2  CUDA  E[n] = Z[n] + Z[n+64],  O[n] = (Z[n] - Z[n+64]) exp(-2 pi i n/128)      n in [0,64)
//   stage 2  GEMM  X[k1 + 256 (2j)]   = sum_n E[k1,n] exp(-2 pi i n j/64)
//                  X[k1 + 256 (2j+1)] = sum_n O[k1,n] exp(-2 pi i n j/64)   (E/O = A operand, written back into the
//                  TMEM columns of the stage-1 accumulators and read from there; constants = B, resident in smem)
//   row 128  CUDA  X[128 + 256 k2] from Y[n2,128] = sum_m (-1)^m U[m,n2] + X[128,n2]
//   power, Hermitian bin map, mel filterbank in moment form (partial moments per <= 43-bin piece by the workers, the
//   64 filters + 10 log10(max(1e-10, .)) + store by the two helper warps).
// Warp roles: 16 workers (112 registers after setmaxnreg), MMA issuer, bulk-copy producer, 2 helpers (32 registers).
// A CTA owns a run of CONSECUTIVE frames; the workers request a frame's samples two 16-row chunks ahead of the fold
// (nothing of the frame is held in registers beyond that), and the fp16 block scale is taken from the half-frame shared
// with the previous frame, checked once the other half has been seen, and the stage-1 attempt repeated if it was too large.
// Input: fp32 mono (IN = 0) or interleaved 16-bit PCM with the channel mean fused into the loader (IN = 1, 2, 4, any).
// GEMM operands are split x = hi + lo (two 16-bit floats) and multiplied as hi*hi + lo*hi + hi*lo with fp32
// accumulation in TMEM: ~2^-17 relative with bf16 halves, ~2^-22 with fp16 halves (SEDB_SPLIT_FP16=1, which adds a
// per-frame power-of-two block scale so that the fp16 range is never exceeded).
#pragma once
#include "umma.cuh"
#include <type_traits>

// cp.async.bulk.prefetch.L2 of the next frame's new half by the copy warp (0 disables it for A/B measurements)
#ifndef SEDB_L2_PREFETCH
#define SEDB_L2_PREFETCH 1
#endif
// Development switches (defaults = the shipped configuration; measurements in profiles/r2_notes.md section 6):
//   SEDB_CONSUMER_FENCE  fence.proxy.async by the MMA warp after its wait instead of by the 512 producers (which drains
//                        their prefetched loads at every chunk)          SEDB_TAIL_FENCE  producer-side fence for the last chunk only
//   SEDB_KAHEAD          chunks of loads in flight ahead of the fold      SEDB_HROW_PIPE   window row factors fetched a chunk ahead
//   SEDB_TW_UNROLL       unroll of the twiddle loop (4 spills)            SEDB_MMA_UNROLL  unroll of the MMA warp's chunk loop
//   SEDB_ROW128_FOLD     row-128 sums over folded inputs (64 terms)       SEDB_ROW128_EARLY  ... run in the stage-1 MMA drain
//   SEDB_INCR_FRAME      incremental (clip, frame) instead of a 64-bit division per frame
//   SEDB_RELAX_NS        nanosleep between the polls of the copy / helper warps
//   -DSEDB_PROF_FOLD / -DSEDB_PROF_MMA / -DSEDB_DEBUG_SCALE  extra in-kernel counters and a per-frame printf
#ifndef SEDB_CONSUMER_FENCE
#define SEDB_CONSUMER_FENCE 1
#endif
#ifndef SEDB_TAIL_FENCE
#define SEDB_TAIL_FENCE 0
#endif
#ifndef SEDB_TW_UNROLL
#define SEDB_TW_UNROLL 2
#endif
#ifndef SEDB_RELAX_NS
#define SEDB_RELAX_NS 200
#endif
#ifndef SEDB_INCR_FRAME
#define SEDB_INCR_FRAME 1
#endif
#ifndef SEDB_ROW128_FOLD
#define SEDB_ROW128_FOLD 1
#endif
#ifndef SEDB_ROW128_EARLY
#define SEDB_ROW128_EARLY 0
#endif
#ifndef SEDB_MMA_UNROLL
#define SEDB_MMA_UNROLL 1
#endif
#ifndef SEDB_KAHEAD
#define SEDB_KAHEAD 2
#endif
#ifndef SEDB_HROW_PIPE
#define SEDB_HROW_PIPE 1
#endif

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
constexpr int kB2ArrBytes = 128 * 64 * 2;                  // one resident stage-2 constant array [N=128][K=64] (16 KB)
constexpr int kB2Bytes = 4 * kB2ArrBytes;                  // breH | breL | bimH | bimL
constexpr int kA1ArrBytes = 128 * 16 * 2;                  // one stage-1 constant array per K-chunk (4 KB)
constexpr int kA1ChunkBytes = 4 * kA1ArrBytes;             // cH | cL | sH | sL  (16 KB)
constexpr int kB1Sbo = 144;                                // padded stride between 8-sample groups (bank spread)
constexpr int kB1Lbo = 16 * kB1Sbo;                        // 2304: stride between groups of 8 rows (K)
constexpr int kB1ArrBytes = 2 * kB1Lbo;                    // 4608
constexpr int kB1SlotBytes = 4 * kB1ArrBytes;              // 18432: UH UL VH VL of one stage-1 chunk
constexpr int kNumSlots = 4;
constexpr int kA1RingBytes = kA1ChunkBytes * kNumSlots;    // 65536: stage-1 constant chunks (bulk copies, run ahead of the frame)
constexpr int kB1RingBytes = kB1SlotBytes * kNumSlots;     // 73728: stage-1 data operand chunks; the power spectrum aliases it
constexpr int kRingBytes = kA1RingBytes + kB1RingBytes;    // 139264
constexpr int kA2ArrBytes = 128 * 16 * 2;                  // 4 KB
static_assert((kNumSlots & (kNumSlots - 1)) == 0, "slot index uses a mask");

constexpr int kOffB2 = 0;
constexpr int kOffRing = kOffB2 + kB2Bytes;                // 65536
constexpr int kOffAlt = kOffRing + kRingBytes;             // per-row alternating partial sums  [16][128] floats
constexpr int kOffX128 = kOffAlt + 16 * 128 * 4;           // X[128, n2]      128 floats
constexpr int kOffV = kOffX128 + 512;                      // Y[n2,128]       128 floats
constexpr int kOffCs = kOffV + 512;                        // exp(-2 pi i j/256) 256 float2
constexpr int kMelMaxPieces = 576;                         // host_tables.h: make_mel_moment_tables
constexpr int kMelPieceLen = 43;                           // host_tables.h: longest piece (checked in sedb.cu)
constexpr int kMelTabEntries = kMelMaxPieces / 2 + 80;
constexpr int kOffMelTab = kOffCs + 2048;                  // segment table (320 x int4)
constexpr int kOffMelPart = kOffMelTab + kMelTabEntries * 16;   // partial moments (2 x kMelMaxPieces floats)
constexpr int kOffMelCoef = kOffMelPart + 2 * kMelMaxPieces * 4;   // 64 x float4 line coefficients
constexpr int kOffNorm = kOffMelCoef + 1024;               // mean[64], std[64] (when given)
constexpr int kOffRed = kOffNorm + 512;                    // [0,16) abs-max scratch (second half), [20] frame scale for the helpers,
                                                           // [24] redo flag, [32,48) abs-max scratch (first half)
constexpr int kOffHannRow = kOffRed + 192;                 // window factors per row m: {sin, cos}(pi (128 m - 544)/31679), 257 x float2
constexpr int kOffHannLane = kOffHannRow + 2064;           // window factors per column n2: cos[128] then sin[128] of pi n2/31679
constexpr int kOffE1Tw = kOffHannLane + 1024;              // exp(-2 pi i n/128), n < 64: re[64] then im[64]
constexpr int kOffBars = kOffE1Tw + 512;                   // mbarriers
constexpr int kOffTmem = kOffBars + 256;                   // tmem base address
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kB1RingBytes >= (kBins + 3) * 4, "power spectrum must fit in the data operand ring");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr int kWorkerWarps = 16;
constexpr int kWorkerThreads = kWorkerWarps * 32;
constexpr int kMmaWarp = 16;
constexpr int kCopyWarp = 17;
constexpr int kThreads = 640;                // 16 workers + MMA + copy + 2 helper warps (register allocation is per 4 warps anyway)
constexpr int kWorkerRegs = 112;             // setmaxnreg: the service warp group gives its registers to the workers
constexpr int kServiceRegs = 32;              // (4 x 32 + 16 x 112) x 32 = the 96 x 640 registers of the launch

struct LogmelParams {
    const float* wave;        // IN 0: [B, wave_stride] fp32 mono
    const int16_t* pcm;       // IN 1: [B, wave_stride, n_channels] interleaved 16-bit PCM (WAV data order); the kernel
                              //       forms the mono mix mean_ch(s / 32768) (dataset_utils.py:67-74 with audio_channels = 1)
    int n_channels;           // IN 1 only
    float pcm_scale;          // IN 1 only: 1 / (32768 n_channels)
    long long wave_stride;    // samples (sample frames) between clips
    int n_samples;            // valid samples per clip
    int n_clips;
    int n_frames;             // T = 1 + n_samples / hop
    const uint8_t* a1;        // stage-1 constants, 8 chunks x 16 KB, canonical K-major, split hi/lo
    const uint8_t* b2;        // stage-2 constants, 4 x 16 KB
    const float* hann;        // factored np.hanning(31680): sin^2(pi i/31679), i = 128 m + n2 - 544 (host_tables.h: make_hann_factors)
    const float* mel_w;       // 64 x {ar, br, af, bf}: line coefficients of the moment form (host_tables.h)
    const int4* mel_tab;      // balanced piece table + per-segment slot ranges (host_tables.h)
    const float* norm;        // nullable: mean[64] then std[64]  (SpectogramDataset.transform, logMel mode)
    float* out;               // MODE 0: [B, T, 64] fp32 log-mel
    float2* spec;             // MODE 1: [B, T, 16385] complex64 STFT
    unsigned long long* prof; // nullable: per-phase cycle counters (diagnostics)
};

// named barrier among the worker threads only
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ int reflect_index(int j, int n) {
    // np.pad(mode='reflect') for a pad shorter than the signal: -1 -> 1, n -> n-2
    if (j < 0) j = -j;
    if (j >= n) j = 2 * (n - 1) - j;
    return j;
}

// 8 consecutive 32-bit TMEM columns of this warp's 32 lanes
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float* v) {
    uint32_t r[2];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r[0]);
    v[1] = __uint_as_float(r[1]);
}
__device__ __forceinline__ void split4(const float2* x, uint2& hi, uint2& lo) {
    split_pack2(x[0], hi.x, lo.x);
    split_pack2(x[1], hi.y, lo.y);
}

// Mel filterbank over a power spectrum in shared memory, in moment form (host_tables.h): per segment between two mel
// points S0 = sum P_k and S1 = sum (k - kb) P_k.  The bins are cut into pieces of <= kMelPieceLen bins; thread `idx`
// accumulates the pieces at table positions idx, idx + nthreads, ... (no cross-lane reduction; the positions of a warp
// start on different banks where the host table could arrange it) and writes the partial moments to the piece's slot.
// The slots of a segment are then added in order by mel_finalize(), so the result does not depend on scheduling.
__device__ __forceinline__ void mel_partials(const float* __restrict__ p_s, const int4* __restrict__ tab_s,
                                             float* __restrict__ part_s, int idx, int nthreads) {
    const int2* pieces = reinterpret_cast<const int2*>(tab_s);
    for (int i = idx; i < kMelMaxPieces; i += nthreads) {
        const int2 e = pieces[i];
        const int k0 = e.x & 0xffff, len = e.x >> 16;
        if (len == 0) continue;                                      // unused thread position
        const int slot = e.y >> 16;
        const float* p = p_s + k0;
        // straight-line over the maximum piece length (bins past the piece are read and discarded: they are still inside
        // the spectrum's buffer) in two batches, so that the loads of a batch are in flight before its first add waits
        // without holding 43 registers; two accumulator pairs; sum (k - kb) P = (k0 - kb) sum P + sum j P with compile-time j
        const float kf0 = static_cast<float>(k0 - (e.y & 0xffff));
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        constexpr int kHalf = (kMelPieceLen + 1) / 2;
#pragma unroll
        for (int h = 0; h < kMelPieceLen; h += kHalf) {
            float v[kHalf];
#pragma unroll
            for (int j = 0; j < kHalf; ++j) v[j] = (h + j < kMelPieceLen && h + j < len) ? p[h + j] : 0.f;
#pragma unroll
            for (int j = 0; j < kHalf; ++j) {
                if (h + j < kMelPieceLen) {
                    if ((j & 1) == 0) {
                        a0 += v[j];
                        b0 = fmaf(static_cast<float>(h + j), v[j], b0);
                    } else {
                        a1 += v[j];
                        b1 = fmaf(static_cast<float>(h + j), v[j], b1);
                    }
                }
            }
        }
        const float s0 = a0 + a1;
        part_s[slot] = s0;
        part_s[kMelMaxPieces + slot] = fmaf(kf0, s0, b0 + b1);
    }
}
// TPF threads per filter (64 * TPF threads in total): thread (m, sub) adds every TPF-th partial slot of segments m and
// m+1, the TPF lanes are combined with a fixed shuffle tree, and lane sub == 0 converts to dB, normalises and stores.
template <int TPF>
__device__ __forceinline__ void mel_finalize(const float* __restrict__ part_s, const int4* __restrict__ tab_s,
                                             const float* __restrict__ coef_s, const float* __restrict__ norm,
                                             float inv_scale2, float* __restrict__ out_row, int tid) {
    const int m = tid / TPF, sub = tid % TPF;
    float m0[2] = {0.f, 0.f}, m1[2] = {0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const int4 f = tab_s[kMelMaxPieces / 2 + m + d];
        for (int i = sub; i < f.y; i += TPF) {
            m0[d] += part_s[f.x + i];
            m1[d] += part_s[kMelMaxPieces + f.x + i];
        }
    }
#pragma unroll
    for (int o = 1; o < TPF; o <<= 1) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            m0[d] += __shfl_xor_sync(0xffffffffu, m0[d], o);
            m1[d] += __shfl_xor_sync(0xffffffffu, m1[d], o);
        }
    }
    if (sub == 0) {
        const float4 cf = reinterpret_cast<const float4*>(coef_s)[m];
        const float a = fmaf(cf.x, m1[0], cf.y * m0[0]) + fmaf(cf.z, m1[1], cf.w * m0[1]);
        float db = 10.0f * log10f(fmaxf(1e-10f, a * inv_scale2));        // librosa.power_to_db(ref=1, amin=1e-10)
        if (norm != nullptr) db = (db - norm[m]) / norm[kMel + m];       // spectograms_dataset.py:105
        out_row[m] = db;
    }
}

#define SEDB_PROF(i)                                                             \
    do {                                                                         \
        if (prm.prof != nullptr && tid == 0) {                                   \
            const long long now__ = clock64();                                   \
            atomicAdd(prm.prof + (i), static_cast<unsigned long long>(now__ - tprev)); \
            tprev = now__;                                                       \
        }                                                                        \
    } while (0)

// ---- frame loads ----------------------------------------------------------------------------------------
// The loads of a frame are issued a few chunks ahead of their use (see the worker loop).  They are volatile asm so
// that neither nvcc nor ptxas gathers them at the top of the frame again (64 live registers and an LSU queue that
// back-pressures every shared-memory access behind it).
__device__ __forceinline__ float4 ldg_nc_f4(const void* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_nc_u2(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// 4 consecutive sample frames of interleaved 16-bit PCM (C in {1, 2, 4}: 8 C bytes, aligned) -> mono floats.  The raw
// load and the conversion are separate so that the loads stay in flight while earlier chunks are processed.
__device__ __forceinline__ float s16lo(uint32_t w) { return static_cast<float>(static_cast<short>(w & 0xffffu)); }
__device__ __forceinline__ float s16hi(uint32_t w) { return static_cast<float>(static_cast<int>(w) >> 16); }
template <int C>
struct PcmRaw {
    uint32_t w[2 * C];
};
template <int C>
__device__ __forceinline__ PcmRaw<C> pcm_zero_raw() {
    PcmRaw<C> r;
#pragma unroll
    for (int i = 0; i < 2 * C; ++i) r.w[i] = 0u;
    return r;
}
template <int C>
__device__ __forceinline__ PcmRaw<C> pcm_load_raw(const int16_t* __restrict__ p) {
    PcmRaw<C> r;
    if (C == 1) {
        const uint2 v = ldg_nc_u2(p);
        r.w[0] = v.x; r.w[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < C / 2; ++q) {
            const uint4 v = ldg_nc_u4(reinterpret_cast<const uint4*>(p) + q);
            r.w[4 * q] = v.x; r.w[4 * q + 1] = v.y; r.w[4 * q + 2] = v.z; r.w[4 * q + 3] = v.w;
        }
    }
    return r;
}
template <int C>
__device__ __forceinline__ float4 pcm_to_mono4(const PcmRaw<C>& r, float scale) {
    float v[4];
    if (C == 1) {
        v[0] = s16lo(r.w[0]); v[1] = s16hi(r.w[0]); v[2] = s16lo(r.w[1]); v[3] = s16hi(r.w[1]);
    } else if (C == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = s16lo(r.w[e]) + s16hi(r.w[e]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            v[e] = (s16lo(r.w[2 * e]) + s16hi(r.w[2 * e])) + (s16lo(r.w[2 * e + 1]) + s16hi(r.w[2 * e + 1]));
    }
    return make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
}

constexpr int kInPcmAny = 16;   // IN value: 16-bit PCM with a run-time channel count (scalar loads)

// MODE 0: log-mel output; 1: complex STFT output.  IN 0: fp32 mono; 1, 2, 4: 16-bit PCM with that many interleaved
// channels (vector loads); kInPcmAny: 16-bit PCM, any channel count.
template <int MODE, int IN>
__global__ void __launch_bounds__(kThreads, 1) logmel_fused_kernel(const LogmelParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* b2_s = smem + kOffB2;
    uint8_t* a1ring = smem + kOffRing;
    uint8_t* b1ring = a1ring + kA1RingBytes;
    float* p_s = reinterpret_cast<float*>(b1ring);
    float* alt_s = reinterpret_cast<float*>(smem + kOffAlt);
    float* x128_s = reinterpret_cast<float*>(smem + kOffX128);
    float* v_s = reinterpret_cast<float*>(smem + kOffV);
    float2* cs_s = reinterpret_cast<float2*>(smem + kOffCs);
    int4* mel_tab_s = reinterpret_cast<int4*>(smem + kOffMelTab);
    float* part_s = reinterpret_cast<float*>(smem + kOffMelPart);
    float* coef_s = reinterpret_cast<float*>(smem + kOffMelCoef);
    float* norm_s = reinterpret_cast<float*>(smem + kOffNorm);
    float* red_s = reinterpret_cast<float*>(smem + kOffRed);
    float2* hrow_s = reinterpret_cast<float2*>(smem + kOffHannRow);
    float* hlane_s = reinterpret_cast<float*>(smem + kOffHannLane);
    float* e1tw_s = reinterpret_cast<float*>(smem + kOffE1Tw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + kOffTmem);

    uint64_t* full1 = bars + 0;      // [4] stage-1 slot filled (16 worker warps + 1 bulk copy)
    uint64_t* empty1 = bars + 4;     // [4] stage-1 slot consumed (tcgen05.commit)
    uint64_t* full2 = bars + 8;      // [4] stage-2 slot filled (4 worker warps)
    uint64_t* d1_full = bars + 12;   // stage-1 accumulators complete
    uint64_t* d2_full = bars + 13;   // stage-2 accumulators complete
    uint64_t* b2_full = bars + 15;   // resident stage-2 constants landed
    // log-mel mode: the two helper warps take the mel finalize off the workers
    uint64_t* part_full = bars + 17; // mel partial moments (and the frame's scale) are in shared memory (16 worker warps)
    uint64_t* part_free = bars + 18; // helpers have finished the previous frame's finalize
    uint64_t* dec = bars + 19;       // fp16 build: the workers have decided whether the stage-1 attempt stands (redo flag in red_s[24])

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kNumSlots; ++s) {
            mbar_init(&full1[s], kWorkerWarps + 1);
            mbar_init(&empty1[s], 1);
            mbar_init(&full2[s], kWorkerWarps);
        }
        mbar_init(d1_full, 1);
        mbar_init(d2_full, 1);
        mbar_init(b2_full, 1);
        mbar_init(part_full, kWorkerWarps);
        mbar_init(part_free, 2);
        mbar_init(dec, 1);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) tmem_alloc<512>(tmem_ptr_s);
    for (int i = tid; i < 256; i += kThreads) {
        float s, c;
        sincospif(static_cast<float>(i) * (1.0f / 128.0f), &s, &c);      // exp(-2 pi i j/256) = (c, -s)
        cs_s[i] = make_float2(c, -s);
        if ((i & 1) == 0 && i < 128) {
            e1tw_s[i >> 1] = c;
            e1tw_s[64 + (i >> 1)] = -s;
        }
    }
    for (int i = tid; i < 2 * 257 + 256; i += kThreads) {
        const float v = prm.hann[i];
        if (i < 2 * 257) reinterpret_cast<float*>(hrow_s)[i] = v; else hlane_s[i - 2 * 257] = v;
    }
    for (int i = tid; i < kMelTabEntries; i += kThreads) mel_tab_s[i] = prm.mel_tab[i];
    for (int i = tid; i < 4 * kMel; i += kThreads) coef_s[i] = prm.mel_w[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;
    // Set-up done (tables above are the context's own constants).  Launched with programmatic stream serialization, the
    // kernel may have been placed before its predecessor finished: wait for it before touching the caller's buffers.
    pdl_entry();
    if (prm.norm != nullptr) {
        for (int i = tid; i < 2 * kMel; i += kThreads) norm_s[i] = prm.norm[i];
        __syncthreads();
    }

    // Frame schedule: CTA b owns the CONSECUTIVE frames [f0, f0 + n_iter) of the flattened (clip, frame) sequence.  Two
    // consecutive frames of a clip share half their samples (hop = window / 2), so the second read of a sample hits L2
    // right after the first, and -- fp16 build -- the abs-max of the shared half is known before the frame is loaded,
    // which lets the fold start on the first samples that arrive (see "block scale" in the worker loop).
    const long long total_frames = static_cast<long long>(prm.n_clips) * prm.n_frames;
    const long long per_cta = total_frames / gridDim.x;
    const int rem_cta = static_cast<int>(total_frames - per_cta * gridDim.x);
    const long long f0 = per_cta * blockIdx.x + min(static_cast<int>(blockIdx.x), rem_cta);
    const int n_iter = static_cast<int>(per_cta) + (static_cast<int>(blockIdx.x) < rem_cta ? 1 : 0);
    volatile int* redo_s = reinterpret_cast<volatile int*>(red_s + 24);
    (void)redo_s;

    // ======================================================================== bulk-copy producer warp
    // register re-allocation between warp groups: 4 service warps x 32 regs + 16 worker warps x 112 regs = the 96 x 640
    // registers the CTA was launched with
    // (issued at the top of each role's own branch: ptxas allocates per region between setmaxnreg and the join)
    if (warp >= kWorkerWarps + 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kServiceRegs));
        // ==================================================================== helper warps (log-mel mode): the mel
        // finalize of every frame (partial moments of the workers -> 64 filters -> dB -> store) runs here, off the
        // workers' critical path: 8 passes of 8 filters x 8 lanes.
        if (MODE == 0) {
            const int hid = tid - (kWorkerWarps + 2) * 32;            // 0..63
            for (int it = 0; it < n_iter; ++it) {
                mbar_wait_relaxed(part_full, it & 1, SEDB_RELAX_NS);
                const long long f = f0 + it;
                float* out_row = prm.out + f * kMel;                  // (clip * T + t) * 64 = f * 64
                const float inv_scale2 = red_s[20];
#pragma unroll 1
                for (int pass = 0; pass < 8; ++pass)
                    mel_finalize<8>(part_s, mel_tab_s, coef_s, prm.norm != nullptr ? norm_s : nullptr, inv_scale2, out_row,
                                    pass * 64 + hid);
                __syncwarp();
                if (lane == 0) mbar_arrive(part_free);
            }
        }
    } else if (warp == kCopyWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kServiceRegs));
        if (elect_one()) {
            mbar_arrive_expect_tx(b2_full, kB2Bytes);
            for (int a = 0; a < 4; ++a)
                bulk_g2s(b2_s + a * kB2ArrBytes, prm.b2 + a * kB2ArrBytes, kB2ArrBytes, b2_full);
        }
        __syncwarp();
        // One pass over the 8 constant chunks per stage-1 ATTEMPT (8 chunks = two revolutions of the 4-slot ring, so slot
        // and parity of chunk c are the same in every attempt).  fp16 build: a frame whose provisional block scale turns
        // out too large is folded again (worker loop), which consumes another 8 chunks; the workers' decision arrives on
        // `dec` right after the fold, long before the next frame needs its first constants.
        int attempt = 0;
        (void)attempt;                                   // (bf16 build: no block scale, no second attempts)
        for (int it = 0; it < n_iter;) {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                const int s = c & (kNumSlots - 1);
                const int u = c / kNumSlots;
                mbar_wait_relaxed(&empty1[s], (u & 1) ^ 1, SEDB_RELAX_NS);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full1[s], kA1ChunkBytes);
                    bulk_g2s(a1ring + s * kA1ChunkBytes, prm.a1 + c * kA1ChunkBytes, kA1ChunkBytes, &full1[s]);
                }
                __syncwarp();
            }
            // pull the new half of the next frame towards L2 while this one is being processed (the other half is this
            // frame's second half; a clip's first frame takes everything up to the end of its window)
            if (SEDB_L2_PREFETCH && it + 1 < n_iter) {
                const long long f = f0 + it + 1;
                const int clip = static_cast<int>(f / prm.n_frames);
                const int t = static_cast<int>(f - static_cast<long long>(clip) * prm.n_frames);
                const long long lo = (t == 0) ? 0 : static_cast<long long>(t) * kHop;
                long long hi = static_cast<long long>(t) * kHop + kHop;
                if (hi > prm.n_samples) hi = prm.n_samples;
                const long long bytes_per = (IN == 0) ? 4 : 2 * prm.n_channels;
                const char* src = (IN == 0)
                    ? reinterpret_cast<const char*>(prm.wave + static_cast<long long>(clip) * prm.wave_stride + lo)
                    : reinterpret_cast<const char*>(prm.pcm + (static_cast<long long>(clip) * prm.wave_stride + lo) * prm.n_channels);
                const long long nbytes = ((hi - lo) * bytes_per) & ~15LL;
                if (nbytes > 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                    if (elect_one())
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(static_cast<uint32_t>(nbytes)) : "memory");
                    __syncwarp();
                }
            }
#if SEDB_SPLIT_FP16
            mbar_wait_relaxed(dec, attempt & 1, SEDB_RELAX_NS);
            ++attempt;
            if (*redo_s == 0) ++it;
#else
            ++it;
#endif
        }
    }
    // ======================================================================== MMA issuer warp
    else if (warp == kMmaWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kServiceRegs));
        // The whole warp stays converged and one elected lane issues: every operand is then warp-uniform (the
        // 512-column allocation starts at TMEM address 0) and each MMA costs a handful of uniform-datapath
        // instructions.  A lane-divergent `if (lane == 0)` region makes ptxas wrap every tcgen05.mma in a vote loop.
        if (tmem != 0) __trap();
        constexpr uint32_t idesc1 = make_idesc(kSplitFmt, kMajorK, kMajorMN, 128, 128);
        constexpr uint32_t idesc2 = make_idesc(kSplitFmt, kMajorK, kMajorK, 128, 128);
        const uint32_t b2_a = smem_u32(b2_s);
        // descriptors for slot 0 / K-chunk 0; other slots and chunks add to the start-address field (16-byte units)
        const uint64_t dA1 = make_smem_desc(smem_u32(a1ring), 2048, 128);               // cH; cL, sH, sL follow
        const uint64_t dB1 = make_smem_desc(smem_u32(b1ring), kB1Lbo, kB1Sbo);          // uH; uL, vH, vL follow
        const uint64_t dB2 = make_smem_desc(b2_a, 2048, 128);                           // reH; reL, imH, imL follow
        constexpr uint32_t kA1Step = kA1ArrBytes >> 4, kB1Step = kB1ArrBytes >> 4;
        constexpr uint32_t kB2Step = kB2ArrBytes >> 4, kA1SlotStep = kA1ChunkBytes >> 4, kB1SlotStep = kB1SlotBytes >> 4;
        mbar_wait(b2_full, 0);
        int attempt = 0;
        (void)attempt;
        for (int it = 0; it < n_iter; ++it) {
            // ---------------- stage 1: 8 K-chunks of 16 folded rows (m); repeated when the workers reject the attempt
            for (;;) {
                constexpr int kMmaUnroll = SEDB_MMA_UNROLL;          // 8: slot, parity and descriptors are immediates
#pragma unroll kMmaUnroll
                for (int c = 0; c < 8; ++c) {
                    const int s = c & (kNumSlots - 1);
                    const int u = c / kNumSlots;
#ifdef SEDB_PROF_MMA
                    long long tm0 = clock64();
#endif
                    mbar_wait(&full1[s], u & 1);
#ifdef SEDB_PROF_MMA
                    long long tm1 = clock64();
                    if (prm.prof != nullptr && lane == 0) atomicAdd(prm.prof + (c == 7 ? 13 : 12), static_cast<unsigned long long>(tm1 - tm0));
#endif
#if SEDB_CONSUMER_FENCE
                    if (!SEDB_TAIL_FENCE || c < 7) fence_proxy_async_smem();
#endif
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a = dA1 + s * kA1SlotStep, bb = dB1 + s * kB1SlotStep;
                        const uint32_t acc = (c > 0) ? 1u : 0u;
                        umma_f16(0, a, bb, idesc1, acc);                                   // cH uH
                        umma_f16(0, a + kA1Step, bb, idesc1, 1u);                          // cL uH
                        umma_f16(0, a, bb + kB1Step, idesc1, 1u);                          // cH uL
                        umma_f16(128, a + 2 * kA1Step, bb + 2 * kB1Step, idesc1, acc);     // sH vH
                        umma_f16(128, a + 3 * kA1Step, bb + 2 * kB1Step, idesc1, 1u);      // sL vH
                        umma_f16(128, a + 2 * kA1Step, bb + 3 * kB1Step, idesc1, 1u);      // sH vL
                        umma_commit(&empty1[s]);
                        if (c == 7) umma_commit(d1_full);
                    }
                    __syncwarp();
#ifdef SEDB_PROF_MMA
                    if (prm.prof != nullptr && lane == 0) atomicAdd(prm.prof + 14, static_cast<unsigned long long>(clock64() - tm1));
#endif
                }
#if SEDB_SPLIT_FP16
                mbar_wait(dec, attempt & 1);
                ++attempt;
                if (*redo_s == 0) break;
#else
                break;
#endif
            }
            // ---------------- stage 2: 4 K-chunks of 16 columns (n), even and odd outputs; A operand from TMEM
            constexpr int kMmaUnroll2 = SEDB_MMA_UNROLL >= 4 ? 4 : 1;
#pragma unroll kMmaUnroll2
            for (int j = 0; j < 4; ++j) {
                mbar_wait(&full2[j], it & 1);
                tc_fence_after();
                if (elect_one()) {
                    // E: re hi/lo at columns 16 j / 16 j + 8, im hi/lo at 64 + ...; O: 128 + ..., 192 + ...
                    const uint32_t a = 16 * j;
                    const uint64_t bre = dB2 + j * ((2 * 2048) >> 4);  // 16 n = 2 K-groups
                    const uint64_t bim = bre + 2 * kB2Step;
                    const uint32_t acc = (j > 0) ? 1u : 0u;
#pragma unroll
                    for (int par = 0; par < 2; ++par) {
                        const uint32_t d = 256 + 128 * par;
                        const uint32_t ap = a + 128 * par;
                        umma_f16_ts(d, ap, bre, idesc2, acc);                  // reH breH
                        umma_f16_ts(d, ap + 8, bre, idesc2, 1u);               // reL breH
                        umma_f16_ts(d, ap, bre + kB2Step, idesc2, 1u);         // reH breL
                        umma_f16_ts(d, ap + 64, bim, idesc2, 1u);              // imH bimH
                        umma_f16_ts(d, ap + 64 + 8, bim, idesc2, 1u);          // imL bimH
                        umma_f16_ts(d, ap + 64, bim + kB2Step, idesc2, 1u);    // imH bimL
                    }
                    if (j == 3) umma_commit(d2_full);
                }
                __syncwarp();
            }
        }
    }
    // ======================================================================== 16 worker warps
    else if (warp < kWorkerWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWorkerRegs));
        const int q = warp & 3;                 // TMEM lane quarter
        const int sub = warp >> 2;              // which 4 of a chunk's 16 columns (stage-2 operand), which 16 of 64 (output)
        const int k1 = q * 32 + lane;           // this thread's stage-1 output row / stage-2 A row
        const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const float sgn_k1 = (k1 & 1) ? -1.f : 1.f;

        // per-thread twiddle constants, W = exp(-2 pi i k1/32768): W^j (j=1..3), anchors W^(16 c + 4 sub), W^64
        // (the j-twiddles are kept as pairs (j, j+1) for packed arithmetic; W^0 = 1)
        float2 wre01, wim01, wre23, wim23, anc[4], w64;
        {
            float s, c;
            float2 wj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sincospif(static_cast<float>(k1 * ((j & 1) + 8 * (j >> 1))) * (1.0f / 16384.0f), &s, &c);
                wj[j] = make_float2(c, -s);
                sincospif(static_cast<float>(k1 * (16 * j + 2 * sub)) * (1.0f / 16384.0f), &s, &c);
                anc[j] = make_float2(c, -s);
            }
            wre01 = f2(1.f, wj[1].x); wim01 = f2(0.f, wj[1].y);
            wre23 = f2(wj[2].x, wj[3].x); wim23 = f2(wj[2].y, wj[3].y);
            sincospif(static_cast<float>(k1 * 64) * (1.0f / 16384.0f), &s, &c);
            w64 = make_float2(c, -s);
        }

        // stage-1 producer mapping: warp -> row r of the 16-row chunk, lane -> 4 consecutive samples
        const int r = warp;
        const uint32_t b1_off = (lane >> 1) * kB1Sbo + (r >> 3) * kB1Lbo + (r & 7) * 16 + (lane & 1) * 8;
        const float alt_sign = (r & 1) ? -1.f : 1.f;

        // shared-window address of the power spectrum, opaque to the optimiser so that it stays in a register
        uint32_t p_s_addr;
        asm volatile("mov.u32 %0, %1;" : "=r"(p_s_addr) : "r"(smem_u32(p_s)));
        // raw form of four consecutive samples between the load and its use: the loaded words themselves
        constexpr int PC = (IN == 1 || IN == 2 || IN == 4) ? IN : 0;          // vector PCM loader
        using Raw = typename std::conditional<PC != 0, PcmRaw<(PC != 0 ? PC : 1)>, float4>::type;
        constexpr int kAhead = (PC == 4) ? 1 : SEDB_KAHEAD;                                             // chunks in flight ahead of the fold

        // fp16 build: exponent of the block scale that one half of a frame would get on its own: 2^e |x| < 64 with e
        // rounded down to even (sqrt(scale) rides on the window factors); silence gets the largest scale
        auto ebucket = [](float mx) -> int {
            int e = 60;
            if (mx > 0.f) e = 5 - (static_cast<int>((__float_as_uint(mx) >> 23) & 0xff) - 127);
            if (!(mx < 3.0e38f)) e = 0;                                       // inf / nan: nothing to preserve
            return max(-56, min(60, e)) & ~1;
        };
        auto absmax4 = [](float m, const float4& v) {
            return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        };
        int e_hist = 0;                         // ebucket of the previous frame's second half = this frame's first half
        int attempt = 0;                        // stage-1 attempts so far (parity of d1_full / dec)

        // ---------------------------------------------------------------- frame loader
        // chunk c, this thread: rows m = 16 c + r (first half of the frame) and 256 - m (second half; row 128 for
        // m = 0), samples n2 = 4 lane .. +3.  Interior frames (every in-window sample inside the clip, aligned rows:
        // all but the first and last frames of a clip) load vectors off two base pointers; only chunk 0 touches rows
        // that are partly outside the window.  Edge frames take the general path (reflect padding, scalar loads).
        const int L = prm.n_samples;
        const int na = 128 * r + 4 * lane, nb = 128 * (r == 0 ? 128 : 256 - r) + 4 * lane;
        const bool live_a0 = na >= kLpad, live_b0 = nb < kLpad + kWin;
        const int C = (IN == 0) ? 1 : (IN == kInPcmAny ? prm.n_channels : IN);
        const float ps = prm.pcm_scale;
        const int fbytes = (IN == 0) ? 4 : 2 * C;                               // bytes per sample frame
        const int pb_off = 128 * (256 - 2 * r) * fbytes;                        // row 256 - r relative to row r, bytes

        long long tprev = clock64();
#if SEDB_INCR_FRAME
        int clip = static_cast<int>(f0 / prm.n_frames);
        int t = static_cast<int>(f0 - static_cast<long long>(clip) * prm.n_frames) - 1;
#endif
        for (int it = 0; it < n_iter; ++it) {
#if SEDB_INCR_FRAME
            if (++t == prm.n_frames) {                                          // (no 64-bit division per frame)
                t = 0;
                ++clip;
            }
#else
            const long long f = f0 + it;
            const int clip = static_cast<int>(f / prm.n_frames);
            const int t = static_cast<int>(f - static_cast<long long>(clip) * prm.n_frames);
#endif
            const bool inside = t >= 1 && static_cast<long long>(t) * kHop + (kWin / 2) <= L;
            const char* ybase;                                                  // clip start
            if (IN == 0) ybase = reinterpret_cast<const char*>(prm.wave + static_cast<long long>(clip) * prm.wave_stride);
            else ybase = reinterpret_cast<const char*>(prm.pcm + static_cast<long long>(clip) * prm.wave_stride * C);
            const bool fast = inside && IN != kInPcmAny && ((reinterpret_cast<uintptr_t>(ybase) & 15) == 0);
            // row r of the frame, this lane's samples (row 16 c + r: + 2048 c sample frames; row 256 - 16 c - r: pb_off - 2048 c)
            const char* pa = ybase + (static_cast<long long>(t) * kHop - kPadRefl + 4 * lane + 128 * r) * fbytes;
            // general path: 4 sample frames from frame position n0, zero outside the window, reflect outside the clip
            auto load_general = [&](int n0) -> Raw {
                // (opaque to the optimiser: otherwise the reflected addresses of all 64 scalar loads of an edge frame are
                // formed at the top of the frame and spill the workers' registers in every frame)
                asm volatile("" : "+r"(n0));
                const int j0 = t * kHop + n0 - kPadRefl;
                const bool dead = n0 < kLpad - 3 || n0 >= kLpad + kWin;           // window is zero
                if constexpr (PC != 0) {
                    Raw q = pcm_zero_raw<PC>();
                    if (!dead) {
                        const int16_t* y = reinterpret_cast<const int16_t*>(ybase);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int16_t* src = y + static_cast<long long>(reflect_index(j0 + e, L)) * PC;
#pragma unroll
                            for (int ch = 0; ch < PC; ++ch) {
                                const uint32_t v = static_cast<uint16_t>(__ldg(src + ch));
                                q.w[(e * PC + ch) >> 1] |= v << (16 * ((e * PC + ch) & 1));
                            }
                        }
                    }
                    return q;
                } else if constexpr (IN == 0) {
                    if (dead) return make_float4(0.f, 0.f, 0.f, 0.f);
                    const float* y = reinterpret_cast<const float*>(ybase);
                    if (((reinterpret_cast<uintptr_t>(y) & 15) == 0) && j0 >= 0 && j0 + 4 <= L && (j0 & 3) == 0)
                        return ldg_nc_f4(y + j0);
                    float4 v;
                    v.x = __ldg(y + reflect_index(j0 + 0, L));
                    v.y = __ldg(y + reflect_index(j0 + 1, L));
                    v.z = __ldg(y + reflect_index(j0 + 2, L));
                    v.w = __ldg(y + reflect_index(j0 + 3, L));
                    return v;
                } else {
                    if (dead) return make_float4(0.f, 0.f, 0.f, 0.f);
                    const int16_t* y = reinterpret_cast<const int16_t*>(ybase);
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int16_t* src = y + static_cast<long long>(reflect_index(j0 + e, L)) * C;
                        int acc = 0;
                        for (int ch = 0; ch < C; ++ch) acc += __ldg(src + ch);
                        v[e] = static_cast<float>(acc) * ps;
                    }
                    return make_float4(v[0], v[1], v[2], v[3]);
                }
            };
            auto load_vec = [&](const char* p) -> Raw {
                if constexpr (PC != 0) return pcm_load_raw<PC>(reinterpret_cast<const int16_t*>(p));
                else return ldg_nc_f4(p);
            };
            auto zero_raw = [&]() -> Raw {
                if constexpr (PC != 0) return pcm_zero_raw<PC>();
                else return make_float4(0.f, 0.f, 0.f, 0.f);
            };
            // fast path (c is a compile-time constant at every call)
            auto load_a = [&](int c) -> Raw {
                if (c == 0 && !live_a0) return zero_raw();
                return load_vec(pa + static_cast<long long>(2048 * c) * fbytes);
            };
            auto load_b = [&](int c) -> Raw {
                const char* pb = pa + pb_off;
                if (c == 0) return live_b0 ? load_vec(r == 0 ? pb - static_cast<long long>(128 * 128) * fbytes : pb) : zero_raw();
                return load_vec(pb - static_cast<long long>(2048 * c) * fbytes);
            };
            // general path (edge frames, unaligned clips, run-time channel counts): a compact loop over the chunks
            auto load_a_gen = [&](int c) -> Raw { return load_general(128 * (16 * c + r) + 4 * lane); };
            auto load_b_gen = [&](int c) -> Raw {
                const int m = 16 * c + r;
                return load_general(128 * ((m == 0) ? 128 : 256 - m) + 4 * lane);
            };
            auto cvt = [&](const Raw& q) -> float4 {
                if constexpr (PC != 0) return pcm_to_mono4<PC>(q, ps);
                else return q;
            };

            // ---------------------------------------------------------------- block scale (fp16 halves)
            // The scale 2^e keeps 2^e |x| < 64 over the whole frame (then no intermediate exceeds 2^15 < 65504) with e the
            // largest even exponent that does: a function of the frame's abs-max alone, e = min(eA, eB) with eA / eB the
            // exponents its first / second half would get on their own.  The first half of a frame is the second half of
            // the previous frame of the clip, which this CTA has just processed, so eA is known BEFORE the frame is
            // loaded: the fold runs with eA on the samples as they arrive and tracks the abs-max of the second half; if
            // that half turns out louder (eB < eA: an onset that crosses a scale step) the stage-1 attempt is dropped and
            // repeated with eB.  Frames without history (first of the CTA or of a clip) find both maxima in a pass of
            // their own first.  Either way the frame gets the same scale, so the result does not depend on the route.
            int e = 0, eB = 0;
            bool provisional = false;
            (void)provisional;
#if SEDB_SPLIT_FP16
            if (it > 0 && t >= 1) {
                e = e_hist;
                provisional = true;
            } else {
                float mA = 0.f, mB = 0.f;
                if (fast) {
#pragma unroll
                    for (int c0 = 0; c0 < 8; c0 += 4) {                          // two batches: rare path, few registers
                        Raw qa[4], qb[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            qa[c] = load_a(c0 + c);
                            qb[c] = load_b(c0 + c);
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            mA = absmax4(mA, cvt(qa[c]));
                            mB = absmax4(mB, cvt(qb[c]));
                        }
                    }
                } else {
#pragma unroll 1
                    for (int c = 0; c < 8; c += 2) {                         // two chunks of loads in flight
                        const Raw a0 = load_a_gen(c), b0 = load_b_gen(c), a1 = load_a_gen(c + 1), b1 = load_b_gen(c + 1);
                        mA = absmax4(absmax4(mA, cvt(a0)), cvt(a1));
                        mB = absmax4(absmax4(mB, cvt(b0)), cvt(b1));
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, o));
                    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, o));
                }
                if (lane == 0) {
                    red_s[warp] = mB;
                    red_s[32 + warp] = mA;
                }
                worker_sync();
                mA = red_s[32 + (lane & 15)];
                mB = red_s[lane & 15];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, o));
                    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, o));
                }
                eB = ebucket(mB);
                e = min(ebucket(mA), eB);
            }
#endif
#ifdef SEDB_DEBUG_SCALE
            if (tid == 0) printf("SCALE f %lld t %d fast %d prov %d e %d eB %d\n", f0 + it, t, int(fast), int(provisional), e, eB);
#endif
            float scale, inv_scale;
            for (;;) {                                                          // stage-1 attempts (one, except after a rejected scale)
                scale = __uint_as_float(static_cast<uint32_t>(127 + e) << 23);
                inv_scale = __uint_as_float(static_cast<uint32_t>(127 - e) << 23);
                const float sqrt_scale = __uint_as_float(static_cast<uint32_t>(127 + e / 2) << 23);
                // ------------------------------------------------------------ stage 1: load, window, fold, split
                Raw qa[8], qb[8];
                if (fast) {
#pragma unroll
                    for (int c = 0; c < kAhead; ++c) {
                        qa[c] = load_a(c);
                        qb[c] = load_b(c);
                    }
                }
                // the previous frame's mel partial sums read the power spectrum, which aliases the operand ring
                worker_sync();
                SEDB_PROF(0);   // frame setup (+ the abs-max pass of a frame without history)
                float2 alt01 = f2s(0.f), alt23 = f2s(0.f);
                float mB = 0.f;
                // window factors of this thread's four columns
                const float4 hc = *reinterpret_cast<const float4*>(hlane_s + 4 * lane);
                const float4 hs = *reinterpret_cast<const float4*>(hlane_s + 128 + 4 * lane);
                // (sqrt(scale) is a power of two: scaling the column factors once is bit-identical to scaling every row factor)
                const float2 sq2 = f2s(sqrt_scale);
                const float2 hc01 = f2mul(sq2, f2(hc.x, hc.y)), hc23 = f2mul(sq2, f2(hc.z, hc.w));
                const float2 hs01 = f2mul(sq2, f2(hs.x, hs.y)), hs23 = f2mul(sq2, f2(hs.z, hs.w));
                // row factors, fetched one chunk ahead of their use
#if SEDB_HROW_PIPE
                float2 ra_nx = hrow_s[r], rb_nx = hrow_s[r == 0 ? 128 : 256 - r];
#endif
                // one chunk: window, even/odd fold, hi/lo split, operand store.  Window w = sin^2(phi_m + phi_n2) from the
                // factor tables (no per-frame window traffic from L2); packed fp32 arithmetic throughout
                auto fold_chunk = [&](int c, float4 xa, float4 xb) {
#ifdef SEDB_PROF_FOLD
                    long long tq0 = clock64();
                    asm volatile("mov.b32 %0, %0;\n\tmov.b32 %1, %1;" : "+f"(xa.x), "+f"(xb.x));
                    if (prm.prof != nullptr && tid == 0) atomicAdd(prm.prof + 12, static_cast<unsigned long long>(clock64() - tq0));
#endif
                    const int s = c % kNumSlots;
                    const int u = c / kNumSlots;
                    const int m = 16 * c + r;
                    mB = absmax4(mB, xb);
#if SEDB_HROW_PIPE
                    const float2 ra = ra_nx, rb = rb_nx;
                    if (c < 7) {
                        ra_nx = hrow_s[m + 16];
                        rb_nx = hrow_s[240 - m];                    // 256 - (m + 16)
                    }
#else
                    const float2 ra = hrow_s[m], rb = hrow_s[m == 0 ? 128 : 256 - m];
#endif
                    // window weight pair scale * sin^2(phi_m + phi_n2)
                    auto win2 = [](float2 rw, float2 cn, float2 sn) {
                        const float2 t = f2fma(f2s(rw.y), sn, f2mul(f2s(rw.x), cn));
                        return f2mul(t, t);
                    };
                    // U = a wa + b wb, V = a wa - b wb with the second product fused into the add, written out so that
                    // the two instances of this code (unrolled / edge-frame loop) cannot be contracted differently: a
                    // clip's result must not depend on its alignment or its place in the batch
                    const float2 wb01 = win2(rb, hc01, hs01), wb23 = win2(rb, hc23, hs23);
                    const float2 a01 = f2mul(f2(xa.x, xa.y), win2(ra, hc01, hs01));
                    const float2 a23 = f2mul(f2(xa.z, xa.w), win2(ra, hc23, hs23));
                    const float2 xb01 = f2(xb.x, xb.y), xb23 = f2(xb.z, xb.w);
                    float2 u01, u23, v01, v23;
                    if (m == 0) {                                   // warp-uniform: U[0] = X[0], V[0] = 0, keep X[128]
                        u01 = a01; u23 = a23;
                        v01 = v23 = f2s(0.f);
                        const float2 b01 = f2mul(xb01, wb01), b23 = f2mul(xb23, wb23);
                        *reinterpret_cast<float4*>(x128_s + 4 * lane) = make_float4(b01.x, b01.y, b23.x, b23.y);
                    } else {
                        u01 = f2fma(xb01, wb01, a01); u23 = f2fma(xb23, wb23, a23);
                        v01 = f2fma(f2neg(xb01), wb01, a01); v23 = f2fma(f2neg(xb23), wb23, a23);
                    }
                    alt01 = f2add(alt01, u01);
                    alt23 = f2add(alt23, u23);
                    uint32_t uh[2], ul[2], vh[2], vl[2];
                    split_pack2(u01, uh[0], ul[0]);
                    split_pack2(u23, uh[1], ul[1]);
                    split_pack2(v01, vh[0], vl[0]);
                    split_pack2(v23, vh[1], vl[1]);
#ifdef SEDB_PROF_FOLD
                    tq0 = clock64();
#endif
                    mbar_wait(&empty1[s], (u & 1) ^ 1);
#ifdef SEDB_PROF_FOLD
                    if (prm.prof != nullptr && tid == 0) atomicAdd(prm.prof + 13, static_cast<unsigned long long>(clock64() - tq0));
                    tq0 = clock64();
#endif
                    uint8_t* dst = b1ring + s * kB1SlotBytes + b1_off;
                    *reinterpret_cast<uint2*>(dst + 0 * kB1ArrBytes) = make_uint2(uh[0], uh[1]);
                    *reinterpret_cast<uint2*>(dst + 1 * kB1ArrBytes) = make_uint2(ul[0], ul[1]);
                    *reinterpret_cast<uint2*>(dst + 2 * kB1ArrBytes) = make_uint2(vh[0], vh[1]);
                    *reinterpret_cast<uint2*>(dst + 3 * kB1ArrBytes) = make_uint2(vl[0], vl[1]);
#if !SEDB_CONSUMER_FENCE
                    fence_proxy_async_smem();
#elif SEDB_TAIL_FENCE
                    if (c == 7) fence_proxy_async_smem();             // last chunk: keep the fence off the MMA warp's tail
#endif
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full1[s]);
#ifdef SEDB_PROF_FOLD
                    if (prm.prof != nullptr && tid == 0) atomicAdd(prm.prof + 14, static_cast<unsigned long long>(clock64() - tq0));
#endif
                };
                if (fast) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (c + kAhead < 8) {
                            qa[c + kAhead] = load_a(c + kAhead);
                            qb[c + kAhead] = load_b(c + kAhead);
                        }
                        fold_chunk(c, cvt(qa[c]), cvt(qb[c]));
                    }
                } else {                                              // edge frames: a compact loop, one chunk requested ahead
                    Raw na = load_a_gen(0), nb = load_b_gen(0);
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        const Raw a = na, b = nb;
                        if (c < 7) {
                            na = load_a_gen(c + 1);
                            nb = load_b_gen(c + 1);
                        }
                        fold_chunk(c, cvt(a), cvt(b));
                    }
                }
                // alternating row sums for k1 = 128: row r of every chunk has parity r
                *reinterpret_cast<float4*>(alt_s + r * 128 + 4 * lane) =
                    make_float4(alt_sign * alt01.x, alt_sign * alt01.y, alt_sign * alt23.x, alt_sign * alt23.y);
#if SEDB_SPLIT_FP16
                if (provisional) {
                    // non-negative floats order like their bit patterns: one REDUX instead of a shuffle tree
                    const uint32_t mw = __reduce_max_sync(0xffffffffu, __float_as_uint(mB));
                    if (lane == 0) red_s[warp] = __uint_as_float(mw);
                }
#endif
                SEDB_PROF(1);   // load / fold / split / store
                worker_sync();                                        // x128_s / alt_s (/ red_s) visible to all workers
#if SEDB_SPLIT_FP16
                bool redo = false;
                if (provisional) {
                    mB = red_s[lane & 15];
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, o));
                    eB = ebucket(mB);
                    redo = eB < e;
                }
                if (tid == 0) {
                    *redo_s = redo ? 1 : 0;
                    mbar_arrive(dec);                                 // release: the MMA and copy warps read the flag
                }
                if (!redo) break;
                mbar_wait(d1_full, attempt & 1);                      // the rejected attempt's MMAs (every phase is consumed)
                ++attempt;
                e = eB;
                provisional = false;
#else
                break;
#endif
            }
            e_hist = eB;

            // ---------------------------------------------------------------- twiddle, radix-2, stage-2 A operand
            // (the row-128 input Y[n2,128] only needs the fold's partial sums: formed while the last stage-1 MMAs drain)
#if SEDB_ROW128_FOLD
            // Y[n2] and Y[128 - n2] meet the same |cos|, |sin| (the frequency index 2 k2 + 1 is odd): the 128-term sums of
            // the row-128 step become 64-term sums over A[n2] = Y[n2] - Y[128 - n2] (cos part) and B[n2] = Y[n2] + Y[128 - n2]
            // (sin part); Y[64] only has a sin term, (-1)^k2.  v_s = A[0..63] | B[0..63], red_s[28] = Y[64].
            if (tid <= 64) {
                auto ysum = [&](int n2) {
                    float acc = x128_s[n2];
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) acc += alt_s[rr * 128 + n2];
                    return acc;                                       // Y[n2,128] (scaled)
                };
                const float y0 = ysum(tid);
                if (tid == 64) {
                    red_s[28] = y0;
                } else if (tid == 0) {
                    v_s[0] = y0;
                    v_s[64] = 0.f;
                } else {
                    const float y1 = ysum(128 - tid);
                    v_s[tid] = y0 - y1;
                    v_s[64 + tid] = y0 + y1;
                }
            }
#else
            if (tid < 128) {
                float acc = x128_s[tid];
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) acc += alt_s[rr * 128 + tid];
                v_s[tid] = acc;                                       // Y[n2,128] (scaled)
            }
#endif
            // row k1 = 128: X[128 + 256 k2] = sum_n2 Y[n2,128] exp(-2 pi i n2 (2 k2+1)/256), k2 in [0,64); returns |X|^2
            float2* spec_row = nullptr;
            if (MODE == 1) spec_row = prm.spec + (static_cast<long long>(clip) * prm.n_frames + t) * kBins;
            auto row128 = [&]() -> float {
                float p128;

                const int k2 = tid >> 3;
                const int part = tid & 7;
                float ar = 0.f, ai = 0.f;
                const int mm = 2 * k2 + 1;
#if SEDB_ROW128_FOLD
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int n2 = part + 8 * i;                      // interleaved: table reads spread over banks
                    const float2 w = cs_s[(n2 * mm) & 255];
                    ar = fmaf(v_s[n2], w.x, ar);
                    ai = fmaf(v_s[64 + n2], w.y, ai);
                }
                if (part == 0) ai += (k2 & 1) ? red_s[28] : -red_s[28];   // - i Y[64] sin(pi (2 k2 + 1) / 2)
#else
#pragma unroll 8
                for (int i = 0; i < 16; ++i) {
                    const int n2 = part + 8 * i;                      // interleaved: table reads spread over banks
                    const float2 w = cs_s[(n2 * mm) & 255];
                    const float v = v_s[n2];
                    ar = fmaf(v, w.x, ar);
                    ai = fmaf(v, w.y, ai);
                }
#endif
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    ar += __shfl_xor_sync(0xffffffffu, ar, o);
                    ai += __shfl_xor_sync(0xffffffffu, ai, o);
                }
                p128 = ar * ar + ai * ai;
                if (MODE == 1 && part == 0) spec_row[128 + 256 * k2] = make_float2(ar * inv_scale, ai * inv_scale);
                return p128;
            };
#if SEDB_ROW128_EARLY
            // in the stage-1 MMA drain (the workers would only wait there); the result is parked in alt_s -- free once the
            // sums above are through -- by the thread that stores it into the spectrum later
            worker_sync();                                            // v_s (written by warps 0-2) visible, alt_s read
            {
                const float p128e = row128();
                if (MODE == 0 && (tid & 7) == 0) alt_s[tid >> 3] = p128e;
            }
#endif
            mbar_wait(d1_full, attempt & 1);
            ++attempt;
            tc_fence_after();
            SEDB_PROF(2);   // wait for stage-1 MMAs
            constexpr int kTwUnroll = SEDB_TW_UNROLL;
#pragma unroll kTwUnroll
            for (int c = 0; c < 4; ++c) {
                // chunk c = columns n in [16 c, 16 c + 16) (and n + 64); this thread owns n = na + {0, 1, 8, 9}.
                // The stage-2 A operand goes back into the very TMEM columns the thread has just read (two K elements
                // per 32-bit column): array a of chunk c lives at column 64 (a >> 1) + 16 c + 8 (a & 1), this thread's
                // K indices 4 sub .. 4 sub + 3 are the columns + 2 sub, + 2 sub + 1.  No shared memory, no proxy fence.
                const int na = 16 * c + 2 * sub;
                float c0[4], c1[4], s0[4], s1[4];
                tmem_ld2(tlane + na, c0);
                tmem_ld2(tlane + na + 8, c0 + 2);
                tmem_ld2(tlane + 64 + na, c1);
                tmem_ld2(tlane + 64 + na + 8, c1 + 2);
                tmem_ld2(tlane + 128 + na, s0);
                tmem_ld2(tlane + 128 + na + 8, s0 + 2);
                tmem_ld2(tlane + 192 + na, s1);
                tmem_ld2(tlane + 192 + na + 8, s1 + 2);
                tmem_ld_wait();
                // packed fp32: pairs of columns (na, na + 1) and (na + 8, na + 9)
                const float2 an = (c == 0) ? anc[0] : (c == 1) ? anc[1] : (c == 2) ? anc[2] : anc[3];   // registers, no local array
                const float2 ancx = f2s(an.x), ancy = f2s(an.y), w64x = f2s(w64.x), w64y = f2s(w64.y);
                float2 er[2], ei[2], orr[2], oi[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 wre = h ? wre23 : wre01, wim = h ? wim23 : wim01;
                    const float2 xlo = *reinterpret_cast<const float2*>(x128_s + na + 8 * h);
                    const float2 xhi = *reinterpret_cast<const float2*>(x128_s + 64 + na + 8 * h);
                    const float2 cr = *reinterpret_cast<const float2*>(e1tw_s + na + 8 * h);
                    const float2 ci = *reinterpret_cast<const float2*>(e1tw_s + 64 + na + 8 * h);
                    const float2 t0r = f2fma(f2neg(ancy), wim, f2mul(ancx, wre));      // tw0 = anc * W^j
                    const float2 t0i = f2fma(ancy, wre, f2mul(ancx, wim));
                    const float2 t1r = f2fma(f2neg(w64y), t0i, f2mul(w64x, t0r));      // tw1 = tw0 * W^64
                    const float2 t1i = f2fma(w64x, t0i, f2mul(w64y, t0r));
                    const float2 y0 = f2fma(f2s(sgn_k1), xlo, f2(c0[2 * h], c0[2 * h + 1]));
                    const float2 y1 = f2fma(f2s(sgn_k1), xhi, f2(c1[2 * h], c1[2 * h + 1]));
                    const float2 v0 = f2(s0[2 * h], s0[2 * h + 1]), v1 = f2(s1[2 * h], s1[2 * h + 1]);
                    const float2 z0r = f2fma(f2neg(v0), t0i, f2mul(y0, t0r)), z0i = f2fma(v0, t0r, f2mul(y0, t0i));
                    const float2 z1r = f2fma(f2neg(v1), t1i, f2mul(y1, t1r)), z1i = f2fma(v1, t1r, f2mul(y1, t1i));
                    er[h] = f2add(z0r, z1r);
                    ei[h] = f2add(z0i, z1i);
                    const float2 dr = f2sub(z0r, z1r), di = f2sub(z0i, z1i);
                    orr[h] = f2fma(f2neg(di), ci, f2mul(dr, cr));
                    oi[h] = f2fma(di, cr, f2mul(dr, ci));
                }
                const uint32_t ta = tlane + 16 * c + 2 * sub;
                uint2 h, l;
                split4(er, h, l);
                tmem_st2(ta, h.x, h.y);
                tmem_st2(ta + 8, l.x, l.y);
                split4(ei, h, l);
                tmem_st2(ta + 64, h.x, h.y);
                tmem_st2(ta + 64 + 8, l.x, l.y);
                split4(orr, h, l);
                tmem_st2(ta + 128, h.x, h.y);
                tmem_st2(ta + 128 + 8, l.x, l.y);
                split4(oi, h, l);
                tmem_st2(ta + 192, h.x, h.y);
                tmem_st2(ta + 192 + 8, l.x, l.y);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full2[c]);
            }
            SEDB_PROF(3);   // twiddle / radix-2 / split
#if !SEDB_ROW128_EARLY
            // (in the stage-2 MMA drain; the power goes to the spectrum after d2_full: it aliases the operand ring until then)
            worker_sync();                                            // v_s (written by warps 0-2) visible
            const float p128 = row128();
#endif
            SEDB_PROF(6);   // row 128
            // ---------------------------------------------------------------- power spectrum / complex output
            mbar_wait(d2_full, it & 1);
            tc_fence_after();
            SEDB_PROF(4);   // wait for stage-2 MMAs
            const bool mirrored = (sub >= 2);                         // k2 = 2 j + par >= 64
#pragma unroll 1
            for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const int j0 = 16 * sub + 8 * hb;
                    float re[8], im[8];
                    tmem_ld8(tlane + 256 + 128 * par + j0, re);
                    tmem_ld8(tlane + 256 + 128 * par + 64 + j0, im);
                    tmem_ld_wait();
                    float pw[8];
                    if (MODE == 0) {
#pragma unroll
                        for (int i = 0; i < 8; i += 2) {
                            const float2 q = f2fma(f2(re[i], re[i + 1]), f2(re[i], re[i + 1]),
                                                   f2mul(f2(im[i], im[i + 1]), f2(im[i], im[i + 1])));
                            pw[i] = q.x;
                            pw[i + 1] = q.y;
                        }
                    }
                    if (MODE == 0) {
                        // bin of element i: k = kb + 512 i, stored at k (k2 < 64) or at the Hermitian mirror 32768 - k; the
                        // column k1 = 0 has no mirror except the Nyquist bin.  Explicit shared-window addresses with
                        // immediate offsets (as `p_s[bin]` under a branch ptxas re-derives the window base for every store)
                        const int kb = k1 + 256 * par + 512 * j0;
                        if (!mirrored) {
                            const uint32_t a0 = p_s_addr + 4u * static_cast<uint32_t>(kb);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 2048u * i), "f"(pw[i]) : "memory");
                        } else {
                            const uint32_t a0 = p_s_addr + 4u * static_cast<uint32_t>(kNfft - kb);
                            const bool nyq = (par == 0 && j0 == 32);          // k = 16384 for i = 0 in the column k1 = 0
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (k1 >= 1 || (i == 0 && nyq))
                                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 - 2048u * i), "f"(pw[i]) : "memory");
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int k = k1 + 256 * (2 * (j0 + i) + par);
                            int bin;
                            float sgn = 1.f;
                            if (!mirrored) bin = k;
                            else if (k1 >= 1) { bin = kNfft - k; sgn = -1.f; }        // Hermitian mirror (conjugate)
                            else bin = (k == kNfft / 2) ? k : -1;
                            if (bin >= 0) spec_row[bin] = make_float2(re[i] * inv_scale, sgn * im[i] * inv_scale);
                        }
                    }
                }
            }
#if SEDB_ROW128_EARLY
            if (MODE == 0 && (tid & 7) == 0) p_s[128 + 256 * (tid >> 3)] = alt_s[tid >> 3];
#else
            if (MODE == 0 && (tid & 7) == 0) p_s[128 + 256 * (tid >> 3)] = p128;
#endif
            SEDB_PROF(5);   // power spectrum
            tc_fence_before();
            if (MODE == 0) {
                if (tid < 3) p_s[kBins + tid] = 0.f;                  // padding read by the vectorised mel bands
                worker_sync();                                        // power spectrum complete
                SEDB_PROF(8);
                if (it > 0) mbar_wait(part_free, (it - 1) & 1);       // previous frame's finalize has read its moments
                mel_partials(p_s, mel_tab_s, part_s, tid, kWorkerThreads);
                if (tid == 0) red_s[20] = inv_scale * inv_scale;      // the frame's block scale, for the finalize
                SEDB_PROF(9);
                __syncwarp();
                if (lane == 0) mbar_arrive(part_full);                // finalize, dB and the store run on the helper warps
                SEDB_PROF(10);
            }
            if (MODE != 0) {
                worker_sync();                                        // stores of this frame done
            }
            SEDB_PROF(7);   // mel + dB
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
    int4* mel_tab_s = reinterpret_cast<int4*>(smem + ((kBins * 4 + 16 + 15) / 16) * 16);
    float* part_s = reinterpret_cast<float*>(mel_tab_s + kMelTabEntries);
    float* coef_s = part_s + 2 * kMelMaxPieces;
    const int tid = threadIdx.x;
    for (int i = tid; i < kMelTabEntries; i += 256) mel_tab_s[i] = mel_tab[i];
    for (int i = tid; i < 4 * kMel; i += 256) coef_s[i] = mel_w[i];
    if (tid < 3) p_s[kBins + tid] = 0.f;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float2* x = spec + row * kBins;
        for (int k = tid; k < kBins; k += 256) {
            const float2 v = x[k];
            p_s[k] = v.x * v.x + v.y * v.y;
        }
        __syncthreads();
        mel_partials(p_s, mel_tab_s, part_s, tid, 256);
        __syncthreads();
        mel_finalize<256 / kMel>(part_s, mel_tab_s, coef_s, norm, 1.0f, out + row * kMel, tid);
        __syncthreads();
    }
}

}  // namespace sedb
