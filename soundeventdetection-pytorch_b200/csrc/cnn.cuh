// Conv-BN-ReLU(-Pool) blocks of the reference CNNs as implicit-GEMM tcgen05 kernels.
//
//   ConvBlock / Cnn_AvgPooling   models/spectogram_models.py:128-205   (3x3 conv, BN2d, ReLU, AvgPool2d)
//   M5                           models/waveform_models.py:9-71        (k=3 conv1d, BN1d, ReLU, MaxPool1d(4))
//
// Activation layout ("blocked planes"): per image, per group of 8 channels, a plane of S pixels x 8 channels
// (16 B per pixel).  Pixels are indexed by the *padded* linear index v = (h+1)*(W+2) + (w+1) (2-D) or v = pos+1 (1-D)
// behind `lead` zero pixels; padding pixels are zero and are never written.  With this layout
//   * one 1-D bulk copy (TMA engine) stages a K-group of a band of pixels (+halo) into shared memory, already in
//     the tcgen05 canonical K-major layout (row = pixel, 16 B = 8 channels), and
//   * the A operand of every filter tap is the same shared-memory patch with the descriptor start address
//     shifted by (dh*(W+2)+dw) pixels -- no im2col copies.
// GEMM view: M = pixels (128 per tile, up to 4 tiles per CTA pass), N = C_out tile (<=128), K = taps x C_in.
//
// Two operand regimes (template AMODE):
//   AMODE 0 (inference, BN folded): activations are ONE fp16 plane set, weights are fp16 hi + lo, two MMAs per
//            product (a wH + a wL) -- or ONE MMA against the N-concatenated [wH | wL] block when 2 C_out_tile <= 128
//            columns; fp32 accumulation in TMEM; epilogue = folded-BN scale/shift + ReLU (+ pooling) in fp32, fp16 store.
//            Measured worst case on frame probabilities: 3e-4 against the 1e-3 gate (weights are exact to 2^-22, only
//            the activations carry fp16 rounding).
//   AMODE 1 (training / gradients): activations are bf16 hi + lo plane sets (bf16 keeps the range of fp32, which the
//            gradients need), weights bf16 hi + lo, three MMAs (aH wH + aL wH + aH wL); epilogue = raw fp32 store
//            (batch-statistics BatchNorm, ReLU and pooling are separate kernels in train mode).
#pragma once
#include "umma.cuh"
#include "conv_issue.cuh"

namespace sedb {

#define CONV_PROF(i)                                                                   \
    do {                                                                               \
        if (p.prof != nullptr && (threadIdx.x & 31) == 0) {                             \
            const long long now__ = clock64();                                         \
            atomicAdd(p.prof + (i), static_cast<unsigned long long>(now__ - tprev));   \
            tprev = now__;                                                             \
        }                                                                              \
    } while (0)

constexpr int kConvCopyWarps = 4;          // a warp issues one cp.async.bulk per ~500-700 cycles whatever its size
                                           // (tests/dev/bulk_rate.py), so the copy jobs are dealt round-robin to 4 warps
constexpr int kConvThreads = 32 * (9 + kConvCopyWarps);   // 8 epilogue warps + MMA warp + copy warps
constexpr int kConvMaxTiles = 4;            // M tiles (128 pixels) per work item
constexpr int kConvMaxWSlots = 6;
constexpr int kConvMaxWSlotBytes = 73728;  // weight ring slot: `kpb` consecutive taps of one 16-channel K-step, [hi | lo] x [cout_tile][16] each
constexpr int kConvMaxKSteps = 8;          // 16-channel K-steps per input-channel chunk (cin_chunk <= 128)
constexpr int kConvLead = 8;

struct ConvParams {
    const uint8_t* in;
    uint8_t* out;            // AMODE 0: fp16 blocked planes; AMODE 1: fp32 blocked planes (32 B per pixel per 8-channel group)
    const uint8_t* wpack;
    const float* scale;      // folded BN scale  [cout]          (AMODE 0)
    const float* shift;      // folded BN shift (+ conv bias) [cout]
    int n_img, n_bands, n_tiles;
    int mode;                // 0: 2-D 3x3, 1: 1-D k=3
    int H, W, Wp;
    int cin, cout, cin_chunk, n_kchunks, cout_tile, n_ntiles;
    int n_nsub, cout_sub;    // an N tile is processed as n_nsub work items of cout_sub columns (small batches: more CTAs)
    int fuse;                // 1: one MMA against [wH | wL] (N = 2 cout_tile); needs n_nsub == 1
    int S_in, S_out;
    int R;                   // rows per band (2-D) ; positions per band = 128*n_tiles (1-D)
    int P, halo;             // patch pixels, halo pixels in front of the band
    int pool;                // 1 none, 2 avg 2x2 (2-D), 4 max 4 (1-D)
    int Ho, Wo, Wpo;
    int ntaps;
    int tapoff[9];           // tap offsets in pixels relative to the patch start
    int patch_bytes;         // narr * (cin_chunk/8) * P * 16   (narr = 1 + AMODE)
    int n_wslots;            // weight ring slots (<= kConvMaxWSlots)
    int kpb;                 // taps per weight ring slot (divides ntaps)
    int wslot_bytes;         // kpb * cout_tile * 64
    int stage_bytes;         // pooling stage (0 without pooling)
    int cstep;               // channels staged per pooling pass (16 or 32)
    int w_nrep;              // replicas of the packed weights in global memory (CTA b streams replica b % w_nrep): every
    long long w_rep_bytes;   // CTA reads the same blocks at the same time, replicas spread those reads over more L2 slices
    unsigned long long* prof; // nullable diagnostics: [0] epilogue wait, [1] epilogue work, [2] mma wait patch,
                              // [3] mma issue, [4] mma wait tmem, [5] copy wait patch_free, [6] items, [7] mma wait weights
};

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

// 8 channels of one pixel -> bf16 hi + lo halves (training planes)
__device__ __forceinline__ void store_split8(uint8_t* hi_ptr, uint8_t* lo_ptr, const float* y) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(y[2 * i]), h1 = __float2bfloat16_rn(y[2 * i + 1]);
        __nv_bfloat16 l0 = __float2bfloat16_rn(y[2 * i] - __bfloat162float(h0));
        __nv_bfloat16 l1 = __float2bfloat16_rn(y[2 * i + 1] - __bfloat162float(h1));
        h[i] = static_cast<uint32_t>(*reinterpret_cast<uint16_t*>(&h0)) |
               (static_cast<uint32_t>(*reinterpret_cast<uint16_t*>(&h1)) << 16);
        l[i] = static_cast<uint32_t>(*reinterpret_cast<uint16_t*>(&l0)) |
               (static_cast<uint32_t>(*reinterpret_cast<uint16_t*>(&l1)) << 16);
    }
    *reinterpret_cast<uint4*>(hi_ptr) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_ptr) = make_uint4(l[0], l[1], l[2], l[3]);
}
// 8 channels of one pixel -> fp16 (inference planes); saturates instead of overflowing to inf
__device__ __forceinline__ void store_h8(uint8_t* ptr, const float* y) {
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 v = __floats2half2_rn(fminf(y[2 * i], 65504.f), fminf(y[2 * i + 1], 65504.f));
        h[i] = *reinterpret_cast<const uint32_t*>(&v);
    }
    *reinterpret_cast<uint4*>(ptr) = make_uint4(h[0], h[1], h[2], h[3]);
}
__device__ __forceinline__ void load_h8(const uint8_t* ptr, float* y) {
    const uint4 v = *reinterpret_cast<const uint4*>(ptr);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        y[2 * i] = f.x;
        y[2 * i + 1] = f.y;
    }
}

#ifndef SEDB_CONV_PDL
#define SEDB_CONV_PDL 1
#endif
// shared memory: [patch | weight ring | pooling stage | scale/shift | barriers | tmem ptr].  The accumulators are double
// buffered in TMEM (columns 0..255 / 256..511 for even / odd work items) so that the patch load and the MMAs of item
// t+1 overlap the epilogue of item t.
template <int AMODE>
__global__ void __launch_bounds__(kConvThreads, 1) conv_umma_kernel(const ConvParams p) {
    constexpr int kNarr = 1 + AMODE;                       // activation plane sets (AMODE 1: hi, lo)
    constexpr uint32_t kFmt = AMODE ? kFmtBF16 : kFmtF16;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* patch = smem;
    uint8_t* wring = smem + ((p.patch_bytes + 127) / 128) * 128;
    float* stage = reinterpret_cast<float*>(wring + p.n_wslots * p.wslot_bytes);
    float* sc_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stage) + p.stage_bytes);
    float* sh_s = sc_s + p.cout;
    uint64_t* bars = reinterpret_cast<uint64_t*>(
        (reinterpret_cast<uintptr_t>(sh_s + p.cout) + 15) & ~static_cast<uintptr_t>(15));
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 32);

    // The patch is filled and released per 16-channel K-step (the MMA loop is K-step major, taps inner): as soon as
    // the MMAs of a K-step are done its two K-groups are reloaded for the next work item, so the patch transfer of
    // item t+1 overlaps the remaining MMAs of item t instead of being exposed between items.
    uint64_t* wfull = bars + 0;        // [6]
    uint64_t* wempty = bars + 6;       // [6]
    uint64_t* patch_full = bars + 12;  // [8] per K-step
    uint64_t* patch_free = bars + 20;  // [8] per K-step
    uint64_t* acc_full = bars + 28;    // [2]
    uint64_t* epi_done = bars + 30;    // [2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kConvMaxWSlots; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], 1);
        }
        for (int s = 0; s < kConvMaxKSteps; ++s) {
            mbar_init(&patch_full[s], 1);
            mbar_init(&patch_free[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&epi_done[s], 8);
        }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_ptr_s);
    if (AMODE == 0)
        for (int i = tid; i < p.cout; i += kConvThreads) {
            sc_s[i] = p.scale[i];
            sh_s[i] = p.shift[i];
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;
    // set-up done: let the next kernel's CTAs be placed as SMs free up, then wait for the previous kernel's results (and
    // for its reads of the planes this layer overwrites).  Both are no-ops for a launch without the attribute.
    if (SEDB_CONV_PDL) pdl_entry();

    const int nsplit = p.n_ntiles * p.n_nsub;
    const int items_total = p.n_img * p.n_bands * nsplit;
    const int n_items = (static_cast<int>(blockIdx.x) < items_total)
                            ? (items_total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                  static_cast<int>(gridDim.x)
                            : 0;
    const int kg_chunk = p.cin_chunk / 8;
    const int ks_chunk = p.cin_chunk / 16;
    const int wblock_bytes = p.cout_tile * 64;
    const int blocks_per_kc = p.ntaps * ks_chunk;
    const int group_bytes = p.kpb * wblock_bytes;

    auto decode = [&](int it, int& img, int& band, int& ntile, int& nsub) {
        int item = blockIdx.x + it * gridDim.x;
        nsub = item % p.n_nsub;
        item /= p.n_nsub;
        ntile = item % p.n_ntiles;
        item /= p.n_ntiles;
        band = item % p.n_bands;
        img = item / p.n_bands;
    };
    // first pixel of a band: 2-D bands start at the first REAL pixel of their first row (column 1 of the padded row), so
    // that band pixel pp sits in column pp % Wp of the image and 2x2 pooling pairs are (even, odd) lane pairs
    auto band_v0 = [&](int band) { return p.mode == 0 ? (p.R * band + 1) * p.Wp + 1 : 1 + band * 128 * p.n_tiles; };

    if (warp >= 9) {
        // ================================================================ bulk-copy producers (converged warps, one
        // elected lane issues; see the note in the MMA issuer).  The copy jobs of an item -- per K-step the patch
        // groups, then the weight groups -- are dealt round-robin to the copy warps; every job waits on its own slot
        // barrier, so the warps need no ordering among themselves.
        {
            const int cw = warp - 9;
            int gw = 0, job = 0;
            long long tprev = clock64();
            for (int it = 0; it < n_items; ++it) {
                int img, band, ntile, nsub;
                decode(it, img, band, ntile, nsub);
                const int v0 = band_v0(band);
                for (int kc = 0; kc < p.n_kchunks; ++kc) {
                    const int gk = it * p.n_kchunks + kc;
                    const uint8_t* wsrc = p.wpack + static_cast<long long>(blockIdx.x % p.w_nrep) * p.w_rep_bytes +
                                          static_cast<long long>(ntile * p.n_kchunks + kc) * blocks_per_kc * wblock_bytes;
                    for (int ks = 0; ks < ks_chunk; ++ks) {
                        if ((job++ % kConvCopyWarps) == cw) {
                            tprev = clock64();
                            if (gk > 0) mbar_wait(&patch_free[ks], (gk - 1) & 1);
                            if (cw == 0) { CONV_PROF(5); }
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&patch_full[ks], kNarr * 2 * p.P * 16);
#pragma unroll
                                for (int arr = 0; arr < kNarr; ++arr)
                                    for (int h = 0; h < 2; ++h) {
                                        const int kgl = 2 * ks + h;
                                        const long long plane =
                                            (static_cast<long long>(img) * kNarr + arr) * (p.cin / 8) + kc * kg_chunk + kgl;
                                        const uint8_t* src = p.in + (plane * p.S_in + kConvLead + v0 - p.halo) * 16;
                                        bulk_g2s(patch + (arr * kg_chunk + kgl) * p.P * 16, src, p.P * 16, &patch_full[ks]);
                                    }
                            }
                            __syncwarp();
                        }
                        for (int b = 0; b < p.ntaps; b += p.kpb, ++gw) {
                            if ((job++ % kConvCopyWarps) != cw) continue;
                            const int s = gw % p.n_wslots, u = gw / p.n_wslots;
                            mbar_wait(&wempty[s], (u & 1) ^ 1);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&wfull[s], group_bytes);
                                bulk_g2s(wring + s * p.wslot_bytes,
                                         wsrc + static_cast<long long>(ks * p.ntaps + b) * wblock_bytes, group_bytes, &wfull[s]);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ================================================================ MMA issuer: the warp stays converged and one
        // elected lane issues, so that all operands are warp-uniform (TMEM base of the 512-column allocation is 0)
        if (tmem != 0) __trap();
        {
            const uint32_t idesc = make_idesc(kFmt, kMajorK, kMajorK, 128, p.cout_sub);
            const uint32_t idesc_cat = make_idesc(kFmt, kMajorK, kMajorK, 128, 2 * p.cout_tile);
            const uint32_t a_lbo = p.P * 16;
            const uint32_t b_lbo = p.cout_tile * 32;               // [hi | lo] rows of one K-group
            const uint64_t a_base = make_smem_desc(smem_u32(patch), a_lbo, 128);
            const uint64_t b_base = make_smem_desc(smem_u32(wring), b_lbo, 128);
            // descriptors as (lo, hi) words: operands differ in the start-address field of lo only (16-byte units)
            const uint32_t a_lo0 = static_cast<uint32_t>(a_base), a_hi = static_cast<uint32_t>(a_base >> 32);
            const uint32_t b_lo0 = static_cast<uint32_t>(b_base), b_hi = static_cast<uint32_t>(b_base >> 32);
            const uint32_t a_lo_delta = kg_chunk * p.P;            // lo half of the patch, in 16-byte units (AMODE 1)
            const uint32_t b_lo_delta = p.cout_tile;               // lo rows of a weight block, in 16-byte units
            const uint32_t b_block = p.cout_tile * 4;              // one (tap, K-step) block, in 16-byte units
            const uint32_t ct = p.fuse ? 2 * p.cout_tile : p.cout_sub;   // accumulator columns per tile
            const uint32_t wslot16 = p.wslot_bytes >> 4;
            const bool fuse = p.fuse != 0;
            const int n_tiles = p.n_tiles, kpb = p.kpb, ntaps = p.ntaps;
            int ws = 0;                                         // weight ring slot and its phase
            uint32_t wph = 0;
            long long tprev = clock64();
            for (int it = 0; it < n_items; ++it) {
                int img, band, ntile, nsub;
                decode(it, img, band, ntile, nsub);
                const uint32_t acc_base = 256 * (it & 1);
                const uint32_t b_sub = b_lo0 + nsub * p.cout_sub;
                tprev = clock64();
                if (it >= 2) {                                  // this accumulator buffer was last drained by item it-2
                    mbar_wait(&epi_done[it & 1], ((it - 2) >> 1) & 1);
                    tc_fence_after();
                }
                CONV_PROF(4);
                uint32_t first = 0u;                            // 0 for the very first MMA of every tile of the item
                for (int kc = 0; kc < p.n_kchunks; ++kc) {
                    const int gk = it * p.n_kchunks + kc;
                    for (int ks = 0; ks < ks_chunk; ++ks) {
                        CONV_PROF(3);                                   // MMA issue since the last point
                        mbar_wait(&patch_full[ks], gk & 1);
                        tc_fence_after();
                        CONV_PROF(2);                                   // wait for this K-step's patch groups
                        const uint32_t a_ks = a_lo0 + 2 * ks * p.P;
                        // one asm statement per weight slot (kpb = 9, 3 or 1 taps): see conv_issue.cuh
                        for (int tap0 = 0; tap0 < ntaps; tap0 += kpb) {
                            CONV_PROF(3);
                            mbar_wait(&wfull[ws], wph);
                            tc_fence_after();
                            CONV_PROF(7);                               // wait for this slot's weights
                            if (elect_one()) {
                                const uint32_t b_cur = b_sub + ws * wslot16;
                                const uint32_t fst = first | static_cast<uint32_t>(tap0);
#define SEDB_SLOT_ARGS acc_base, ct, a_ks, a_hi, b_cur, b_hi, b_block, b_lo_delta, a_lo_delta, idesc, idesc_cat, fst, n_tiles
#define SEDB_SLOT_CALL(NT, ...)                                                              \
    do {                                                                                     \
        if (AMODE == 0) {                                                                    \
            if (fuse) conv_slot_a0_fused_##NT(SEDB_SLOT_ARGS, __VA_ARGS__);                  \
            else conv_slot_a0_split_##NT(SEDB_SLOT_ARGS, __VA_ARGS__);                       \
        } else {                                                                             \
            if (fuse) conv_slot_a1_fused_##NT(SEDB_SLOT_ARGS, __VA_ARGS__);                  \
            else conv_slot_a1_split_##NT(SEDB_SLOT_ARGS, __VA_ARGS__);                       \
        }                                                                                    \
    } while (0)
                                if (kpb == 9) {
                                    SEDB_SLOT_CALL(9, p.tapoff[0], p.tapoff[1], p.tapoff[2], p.tapoff[3], p.tapoff[4],
                                                   p.tapoff[5], p.tapoff[6], p.tapoff[7], p.tapoff[8]);
                                } else if (kpb == 3) {
                                    if (tap0 == 0) SEDB_SLOT_CALL(3, p.tapoff[0], p.tapoff[1], p.tapoff[2]);
                                    else if (tap0 == 3) SEDB_SLOT_CALL(3, p.tapoff[3], p.tapoff[4], p.tapoff[5]);
                                    else SEDB_SLOT_CALL(3, p.tapoff[6], p.tapoff[7], p.tapoff[8]);
                                } else {
                                    SEDB_SLOT_CALL(1, p.tapoff[tap0]);
                                }
#undef SEDB_SLOT_CALL
#undef SEDB_SLOT_ARGS
                                umma_commit(&wempty[ws]);
                            }
                            __syncwarp();
                            if (++ws == p.n_wslots) {
                                ws = 0;
                                wph ^= 1u;
                            }
                        }
                        first = 1u;
                        if (elect_one()) umma_commit(&patch_free[ks]);      // this K-step's patch groups may be refilled
                        __syncwarp();
                    }
                    if (kc + 1 == p.n_kchunks) {
                        if (elect_one()) umma_commit(&acc_full[it & 1]);
                        __syncwarp();
                    }
                    CONV_PROF(3);
                }
            }
        }
    } else {
        // ================================================================ epilogue warps
        const int q = warp & 3, grp = warp >> 2;
        const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const int etid = tid;                                     // 0..255
        long long tprev = clock64();
        for (int it = 0; it < n_items; ++it) {
            int img, band, ntile, nsub;
            decode(it, img, band, ntile, nsub);
            const int v0 = band_v0(band);
            const int n0 = ntile * p.cout_tile + nsub * p.cout_sub;
            // number of valid band pixels
            int rows_eff = 1, npix;
            if (p.mode == 0) {
                rows_eff = min(p.R, p.H - p.R * band);
                npix = rows_eff * p.Wp;
            } else {
                npix = min(128 * p.n_tiles, p.W - band * 128 * p.n_tiles);
            }
            mbar_wait(&acc_full[it & 1], (it >> 1) & 1);
            tc_fence_after();
            if (tid == 0) { CONV_PROF(0); }
            const uint32_t tacc = tlane + 256 * (it & 1);
            const int tile_cols = p.fuse ? 2 * p.cout_tile : p.cout_sub;
            const long long out_img = static_cast<long long>(img) * (p.cout / 8);
            // 16 accumulator columns of tile m starting at channel c0 (the [wH | wL] halves added when fused)
            auto load16 = [&](int m, int c0, float* acc) {
                tmem_ld16(tacc + m * tile_cols + c0, acc);
                if (p.fuse) {
                    float acc2[16];
                    tmem_ld16(tacc + m * tile_cols + p.cout_tile + c0, acc2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] += acc2[i];
                }
                tmem_ld_wait();
            };
            if (AMODE == 1 || p.pool == 1) {
                for (int m = grp; m < p.n_tiles; m += 2) {
                    const int pp = 128 * m + 32 * q + lane;
                    bool valid = pp < npix;
                    if (p.mode == 0) valid = valid && (pp % p.Wp) < p.W;
                    const long long vout = kConvLead + v0 + pp;
                    for (int c0 = 0; c0 < p.cout_sub; c0 += 16) {
                        float acc[16];
                        load16(m, c0, acc);
                        if (valid) {
                            if (AMODE == 0) {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    acc[i] = fmaxf(0.f, fmaf(acc[i], sc_s[n0 + c0 + i], sh_s[n0 + c0 + i]));
#pragma unroll
                                for (int g2 = 0; g2 < 2; ++g2) {
                                    const long long kg = (n0 + c0) / 8 + g2;
                                    store_h8(p.out + ((out_img + kg) * p.S_out + vout) * 16, acc + 8 * g2);
                                }
                            } else {
#pragma unroll
                                for (int g2 = 0; g2 < 2; ++g2) {
                                    const long long kg = (n0 + c0) / 8 + g2;
                                    float4* o = reinterpret_cast<float4*>(p.out + ((out_img + kg) * p.S_out + vout) * 32);
                                    o[0] = make_float4(acc[8 * g2], acc[8 * g2 + 1], acc[8 * g2 + 2], acc[8 * g2 + 3]);
                                    o[1] = make_float4(acc[8 * g2 + 4], acc[8 * g2 + 5], acc[8 * g2 + 6], acc[8 * g2 + 7]);
                                }
                            }
                        }
                    }
                }
            } else {
                if (p.mode == 0) {
                    // 2x2 average pooling.  Horizontal pairs are (even, odd) lanes: one shuffle; the pair sums of `cstep`
                    // channels go to shared memory (row-major over pair index pp / 2), and the vertical partner of pair
                    // j is pair j + Wp / 2.
                    const int cstep = p.cstep, sstr = cstep + 1;
                    for (int c0 = 0; c0 < p.cout_sub; c0 += cstep) {
                        for (int m = grp; m < p.n_tiles; m += 2) {
                            const int pp = 128 * m + 32 * q + lane;
                            float* st = stage + (pp >> 1) * sstr;
                            for (int c1 = 0; c1 < cstep; c1 += 16) {
                                float acc[16];
                                load16(m, c0 + c1, acc);
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    const float y = fmaxf(0.f, fmaf(acc[i], sc_s[n0 + c0 + c1 + i], sh_s[n0 + c0 + c1 + i]));
                                    acc[i] = y + __shfl_xor_sync(0xffffffffu, y, 1);
                                }
                                if ((lane & 1) == 0) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) st[c1 + i] = acc[i];
                                }
                            }
                        }
                        epi_sync();
                        const int ngrp = cstep / 8;                      // 8-channel groups per pass
                        const int npool = (p.R / 2) * p.Wo, half_wp = p.Wp >> 1;
                        for (int u = etid; u < npool * ngrp; u += 256) {
                            const int g8 = u % ngrp, pix = u / ngrp;
                            const int r2 = pix / p.Wo, w2 = pix % p.Wo;
                            const int ho = (p.R / 2) * band + r2;
                            if (r2 < rows_eff / 2 && ho < p.Ho) {
                                const float* a = stage + (r2 * p.Wp + w2) * sstr + g8 * 8;
                                const float* b = a + half_wp * sstr;
                                float y[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) y[i] = 0.25f * (a[i] + b[i]);
                                const long long vout = kConvLead + (ho + 1) * p.Wpo + w2 + 1;
                                const long long kg = (n0 + c0) / 8 + g8;
                                store_h8(p.out + ((out_img + kg) * p.S_out + vout) * 16, y);
                            }
                        }
                        epi_sync();
                    }
                } else {
                    // max over 4 consecutive positions (1-D): lanes 4j .. 4j + 3, two shuffles, no shared memory
                    for (int m = grp; m < p.n_tiles; m += 2) {
                        const int pp = 128 * m + 32 * q + lane;
                        const int po = band * 32 * p.n_tiles + (pp >> 2);
                        for (int c0 = 0; c0 < p.cout_sub; c0 += 16) {
                            float acc[16];
                            load16(m, c0, acc);
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float y = fmaxf(0.f, fmaf(acc[i], sc_s[n0 + c0 + i], sh_s[n0 + c0 + i]));
                                y = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, 1));
                                acc[i] = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, 2));
                            }
                            if ((lane & 3) == 0 && po < p.Wo) {
                                const long long vout = kConvLead + 1 + po;
#pragma unroll
                                for (int g2 = 0; g2 < 2; ++g2) {
                                    const long long kg = (n0 + c0) / 8 + g2;
                                    store_h8(p.out + ((out_img + kg) * p.S_out + vout) * 16, acc + 8 * g2);
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&epi_done[it & 1]);
            if (tid == 0) { CONV_PROF(1); if (p.prof) atomicAdd(p.prof + 6, 1ull); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------------
// First layer of Cnn_AvgPooling (C_in = audio_channels = 1, K = 9): CUDA cores, fp32.
// x: [n_img, H, W] fp32; w: [cout][9]; one thread per output pixel.
// (A variant with one thread per (pixel, 8-channel group) -- warp-uniform weight reads, 512-byte warp stores -- was 2.5x
// slower: the nine input loads and the index arithmetic are then repeated per group.)
// RAW 0: folded BN + ReLU, fp16 blocked planes (inference); RAW 1: raw convolution, fp32 blocked planes (training).
template <int RAW>
__global__ void __launch_bounds__(256) conv_in2d_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        uint8_t* __restrict__ out, int n_img, int H, int W, int cout,
                                                        int S_out) {
    pdl_entry();
    extern __shared__ __align__(128) uint8_t smem[];
    float* w_s = reinterpret_cast<float*>(smem);     // [cout][9]
    float* sc_s = w_s + cout * 9;
    float* sh_s = sc_s + cout;
    for (int i = threadIdx.x; i < cout * 9; i += blockDim.x) w_s[i] = w[i];
    if (!RAW)
        for (int i = threadIdx.x; i < cout; i += blockDim.x) {
            sc_s[i] = scale[i];
            sh_s[i] = shift[i];
        }
    __syncthreads();
    const long long total = static_cast<long long>(n_img) * H * W;
    const int Wp = W + 2;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int wq = static_cast<int>(idx % W);
        const int h = static_cast<int>((idx / W) % H);
        const int img = static_cast<int>(idx / (static_cast<long long>(W) * H));
        const float* xi = x + static_cast<long long>(img) * H * W;
        float v[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = h + kh - 1, ww = wq + kw - 1;
                v[kh * 3 + kw] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
            }
        const long long vout = kConvLead + (h + 1) * Wp + wq + 1;
        const long long out_img = static_cast<long long>(img) * (cout / 8);
        for (int kg = 0; kg < cout / 8; ++kg) {
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = kg * 8 + i;
                float a = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t) a = fmaf(v[t], w_s[c * 9 + t], a);
                y[i] = RAW ? a : fmaxf(0.f, fmaf(a, sc_s[c], sh_s[c]));
            }
            if (RAW) {
                float4* o = reinterpret_cast<float4*>(out + ((out_img + kg) * S_out + vout) * 32);
                o[0] = make_float4(y[0], y[1], y[2], y[3]);
                o[1] = make_float4(y[4], y[5], y[6], y[7]);
            } else {
                store_h8(out + ((out_img + kg) * S_out + vout) * 16, y);
            }
        }
    }
}

// The same layer with FOUR vertically adjacent output pixels (rows h0..h0+3 of one column) per thread.  The one-pixel
// kernel issues one broadcast LDS per FMA (288 per pixel: it is bound by the LSU at 92 us per 256 clips); here a tap's
// eight weights of a channel group come in with two LDS.128 and feed 32 FMAs (packed: 16 issue slots), the 6 x 3 input
// window is loaded once, and the lanes of a warp are neighbours along W, so every store instruction of a warp writes 512
// contiguous bytes of a plane row.  Weights in shared memory as [tap][cout].  The arithmetic per output (tap order, fmaf
// chain, folded BN) is the one-pixel kernel's, so both produce identical planes.
template <int RAW>
__global__ void __launch_bounds__(256, 3) conv_in2d_px4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, uint8_t* __restrict__ out,
                                                            int n_img, int H, int W, int cout, int S_out) {
    pdl_entry();
    extern __shared__ __align__(128) uint8_t smem[];
    float* w_s = reinterpret_cast<float*>(smem);     // [9][cout]
    float* sc_s = w_s + cout * 9;
    float* sh_s = sc_s + cout;
    for (int i = threadIdx.x; i < cout * 9; i += blockDim.x) w_s[(i % 9) * cout + i / 9] = w[i];
    if (!RAW)
        for (int i = threadIdx.x; i < cout; i += blockDim.x) {
            sc_s[i] = scale[i];
            sh_s[i] = shift[i];
        }
    __syncthreads();
    const int H4 = (H + 3) / 4;
    const long long total = static_cast<long long>(n_img) * H4 * W;
    const int Wp = W + 2;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int wq = static_cast<int>(idx % W);
        const int h0 = static_cast<int>((idx / W) % H4) * 4;
        const int img = static_cast<int>(idx / (static_cast<long long>(W) * H4));
        const float* xi = x + static_cast<long long>(img) * H * W;
        float v[6][3];                                   // input rows h0-1..h0+4, columns wq-1..wq+1 (zero outside the image)
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {
            const int hh = h0 + rr - 1;
            const bool row_ok = hh >= 0 && hh < H;
            const float* xr = xi + static_cast<long long>(hh) * W + wq;
            v[rr][0] = (row_ok && wq > 0) ? __ldg(xr - 1) : 0.f;
            v[rr][1] = row_ok ? __ldg(xr) : 0.f;
            v[rr][2] = (row_ok && wq + 1 < W) ? __ldg(xr + 1) : 0.f;
        }
        const long long vout = kConvLead + static_cast<long long>(h0 + 1) * Wp + wq + 1;
        const long long out_img = static_cast<long long>(img) * (cout / 8);
        for (int kg = 0; kg < cout / 8; ++kg) {
            float2 a[4][4];                              // packed fp32: channel pairs (two FMAs per issue slot)
#pragma unroll
            for (int px = 0; px < 4; ++px)
#pragma unroll
                for (int i = 0; i < 4; ++i) a[px][i] = f2s(0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 wa = *reinterpret_cast<const float4*>(w_s + t * cout + kg * 8);
                const float4 wb = *reinterpret_cast<const float4*>(w_s + t * cout + kg * 8 + 4);
                const float2 wt[4] = {f2(wa.x, wa.y), f2(wa.z, wa.w), f2(wb.x, wb.y), f2(wb.z, wb.w)};
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const float2 vv = f2s(v[px + t / 3][t % 3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[px][i] = f2fma(vv, wt[i], a[px][i]);
                }
            }
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                if (h0 + px >= H) break;
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = kg * 8 + i;
                    const float ai = (i & 1) ? a[px][i >> 1].y : a[px][i >> 1].x;
                    y[i] = RAW ? ai : fmaxf(0.f, fmaf(ai, sc_s[c], sh_s[c]));
                }
                const long long pix = (out_img + kg) * S_out + vout + static_cast<long long>(px) * Wp;
                if (RAW) {
                    float4* o = reinterpret_cast<float4*>(out + pix * 32);
                    o[0] = make_float4(y[0], y[1], y[2], y[3]);
                    o[1] = make_float4(y[4], y[5], y[6], y[7]);
                } else {
                    store_h8(out + pix * 16, y);
                }
            }
        }
    }
}

// Head of Cnn_AvgPooling (spectogram_models.py:193-205): mean over freq, Linear, sigmoid, x ratio time repeat.
// One warp per (image, time step).  in: blocked planes of the last block (C, Hf, Wf): SPLIT 0 = fp16 planes
// (inference), SPLIT 1 = bf16 hi + lo planes (training).  Lane l handles the 8-channel groups l, l + 32, ...
template <int SPLIT>
__global__ void __launch_bounds__(256) head2d_kernel(const uint8_t* __restrict__ in, const float* __restrict__ fc_w,
                                                     const float* __restrict__ fc_b, float* __restrict__ logits,
                                                     float* __restrict__ probs, int n_img, int C, int Hf, int Wf,
                                                     int S_in, int classes, int ratio) {
    pdl_entry();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= n_img * Hf) return;
    const int img = warp_global / Hf, h = warp_global % Hf;
    const int Wp = Wf + 2;
    const int nkg = C / 8;
    const long long img_base = static_cast<long long>(img) * (SPLIT ? 2 : 1) * nkg;
    const float inv_w = 1.0f / static_cast<float>(Wf);
    for (int cls = 0; cls < classes; ++cls) {
        float acc = 0.f;
        for (int kg = lane; kg < nkg; kg += 32) {
            float s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = 0.f;
            for (int wq = 0; wq < Wf; ++wq) {
                const long long v = kConvLead + (h + 1) * Wp + wq + 1;
                float y[8];
                if (SPLIT) {
                    const uint4 a = *reinterpret_cast<const uint4*>(in + ((img_base + kg) * S_in + v) * 16);
                    const uint4 b = *reinterpret_cast<const uint4*>(in + ((img_base + nkg + kg) * S_in + v) * 16);
                    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        y[2 * i] = __uint_as_float(wa[i] << 16) + __uint_as_float(wb[i] << 16);
                        y[2 * i + 1] = __uint_as_float(wa[i] & 0xffff0000u) + __uint_as_float(wb[i] & 0xffff0000u);
                    }
                } else {
                    load_h8(in + ((img_base + kg) * S_in + v) * 16, y);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] += y[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = fmaf(s[i] * inv_w, fc_w[cls * C + kg * 8 + i], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc += fc_b[cls];
        const float pr = 1.0f / (1.0f + __expf(-acc));
        for (int r = lane; r < ratio; r += 32) {
            const long long o = (static_cast<long long>(img) * Hf * ratio + static_cast<long long>(h) * ratio + r) * classes + cls;
            if (logits) logits[o] = acc;
            if (probs) probs[o] = pr;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// M5 conv_block1 on tensor cores: Conv1d(1->64, k=79, s=4, p=39) + BN + ReLU + MaxPool1d(4)
// (waveform_models.py:14-19).  GEMM view: M = output positions, K = 80 taps (79 + one zero), N = 64 channels.
// The im2col matrix A[p, j] = x[4p - 39 + j] is never built: the samples of a tile are staged once in shared memory
// as bf16 hi/lo, and the tcgen05 K-major descriptor walks them with OVERLAPPING strides -- 16 bytes (8 samples)
// between consecutive rows and 16 bytes between the two K-groups of an MMA.  A 16-byte row pitch is a conv stride of
// 8 samples, so even and odd output positions form two interleaved GEMMs (parity r reads a copy of the samples shifted
// by 4r), whose accumulators are max-combined in the epilogue as the first half of the MaxPool.
// One work item = 256 consecutive positions of one frame (64 pooled positions); 8 worker warps stage + drain, one
// warp issues; staging and TMEM accumulators are double buffered.
constexpr int kFrontThreads = 288;
constexpr int kFrontTilePos = 256;
constexpr int kFrontSamples = 1120;                        // per parity: 8*127 + 80 = 1096 halves, padded
constexpr int kFrontStageBytes = 4 * kFrontSamples * 2;    // {parity 0, parity 1} x {hi, lo}
constexpr int kFrontWBytes = 2 * 64 * 80 * 2;              // weights hi | lo, canonical K-major [64][80]
constexpr int kFrontSmem = 2 * kFrontStageBytes + kFrontWBytes + 2 * 64 * 4 + 16 * 8 + 16 + 128;

__global__ void __launch_bounds__(kFrontThreads, 1) m5_front_kernel(const float* __restrict__ x,
                                                                    const uint8_t* __restrict__ wpack,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift,
                                                                    uint8_t* __restrict__ out, int n_frames, int L_in,
                                                                    int L_conv, int L_out, int S_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stage = smem;                                          // [2 buffers][parity][hi|lo][kFrontSamples] bf16
    uint8_t* w_s = smem + 2 * kFrontStageBytes;
    float* sc_s = reinterpret_cast<float*>(w_s + kFrontWBytes);
    float* sh_s = sc_s + 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sh_s + 64) + 15) & ~static_cast<uintptr_t>(15));
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 8);
    uint64_t* full = bars + 0;       // [2] staging buffer filled (8 warps)
    uint64_t* acc_full = bars + 2;   // [2] accumulators complete (commit)
    uint64_t* epi_done = bars + 4;   // [2] accumulators drained (8 warps)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 8);
            mbar_init(&acc_full[i], 1);
            mbar_init(&epi_done[i], 8);
        }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_ptr_s);
    for (int i = tid; i < kFrontWBytes / 16; i += kFrontThreads)
        reinterpret_cast<uint4*>(w_s)[i] = reinterpret_cast<const uint4*>(wpack)[i];
    if (tid < 64) {
        sc_s[tid] = scale[tid];
        sh_s[tid] = shift[tid];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_s;

    const int tiles_per_frame = (L_conv + kFrontTilePos - 1) / kFrontTilePos;
    const long long total = static_cast<long long>(n_frames) * tiles_per_frame;
    const int n_items = (blockIdx.x < total) ? static_cast<int>((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 8) {
        // ================================================================ MMA issuer (converged warp, elected lane)
        if (tmem != 0) __trap();
        const uint32_t idesc = make_idesc(kFmtBF16, kMajorK, kMajorK, 128, 64);
        const uint64_t a_base = make_smem_desc(smem_u32(stage), 16, 128);          // rows 16 B apart, K-groups 16 B apart
        const uint64_t b_base = make_smem_desc(smem_u32(w_s), 64 * 16, 128);
        constexpr uint32_t kArr = (kFrontSamples * 2) >> 4;                          // one staged array, 16-byte units
        constexpr uint32_t kWLo = (64 * 80 * 2) >> 4;
        for (int it = 0; it < n_items; ++it) {
            const int buf = it & 1;
            if (it >= 2) {
                mbar_wait(&epi_done[buf], ((it - 2) >> 1) & 1);
                tc_fence_after();
            }
            mbar_wait(&full[buf], (it >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint64_t aH = a_base + buf * (4 * kArr) + r * (2 * kArr);
                    const uint64_t aL = aH + kArr;
                    const uint32_t d = buf * 128 + r * 64;
#pragma unroll
                    for (int ks = 0; ks < 5; ++ks) {
                        const uint64_t bH = b_base + ks * ((2 * 64 * 16) >> 4);
                        umma_f16(d, aH + 2 * ks, bH, idesc, ks > 0 ? 1u : 0u);      // 16 taps = 32 B further along each row
                        umma_f16(d, aL + 2 * ks, bH, idesc, 1u);
                        umma_f16(d, aH + 2 * ks, bH + kWLo, idesc, 1u);
                    }
                }
                umma_commit(&acc_full[buf]);
            }
            __syncwarp();
        }
    } else {
        // ================================================================ 8 worker warps: stage samples, drain tiles
        const int q = warp & 3, grp = warp >> 2;               // TMEM lane quarter; channel half (32 channels)
        auto stage_item = [&](int it) {
            const long long item = blockIdx.x + static_cast<long long>(it) * gridDim.x;
            const int frame = static_cast<int>(item / tiles_per_frame);
            const int tile = static_cast<int>(item - static_cast<long long>(frame) * tiles_per_frame);
            const float* xf = x + static_cast<long long>(frame) * L_in;
            const int base = 4 * tile * kFrontTilePos - 39;   // sample index of tap 0 of the tile's first position
            __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(stage + (it & 1) * kFrontStageBytes);
            for (int i = tid; i < 2 * kFrontSamples; i += 256) {
                const int r = i / kFrontSamples, k = i - r * kFrontSamples;
                const int gi = base + 4 * r + k;
                const float v = (gi >= 0 && gi < L_in) ? __ldg(xf + gi) : 0.f;
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                st[(2 * r + 0) * kFrontSamples + k] = h;
                st[(2 * r + 1) * kFrontSamples + k] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[it & 1]);
        };
        if (n_items > 0) stage_item(0);
        for (int it = 0; it < n_items; ++it) {
            if (it + 1 < n_items) stage_item(it + 1);          // buffer (it+1)&1 was released by the MMAs of item it-1
            const long long item = blockIdx.x + static_cast<long long>(it) * gridDim.x;
            const int frame = static_cast<int>(item / tiles_per_frame);
            const int tile = static_cast<int>(item - static_cast<long long>(frame) * tiles_per_frame);
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
            // lane = row qrow of the tile: positions 2 qrow (parity 0) and 2 qrow + 1 (parity 1); this warp handles
            // channels [32 grp, 32 grp + 32)
            const uint32_t t0 = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * 128 + 32 * grp;
            float y[32];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                float a0[16], a1[16];
                tmem_ld16(t0 + 16 * h2, a0);
                tmem_ld16(t0 + 64 + 16 * h2, a1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = 32 * grp + 16 * h2 + i;
                    const float v0 = fmaf(a0[i], sc_s[c], sh_s[c]), v1 = fmaf(a1[i], sc_s[c], sh_s[c]);
                    y[16 * h2 + i] = fmaxf(0.f, fmaxf(v0, v1));                 // ReLU and the first pooling pair
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&epi_done[buf]);
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], __shfl_xor_sync(0xffffffffu, y[i], 1));   // rows 2g, 2g+1
            const int qrow = q * 32 + lane;
            const int j = tile * (kFrontTilePos / 4) + (qrow >> 1);               // pooled position
            if ((lane & 1) == 0 && j < L_out) {
                const long long v = kConvLead + 1 + j;
                const long long img = static_cast<long long>(frame) * 8;
#pragma unroll
                for (int g2 = 0; g2 < 4; ++g2) {
                    const int kg = 4 * grp + g2;
                    store_h8(out + ((img + kg) * S_out + v) * 16, y + 8 * g2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// Conv1d(1->64, k=79) weight [64][79] fp32 -> {hi, lo} x canonical K-major [64][80] bf16 (tap 79 = 0)
__global__ void pack_front_weight_kernel(const float* __restrict__ w, uint8_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 64 * 80) return;
    const int c = idx / 80, k = idx % 80;
    const float v = (k < 79) ? w[c * 79 + k] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const int off = (k / 8) * (64 * 16) + c * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(out + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(out + 64 * 80 * 2 + off) = l;
}

// Head of M5 (waveform_models.py:66-67): mean over time, Linear.  One warp per frame; in: fp16 blocked planes.
__global__ void __launch_bounds__(256) head1d_kernel(const uint8_t* __restrict__ in, const float* __restrict__ fc_w,
                                                     const float* __restrict__ fc_b, float* __restrict__ logits,
                                                     int n, int C, int Lf, int S_in, int classes) {
    pdl_entry();
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wg >= n) return;
    const int nkg = C / 8;
    const long long img_base = static_cast<long long>(wg) * nkg;
    const float inv_l = 1.0f / static_cast<float>(Lf);
    for (int cls = 0; cls < classes; ++cls) {
        float acc = 0.f;
        for (int kg = lane; kg < nkg; kg += 32) {
            float s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = 0.f;
            for (int pos = 0; pos < Lf; ++pos) {
                float y[8];
                load_h8(in + ((img_base + kg) * S_in + kConvLead + 1 + pos) * 16, y);
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] += y[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = fmaf(s[i] * inv_l, fc_w[cls * C + kg * 8 + i], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) logits[static_cast<long long>(wg) * classes + cls] = acc + fc_b[cls];
    }
}

// ---- optimizer -----------------------------------------------------------------------------------------
// torch.optim.Adam(amsgrad=True) over a flat buffer (train.py:85); HBM-bound: 20 B read + 16 B written per parameter.
__global__ void __launch_bounds__(256) adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ m, float* __restrict__ v,
                                                           float* __restrict__ vmax, long long n, float step_size,
                                                           float beta1, float beta2, float inv_bc2_sqrt, float eps,
                                                           float weight_decay, float grad_scale) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float pi = p[i];
        float gi = g[i] * grad_scale;
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);          // lerp(m, g, 1 - b1)
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
        const float vm = fmaxf(vmax[i], vi);
        m[i] = mi;
        v[i] = vi;
        vmax[i] = vm;
        const float denom = sqrtf(vm) * inv_bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

// ---- parameter preparation ---------------------------------------------------------------------------
// BN fold: scale = gamma / sqrt(var + eps), shift = beta - mean*scale (+ bias*scale)
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ bias, float eps, int n, float* __restrict__ scale,
                               float* __restrict__ shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = beta[i] - mean[i] * s + (bias ? bias[i] * s : 0.f);
}

// Conv weight [cout][cin][ntaps] fp32 -> blocks [ntile][kc][ks][tap] of [hi | lo] x canonical K-major [cout_tile][16]:
// inside each 8-channel K-group the cout_tile hi rows are followed by the cout_tile lo rows, so one block serves as two
// N = cout_tile operands (or sub-ranges of them) or as ONE N = 2 cout_tile operand.  fp16 != 0: fp16 halves (inference),
// else bf16 halves (training).  transpose != 0 packs the data-gradient convolution of the same layer instead: output
// channel = ci, input channel = co, taps reversed (w'[ci][co][t] = w[co][ci][ntaps - 1 - t]); cout/cin/cout_tile/cin_chunk
// then describe that transposed convolution.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, int cout, int cin,
                                        int ntaps, int cout_tile, int cin_chunk, int fp16, int transpose, int nrep) {
    const long long total = static_cast<long long>(cout) * cin * ntaps;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int tap = static_cast<int>(idx % ntaps);
    const int ci = static_cast<int>((idx / ntaps) % cin);
    const int co = static_cast<int>(idx / (static_cast<long long>(ntaps) * cin));
    const float v = transpose ? w[(static_cast<long long>(ci) * cout + co) * ntaps + (ntaps - 1 - tap)] : w[idx];
    uint16_t h16, l16;
    if (fp16) {
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        h16 = *reinterpret_cast<const uint16_t*>(&h);
        l16 = *reinterpret_cast<const uint16_t*>(&l);
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        h16 = *reinterpret_cast<const uint16_t*>(&h);
        l16 = *reinterpret_cast<const uint16_t*>(&l);
    }
    const int ntile = co / cout_tile, n = co % cout_tile;
    const int kc = ci / cin_chunk, cil = ci % cin_chunk;
    const int ks = cil / 16, k = cil % 16;
    const int n_kchunks = cin / cin_chunk, ks_chunk = cin_chunk / 16;
    const long long block = ((static_cast<long long>(ntile) * n_kchunks + kc) * ks_chunk + ks) * ntaps + tap;   // K-step major
    const long long base = block * (cout_tile * 64);
    const int off = (k / 8) * (cout_tile * 32) + n * 16 + (k % 8) * 2;
    const long long rep_bytes = total * 4;                      // hi + lo, 2 bytes each
    for (int r = 0; r < nrep; ++r) {
        *reinterpret_cast<uint16_t*>(out + r * rep_bytes + base + off) = h16;
        *reinterpret_cast<uint16_t*>(out + r * rep_bytes + base + cout_tile * 16 + off) = l16;
    }
}

// Workspace hygiene: padding pixels of every activation plane must be zero (the kernels never write them).  The host
// zeroes a workspace whenever its (pointer, geometry) differs from the one it zeroed last; nothing in-band is trusted.
__global__ void ws_zero_kernel(uint4* __restrict__ ws, long long n16) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        ws[i] = z;
}

}  // namespace sedb
