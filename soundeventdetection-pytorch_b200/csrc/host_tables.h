// Host-side generation of the constant tables the kernels consume: split DFT matrices in the tcgen05
// canonical shared-memory layout, and the Slaney mel filterbank of the reference
// (librosa.filters.mel as called at dataset/spectogram/preprocess.py:13-18).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace sedb_host {

constexpr double kPi = 3.14159265358979323846;

inline uint16_t to_split_bits(float x, bool fp16) {
    if (fp16) {
        __half h = __float2half_rn(x);
        uint16_t u;
        std::memcpy(&u, &h, 2);
        return u;
    }
    __nv_bfloat16 b = __float2bfloat16_rn(x);
    uint16_t u;
    std::memcpy(&u, &b, 2);
    return u;
}
inline float from_split_bits(uint16_t u, bool fp16) {
    if (fp16) {
        __half h;
        std::memcpy(&h, &u, 2);
        return __half2float(h);
    }
    __nv_bfloat16 b;
    std::memcpy(&b, &u, 2);
    return __bfloat162float(b);
}
inline void split_hi_lo(double v, bool fp16, uint16_t& hi, uint16_t& lo) {
    const float x = static_cast<float>(v);
    hi = to_split_bits(x, fp16);
    lo = to_split_bits(x - from_split_bits(hi, fp16), fp16);
}

// Stage-1 constants (even/odd folded real DFT over n1, K = 128): 8 K-chunks x {cH, cL, sH, sL} x
// [128 rows k1][16 k = m - 16 c], K-major canonical (byte = (k/8)*2048 + row*16 + (k%8)*2);
// c = cos(2 pi k1 m/256) multiplies U[m] = X[m] + X[256-m], s = -sin(2 pi k1 m/256) multiplies V[m] = X[m] - X[256-m].
inline std::vector<uint8_t> make_stage1_constants(bool fp16) {
    std::vector<uint8_t> buf(8 * 4 * 4096);
    for (int c = 0; c < 8; ++c)
        for (int row = 0; row < 128; ++row)
            for (int kk = 0; kk < 16; ++kk) {
                const int m = 16 * c + kk;
                const double ang = 2.0 * kPi * static_cast<double>((row * m) % 256) / 256.0;
                uint16_t ch, cl, sh, sl;
                split_hi_lo(std::cos(ang), fp16, ch, cl);
                split_hi_lo(-std::sin(ang), fp16, sh, sl);
                const size_t off = static_cast<size_t>(c) * 16384 + (kk / 8) * 2048 + row * 16 + (kk % 8) * 2;
                std::memcpy(&buf[off + 0 * 4096], &ch, 2);
                std::memcpy(&buf[off + 1 * 4096], &cl, 2);
                std::memcpy(&buf[off + 2 * 4096], &sh, 2);
                std::memcpy(&buf[off + 3 * 4096], &sl, 2);
            }
    return buf;
}

// Stage-2 constants (64-point complex DFT shared by the even and odd radix-2 halves): {breH, breL, bimH, bimL} x
// [128 rows = output column][64 k = n], K-major canonical.  Output columns 0..63 are real parts, 64..127 imaginary:
//   bre (multiplies the real part of the A operand) = [cos | -sin],  bim (imaginary part) = [sin | cos].
inline std::vector<uint8_t> make_stage2_constants(bool fp16) {
    std::vector<uint8_t> buf(4 * 16384);
    for (int col = 0; col < 128; ++col)
        for (int k = 0; k < 64; ++k) {
            const int j = col & 63;
            const double ang = 2.0 * kPi * static_cast<double>((k * j) % 64) / 64.0;
            const double cv = std::cos(ang), sv = std::sin(ang);
            const double bre = (col < 64) ? cv : -sv;
            const double bim = (col < 64) ? sv : cv;
            uint16_t rh, rl, ih, il;
            split_hi_lo(bre, fp16, rh, rl);
            split_hi_lo(bim, fp16, ih, il);
            const size_t off = static_cast<size_t>(k / 8) * 2048 + col * 16 + (k % 8) * 2;
            std::memcpy(&buf[off + 0 * 16384], &rh, 2);
            std::memcpy(&buf[off + 1 * 16384], &rl, 2);
            std::memcpy(&buf[off + 2 * 16384], &ih, 2);
            std::memcpy(&buf[off + 3 * 16384], &il, 2);
        }
    return buf;
}

// np.hanning(win) centre-padded with zeros to n_fft (librosa util.pad_center), float32.
inline std::vector<float> make_hann_padded(int win, int n_fft) {
    std::vector<float> w(n_fft, 0.f);
    const int lpad = (n_fft - win) / 2;
    for (int i = 0; i < win; ++i)
        w[lpad + i] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * kPi * static_cast<double>(i) / (win - 1)));
    return w;
}

// ---- librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney') --------------------
inline double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return (f >= min_log_hz) ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
inline double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return (m >= min_log_mel) ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// Returns the dense (n_bins x n_mels) float32 row-major matrix == MEL_FILTER_BANK_MATRIX.
inline std::vector<float> make_mel_matrix(int sr, int n_fft, int n_mels, double fmin, double fmax) {
    const int n_bins = n_fft / 2 + 1;
    std::vector<double> mel_f(n_mels + 2);
    const double lo = hz_to_mel(fmin), hi = hz_to_mel(fmax);
    const double step = (hi - lo) / (n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(i == n_mels + 1 ? hi : lo + step * i);
    const double val = 1.0 / (static_cast<double>(n_fft) * (1.0 / sr));      // np.fft.rfftfreq
    std::vector<float> w(static_cast<size_t>(n_bins) * n_mels, 0.f);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < n_bins; ++k) {
            const double fk = k * val;
            const double lower = -(mel_f[i] - fk) / fd0;
            const double upper = (mel_f[i + 2] - fk) / fd1;
            const double tri = std::fmax(0.0, std::fmin(lower, upper));
            const float tri32 = static_cast<float>(tri);
            w[static_cast<size_t>(k) * n_mels + i] = static_cast<float>(static_cast<double>(tri32) * enorm);
        }
    }
    return w;
}

// Per-filter line coefficients: weight(k) = max(0, min(a_r k + b_r, a_f k + b_f)), the triangle of
// librosa.filters.mel with the Slaney area normalisation folded in (evaluated in fp32 by the kernels; agrees with the
// float32 matrix to ~1e-6 of the peak weight).
inline std::vector<float> make_mel_coefficients(int sr, int n_fft, int n_mels, double fmin, double fmax) {
    std::vector<double> mel_f(n_mels + 2);
    const double lo = hz_to_mel(fmin), hi = hz_to_mel(fmax);
    const double step = (hi - lo) / (n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(i == n_mels + 1 ? hi : lo + step * i);
    const double val = 1.0 / (static_cast<double>(n_fft) * (1.0 / sr));
    std::vector<float> c(static_cast<size_t>(n_mels) * 4);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        c[4 * i + 0] = static_cast<float>(val * enorm / fd0);
        c[4 * i + 1] = static_cast<float>(-mel_f[i] * enorm / fd0);
        c[4 * i + 2] = static_cast<float>(-val * enorm / fd1);
        c[4 * i + 3] = static_cast<float>(mel_f[i + 2] * enorm / fd1);
    }
    return c;
}

struct MelTabEntry {
    int x, y, z, w;
};
constexpr int kMelRows = 16;          // work rows (one per worker warp of the fused kernel)
constexpr int kMelSegsPerRow = 16;
constexpr int kMelTabEntries = kMelRows * kMelSegsPerRow + 64;

// Compact, load-balanced form of the filterbank for the kernels:
//   weights : per filter the contiguous run of non-zero weights, extended down to a multiple of 4 bins and padded
//             to a multiple of 4 entries (float4 loads);
//   tab[row*16 + s] = {first bin, filter index, float4 count, partial-sum slot}  -- segments of work row `row`
//             (a filter may be split over consecutive rows; count 0 terminates a row);
//   tab[256 + m]    = {first slot, number of slots, 0, 0} of filter m (partials are summed in slot order).
inline bool make_mel_segments(const std::vector<float>& dense, int n_bins, int n_mels, std::vector<MelTabEntry>& tab,
                              std::vector<float>& weights, int& n_slots) {
    tab.assign(kMelTabEntries, MelTabEntry{0, 0, 0, 0});
    weights.clear();
    std::vector<int> first(n_mels), n4(n_mels), woff(n_mels);
    long long total_cost = 0;
    auto cost = [](int items) { return (items + 31) / 32 + 2; };
    for (int m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < n_bins; ++k)
            if (dense[static_cast<size_t>(k) * n_mels + m] != 0.f) {
                if (lo < 0) lo = k;
                hi = k;
            }
        if (lo < 0) { lo = 0; hi = -1; }
        lo &= ~3;
        int cnt = hi - lo + 1;
        if (cnt < 0) cnt = 0;
        const int cnt_pad = (cnt + 3) & ~3;
        first[m] = lo;
        n4[m] = cnt_pad / 4;
        woff[m] = static_cast<int>(weights.size());
        for (int i = 0; i < cnt_pad; ++i) {
            const int k = lo + i;
            weights.push_back(k < n_bins ? dense[static_cast<size_t>(k) * n_mels + m] : 0.f);
        }
        total_cost += cost(n4[m]);
    }
    const long long quota = (total_cost + kMelRows - 1) / kMelRows + 1;
    int row = 0, seg = 0, slot = 0;
    long long used = 0;
    for (int m = 0; m < n_mels; ++m) {
        int done = 0;
        tab[kMelRows * kMelSegsPerRow + m] = MelTabEntry{slot, 0, 0, 0};
        if (n4[m] == 0) continue;
        while (done < n4[m]) {
            if ((used >= quota || seg >= kMelSegsPerRow) && row + 1 < kMelRows) {
                ++row;
                seg = 0;
                used = 0;
            }
            if (seg >= kMelSegsPerRow) return false;
            long long room = (quota - used - 2) * 32;                 // float4 items that still fit this row
            if (room < 32) room = 32;
            int take = n4[m] - done;
            if (take > room && row + 1 < kMelRows) take = static_cast<int>(room);
            tab[row * kMelSegsPerRow + seg] = MelTabEntry{first[m] + 4 * done, m, take, slot};
            tab[kMelRows * kMelSegsPerRow + m].y += 1;
            used += cost(take);
            done += take;
            ++seg;
            ++slot;
        }
    }
    n_slots = slot;
    return slot <= 128;
}

}  // namespace sedb_host
