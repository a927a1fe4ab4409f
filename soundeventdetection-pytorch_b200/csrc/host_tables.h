// Host-side generation of the constant tables the kernels consume: split DFT matrices in the tcgen05
// canonical shared-memory layout, and the Slaney mel filterbank of the reference
// (librosa.filters.mel as called at dataset/spectogram/preprocess.py:13-18).
#pragma once
#include <algorithm>
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
            // K position k of chunk k/16 holds column n = 16 (k/16) + 2 s + (jj & 1) + 8 (jj >> 1), s = (k%16)/4,
            // jj = k%4: the order in which the worker threads lay their columns into the TMEM A operand (logmel.cuh)
            const int p16 = k % 16, sq = p16 / 4, jj = p16 % 4;
            const int n = 16 * (k / 16) + 2 * sq + (jj & 1) + 8 * (jj >> 1);
            const int j = col & 63;
            const double ang = 2.0 * kPi * static_cast<double>((n * j) % 64) / 64.0;
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

// np.hanning(win) centre-padded with zeros to n_fft (librosa util.pad_center) in factored form: w[128 m + n2] = sin^2(phi_m + psi_n2) with
// phi_m = pi (128 m - lpad)/(win - 1), psi_n2 = pi n2/(win - 1) (0.5 - 0.5 cos(2x) = sin^2 x; relative accuracy is kept
// at the window's ends).  Layout: 257 x {sin phi_m, cos phi_m}, then cos psi[128], then sin psi[128].
inline std::vector<float> make_hann_factors(int win, int n_fft) {
    const int lpad = (n_fft - win) / 2;
    std::vector<float> t(2 * 257 + 256);
    for (int m = 0; m <= 256; ++m) {
        const double phi = kPi * static_cast<double>(128 * m - lpad) / (win - 1);
        t[2 * m] = static_cast<float>(std::sin(phi));
        t[2 * m + 1] = static_cast<float>(std::cos(phi));
    }
    for (int n = 0; n < 128; ++n) {
        const double psi = kPi * static_cast<double>(n) / (win - 1);
        t[2 * 257 + n] = static_cast<float>(std::cos(psi));
        t[2 * 257 + 128 + n] = static_cast<float>(std::sin(psi));
    }
    return t;
}

// ---- polyphase filters of the sample-rate converter (csrc/resample.cuh) ---------------------------------
// Kaiser-windowed sinc with resampy's `kaiser_best` design (the default of librosa.resample before librosa 0.10):
// 64 zero crossings, roll-off 0.9475937167399596, beta 14.769656459379492, in exact polyphase form for the reduced
// rates lo / ln.  Returns the table transposed, [taps][ln] float32 (taps = 2 width + lo), computed in float64.
constexpr int kResampleZeros = 64;
constexpr double kResampleRolloff = 0.9475937167399596;
constexpr double kResampleBeta = 14.769656459379492;
inline double bessel_i0(double x) {                  // power series; converges fast for x <= ~15
    double sum = 1.0, term = 1.0;
    const double q = 0.25 * x * x;
    for (int k = 1; k < 200; ++k) {
        term *= q / (static_cast<double>(k) * k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}
// Compact form for the kernel: a phase's coefficients outside a window of `span` <= 2 width + 1 taps are the clamped
// tails of the Kaiser window (|h| < 1e-21), so only [first[p], first[p] + span) of each phase is kept:
// hc[k * ln + p] = h[(first[p] + k) * ln + p].
inline void compact_resample_filters(const std::vector<float>& h, int ln, int taps, std::vector<float>& hc,
                                     std::vector<int>& first, int& span) {
    first.assign(ln, 0);
    span = 1;
    std::vector<int> last(ln, 0);
    for (int p = 0; p < ln; ++p) {
        int a = taps, b = -1;
        for (int k = 0; k < taps; ++k)
            if (std::fabs(h[static_cast<size_t>(k) * ln + p]) > 1e-12f) {
                if (k < a) a = k;
                b = k;
            }
        if (b < 0) a = b = 0;
        first[p] = a;
        last[p] = b;
        if (b - a + 1 > span) span = b - a + 1;
    }
    if (4 * span > 3 * taps) {                      // nothing worth skipping (small Lo): keep every tap, no per-phase offsets
        span = taps;
        first.assign(ln, 0);
    }
    for (int p = 0; p < ln; ++p)
        if (first[p] + span > taps) first[p] = taps - span;
    hc.assign(static_cast<size_t>(span) * ln, 0.f);
    for (int p = 0; p < ln; ++p)
        for (int k = 0; k < span; ++k) hc[static_cast<size_t>(k) * ln + p] = h[static_cast<size_t>(first[p] + k) * ln + p];
}

// Tensor-core form (csrc/resample.cuh: resample_umma_kernel): per 16-tap K-step one block of [hi | lo] fp16 halves of
// the filters scaled by 2^shift, each half npad x 16 in the canonical K-major no-swizzle operand layout
// byte(n, k) = (n / 8) 128 + (k / 8) (npad / 8) 128 + (n % 8) 16 + (k % 8) 2  (rows n >= ln and taps >= `taps` are zero).
inline std::vector<uint8_t> pack_resample_filters(const std::vector<float>& h, int ln, int taps, int shift, int& npad,
                                                  int& nks) {
    npad = (ln + 15) / 16 * 16;
    nks = (taps + 15) / 16;
    const size_t half = static_cast<size_t>(npad) * 32, chunk = 2 * half;
    std::vector<uint8_t> out(chunk * nks, 0);
    const float sc = std::ldexp(1.0f, shift);
    for (int ks = 0; ks < nks; ++ks)
        for (int n = 0; n < ln; ++n)
            for (int k = 0; k < 16; ++k) {
                const int tap = 16 * ks + k;
                if (tap >= taps) continue;
                const float v = h[static_cast<size_t>(tap) * ln + n] * sc;
                const __half hi = __float2half_rn(v);
                const __half lo = __float2half_rn(v - __half2float(hi));
                const size_t off = static_cast<size_t>(n / 8) * 128 + static_cast<size_t>(k / 8) * (npad / 8) * 128 +
                                   (n % 8) * 16 + (k % 8) * 2;
                std::memcpy(&out[ks * chunk + off], &hi, 2);
                std::memcpy(&out[ks * chunk + half + off], &lo, 2);
            }
    return out;
}

inline std::vector<float> make_resample_filters(int lo, int ln, int& width, int& taps) {
    const double base = static_cast<double>(lo < ln ? lo : ln) * kResampleRolloff;
    width = static_cast<int>(std::ceil(kResampleZeros * static_cast<double>(lo) / base));
    taps = 2 * width + lo;
    std::vector<float> h(static_cast<size_t>(taps) * ln);
    const double i0b = bessel_i0(kResampleBeta);
    for (int p = 0; p < ln; ++p)
        for (int k = 0; k < taps; ++k) {
            double t = (static_cast<double>(k - width) / lo - static_cast<double>(p) / ln) * base;
            if (t < -kResampleZeros) t = -kResampleZeros;
            if (t > kResampleZeros) t = kResampleZeros;
            const double r = t / kResampleZeros;
            const double win = bessel_i0(kResampleBeta * std::sqrt(std::fmax(0.0, 1.0 - r * r))) / i0b;
            const double sinc = (t == 0.0) ? 1.0 : std::sin(kPi * t) / (kPi * t);
            h[static_cast<size_t>(k) * ln + p] = static_cast<float>(sinc * win * (base / lo));
        }
    return h;
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

// ---- moment form of the triangular filterbank ----------------------------------------------------------------------
// Between two consecutive mel points every filter weight is a straight line in the bin index, so
//     mel_m = sum_k P_k w_m(k) = ar_m S1_m + br_m S0_m + af_m S1_{m+1} + bf_m S0_{m+1}
// with the per-segment moments S0_s = sum P_k, S1_s = sum (k - kb_s) P_k over the bins kb_s <= k < kb_{s+1} of segment
// s (segment m carries the rising edge of filter m, segment m+1 its falling edge).  Each bin is touched once instead
// of once per overlapping filter and no weight table is read.  The bins are cut into pieces of <= 35 bins, one per
// thread, so the accumulation needs no cross-lane reduction.  The lines are those of librosa.filters.mel (Slaney area
// normalisation folded in), evaluated in fp32 by the kernels: they agree with the float32 matrix to ~1e-6 of the peak.
struct MelTabEntry {
    int x, y, z, w;
};
constexpr int kMelSegments = 65;
constexpr int kMelPieceLen = 43;      // bins per piece (global grid): odd, so the grid starts cycle through all banks; <= 512 pieces
constexpr int kMelMaxPieces = 576;
constexpr int kMelTabEntries = kMelMaxPieces / 2 + 80;      // int4 entries: 2 pieces each, then 65 segment ranges

// One piece per thread position i: tab[i/2].{x,y} or .{z,w} = {first bin | length << 16, first bin of the segment |
// slot << 16} (length 0: unused position); a piece's partial moments go to its slot (pieces are numbered in segment order).  tab[kMelMaxPieces/2 + s] = {first piece, number of pieces, 0, 0} of segment s.
// coef[4 m .. +3] = {ar, br, af, bf} of filter m (br, bf already expressed in segment-local bin coordinates).
inline bool make_mel_moment_tables(int sr, int n_fft, int n_mels, double fmin, double fmax,
                                   std::vector<MelTabEntry>& tab, std::vector<float>& coef, int& n_pieces) {
    if (n_mels + 1 != kMelSegments) return false;
    const int n_bins = n_fft / 2 + 1;
    std::vector<double> mel_f(n_mels + 2);
    const double lo = hz_to_mel(fmin), hi = hz_to_mel(fmax);
    const double step = (hi - lo) / (n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(i == n_mels + 1 ? hi : lo + step * i);
    const double val = 1.0 / (static_cast<double>(n_fft) * (1.0 / sr));
    std::vector<int> kb(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) {
        int k = static_cast<int>(std::ceil(mel_f[i] / val - 1e-9));
        if (k < 0) k = 0;
        if (k > n_bins) k = n_bins;
        kb[i] = k;
    }
    coef.assign(static_cast<size_t>(n_mels) * 4, 0.f);
    for (int m = 0; m < n_mels; ++m) {
        const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1];
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        const double ar = val * enorm / fd0, br = -mel_f[m] * enorm / fd0;
        const double af = -val * enorm / fd1, bf = mel_f[m + 2] * enorm / fd1;
        coef[4 * m + 0] = static_cast<float>(ar);
        coef[4 * m + 1] = static_cast<float>(ar * kb[m] + br);
        coef[4 * m + 2] = static_cast<float>(af);
        coef[4 * m + 3] = static_cast<float>(af * kb[m + 1] + bf);
    }
    tab.assign(kMelTabEntries, MelTabEntry{0, 0, 0, 0});
    struct Piece { int k, len, seg_first_bin, slot; };
    std::vector<Piece> pieces;
    int piece = 0;
    for (int s = 0; s < kMelSegments; ++s) {
        const int first_piece = piece;
        // pieces are cut on a global grid of kMelPieceLen bins (and at the segment ends): the grid starts cycle evenly
        // through the 32 banks, only the 65 segment starts fall where they fall
        for (int k = kb[s]; k < kb[s + 1];) {
            if (piece >= kMelMaxPieces) return false;
            int end = (k / kMelPieceLen + 1) * kMelPieceLen;
            if (end > kb[s + 1]) end = kb[s + 1];
            pieces.push_back(Piece{k, end - k, kb[s], piece});
            ++piece;
            k = end;
        }
        tab[kMelMaxPieces / 2 + s] = MelTabEntry{first_piece, piece - first_piece, 0, 0};
    }
    // Thread positions: the 32 pieces a warp walks in lock step should start on 32 different shared-memory banks
    // (first bin mod 32 all distinct), which the segment boundaries spoil in segment order.  Greedy fill: every warp takes
    // one piece of each residue class while the class lasts; what is left over goes to the free lanes.
    std::vector<std::vector<Piece>> by_res(32);
    for (const Piece& q : pieces) by_res[q.k & 31].push_back(q);
    std::vector<Piece> order(kMelMaxPieces, Piece{0, 0, 0, 0});
    const int n_warps = piece <= 512 ? 16 : (piece + 31) / 32; // one pass of a 512-thread CTA, with slack lanes if possible
    std::vector<int> load(n_warps, 0);
    std::vector<std::vector<int>> res_cnt(n_warps, std::vector<int>(32, 0));
    std::vector<int> cls(32);
    for (int r = 0; r < 32; ++r) cls[r] = r;
    std::sort(cls.begin(), cls.end(), [&](int a, int b) { return by_res[a].size() > by_res[b].size(); });
    for (int r : cls)                                          // big classes first; each piece to the warp that has the
        for (const Piece& q : by_res[r]) {                     // fewest pieces of its class, then the fewest pieces
            int best = -1;
            for (int w = 0; w < n_warps; ++w) {
                if (load[w] >= 32) continue;
                if (best < 0 || res_cnt[w][r] < res_cnt[best][r] ||
                    (res_cnt[w][r] == res_cnt[best][r] && load[w] < load[best]))
                    best = w;
            }
            if (best < 0) return false;
            order[best * 32 + load[best]] = q;
            ++load[best];
            ++res_cnt[best][r];
        }
    for (int i = 0; i < kMelMaxPieces; ++i) {
        const Piece& q = order[i];
        MelTabEntry& e = tab[i / 2];
        const int a = q.k | (q.len << 16), b = q.seg_first_bin | (q.slot << 16);
        if (i & 1) { e.z = a; e.w = b; }
        else       { e.x = a; e.y = b; }
    }
    n_pieces = piece;
    return true;
}

}  // namespace sedb_host
