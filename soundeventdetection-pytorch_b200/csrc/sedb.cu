// libsedb.so -- C ABI (include/sedb.h) over the hand-written sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sedb.h"
#include "host_tables.h"
#include "logmel.cuh"
#include "resample.cuh"
#include "probe.cuh"
#include "cnn.cuh"
#include "cnn_train.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
unsigned long long* g_prof = nullptr;   // device buffer of 8 x 16 counters when phase profiling is enabled
int g_conv_layer = 0;

int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CUDA_TRY(expr)                                                                    \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) return fail("%s: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)

constexpr bool kFp16 = SEDB_SPLIT_FP16 != 0;

}  // namespace

struct sedb_ctx {
    int device = -1;
    int num_sms = 0;
    uint8_t* a1 = nullptr;        // stage-1 DFT constants
    uint8_t* b2 = nullptr;        // stage-2 DFT constants
    float* hann = nullptr;        // factored Hann window tables (make_hann_factors)
    float* mel_w = nullptr;       // per-filter line coefficients {a_r, b_r, a_f, b_f}
    int4* mel_tab = nullptr;      // per-filter band table
    // host-buffer pipeline state (sedb_logmel_host_f32 / sedb_sed_host_f32)
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    float* stage[2] = {nullptr, nullptr};
    size_t stage_elems = 0;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    float* d_out = nullptr;
    size_t d_out_elems = 0;
    float* d_norm = nullptr;
    void* d_ws[2] = {nullptr, nullptr};   // CNN workspaces of the host pipeline: [0] full groups, [1] the odd first group
    size_t d_ws_bytes[2] = {0, 0};        // (the activation planes keep their zero padding per geometry, see ws_zero_kernel)
    float* d_probs = nullptr;
    size_t d_probs_elems = 0;
    // polyphase filter tables of the sample-rate converter, one per reduced rate pair (built on first use)
    struct ResampleTable {
        int lo = 0, ln = 0, width = 0, taps = 0, span = 0;
        float* h = nullptr;          // compact coefficients [span][ln], then first[ln] as int
        uint8_t* hpack = nullptr;    // tensor-core form (pack_resample_filters); null when the rate pair does not qualify
        int npad = 0, nks = 0;
        size_t umma_smem = 0;
    };
    std::vector<ResampleTable> resample_tables;
};

#include "cnn_host.inl"
#include "cnn_train_host.inl"
static_assert(sedb::kMelPieceLen == sedb_host::kMelPieceLen && sedb::kMelMaxPieces == sedb_host::kMelMaxPieces,
              "mel piece geometry: kernels and host tables must agree");

extern "C" {

int sedb_version(void) { return SEDB_ABI_VERSION; }
const char* sedb_last_error(void) { return g_err.c_str(); }
int sedb_split_is_fp16(void) { return kFp16 ? 1 : 0; }
long long sedb_launch_count(void) { return g_launches.load(); }

int sedb_check_config(int sample_rate, int frame_size, int hop_size, int nfft, int mel_bins, float fmin, float fmax) {
    if (sample_rate != SEDB_SAMPLE_RATE || frame_size != SEDB_FRAME_SIZE || hop_size != SEDB_HOP_SIZE ||
        nfft != SEDB_NFFT || mel_bins != SEDB_MEL_BINS || fmin != SEDB_MEL_FMIN || fmax != SEDB_MEL_FMAX)
        return fail("configuration mismatch: library is specialised for sr=%d frame=%d hop=%d nfft=%d mel=%d "
                    "fmin=%g fmax=%g, caller has sr=%d frame=%d hop=%d nfft=%d mel=%d fmin=%g fmax=%g",
                    SEDB_SAMPLE_RATE, SEDB_FRAME_SIZE, SEDB_HOP_SIZE, SEDB_NFFT, SEDB_MEL_BINS, SEDB_MEL_FMIN,
                    SEDB_MEL_FMAX, sample_rate, frame_size, hop_size, nfft, mel_bins, fmin, fmax);
    return 0;
}

long long sedb_num_frames(long long n_samples) { return 1 + n_samples / SEDB_HOP_SIZE; }

int sedb_mel_filterbank(float* out_host) {
    if (!out_host) return fail("sedb_mel_filterbank: null output");
    std::vector<float> w = sedb_host::make_mel_matrix(SEDB_SAMPLE_RATE, SEDB_NFFT, SEDB_MEL_BINS, SEDB_MEL_FMIN,
                                                      SEDB_MEL_FMAX);
    std::memcpy(out_host, w.data(), w.size() * sizeof(float));
    return 0;
}

int sedb_destroy(sedb_ctx_t* c);

int sedb_create(sedb_ctx_t** out_ctx) {
    if (!out_ctx) return fail("sedb_create: null output");
    *out_ctx = nullptr;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail("sedb_create: device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", dev,
                    prop.major, prop.minor);
    sedb_ctx* c = new (std::nothrow) sedb_ctx();
    if (!c) return fail("sedb_create: out of host memory");
    struct Guard {                                   // a failure below must not leak the partial context
        sedb_ctx* c;
        ~Guard() { if (c) sedb_destroy(c); }
    } guard{c};
    c->device = dev;
    c->num_sms = prop.multiProcessorCount;
    std::vector<uint8_t> a1 = sedb_host::make_stage1_constants(kFp16);
    std::vector<uint8_t> b2 = sedb_host::make_stage2_constants(kFp16);
    std::vector<sedb_host::MelTabEntry> tab;
    std::vector<float> wts;
    int mel_slots = 0;
    if (!sedb_host::make_mel_moment_tables(SEDB_SAMPLE_RATE, SEDB_NFFT, SEDB_MEL_BINS, SEDB_MEL_FMIN, SEDB_MEL_FMAX, tab,
                                           wts, mel_slots))
        return fail("sedb_create: mel work table does not fit");
    CUDA_TRY(cudaMalloc(&c->a1, a1.size()));
    CUDA_TRY(cudaMalloc(&c->b2, b2.size()));
    std::vector<float> hann = sedb_host::make_hann_factors(SEDB_FRAME_SIZE, SEDB_NFFT);
    CUDA_TRY(cudaMalloc(&c->hann, hann.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpy(c->hann, hann.data(), hann.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->mel_w, wts.size() * sizeof(float)));
    CUDA_TRY(cudaMalloc(&c->mel_tab, tab.size() * sizeof(int4)));
    CUDA_TRY(cudaMemcpy(c->a1, a1.data(), a1.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->b2, b2.data(), b2.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->mel_w, wts.data(), wts.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->mel_tab, tab.data(), tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaFuncSetAttribute(sedb::logmel_fused_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  sedb::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(sedb::logmel_fused_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  sedb::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(sedb::logmel_fused_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  sedb::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(sedb::logmel_fused_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  sedb::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(sedb::logmel_fused_kernel<0, sedb::kInPcmAny>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, sedb::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(sedb::logmel_fused_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  sedb::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(sedb::power_mel_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  80 * 1024));
    if (int rc = sedb_cnn_kernels_init()) return rc;
    guard.c = nullptr;
    *out_ctx = c;
    return 0;
}

int sedb_destroy(sedb_ctx_t* c) {
    if (!c) return 0;
    cudaFree(c->a1);
    cudaFree(c->b2);
    cudaFree(c->hann);
    cudaFree(c->mel_w);
    cudaFree(c->mel_tab);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->stage[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    }
    cudaFree(c->d_out);
    cudaFree(c->d_norm);
    cudaFree(c->d_ws[0]);
    cudaFree(c->d_ws[1]);
    cudaFree(c->d_probs);
    for (auto& t : c->resample_tables) {
        cudaFree(t.h);
        cudaFree(t.hpack);
    }
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    delete c;
    return 0;
}

// in_fmt 0: wave = float32 mono [n_clips, wave_stride]; 1: wave = int16 PCM [n_clips, wave_stride, n_channels]
static int launch_logmel(sedb_ctx_t* c, int mode, const void* wave, long long n_clips, long long n_samples,
                         long long wave_stride, const float* norm, float* out, float* spec, cudaStream_t st,
                         int in_fmt = 0, int n_channels = 1) {
    if (!c) return fail("null context");
    if (n_clips < 0) return fail("negative clip count");
    if (n_clips == 0) return 0;
    if (!wave || (mode == 0 ? !out : !spec)) return fail("null buffer");
    if (n_samples <= sedb::kPadRefl)
        return fail("n_samples=%lld: reflect padding (center=True, n_fft=%d) needs more than %d samples", n_samples,
                    SEDB_NFFT, sedb::kPadRefl);
    if (n_samples > 2000000000LL) return fail("n_samples too large");
    if (wave_stride < n_samples) return fail("wave_stride < n_samples");
    if (in_fmt == 1 && (n_channels < 1 || n_channels > 16)) return fail("n_channels=%d: supported 1..16", n_channels);
    sedb::LogmelParams p;
    p.wave = in_fmt == 0 ? static_cast<const float*>(wave) : nullptr;
    p.pcm = in_fmt == 1 ? static_cast<const int16_t*>(wave) : nullptr;
    p.n_channels = n_channels;
    p.pcm_scale = 1.0f / (32768.0f * static_cast<float>(n_channels));
    p.wave_stride = wave_stride;
    p.n_samples = static_cast<int>(n_samples);
    p.n_clips = static_cast<int>(n_clips);
    p.n_frames = static_cast<int>(sedb_num_frames(n_samples));
    p.a1 = c->a1;
    p.b2 = c->b2;
    p.hann = c->hann;
    p.mel_w = c->mel_w;
    p.mel_tab = c->mel_tab;
    p.norm = norm;
    p.out = out;
    p.spec = reinterpret_cast<float2*>(spec);
    p.prof = g_prof;
    const long long total = n_clips * p.n_frames;
    const int grid = static_cast<int>(total < c->num_sms ? total : c->num_sms);
    const dim3 g(static_cast<unsigned>(grid)), b(sedb::kThreads);
    cudaError_t le = cudaSuccess;
    if (mode == 0 && in_fmt == 0)
        le = launch_pdl(sedb::logmel_fused_kernel<0, 0>, g, b, sedb::kSmemBytes, st, p);
    else if (mode == 0 && n_channels == 1)
        le = launch_pdl(sedb::logmel_fused_kernel<0, 1>, g, b, sedb::kSmemBytes, st, p);
    else if (mode == 0 && n_channels == 2)
        le = launch_pdl(sedb::logmel_fused_kernel<0, 2>, g, b, sedb::kSmemBytes, st, p);
    else if (mode == 0 && n_channels == 4)
        le = launch_pdl(sedb::logmel_fused_kernel<0, 4>, g, b, sedb::kSmemBytes, st, p);
    else if (mode == 0)
        le = launch_pdl(sedb::logmel_fused_kernel<0, sedb::kInPcmAny>, g, b, sedb::kSmemBytes, st, p);
    else if (in_fmt == 0)
        le = launch_pdl(sedb::logmel_fused_kernel<1, 0>, g, b, sedb::kSmemBytes, st, p);
    else
        return fail("the complex STFT output takes float32 input");
    g_launches.fetch_add(1);
    CUDA_TRY(le);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static long long gcd_ll(long long a, long long b) {
    while (b) {
        const long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}

long long sedb_resample_num_samples(long long n_in, int sr_in, int sr_out) {
    if (n_in < 0 || sr_in <= 0 || sr_out <= 0) return -1;
    const long long g = gcd_ll(sr_in, sr_out), lo = sr_in / g, ln = sr_out / g;
    return (n_in * ln + lo - 1) / lo;                       // librosa.resample: ceil(n * target_sr / orig_sr)
}

int sedb_resample_filters(int sr_in, int sr_out, float* out_host, int* width, int* taps, int* phases) {
    if (sr_in <= 0 || sr_out <= 0) return fail("sample rates must be positive");
    const long long g = gcd_ll(sr_in, sr_out);
    const int lo = static_cast<int>(sr_in / g), ln = static_cast<int>(sr_out / g);
    if (lo > 4096 || ln > 4096) return fail("rates %d -> %d reduce to %d / %d", sr_in, sr_out, lo, ln);
    int w = 0, t = 0;
    std::vector<float> h = sedb_host::make_resample_filters(lo, ln, w, t);
    if (width) *width = w;
    if (taps) *taps = t;
    if (phases) *phases = ln;
    if (out_host) std::memcpy(out_host, h.data(), h.size() * sizeof(float));
    return 0;
}

int sedb_resample_f32(sedb_ctx_t* c, const float* in_dev, long long n_clips, long long n_in, long long in_stride,
                      int sr_in, int sr_out, float* out_dev, long long out_stride, void* stream) {
    if (!c) return fail("null context");
    if (n_clips < 0 || n_in < 0) return fail("negative size");
    if (sr_in <= 0 || sr_out <= 0) return fail("sample rates must be positive");
    if (n_clips == 0 || n_in == 0) return 0;
    if (!in_dev || !out_dev) return fail("null buffer");
    if (n_clips > 65535) return fail("n_clips=%lld: at most 65535 clips per call", n_clips);
    const long long g = gcd_ll(sr_in, sr_out);
    const int lo = static_cast<int>(sr_in / g), ln = static_cast<int>(sr_out / g);
    const long long n_out = sedb_resample_num_samples(n_in, sr_in, sr_out);
    if (n_in > 2000000000LL || n_out > 2000000000LL) return fail("clip too long");
    if (in_stride < n_in || out_stride < n_out) return fail("stride shorter than the clip");
    if (lo > 4096 || ln > 4096)
        return fail("rates %d -> %d reduce to %d / %d: more than 4096 phases or samples per block", sr_in, sr_out, lo, ln);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (lo == ln) {
        CUDA_TRY(cudaMemcpy2DAsync(out_dev, out_stride * sizeof(float), in_dev, in_stride * sizeof(float),
                                   n_in * sizeof(float), n_clips, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const sedb_ctx::ResampleTable* tab = nullptr;
    for (const auto& t : c->resample_tables)
        if (t.lo == lo && t.ln == ln) tab = &t;
    if (!tab) {                                              // first use of this rate pair (allocates: not under graph capture)
        sedb_ctx::ResampleTable t;
        t.lo = lo;
        t.ln = ln;
        std::vector<float> hfull = sedb_host::make_resample_filters(lo, ln, t.width, t.taps), h;
        std::vector<int> first;
        sedb_host::compact_resample_filters(hfull, ln, t.taps, h, first, t.span);
        const size_t ncoef = h.size();
        h.resize(ncoef + first.size());
        std::memcpy(h.data() + ncoef, first.data(), first.size() * sizeof(int));
        CUDA_TRY(cudaMalloc(&t.h, h.size() * sizeof(float)));
        cudaError_t e = cudaMemcpy(t.h, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(t.h);
            return fail("cudaMemcpy(resample filters): %s", cudaGetErrorString(e));
        }
        // tensor-core path: N = phases <= 256, the tile's input span + rings must fit shared memory
        {
            int npad = 0, nks = 0;
            std::vector<uint8_t> hp = sedb_host::pack_resample_filters(hfull, ln, t.taps, sedb::kRsFilterShift, npad, nks);
            const size_t span_b = ((static_cast<size_t>(127) * lo + 16 * nks) * 4 + 127) / 128 * 128;
            const size_t smem = span_b + static_cast<size_t>(sedb::kRsASlots) * 8192 + static_cast<size_t>(sedb::kRsBSlots) * npad * 64 + 256;
            if (npad >= 64 && npad <= 256 && smem <= 220 * 1024) {       // (few phases: the CUDA-core FIR is faster)
                e = cudaMalloc(&t.hpack, hp.size());
                if (e == cudaSuccess) e = cudaMemcpy(t.hpack, hp.data(), hp.size(), cudaMemcpyHostToDevice);
                if (e != cudaSuccess) {
                    cudaFree(t.h);
                    cudaFree(t.hpack);
                    return fail("resample filters (tensor-core form): %s", cudaGetErrorString(e));
                }
                t.npad = npad;
                t.nks = nks;
                t.umma_smem = smem;
            }
        }
        c->resample_tables.push_back(t);
        tab = &c->resample_tables.back();
    }
    const long long n_blocks_all = (n_out + ln - 1) / ln;
    static const bool force_fir = std::getenv("SEDB_RESAMPLE_FIR") != nullptr;
    if (tab->hpack && !force_fir) {
        sedb::ResampleUmmaParams q;
        q.x = in_dev;
        q.y = out_dev;
        q.hpack = tab->hpack;
        q.in_stride = in_stride;
        q.out_stride = out_stride;
        q.n_in = static_cast<int>(n_in);
        q.n_out = static_cast<int>(n_out);
        q.lo = lo;
        q.ln = ln;
        q.width = tab->width;
        q.npad = tab->npad;
        q.nks = tab->nks;
        q.tiles_per_clip = static_cast<int>((n_blocks_all + 127) / 128);
        const long long tiles = static_cast<long long>(q.tiles_per_clip) * n_clips;
        if (tiles < (1LL << 31)) {
            q.n_tiles = static_cast<int>(tiles);
            CUDA_TRY(cudaFuncSetAttribute(sedb::resample_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(tab->umma_smem)));
            const int grid = static_cast<int>(tiles < c->num_sms ? tiles : c->num_sms);
            sedb::resample_umma_kernel<<<grid, sedb::kRsThreads, tab->umma_smem, st>>>(q);
            g_launches.fetch_add(1);
            CUDA_TRY(cudaGetLastError());
            return 0;
        }
    }
    sedb::ResampleParams p;
    p.x = in_dev;
    p.y = out_dev;
    p.h = tab->h;
    p.first = reinterpret_cast<const int*>(tab->h + static_cast<size_t>(tab->span) * ln);
    p.span = tab->span;
    p.in_stride = in_stride;
    p.out_stride = out_stride;
    p.n_in = static_cast<int>(n_in);
    p.n_out = static_cast<int>(n_out);
    p.lo = lo;
    p.ln = ln;
    p.width = tab->width;
    p.taps = tab->taps;
    // tile: about 1024 (phase, block-group) items, input span at most 12288 floats of shared memory
    const int per = sedb::kResampleBlocksPerThread;
    int nb = per * ((1024 + ln - 1) / ln);
    const int nb_cap = ((12288 - p.taps) / lo + 1) / per * per;
    if (nb > nb_cap) nb = nb_cap;
    if (nb < per) nb = per;
    p.nb = nb;
    const size_t smem = (static_cast<size_t>(nb - 1) * lo + p.taps) * sizeof(float);
    if (smem > 200 * 1024) return fail("rates %d -> %d: filter span does not fit shared memory", sr_in, sr_out);
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(sedb::resample_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    const long long n_blocks = (n_out + ln - 1) / ln;
    const dim3 grid(static_cast<unsigned>((n_blocks + nb - 1) / nb), static_cast<unsigned>(n_clips));
    sedb::resample_fir_kernel<<<grid, sedb::kResampleThreads, smem, st>>>(p);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sedb_logmel_f32(sedb_ctx_t* ctx, const float* wave_dev, long long n_clips, long long n_samples,
                    long long wave_stride, const float* norm_dev, float* out_dev, void* stream) {
    return launch_logmel(ctx, 0, wave_dev, n_clips, n_samples, wave_stride, norm_dev, out_dev, nullptr,
                         static_cast<cudaStream_t>(stream));
}

int sedb_stft_c64(sedb_ctx_t* ctx, const float* wave_dev, long long n_clips, long long n_samples,
                  long long wave_stride, float* spec_dev, void* stream) {
    return launch_logmel(ctx, 1, wave_dev, n_clips, n_samples, wave_stride, nullptr, nullptr, spec_dev,
                         static_cast<cudaStream_t>(stream));
}

int sedb_power_mel_db_f32(sedb_ctx_t* c, const float* spec_dev, long long rows, const float* norm_dev, float* out_dev,
                          void* stream) {
    if (!c) return fail("null context");
    if (rows < 0) return fail("negative row count");
    if (rows == 0) return 0;
    if (!spec_dev || !out_dev) return fail("null buffer");
    const int grid = static_cast<int>(rows < 4LL * c->num_sms ? rows : 4LL * c->num_sms);
    sedb::power_mel_db_kernel<<<grid, 256, 80 * 1024, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(spec_dev), rows, c->mel_w, c->mel_tab, norm_dev, out_dev);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- host-buffer pipelines -------------------------------------------------------------------------------
static int ensure_pipeline(sedb_ctx_t* c, size_t stage_elems, size_t out_elems) {
    if (!c->s_copy) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaMalloc(&c->d_norm, 2 * SEDB_MEL_BINS * sizeof(float)));
    }
    if (stage_elems > c->stage_elems) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(c->stage[i]);
            c->stage[i] = nullptr;
            CUDA_TRY(cudaMalloc(&c->stage[i], stage_elems * sizeof(float)));   // (sized in 4-byte units)
        }
        c->stage_elems = stage_elems;
    }
    if (out_elems > c->d_out_elems) {
        cudaFree(c->d_out);
        c->d_out = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_out, out_elems * sizeof(float)));
        c->d_out_elems = out_elems;
    }
    return 0;
}

// Shared body: H2D in chunks of clips (double buffered) overlapped with the fused log-mel kernel; optional CNN
// on the resident log-mel image; one D2H of the result.
static int run_host_pipeline(sedb_ctx_t* c, sedb_cnn_t* cnn, const void* wave_host, long long n_clips,
                             long long n_samples, long long wave_stride, const float* norm_host, float* result_host,
                             int in_fmt = 0, int n_channels = 1) {
    if (!c) return fail("null context");
    if (!wave_host || !result_host) return fail("null buffer");
    if (n_clips <= 0) return n_clips == 0 ? 0 : fail("negative clip count");
    if (n_samples <= sedb::kPadRefl) return fail("n_samples must exceed %d", sedb::kPadRefl);
    if (wave_stride < n_samples) return fail("wave_stride < n_samples");
    const long long T = sedb_num_frames(n_samples);
    if (in_fmt == 1 && (n_channels < 1 || n_channels > 16)) return fail("n_channels=%d: supported 1..16", n_channels);
    // bytes per sample frame: float32 mono or interleaved 16-bit PCM
    const size_t fb = in_fmt == 0 ? sizeof(float) : 2 * static_cast<size_t>(n_channels);
    // chunk ~ 64 MB of waveform: long enough to saturate PCIe, short enough to overlap with compute
    long long chunk = (64LL << 20) / (n_samples * static_cast<long long>(fb));
    if (chunk < 1) chunk = 1;
    if (chunk > n_clips) chunk = n_clips;
    const size_t samples_padded = (static_cast<size_t>(n_samples) + 7) & ~static_cast<size_t>(7);   // 16-byte rows
    if (int rc = ensure_pipeline(c, (static_cast<size_t>(chunk) * samples_padded * fb + 3) / 4,
                                 static_cast<size_t>(n_clips) * T * SEDB_MEL_BINS))
        return rc;
    const float* norm_dev = nullptr;
    if (norm_host) {
        CUDA_TRY(cudaMemcpyAsync(c->d_norm, norm_host, 2 * SEDB_MEL_BINS * sizeof(float), cudaMemcpyHostToDevice,
                                 c->s_comp));
        norm_dev = c->d_norm;
    }
    // The CNN runs per group of kCnnGroup clips behind the log-mel of their chunks, overlapped with the copies still to
    // come; groups are aligned to the end of the batch (the odd remainder goes first), so that only one full group's
    // forward pass is exposed after the last host->device copy.
    const long long kCnnGroup = 64;
    long long cnn_done = 0;                                            // clips already handed to the CNN
    long long cnn_next = (n_clips % kCnnGroup) ? (n_clips % kCnnGroup) : (n_clips < kCnnGroup ? n_clips : kCnnGroup);
    int buf = 0;
    for (long long c0 = 0; c0 < n_clips; c0 += chunk, buf ^= 1) {
        const long long nc = (n_clips - c0 < chunk) ? (n_clips - c0) : chunk;
        // the staging buffer may still be read by the kernel launched two chunks ago
        CUDA_TRY(cudaStreamWaitEvent(c->s_copy, c->ev_done[buf], 0));
        CUDA_TRY(cudaMemcpy2DAsync(c->stage[buf], samples_padded * fb,
                                   static_cast<const uint8_t*>(wave_host) + static_cast<size_t>(c0 * wave_stride) * fb,
                                   static_cast<size_t>(wave_stride) * fb, static_cast<size_t>(n_samples) * fb,
                                   static_cast<size_t>(nc), cudaMemcpyHostToDevice, c->s_copy));
        CUDA_TRY(cudaEventRecord(c->ev_h2d[buf], c->s_copy));
        CUDA_TRY(cudaStreamWaitEvent(c->s_comp, c->ev_h2d[buf], 0));
        if (int rc = launch_logmel(c, 0, c->stage[buf], nc, n_samples, static_cast<long long>(samples_padded), norm_dev,
                                   c->d_out + c0 * T * SEDB_MEL_BINS, nullptr, c->s_comp, in_fmt, n_channels))
            return rc;
        CUDA_TRY(cudaEventRecord(c->ev_done[buf], c->s_comp));
        while (cnn && cnn_done + cnn_next <= c0 + nc && cnn_done + cnn_next < n_clips) {
            if (int rc = sedb_cnn_forward_group(c, cnn, cnn_done, cnn_next, n_clips, T, cnn_next == kCnnGroup ? 0 : 1))
                return rc;
            cnn_done += cnn_next;
            cnn_next = kCnnGroup;
        }
    }
    if (!cnn) {
        CUDA_TRY(cudaMemcpyAsync(result_host, c->d_out, static_cast<size_t>(n_clips) * T * SEDB_MEL_BINS * sizeof(float),
                                 cudaMemcpyDeviceToHost, c->s_comp));
    } else {
        if (int rc = sedb_cnn_forward_group(c, cnn, cnn_done, n_clips - cnn_done, n_clips, T,
                                            n_clips - cnn_done == kCnnGroup ? 0 : 1))
            return rc;
        if (int rc = sedb_cnn_results_to_host(c, cnn, n_clips, T, result_host)) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(c->s_comp));
    return 0;
}

int sedb_logmel_host_f32(sedb_ctx_t* ctx, const float* wave_host, long long n_clips, long long n_samples,
                         long long wave_stride, const float* norm_host, float* out_host) {
    return run_host_pipeline(ctx, nullptr, wave_host, n_clips, n_samples, wave_stride, norm_host, out_host);
}

int sedb_sed_host_f32(sedb_ctx_t* ctx, sedb_cnn_t* cnn, const float* wave_host, long long n_clips,
                      long long n_samples, long long wave_stride, const float* norm_host, float* probs_host) {
    if (!cnn) return fail("null cnn handle");
    return run_host_pipeline(ctx, cnn, wave_host, n_clips, n_samples, wave_stride, norm_host, probs_host);
}

/* ---- 16-bit PCM input (the WAV data chunk as it is on disk), channel mean fused into the loader --------------- */
int sedb_logmel_pcm16(sedb_ctx_t* ctx, const int16_t* pcm_dev, long long n_clips, long long n_samples,
                      long long clip_stride, int n_channels, const float* norm_dev, float* out_dev, void* stream) {
    return launch_logmel(ctx, 0, pcm_dev, n_clips, n_samples, clip_stride, norm_dev, out_dev, nullptr,
                         static_cast<cudaStream_t>(stream), 1, n_channels);
}

int sedb_logmel_host_pcm16(sedb_ctx_t* ctx, const int16_t* pcm_host, long long n_clips, long long n_samples,
                           long long clip_stride, int n_channels, const float* norm_host, float* out_host) {
    return run_host_pipeline(ctx, nullptr, pcm_host, n_clips, n_samples, clip_stride, norm_host, out_host, 1, n_channels);
}

int sedb_sed_host_pcm16(sedb_ctx_t* ctx, sedb_cnn_t* cnn, const int16_t* pcm_host, long long n_clips,
                        long long n_samples, long long clip_stride, int n_channels, const float* norm_host,
                        float* probs_host) {
    if (!cnn) return fail("null cnn handle");
    return run_host_pipeline(ctx, cnn, pcm_host, n_clips, n_samples, clip_stride, norm_host, probs_host, 1, n_channels);
}

int sedb_adam_amsgrad_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                           float* max_exp_avg_sq_dev, long long n, float lr, float beta1, float beta2, float eps,
                           float weight_decay, long long step, float grad_scale, void* stream) {
    if (n < 0 || step < 1) return fail("sedb_adam_amsgrad_step: bad size or step");
    if (n == 0) return 0;
    if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || !max_exp_avg_sq_dev) return fail("null buffer");
    const double bc1 = 1.0 - std::pow(static_cast<double>(beta1), static_cast<double>(step));
    const double bc2 = 1.0 - std::pow(static_cast<double>(beta2), static_cast<double>(step));
    const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    const float inv_bc2_sqrt = static_cast<float>(1.0 / std::sqrt(bc2));
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    sedb::adam_amsgrad_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, max_exp_avg_sq_dev, n, step_size, beta1, beta2, inv_bc2_sqrt,
        eps, weight_decay, grad_scale);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sedb_debug_plan_layer(int cin, int cout, int pool, int mode, int ntaps, int H, int W, int amode, long long n_img,
                          int num_sms, int* out8) {
    if (!out8) return fail("null output");
    UmmaLayer L;
    L.cin = cin; L.cout = cout; L.pool = pool; L.mode = mode; L.ntaps = ntaps;
    L.cin_chunk = cin > 128 ? 128 : cin;
    L.cout_tile = cout > 128 ? 128 : cout;
    sedb::ConvParams p;
    const int S = plan_umma_layer(L, H, W, amode, n_img, num_sms, p);
    if (S < 0) return fail("layer %d->%d at %d x %d cannot be planned", cin, cout, H, W);
    out8[0] = p.n_tiles; out8[1] = p.n_nsub; out8[2] = p.fuse; out8[3] = p.n_bands; out8[4] = p.R;
    out8[5] = static_cast<int>(conv_smem_bytes(p)); out8[6] = S; out8[7] = p.kpb * 100 + p.n_wslots * 10 + (p.cstep == 32);
    return 0;
}

int sedb_debug_umma_rate(int N, int b_major, int n_acc, int reps, int lbo_a, int lbo_b, int grid,
                         unsigned long long* cycles_host) {
    if (!cycles_host || N < 16 || N > 128 || n_acc < 1 || n_acc > 4 || reps < 1) return fail("bad arguments");
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 8));
    CUDA_TRY(cudaFuncSetAttribute(sedb::umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    sedb::umma_rate_kernel<<<grid, 128, 64 * 1024>>>(d, N, b_major, n_acc, reps, lbo_a, lbo_b);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(cycles_host, d, 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

int sedb_debug_bulk_rate(int bytes, int depth, int reps, int nsrc, int spin, int grid, int nwarps,
                         unsigned long long* cycles_host) {
    if (!cycles_host || bytes < 16 || bytes % 16 || depth < 1 || depth > 16 || nwarps < 1 || nwarps > 8 ||
        bytes * depth * nwarps > 200 * 1024 || reps < 1 || nsrc < 1)
        return fail("bad arguments");
    unsigned long long* d = nullptr;
    uint8_t* src = nullptr;
    CUDA_TRY(cudaMalloc(&d, 8));
    CUDA_TRY(cudaMalloc(&src, static_cast<size_t>(bytes) * nsrc));
    CUDA_TRY(cudaMemset(src, 0, static_cast<size_t>(bytes) * nsrc));
    CUDA_TRY(cudaFuncSetAttribute(sedb::bulk_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int i = 0; i < 2; ++i)
        sedb::bulk_rate_kernel<<<grid, 32 * nwarps, bytes * depth * nwarps>>>(d, src, bytes, depth, reps, nsrc, spin);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(cycles_host, d, 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(src);
    return 0;
}

int sedb_debug_phase_profile(int enable, unsigned long long* out_host16) {
    g_conv_layer = 0;
    if (enable && !g_prof) {
        CUDA_TRY(cudaMalloc(&g_prof, 128 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(g_prof, 0, 128 * sizeof(unsigned long long)));
    }
    if (out_host16 && g_prof) {
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(out_host16, g_prof, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemset(g_prof, 0, 128 * sizeof(unsigned long long)));
    }
    if (!enable && g_prof) {
        cudaFree(g_prof);
        g_prof = nullptr;
    }
    return 0;
}

int sedb_debug_umma_probe(const float* a_dev, const float* b_dev, float* d_dev, int N, int K, int a_major,
                          int b_major, int pad, int neg_b, int swap_lbo_sbo, void* stream) {
    if (!a_dev || !b_dev || !d_dev) return fail("null buffer");
    if (N < 16 || N > 256 || N % 16 || K < 16 || K > 64 || K % 16 || pad < 0 || pad % 16)
        return fail("probe supports N in [16,256] step 16, K in {16,32,48,64}, pad multiple of 16");
    const int smem = (K / 8) * (16 + N / 8) * (128 + pad);
    CUDA_TRY(cudaFuncSetAttribute(sedb::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    sedb::umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a_dev, b_dev, d_dev, N, K, a_major,
                                                                                b_major, pad, neg_b, swap_lbo_sbo);
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
