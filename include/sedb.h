/* sedb.h -- C ABI of libsedb.so: the B200-native hot path of ariel415el/SoundEventDetection-Pytorch.
 *
 * The reference is pure Python and has no FFI layer; the boundary it exposes for this path is the
 * set of Python callables cited below (paths relative to the reference repository).  Each entry
 * point here is what a binding for that callable would call.  Conventions:
 *   - every function returns 0 on success and a non-zero code on failure; the message is available
 *     from sedb_last_error() (thread-local).  Nothing throws or aborts across the boundary.
 *   - "dev" pointers are CUDA device pointers on the device that was current when the context was
 *     created; "host" pointers are ordinary (ideally pinned) host memory.
 *   - the caller owns every input/output buffer; the context owns only constant tables and packed
 *     weight copies.  Device entry points are asynchronous and stream-ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream); they never synchronise.
 *   - a context may be used from one host thread at a time; one context per GPU.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SEDB_H_
#define SEDB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEDB_ABI_VERSION 1

/* configuration the kernels are specialised for (dataset/common_config.py:2-8,
 * dataset/spectogram/spectogram_configs.py:5-8) */
#define SEDB_SAMPLE_RATE 48000
#define SEDB_FRAME_SIZE 31680
#define SEDB_HOP_SIZE 15840
#define SEDB_NFFT 32768
#define SEDB_NUM_BINS 16385
#define SEDB_MEL_BINS 64
#define SEDB_MEL_FMIN 20.0f
#define SEDB_MEL_FMAX 24000.0f

typedef struct sedb_ctx sedb_ctx_t;
typedef struct sedb_cnn sedb_cnn_t;
typedef struct sedb_m5 sedb_m5_t;

/* ---- library / configuration ------------------------------------------------------------------ */
int sedb_version(void);
const char* sedb_last_error(void);
/* 1 if the library was built with fp16 split operands (+ block scaling), 0 for bf16 split operands */
int sedb_split_is_fp16(void);
/* Refuses (non-zero) when the caller's dataset/spectogram/spectogram_configs.py constants differ from
 * the compile-time specialisation above. */
int sedb_check_config(int sample_rate, int frame_size, int hop_size, int nfft, int mel_bins, float fmin,
                      float fmax);
/* T = 1 + n_samples / hop  (librosa.stft(center=True) as called at dataset/spectogram/preprocess.py:25-33) */
long long sedb_num_frames(long long n_samples);

int sedb_create(sedb_ctx_t** out_ctx);
int sedb_destroy(sedb_ctx_t* ctx);

/* ---- log-mel feature extraction ---------------------------------------------------------------- */
/* MEL_FILTER_BANK_MATRIX (dataset/spectogram/preprocess.py:13-18): writes the (16385, 64) float32
 * row-major matrix the kernels use to host memory.  Host-only, no device needed. */
int sedb_mel_filterbank(float* out_host);

/* Fused multichannel_complex_to_log_mel(multichannel_stft(x)) (preprocess.py:21-45) for a batch of mono
 * clips.  wave_dev: [n_clips, wave_stride] float32 (n_samples valid per clip, n_samples > 16384);
 * norm_dev: NULL or mean[64] followed by std[64] (SpectogramDataset.transform, spectograms_dataset.py:104-105);
 * out_dev: [n_clips, T, 64] float32.
 * Accuracy contract ("dynamic-range window").  The reference's DFT runs in float64; this one runs on tensor cores with
 * split 16-bit operands and fp32 accumulation, which carries an error relative to the frame's TOTAL energy: ~2^-22 with
 * fp16 halves plus a per-frame power-of-two block scale (the default build, sedb_split_is_fp16() == 1), ~2^-17 with bf16
 * halves.  Every mel bin within 100 dB (fp16 build; 75 dB for the bf16 build) of the loudest mel bin of the SAME frame is
 * within 1e-2 dB of the reference; a bin further down is reported no lower than the reference minus 1e-2 dB and at most at
 * the DFT's own noise floor (about 100 dB below the frame's loudest bin).  Real recordings stay inside the window (16-bit
 * PCM quantisation noise under a full-scale tone sits 110-125 dB down: measured error there ~0.1 dB, see
 * tests/test_gpu_logmel.py::test_dynamic_range_contract_*).
 * Reproducibility.  A frame's result is a function of the frame's samples alone: bit-identical whichever batch the clip
 * arrives in, wherever it sits in it and however its buffer is aligned (the block scale is the exponent bucket of the
 * frame's abs-max; the kernel usually knows it from the half the frame shares with the previous one and repeats the first
 * DFT stage of the few frames where the other half is louder by a scale step). */
int sedb_logmel_f32(sedb_ctx_t* ctx, const float* wave_dev, long long n_clips, long long n_samples,
                    long long wave_stride, const float* norm_dev, float* out_dev, void* stream);

/* multichannel_stft (preprocess.py:21-36): spec_dev: [n_clips, T, 16385] complex64 (interleaved re,im). */
int sedb_stft_c64(sedb_ctx_t* ctx, const float* wave_dev, long long n_clips, long long n_samples,
                  long long wave_stride, float* spec_dev, void* stream);

/* multichannel_complex_to_log_mel (preprocess.py:39-45) on an existing complex64 spectrogram:
 * spec_dev: [rows, 16385] complex64; out_dev: [rows, 64] float32. */
int sedb_power_mel_db_f32(sedb_ctx_t* ctx, const float* spec_dev, long long rows, const float* norm_dev,
                          float* out_dev, void* stream);

/* Host-buffer variant of sedb_logmel_f32: copies clips host->device in chunks overlapped with compute on
 * internal streams and copies the log-mel image back; synchronises before returning.  wave_host and
 * out_host should be pinned for full PCIe bandwidth. */
int sedb_logmel_host_f32(sedb_ctx_t* ctx, const float* wave_host, long long n_clips, long long n_samples,
                         long long wave_stride, const float* norm_host, float* out_host);

/* ---- sample-rate conversion in front of the path ---------------------------------------------------
 * Replaces the resampling branch of read_multichannel_audio (dataset/dataset_utils.py:77-84:
 * `librosa.resample(x, orig_sr=fs, target_sr=target_fs)` per channel) for files that are not at the
 * working rate.  Band-limited interpolation with a Kaiser-windowed sinc in exact polyphase form, resampy's
 * `kaiser_best` design (64 zero crossings, roll-off 0.9475937, beta 14.7697: librosa's default before
 * 0.10; from 0.10 on the default is `soxr_hq`, which is not reproduced).  The reference pins no librosa
 * version, so parity for this row is unpinned; the oracle (oracle/resample_ref.py) is checked against
 * torchaudio's sinc_interp_kaiser with the same parameters.
 * in: [n_clips, in_stride] float32 device, n_in valid samples per clip; out: [n_clips, out_stride] with
 * sedb_resample_num_samples(n_in, sr_in, sr_out) = ceil(n_in * sr_out / sr_in) samples per clip.  The
 * first call for a rate pair builds and uploads its filter table (allocates and synchronises: do it
 * outside graph capture).  Rates that reduce to more than 4096 : 4096 are refused.  Rate pairs with 64-256 phases
 * (44.1 -> 48 kHz: 160) run as a tcgen05 GEMM against a Toeplitz view of the input, the others as a CUDA-core
 * polyphase FIR; the environment variable SEDB_RESAMPLE_FIR forces the latter. */
long long sedb_resample_num_samples(long long n_in, int sr_in, int sr_out);
/* The filter table the converter uses for a rate pair, [taps][phases] float32 (host; no GPU needed): tap k of
 * phase p weighs x[i * Lo + k - width] in y[i * Ln + p].  out_host may be null to query the sizes. */
int sedb_resample_filters(int sr_in, int sr_out, float* out_host, int* width, int* taps, int* phases);
int sedb_resample_f32(sedb_ctx_t* ctx, const float* in_dev, long long n_clips, long long n_in, long long in_stride,
                      int sr_in, int sr_out, float* out_dev, long long out_stride, void* stream);

/* ---- 16-bit PCM input: read_multichannel_audio's channel handling fused into the loader -------------
 * Replaces dataset/dataset_utils.py:63-74 (soundfile.read of a PCM_16 file = int16 / 32768, then
 * `.mean(1)` because common_config.audio_channels == 1) followed by the fused log-mel above.
 * pcm: [n_clips, clip_stride, n_channels] interleaved int16, i.e. the WAV data chunk as stored on disk
 * (n_samples valid sample frames per clip, n_channels in 1..16; 1, 2 and 4 take the vector path).
 * Halves the bytes a clip costs in HBM and over PCIe compared with float32 mono. */
int sedb_logmel_pcm16(sedb_ctx_t* ctx, const int16_t* pcm_dev, long long n_clips, long long n_samples,
                      long long clip_stride, int n_channels, const float* norm_dev, float* out_dev, void* stream);
int sedb_logmel_host_pcm16(sedb_ctx_t* ctx, const int16_t* pcm_host, long long n_clips, long long n_samples,
                           long long clip_stride, int n_channels, const float* norm_host, float* out_host);

/* ---- spectrogram CNN: Cnn_AvgPooling (models/spectogram_models.py:163-205) ----------------------- */
/* channels[i], pools[i]: model_config entries (spectogram_models.py:7, main.py:35); input channels = 1. */
int sedb_cnn_create(sedb_ctx_t* ctx, const int* channels, const int* pools, int n_blocks, int classes_num,
                    sedb_cnn_t** out);
int sedb_cnn_destroy(sedb_cnn_t* cnn);
/* tensors_dev: device float32 pointers in state_dict order per block:
 *   conv1.weight, conv2.weight, bn1.{weight,bias,running_mean,running_var}, bn2.{weight,bias,running_mean,running_var}
 * followed by event_fc.weight, event_fc.bias.  BatchNorm (eps 1e-5) is folded into per-channel scale/shift.
 * Must be called again after the parameters change. */
int sedb_cnn_load(sedb_cnn_t* cnn, const float* const* tensors_dev, int n_tensors, void* stream);
/* frames produced for T input frames: ratio * floor-pooled(T)  (spectogram_models.py:185-202) */
long long sedb_cnn_out_frames(const sedb_cnn_t* cnn, long long T);
size_t sedb_cnn_workspace_bytes(const sedb_cnn_t* cnn, long long n_clips, long long T);
/* The workspace holds the activation planes; their padding pixels must be zero and the kernels never write them.  The
 * handle remembers (host-side) for which (workspace pointer, geometry) pairs it has zeroed the padding and zeroes again
 * only when the pair is new.  It never trusts the memory contents: a caller that frees/re-allocates a workspace, or
 * lets anything else write into it between calls, must call this first (workspace_dev == NULL forgets all of them). */
int sedb_cnn_workspace_invalidate(sedb_cnn_t* cnn, const void* workspace_dev);
/* x_dev: [n_clips, 1, T, 64] float32 log-mel; logits_dev / probs_dev (either may be NULL):
 * [n_clips, out_frames, classes] float32 = forward() / logits() of the reference module. */
int sedb_cnn_forward(sedb_cnn_t* cnn, const float* x_dev, long long n_clips, long long T, float* logits_dev,
                     float* probs_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- training step of Cnn_AvgPooling (train.py:96-103) -------------------------------------------------------------
 * Replaces `model.train(); out = model(x); loss = criterion(out, y); loss.backward()` for the spectrogram CNN:
 * train-mode forward with batch-statistics BatchNorm (models/spectogram_models.py:153-160; eps 1e-5, running
 * statistics updated in place with `momentum`, unbiased variance -- torch.nn.BatchNorm2d defaults), then the backward
 * pass from dL/dlogits to every parameter gradient.  Convolutions (forward, data gradient, weight gradient) run on
 * tcgen05 with bf16 hi+lo split operands and fp32 accumulation; BN statistics are accumulated in fp64.
 * tensors_dev: the same list as sedb_cnn_load (conv weights, BN weight/bias/running_mean/running_var, event_fc), read
 * directly -- no load step; running_mean / running_var are written.
 * The workspace (sedb_cnn_train_workspace_bytes) carries the forward activations to the backward call: call backward
 * with the same workspace, shape and x_dev, before the next forward.  Same zero-padding contract as
 * sedb_cnn_forward (sedb_cnn_workspace_invalidate also forgets training workspaces).
 * Streams.  Both calls are asynchronous on `stream`.  Internally the weight packing (forward) and the weight-gradient
 * GEMMs (backward) run on a stream owned by the handle, forked from and joined back into `stream` with events before the
 * call returns control of the results to `stream`; under stream capture the fork / join becomes graph edges.  One
 * forward / backward pair per handle at a time. */
size_t sedb_cnn_train_workspace_bytes(sedb_cnn_t* cnn, long long n_clips, long long T);
int sedb_cnn_train_forward(sedb_cnn_t* cnn, float* const* tensors_dev, int n_tensors, const float* x_dev,
                           long long n_clips, long long T, float momentum, float* logits_dev, void* workspace_dev,
                           size_t workspace_bytes, void* stream);
/* dlogits_dev: [n_clips, out_frames, classes] = dL/d(forward output).  grads_dev: one float32 buffer per parameter in
 * module.parameters() order -- per block conv1.weight, conv2.weight, bn1.weight, bn1.bias, bn2.weight, bn2.bias, then
 * event_fc.weight, event_fc.bias -- each OVERWRITTEN with the gradient (they may be views of one flat bucket). */
int sedb_cnn_train_backward(sedb_cnn_t* cnn, float* const* tensors_dev, int n_tensors, const float* x_dev,
                            const float* dlogits_dev, long long n_clips, long long T, float* const* grads_dev,
                            int n_grads, void* workspace_dev, size_t workspace_bytes, void* stream);
/* WeightedBCE (utils/common.py:11-30, multi_frame=True): binary_cross_entropy_with_logits(output[:, :N],
 * target[:, :N], pos_weight) with N = min(F_out, F_tgt), mean reduction.  loss_dev[0] receives the loss, dlogits_dev
 * ([B, F_out, K], nullable) grad_scale * dloss/doutput (zero beyond frame N).  Either output may be NULL. */
int sedb_bce_with_logits(const float* logits_dev, const float* target_dev, long long B, long long F_out,
                         long long F_tgt, int K, float pos_weight, float grad_scale, float* loss_dev,
                         float* dlogits_dev, void* stream);

/* ---- waveform CNN: M5 (models/waveform_models.py:9-71) ------------------------------------------ */
int sedb_m5_create(sedb_ctx_t* ctx, int classes_num, sedb_m5_t** out);
int sedb_m5_destroy(sedb_m5_t* m5);
/* tensors_dev in state_dict order: for each Conv1d+BatchNorm1d pair
 *   conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var   (9 pairs)
 * then fc.weight, fc.bias. */
int sedb_m5_load(sedb_m5_t* m5, const float* const* tensors_dev, int n_tensors, void* stream);
size_t sedb_m5_workspace_bytes(const sedb_m5_t* m5, long long n_frames);
int sedb_m5_workspace_invalidate(sedb_m5_t* m5, const void* workspace_dev);   /* see sedb_cnn_workspace_invalidate */
/* x_dev: [n_frames, 1, 31680] float32; logits_dev: [n_frames, classes] float32. */
int sedb_m5_forward(sedb_m5_t* m5, const float* x_dev, long long n_frames, float* logits_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- training step support (train.py:85,101-103) ------------------------------------------------------------ */
/* One fused optimizer update over a flat float32 parameter buffer with the semantics of
 * torch.optim.Adam(lr, betas, eps, weight_decay, amsgrad=True) as constructed at train.py:85:
 *   g = grad * grad_scale (+ weight_decay * p);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  vmax = max(vmax, v)
 *   p -= lr / (1 - b1^step) * m / (sqrt(vmax) / sqrt(1 - b2^step) + eps)
 * grad_scale folds the 1/world_size of the data-parallel gradient mean into the update; step counts from 1. */
int sedb_adam_amsgrad_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                           float* max_exp_avg_sq_dev, long long n, float lr, float beta1, float beta2, float eps,
                           float weight_decay, long long step, float grad_scale, void* stream);

/* The same update with the step counter and the learning rate in device memory, so that a captured CUDA graph of the
 * whole training step can be replayed: state_dev[0] = number of steps done so far (float, incremented by the call),
 * state_dev[1] = lr (train.py:108-110 decays it on the host every 200 iterations: write the new value there);
 * hyper_dev: 2 floats of scratch. */
int sedb_adam_amsgrad_step_dev(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                               float* max_exp_avg_sq_dev, long long n, float* state_dev, float* hyper_dev, float beta1,
                               float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* ---- end-to-end: waveform -> log-mel -> CNN -> frame probabilities -------------------------------- */
/* infer.py:27-33 intent.  wave_host: [n_clips, wave_stride] float32 host; probs_host: [n_clips, out_frames,
 * classes] float32 host.  H2D chunks, log-mel, CNN and the D2H of the probabilities are pipelined. */
int sedb_sed_host_f32(sedb_ctx_t* ctx, sedb_cnn_t* cnn, const float* wave_host, long long n_clips,
                      long long n_samples, long long wave_stride, const float* norm_host, float* probs_host);

/* the same from 16-bit PCM host buffers (see sedb_logmel_pcm16) */
int sedb_sed_host_pcm16(sedb_ctx_t* ctx, sedb_cnn_t* cnn, const int16_t* pcm_host, long long n_clips,
                        long long n_samples, long long clip_stride, int n_channels, const float* norm_host,
                        float* probs_host);

/* ---- diagnostics ---------------------------------------------------------------------------------- */
/* One-CTA tcgen05 GEMM probe used by the GPU tests to pin the shared-memory descriptor conventions:
 * D[128,N] = A[128,K] * B[K,N] with bf16-rounded operands.  a_dev/b_dev/d_dev are float32 row-major.
 * a_major/b_major: 0 = K-major, 1 = MN-major canonical layout; pad: extra bytes added to the 8-row group
 * stride; neg_b: set the B-negate bit. */
int sedb_debug_umma_probe(const float* a_dev, const float* b_dev, float* d_dev, int N, int K, int a_major,
                          int b_major, int pad, int neg_b, int swap_lbo_sbo, void* stream);
/* Workspace layout of the training step for (n_clips, T): out = {layers, G offset, partial-sum offset, bytes} then per
 * conv layer {C_out, H, W, pool, Z offset, Z plane pixels, A offset, A plane pixels, dZ offset, dZ plane pixels, offset of
 * the statistics (in doubles), weight-gradient chunks * 1000 + band pixels}.  Used by the tests to compare intermediates. */
int sedb_debug_train_layout(sedb_cnn_t* cnn, long long n_clips, long long T, long long* out, int max_out);
/* Host-only: the decomposition the planner picks for one tensor-core conv layer (cin -> cout, pool, mode 0 = 3x3 2-D /
 * 1 = k3 1-D, input H x W, amode 0 = inference fp16 / 1 = training bf16 split) over n_img images on num_sms SMs.
 * out8 = {M tiles per item, N sub-items, fused [wH|wL] MMA, bands per image, rows per band, smem bytes, input plane
 * pixels, kpb*100 + weight slots*10 + (32-channel pooling pass)}. */
int sedb_debug_plan_layer(int cin, int cout, int pool, int mode, int ntaps, int H, int W, int amode, long long n_img,
                          int num_sms, int* out8);
/* tcgen05.mma issue-to-completion throughput probe: `grid` CTAs each issue `reps` MMAs of 128 x N x 16 over n_acc
 * accumulator tiles; cycles_host receives block 0's total cycles. */
int sedb_debug_umma_rate(int N, int b_major, int n_acc, int reps, int lbo_a, int lbo_b, int grid,
                         unsigned long long* cycles_host);
/* cp.async.bulk global -> shared throughput probe: `grid` CTAs, `nwarps` issuing warps each keeping `depth` copies of
 * `bytes` in flight, `reps` copies per CTA over nsrc distinct source blocks; cycles_host receives block 0's total cycles. */
int sedb_debug_bulk_rate(int bytes, int depth, int reps, int nsrc, int spin, int grid, int nwarps,
                         unsigned long long* cycles_host);
/* Per-phase cycle counters of logmel_fused_kernel (thread 0 of every CTA, summed): enable != 0 switches the
 * instrumentation on; out_host16 (nullable) receives and clears 128 counters (16 for the log-mel kernel, then 16 per
 * conv layer of the next CNN forward); enable == 0 switches it off. */
int sedb_debug_phase_profile(int enable, unsigned long long* out_host16);
/* Number of kernel launches issued through this library since load (bench.py's gpu_launches). */
long long sedb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SEDB_H_ */
