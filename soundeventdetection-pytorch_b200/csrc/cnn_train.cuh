// Training step of Cnn_AvgPooling (reference train.py:96-103, models/spectogram_models.py:153-160, utils/common.py:16-30):
// train-mode forward with batch-statistics BatchNorm, WeightedBCE on the logits, and the backward pass down to every
// parameter gradient.  The convolutions (forward, data gradient, weight gradient) run on tcgen05 tensor cores; the
// BatchNorm / ReLU / pooling / head / loss pieces are HBM-(really L2-)bound element-wise kernels around them.
//
// Data flow per conv layer l (conv -> BN(batch stats) -> ReLU [-> AvgPool2]):
//   forward   Z_l = conv(A_{l-1}, W_l)                       conv_umma_kernel<1> (raw fp32 planes)        [conv_in2d<1> for l = 0]
//             sums_l = {sum Z, sum Z^2} per channel          bn_stats_kernel (fp64 accumulation)
//             A_l = pool(relu(gamma (Z - mean) rstd + beta)) bn_apply_kernel (bf16 hi + lo planes; running stats updated)
//   backward  G_l = dL/dA_l (fp32 planes; from the head or from the next layer's data gradient)
//             g = G_l (un-pooled, ReLU mask);  s1 = sum g, s2 = sum g xhat                bn_bwd_reduce_kernel
//             dZ_l = gamma rstd (g - s1/n - xhat s2/n); dgamma = s2, dbeta = s1           bn_bwd_apply_kernel (bf16 hi + lo planes)
//             dW_l[co][ci][tap] = sum_pixels dZ_l[p][co] A_{l-1}[p + tap][ci]              wgrad_umma_kernel + wgrad_finalize_kernel
//             G_{l-1} = conv(dZ_l, W_l rotated/transposed)                                 conv_umma_kernel<1> (dgrad)
// Plane layouts are those of cnn.cuh: bf16 planes [img][hi|lo][C/8][S][8], fp32 planes [img][C/8][S][8], padded pixel
// index behind kConvLead lead pixels; padding of the bf16 planes is zero and never written.
#pragma once
#include "cnn.cuh"

namespace sedb {

constexpr float kBnEps = 1e-5f;

// Per-channel constants of one BatchNorm from its batch sums: y = a z + b with a = gamma rstd, b = beta - mean a.
// ONE definition with pinned roundings, used by the forward kernel and by both backward kernels: the ReLU mask of the
// backward pass (a z + b > 0, evaluated with fmaf exactly as the forward did) must agree with the forward bit for bit.
__device__ __forceinline__ void bn_consts(double sum, double sumsq, double inv_n, float gamma, float beta, float& a,
                                          float& b, float& mean_f, float& rstd) {
    const double mean = sum * inv_n;
    double var = sumsq * inv_n - mean * mean;                     // biased variance (what BN normalises with)
    if (var < 0.0) var = 0.0;
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kBnEps)));
    mean_f = static_cast<float>(mean);
    a = __fmul_rn(gamma, rstd);
    b = __fmaf_rn(-mean_f, a, beta);
}
__device__ __forceinline__ void bn_channel_consts(const double* __restrict__ sums, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, int C, double inv_n, float* a_s,
                                                  float* b_s, float* mean_s, float* rstd_s) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a, b, mean, rstd;
        bn_consts(sums[c], sums[C + c], inv_n, gamma[c], beta[c], a, b, mean, rstd);
        a_s[c] = a;
        b_s[c] = b;
        if (mean_s) mean_s[c] = mean;
        if (rstd_s) rstd_s[c] = rstd;
    }
}

// ---- forward: per-channel sum and sum of squares of a conv output (fp32 planes) --------------------------------------
// grid = (chunks, C/8): block (x, kg) walks its share of the (image, pixel) pairs of group kg.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ z, int n_img, int C, int H, int W, int S,
                                                       double* __restrict__ sums) {
    pdl_entry();
    const int kg = blockIdx.y, nkg = C / 8, Wp = W + 2;
    const int total = n_img * H * W;                 // < 2^31 (checked by the host)
    float s[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int w = idx % W, h = (idx / W) % H;
        const long long img = idx / (W * H);
        const float4* p = reinterpret_cast<const float4*>(z + ((img * nkg + kg) * S + kConvLead + (h + 1) * Wp + w + 1) * 8);
        const float4 v0 = p[0], v1 = p[1];
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i] += v[i];
            ss[i] = fmaf(v[i], v[i], ss[i]);
        }
    }
    __shared__ double red[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double a = s[i], b = ss[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            red[warp][i] = a;
            red[warp][8 + i] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
        const int c = kg * 8 + (threadIdx.x & 7);
        atomicAdd(sums + (threadIdx.x < 8 ? c : C + c), a);
    }
}

// ---- forward: BN (batch statistics) + ReLU (+ 2x2 average pooling) -> bf16 hi + lo planes ------------------------------
// One thread per (image, output pixel, 8-channel group).  Block 0 also updates running_mean / running_var
// (momentum 0.1, unbiased variance: torch.nn.BatchNorm2d defaults, as in the reference's ConvBlock).
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                                       float momentum, int n_img, int C, int H, int W, int S_z, int pool,
                                                       int S_out, uint8_t* __restrict__ out) {
    pdl_entry();
    extern __shared__ float bn_s[];
    float* a_s = bn_s;
    float* b_s = bn_s + C;
    const double n = static_cast<double>(n_img) * H * W;
    bn_channel_consts(sums, gamma, beta, C, 1.0 / n, a_s, b_s, nullptr, nullptr);
    if (blockIdx.x == 0 && running_mean != nullptr) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const double mean = sums[c] / n;
            double var = sums[C + c] / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
        }
    }
    __syncthreads();
    const int nkg = C / 8, Wp = W + 2;
    const int Ho = H / pool, Wo = W / pool, Wpo = Wo + 2;
    const int total = n_img * Ho * Wo * nkg;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int wo = idx % Wo;
        const int ho = (idx / Wo) % Ho;
        const int kg = (idx / (Wo * Ho)) % nkg;
        const long long img = idx / (Wo * Ho * nkg);
        const float* zp = z + ((img * nkg + kg) * S_z + kConvLead) * 8;
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = 0.f;
        for (int dy = 0; dy < pool; ++dy)
            for (int dx = 0; dx < pool; ++dx) {
                const int v = (ho * pool + dy + 1) * Wp + wo * pool + dx + 1;
                const float4 v0 = *reinterpret_cast<const float4*>(zp + static_cast<long long>(v) * 8);
                const float4 v1 = *reinterpret_cast<const float4*>(zp + static_cast<long long>(v) * 8 + 4);
                const float t[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] += fmaxf(0.f, __fmaf_rn(t[i], a_s[kg * 8 + i], b_s[kg * 8 + i]));
            }
        if (pool == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] *= 0.25f;
        }
        const long long vout = kConvLead + (ho + 1) * Wpo + wo + 1;
        uint8_t* hi = out + (((img * 2) * nkg + kg) * S_out + vout) * 16;
        uint8_t* lo = out + (((img * 2 + 1) * nkg + kg) * S_out + vout) * 16;
        store_split8(hi, lo, y);
    }
}

// ---- loss: binary cross-entropy on logits with a positive-class weight (utils/common.py:16-30, multi_frame=True) -------
// logits [B, F_out, K], target [B, F_tgt, K]; the first N = min(F_out, F_tgt) frames count; mean over B N K elements.
// Writes loss[0] and dlogits [B, F_out, K] (= grad_scale dloss/dlogits, zero beyond frame N).  One block: the reduction
// order is fixed, the loss is bit-reproducible.
__global__ void __launch_bounds__(1024) bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                          int B, int F_out, int F_tgt, int K, float pos_weight,
                                                          float grad_scale, float* __restrict__ loss,
                                                          float* __restrict__ dlogits) {
    pdl_entry();
    const int N = min(F_out, F_tgt);
    const long long count = static_cast<long long>(B) * N * K;
    const float inv = 1.0f / static_cast<float>(count);
    double acc = 0.0;
    const long long total = static_cast<long long>(B) * F_out * K;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        const int k = static_cast<int>(i % K);
        const int f = static_cast<int>((i / K) % F_out);
        const long long b = i / (static_cast<long long>(K) * F_out);
        float g = 0.f;
        if (f < N) {
            const float x = logits[i];
            const float y = target[(b * F_tgt + f) * K + k];
            const float lw = 1.f + (pos_weight - 1.f) * y;
            const float sp = log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f);      // softplus(-x)
            acc += static_cast<double>((1.f - y) * x + lw * sp);
            const float sig = 1.f / (1.f + expf(-x));
            g = ((1.f - y) - lw * (1.f - sig)) * inv * grad_scale;
        }
        if (dlogits) dlogits[i] = g;
    }
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        acc = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && loss) loss[0] = static_cast<float>(acc / static_cast<double>(count));
    }
}

// ---- head backward: dlogits -> d event_fc.{weight, bias} and G = dL/dA_last (fp32 planes) ------------------------------
// forward (spectogram_models.py:193-200): logits[b, h ratio + r, k] = fc_b[k] + sum_c fc_w[k][c] mean_w A[b, c, h, w].
// One warp per (image, time step); d fc_w / d fc_b are accumulated with float atomics (zeroed by the caller).
__global__ void __launch_bounds__(256) head2d_bwd_kernel(const uint8_t* __restrict__ act, const float* __restrict__ fc_w,
                                                         const float* __restrict__ dlogits, float* __restrict__ d_fc_w,
                                                         float* __restrict__ d_fc_b, float* __restrict__ g_out, int n_img,
                                                         int C, int Hf, int Wf, int S_in, int S_g, int classes, int ratio) {
    pdl_entry();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= n_img * Hf) return;
    const int img = warp_global / Hf, h = warp_global % Hf;
    const int Wp = Wf + 2, nkg = C / 8;
    const float inv_w = 1.0f / static_cast<float>(Wf);
    // activation means of this (image, time step): lane -> groups lane, lane + 32, ...
    for (int kg = lane; kg < nkg; kg += 32) {
        float m[8], gsum[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = gsum[i] = 0.f;
        for (int wq = 0; wq < Wf; ++wq) {
            const long long v = kConvLead + (h + 1) * Wp + wq + 1;
            const uint4 a = *reinterpret_cast<const uint4*>(act + (((static_cast<long long>(img) * 2) * nkg + kg) * S_in + v) * 16);
            const uint4 b = *reinterpret_cast<const uint4*>(act + (((static_cast<long long>(img) * 2 + 1) * nkg + kg) * S_in + v) * 16);
            const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                m[2 * i] += __uint_as_float(wa[i] << 16) + __uint_as_float(wb[i] << 16);
                m[2 * i + 1] += __uint_as_float(wa[i] & 0xffff0000u) + __uint_as_float(wb[i] & 0xffff0000u);
            }
        }
        for (int k = 0; k < classes; ++k) {
            float gk = 0.f;
            for (int r = 0; r < ratio; ++r)
                gk += dlogits[(static_cast<long long>(img) * Hf * ratio + static_cast<long long>(h) * ratio + r) * classes + k];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                atomicAdd(d_fc_w + static_cast<long long>(k) * C + kg * 8 + i, gk * m[i] * inv_w);
                gsum[i] = fmaf(gk, fc_w[static_cast<long long>(k) * C + kg * 8 + i], gsum[i]);
            }
            if (kg == 0) atomicAdd(d_fc_b + k, gk);
        }
        const float4 g0 = make_float4(gsum[0] * inv_w, gsum[1] * inv_w, gsum[2] * inv_w, gsum[3] * inv_w);
        const float4 g1 = make_float4(gsum[4] * inv_w, gsum[5] * inv_w, gsum[6] * inv_w, gsum[7] * inv_w);
        for (int wq = 0; wq < Wf; ++wq) {
            const long long v = kConvLead + (h + 1) * Wp + wq + 1;
            float4* o = reinterpret_cast<float4*>(g_out + ((static_cast<long long>(img) * nkg + kg) * S_g + v) * 8);
            o[0] = g0;
            o[1] = g1;
        }
    }
}

// gradient of the post-ReLU activation of pixel (h, w) for 8 channels: from the pooled gradient planes when the layer
// is followed by 2x2 average pooling (floor mode: the odd last row / column gets no gradient)
__device__ __forceinline__ void load_act_grad(const float* __restrict__ g, long long plane_base, int S_g, int h, int w,
                                              int pool, int Ho, int Wo, float* out8) {
    int hh = h, ww = w;
    float f = 1.f;
    if (pool == 2) {
        hh = h >> 1;
        ww = w >> 1;
        f = 0.25f;
        if (hh >= Ho || ww >= Wo) {
#pragma unroll
            for (int i = 0; i < 8; ++i) out8[i] = 0.f;
            return;
        }
    }
    const int Wg = (pool == 2 ? Wo : Wo) + 2;
    const float4* p = reinterpret_cast<const float4*>(g + (plane_base * S_g + kConvLead + (hh + 1) * Wg + ww + 1) * 8);
    const float4 v0 = p[0], v1 = p[1];
    out8[0] = v0.x * f; out8[1] = v0.y * f; out8[2] = v0.z * f; out8[3] = v0.w * f;
    out8[4] = v1.x * f; out8[5] = v1.y * f; out8[6] = v1.z * f; out8[7] = v1.w * f;
}

// ---- backward: the two BatchNorm reductions  s1 = sum g, s2 = sum g xhat  (g = masked activation gradient) ------------
// grid = (chunks, C/8); bsum[0..C) = s1, bsum[C..2C) = s2 (fp64, zeroed by the caller)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ z, const float* __restrict__ g,
                                                            const double* __restrict__ sums, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int n_img, int C, int H, int W,
                                                            int S_z, int pool, int S_g, double* __restrict__ bsum) {
    pdl_entry();
    __shared__ float a_s[8], b_s[8], mean_s[8], rstd_s[8];
    __shared__ double red[8][16];
    const int kg = blockIdx.y, nkg = C / 8, Wp = W + 2;
    const int Ho = H / pool, Wo = W / pool;
    const double n = static_cast<double>(n_img) * H * W;
    if (threadIdx.x < 8) {
        const int c = kg * 8 + threadIdx.x;
        bn_consts(sums[c], sums[C + c], 1.0 / n, gamma[c], beta[c], a_s[threadIdx.x], b_s[threadIdx.x], mean_s[threadIdx.x],
                  rstd_s[threadIdx.x]);
    }
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    const int total = n_img * H * W;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int w = idx % W, h = (idx / W) % H;
        const long long img = idx / (W * H);
        const float4* p = reinterpret_cast<const float4*>(z + ((img * nkg + kg) * S_z + kConvLead + (h + 1) * Wp + w + 1) * 8);
        const float4 v0 = p[0], v1 = p[1];
        const float zz[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float gg[8];
        load_act_grad(g, img * nkg + kg, S_g, h, w, pool, Ho, Wo, gg);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float gm = (__fmaf_rn(zz[i], a_s[i], b_s[i]) > 0.f) ? gg[i] : 0.f;
            s1[i] += gm;
            s2[i] = fmaf(gm, (zz[i] - mean_s[i]) * rstd_s[i], s2[i]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double a = s1[i], b = s2[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            red[warp][i] = a;
            red[warp][8 + i] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
        const int c = kg * 8 + (threadIdx.x & 7);
        atomicAdd(bsum + (threadIdx.x < 8 ? c : C + c), a);
    }
}

// ---- backward: dZ = gamma rstd (g - s1/n - xhat s2/n) -> bf16 hi + lo planes; dgamma = s2, dbeta = s1 -------------------
// grid = (chunks, C/8).  FIRST != 0 (block0.conv1, C_in = 1): dZ is not stored; the weight gradient
// dW[co][tap] = sum dZ[p][co] x[p + tap] is accumulated right here on CUDA cores (float atomics into d_w, zeroed by
// the caller).
template <int FIRST>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ g,
                                                           const double* __restrict__ sums, const double* __restrict__ bsum,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int n_img, int C, int H, int W, int S_z, int pool, int S_g,
                                                           int S_dz, uint8_t* __restrict__ dz, float* __restrict__ d_gamma,
                                                           float* __restrict__ d_beta, const float* __restrict__ x_in,
                                                           float* __restrict__ d_w) {
    pdl_entry();
    __shared__ float a_s[8], b_s[8], mean_s[8], rstd_s[8], m1_s[8], m2_s[8];
    __shared__ float wred[8][72];
    const int kg = blockIdx.y, nkg = C / 8, Wp = W + 2;
    const int Ho = H / pool, Wo = W / pool;
    const double n = static_cast<double>(n_img) * H * W;
    if (threadIdx.x < 8) {
        const int c = kg * 8 + threadIdx.x;
        bn_consts(sums[c], sums[C + c], 1.0 / n, gamma[c], beta[c], a_s[threadIdx.x], b_s[threadIdx.x], mean_s[threadIdx.x],
                  rstd_s[threadIdx.x]);
        m1_s[threadIdx.x] = static_cast<float>(bsum[c] / n);
        m2_s[threadIdx.x] = static_cast<float>(bsum[C + c] / n);
        if (blockIdx.x == 0) {
            d_beta[c] = static_cast<float>(bsum[c]);
            d_gamma[c] = static_cast<float>(bsum[C + c]);
        }
    }
    __syncthreads();
    float wacc[FIRST ? 72 : 1];
    if (FIRST) {
#pragma unroll
        for (int i = 0; i < 72; ++i) wacc[i] = 0.f;
    }
    const int total = n_img * H * W;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int w = idx % W, h = (idx / W) % H;
        const long long img = idx / (W * H);
        const long long v = kConvLead + (h + 1) * Wp + w + 1;
        const float4* p = reinterpret_cast<const float4*>(z + ((img * nkg + kg) * S_z + v) * 8);
        const float4 v0 = p[0], v1 = p[1];
        const float zz[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float gg[8], d[8];
        load_act_grad(g, img * nkg + kg, S_g, h, w, pool, Ho, Wo, gg);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float gm = (__fmaf_rn(zz[i], a_s[i], b_s[i]) > 0.f) ? gg[i] : 0.f;
            const float xh = (zz[i] - mean_s[i]) * rstd_s[i];
            d[i] = a_s[i] * (gm - m1_s[i] - xh * m2_s[i]);
        }
        if (FIRST) {
            const float* xi = x_in + img * H * W;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int hh = h + kh - 1, ww = w + kw - 1;
                    const float xv = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + hh * W + ww) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) wacc[i * 9 + kh * 3 + kw] = fmaf(d[i], xv, wacc[i * 9 + kh * 3 + kw]);
                }
        } else {
            uint8_t* hi = dz + (((img * 2) * nkg + kg) * S_dz + v) * 16;
            uint8_t* lo = dz + (((img * 2 + 1) * nkg + kg) * S_dz + v) * 16;
            store_split8(hi, lo, d);
        }
    }
    if (FIRST) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 72; ++i) {
            float a = wacc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) wred[warp][i] = a;
        }
        __syncthreads();
        if (threadIdx.x < 72) {
            float a = 0.f;
            for (int w = 0; w < 8; ++w) a += wred[w][threadIdx.x];
            atomicAdd(d_w + static_cast<long long>(kg) * 72 + threadIdx.x, a);     // [co][tap] with co = kg 8 + i
        }
    }
}

// ---- weight gradient on tensor cores -----------------------------------------------------------------------------------
// dW[co][ci][tap] = sum_{img, pixel p} dZ[p][co] X[p + off(tap)][ci].  GEMM view: M = co (128 rows; rows beyond C_out read
// whatever follows in shared memory and are never drained), N = ci, K = pixels (16 per MMA).  Both operands come from
// the blocked planes UNCHANGED: a plane ([pixel][8 channels], 16 B per pixel) is at once the K-major operand of the
// forward conv (row = pixel) and the canonical MN-major operand of this GEMM (MN = channel, K = pixel, 8 K per 128 B):
// LBO = 128 B between groups of 8 pixels, SBO = the plane stride between groups of 8 channels.  A filter tap is the X
// operand shifted by off(tap) pixels.
// CTA (pc, tg, mt): pixel chunk pc (a strided set of (image, band) items), tap row tg (3 taps), 128 x 128 output tile mt.
// Accumulators: 3 taps x N columns of TMEM, kept over all items; drained once into part[pc][tap][ci][co].
constexpr int kWgThreads = 192;            // 4 loader/drain warps + MMA warp + 1 spare (TMEM alloc)
struct WgradParams {
    const uint8_t* dz;       // bf16 hi + lo planes of dZ (C_out channels), plane size S_dz
    const uint8_t* x;        // bf16 hi + lo planes of the layer input (C_in channels), plane size S_x
    float* part;             // [n_pc][9][cin][cout] partial sums
    int n_img, H, W, Wp;
    int cout, cin, S_dz, S_x;
    int Pb;                  // band pixels per item (multiple of 16)
    int n_bands;             // bands per image: ceil((H * Wp) / Pb) over the padded pixel range of the image rows
    int n_pc;                // pixel chunks (grid.x)
    int Px;                  // X patch pixels per item (Pb + 2, rounded up to 8)
    int stage_bytes;         // one stage: dZ patch + X patch
    int n_stages;            // 2..4 stages in flight (the loaders run ahead of the MMAs by n_stages - 1 items)
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_umma_kernel(const WgradParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[10];
    __shared__ uint32_t tmem_ptr_s;
    uint64_t* full = bars + 0;       // [4] stage loaded (4 loader warps)
    uint64_t* empty = bars + 4;      // [4] stage consumed (commit)
    uint64_t* done = bars + 8;       // all MMAs complete
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pc = blockIdx.x, tg = blockIdx.y;
    const int mt = blockIdx.z / ((p.cin + 127) / 128), nt_ = blockIdx.z % ((p.cin + 127) / 128);
    const int co0 = mt * 128, ci0 = nt_ * 128;
    const int Mrows = min(128, p.cout - co0), N = min(128, p.cin - ci0);
    const int nkg_o = p.cout / 8, nkg_i = p.cin / 8;
    const int mkg = Mrows / 8, nkg = N / 8;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&full[i], 4);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 5) tmem_alloc<512>(&tmem_ptr_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_ptr_s;
    pdl_entry();                                       // set-up done (barriers, TMEM); see umma.cuh
    const int items_total = p.n_img * p.n_bands;
    const int n_items = (pc < items_total) ? (items_total - pc + p.n_pc - 1) / p.n_pc : 0;
    // stage layout: dZ [half][mkg][Pb][16 B] | X [half][nkg][Px][16 B]
    const int dz_half = mkg * p.Pb * 16, x_half = nkg * p.Px * 16;
    const int x_off = 2 * dz_half;
    const int dh = tg - 1;

    if (warp < 4) {
        // ---------------------------------------------------------------- loaders: planes -> shared memory with cp.async
        // (16 bytes per copy, no register staging: a thread fires every copy of the item and waits once; the
        // register-staged loop this replaces paid one L2 round trip per plane row, ~14 us per item)
        for (int it = 0; it < n_items; ++it) {
            const int item = pc + it * p.n_pc;
            const int img = item / p.n_bands, band = item % p.n_bands;
            const int v0 = p.Wp + band * p.Pb;                 // first padded pixel of the band (row 1, col 0 = image row 0)
            const int buf = it % p.n_stages, use = it / p.n_stages;
            if (use > 0) mbar_wait(&empty[buf], (use - 1) & 1);
            uint8_t* st = smem + buf * p.stage_bytes;
            // dZ patch: pixels [v0, v0 + Pb) of every (half, plane) row; a warp takes whole rows, lanes run along pixels
            for (int row = warp; row < 2 * mkg; row += 4) {
                const int half = row / mkg, pl = row - half * mkg;
                const long long plane = (static_cast<long long>(img) * 2 + half) * nkg_o + (co0 >> 3) + pl;
                const uint4* src = reinterpret_cast<const uint4*>(p.dz + (plane * p.S_dz + kConvLead + v0) * 16);
                uint4* dst = reinterpret_cast<uint4*>(st + half * dz_half + pl * p.Pb * 16);
                for (int px = lane; px < p.Pb; px += 32) cp_async16(dst + px, src + px);
            }
            // X patch: pixels [v0 + dh Wp - 1, ... + Px)
            const int xs = v0 + dh * p.Wp - 1;
            for (int row = warp; row < 2 * nkg; row += 4) {
                const int half = row / nkg, pl = row - half * nkg;
                const long long plane = (static_cast<long long>(img) * 2 + half) * nkg_i + (ci0 >> 3) + pl;
                const uint4* src = reinterpret_cast<const uint4*>(p.x + (plane * p.S_x + kConvLead + xs) * 16);
                uint4* dst = reinterpret_cast<uint4*>(st + x_off + half * x_half + pl * p.Px * 16);
                for (int px = lane; px < p.Px; px += 32) cp_async16(dst + px, src + px);
            }
            cp_async_wait_all();                               // all of this thread's copies have landed
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[buf]);
        }
        // ---------------------------------------------------------------- drain the accumulators
        mbar_wait(done, 0);
        tc_fence_after();
        const int co = warp * 32 + lane;
        const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        for (int t = 0; t < 3; ++t) {
            const int tap = tg * 3 + t;
            for (int c0 = 0; c0 < N; c0 += 16) {
                float acc[16];
                tmem_ld16(tl + t * 128 + c0, acc);
                tmem_ld_wait();
                if (co < Mrows) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        p.part[((static_cast<long long>(pc) * 9 + tap) * p.cin + ci0 + c0 + i) * p.cout + co0 + co] =
                            (n_items > 0) ? acc[i] : 0.f;
                }
            }
        }
    } else if (warp == 4) {
        // ---------------------------------------------------------------- MMA issuer (converged warp, elected lane)
        if (tmem != 0) __trap();
        const uint32_t idesc = make_idesc(kFmtBF16, kMajorMN, kMajorMN, 128, N);
        for (int it = 0; it < n_items; ++it) {
            const int buf = it % p.n_stages;
            mbar_wait(&full[buf], (it / p.n_stages) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = smem_u32(smem + buf * p.stage_bytes);
                // MN-major: LBO (K groups of 8 pixels) = 128 B, SBO (groups of 8 channels) = plane stride
                const uint64_t aH = make_smem_desc(st, 128, p.Pb * 16);
                const uint64_t aL = make_smem_desc(st + dz_half, 128, p.Pb * 16);
                const uint64_t bH = make_smem_desc(st + x_off, 128, p.Px * 16);
                const uint64_t bL = make_smem_desc(st + x_off + x_half, 128, p.Px * 16);
                for (int ks = 0; ks < p.Pb / 16; ++ks) {
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    const uint32_t ka = ks * 16;                       // 16 pixels = 16 x 16 B = 16 units
#pragma unroll
                    for (int t = 0; t < 3; ++t) {                      // tap dw = t - 1: X shifted by t pixels
                        umma_f16(t * 128, aH + ka, bH + ka + t, idesc, acc);
                        umma_f16(t * 128, aL + ka, bH + ka + t, idesc, 1u);
                        umma_f16(t * 128, aH + ka, bL + ka + t, idesc, 1u);
                    }
                }
                umma_commit(&empty[buf]);
                if (it + 1 == n_items) umma_commit(done);
            }
            __syncwarp();
        }
        if (n_items == 0) {
            if (elect_one()) mbar_arrive(done);
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// dW[co][ci][tap] = sum_pc part[pc][tap][ci][co]  (fixed order: bit-reproducible), written into the gradient buffers.
// One launch for all tensor-core layers of the model: blockIdx.y = layer.
constexpr int kTrainMaxLayers = 16;
struct WgradFinalizeAll {
    int n;
    struct {
        const float* part;
        float* d_w;
        int n_pc, cout, cin;
    } L[kTrainMaxLayers];
};
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const WgradFinalizeAll f) {
    pdl_entry();
    const auto& L = f.L[blockIdx.y];
    const long long total = static_cast<long long>(L.cout) * L.cin * 9;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(idx % L.cout);
        const int ci = static_cast<int>((idx / L.cout) % L.cin);
        const int tap = static_cast<int>(idx / (static_cast<long long>(L.cout) * L.cin));
        float a = 0.f;
        for (int pc = 0; pc < L.n_pc; ++pc) a += L.part[((static_cast<long long>(pc) * 9 + tap) * L.cin + ci) * L.cout + co];
        L.d_w[(static_cast<long long>(co) * L.cin + ci) * 9 + tap] = a;
    }
}

// bf16 hi|lo packs of every tensor-core layer's current weights in ONE launch (blockIdx.y = layer): the forward
// convolution and its data-gradient transpose (see pack_conv_weight_kernel for the layout)
struct PackTrainAll {
    int n;
    struct {
        const float* w;
        uint8_t* fwd;
        uint8_t* dgrad;
        int cout, cin, ct, ck, dct, dck;          // forward: cout tile / cin chunk; data gradient: its cout (= cin) tile / cin (= cout) chunk
    } L[kTrainMaxLayers];
};
__device__ __forceinline__ void pack_store_bf16(uint8_t* out, float v, int co, int ci, int tap, int cout_tile, int cin_chunk,
                                                int cin) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const int ntile = co / cout_tile, n = co % cout_tile;
    const int kc = ci / cin_chunk, cil = ci % cin_chunk;
    const int ks = cil / 16, k = cil % 16;
    const int n_kchunks = cin / cin_chunk, ks_chunk = cin_chunk / 16;
    const long long block = ((static_cast<long long>(ntile) * n_kchunks + kc) * ks_chunk + ks) * 9 + tap;
    const long long base = block * (cout_tile * 64);
    const int off = (k / 8) * (cout_tile * 32) + n * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(out + base + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(out + base + cout_tile * 16 + off) = l;
}
__global__ void __launch_bounds__(256) pack_train_weights_kernel(const PackTrainAll f) {
    pdl_entry();
    const auto& L = f.L[blockIdx.y];
    const long long total = static_cast<long long>(L.cout) * L.cin * 9;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(idx % 9);
        const int ci = static_cast<int>((idx / 9) % L.cin);
        const int co = static_cast<int>(idx / (9LL * L.cin));
        const float v = L.w[idx];
        pack_store_bf16(L.fwd, v, co, ci, tap, L.ct, L.ck, L.cin);
        pack_store_bf16(L.dgrad, v, ci, co, 8 - tap, L.dct, L.dck, L.cout);     // w'[ci][co][8 - tap] = w[co][ci][tap]
    }
}

// ---- Adam(amsgrad) with the step count and learning rate in device memory (CUDA-graph friendly) ----------------------
// state[0] = step (as float, incremented here), state[1] = lr; hyper[0..1] receive step_size and 1/sqrt(bc2)
__global__ void adam_prepare_kernel(float* __restrict__ state, float beta1, float beta2, float* __restrict__ hyper) {
    pdl_entry();
    const double step = static_cast<double>(state[0]) + 1.0;
    state[0] = static_cast<float>(step);
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    hyper[0] = static_cast<float>(static_cast<double>(state[1]) / bc1);
    hyper[1] = static_cast<float>(1.0 / sqrt(bc2));
}
__global__ void __launch_bounds__(256) adam_amsgrad_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                               float* __restrict__ m, float* __restrict__ v,
                                                               float* __restrict__ vmax, long long n,
                                                               const float* __restrict__ hyper, float beta1, float beta2,
                                                               float eps, float weight_decay, float grad_scale) {
    pdl_entry();
    const float step_size = hyper[0], inv_bc2_sqrt = hyper[1];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float pi = p[i];
        float gi = g[i] * grad_scale;
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
        const float vm = fmaxf(vmax[i], vi);
        m[i] = mi;
        v[i] = vi;
        vmax[i] = vm;
        p[i] = pi - step_size * (mi / (sqrtf(vm) * inv_bc2_sqrt + eps));
    }
}

}  // namespace sedb
