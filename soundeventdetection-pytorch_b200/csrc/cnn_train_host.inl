// Host side of the Cnn_AvgPooling training step (included by sedb.cu after cnn_host.inl): per-shape plan, buffers,
// launches of cnn_train.cuh.  Reference: train.py:96-103, models/spectogram_models.py:153-160,185-202.

#ifndef SEDB_WGRAD_SIDE_STREAM
#define SEDB_WGRAD_SIDE_STREAM 1
#endif

namespace {

struct TrainLayer {               // conv layer l = 0 .. nl (0 = block0.conv1 on CUDA cores)
    int cin = 0, cout = 0, H = 0, W = 0;
    int pool = 1;                 // pooling applied by this layer's BN/ReLU kernel (2 after the second conv of a pooled block)
    int block = 0, which = 0;     // ConvBlock index, 0 = conv1/bn1, 1 = conv2/bn2
    PlaneGeom Z, A, dZ;           // conv output (fp32), post BN/ReLU/pool activation (bf16 hi+lo), dL/dZ (bf16 hi+lo)
    sedb::ConvParams fwd, dgrad;  // l >= 1
    sedb::WgradParams wg;         // l >= 1
    size_t wg_smem = 0;
    int wg_tiles = 1;
    size_t sums_off = 0;          // doubles: [2C] forward sums, then [2C] backward sums
    size_t part_off = 0;          // weight-gradient partial sums of this layer
};

struct TrainPlan {
    std::vector<TrainLayer> L;
    size_t stats_bytes = 0, g_off = 0, g_bytes = 0, part_off = 0, part_bytes = 0, ws_bytes = 0;
    int Hf = 0, Wf = 0;
    unsigned long long tag = 0;
};

// weight-gradient decomposition of one layer: band size from the shared-memory budget, pixel chunks from the SM count
void plan_wgrad(TrainLayer& T, long long n_img, int num_sms) {
    sedb::WgradParams& w = T.wg;
    w = sedb::WgradParams{};
    w.n_img = static_cast<int>(n_img);
    w.H = T.H;
    w.W = T.W;
    w.Wp = T.W + 2;
    w.cout = T.cout;
    w.cin = T.cin;
    const int mkg = std::min(128, T.cout) / 8, nkg = std::min(128, T.cin) / 8;
    const int pixels = T.H * w.Wp;                                   // padded pixel range of the image rows
    int Pb = std::min(128, round_up(pixels, 16));
    for (;; Pb -= 16) {
        const int Px = round_up(Pb + 2, 8);
        const size_t stage = static_cast<size_t>(2) * (mkg * Pb + nkg * Px) * 16;
        const size_t slack = static_cast<size_t>(16) * Pb * 16 + 1024;      // M = 128 rows are read whatever C_out is
        if (2 * stage + slack <= 220 * 1024 || Pb <= 16) {
            w.Pb = Pb;
            w.Px = Px;
            w.stage_bytes = static_cast<int>(stage);
            w.n_stages = static_cast<int>(std::max<size_t>(2, std::min<size_t>(4, (220 * 1024 - slack) / stage)));
            T.wg_smem = w.n_stages * stage + slack;
            break;
        }
    }
    w.n_bands = (pixels + w.Pb - 1) / w.Pb;
    T.wg_tiles = ((T.cout + 127) / 128) * ((T.cin + 127) / 128);
    const long long items = n_img * w.n_bands;
    long long n_pc = (items + 3) / 4;                                // ~4 items per CTA: latency matters more than the drain
    const long long cap = std::max(1, num_sms / (3 * T.wg_tiles));
    w.n_pc = static_cast<int>(std::max<long long>(1, std::min(n_pc, cap)));
}

int wgrad_plane_S(const TrainLayer& T) {     // pixels a plane must hold for the weight-gradient patches of this layer
    return round_up(sedb::kConvLead + T.wg.Wp + T.wg.n_bands * T.wg.Pb + T.wg.Wp + 2 + 16, 8);
}

}  // namespace

struct sedb_cnn_train {
    std::map<std::pair<long long, long long>, TrainPlan> plans;
    std::vector<uint8_t*> wpack_fwd, wpack_dgrad;     // per umma layer: bf16 hi|lo packs (forward / data-gradient convolution)
    ZeroedSet zeroed;
    // The weight-gradient GEMM of a layer only needs dZ_l and A_{l-1}; nothing downstream of it but the final reduction
    // reads its output.  It runs on a stream of its own, forked off the caller's stream once dZ_l is written and joined
    // before wgrad_finalize, so the (latency-bound) weight gradients overlap the data-gradient / BatchNorm chain.  The
    // fork / join is by events, which a stream capture (the trainer's CUDA graph) turns into graph edges.
    cudaStream_t s_wgrad = nullptr;
    std::vector<cudaEvent_t> ev_fork;                  // one per umma layer
    cudaEvent_t ev_join = nullptr;
};

static void sedb_cnn_train_free(sedb_cnn* m) {
    if (!m || !m->train) return;
    for (auto p : m->train->wpack_fwd) cudaFree(p);
    for (auto p : m->train->wpack_dgrad) cudaFree(p);
    for (auto e : m->train->ev_fork) cudaEventDestroy(e);
    if (m->train->ev_join) cudaEventDestroy(m->train->ev_join);
    if (m->train->s_wgrad) cudaStreamDestroy(m->train->s_wgrad);
    delete m->train;
    m->train = nullptr;
}

static void sedb_cnn_train_invalidate(sedb_cnn* m, const void* ws) {
    if (m && m->train) m->train->zeroed.drop(ws);
}

static int cnn_train_state(sedb_cnn* m) {
    if (m->train) return 0;
    sedb_cnn_train* t = new (std::nothrow) sedb_cnn_train();
    if (!t) return fail("out of host memory");
    m->train = t;
    for (const UmmaLayer& L : m->layers) {
        uint8_t *a = nullptr, *b = nullptr;
        CUDA_TRY(cudaMalloc(&a, L.pack_bytes()));
        t->wpack_fwd.push_back(a);
        CUDA_TRY(cudaMalloc(&b, L.pack_bytes()));
        t->wpack_dgrad.push_back(b);
    }
    CUDA_TRY(cudaFuncSetAttribute(sedb::wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CUDA_TRY(cudaStreamCreateWithFlags(&t->s_wgrad, cudaStreamNonBlocking));
    for (size_t i = 0; i < m->layers.size(); ++i) {
        cudaEvent_t e = nullptr;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        t->ev_fork.push_back(e);
    }
    CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
    return 0;
}

static int cnn_make_train_plan(const sedb_cnn* m, long long n_clips, long long T, TrainPlan& plan) {
    const int nl = static_cast<int>(m->layers.size());
    const int num_sms = m->ctx->num_sms;
    plan.L.assign(nl + 1, TrainLayer{});
    int H = static_cast<int>(T), W = SEDB_MEL_BINS;
    size_t sums = 0;
    for (int l = 0; l <= nl; ++l) {
        TrainLayer& t = plan.L[l];
        t.block = l / 2;                           // l = 0: conv1 of block 0; odd l: conv2; even l: conv1
        t.which = l % 2;
        t.cin = (l == 0) ? 1 : m->layers[l - 1].cin;
        t.cout = (l == 0) ? m->channels[0] : m->layers[l - 1].cout;
        t.H = H;
        t.W = W;
        t.pool = (t.which == 1) ? m->pools[t.block] : 1;
        t.sums_off = sums;
        sums += static_cast<size_t>(4) * t.cout;
        if (l >= 1) {
            UmmaLayer L = m->layers[l - 1];
            L.pool = 1;                                              // pooling happens after BN/ReLU in train mode
            const int S_in = plan_umma_layer(L, H, W, 1, n_clips, num_sms, t.fwd);
            if (S_in < 0) return fail("unsupported feature-map size %d x %d", H, W);
            UmmaLayer D = L;                                         // data gradient: C_out -> C_in, same spatial size
            D.cin = L.cout;
            D.cout = L.cin;
            D.cin_chunk = D.cin > 128 ? 128 : D.cin;
            D.cout_tile = D.cout > 128 ? 128 : D.cout;
            const int S_dz = plan_umma_layer(D, H, W, 1, n_clips, num_sms, t.dgrad);
            if (S_dz < 0) return fail("unsupported feature-map size %d x %d", H, W);
            plan_wgrad(t, n_clips, num_sms);
            const int S_wg = wgrad_plane_S(t);
            plan.L[l - 1].A.S = std::max(S_in, S_wg);
            t.dZ.S = std::max(S_dz, S_wg);
        } else {
            t.dZ.S = 0;                                              // block0.conv1: dZ is consumed in registers
        }
        t.Z.C = t.cout; t.Z.H = H; t.Z.W = W; t.Z.elt = 4;
        t.Z.S = final_plane_S(0, H, W);
        t.dZ.C = t.cout; t.dZ.H = H; t.dZ.W = W; t.dZ.elt = 4;
        const int Ho = H / t.pool, Wo = W / t.pool;
        if (Ho < 1 || Wo < 1) return fail("input of %lld frames is too short for this model's pooling", T);
        t.A.C = t.cout; t.A.H = Ho; t.A.W = Wo; t.A.elt = 4;
        H = Ho;
        W = Wo;
    }
    plan.L[nl].A.S = final_plane_S(0, H, W);
    plan.Hf = H;
    plan.Wf = W;
    size_t off = 0;
    plan.stats_bytes = (sums * sizeof(double) + 127) / 128 * 128;
    off += plan.stats_bytes;
    unsigned long long tag = mix_tag(0x7A1Eull, static_cast<unsigned long long>(n_clips));
    tag = mix_tag(tag, static_cast<unsigned long long>(T));
    auto place = [&](PlaneGeom& g) {
        g.offset = off;
        off += (g.bytes_per_img() * static_cast<size_t>(n_clips) + 127) / 128 * 128;
        tag = mix_tag(tag, (static_cast<unsigned long long>(g.C) << 40) ^ (static_cast<unsigned long long>(g.S) << 8) ^ g.W);
    };
    size_t g_max = 0, part_max = 0;
    for (int l = 0; l <= nl; ++l) {
        TrainLayer& t = plan.L[l];
        place(t.Z);
        place(t.A);
        if (l >= 1) place(t.dZ);
        // gradient planes w.r.t. this layer's activation A_l (fp32, geometry of A)
        const size_t gbytes = static_cast<size_t>(n_clips) * (t.cout / 8) * final_plane_S(0, t.A.H, t.A.W) * 32;
        g_max = std::max(g_max, gbytes);
    }
    plan.g_off = off;
    plan.g_bytes = (g_max + 127) / 128 * 128;
    off += plan.g_bytes;
    plan.part_off = off;
    for (int l = 1; l <= nl; ++l) {
        TrainLayer& t = plan.L[l];
        t.part_off = off;
        off += (static_cast<size_t>(t.wg.n_pc) * 9 * t.cin * t.cout * 4 + 127) / 128 * 128;
    }
    plan.part_bytes = off - plan.part_off;
    (void)part_max;
    plan.ws_bytes = off;
    plan.tag = tag | 1ull;
    return 0;
}

static int cnn_get_train_plan(sedb_cnn* m, long long n_clips, long long T, const TrainPlan** out) {
    if (int rc = cnn_train_state(m)) return rc;
    const auto key = std::make_pair(n_clips, T);
    auto it = m->train->plans.find(key);
    if (it == m->train->plans.end()) {
        TrainPlan plan;
        if (int rc = cnn_make_train_plan(m, n_clips, T, plan)) return rc;
        if (m->train->plans.size() >= 16) m->train->plans.clear();
        it = m->train->plans.emplace(key, std::move(plan)).first;
    }
    *out = &it->second;
    return 0;
}

static int ew_blocks(long long total, int cap = 148 * 8) {
    long long b = (total + 255) / 256;
    return static_cast<int>(std::max<long long>(1, std::min<long long>(b, cap)));
}

extern "C" {

size_t sedb_cnn_train_workspace_bytes(sedb_cnn_t* m, long long n_clips, long long T) {
    if (!m || n_clips <= 0 || T <= 0) return 0;
    const TrainPlan* plan = nullptr;
    if (cnn_get_train_plan(m, n_clips, T, &plan)) return 0;
    return plan->ws_bytes;
}

int sedb_debug_train_layout(sedb_cnn_t* m, long long n_clips, long long T, long long* out, int max_out) {
    if (!m || !out) return fail("null argument");
    const TrainPlan* plan = nullptr;
    if (int rc = cnn_get_train_plan(m, n_clips, T, &plan)) return rc;
    const int n = static_cast<int>(plan->L.size());
    if (max_out < 4 + 12 * n) return fail("need %d entries", 4 + 12 * n);
    out[0] = n;
    out[1] = static_cast<long long>(plan->g_off);
    out[2] = static_cast<long long>(plan->part_off);
    out[3] = static_cast<long long>(plan->ws_bytes);
    for (int l = 0; l < n; ++l) {
        const TrainLayer& t = plan->L[l];
        long long* o = out + 4 + 12 * l;
        o[0] = t.cout; o[1] = t.H; o[2] = t.W; o[3] = t.pool;
        o[4] = static_cast<long long>(t.Z.offset); o[5] = t.Z.S;
        o[6] = static_cast<long long>(t.A.offset); o[7] = t.A.S;
        o[8] = static_cast<long long>(t.dZ.offset); o[9] = t.dZ.S;
        o[10] = static_cast<long long>(t.sums_off); o[11] = t.wg.n_pc * 1000 + t.wg.Pb;
    }
    return 0;
}

int sedb_cnn_train_forward(sedb_cnn_t* m, float* const* t, int n_tensors, const float* x_dev, long long n_clips,
                           long long T, float momentum, float* logits_dev, void* workspace_dev, size_t workspace_bytes,
                           void* stream) {
    if (!m || !t || !x_dev || !logits_dev || !workspace_dev) return fail("sedb_cnn_train_forward: null argument");
    if (n_tensors != 10 * m->n_blocks + 2)
        return fail("sedb_cnn_train_forward: expected %d tensors, got %d", 10 * m->n_blocks + 2, n_tensors);
    for (int i = 0; i < n_tensors; ++i)
        if (!t[i]) return fail("sedb_cnn_train_forward: tensor %d is null", i);
    if (n_clips < 1 || T < 1 || n_clips > (1 << 20) || T > (1 << 20) ||
        n_clips * T * SEDB_MEL_BINS * m->channels[0] / 8 >= (1LL << 31))
        return fail("sedb_cnn_train_forward: bad shape (the training kernels index up to 2^31 (pixel, 8-channel group) units)");
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 127) return fail("workspace must be 128-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TrainPlan* planp = nullptr;
    if (int rc = cnn_get_train_plan(m, n_clips, T, &planp)) return rc;
    const TrainPlan& plan = *planp;
    if (workspace_bytes < plan.ws_bytes)
        return fail("sedb_cnn_train_forward: workspace has %zu bytes, needs %zu", workspace_bytes, plan.ws_bytes);
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    if (int rc = prepare_workspace(m->train->zeroed, ws, plan.ws_bytes, plan.tag, st)) return rc;
    CUDA_TRY(cudaMemsetAsync(ws, 0, plan.stats_bytes, st));
    double* stats = reinterpret_cast<double*>(ws);
    const int n_img = static_cast<int>(n_clips);
    const int nl = static_cast<int>(m->layers.size());
    // bf16 hi|lo packs of the current weights: forward convolution and its data-gradient transpose (one launch)
    if (nl > sedb::kTrainMaxLayers) return fail("too many conv layers for the training step");
    bool pack_forked = false;
    {
        sedb::PackTrainAll pk;
        pk.n = nl;
        long long max_total = 0;
        for (int i = 0; i < nl; ++i) {
            const UmmaLayer& L = m->layers[i];
            const int b = (i + 1) / 2, which = (i + 1) % 2;
            pk.L[i].w = t[10 * b + which];
            pk.L[i].fwd = m->train->wpack_fwd[i];
            pk.L[i].dgrad = m->train->wpack_dgrad[i];
            pk.L[i].cout = L.cout;
            pk.L[i].cin = L.cin;
            pk.L[i].ct = L.cout_tile;
            pk.L[i].ck = L.cin_chunk;
            pk.L[i].dct = L.cin > 128 ? 128 : L.cin;
            pk.L[i].dck = L.cout > 128 ? 128 : L.cout;
            max_total = std::max(max_total, static_cast<long long>(L.cout) * L.cin * L.ntaps);
        }
        dim3 pgrid(static_cast<unsigned>(std::min<long long>((max_total + 255) / 256, 148)), nl);
        if (SEDB_WGRAD_SIDE_STREAM) {
            // the packs are first needed by layer 1: they are written on the side stream while layer 0 runs
            CUDA_TRY(cudaEventRecord(m->train->ev_fork[0], st));
            CUDA_TRY(cudaStreamWaitEvent(m->train->s_wgrad, m->train->ev_fork[0], 0));
            sedb::pack_train_weights_kernel<<<pgrid, 256, 0, m->train->s_wgrad>>>(pk);
            CUDA_TRY(cudaEventRecord(m->train->ev_join, m->train->s_wgrad));
            pack_forked = true;
        } else {
            CUDA_TRY(launch_pdl(sedb::pack_train_weights_kernel, dim3(pgrid), dim3(256), 0, st, pk));
        }
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaGetLastError());
    for (int l = 0; l <= nl; ++l) {
        if (l == 1 && pack_forked) CUDA_TRY(cudaStreamWaitEvent(st, m->train->ev_join, 0));
        const TrainLayer& tl = plan.L[l];
        float* const* q = t + 10 * tl.block;
        const float* gamma = q[2 + 4 * tl.which];
        const float* beta = q[3 + 4 * tl.which];
        float* rmean = q[4 + 4 * tl.which];
        float* rvar = q[5 + 4 * tl.which];
        if (l == 0) {
            const bool px4 = true;                  // four rows per thread
            const long long total = static_cast<long long>(n_img) * (px4 ? (tl.H + 3) / 4 : tl.H) * tl.W;
            long long blocks = (total + 255) / 256;
            if (!px4 && blocks > 148LL * 16) blocks = 148LL * 16;
            const size_t smem = static_cast<size_t>(tl.cout) * 11 * sizeof(float);
            if (px4)
                CUDA_TRY(launch_pdl(sedb::conv_in2d_px4_kernel<1>, dim3(static_cast<int>(blocks)), dim3(256), smem, st, 
                    x_dev, q[0], nullptr, nullptr, ws + tl.Z.offset, n_img, tl.H, tl.W, tl.cout, tl.Z.S));
            else
                CUDA_TRY(launch_pdl(sedb::conv_in2d_kernel<1>, dim3(static_cast<int>(blocks)), dim3(256), smem, st, 
                    x_dev, q[0], nullptr, nullptr, ws + tl.Z.offset, n_img, tl.H, tl.W, tl.cout, tl.Z.S));
            g_launches.fetch_add(1);
        } else {
            if (int rc = launch_umma_layer<1>(m->ctx, m->train->wpack_fwd[l - 1], 1, 0, nullptr, nullptr, tl.fwd,
                                              ws + plan.L[l - 1].A.offset, ws + tl.Z.offset, n_img, plan.L[l - 1].A.S,
                                              tl.Z.S, st))
                return rc;
        }
        const long long px = static_cast<long long>(n_img) * tl.H * tl.W;
        dim3 sgrid(ew_blocks(px, 64), tl.cout / 8);
        CUDA_TRY(launch_pdl(sedb::bn_stats_kernel, dim3(sgrid), dim3(256), 0, st, reinterpret_cast<const float*>(ws + tl.Z.offset), n_img, tl.cout, tl.H,
                                                    tl.W, tl.Z.S, stats + tl.sums_off));
        const long long units = static_cast<long long>(n_img) * tl.A.H * tl.A.W * (tl.cout / 8);
        CUDA_TRY(launch_pdl(sedb::bn_apply_kernel, dim3(ew_blocks(units)), dim3(256), 2 * tl.cout * sizeof(float), st, 
            reinterpret_cast<const float*>(ws + tl.Z.offset), stats + tl.sums_off, gamma, beta, rmean, rvar, momentum, n_img,
            tl.cout, tl.H, tl.W, tl.Z.S, tl.pool, tl.A.S, ws + tl.A.offset));
        g_launches.fetch_add(2);
        CUDA_TRY(cudaGetLastError());
    }
    {
        const TrainLayer& tl = plan.L[nl];
        const long long warps = static_cast<long long>(n_img) * plan.Hf;
        const int blocks = static_cast<int>((warps * 32 + 255) / 256);
        CUDA_TRY(launch_pdl(sedb::head2d_kernel<1>, dim3(blocks), dim3(256), 0, st, ws + tl.A.offset, t[10 * m->n_blocks], t[10 * m->n_blocks + 1],
                                                      logits_dev, nullptr, n_img, tl.cout, plan.Hf, plan.Wf, tl.A.S,
                                                      m->classes, m->ratio));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int sedb_cnn_train_backward(sedb_cnn_t* m, float* const* t, int n_tensors, const float* x_dev, const float* dlogits_dev,
                            long long n_clips, long long T, float* const* grads, int n_grads, void* workspace_dev,
                            size_t workspace_bytes, void* stream) {
    if (!m || !t || !x_dev || !dlogits_dev || !grads || !workspace_dev) return fail("sedb_cnn_train_backward: null argument");
    if (n_tensors != 10 * m->n_blocks + 2 || n_grads != 6 * m->n_blocks + 2)
        return fail("sedb_cnn_train_backward: expected %d tensors and %d gradients, got %d and %d", 10 * m->n_blocks + 2,
                    6 * m->n_blocks + 2, n_tensors, n_grads);
    for (int i = 0; i < n_grads; ++i)
        if (!grads[i]) return fail("sedb_cnn_train_backward: gradient %d is null", i);
    if (!m->train) return fail("sedb_cnn_train_backward: no forward pass has been run");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const TrainPlan* planp = nullptr;
    if (int rc = cnn_get_train_plan(m, n_clips, T, &planp)) return rc;
    const TrainPlan& plan = *planp;
    if (workspace_bytes < plan.ws_bytes) return fail("sedb_cnn_train_backward: workspace too small");
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    if (!m->train->zeroed.has(ws, plan.tag))
        return fail("sedb_cnn_train_backward: the workspace does not hold the forward pass of this shape");
    double* stats = reinterpret_cast<double*>(ws);
    float* G = reinterpret_cast<float*>(ws + plan.g_off);
    sedb::WgradFinalizeAll fin;
    fin.n = static_cast<int>(m->layers.size());
    const int n_img = static_cast<int>(n_clips);
    const int nl = static_cast<int>(m->layers.size());
    const int nb = m->n_blocks;
    const int Cl = m->channels[nb - 1];
    float* d_fc_w = grads[6 * nb];
    float* d_fc_b = grads[6 * nb + 1];
    CUDA_TRY(cudaMemsetAsync(d_fc_w, 0, static_cast<size_t>(m->classes) * Cl * sizeof(float), st));
    CUDA_TRY(cudaMemsetAsync(d_fc_b, 0, m->classes * sizeof(float), st));
    CUDA_TRY(cudaMemsetAsync(grads[0], 0, static_cast<size_t>(m->channels[0]) * 9 * sizeof(float), st));   // block0.conv1: atomics
    {
        const TrainLayer& tl = plan.L[nl];
        const long long warps = static_cast<long long>(n_img) * plan.Hf;
        const int blocks = static_cast<int>((warps * 32 + 255) / 256);
        CUDA_TRY(launch_pdl(sedb::head2d_bwd_kernel, dim3(blocks), dim3(256), 0, st, ws + tl.A.offset, t[10 * nb], dlogits_dev, d_fc_w, d_fc_b, G, n_img,
                                                       tl.cout, plan.Hf, plan.Wf, tl.A.S, final_plane_S(0, plan.Hf, plan.Wf),
                                                       m->classes, m->ratio));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    bool forked = false;
    for (int l = nl; l >= 0; --l) {
        const TrainLayer& tl = plan.L[l];
        float* const* q = t + 10 * tl.block;
        const float* gamma = q[2 + 4 * tl.which];
        const float* beta = q[3 + 4 * tl.which];
        float* d_w = grads[6 * tl.block + tl.which];
        float* d_gamma = grads[6 * tl.block + 2 + 2 * tl.which];
        float* d_beta = grads[6 * tl.block + 3 + 2 * tl.which];
        const float* Z = reinterpret_cast<const float*>(ws + tl.Z.offset);
        const int S_g = final_plane_S(0, tl.A.H, tl.A.W);
        const long long px = static_cast<long long>(n_img) * tl.H * tl.W;
        dim3 rgrid(ew_blocks(px, 64), tl.cout / 8);
        CUDA_TRY(launch_pdl(sedb::bn_bwd_reduce_kernel, dim3(rgrid), dim3(256), 0, st, Z, G, stats + tl.sums_off, gamma, beta, n_img, tl.cout, tl.H, tl.W,
                                                         tl.Z.S, tl.pool, S_g, stats + tl.sums_off + 2 * tl.cout));
        dim3 agrid(ew_blocks(px, 128), tl.cout / 8);
        if (l == 0) {
            CUDA_TRY(launch_pdl(sedb::bn_bwd_apply_kernel<1>, dim3(agrid), dim3(256), 0, st, Z, G, stats + tl.sums_off, stats + tl.sums_off + 2 * tl.cout,
                                                               gamma, beta, n_img, tl.cout, tl.H, tl.W, tl.Z.S, tl.pool, S_g,
                                                               0, nullptr, d_gamma, d_beta, x_dev, d_w));
            g_launches.fetch_add(2);
            CUDA_TRY(cudaGetLastError());
            break;
        }
        if (l > sedb::kTrainMaxLayers) return fail("too many conv layers for the training step");
        CUDA_TRY(launch_pdl(sedb::bn_bwd_apply_kernel<0>, dim3(agrid), dim3(256), 0, st, Z, G, stats + tl.sums_off, stats + tl.sums_off + 2 * tl.cout, gamma,
                                                           beta, n_img, tl.cout, tl.H, tl.W, tl.Z.S, tl.pool, S_g, tl.dZ.S,
                                                           ws + tl.dZ.offset, d_gamma, d_beta, nullptr, nullptr));
        g_launches.fetch_add(2);
        // weight gradient
        sedb::WgradParams wp = tl.wg;
        wp.dz = ws + tl.dZ.offset;
        wp.x = ws + plan.L[l - 1].A.offset;
        wp.part = reinterpret_cast<float*>(ws + tl.part_off);
        wp.S_dz = tl.dZ.S;
        wp.S_x = plan.L[l - 1].A.S;
        dim3 wgrid(wp.n_pc, 3, tl.wg_tiles);
        if (SEDB_WGRAD_SIDE_STREAM) {
            CUDA_TRY(cudaEventRecord(m->train->ev_fork[l - 1], st));                 // dZ_l is complete on the caller's stream
            CUDA_TRY(cudaStreamWaitEvent(m->train->s_wgrad, m->train->ev_fork[l - 1], 0));
            sedb::wgrad_umma_kernel<<<wgrid, sedb::kWgThreads, tl.wg_smem, m->train->s_wgrad>>>(wp);
            forked = true;
        } else {
            CUDA_TRY(launch_pdl(sedb::wgrad_umma_kernel, dim3(wgrid), dim3(sedb::kWgThreads), tl.wg_smem, st, wp));
        }
        fin.L[l - 1].part = wp.part;
        fin.L[l - 1].d_w = d_w;
        fin.L[l - 1].n_pc = wp.n_pc;
        fin.L[l - 1].cout = tl.cout;
        fin.L[l - 1].cin = tl.cin;
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        // data gradient: G = dL/dA_{l-1}
        if (int rc = launch_umma_layer<1>(m->ctx, m->train->wpack_dgrad[l - 1], 1, 0, nullptr, nullptr, tl.dgrad,
                                          ws + tl.dZ.offset, reinterpret_cast<uint8_t*>(G), n_img, tl.dZ.S,
                                          final_plane_S(0, tl.H, tl.W), st))
            return rc;
    }
    if (forked) {                                                    // join: the weight-gradient stream's partial sums
        CUDA_TRY(cudaEventRecord(m->train->ev_join, m->train->s_wgrad));
        CUDA_TRY(cudaStreamWaitEvent(st, m->train->ev_join, 0));
    }
    if (nl > 0) {                                                    // all weight gradients: partial sums -> gradient buffers
        long long max_total = 0;
        for (int i = 0; i < nl; ++i) max_total = std::max(max_total, static_cast<long long>(fin.L[i].cout) * fin.L[i].cin * 9);
        dim3 fgrid(static_cast<unsigned>(std::min<long long>((max_total + 255) / 256, 148)), nl);
        CUDA_TRY(launch_pdl(sedb::wgrad_finalize_kernel, dim3(fgrid), dim3(256), 0, st, fin));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int sedb_bce_with_logits(const float* logits_dev, const float* target_dev, long long B, long long F_out, long long F_tgt,
                         int K, float pos_weight, float grad_scale, float* loss_dev, float* dlogits_dev, void* stream) {
    if (!logits_dev || !target_dev || (!loss_dev && !dlogits_dev)) return fail("sedb_bce_with_logits: null argument");
    if (B < 1 || F_out < 1 || F_tgt < 1 || K < 1 || B * F_out * K > (1LL << 30)) return fail("sedb_bce_with_logits: bad shape");
    CUDA_TRY(launch_pdl(sedb::bce_logits_kernel, dim3(1), dim3(1024), 0, static_cast<cudaStream_t>(stream), 
        logits_dev, target_dev, static_cast<int>(B), static_cast<int>(F_out), static_cast<int>(F_tgt), K, pos_weight,
        grad_scale, loss_dev, dlogits_dev));
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int sedb_adam_amsgrad_step_dev(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                               float* max_exp_avg_sq_dev, long long n, float* state_dev, float* hyper_dev, float beta1,
                               float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
    if (n < 0) return fail("sedb_adam_amsgrad_step_dev: bad size");
    if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || !max_exp_avg_sq_dev || !state_dev || !hyper_dev)
        return fail("null buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(launch_pdl(sedb::adam_prepare_kernel, dim3(1), dim3(1), 0, st, state_dev, beta1, beta2, hyper_dev));
    g_launches.fetch_add(1);
    if (n > 0) {
        long long blocks = (n + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        CUDA_TRY(launch_pdl(sedb::adam_amsgrad_dev_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, st, param_dev, grad_dev, exp_avg_dev,
                                                                               exp_avg_sq_dev, max_exp_avg_sq_dev, n,
                                                                               hyper_dev, beta1, beta2, eps, weight_decay,
                                                                               grad_scale));
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
