// Host side of the CNN paths (included by sedb.cu): handles, geometry planning, parameter packing, launches.

namespace {

struct PlaneGeom {          // one blocked activation buffer
    int C = 0, H = 0, W = 0;
    int S = 0;              // pixels per 8-channel plane (incl. lead/tail slack)
    int elt = 2;            // bytes per element: 2 = fp16 planes, 4 = bf16 hi + lo plane sets or fp32 planes
    size_t offset = 0;      // byte offset inside the workspace
    size_t bytes_per_img() const { return static_cast<size_t>(C / 8) * S * 8 * elt; }
};

struct UmmaLayer {          // a conv layer executed by conv_umma_kernel
    int cin = 0, cout = 0, pool = 1, mode = 0, ntaps = 9;
    uint8_t* wpack = nullptr;
    float* scale = nullptr;
    float* shift = nullptr;
    int cin_chunk = 0, cout_tile = 0;
    int nrep = 1;             // replicas of the packed weights (ConvParams::w_nrep)
    size_t pack_bytes() const { return static_cast<size_t>(cout) * cin * ntaps * 4; }
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }
int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}
const int kFuseMaxCout = env_int("SEDB_FUSE_MAX_COUT", 64);
const size_t kWeightReplicas = static_cast<size_t>(env_int("SEDB_WEIGHT_REPLICAS", 8));

size_t conv_smem_bytes(const sedb::ConvParams& p) {
    const size_t patch = (static_cast<size_t>(p.patch_bytes) + 127) / 128 * 128;
    return patch + static_cast<size_t>(p.n_wslots) * p.wslot_bytes + p.stage_bytes +
           static_cast<size_t>(2) * p.cout * 4 + 16 + 32 * 8 + 16 + 128;
}

// Weight-ring geometry and pooling stage next to the patch: prefer several K-steps per slot (fewer barrier round
// trips for the MMA issuer), then as many slots as fit.
bool conv_fit_smem(sedb::ConvParams& p) {
    const int wblock = p.cout_tile * 64;
    const bool staged = (p.pool == 2 && p.mode == 0);             // 2x2 pooling exchanges pair sums through shared memory
    for (int cstep = 64; cstep >= 16; cstep /= 2) {
        if (staged && p.cout_sub % cstep) continue;
        p.cstep = cstep;
        p.stage_bytes = staged ? 64 * p.n_tiles * (cstep + 1) * 4 : 0;
        for (int kpb = p.ntaps; kpb >= 1; --kpb) {                 // taps per slot: all of them, 3 or 1 (conv_issue.cuh)
            if (p.ntaps % kpb || (kpb != p.ntaps && kpb != 3 && kpb != 1) || kpb * wblock > sedb::kConvMaxWSlotBytes)
                continue;
            p.kpb = kpb;
            p.wslot_bytes = kpb * wblock;
            for (int slots = sedb::kConvMaxWSlots; slots >= (kpb == p.ntaps ? 2 : 3); --slots) {
                p.n_wslots = slots;
                if (conv_smem_bytes(p) <= 227 * 1024) return true;
            }
        }
        if (!staged) break;
    }
    p.cstep = 16;
    p.stage_bytes = staged ? 64 * p.n_tiles * 17 * 4 : 0;
    p.kpb = 1;
    p.wslot_bytes = wblock;
    p.n_wslots = 2;
    return conv_smem_bytes(p) <= 227 * 1024;
}

// Geometry of one candidate decomposition (band of <= max_tiles M tiles, N tile split into n_nsub work items);
// returns false when it does not fit.  S_in receives the plane size the *input* buffer must have.
bool plan_candidate(const UmmaLayer& L, int H, int W, int amode, int max_tiles, int n_nsub, int fuse, sedb::ConvParams& p,
                    int& S_in) {
    p = sedb::ConvParams{};
    p.mode = L.mode;
    p.H = H;
    p.W = W;
    p.Wp = W + 2;
    p.cin = L.cin;
    p.cout = L.cout;
    p.cin_chunk = L.cin_chunk;
    p.n_kchunks = L.cin / L.cin_chunk;
    p.cout_tile = L.cout_tile;
    p.n_ntiles = L.cout / L.cout_tile;
    p.pool = L.pool;
    p.ntaps = L.ntaps;
    p.n_nsub = n_nsub;
    p.cout_sub = L.cout_tile / n_nsub;
    if (p.cout_sub % 16 || p.cout_sub < 16) return false;
    // [wH | wL] as ONE operand (N = 2 cout_tile <= 128): a property of the layer, never of the batch size, so that a
    // clip's result does not depend on the decomposition chosen for the batch it arrives in (the fused form adds the two
    // partial sums in the epilogue, the unfused form in the accumulator: different roundings)
    if (fuse != (L.cout_tile <= kFuseMaxCout ? 1 : 0)) return false;
    if (fuse && n_nsub != 1) return false;
    p.fuse = fuse;
    const int tile_cols = p.fuse ? 2 * p.cout_tile : p.cout_sub;
    if (max_tiles * tile_cols > 256) max_tiles = 256 / tile_cols;       // accumulators are double buffered in TMEM
    if (max_tiles < 1) return false;
    if (L.mode == 0) {
        p.halo = p.Wp + 1;
        int R = (128 * max_tiles) / p.Wp;
        if (L.pool == 2) R &= ~1;
        const int Hcap = (L.pool == 2) ? round_up(H, 2) : H;
        if (R > Hcap) R = Hcap;
        if (R < L.pool) return false;
        p.R = R;
        p.n_tiles = (R * p.Wp + 127) / 128;
        p.n_bands = (H + R - 1) / R;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) p.tapoff[kh * 3 + kw] = p.halo + (kh - 1) * p.Wp + (kw - 1);
        p.Ho = H / L.pool;
        p.Wo = W / L.pool;
        p.Wpo = p.Wo + 2;
    } else {
        p.halo = 1;
        p.n_tiles = (W + 127) / 128;
        if (p.n_tiles > max_tiles) p.n_tiles = max_tiles;
        p.R = 0;
        p.n_bands = (W + 128 * p.n_tiles - 1) / (128 * p.n_tiles);
        for (int k = 0; k < 3; ++k) p.tapoff[k] = k;
        p.Ho = 1;
        p.Wo = W / L.pool;
        p.Wpo = p.Wo + 2;
    }
    p.P = 128 * p.n_tiles + 2 * p.halo;
    p.patch_bytes = (1 + amode) * (L.cin_chunk / 8) * p.P * 16;
    if (!conv_fit_smem(p)) return false;
    const int v0_last = (L.mode == 0) ? (p.R * (p.n_bands - 1) + 1) * p.Wp + 1 : 1 + (p.n_bands - 1) * 128 * p.n_tiles;
    S_in = round_up(sedb::kConvLead + v0_last - p.halo + p.P, 8);
    return true;
}

// Measured tcgen05.mma cost (cycles, 128 x N x 16, no-swizzle operands from shared memory; profiles/r1_phase_cycles.md)
double mma_cycles(int N) { return N <= 32 ? 40.6 : N <= 64 ? 48.6 : N <= 128 ? 64.7 : 129.4; }

// Estimated kernel time (cycles) of a candidate: waves of work items over the SMs, each item bound by its MMAs or by the
// L2 -> shared-memory stream of its patch and weights, plus a fixed pipeline fill/drain per item.
double candidate_cost(const sedb::ConvParams& p, int amode, long long n_img, int num_sms) {
    const double k_steps = static_cast<double>(p.cin / 16) * p.ntaps;
    double per_tile;
    if (p.fuse) per_tile = mma_cycles(2 * p.cout_tile) + (amode ? mma_cycles(p.cout_sub) : 0.0);
    else per_tile = (amode ? 3.0 : 2.0) * mma_cycles(p.cout_sub);
    const double mma = p.n_tiles * k_steps * per_tile;
    const double bytes = static_cast<double>(p.patch_bytes) * p.n_kchunks + k_steps * p.cout_tile * 64.0;
    const long long items = n_img * p.n_bands * p.n_ntiles * p.n_nsub;
    const long long waves = (items + num_sms - 1) / num_sms;
    // L2 -> SM: ~6300 B/clk for the whole chip (B300_MICROARCH.md, LTS cap), ~64 B/clk for one SM's bulk copies
    const double active = static_cast<double>(std::min<long long>(items, num_sms));
    const double stream = bytes / std::min(64.0, 6300.0 / active);
    const double epi = p.n_tiles * (p.cout_sub / 16.0) * (p.pool != 1 ? 350.0 : 220.0);
    const double item = std::max(std::max(mma, stream), epi) + 2500.0;
    return static_cast<double>(waves) * item + 6000.0;                   // + launch/prologue
}

// Picks the decomposition of one tensor-core conv layer for an input of H x W (2-D) or length W (1-D) over n_img
// images and returns the plane size S the *input* buffer must have (-1: unsupported).
int plan_umma_layer(const UmmaLayer& L, int H, int W, int amode, long long n_img, int num_sms, sedb::ConvParams& best) {
    double best_cost = 0.0;
    int best_S = -1;
    for (int cand = 0; cand < 6; ++cand) {
        const int n_nsub = 1 << (cand >> 1), fuse = cand & 1;
        int last_tiles = -1;
        for (int max_tiles = sedb::kConvMaxTiles; max_tiles >= 1; --max_tiles) {
            sedb::ConvParams p;
            int S = 0;
            if (!plan_candidate(L, H, W, amode, max_tiles, n_nsub, fuse, p, S)) continue;
            if (p.n_tiles == last_tiles) continue;
            last_tiles = p.n_tiles;
            const double c = candidate_cost(p, amode, n_img, num_sms);
            if (best_S < 0 || c < best_cost * 0.999) {
                best_cost = c;
                best_S = S;
                best = p;
            }
        }
    }
    return best_S;
}

int final_plane_S(int mode, int H, int W) {
    return round_up(sedb::kConvLead + (mode == 0 ? (H + 2) * (W + 2) : W + 2) + 8, 8);
}

int alloc_layer_params(UmmaLayer& L) {
    L.cin_chunk = L.cin > 128 ? 128 : L.cin;
    L.cout_tile = L.cout > 128 ? 128 : L.cout;
    if (L.cin % 16 || L.cin % L.cin_chunk || L.cout % 16 || L.cout % L.cout_tile)
        return fail("conv layer %d->%d: channel counts must be multiples of 16 (and of 128 above 128)", L.cin, L.cout);
    L.nrep = static_cast<int>(std::max<size_t>(1, std::min<size_t>(kWeightReplicas, (8u << 20) / L.pack_bytes())));
    CUDA_TRY(cudaMalloc(&L.wpack, L.pack_bytes() * L.nrep));
    CUDA_TRY(cudaMalloc(&L.scale, L.cout * sizeof(float)));
    CUDA_TRY(cudaMalloc(&L.shift, L.cout * sizeof(float)));
    return 0;
}
void free_layer_params(UmmaLayer& L) {
    cudaFree(L.wpack);
    cudaFree(L.scale);
    cudaFree(L.shift);
    L.wpack = nullptr;
    L.scale = L.shift = nullptr;
}

// Launch with programmatic stream serialization: the kernel may be placed while its predecessor in the stream is still
// running (after the predecessor's `griddepcontrol.launch_dependents`); it must execute `griddepcontrol.wait` before it
// touches global memory.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int AMODE>
int launch_umma_layer(const sedb_ctx* c, const uint8_t* wpack, int w_nrep, size_t w_rep_bytes, const float* scale,
                      const float* shift, sedb::ConvParams p, const uint8_t* in, uint8_t* out, int n_img, int S_in,
                      int S_out, cudaStream_t st) {
    p.in = in;
    p.out = out;
    p.wpack = wpack;
    p.w_nrep = w_nrep;
    p.w_rep_bytes = static_cast<long long>(w_rep_bytes);
    p.scale = scale;
    p.shift = shift;
    p.n_img = n_img;
    p.S_in = S_in;
    p.S_out = S_out;
    p.prof = g_prof ? g_prof + 16 * (1 + (g_conv_layer++ % 7)) : nullptr;
    const long long items = static_cast<long long>(n_img) * p.n_bands * p.n_ntiles * p.n_nsub;
    if (items <= 0) return 0;
    const int grid = static_cast<int>(items < c->num_sms ? items : c->num_sms);
    const size_t smem = conv_smem_bytes(p);
    if (smem > 227 * 1024) return fail("conv layer needs %zu bytes of shared memory", smem);
    if (SEDB_CONV_PDL) {
        // Programmatic dependent launch between the layers (inference and training): the kernel's set-up (barriers, TMEM allocation,
        // folded-BN constants) runs on every SM the previous layer has already left; `griddepcontrol.wait` in the kernel
        // holds all global-memory traffic until the previous grid has completed and flushed.
        CUDA_TRY(launch_pdl(sedb::conv_umma_kernel<AMODE>, dim3(static_cast<unsigned>(grid)), dim3(sedb::kConvThreads), smem, st, p));
    } else {
        sedb::conv_umma_kernel<AMODE><<<grid, sedb::kConvThreads, smem, st>>>(p);
    }
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int fold_and_pack(UmmaLayer& L, const float* w, const float* bias, const float* gamma, const float* beta,
                  const float* mean, const float* var, cudaStream_t st) {
    const long long total = static_cast<long long>(L.cout) * L.cin * L.ntaps;
    sedb::pack_conv_weight_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(w, L.wpack, L.cout, L.cin,
                                                                                      L.ntaps, L.cout_tile, L.cin_chunk, 1, 0, L.nrep);
    sedb::bn_fold_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(gamma, beta, mean, var, bias, 1e-5f, L.cout, L.scale,
                                                              L.shift);
    g_launches.fetch_add(2);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Host-side record of which workspaces hold zeroed padding for which geometry.  Nothing in-band is trusted: a caller
// that lets anything else write a workspace between calls (or frees and re-allocates it) must invalidate it.
struct ZeroedSet {
    struct Entry { const void* ptr; unsigned long long tag; };
    std::vector<Entry> entries;
    bool has(const void* ptr, unsigned long long tag) const {
        for (const auto& e : entries)
            if (e.ptr == ptr && e.tag == tag) return true;
        return false;
    }
    void put(const void* ptr, unsigned long long tag) {
        for (auto& e : entries)
            if (e.ptr == ptr) { e.tag = tag; return; }
        if (entries.size() >= 8) entries.erase(entries.begin());
        entries.push_back({ptr, tag});
    }
    void drop(const void* ptr) {
        if (!ptr) { entries.clear(); return; }
        for (size_t i = 0; i < entries.size(); ++i)
            if (entries[i].ptr == ptr) { entries.erase(entries.begin() + i); return; }
    }
};

int prepare_workspace(ZeroedSet& z, void* ws, size_t bytes, unsigned long long tag, cudaStream_t st) {
    if (z.has(ws, tag)) return 0;
    sedb::ws_zero_kernel<<<592, 256, 0, st>>>(reinterpret_cast<uint4*>(ws), static_cast<long long>(bytes / 16));
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    z.put(ws, tag);
    return 0;
}

unsigned long long mix_tag(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h;
}

}  // namespace

// ============================================================================================ Cnn_AvgPooling
struct CnnPlan {
    std::vector<PlaneGeom> planes;            // planes[0] = output of block0.conv1, planes[i+1] = output of layers[i]
    std::vector<sedb::ConvParams> params;     // per umma layer
    size_t ws_bytes = 0;
    int Hf = 0, Wf = 0;
    unsigned long long tag = 0;
};

struct sedb_cnn {
    sedb_ctx* ctx = nullptr;
    int n_blocks = 0, classes = 0;
    std::vector<int> channels, pools;
    int ratio = 1;                    // 2 ** num_pools (spectogram_models.py:167-172,200)
    // block 0 conv1 (C_in = 1)
    float* w_in = nullptr;            // [C0][9]
    float* scale_in = nullptr;
    float* shift_in = nullptr;
    std::vector<UmmaLayer> layers;    // block0.conv2, block1.conv1, block1.conv2, ...
    float* fc_w = nullptr;            // [classes][C_last]
    float* fc_b = nullptr;
    bool loaded = false;
    std::map<std::pair<long long, long long>, CnnPlan> plans;   // (n_clips, T) -> plan, built once per shape
    ZeroedSet zeroed;
    struct sedb_cnn_train* train = nullptr;                       // training-step state (cnn_train_host.inl), lazily built
};

static int cnn_make_plan(const sedb_cnn* m, long long n_clips, long long T, CnnPlan& plan) {
    const int nl = static_cast<int>(m->layers.size());
    plan.planes.assign(nl + 1, PlaneGeom{});
    plan.params.assign(nl, sedb::ConvParams{});
    int H = static_cast<int>(T), W = SEDB_MEL_BINS;
    plan.planes[0].C = m->channels[0];
    plan.planes[0].H = H;
    plan.planes[0].W = W;
    for (int i = 0; i < nl; ++i) {
        const UmmaLayer& L = m->layers[i];
        const int S_in = plan_umma_layer(L, H, W, 0, n_clips, m->ctx->num_sms, plan.params[i]);
        if (S_in < 0) return fail("unsupported feature-map size %d x %d", H, W);
        plan.planes[i].S = S_in;
        H = plan.params[i].Ho;
        W = plan.params[i].Wo;
        if (L.pool == 1) { H = plan.params[i].H; W = plan.params[i].W; }
        if (H < 1 || W < 1) return fail("input of %lld frames is too short for this model's pooling", T);
        plan.planes[i + 1].C = L.cout;
        plan.planes[i + 1].H = H;
        plan.planes[i + 1].W = W;
    }
    plan.planes[nl].S = final_plane_S(0, H, W);
    plan.Hf = H;
    plan.Wf = W;
    size_t off = 0;
    unsigned long long tag = mix_tag(0x5EDBull, static_cast<unsigned long long>(n_clips));
    tag = mix_tag(tag, static_cast<unsigned long long>(T));
    for (auto& g : plan.planes) {
        g.offset = off;
        off += (g.bytes_per_img() * static_cast<size_t>(n_clips) + 127) / 128 * 128;
        tag = mix_tag(tag, (static_cast<unsigned long long>(g.C) << 40) ^ (static_cast<unsigned long long>(g.S) << 8) ^ g.W);
    }
    plan.ws_bytes = off;
    plan.tag = tag | 1ull;
    return 0;
}

// plan for (n_clips, T), built on first use and kept in the handle (a forward call then costs the launches only)
static int cnn_get_plan(sedb_cnn* m, long long n_clips, long long T, const CnnPlan** out) {
    const auto key = std::make_pair(n_clips, T);
    auto it = m->plans.find(key);
    if (it == m->plans.end()) {
        CnnPlan plan;
        if (int rc = cnn_make_plan(m, n_clips, T, plan)) return rc;
        if (m->plans.size() >= 64) m->plans.clear();
        it = m->plans.emplace(key, std::move(plan)).first;
    }
    *out = &it->second;
    return 0;
}

static int sedb_cnn_kernels_init() {
    CUDA_TRY(cudaFuncSetAttribute(sedb::conv_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(sedb::conv_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(sedb::m5_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sedb::kFrontSmem));
    return 0;
}

static void sedb_cnn_train_free(sedb_cnn* m);
static void sedb_cnn_train_invalidate(sedb_cnn* m, const void* ws);

extern "C" {

int sedb_cnn_destroy(sedb_cnn_t* m);

int sedb_cnn_create(sedb_ctx_t* ctx, const int* channels, const int* pools, int n_blocks, int classes_num,
                    sedb_cnn_t** out) {
    if (!ctx || !channels || !pools || !out) return fail("sedb_cnn_create: null argument");
    *out = nullptr;
    if (n_blocks < 1 || n_blocks > 8 || classes_num < 1) return fail("sedb_cnn_create: bad model_config");
    for (int i = 0; i < n_blocks; ++i) {
        if (channels[i] % 16 || channels[i] < 16 || channels[i] > 1024)
            return fail("sedb_cnn_create: channels[%d]=%d must be a multiple of 16 in [16,1024]", i, channels[i]);
        if (channels[i] > 128 && channels[i] % 128)
            return fail("sedb_cnn_create: channels[%d]=%d above 128 must be a multiple of 128", i, channels[i]);
        if (pools[i] != 1 && pools[i] != 2) return fail("sedb_cnn_create: pool size must be 1 or 2");
    }
    sedb_cnn* m = new (std::nothrow) sedb_cnn();
    if (!m) return fail("out of host memory");
    // every failure below releases what has been allocated so far (the caller never sees a partial object)
    struct Guard {
        sedb_cnn* m;
        ~Guard() { if (m) sedb_cnn_destroy(m); }
    } guard{m};
    m->ctx = ctx;
    m->n_blocks = n_blocks;
    m->classes = classes_num;
    m->channels.assign(channels, channels + n_blocks);
    m->pools.assign(pools, pools + n_blocks);
    // spectogram_models.py:167-172: num_pools starts at 1 and counts pool==2 among blocks 1..n-1
    int num_pools = 1;
    for (int i = 1; i < n_blocks; ++i)
        if (pools[i] == 2) ++num_pools;
    m->ratio = 1 << num_pools;
    const int C0 = channels[0];
    CUDA_TRY(cudaMalloc(&m->w_in, C0 * 9 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->scale_in, C0 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->shift_in, C0 * sizeof(float)));
    for (int b = 0; b < n_blocks; ++b) {
        if (b > 0) {
            m->layers.emplace_back();
            UmmaLayer& L = m->layers.back();
            L.cin = channels[b - 1];
            L.cout = channels[b];
            L.pool = 1;
            if (int rc = alloc_layer_params(L)) return rc;
        }
        m->layers.emplace_back();
        UmmaLayer& L = m->layers.back();
        L.cin = channels[b];
        L.cout = channels[b];
        L.pool = pools[b];
        if (int rc = alloc_layer_params(L)) return rc;
    }
    CUDA_TRY(cudaMalloc(&m->fc_w, static_cast<size_t>(classes_num) * channels[n_blocks - 1] * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->fc_b, classes_num * sizeof(float)));
    guard.m = nullptr;
    *out = m;
    return 0;
}

int sedb_cnn_destroy(sedb_cnn_t* m) {
    if (!m) return 0;
    sedb_cnn_train_free(m);
    cudaFree(m->w_in);
    cudaFree(m->scale_in);
    cudaFree(m->shift_in);
    for (auto& L : m->layers) free_layer_params(L);
    cudaFree(m->fc_w);
    cudaFree(m->fc_b);
    delete m;
    return 0;
}

int sedb_cnn_load(sedb_cnn_t* m, const float* const* t, int n_tensors, void* stream) {
    if (!m || !t) return fail("sedb_cnn_load: null argument");
    if (n_tensors != 10 * m->n_blocks + 2)
        return fail("sedb_cnn_load: expected %d tensors (10 per block + event_fc weight/bias), got %d",
                    10 * m->n_blocks + 2, n_tensors);
    for (int i = 0; i < n_tensors; ++i)
        if (!t[i]) return fail("sedb_cnn_load: tensor %d is null", i);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int li = 0;
    for (int b = 0; b < m->n_blocks; ++b) {
        const float* const* q = t + 10 * b;   // conv1.w conv2.w bn1.{w,b,rm,rv} bn2.{w,b,rm,rv}
        if (b == 0) {
            CUDA_TRY(cudaMemcpyAsync(m->w_in, q[0], m->channels[0] * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
            sedb::bn_fold_kernel<<<(m->channels[0] + 127) / 128, 128, 0, st>>>(q[2], q[3], q[4], q[5], nullptr, 1e-5f,
                                                                             m->channels[0], m->scale_in, m->shift_in);
            g_launches.fetch_add(1);
        } else {
            if (int rc = fold_and_pack(m->layers[li++], q[0], nullptr, q[2], q[3], q[4], q[5], st)) return rc;
        }
        if (int rc = fold_and_pack(m->layers[li++], q[1], nullptr, q[6], q[7], q[8], q[9], st)) return rc;
    }
    const int Cl = m->channels[m->n_blocks - 1];
    CUDA_TRY(cudaMemcpyAsync(m->fc_w, t[10 * m->n_blocks], static_cast<size_t>(m->classes) * Cl * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(m->fc_b, t[10 * m->n_blocks + 1], m->classes * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaGetLastError());
    m->loaded = true;
    return 0;
}

long long sedb_cnn_out_frames(const sedb_cnn_t* m, long long T) {
    if (!m) return -1;
    long long H = T;
    for (int b = 0; b < m->n_blocks; ++b) H /= m->pools[b];
    return H * m->ratio;
}

size_t sedb_cnn_workspace_bytes(const sedb_cnn_t* m, long long n_clips, long long T) {
    if (!m || n_clips <= 0 || T <= 0) return 0;
    const CnnPlan* plan = nullptr;
    if (cnn_get_plan(const_cast<sedb_cnn_t*>(m), n_clips, T, &plan)) return 0;
    return plan->ws_bytes;
}

int sedb_cnn_workspace_invalidate(sedb_cnn_t* m, const void* workspace_dev) {
    if (!m) return fail("sedb_cnn_workspace_invalidate: null handle");
    m->zeroed.drop(workspace_dev);
    sedb_cnn_train_invalidate(m, workspace_dev);
    return 0;
}

int sedb_cnn_forward(sedb_cnn_t* m, const float* x_dev, long long n_clips, long long T, float* logits_dev,
                     float* probs_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!m || !x_dev || !workspace_dev) return fail("sedb_cnn_forward: null argument");
    if (!logits_dev && !probs_dev) return fail("sedb_cnn_forward: no output requested");
    if (!m->loaded) return fail("sedb_cnn_forward: parameters not loaded (call sedb_cnn_load)");
    if (n_clips < 0 || T < 1 || n_clips > (1 << 20) || T > (1 << 20)) return fail("sedb_cnn_forward: bad shape");
    if (n_clips == 0) return 0;
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 127) return fail("workspace must be 128-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CnnPlan* planp = nullptr;
    if (int rc = cnn_get_plan(m, n_clips, T, &planp)) return rc;
    const CnnPlan& plan = *planp;
    if (workspace_bytes < plan.ws_bytes)
        return fail("sedb_cnn_forward: workspace has %zu bytes, needs %zu", workspace_bytes, plan.ws_bytes);
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    if (int rc = prepare_workspace(m->zeroed, ws, plan.ws_bytes, plan.tag, st)) return rc;
    const int n_img = static_cast<int>(n_clips);
    {   // block 0 conv1
        const PlaneGeom& g = plan.planes[0];
        const bool px4 = true;                      // four rows per thread
        const long long total = static_cast<long long>(n_img) * (px4 ? (g.H + 3) / 4 : g.H) * g.W;
        long long blocks = (total + 255) / 256;
        if (!px4 && blocks > 148LL * 16) blocks = 148LL * 16;     // (px4: one item per thread, no ragged second pass)
        const size_t smem = static_cast<size_t>(g.C) * 11 * sizeof(float);
        if (px4)
            sedb::conv_in2d_px4_kernel<0><<<static_cast<int>(blocks), 256, smem, st>>>(
                x_dev, m->w_in, m->scale_in, m->shift_in, ws + g.offset, n_img, g.H, g.W, g.C, g.S);
        else
            sedb::conv_in2d_kernel<0><<<static_cast<int>(blocks), 256, smem, st>>>(
                x_dev, m->w_in, m->scale_in, m->shift_in, ws + g.offset, n_img, g.H, g.W, g.C, g.S);
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    for (size_t i = 0; i < m->layers.size(); ++i) {
        const UmmaLayer& L = m->layers[i];
        if (int rc = launch_umma_layer<0>(m->ctx, L.wpack, L.nrep, L.pack_bytes(), L.scale, L.shift, plan.params[i], ws + plan.planes[i].offset,
                                          ws + plan.planes[i + 1].offset, n_img, plan.planes[i].S, plan.planes[i + 1].S, st))
            return rc;
    }
    {
        const PlaneGeom& g = plan.planes.back();
        const long long warps = static_cast<long long>(n_img) * plan.Hf;
        const int blocks = static_cast<int>((warps * 32 + 255) / 256);
        CUDA_TRY(launch_pdl(sedb::head2d_kernel<0>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st,
                            static_cast<const uint8_t*>(ws + g.offset), static_cast<const float*>(m->fc_w),
                            static_cast<const float*>(m->fc_b), logits_dev, probs_dev, n_img, g.C, plan.Hf, plan.Wf, g.S,
                            m->classes, m->ratio));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"

// CNN stage of the host pipelines: clips [clip0, clip0 + n_group) of the log-mel image resident in ctx->d_out ->
// probabilities in ctx->d_probs (sized for total_clips).  Called per group of clips while later waveform chunks are
// still in flight, so that only the last group's forward pass is exposed after the last host->device copy.
static int sedb_cnn_forward_group(sedb_ctx_t* c, sedb_cnn_t* cnn, long long clip0, long long n_group, long long total_clips,
                                  long long T, int slot) {
    const size_t need = sedb_cnn_workspace_bytes(cnn, n_group, T);
    if (need == 0) return fail("cannot plan the CNN for %lld clips x %lld frames", n_group, T);
    if (need > c->d_ws_bytes[slot]) {
        CUDA_TRY(cudaStreamSynchronize(c->s_comp));        // an earlier group may still be using the old workspace
        cnn->zeroed.drop(c->d_ws[slot]);
        cudaFree(c->d_ws[slot]);
        c->d_ws[slot] = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_ws[slot], need));
        c->d_ws_bytes[slot] = need;
    }
    const size_t per_clip = static_cast<size_t>(sedb_cnn_out_frames(cnn, T)) * cnn->classes;
    const size_t n_out = static_cast<size_t>(total_clips) * per_clip;
    if (n_out > c->d_probs_elems) {
        CUDA_TRY(cudaStreamSynchronize(c->s_comp));
        cudaFree(c->d_probs);
        c->d_probs = nullptr;
        CUDA_TRY(cudaMalloc(&c->d_probs, n_out * sizeof(float)));
        c->d_probs_elems = n_out;
    }
    return sedb_cnn_forward(cnn, c->d_out + clip0 * T * SEDB_MEL_BINS, n_group, T, nullptr, c->d_probs + clip0 * per_clip,
                            c->d_ws[slot], c->d_ws_bytes[slot], c->s_comp);
}
static int sedb_cnn_results_to_host(sedb_ctx_t* c, sedb_cnn_t* cnn, long long n_clips, long long T, float* probs_host) {
    const size_t n_out = static_cast<size_t>(n_clips) * sedb_cnn_out_frames(cnn, T) * cnn->classes;
    CUDA_TRY(cudaMemcpyAsync(probs_host, c->d_probs, n_out * sizeof(float), cudaMemcpyDeviceToHost, c->s_comp));
    return 0;
}

// ============================================================================================ M5
struct M5Plan {
    std::vector<PlaneGeom> planes;
    std::vector<sedb::ConvParams> params;
    size_t ws_bytes = 0;
    int Lf = 0;
    unsigned long long tag = 0;
};

struct sedb_m5 {
    sedb_ctx* ctx = nullptr;
    int classes = 0;
    float* w_in = nullptr;            // conv_block1 conv: [64][79]
    uint8_t* w_front = nullptr;       // the same, packed bf16 hi|lo for m5_front_kernel
    float* scale_in = nullptr;
    float* shift_in = nullptr;
    std::vector<UmmaLayer> layers;    // the 8 k=3 convolutions
    float* fc_w = nullptr;            // [classes][256]
    float* fc_b = nullptr;
    bool loaded = false;
    std::map<long long, M5Plan> plans;
    ZeroedSet zeroed;
};

static const int kM5Cin[8] = {64, 64, 64, 64, 64, 128, 128, 256};
static const int kM5Cout[8] = {64, 64, 64, 64, 128, 128, 256, 256};
static const int kM5Pool[8] = {1, 4, 1, 4, 1, 4, 1, 1};          // waveform_models.py:22-56
static const int kM5FrameLen = SEDB_FRAME_SIZE;                  // M5 input length (waveform_configs.frame_size)

static int m5_make_plan(const sedb_m5* m, long long n_frames, M5Plan& plan) {
    plan.planes.assign(9, PlaneGeom{});
    plan.params.assign(8, sedb::ConvParams{});
    int L = ((kM5FrameLen + 2 * 39 - 79) / 4 + 1) / 4;            // conv k79 s4 p39 -> 7920, MaxPool(4) -> 1980
    plan.planes[0].C = 64;
    plan.planes[0].H = 1;
    plan.planes[0].W = L;
    for (int i = 0; i < 8; ++i) {
        const int S_in = plan_umma_layer(m->layers[i], 1, L, 0, n_frames, m->ctx->num_sms, plan.params[i]);
        if (S_in < 0) return fail("m5 planning failed");
        plan.planes[i].S = S_in;
        if (m->layers[i].pool != 1) L = plan.params[i].Wo;
        plan.planes[i + 1].C = m->layers[i].cout;
        plan.planes[i + 1].H = 1;
        plan.planes[i + 1].W = L;
    }
    plan.planes[8].S = final_plane_S(1, 1, L);
    plan.Lf = L;
    size_t off = 0;
    unsigned long long tag = mix_tag(0x35ull, static_cast<unsigned long long>(n_frames));
    for (auto& g : plan.planes) {
        g.offset = off;
        off += (g.bytes_per_img() * static_cast<size_t>(n_frames) + 127) / 128 * 128;
        tag = mix_tag(tag, (static_cast<unsigned long long>(g.C) << 40) ^ (static_cast<unsigned long long>(g.S) << 8) ^ g.W);
    }
    plan.ws_bytes = off;
    plan.tag = tag | 1ull;
    return 0;
}

static int m5_get_plan(sedb_m5* m, long long n_frames, const M5Plan** out) {
    auto it = m->plans.find(n_frames);
    if (it == m->plans.end()) {
        M5Plan plan;
        if (int rc = m5_make_plan(m, n_frames, plan)) return rc;
        if (m->plans.size() >= 64) m->plans.clear();
        it = m->plans.emplace(n_frames, std::move(plan)).first;
    }
    *out = &it->second;
    return 0;
}

extern "C" {

int sedb_m5_destroy(sedb_m5_t* m);

int sedb_m5_create(sedb_ctx_t* ctx, int classes_num, sedb_m5_t** out) {
    if (!ctx || !out) return fail("sedb_m5_create: null argument");
    *out = nullptr;
    if (classes_num < 1) return fail("sedb_m5_create: classes_num must be positive");
    sedb_m5* m = new (std::nothrow) sedb_m5();
    if (!m) return fail("out of host memory");
    struct Guard {
        sedb_m5* m;
        ~Guard() { if (m) sedb_m5_destroy(m); }
    } guard{m};
    m->ctx = ctx;
    m->classes = classes_num;
    CUDA_TRY(cudaMalloc(&m->w_in, 64 * 80 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->w_front, sedb::kFrontWBytes));
    CUDA_TRY(cudaMalloc(&m->scale_in, 64 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->shift_in, 64 * sizeof(float)));
    for (int i = 0; i < 8; ++i) {
        m->layers.emplace_back();
        UmmaLayer& L = m->layers.back();
        L.cin = kM5Cin[i];
        L.cout = kM5Cout[i];
        L.pool = kM5Pool[i];
        L.mode = 1;
        L.ntaps = 3;
        if (int rc = alloc_layer_params(L)) return rc;
    }
    CUDA_TRY(cudaMalloc(&m->fc_w, static_cast<size_t>(classes_num) * 256 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&m->fc_b, classes_num * sizeof(float)));
    guard.m = nullptr;
    *out = m;
    return 0;
}

int sedb_m5_destroy(sedb_m5_t* m) {
    if (!m) return 0;
    cudaFree(m->w_in);
    cudaFree(m->w_front);
    cudaFree(m->scale_in);
    cudaFree(m->shift_in);
    for (auto& L : m->layers) free_layer_params(L);
    cudaFree(m->fc_w);
    cudaFree(m->fc_b);
    delete m;
    return 0;
}

int sedb_m5_load(sedb_m5_t* m, const float* const* t, int n_tensors, void* stream) {
    if (!m || !t) return fail("sedb_m5_load: null argument");
    if (n_tensors != 9 * 6 + 2) return fail("sedb_m5_load: expected 56 tensors (6 per conv+bn pair, fc weight/bias), got %d", n_tensors);
    for (int i = 0; i < n_tensors; ++i)
        if (!t[i]) return fail("sedb_m5_load: tensor %d is null", i);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // pair 0: conv_block1 (k=79): conv.w conv.b bn.w bn.b bn.rm bn.rv
    CUDA_TRY(cudaMemcpyAsync(m->w_in, t[0], 64 * 79 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    sedb::bn_fold_kernel<<<1, 128, 0, st>>>(t[2], t[3], t[4], t[5], t[1], 1e-5f, 64, m->scale_in, m->shift_in);
    sedb::pack_front_weight_kernel<<<(64 * 80 + 255) / 256, 256, 0, st>>>(m->w_in, m->w_front);
    g_launches.fetch_add(2);
    for (int i = 0; i < 8; ++i) {
        const float* const* q = t + 6 * (i + 1);
        if (int rc = fold_and_pack(m->layers[i], q[0], q[1], q[2], q[3], q[4], q[5], st)) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(m->fc_w, t[54], static_cast<size_t>(m->classes) * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(m->fc_b, t[55], m->classes * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaGetLastError());
    m->loaded = true;
    return 0;
}

size_t sedb_m5_workspace_bytes(const sedb_m5_t* m, long long n_frames) {
    if (!m || n_frames <= 0) return 0;
    const M5Plan* plan = nullptr;
    if (m5_get_plan(const_cast<sedb_m5_t*>(m), n_frames, &plan)) return 0;
    return plan->ws_bytes;
}

int sedb_m5_workspace_invalidate(sedb_m5_t* m, const void* workspace_dev) {
    if (!m) return fail("sedb_m5_workspace_invalidate: null handle");
    m->zeroed.drop(workspace_dev);
    return 0;
}

int sedb_m5_forward(sedb_m5_t* m, const float* x_dev, long long n_frames, float* logits_dev, void* workspace_dev,
                    size_t workspace_bytes, void* stream) {
    if (!m || !x_dev || !logits_dev || !workspace_dev) return fail("sedb_m5_forward: null argument");
    if (!m->loaded) return fail("sedb_m5_forward: parameters not loaded (call sedb_m5_load)");
    if (n_frames < 0 || n_frames > (1 << 24)) return fail("sedb_m5_forward: bad frame count");
    if (n_frames == 0) return 0;
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 127) return fail("workspace must be 128-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const M5Plan* planp = nullptr;
    if (int rc = m5_get_plan(m, n_frames, &planp)) return rc;
    const M5Plan& plan = *planp;
    if (workspace_bytes < plan.ws_bytes)
        return fail("sedb_m5_forward: workspace has %zu bytes, needs %zu", workspace_bytes, plan.ws_bytes);
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    if (int rc = prepare_workspace(m->zeroed, ws, plan.ws_bytes, plan.tag, st)) return rc;
    const int n = static_cast<int>(n_frames);
    {
        const PlaneGeom& g = plan.planes[0];
        const int L_conv = (kM5FrameLen + 2 * 39 - 79) / 4 + 1;                 // 7920
        const long long items = static_cast<long long>(n) * ((L_conv + sedb::kFrontTilePos - 1) / sedb::kFrontTilePos);
        const int grid = static_cast<int>(items < m->ctx->num_sms ? items : m->ctx->num_sms);
        sedb::m5_front_kernel<<<grid, sedb::kFrontThreads, sedb::kFrontSmem, st>>>(
            x_dev, m->w_front, m->scale_in, m->shift_in, ws + g.offset, n, kM5FrameLen, L_conv, g.W, g.S);
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    for (int i = 0; i < 8; ++i) {
        const UmmaLayer& L = m->layers[i];
        if (int rc = launch_umma_layer<0>(m->ctx, L.wpack, L.nrep, L.pack_bytes(), L.scale, L.shift, plan.params[i], ws + plan.planes[i].offset,
                                          ws + plan.planes[i + 1].offset, n, plan.planes[i].S, plan.planes[i + 1].S, st))
            return rc;
    }
    {
        const PlaneGeom& g = plan.planes[8];
        const int blocks = (n * 32 + 255) / 256;
        CUDA_TRY(launch_pdl(sedb::head1d_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st,
                            static_cast<const uint8_t*>(ws + g.offset), static_cast<const float*>(m->fc_w),
                            static_cast<const float*>(m->fc_b), logits_dev, n, g.C, plan.Lf, g.S, m->classes));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"
