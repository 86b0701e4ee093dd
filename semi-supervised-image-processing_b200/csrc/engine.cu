// C ABI of the library (include/fx_b200.h): handle lifetime, weight folding/packing, the
// forward schedule of the truncated ResNet-18, host-buffer entry point.
//
// Reference lines replaced: load_model (src/feature_extraction.py:210-227) -> fx_create +
// fx_load_weights; the body of the batch loop (src/feature_extraction.py:289-294) -> fx_embed /
// fx_embed_host.  The network topology follows torchvision/models/resnet.py:166-282 with
// layers=[2,2,2,2] (resnet.py:684-705) and children()[:-1] (fc dropped, avgpool kept).
#include <cmath>
#include <cstring>
#include <mutex>

#include "fx_common.cuh"

namespace fx {

int tc_conv_packed(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual,
                   __nv_bfloat16* out, float* out_f32, int n, int relu, cudaStream_t stream, const PackedLayer* ds = nullptr,
                   __nv_bfloat16* ds_out = nullptr);
int tc_tma_probe(fx_engine* e, const void* base, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                 const uint32_t* estr, int swizzle, const int* coords, int bytes, uint8_t* out_dev, cudaStream_t stream);

int tc_umma_shift_probe(fx_engine* e, const void* a_dev, const void* b_dev, int kb_elems, int shift_rows, int base_offset,
                        float* out_dev, cudaStream_t stream);

int tc_mma_rate(fx_engine* e, int n_cols, int rowb, int shift_rows, int tap_stride_rows, int iters, float* out_dev, cudaStream_t stream);

static std::mutex g_err_mutex;
static std::string g_create_error;

int set_error(fx_engine* e, int code, const std::string& msg) {
    if (e) {
        e->err = msg;
    } else {
        std::lock_guard<std::mutex> lk(g_err_mutex);
        g_create_error = msg;
    }
    return code;
}

void lane_store(fx_engine* e);
int select_lane(fx_engine* e, int lane);

// Input spatial size of every conv layer of the 224x224 network, in fx_load_weights order.
static const int kLayerHin[kNumLayers] = {224, 56, 56, 56, 56, 56, 28, 56, 28, 28, 28, 14, 28, 14, 14, 14, 7, 14, 7, 7};

static inline float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static void free_layer(PackedLayer& L) {
    cudaFree(L.w_f32);
    cudaFree(L.w_bf16);
    cudaFree(L.bias);
    cudaFree(L.w_h16);
    cudaFree(L.w_l16);
    L.w_f32 = nullptr;
    L.w_bf16 = nullptr;
    L.bias = nullptr;
    L.w_h16 = L.w_l16 = nullptr;
}

// Fold BN into the conv (fp64), reorder OIHW -> O,KH,KW,I, upload in the engine's precision.
static int pack_layer(fx_engine* e, const fx_conv_bn& src, int hin, int win, PackedLayer& L) {
    if (!src.weight || !src.gamma || !src.beta || !src.mean || !src.var)
        return set_error(e, FX_ERR_INVALID, "layer table holds a null pointer");
    if (src.cout <= 0 || src.cin <= 0 || src.kh <= 0 || src.kw <= 0 || src.stride <= 0 || src.pad < 0)
        return set_error(e, FX_ERR_INVALID, "layer table holds a non-positive dimension");
    free_layer(L);
    LayerGeom& g = L.g;
    g.cin = src.cin;
    g.cout = src.cout;
    g.kh = src.kh;
    g.kw = src.kw;
    g.stride = src.stride;
    g.pad = src.pad;
    g.hin = hin;
    g.win = win;
    g.hout = (hin + 2 * src.pad - src.kh) / src.stride + 1;
    g.wout = (win + 2 * src.pad - src.kw) / src.stride + 1;
    const bool bf16 = e->precision == FX_PRECISION_BF16;
    const int taps = g.kh * g.kw;
    L.host_w.assign((size_t)g.cout * taps * g.cin, 0.f);
    L.host_b.assign(g.cout, 0.f);
    for (int o = 0; o < g.cout; ++o) {
        const double s = (double)src.gamma[o] / std::sqrt((double)src.var[o] + (double)src.eps);
        L.host_b[o] = (float)((double)src.beta[o] - (double)src.mean[o] * s);
        for (int i = 0; i < g.cin; ++i)
            for (int t = 0; t < taps; ++t) {
                const float w = (float)((double)src.weight[((size_t)o * g.cin + i) * taps + t] * s);
                L.host_w[((size_t)o * taps + t) * g.cin + i] = bf16 ? bf16_round(w) : w;
            }
    }
    FX_CUDA(e, cudaMalloc(&L.bias, sizeof(float) * g.cout));
    FX_CUDA(e, cudaMemcpy(L.bias, L.host_b.data(), sizeof(float) * g.cout, cudaMemcpyHostToDevice));
    const bool stem = g.cin == 3;
    if (!bf16) {
        // fp32 pack [cout][kh][kw][cin_pad]; the stem pads cin 3 -> 4 to match the staging tensor
        const int cp = stem ? kIn0C : g.cin;
        if (!stem && g.cin % 4 != 0) return set_error(e, FX_ERR_UNSUPPORTED, "fp32 path: cin must be 3 or a multiple of 4");
        if (g.cout % 4 != 0) return set_error(e, FX_ERR_UNSUPPORTED, "fp32 path: cout must be a multiple of 4");
        std::vector<float> w((size_t)g.cout * taps * cp, 0.f);
        for (int o = 0; o < g.cout; ++o)
            for (int t = 0; t < taps; ++t)
                for (int i = 0; i < g.cin; ++i) w[((size_t)o * taps + t) * cp + i] = L.host_w[((size_t)o * taps + t) * g.cin + i];
        FX_CUDA(e, cudaMalloc(&L.w_f32, sizeof(float) * w.size()));
        FX_CUDA(e, cudaMemcpy(L.w_f32, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
        const bool stem7 = stem && g.kh == 7 && g.kw == 7 && g.cout == 64;
        if (e->tight_tc && (stem7 || (!stem && g.cin % 64 == 0 && g.cout % 64 == 0))) {  // split fp16 pack for the tensor-core tight mode
            std::vector<uint16_t> hi, lo;
            std::vector<float> s2d;
            if (stem7) split_stem_s2d_weights(L.host_w, g.cout, s2d);  // the stem runs as a 4x4 conv over the space-to-depth crop
            L.w_scale_log2 = split_pack_weights(stem7 ? s2d : L.host_w, hi, lo);
            FX_CUDA(e, cudaMalloc(&L.w_h16, 2 * hi.size()));
            FX_CUDA(e, cudaMalloc(&L.w_l16, 2 * lo.size()));
            FX_CUDA(e, cudaMemcpy(L.w_h16, hi.data(), 2 * hi.size(), cudaMemcpyHostToDevice));
            FX_CUDA(e, cudaMemcpy(L.w_l16, lo.data(), 2 * lo.size(), cudaMemcpyHostToDevice));
        }
        return FX_OK;
    }
    // bf16 GEMM-B pack [cout][K]
    std::vector<__nv_bfloat16> w;
    if (stem) {
        // space-to-depth form (fx_common.cuh): K = (R, S, dy, dx, c) with filter row 2R+dy, column 2S+dx;
        // the 8th row / column and channels 12..15 are zero
        if (g.kh != 7 || g.kw != 7) return set_error(e, FX_ERR_UNSUPPORTED, "stem must be 7x7");
        L.k_bf16 = 16 * kS2dC;
        w.assign((size_t)g.cout * L.k_bf16, __float2bfloat16_rn(0.f));
        for (int o = 0; o < g.cout; ++o)
            for (int r = 0; r < 7; ++r)
                for (int s = 0; s < 7; ++s)
                    for (int i = 0; i < 3; ++i)
                        w[(size_t)o * L.k_bf16 + ((r >> 1) * 4 + (s >> 1)) * kS2dC + ((r & 1) * 2 + (s & 1)) * 3 + i] =
                            __float2bfloat16_rn(L.host_w[((size_t)o * taps + r * 7 + s) * 3 + i]);
    } else {
        L.k_bf16 = taps * g.cin;
        w.resize((size_t)g.cout * L.k_bf16);
        for (size_t i = 0; i < w.size(); ++i) w[i] = __float2bfloat16_rn(L.host_w[i]);
    }
    FX_CUDA(e, cudaMalloc(&L.w_bf16, sizeof(__nv_bfloat16) * w.size()));
    FX_CUDA(e, cudaMemcpy(L.w_bf16, w.data(), sizeof(__nv_bfloat16) * w.size(), cudaMemcpyHostToDevice));
    return FX_OK;
}

// fx_profile_*: bracket launch `slot` (conv layer index, or 20 = avgpool) with events on the launching stream.
struct ProfScope {
    fx_engine* e;
    int slot;
    cudaStream_t s;
    ProfScope(fx_engine* e_, int slot_, cudaStream_t s_) : e(e_), slot(slot_), s(s_) {
        if (e->prof_on) cudaEventRecord(e->prof_ev[2 * slot], s);
    }
    ~ProfScope() {
        if (e->prof_on) cudaEventRecord(e->prof_ev[2 * slot + 1], s);
    }
};

// Tile walk of launch `li` (conv_flat.cu TileWalk).  The stem's and layer 1's tensors (103 MB each at batch 256) do not
// fit in L2, so consecutive launches sweep them in opposite directions: the preprocess kernel and the strided conv of
// layer 2 ascend, hence stem descending, layer1.0.conv1 ascending, ... layer1.1.conv2 descending.  Results do not
// depend on the walk.  FX_SCHED: 0 = every kernel ascending over contiguous runs, 1 = alternate over contiguous runs,
// 2 = alternate, tiles strided over the CTAs (the grid sweeps the tensor in time).
static int tile_sched(int li) {
    static const int mode = [] {
        const char* v = getenv("FX_SCHED");
        return v ? atoi(v) : 2;
    }();
    if (mode == 0 || li > 4) return 0;
    return ((li & 1) ? 0 : 1) | (mode == 2 ? 2 : 0);
}

// One conv layer on NHWC activations in the engine's precision.  Layer 0 reads the staging tensor.
static int run_conv(fx_engine* e, int li, const void* in, const void* residual, void* out, float* out_f32, int n, int relu,
                    cudaStream_t stream) {
    ProfScope ps(e, li, stream);
    const PackedLayer& L = e->layers[li];
    if (e->precision == FX_PRECISION_BF16) {
        if (!out_f32 && flat_supported(L.g))
            return flat_conv(e, L, static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(residual),
                             static_cast<__nv_bfloat16*>(out), n, relu, false, stream, tile_sched(li));
        return tc_conv_packed(e, L, static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(residual),
                              static_cast<__nv_bfloat16*>(out), out_f32, n, relu, stream);
    }
    float* o = out_f32 ? out_f32 : static_cast<float*>(out);
    if (li == 0)
        return simt_conv(e, L, static_cast<const float*>(in), kIn0H, kIn0W, kIn0C, 0, static_cast<const float*>(residual), o, n,
                         relu, stream);
    return simt_conv(e, L, static_cast<const float*>(in), L.g.hin, L.g.win, L.g.cin, L.g.pad, static_cast<const float*>(residual),
                     o, n, relu, stream);
}

// The truncated ResNet-18 on `n` staged images -> fp32 [n][512].
static int forward(fx_engine* e, int n, float* emb, cudaStream_t stream) {
    const bool bf16 = e->precision == FX_PRECISION_BF16;
    void *A = e->act[0], *B = e->act[1], *C = e->act[2];
    int rc;
    // stem: conv1+bn1+relu -> maxpool            (resnet.py:268-271)
    if (bf16) {  // one kernel: the max-pool runs in the conv epilogue
        ProfScope ps(e, 0, stream);
        if ((rc = flat_conv(e, e->layers[0], static_cast<const __nv_bfloat16*>(e->in0), nullptr, static_cast<__nv_bfloat16*>(A), n, 1,
                            true, stream, tile_sched(0))) != FX_OK)
            return rc;
    } else if (e->tight_tc) {
        // tight mode on the tensor cores: fp32 staging -> split space-to-depth (second half of in0) -> conv1 -> max-pool -> C
        ProfScope ps(e, 0, stream);
        char* s2d = static_cast<char*>(e->in0) + (size_t)e->max_batch * kIn0H * kIn0W * kIn0C * 4;
        if ((rc = split_stem(e, e->layers[0], static_cast<const float*>(e->in0), s2d, B, C, n, stream)) != FX_OK) return rc;
    } else {
        ProfScope ps(e, 0, stream);
        if ((rc = run_conv(e, 0, e->in0, nullptr, B, nullptr, n, 1, stream)) != FX_OK) return rc;
        if ((rc = maxpool_3x3s2(e, B, A, n, 112, 112, 64, bf16, stream)) != FX_OK) return rc;
    }
    if (!bf16 && e->tight_tc) {
        // Tight mode on the tensor cores (conv_split.cu): activations are split fp16 (hi plane, lo plane: the bytes of the
        // fp32 tensor).  X = block input, T = conv1 output, Y = block output / downsample branch.
        void *X = C, *T = B, *Y = A;
        int l = 1;
        for (int stage = 0; stage < 4; ++stage)
            for (int blk = 0; blk < 2; ++blk) {
                const bool down = stage > 0 && blk == 0;
                const bool last = stage == 3 && blk == 1;
                auto conv = [&](int idx, const void* in, const void* res, void* out, float* out_f32, int relu) {
                    ProfScope ps(e, idx, stream);
                    return split_conv(e, e->layers[idx], in, res, out_f32 ? nullptr : out, out_f32, n, relu, stream);
                };
                if (!down) {
                    if ((rc = conv(l, X, nullptr, T, nullptr, 1)) != FX_OK) return rc;
                    if ((rc = conv(l + 1, T, X, Y, last ? e->final_f32 : nullptr, 1)) != FX_OK) return rc;
                    std::swap(X, Y);
                    l += 2;
                } else {
                    if ((rc = conv(l, X, nullptr, T, nullptr, 1)) != FX_OK) return rc;
                    if ((rc = conv(l + 2, X, nullptr, Y, nullptr, 0)) != FX_OK) return rc;
                    if ((rc = conv(l + 1, T, Y, X, nullptr, 1)) != FX_OK) return rc;  // X's old contents are dead once both readers ran
                    l += 3;
                }
            }
        ProfScope ps(e, kNumLayers, stream);
        return avgpool_7x7(e, e->final_f32, false, emb, n, 49, kEmbed, stream);
    }
    // four stages of two BasicBlocks            (resnet.py:89-105, 273-276)
    int li = 1;
    for (int stage = 0; stage < 4; ++stage)
        for (int blk = 0; blk < 2; ++blk) {
            const bool down = stage > 0 && blk == 0;
            const bool last = stage == 3 && blk == 1;
            if (!down) {
                // out = relu(conv2(relu(conv1(x))) + x)
                if ((rc = run_conv(e, li, A, nullptr, B, nullptr, n, 1, stream)) != FX_OK) return rc;
                if ((rc = run_conv(e, li + 1, B, A, C, last ? e->final_f32 : nullptr, n, 1, stream)) != FX_OK) return rc;
                std::swap(A, C);
                li += 2;
            } else {
                // out = relu(conv2(relu(conv1(x))) + downsample(x))
                if (bf16) {  // conv1 (3x3/s2) and the 1x1/s2 downsample read the same input: one grouped launch
                    ProfScope ps(e, li, stream);
                    if ((rc = tc_conv_packed(e, e->layers[li], static_cast<const __nv_bfloat16*>(A), nullptr, static_cast<__nv_bfloat16*>(B),
                                             nullptr, n, 1, stream, &e->layers[li + 2], static_cast<__nv_bfloat16*>(C))) != FX_OK)
                        return rc;
                } else {
                    if ((rc = run_conv(e, li, A, nullptr, B, nullptr, n, 1, stream)) != FX_OK) return rc;
                    if ((rc = run_conv(e, li + 2, A, nullptr, C, nullptr, n, 0, stream)) != FX_OK) return rc;
                }
                if ((rc = run_conv(e, li + 1, B, C, A, nullptr, n, 1, stream)) != FX_OK) return rc;
                li += 3;
            }
        }
    // avgpool + flatten                         (resnet.py:278-279; src/feature_extraction.py:293)
    ProfScope ps(e, kNumLayers, stream);
    return avgpool_7x7(e, e->final_f32, false, emb, n, 49, kEmbed, stream);
}

// ---- lanes (fx_common.cuh): the engine's per-batch fields are the active lane's; swap on selection ----
void lane_store(fx_engine* e) {
    fx_engine::Lane& ln = e->lanes[e->cur_lane];
    ln.in0 = e->in0;
    for (int i = 0; i < 3; ++i) ln.act[i] = e->act[i];
    ln.final_f32 = e->final_f32;
    ln.img_dev = e->img_dev;
    ln.img_host = e->img_host;
    ln.img_host_free = e->img_host_free;
    ln.staged = e->staged;
}

static void lane_load(fx_engine* e, int lane) {
    const fx_engine::Lane& ln = e->lanes[lane];
    e->cur_lane = lane;
    e->in0 = ln.in0;
    for (int i = 0; i < 3; ++i) e->act[i] = ln.act[i];
    e->final_f32 = ln.final_f32;
    e->img_dev = ln.img_dev;
    e->img_host = ln.img_host;
    e->img_host_free = ln.img_host_free;
    e->staged = ln.staged;
}

// Allocate the active lane's buffers (its fields must be null).
static int lane_alloc(fx_engine* e) {
    const bool bf16 = e->precision == FX_PRECISION_BF16;
    const size_t esz = bf16 ? 2 : 4, mb = (size_t)e->max_batch;
    // FP32: the fp32 staging tensor, followed by the scratch for its split space-to-depth form (conv_split.cu)
    const size_t in0_bytes = bf16 ? mb * kS2dH * kS2dW * kS2dC * 2 : mb * kIn0H * kIn0W * kIn0C * 4 + split_stem_scratch_bytes(e->max_batch);
    e->act_bytes = mb * 112 * 112 * 64 * esz;  // largest activation: the conv1 output of the unfused (fp32) stem
    int rc = FX_OK;
    auto alloc = [&](void** p, size_t bytes) {
        if (rc != FX_OK) return;
        cudaError_t a = cudaMalloc(p, bytes);
        if (a != cudaSuccess) rc = set_error(e, a == cudaErrorMemoryAllocation ? FX_ERR_NOMEM : FX_ERR_CUDA,
                                             std::string("cudaMalloc: ") + cudaGetErrorString(a));
    };
    alloc(&e->in0, in0_bytes);
    alloc(&e->act[0], mb * 56 * 56 * 64 * esz);
    alloc(&e->act[1], e->act_bytes);
    alloc(&e->act[2], mb * 56 * 56 * 64 * esz);
    alloc(reinterpret_cast<void**>(&e->final_f32), mb * 49 * kEmbed * sizeof(float));
    if (rc != FX_OK) return rc;
    // the pad region of the staging tensor is conv zero padding and is never written again
    FX_CUDA(e, cudaMemset(e->in0, 0, in0_bytes));
    if ((rc = preprocess_lane_init(e)) != FX_OK) return rc;
    e->staged = 0;
    e->lanes[e->cur_lane].allocated = true;
    return FX_OK;
}

int select_lane(fx_engine* e, int lane) {
    if (lane < 0 || lane >= FX_MAX_LANES) return set_error(e, FX_ERR_INVALID, "lane out of range");
    if (lane == e->cur_lane) return FX_OK;
    lane_store(e);
    lane_load(e, lane);
    if (!e->lanes[lane].allocated) {
        int rc = lane_alloc(e);
        lane_store(e);
        if (rc != FX_OK) return rc;
    }
    return FX_OK;
}

}  // namespace fx

using namespace fx;

extern "C" {

const char* fx_version(void) { return "fx_b200 0.1.0 (sm_100a)"; }
int fx_abi_version(void) { return FX_ABI_VERSION; }

const char* fx_last_error(fx_handle h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    return g_create_error.c_str();
}

void fx_destroy(fx_handle e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto& L : e->layers) free_layer(L);
    preprocess_free(e);
    tc_free(e);
    jpeg_free(e);
    lane_store(e);
    for (auto& ln : e->lanes) {
        cudaFree(ln.in0);
        for (auto& a : ln.act) cudaFree(a);
        cudaFree(ln.final_f32);
        cudaFree(ln.img_dev);
        if (ln.img_host) cudaFreeHost(ln.img_host);
        if (ln.img_host_free) cudaEventDestroy(ln.img_host_free);
    }
    for (auto& hs : e->slots) {
        cudaFree(hs.src_dev);
        cudaFree(hs.emb_dev);
        if (hs.copied) cudaEventDestroy(hs.copied);
        if (hs.done) cudaEventDestroy(hs.done);
    }
    for (auto& st : e->lane_stream)
        if (st) cudaStreamDestroy(st);
    for (auto& g : e->step_graph) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    cudaFree(e->head_w);
    cudaFree(e->head_b);
    cudaFree(e->post_scratch);
    for (auto& ev : e->prof_ev)
        if (ev) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

int fx_create(fx_handle* out, int device, int max_batch, int precision) {
    if (!out) return set_error(nullptr, FX_ERR_INVALID, "fx_create: null output pointer");
    *out = nullptr;
    if (max_batch < 1 || max_batch > 65536) return set_error(nullptr, FX_ERR_INVALID, "fx_create: max_batch out of range");
    if (precision != FX_PRECISION_BF16 && precision != FX_PRECISION_FP32)
        return set_error(nullptr, FX_ERR_INVALID, "fx_create: unknown precision");
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return set_error(nullptr, FX_ERR_UNSUPPORTED,
                         std::string("fx_create: no CUDA device (") + cudaGetErrorString(err) + "); this library has no CPU path");
    if (device < 0 || device >= count) return set_error(nullptr, FX_ERR_INVALID, "fx_create: device index out of range");
    cudaDeviceProp prop;
    if ((err = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return set_error(nullptr, FX_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(err));
    if (prop.major != 10)
        return set_error(nullptr, FX_ERR_UNSUPPORTED,
                         "fx_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                             "; this library is built for sm_100a (B200) only");
    if ((err = cudaSetDevice(device)) != cudaSuccess)
        return set_error(nullptr, FX_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(err));

    fx_engine* e = new fx_engine();
    e->device = device;
    e->max_batch = max_batch;
    e->precision = precision;
    e->sm_count = prop.multiProcessorCount;
    auto fail = [&](int rc) {
        set_error(nullptr, rc, e->err);
        fx_destroy(e);
        return rc;
    };
    int rc = lane_alloc(e);  // lane 0
    if (rc != FX_OK) return fail(rc);
    if ((err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail(set_error(e, FX_ERR_CUDA, cudaGetErrorString(err)));
    if (const char* g = getenv("FX_GRAPHS")) e->graphs_on = g[0] != '0';
    if (const char* g = getenv("FX_TIGHT_SIMT")) e->tight_tc = g[0] != '1';
    if (const char* g = getenv("FX_GRAPH_MAX_BATCH")) e->graph_max_batch = atoi(g);
    if ((rc = preprocess_init(e)) != FX_OK) return fail(rc);
    if ((rc = tc_init(e)) != FX_OK) return fail(rc);
    *out = e;
    return FX_OK;
}

int fx_load_weights(fx_handle e, const fx_conv_bn* layers, int n_layers) {
    if (!e) return FX_ERR_INVALID;
    if (!layers || n_layers != kNumLayers)
        return set_error(e, FX_ERR_INVALID, "fx_load_weights: expected the 20 conv+bn groups of resnet18 minus fc");
    FX_CUDA(e, cudaSetDevice(e->device));
    static const int cin[kNumLayers] = {3, 64, 64, 64, 64, 64, 128, 64, 128, 128, 128, 256, 128, 256, 256, 256, 512, 256, 512, 512};
    static const int cout[kNumLayers] = {64, 64, 64, 64, 64, 128, 128, 128, 128, 128, 256, 256, 256, 256, 256, 512, 512, 512, 512, 512};
    static const int ks[kNumLayers] = {7, 3, 3, 3, 3, 3, 3, 1, 3, 3, 3, 3, 1, 3, 3, 3, 3, 1, 3, 3};
    static const int st[kNumLayers] = {2, 1, 1, 1, 1, 2, 1, 2, 1, 1, 2, 1, 2, 1, 1, 2, 1, 2, 1, 1};
    e->weights_loaded = false;
    for (int i = 0; i < kNumLayers; ++i) {
        const fx_conv_bn& s = layers[i];
        if (s.cin != cin[i] || s.cout != cout[i] || s.kh != ks[i] || s.kw != ks[i] || s.stride != st[i] || s.pad != ks[i] / 2)
            return set_error(e, FX_ERR_INVALID, "fx_load_weights: layer " + std::to_string(i) + " does not match resnet18");
        int rc = pack_layer(e, s, kLayerHin[i], kLayerHin[i], e->layers[i]);
        if (rc != FX_OK) return rc;
    }
    e->weights_loaded = true;
    e->weights_epoch++;  // captured step graphs hold the old weight pointers
    return FX_OK;
}

int fx_preprocess_nchw_f32(fx_handle e, const uint8_t* src_dev, const fx_image_desc* descs, int n, float* out_dev, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (n < 0 || n > e->max_batch || (n > 0 && (!src_dev || !descs || !out_dev)))
        return set_error(e, FX_ERR_INVALID, "fx_preprocess_nchw_f32: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    return preprocess_run(e, src_dev, descs, n, PreOut::NCHW_F32, out_dev, static_cast<cudaStream_t>(stream));
}

int fx_preprocess(fx_handle e, const uint8_t* src_dev, const fx_image_desc* descs, int n, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (n < 0 || n > e->max_batch || (n > 0 && (!src_dev || !descs))) return set_error(e, FX_ERR_INVALID, "fx_preprocess: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    e->staged = 0;
    int rc = preprocess_run(e, src_dev, descs, n, e->precision == FX_PRECISION_BF16 ? PreOut::IN0_BF16 : PreOut::IN0_F32, e->in0,
                            static_cast<cudaStream_t>(stream));
    if (rc == FX_OK) e->staged = n;
    return rc;
}

int fx_stage_nchw_f32(fx_handle e, const float* in_dev, int n, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (n < 0 || n > e->max_batch || (n > 0 && !in_dev)) return set_error(e, FX_ERR_INVALID, "fx_stage_nchw_f32: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    e->staged = 0;
    int rc = stage_nchw_run(e, in_dev, n, static_cast<cudaStream_t>(stream));
    if (rc == FX_OK) e->staged = n;
    return rc;
}

int fx_forward(fx_handle e, int n, float* emb_dev, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (!e->weights_loaded) return set_error(e, FX_ERR_STATE, "fx_forward: fx_load_weights has not succeeded");
    if (n < 0 || n > e->staged) return set_error(e, FX_ERR_STATE, "fx_forward: more images requested than staged");
    if (n == 0) return FX_OK;
    if (!emb_dev) return set_error(e, FX_ERR_INVALID, "fx_forward: null output");
    FX_CUDA(e, cudaSetDevice(e->device));
    return forward(e, n, emb_dev, static_cast<cudaStream_t>(stream));
}

int fx_set_transform(fx_handle e, int transform) {
    if (!e) return FX_ERR_INVALID;
    if (transform != FX_TRANSFORM_EXTRACT && transform != FX_TRANSFORM_SQUARE224)
        return set_error(e, FX_ERR_INVALID, "fx_set_transform: unknown transform");
    e->transform = transform;
    return FX_OK;
}

int fx_load_head(fx_handle e, const float* weight, const float* bias, int num_classes) {
    if (!e) return FX_ERR_INVALID;
    if (!weight || !bias || num_classes < 1 || num_classes > FX_MAX_CLASSES)
        return set_error(e, FX_ERR_INVALID, "fx_load_head: null pointer or num_classes outside 1.." + std::to_string(FX_MAX_CLASSES));
    FX_CUDA(e, cudaSetDevice(e->device));
    FX_CUDA(e, cudaDeviceSynchronize());  // a previous head may still be in use
    // (the captured step graphs end at the average pool and the post-processing scratch has nothing to do with the
    // head: neither is touched here)
    cudaFree(e->head_w);
    cudaFree(e->head_b);
    e->head_w = e->head_b = nullptr;
    e->head_classes = 0;
    FX_CUDA(e, cudaMalloc(&e->head_w, sizeof(float) * kEmbed * num_classes));
    FX_CUDA(e, cudaMalloc(&e->head_b, sizeof(float) * num_classes));
    FX_CUDA(e, cudaMemcpy(e->head_w, weight, sizeof(float) * kEmbed * num_classes, cudaMemcpyHostToDevice));
    FX_CUDA(e, cudaMemcpy(e->head_b, bias, sizeof(float) * num_classes, cudaMemcpyHostToDevice));
    e->head_classes = num_classes;
    return FX_OK;
}

int fx_classify(fx_handle e, int n, float* emb_dev, float* logits_dev, float* probs_dev, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (!e->head_classes) return set_error(e, FX_ERR_STATE, "fx_classify: fx_load_head has not succeeded");
    int rc = fx_forward(e, n, emb_dev, stream);
    if (rc != FX_OK || n == 0) return rc;
    return head_run(e, emb_dev, n, logits_dev, probs_dev, static_cast<cudaStream_t>(stream));
}

int fx_select_lane(fx_handle e, int lane) {
    if (!e) return FX_ERR_INVALID;
    FX_CUDA(e, cudaSetDevice(e->device));
    return select_lane(e, lane);
}

// ---- CUDA graph of a whole step for launch-bound batch sizes ---------------------------------
static void graph_destroy(fx_engine::StepGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g = fx_engine::StepGraph();
}

// Capture fx_preprocess (plan hit: one launch, no upload) + the trunk on `s`; find the two kernel nodes whose
// parameters depend on the call.  Returns false (and leaves no graph) if anything about the capture fails.
static bool graph_build(fx_engine* e, fx_engine::StepGraph& g, const uint8_t* src_dev, const fx_image_desc* descs, int n, float* emb_dev,
                        const void* pre_func, cudaStream_t s) {
    graph_destroy(g);
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const uint64_t launches = e->launches;
    e->capturing = true;
    int rc = fx_preprocess(e, src_dev, descs, n, s);
    if (rc == FX_OK) rc = fx_forward(e, n, emb_dev, s);
    e->capturing = false;
    e->launches = launches;  // nothing has run yet
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(s, &graph);
    if (rc != FX_OK || end != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return false;
    }
    g.graph = graph;
    size_t count = 0;
    if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess || cudaGraphGetNodes(graph, nullptr, &count) != cudaSuccess) {
        cudaGetLastError();
        graph_destroy(g);
        return false;
    }
    std::vector<cudaGraphNode_t> nodes(count);
    cudaGraphGetNodes(graph, nodes.data(), &count);
    const void* pool_func = avgpool_kernel_ptr(false);
    for (cudaGraphNode_t node : nodes) {
        cudaGraphNodeType type;
        cudaKernelNodeParams kp;
        if (cudaGraphNodeGetType(node, &type) != cudaSuccess || type != cudaGraphNodeTypeKernel) continue;
        if (cudaGraphKernelNodeGetParams(node, &kp) != cudaSuccess) continue;
        g.kernels++;
        if (kp.func == pre_func) g.pre_node = node;
        if (kp.func == pool_func) g.pool_node = node;
    }
    if (!g.pre_node || !g.pool_node) {
        cudaGetLastError();
        graph_destroy(g);
        return false;
    }
    g.n = n;
    g.plan_serial = e->pre_plan[e->cur_lane].serial;
    g.weights_epoch = e->weights_epoch;
    g.src = src_dev;
    g.emb = emb_dev;
    return true;
}

// Replace kernel parameter `index` of a graph node (the other parameters keep the values captured with the graph).
static bool graph_patch(fx_engine::StepGraph& g, cudaGraphNode_t node, int index, void* value_ptr, int n_params) {
    cudaKernelNodeParams kp;
    if (cudaGraphKernelNodeGetParams(node, &kp) != cudaSuccess) return false;
    void* args[8];
    for (int i = 0; i < n_params; ++i) args[i] = kp.kernelParams[i];
    args[index] = value_ptr;
    kp.kernelParams = args;
    return cudaGraphExecKernelNodeSetParams(g.exec, node, &kp) == cudaSuccess;
}

int fx_embed(fx_handle e, const uint8_t* src_dev, const fx_image_desc* descs, int n, float* emb_dev, void* stream) {
    if (!e) return FX_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // Launch-bound regime (19 launches of a few microseconds each at batch <= 64): replay the step as one CUDA graph.
    // Needs a capturable stream, the lane's preprocess plan to repeat (no descriptor upload inside the graph) and no
    // per-launch profiling events.
    if (e->graphs_on && s != nullptr && s != cudaStreamLegacy && !e->prof_on && n > 0 && n <= e->graph_max_batch && n <= e->max_batch &&
        e->weights_loaded && src_dev && descs && emb_dev) {
        const PreOut mode = e->precision == FX_PRECISION_BF16 ? PreOut::IN0_BF16 : PreOut::IN0_F32;
        const void* pre_func = nullptr;
        if (preprocess_plan_hit(e, descs, n, mode, &pre_func)) {
            FX_CUDA(e, cudaSetDevice(e->device));
            FX_CUDA(e, cudaStreamWaitEvent(s, e->img_host_free, 0));  // the plan's upload may be on another stream
            fx_engine::StepGraph& g = e->step_graph[e->cur_lane];
            const bool usable = g.exec && g.n == n && g.plan_serial == e->pre_plan[e->cur_lane].serial && g.weights_epoch == e->weights_epoch;
            if (!usable && !graph_build(e, g, src_dev, descs, n, emb_dev, pre_func, s)) e->graphs_on = false;  // never try again on this handle
            if (g.exec && e->graphs_on) {
                bool ok = true;
                if (g.src != src_dev) {
                    const void* v = src_dev;
                    ok = graph_patch(g, g.pre_node, 0, &v, e->pre_plan[e->cur_lane].s2d ? 5 : 7);
                    g.src = src_dev;
                }
                if (ok && g.emb != emb_dev) {
                    float* v = emb_dev;
                    ok = graph_patch(g, g.pool_node, 1, &v, 5);
                    g.emb = emb_dev;
                }
                if (ok && cudaGraphLaunch(g.exec, s) == cudaSuccess) {
                    e->staged = n;
                    e->launches += g.kernels;
                    return FX_OK;
                }
                cudaGetLastError();
                graph_destroy(g);
                e->graphs_on = false;
            }
        }
    }
    int rc = fx_preprocess(e, src_dev, descs, n, stream);
    if (rc != FX_OK) return rc;
    return fx_forward(e, n, emb_dev, stream);
}

// Host-buffer path, pipelined over FX_HOST_SLOTS slots: the H2D copy of a batch (copy stream) overlaps the kernels
// of the earlier ones (two lane streams); the D2H of the embeddings follows the kernels on the lane's stream.
}  // extern "C"

namespace fx {

// The two halves of a pipelined host-buffer step, shared by fx_embed_host_async* (engine.cu) and fx_embed_files_async
// (decode.cu).  prepare: the slot's events / device buffers exist, its previous batch is finished, src_dev holds
// `total_bytes`.  The caller then fills hs.src_dev on e->copy_stream and records hs.copied.  compute: the lane's stream
// waits for hs.copied, runs preprocess + trunk, and either copies the rows to emb_host or leaves them at emb_dev_out.
int embed_slot_prepare(fx_engine* e, int slot, size_t total_bytes) {
    fx_engine::HostSlot& hs = e->slots[slot];
    if (!hs.copied) {
        FX_CUDA(e, cudaEventCreateWithFlags(&hs.copied, cudaEventDisableTiming));
        FX_CUDA(e, cudaEventCreateWithFlags(&hs.done, cudaEventDisableTiming));
        FX_CUDA(e, cudaMalloc(&hs.emb_dev, sizeof(float) * kEmbed * e->max_batch));
    }
    if (hs.busy) {  // the caller reuses a slot it never waited on: finish it first
        FX_CUDA(e, cudaEventSynchronize(hs.done));
        hs.busy = false;
    }
    if (total_bytes > hs.cap) {
        cudaFree(hs.src_dev);
        hs.src_dev = nullptr;
        hs.cap = 0;
        const size_t cap = total_bytes + total_bytes / 4 + 256;
        cudaError_t a = cudaMalloc(&hs.src_dev, cap);
        if (a != cudaSuccess) return set_error(e, FX_ERR_NOMEM, std::string("cudaMalloc(h2d staging): ") + cudaGetErrorString(a));
        hs.cap = cap;
    }
    return FX_OK;
}

int embed_slot_compute(fx_engine* e, int slot, const fx_image_desc* descs, int n, float* emb_host, float* emb_dev_out) {
    fx_engine::HostSlot& hs = e->slots[slot];
    const int lane = slot % FX_MAX_LANES;
    if (!e->lane_stream[lane]) FX_CUDA(e, cudaStreamCreateWithFlags(&e->lane_stream[lane], cudaStreamNonBlocking));
    cudaStream_t s = e->lane_stream[lane];
    FX_CUDA(e, cudaStreamWaitEvent(s, hs.copied, 0));
    const int caller_lane = e->cur_lane;  // the caller's selection is restored afterwards
    int rc = select_lane(e, lane);
    if (rc == FX_OK) rc = fx_embed(e, hs.src_dev, descs, n, emb_dev_out ? emb_dev_out : hs.emb_dev, s);
    const int rc2 = select_lane(e, caller_lane);
    if (rc != FX_OK) return rc;
    if (rc2 != FX_OK) return rc2;
    if (!emb_dev_out) FX_CUDA(e, cudaMemcpyAsync(emb_host, hs.emb_dev, sizeof(float) * kEmbed * n, cudaMemcpyDeviceToHost, s));
    FX_CUDA(e, cudaEventRecord(hs.done, s));
    hs.busy = true;
    return FX_OK;
}

}  // namespace fx

extern "C" {

// emb_host: rows are copied back to the host (fx_embed_host_async); emb_dev_out: rows are written straight to that device
// address, e.g. the rank's slot of an all-gather buffer (fx_embed_host_async_dev).  Exactly one of them is given.
static int embed_host_submit(fx_handle e, int slot, const uint8_t* src_host, size_t total_bytes, const fx_image_desc* descs, int n,
                             float* emb_host, float* emb_dev_out) {
    if (!e) return FX_ERR_INVALID;
    if (slot < 0 || slot >= FX_HOST_SLOTS || n < 0 || n > e->max_batch || (n > 0 && (!src_host || !descs || (!emb_host && !emb_dev_out))))
        return set_error(e, FX_ERR_INVALID, "fx_embed_host_async: bad arguments");
    for (int i = 0; i < n; ++i) {
        const size_t need = descs[i].offset + (size_t)descs[i].height * descs[i].width * descs[i].channels;
        if (descs[i].height < 1 || descs[i].width < 1 || need > total_bytes)
            return set_error(e, FX_ERR_INVALID, "fx_embed_host_async: image " + std::to_string(i) + " lies outside the buffer");
    }
    FX_CUDA(e, cudaSetDevice(e->device));
    int rc = embed_slot_prepare(e, slot, total_bytes);
    if (rc != FX_OK || n == 0) return rc;
    fx_engine::HostSlot& hs = e->slots[slot];
    static const int dbg = getenv("FX_DEBUG_E2E") ? atoi(getenv("FX_DEBUG_E2E")) : 0;  // measurement knob: 1 = skip the H2D copy, 2 = skip the D2H copy, 4 = the H2D copy alone
    static const bool rows_only = !(getenv("FX_H2D_ROWS") && getenv("FX_H2D_ROWS")[0] == '0');
    if (!(dbg & 1)) {
        // A uniform batch (same size, constant stride: every batch of a synthetic / pre-decoded dataset) is copied as ONE 2-D
        // copy of just the source rows the crop touches -- Resize(256) + CenterCrop(224) never reads the outer 6 % of the rows
        // on either side (224 x 224 sources: rows 14..209), and the preprocess kernels never load them; 12.5 % fewer PCIe bytes,
        // which is what bounds the end-to-end rate from four GPUs up.  Anything else: the whole buffer, as before.
        bool done = false;
        if (rows_only && n >= 1) {
            const fx_image_desc& d0 = descs[0];
            const size_t img_bytes = (size_t)d0.height * d0.width * d0.channels;
            const size_t stride = n > 1 ? (size_t)(descs[1].offset - descs[0].offset) : img_bytes;
            bool uniform = stride >= img_bytes;
            for (int i = 1; i < n && uniform; ++i)
                uniform = descs[i].height == d0.height && descs[i].width == d0.width && descs[i].channels == d0.channels &&
                          descs[i].offset == d0.offset + (uint64_t)i * stride;
            int lo = 0, hi = 0;
            if (uniform && preprocess_rows_needed(e, d0.height, d0.width, &lo, &hi) == FX_OK && hi > lo && hi - lo < d0.height) {
                const size_t rowb = (size_t)d0.width * d0.channels, skip = d0.offset + (size_t)lo * rowb;
                // (the crop's COLUMNS as well -- one 3-D copy of 588-byte row pieces, 12 % fewer bytes again -- was measured: the copy
                // engine moves such short pieces at 16.4 GB/s against 54.5 GB/s for whole rows; tools/h2d_window_probe.py)
                FX_CUDA(e, cudaMemcpy2DAsync(hs.src_dev + skip, stride, src_host + skip, stride, (size_t)(hi - lo) * rowb, n,
                                             cudaMemcpyHostToDevice, e->copy_stream));
                e->h2d_bytes += (size_t)(hi - lo) * rowb * n;
                done = true;
            }
        }
        if (!done) {
            FX_CUDA(e, cudaMemcpyAsync(hs.src_dev, src_host, total_bytes, cudaMemcpyHostToDevice, e->copy_stream));
            e->h2d_bytes += total_bytes;
        }
    }
    FX_CUDA(e, cudaEventRecord(hs.copied, e->copy_stream));
    if (dbg & 4) {  // measurement knob: the copy alone (what the H2D leg can carry with this copy shape)
        FX_CUDA(e, cudaEventRecord(hs.done, e->copy_stream));
        hs.busy = true;
        return FX_OK;
    }
    rc = embed_slot_compute(e, slot, descs, n, (dbg & 2) ? nullptr : emb_host, (dbg & 2) && !emb_dev_out ? hs.emb_dev : emb_dev_out);
    if (rc != FX_OK) cudaStreamSynchronize(e->copy_stream);  // the slot is not marked busy: the caller's buffer must be free of DMA when the error returns
    return rc;
}

int fx_embed_host_async(fx_handle e, int slot, const uint8_t* src_host, size_t total_bytes, const fx_image_desc* descs, int n,
                        float* emb_host) {
    return embed_host_submit(e, slot, src_host, total_bytes, descs, n, emb_host, nullptr);
}

int fx_embed_host_async_dev(fx_handle e, int slot, const uint8_t* src_host, size_t total_bytes, const fx_image_desc* descs, int n,
                            float* emb_dev) {
    if (e && n > 0 && !emb_dev) return set_error(e, FX_ERR_INVALID, "fx_embed_host_async_dev: null output");
    return embed_host_submit(e, slot, src_host, total_bytes, descs, n, nullptr, emb_dev);
}

int fx_embed_host_wait(fx_handle e, int slot) {
    if (!e) return FX_ERR_INVALID;
    if (slot < 0 || slot >= FX_HOST_SLOTS) return set_error(e, FX_ERR_INVALID, "fx_embed_host_wait: bad slot");
    fx_engine::HostSlot& hs = e->slots[slot];
    if (!hs.busy) return FX_OK;
    FX_CUDA(e, cudaSetDevice(e->device));
    FX_CUDA(e, cudaEventSynchronize(hs.done));
    hs.busy = false;
    return FX_OK;
}

int fx_embed_host(fx_handle e, const uint8_t* src_host, size_t total_bytes, const fx_image_desc* descs, int n, float* emb_host) {
    int rc = fx_embed_host_async(e, 0, src_host, total_bytes, descs, n, emb_host);
    if (rc != FX_OK) return rc;
    return fx_embed_host_wait(e, 0);
}

int fx_column_stats(fx_handle e, const float* emb_dev, int64_t n, int d, double* col_mean_dev, double* col_std_dev, double* col_var_dev,
                    fx_matrix_stats* stats, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (!emb_dev || n < 1 || d < 1 || d > 4096) return set_error(e, FX_ERR_INVALID, "fx_column_stats: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    return post_column_stats(e, emb_dev, n, d, col_mean_dev, col_std_dev, col_var_dev, stats, static_cast<cudaStream_t>(stream));
}

int fx_standardize(fx_handle e, const float* emb_dev, int64_t n, int d, const double* col_mean_dev, const double* col_scale_dev,
                   float* out_dev, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (!emb_dev || !col_mean_dev || !col_scale_dev || !out_dev || n < 1 || d < 1 || d > 4096)
        return set_error(e, FX_ERR_INVALID, "fx_standardize: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    return post_standardize(e, emb_dev, n, d, col_mean_dev, col_scale_dev, out_dev, static_cast<cudaStream_t>(stream));
}

int fx_neighbor_probe(fx_handle e, const float* emb_dev, int64_t n, int d, const int64_t* query_rows_host, int q,
                      int64_t* neighbor_rows_host, float* similarity_host, void* stream) {
    if (!e) return FX_ERR_INVALID;
    if (!emb_dev || !query_rows_host || !neighbor_rows_host || !similarity_host || n < 2 || d < 1 || d > 4096 || q < 1 || q > 32)
        return set_error(e, FX_ERR_INVALID, "fx_neighbor_probe: bad arguments (n >= 2, 1 <= q <= 32, d <= 4096)");
    for (int j = 0; j < q; ++j)
        if (query_rows_host[j] < 0 || query_rows_host[j] >= n) return set_error(e, FX_ERR_INVALID, "fx_neighbor_probe: query row out of range");
    FX_CUDA(e, cudaSetDevice(e->device));
    return post_neighbor_probe(e, emb_dev, n, d, query_rows_host, q, neighbor_rows_host, similarity_host, static_cast<cudaStream_t>(stream));
}

uint64_t fx_launch_count(fx_handle e) { return e ? e->launches : 0; }
uint64_t fx_h2d_bytes(fx_handle e) { return e ? e->h2d_bytes : 0; }

int fx_profile_enable(fx_handle e, int on) {
    if (!e) return FX_ERR_INVALID;
    FX_CUDA(e, cudaSetDevice(e->device));
    if (on && !e->prof_ev[0])
        for (auto& ev : e->prof_ev) FX_CUDA(e, cudaEventCreate(&ev));
    e->prof_on = on != 0;
    return FX_OK;
}

int fx_profile_read(fx_handle e, float* ms, int capacity) {
    if (!e) return FX_ERR_INVALID;
    if (!ms || capacity < kNumLayers + 1 || !e->prof_ev[0]) return set_error(e, FX_ERR_INVALID, "fx_profile_read: profiling was never enabled / buffer too small");
    FX_CUDA(e, cudaSetDevice(e->device));
    for (int i = 0; i <= kNumLayers; ++i) {
        ms[i] = 0.f;
        if (cudaEventSynchronize(e->prof_ev[2 * i + 1]) == cudaSuccess) cudaEventElapsedTime(&ms[i], e->prof_ev[2 * i], e->prof_ev[2 * i + 1]);
    }
    cudaGetLastError();  // an unrecorded slot reports an error we do not care about
    return FX_OK;
}

// ---- test / inspection entry points -------------------------------------------------------

int fx_debug_conv(fx_handle e, const fx_conv_bn* layer, int hin, int win, const float* in_dev, const float* residual_dev, int n,
                  int relu_flags, float* out_dev, void* stream_) {
    if (!e) return FX_ERR_INVALID;
    const int relu = relu_flags & 1;
    const bool f32_out = (relu_flags & 2) != 0 && e->precision == FX_PRECISION_BF16 && layer && layer->cin != 3;
    if (!layer || !in_dev || !out_dev || n < 1 || n > e->max_batch) return set_error(e, FX_ERR_INVALID, "fx_debug_conv: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PackedLayer L;
    int rc = pack_layer(e, *layer, hin, win, L);
    if (rc != FX_OK) {
        free_layer(L);
        return rc;
    }
    const LayerGeom& g = L.g;
    const size_t in_count = (size_t)n * hin * win * g.cin, out_count = (size_t)n * g.hout * g.wout * g.cout;
    const bool bf16 = e->precision == FX_PRECISION_BF16;
    const bool stem = g.cin == 3;
    if (stem && (hin != kCrop || win != kCrop)) {
        free_layer(L);
        return set_error(e, FX_ERR_UNSUPPORTED, "fx_debug_conv: 3-channel input must be 224x224");
    }
    const size_t small_cap = (size_t)e->max_batch * 56 * 56 * 64, big_cap = (size_t)e->max_batch * 112 * 112 * 64;
    if ((!stem && in_count > small_cap) || out_count > big_cap || (residual_dev && out_count > small_cap)) {
        free_layer(L);
        return set_error(e, FX_ERR_INVALID, "fx_debug_conv: activation larger than the engine workspace");
    }
    do {
        const void* in_act = in_dev;
        if (stem) {
            if ((rc = pad_nhwc3_to_in0(e, in_dev, e->in0, bf16, n, stream)) != FX_OK) break;
            in_act = e->in0;
        }
        if (!bf16) {
            if (stem)
                rc = simt_conv(e, L, static_cast<const float*>(in_act), kIn0H, kIn0W, kIn0C, 0, residual_dev, out_dev, n, relu, stream);
            else if (L.w_h16) {  // tight mode on the tensor cores: split the operands, run, join
                if ((rc = f32_to_split(e, in_dev, e->act[0], in_count, stream)) != FX_OK) break;
                if (residual_dev && (rc = f32_to_split(e, residual_dev, e->act[2], out_count, stream)) != FX_OK) break;
                if ((rc = split_conv(e, L, e->act[0], residual_dev ? e->act[2] : nullptr, e->act[1], nullptr, n, relu, stream)) != FX_OK) break;
                rc = split_to_f32(e, e->act[1], out_dev, out_count, stream);
            } else
                rc = simt_conv(e, L, in_dev, hin, win, g.cin, g.pad, residual_dev, out_dev, n, relu, stream);
            break;
        }
        __nv_bfloat16* bin = static_cast<__nv_bfloat16*>(e->act[0]);
        __nv_bfloat16* bres = static_cast<__nv_bfloat16*>(e->act[2]);
        __nv_bfloat16* bout = static_cast<__nv_bfloat16*>(e->act[1]);
        if (!stem) {
            if ((rc = f32_to_bf16(e, in_dev, bin, in_count, stream)) != FX_OK) break;
            in_act = bin;
        }
        if (residual_dev && (rc = f32_to_bf16(e, residual_dev, bres, out_count, stream)) != FX_OK) break;
        if (f32_out) {  // the accumulator itself (+bias, +residual, ReLU), not rounded to bf16: tensor-core accumulation probe
            rc = tc_conv_packed(e, L, static_cast<const __nv_bfloat16*>(in_act), residual_dev ? bres : nullptr, nullptr, out_dev, n, relu,
                                stream);
            break;
        }
        if (flat_supported(L.g))
            rc = flat_conv(e, L, static_cast<const __nv_bfloat16*>(in_act), residual_dev ? bres : nullptr, bout, n, relu, false, stream);
        else
            rc = tc_conv_packed(e, L, static_cast<const __nv_bfloat16*>(in_act), residual_dev ? bres : nullptr, bout, nullptr, n, relu,
                                stream);
        if (rc != FX_OK) break;
        rc = bf16_to_f32(e, bout, out_dev, out_count, stream);
    } while (0);
    cudaError_t serr = cudaStreamSynchronize(stream);  // the packed weights are freed below
    free_layer(L);
    if (rc == FX_OK && serr != cudaSuccess) rc = set_error(e, FX_ERR_CUDA, std::string("fx_debug_conv: ") + cudaGetErrorString(serr));
    return rc;
}

int fx_debug_stem_pool(fx_handle e, const fx_conv_bn* layer, const float* in_dev, int n, float* out_dev, void* stream_) {
    if (!e) return FX_ERR_INVALID;
    if (!layer || !in_dev || !out_dev || n < 1 || n > e->max_batch) return set_error(e, FX_ERR_INVALID, "fx_debug_stem_pool: bad arguments");
    if (e->precision != FX_PRECISION_BF16 && !e->tight_tc)
        return set_error(e, FX_ERR_UNSUPPORTED, "fx_debug_stem_pool: bf16 engines and the tensor-core tight mode only");
    FX_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    PackedLayer L;
    int rc = pack_layer(e, *layer, kCrop, kCrop, L);
    if (e->precision != FX_PRECISION_BF16) {  // tight mode: fp32 staging -> split space-to-depth -> conv1 -> max-pool -> fp32
        if (rc == FX_OK) rc = pad_nhwc3_to_in0(e, in_dev, e->in0, false, n, stream);
        char* s2d = static_cast<char*>(e->in0) + (size_t)e->max_batch * kIn0H * kIn0W * kIn0C * 4;
        if (rc == FX_OK) rc = split_stem(e, L, static_cast<const float*>(e->in0), s2d, e->act[1], e->act[2], n, stream);
        if (rc == FX_OK) rc = split_to_f32(e, e->act[2], out_dev, (size_t)n * 56 * 56 * 64, stream);
        cudaError_t serr = cudaStreamSynchronize(stream);
        free_layer(L);
        if (rc == FX_OK && serr != cudaSuccess) rc = set_error(e, FX_ERR_CUDA, std::string("fx_debug_stem_pool: ") + cudaGetErrorString(serr));
        return rc;
    }
    if (rc == FX_OK && !flat_supported(L.g)) rc = set_error(e, FX_ERR_UNSUPPORTED, "fx_debug_stem_pool: not the 7x7/s2 stem");
    if (rc == FX_OK) rc = pad_nhwc3_to_in0(e, in_dev, e->in0, true, n, stream);
    __nv_bfloat16* bout = static_cast<__nv_bfloat16*>(e->act[0]);
    if (rc == FX_OK) rc = flat_conv(e, L, static_cast<const __nv_bfloat16*>(e->in0), nullptr, bout, n, 1, true, stream);
    if (rc == FX_OK) rc = bf16_to_f32(e, bout, out_dev, (size_t)n * 56 * 56 * 64, stream);
    cudaError_t serr = cudaStreamSynchronize(stream);
    free_layer(L);
    if (rc == FX_OK && serr != cudaSuccess) rc = set_error(e, FX_ERR_CUDA, std::string("fx_debug_stem_pool: ") + cudaGetErrorString(serr));
    return rc;
}

int fx_debug_staging(fx_handle e, int n, void* out_dev, size_t bytes, void* stream_) {
    if (!e) return FX_ERR_INVALID;
    const size_t per = e->precision == FX_PRECISION_BF16 ? (size_t)kS2dH * kS2dW * kS2dC * 2 : (size_t)kIn0H * kIn0W * kIn0C * 4;
    if (n < 0 || n > e->staged || !out_dev || bytes != per * (size_t)n) return set_error(e, FX_ERR_INVALID, "fx_debug_staging: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    FX_CUDA(e, cudaMemcpyAsync(out_dev, e->in0, bytes, cudaMemcpyDeviceToDevice, stream));
    FX_CUDA(e, cudaStreamSynchronize(stream));
    return FX_OK;
}

int fx_debug_folded(fx_handle e, int layer, float* weight_host, float* bias_host) {
    if (!e) return FX_ERR_INVALID;
    if (!e->weights_loaded || layer < 0 || layer >= kNumLayers) return set_error(e, FX_ERR_INVALID, "fx_debug_folded: bad layer");
    const PackedLayer& L = e->layers[layer];
    if (weight_host) std::memcpy(weight_host, L.host_w.data(), sizeof(float) * L.host_w.size());
    if (bias_host) std::memcpy(bias_host, L.host_b.data(), sizeof(float) * L.host_b.size());
    return FX_OK;
}

int fx_debug_tma_probe(fx_handle e, const void* base_dev, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       const uint32_t* elem_strides, int swizzle, const int* coords, int bytes, uint8_t* out_dev) {
    if (!e) return FX_ERR_INVALID;
    FX_CUDA(e, cudaSetDevice(e->device));
    int rc = tc_tma_probe(e, base_dev, dims, strides_bytes, box, elem_strides, swizzle, coords, bytes, out_dev, nullptr);
    if (rc != FX_OK) return rc;
    FX_CUDA(e, cudaDeviceSynchronize());
    return FX_OK;
}

int fx_debug_umma_shift(fx_handle e, const void* a_dev, const void* b_dev, int kb_elems, int shift_rows, int base_offset,
                        float* out_dev) {
    if (!e) return FX_ERR_INVALID;
    if (!a_dev || !b_dev || !out_dev) return set_error(e, FX_ERR_INVALID, "fx_debug_umma_shift: null pointer");
    FX_CUDA(e, cudaSetDevice(e->device));
    int rc = tc_umma_shift_probe(e, a_dev, b_dev, kb_elems, shift_rows, base_offset, out_dev, nullptr);
    if (rc != FX_OK) return rc;
    FX_CUDA(e, cudaDeviceSynchronize());
    return FX_OK;
}

int fx_debug_mma_rate(fx_handle e, int n_cols, int rowb, int shift_rows, int tap_stride_rows, int iters, float* out_dev) {
    if (!e) return FX_ERR_INVALID;
    if (!out_dev || iters < 1) return set_error(e, FX_ERR_INVALID, "fx_debug_mma_rate: bad arguments");
    FX_CUDA(e, cudaSetDevice(e->device));
    int rc = tc_mma_rate(e, n_cols, rowb, shift_rows, tap_stride_rows, iters, out_dev, nullptr);
    if (rc != FX_OK) return rc;
    FX_CUDA(e, cudaDeviceSynchronize());
    return FX_OK;
}

// Host-side pieces of the preprocess, exposed so that CPU-only tests can check them without a GPU.
int fx_host_resized_size(int h, int w, int* oh, int* ow) {
    if (!oh || !ow || h < 1 || w < 1) return FX_ERR_INVALID;
    resized_size(h, w, *oh, *ow);
    return FX_OK;
}
int fx_host_crop_offset(int size) { return crop_offset(size); }
int fx_host_coeffs(int in_size, int out_size, int32_t* xmin, int32_t* count, int32_t* taps, int taps_capacity) {
    if (in_size < 1 || out_size < 1) return FX_ERR_INVALID;
    std::vector<int32_t> mn, ct, kk;
    int ks = 0;
    pillow_coeffs(in_size, out_size, mn, ct, kk, ks);
    if (!xmin || !count || !taps) return ks;
    if (taps_capacity < ks * out_size) return FX_ERR_INVALID;
    std::memcpy(xmin, mn.data(), sizeof(int32_t) * out_size);
    std::memcpy(count, ct.data(), sizeof(int32_t) * out_size);
    std::memcpy(taps, kk.data(), sizeof(int32_t) * kk.size());
    return ks;
}

}  // extern "C"
