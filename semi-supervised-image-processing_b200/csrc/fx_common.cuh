// Shared declarations of the B200 feature-extraction library (internal; the public C ABI is
// include/fx_b200.h).  sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fx_b200.h"

namespace fx {

constexpr int kCrop = FX_CROP;      // 224
constexpr int kResize = FX_RESIZE;  // 256
constexpr int kEmbed = FX_EMBED_DIM;
constexpr int kNumLayers = FX_NUM_CONV_LAYERS;

// conv1 staging buffer: the normalised crop, channel-padded 3 -> 4 and spatially padded so that
// conv1 (7x7, stride 2, pad 3) needs no out-of-bounds handling: pixel (y, x) of the crop lives at
// [y + kIn0Pad][x + kIn0Pad].  230 x 232 x 4; the pad is zero (conv zero padding of the
// NORMALISED tensor, torchvision/models/resnet.py:197).
constexpr int kIn0Pad = 3;
constexpr int kIn0H = kCrop + 6;  // 230
constexpr int kIn0W = kCrop + 8;  // 232: 8-pixel windows starting at 2*ow stay inside the row
constexpr int kIn0C = 4;

// bf16 path: the stem reads the normalised crop in SPACE-TO-DEPTH form so that the 7x7/s2 conv is a
// 4x4/s1 conv the flat kernel can run (conv_flat.cu).  With the crop zero-padded by 3 on every side
// (230 x 230, conv padding of the NORMALISED tensor), s2d pixel (Y, X) holds padded pixels
// (2Y+dy, 2X+dx), dy, dx in {0,1}, as channel (dy*2+dx)*3 + c; channels 12..15 and column 115 are
// zero.  [n][115][116][16] bf16 = 426,880 B per image (same bytes as the fp32 path's layout / 2).
constexpr int kS2dH = 115;
constexpr int kS2dW = 116;
constexpr int kS2dC = 16;

struct LayerGeom {
    int cin, cout, kh, kw, stride, pad;
    int hin, win, hout, wout;  // logical activation sizes
};

// ---------------------------------------------------------------------------------------------
// Preprocess geometry: Pillow's coefficient tables for one (height, width), already composed with
// the centre crop.  Built on the host in double precision (preprocess.cu), cached per size.
// ---------------------------------------------------------------------------------------------
struct GeomTableHost {
    int h, w;
    int ksh, ksv;            // taps per output sample, horizontal / vertical
    int band;                // output rows per thread block
    int max_rows;            // source rows a band touches (upper bound)
    int col_lo, col_hi;      // source pixel columns touched by the crop [lo, hi)
    int row_lo, row_hi;      // source rows touched by the crop [lo, hi)
    int cnt_h, cnt_v;        // largest tap count actually used inside the crop, per axis
    int s2d_rows[2];         // source rows a band of the s2d kernel touches (upper bound), bands of 8 / 16 row pairs
    bool noclip;             // all taps >= 0 and sums small enough that clip8 never clips
    std::vector<int32_t> blob;  // packed device image, layout in preprocess.cu
};

struct GeomEntry {
    int32_t* dev = nullptr;  // device copy of blob
    int ksh = 0, ksv = 0, band = 0, max_rows = 0, col_lo = 0, col_hi = 0, cnt_h = 0, cnt_v = 0;
    int s2d_rows[2] = {0, 0};
    int row_lo = 0, row_hi = 0;  // source rows the crop touches: the only rows the kernels read (and the host path copies)
    bool noclip = false;
};

// Per-image record consumed by the preprocess kernel.
struct ImgDev {
    unsigned long long src_off;
    const int32_t* geom;
    int h, w, c;
    int ksh, ksv, band, col_lo, col_hi;
    int fast;  // 1: RGB image whose taps fit the register-resident fast path (ksh, ksv <= 5, band fits in smem)
};

struct PackedLayer {
    LayerGeom g{};
    // fp32 pack: [cout][kh][kw][cin_pad]  (cin_pad = 4 for conv1, else cin)
    float* w_f32 = nullptr;
    // bf16 pack: GEMM B matrix [cout][K], K ordered (kh, kw, cin) -- conv1: s2d form (4, 4, 16)
    __nv_bfloat16* w_bf16 = nullptr;
    float* bias = nullptr;
    int k_bf16 = 0;
    // tight mode on the tensor cores (conv_split.cu): folded weights * 2^w_scale_log2 as fp16 hi / lo planes, [cout][K]
    void* w_h16 = nullptr;
    void* w_l16 = nullptr;
    int w_scale_log2 = 0;
    std::vector<float> host_w;  // folded, [cout][kh][kw][cin] (bf16-rounded in BF16 precision)
    std::vector<float> host_b;
};

}  // namespace fx

struct fx_engine {
    int device = 0;
    int max_batch = 0;
    int precision = FX_PRECISION_BF16;
    int sm_count = 148;
    bool weights_loaded = false;
    int staged = 0;  // images currently in the conv1 staging buffer
    uint64_t launches = 0;
    uint64_t h2d_bytes = 0;  // bytes the host-buffer entry points have copied host -> device (fx_h2d_bytes)
    std::string err;

    // preprocess
    float* lut_f32 = nullptr;            // [3][256]
    __nv_bfloat16* lut_bf16 = nullptr;   // [3][256]
    std::map<std::pair<int, int>, fx::GeomEntry> geoms;  // key: (height + (transform << 24), width)
    // the coefficient tables live in ONE device arena (bump allocation): a new (height, width) costs a host-side table
    // build and an asynchronous copy, not a cudaMalloc (a device-wide synchronisation) -- ragged datasets bring hundreds
    // of new sizes per batch.  Dropped as a whole when full (preprocess.cu).
    uint8_t* geom_arena = nullptr;
    size_t geom_arena_cap = 0, geom_arena_used = 0;
    std::vector<int32_t*> geom_spill;  // tables that did not fit in what was left of the arena (plain allocations)
    cudaStream_t geom_stream = nullptr;  // every table upload goes through this stream ...
    cudaEvent_t geom_ready = nullptr;    // ... and re-records this event: a consumer stream that waits for it sees every table so far
    int transform = FX_TRANSFORM_EXTRACT;
    fx::ImgDev* img_dev = nullptr;       // [max_batch]
    fx::ImgDev* img_host = nullptr;      // pinned, [max_batch]
    cudaEvent_t img_host_free = nullptr; // img_host may be rewritten once this has fired
    float norm_a[3] = {}, norm_b[3] = {};  // Normalize o ToTensor as one FMA per value (preprocess.cu, NormFma)
    bool norm_fma_ok = false;            // ... proven equal to the bf16 table for all 768 inputs at start-up
    bool s2d_ppb16 = false;              // FX_DEBUG_S2D_PPB=16: bands of 16 row pairs instead of 8 (slower; measurement knob)
    bool pre_force_banded = false;       // FX_DEBUG_PRE_BANDED=1: measurement knob (preprocess.cu)

    // trunk
    fx::PackedLayer layers[fx::kNumLayers];
    void* in0 = nullptr;       // conv1 staging: fp32 [max_batch][230][232][4] or bf16 s2d [max_batch][115][116][16]
    void* act[3] = {nullptr, nullptr, nullptr};  // ping-pong activations (NHWC)
    float* final_f32 = nullptr;  // [max_batch][49][512]
    size_t act_bytes = 0;
    void* tc_state = nullptr;  // tcgen05 path: tensor maps etc. (conv_tc.cu)
    bool tight_tc = true;      // FP32 precision: split-fp16 tensor-core convs (conv_split.cu); FX_TIGHT_SIMT=1 keeps the CUDA-core kernel

    // per-launch timing of fx_forward (fx_profile_*): event pairs for the 20 convs + avgpool
    bool prof_on = false;
    cudaEvent_t prof_ev[2 * (FX_NUM_CONV_LAYERS + 1)] = {};

    // Lanes: independent sets of the per-batch buffers above (in0, act, final_f32, img_dev, img_host, img_host_free,
    // staged).  The fields above are the ACTIVE lane's; fx_select_lane swaps them with the stored copy.  Batches
    // queued on different lanes and different streams overlap on the GPU: every trunk kernel is a persistent
    // one-CTA-per-SM grid, so the CTAs of the other lane's next kernel fill the SMs that a kernel's tail (and the
    // launch gap behind it) would leave idle.  Lane 1 is allocated on first use.
    struct Lane {
        void* in0 = nullptr;
        void* act[3] = {nullptr, nullptr, nullptr};
        float* final_f32 = nullptr;
        fx::ImgDev* img_dev = nullptr;
        fx::ImgDev* img_host = nullptr;
        cudaEvent_t img_host_free = nullptr;
        int staged = 0;
        bool allocated = false;
    } lanes[FX_MAX_LANES];
    int cur_lane = 0;
    // preprocess launch plan of each lane's last batch; reused when the next batch has the same descriptor table
    struct PrePlan {
        bool valid = false, s2d = false;
        int mode = 0, transform = 0, n = 0, kernel = 0, bands = 0, smem = 0, tmp_bytes = 0, rowbuf = 0, ppb = 8;
        std::vector<fx_image_desc> descs;
        uint64_t serial = 0;  // bumped whenever the plan is rebuilt (CUDA-graph cache key)
    } pre_plan[FX_MAX_LANES];
    // CUDA graph of one whole step (preprocess + trunk) per lane, for launch-bound batch sizes: captured once the
    // lane's preprocess plan repeats, replayed with the two call-specific pointers (source images, embedding output)
    // patched into its first and last kernel node (engine.cu)
    struct StepGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t pre_node = nullptr, pool_node = nullptr;
        int n = 0, kernels = 0;
        uint64_t plan_serial = 0, weights_epoch = 0;
        const uint8_t* src = nullptr;
        float* emb = nullptr;
    } step_graph[FX_MAX_LANES];
    bool graphs_on = true;     // FX_GRAPHS=0 disables; switched off for the handle if a capture ever fails
    int graph_max_batch = 64;  // FX_GRAPH_MAX_BATCH: above this a step is GPU-bound and launched directly
    bool capturing = false;
    uint64_t weights_epoch = 0;

    // host-buffer path (fx_embed_host*): FX_HOST_SLOTS pipelined slots (device copies of one batch's packed images and
    // embeddings); slot s computes on lane s % FX_MAX_LANES, whose stream serialises the slots that share its buffers.
    // Four slots keep two batches computing (one per lane) while the H2D copies of the next two are already queued.
    struct HostSlot {
        uint8_t* src_dev = nullptr;
        size_t cap = 0;
        float* emb_dev = nullptr;
        cudaEvent_t copied = nullptr, done = nullptr;
        bool busy = false;
    } slots[FX_HOST_SLOTS];
    // on-device post-processing scratch (postprocess.cu)
    void* post_scratch = nullptr;
    size_t post_cap = 0;
    // classifier head (fx_load_head)
    float* head_w = nullptr;  // [head_classes][512]
    float* head_b = nullptr;
    int head_classes = 0;
    void* jpeg_state = nullptr;  // nvJPEG decode context (decode.cu), created by fx_jpeg_init
    cudaStream_t lane_stream[FX_MAX_LANES] = {};  // kernels + D2H of the host-buffer path, one per lane
    cudaStream_t copy_stream = nullptr;           // H2D of the host-buffer path
};

namespace fx {

int set_error(fx_engine* e, int code, const std::string& msg);
const char* cuda_err_text(cudaError_t err);

#define FX_CUDA(e, call)                                                                             \
    do {                                                                                             \
        cudaError_t _err = (call);                                                                   \
        if (_err != cudaSuccess)                                                                     \
            return fx::set_error((e), FX_ERR_CUDA,                                                   \
                                 std::string(#call) + ": " + cudaGetErrorString(_err));              \
    } while (0)

#define FX_LAUNCH_CHECK(e, name)                                                                     \
    do {                                                                                             \
        cudaError_t _err = cudaGetLastError();                                                       \
        if (_err != cudaSuccess)                                                                     \
            return fx::set_error((e), FX_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(_err)); \
        (e)->launches++;                                                                             \
    } while (0)

// Programmatic dependent launch: the kernels of a step are launched with the stream-serialization attribute; each
// signals `launch_dependents` at once and executes `griddepcontrol.wait` after its set-up (barrier init, TMEM
// allocation, resident-weight loads -- nothing that depends on, or could disturb, an earlier kernel) and BEFORE its
// first dependent global read or any global write.  The next kernel's prologue then overlaps this kernel's tail.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// preprocess.cu
enum class PreOut : int { NCHW_F32 = 0, IN0_BF16 = 1, IN0_F32 = 2 };
int preprocess_init(fx_engine* e);
void preprocess_free(fx_engine* e);
int preprocess_lane_init(fx_engine* e);  // per-lane descriptor buffers -> the active lane's fields
// true if `descs` repeats the active lane's previous batch (no upload needed); *kernel = the kernel that would launch
bool preprocess_plan_hit(fx_engine* e, const fx_image_desc* descs, int n, PreOut mode, const void** kernel);
// source rows [lo, hi) of an h x w image that the current transform's crop touches (geometry cache; FX_OK or an error)
int preprocess_rows_needed(fx_engine* e, int h, int w, int* row_lo, int* row_hi);
const void* avgpool_kernel_ptr(bool in_is_bf16);
int preprocess_run(fx_engine* e, const uint8_t* src_dev, const fx_image_desc* descs, int n, PreOut mode,
                   void* out, cudaStream_t stream);
int stage_nchw_run(fx_engine* e, const float* in_dev, int n, cudaStream_t stream);
// host-only helpers (also exported for tests through the debug ABI)
void pillow_coeffs(int in_size, int out_size, std::vector<int32_t>& xmin, std::vector<int32_t>& cnt,
                   std::vector<int32_t>& kk, int& ksize);
void resized_size(int h, int w, int& oh, int& ow);
int crop_offset(int size);

// trunk_simt.cu (fp32 tight-tolerance path + pooling / layout kernels shared by both paths)
int simt_conv(fx_engine* e, const PackedLayer& L, const float* in, int hin_phys, int win_phys, int cin_phys,
              int pad, const float* residual, float* out, int n, int relu, cudaStream_t stream);
int maxpool_3x3s2(fx_engine* e, const void* in, void* out, int n, int h, int w, int c, bool bf16,
                  cudaStream_t stream);
int head_run(fx_engine* e, const float* emb, int n, float* logits, float* probs, cudaStream_t stream);
int avgpool_7x7(fx_engine* e, const void* in, bool in_is_bf16, float* out, int n, int hw, int c,
                cudaStream_t stream);
int f32_to_bf16(fx_engine* e, const float* in, __nv_bfloat16* out, size_t count, cudaStream_t stream);
int bf16_to_f32(fx_engine* e, const __nv_bfloat16* in, float* out, size_t count, cudaStream_t stream);
int pad_nhwc3_to_in0(fx_engine* e, const float* in_nhwc3, void* in0, bool bf16, int n, cudaStream_t stream);

// postprocess.cu (sanity statistics, StandardScaler, nearest-neighbour probe on the device)
int post_column_stats(fx_engine* e, const float* x, long long n, int d, double* mean_dev, double* std_dev, double* var_dev,
                      fx_matrix_stats* out, cudaStream_t stream);
int post_standardize(fx_engine* e, const float* x, long long n, int d, const double* mean_dev, const double* scale_dev, float* out,
                     cudaStream_t stream);
int post_neighbor_probe(fx_engine* e, const float* x, long long n, int d, const int64_t* qidx_host, int q, int64_t* nbr_host,
                        float* sim_host, cudaStream_t stream);

// decode.cu (GPU JPEG decode feeding the preprocess kernel)
void jpeg_free(fx_engine* e);

// conv_tc.cu (bf16 tcgen05 implicit-GEMM path)
int tc_init(fx_engine* e);
void tc_free(fx_engine* e);
// Runs layer `li` on NHWC bf16 activations.  out_f32 != nullptr: write fp32 instead of bf16.
int tc_conv(fx_engine* e, int li, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
            float* out_f32, int n, int relu, cudaStream_t stream);

// conv_split.cu (tight-tolerance mode on the tensor cores: split fp16 operands, register accumulation)
int split_conv(fx_engine* e, const PackedLayer& L, const void* in, const void* residual, void* out, float* out_f32, int n, int relu,
               cudaStream_t stream);
int f32_to_split(fx_engine* e, const float* in, void* out_split, size_t count, cudaStream_t stream);
int split_to_f32(fx_engine* e, const void* in_split, float* out, size_t count, cudaStream_t stream);
int split_pack_weights(const std::vector<float>& w, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo);
void split_stem_s2d_weights(const std::vector<float>& host_w, int cout, std::vector<float>& out);
size_t split_stem_scratch_bytes(int n);
int split_stem(fx_engine* e, const PackedLayer& L, const float* in0_f32, void* xs2d_split, void* conv_out_split, void* pooled_split, int n,
               cudaStream_t stream);

// conv_flat.cu (bf16 tcgen05 weight-stationary halo-tile path: stem, layer1, layer2 3x3/s1)
bool flat_supported(const LayerGeom& g);
// pool (stem only): also apply the 3x3/s2/p1 max-pool in the epilogue; out is then [n][56][56][64].
// sched: tile walk of the stem / layer-1 kernels (conv_flat.cu TileWalk: bit 0 descending, bit 1 strided over the CTAs).
int flat_conv(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out, int n,
              int relu, bool pool, cudaStream_t stream, int sched = 0);

}  // namespace fx
