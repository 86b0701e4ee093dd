/*
 * fx_b200.h -- C ABI of the B200-native feature-extraction hot path.
 *
 * The reference (Septimus4/semi-supervised-image-processing) is pure Python and has no FFI of its
 * own; the boundary it offers is the module surface of src/feature_extraction.py.  Each entry
 * point below names the reference lines it stands in for.  Everything is `extern "C"`, plain
 * pointers and sizes; no torch / C++ types cross the boundary.
 *
 * Conventions
 *   - every call returns FX_OK (0) or a negative fx_status; nothing throws across the ABI;
 *     fx_last_error(h) gives the text of the last failure on that handle (fx_last_error(NULL)
 *     for a failed fx_create);
 *   - one handle per GPU, not thread-safe per handle, re-entrant across handles;
 *   - "dev" pointers are device memory on the handle's GPU, "host" pointers are host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); device
 *     entry points are asynchronous on it, the *_host entry points synchronise before returning;
 *   - the caller owns every buffer it passes; the library owns its workspace and TMA descriptors.
 *   - there is no CPU fallback: without a usable sm_100 device fx_create fails.
 */
#ifndef FX_B200_H
#define FX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FX_ABI_VERSION 1

typedef struct fx_engine *fx_handle;

typedef enum fx_status {
    FX_OK = 0,
    FX_ERR_INVALID = -1,     /* bad argument (null pointer, n > max_batch, bad layer table ...) */
    FX_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed; text in fx_last_error       */
    FX_ERR_UNSUPPORTED = -3, /* wrong channel count / image too small or too large / no sm_100  */
    FX_ERR_NOMEM = -4,
    FX_ERR_STATE = -5        /* call order: weights not loaded, nothing staged, ...              */
} fx_status;

/* Arithmetic of the trunk.  BF16 = tcgen05 tensor-core path (bf16 operands, fp32 accumulate);
 * FP32 = the tight-tolerance mode: fp32-accurate results (relative L2 <= 1e-5 end to end against the reference's fp32
 * path) from split fp16 operands on the tensor cores with round-to-nearest accumulation in registers (csrc/conv_split.cu);
 * the environment variable FX_TIGHT_SIMT=1 at fx_create selects the plain fp32 CUDA-core kernel instead. */
typedef enum fx_precision { FX_PRECISION_BF16 = 0, FX_PRECISION_FP32 = 1 } fx_precision;

/* Geometry constants of the path: src/feature_extraction.py:64-67. */
#define FX_RESIZE 256
#define FX_CROP 224
#define FX_EMBED_DIM 512
#define FX_NUM_CONV_LAYERS 20
#define FX_MAX_LANES 2
#define FX_HOST_SLOTS 4
#define FX_MAX_CLASSES 32

/* Geometry of the fused preprocess (fx_set_transform). */
#define FX_TRANSFORM_EXTRACT 0   /* Resize(256) -> CenterCrop(224): src/feature_extraction.py:200-207 (default) */
#define FX_TRANSFORM_SQUARE224 1 /* Resize((224,224)), no crop: eval transform of src/training/common.py:111-117 */

/*
 * One decoded image inside a packed uint8 buffer, exactly what PIL hands the reference's
 * transform in preprocess_image (src/feature_extraction.py:233-240): HWC, 8 bits per channel,
 * row pitch = width*channels, no mode conversion.
 * channels == 3: mode "RGB" (the only mode the reference's Normalize accepts).
 * channels == 1: opt-in "gray carriage" of a file whose three stored channels are identical
 *                (R==G==B); the plane is replicated and each replica gets its own mean/std, which
 *                is arithmetically what the reference computes on such a file.  The Python host
 *                only uses it when asked to; a true mode-"L" file still raises as the reference.
 */
typedef struct fx_image_desc {
    uint64_t offset;  /* byte offset of pixel (0,0,0) from the `src` pointer of the call */
    int32_t height;
    int32_t width;
    int32_t channels; /* 3 or 1 */
    int32_t reserved;
} fx_image_desc;

/*
 * One convolution + its BatchNorm, as stored in a torchvision ResNet state_dict
 * (torchvision/models/resnet.py:59-105,197-243): conv weight OIHW fp32 (bias=False), BN affine
 * and running statistics.  The library folds BN into the conv in fp64
 * (s = gamma/sqrt(var+eps), W' = W*s, b' = beta - mean*s), then packs W' for its kernels.
 * All pointers are HOST memory and only read during fx_load_weights.
 */
typedef struct fx_conv_bn {
    const float *weight; /* [cout][cin][kh][kw] */
    const float *gamma;  /* [cout] */
    const float *beta;
    const float *mean;
    const float *var;
    float eps; /* 1e-5 */
    int32_t cout, cin, kh, kw, stride, pad;
} fx_conv_bn;

/* Layer order expected by fx_load_weights (children()[:-1] of resnet18, i.e. what load_model
 * keeps at src/feature_extraction.py:219-225):
 *  0 conv1
 *  1 layer1.0.conv1   2 layer1.0.conv2   3 layer1.1.conv1   4 layer1.1.conv2
 *  5 layer2.0.conv1   6 layer2.0.conv2   7 layer2.0.downsample   8 layer2.1.conv1   9 layer2.1.conv2
 * 10 layer3.0.conv1  11 layer3.0.conv2  12 layer3.0.downsample  13 layer3.1.conv1  14 layer3.1.conv2
 * 15 layer4.0.conv1  16 layer4.0.conv2  17 layer4.0.downsample  18 layer4.1.conv1  19 layer4.1.conv2
 */

const char *fx_version(void);
int fx_abi_version(void);

/* Text of the last error on `h` (or of the last failed fx_create when h == NULL). */
const char *fx_last_error(fx_handle h);

/*
 * Replaces load_model(device) (src/feature_extraction.py:210-227) together with fx_load_weights:
 * binds a GPU, allocates the activation workspace for up to `max_batch` images per call.
 */
int fx_create(fx_handle *out, int device, int max_batch, int precision);
void fx_destroy(fx_handle h);

/* The frozen trunk's parameters: resnet18 minus fc (src/feature_extraction.py:217-225). */
int fx_load_weights(fx_handle h, const fx_conv_bn *layers, int n_layers);

/*
 * Replaces build_transform()/preprocess_image for a batch
 * (src/feature_extraction.py:184-207,233-240): Resize(256) with Pillow's fixed-point antialiased
 * bilinear, CenterCrop(224), ToTensor, Normalize -- one fused kernel.
 * Writes exactly the tensor the reference stacks at src/feature_extraction.py:289:
 * fp32 NCHW [n][3][224][224], bit-identical to the reference transform.
 */
int fx_preprocess_nchw_f32(fx_handle h, const uint8_t *src_dev, const fx_image_desc *descs_host, int n,
                           float *out_dev, void *stream);

/*
 * Same transform, written into the engine's conv1 staging buffer (channel-padded, spatially
 * padded NHWC in the trunk's precision).  Follow with fx_forward.
 */
int fx_preprocess(fx_handle h, const uint8_t *src_dev, const fx_image_desc *descs_host, int n, void *stream);

/*
 * Replaces `model(batch_tensor)` + torch.flatten (src/feature_extraction.py:290-293) on the `n`
 * images staged by the last fx_preprocess / fx_stage_nchw_f32: writes fp32 [n][512] at emb_dev
 * (which may point into a larger gather buffer).
 */
int fx_forward(fx_handle h, int n, float *emb_dev, void *stream);

/* Stage an already normalised fp32 NCHW [n][3][224][224] batch (the reference's own
 * batch_tensor) instead of running the preprocess kernel; lets the trunk be checked alone. */
int fx_stage_nchw_f32(fx_handle h, const float *in_dev, int n, void *stream);

/*
 * Lanes.  A handle owns FX_MAX_LANES independent sets of staging / activation buffers (the weights are shared).
 * fx_select_lane makes `lane` the target of the following fx_preprocess / fx_stage_nchw_f32 / fx_forward / fx_embed
 * calls on this handle.  Batches queued on DIFFERENT lanes and DIFFERENT streams are independent and overlap on the
 * GPU: every trunk kernel is a persistent one-CTA-per-SM grid, so the other lane's kernels fill the SMs that the tail
 * of each kernel (and the launch gap behind it) leaves idle -- about +7 % images/s at batch 256.  The reference loop
 * (src/feature_extraction.py:272-300) is strictly serial; results do not depend on the lane.  Lane 0 exists from
 * fx_create, lane 1 is allocated on first selection.  The host-buffer slots below use lane == slot by themselves.
 */
int fx_select_lane(fx_handle h, int lane);

/*
 * Which deterministic transform the preprocess kernels apply from now on (all lanes).  Both end in ToTensor +
 * Normalize with the ImageNet constants and use Pillow's fixed-point antialiased bilinear resampler, bit-exactly.
 * FX_TRANSFORM_SQUARE224 is the reference's evaluation transform for the classifier-head inference paths
 * (generate_pseudo_labels src/training/semi_supervised.py:44-72, evaluate_model src/training/common.py:439-506,
 * compute_probs src/threshold_sweep.py:21-38): each axis is resampled independently to 224, nothing is cropped.
 */
int fx_set_transform(fx_handle h, int transform);

/*
 * Classifier head of the fine-tuned model (create_model, src/training/common.py:299-304: resnet18 with
 * fc = nn.Linear(512, num_classes)): weight fp32 [num_classes][512], bias fp32 [num_classes] (host pointers).
 */
int fx_load_head(fx_handle h, const float *weight, const float *bias, int num_classes);

/*
 * Replaces `outputs = model(inputs); probs = torch.softmax(outputs, dim=1)` of the three inference loops named
 * above on the n images staged on the active lane: trunk -> emb_dev fp32 [n][512] (required, caller owned) ->
 * logits_dev fp32 [n][num_classes] and probs_dev fp32 [n][num_classes] (either may be NULL).  The head runs in
 * fp32 with a fixed summation order in either precision mode.
 */
int fx_classify(fx_handle h, int n, float *emb_dev, float *logits_dev, float *probs_dev, void *stream);

/* fx_preprocess + fx_forward. */
int fx_embed(fx_handle h, const uint8_t *src_dev, const fx_image_desc *descs_host, int n, float *emb_dev,
             void *stream);

/*
 * The whole of src/feature_extraction.py:289-294 for one batch with HOST buffers: host->device
 * copy of the packed uint8 images (total_bytes), preprocess, trunk, device->host copy of the
 * [n][512] fp32 embeddings.  Synchronous.  Pinned host memory makes the copies asynchronous
 * inside the call but is not required.
 */
int fx_embed_host(fx_handle h, const uint8_t *src_host, size_t total_bytes, const fx_image_desc *descs_host,
                  int n, float *emb_host);

/*
 * The same, pipelined over FX_HOST_SLOTS slots (0..3).  fx_embed_host_async queues the H2D copy of the batch on the
 * library's copy stream, then the kernels and the D2H of the embeddings on the compute stream of lane (slot % 2), and
 * returns; fx_embed_host_wait(slot) blocks until that slot's embeddings are in emb_host.  Cycling through the slots
 * keeps two batches computing (one per lane, overlapping each other's kernel tails) while the copies of the next
 * ones are already in flight.  src_host / emb_host must stay valid (and should be pinned) until the wait returns.
 * fx_embed_host == async on slot 0 + wait.
 */
int fx_embed_host_async(fx_handle h, int slot, const uint8_t *src_host, size_t total_bytes,
                        const fx_image_desc *descs_host, int n, float *emb_host);
int fx_embed_host_wait(fx_handle h, int slot);

/*
 * fx_embed_host_async with the embeddings left ON THE DEVICE: the trunk's last kernel writes the n rows straight to
 * emb_dev, no device->host copy.  Multi-GPU extraction points emb_dev into the rank's slot of the all-gather buffer
 * (buf + (rank * rows_per_rank + rows_done) * 512), so that ONE in-place ncclAllGather assembles [N,512] and only the
 * rank that writes the artifacts copies it to the host (replaces the per-batch `.cpu().numpy()` of
 * src/feature_extraction.py:293-294 and the concatenation at :305).  Pair with fx_embed_host_wait(slot).
 */
int fx_embed_host_async_dev(fx_handle h, int slot, const uint8_t *src_host, size_t total_bytes,
                            const fx_image_desc *descs_host, int n, float *emb_dev);

/*
 * ---- GPU JPEG decode feeding the preprocess kernel ----
 * Replaces `Image.open(path)` + Pillow's libjpeg decode inside the reference's serial loop
 * (src/feature_extraction.py:238, 276-284) for baseline 8-bit three-component JPEGs with a complete bitstream: nvJPEG
 * (hardware JPEG engines when available, else its CUDA decoder; loaded with dlopen at fx_jpeg_init) decodes a whole batch
 * straight into the device image buffer the preprocess kernel reads.  Every other file is reported as FX_FILE_HOST_DECODE
 * and must be decoded by the caller with the reference's own Pillow call, which keeps the per-file
 * (UnidentifiedImageError, OSError) semantics of :281-284.  NOT bit-exact against libjpeg-turbo (IDCT / chroma
 * up-sampling differ; measured in profiles/r02_nvjpeg_tolerance.md): opt-in, the host decode pool is the default.
 */
#define FX_JPEG_BACKEND_AUTO 0     /* hardware engines if the handle can be created for them, else the CUDA decoder */
#define FX_JPEG_BACKEND_HARDWARE 1 /* NVJPEG_BACKEND_HARDWARE */
#define FX_JPEG_BACKEND_GPU 2      /* NVJPEG_BACKEND_GPU_HYBRID */

#define FX_FILE_GPU_JPEG 0    /* read into the slot's bitstream buffer, will be decoded on the GPU */
#define FX_FILE_HOST_DECODE 1 /* readable, but not a JPEG this path takes: the caller decodes it (or reports its failure) */
#define FX_FILE_UNREADABLE 2  /* stat/open/read failed: the caller's own open() reports the reference's error */

typedef struct fx_file_info {
    int32_t status; /* FX_FILE_* */
    int32_t height, width, components;
    int32_t subsampling; /* nvjpegChromaSubsampling_t value, -1 unknown */
    int32_t encoding;    /* SOF marker: 0xC0 baseline, 0xC2 progressive, ... */
    int32_t precision;   /* bits per sample */
    int32_t reserved;
    uint64_t offset, length; /* position / size of the bitstream in the slot's page-locked buffer; length 0 for files that do
                              * not start with a JPEG SOI marker (or exceed 256 MiB): those are not read at all */
} fx_file_info;

/* Load nvJPEG and create the decoder for `backend` (idempotent; FX_ERR_UNSUPPORTED when nvJPEG cannot be had). */
int fx_jpeg_init(fx_handle h, int backend);
/* FX_JPEG_BACKEND_HARDWARE / _GPU the decoder runs on, FX_JPEG_BACKEND_AUTO (0) before fx_jpeg_init. */
int fx_jpeg_backend(fx_handle h);
/* Frame header + eligibility of one bitstream in host memory (no decode). */
int fx_jpeg_probe(fx_handle h, const uint8_t *data, size_t length, fx_file_info *info);
/* Read the n files of one batch into slot `slot`'s page-locked bitstream buffer with a native thread pool
 * (FX_IO_THREADS, default 8) and fill info[i].  Waits for the slot's previous batch first. */
int fx_jpeg_read_files(fx_handle h, int slot, const char *const *paths, int n, fx_file_info *info);
/* Decode n bitstreams (host memory) to interleaved RGB at dst_dev + descs[i].offset (pitch = width * 3) on `stream`;
 * synchronises the stream before returning.  Test / study entry point. */
int fx_jpeg_decode(fx_handle h, const uint8_t *const *data, const size_t *lengths, int n, uint8_t *dst_dev,
                   const fx_image_desc *descs, void *stream);
/*
 * One pipelined step (as fx_embed_host_async / _dev) over the files last read into `slot` by fx_jpeg_read_files:
 * entries with info[i].status == FX_FILE_GPU_JPEG are decoded by nvJPEG, for the others host_pixels[i] holds the
 * caller-decoded HWC uint8 pixels; descs lays the decoded images out in the slot's device buffer (total_bytes).
 * `info` / `descs` list only the files that take part (failures already dropped), in row order.  Exactly one of
 * emb_host (rows copied back) / emb_dev (rows left on the device) is non-NULL.  FX_ERR_UNSUPPORTED when nvJPEG rejects
 * the batch: nothing was queued, re-submit the batch host-decoded.  Pair with fx_embed_host_wait(slot).
 */
int fx_embed_files_async(fx_handle h, int slot, const fx_file_info *info, const uint8_t *const *host_pixels,
                         const fx_image_desc *descs, int n, size_t total_bytes, float *emb_host, float *emb_dev);

/*
 * ---- on-device post-processing of an [n][d] fp32 embedding matrix ----
 * At N = 1M the reference's numpy / scikit-learn post-processing scans 2 GB on the host several times; these are the
 * same reductions as single HBM-bound passes over the (gathered) device buffer.  d <= 4096.  Column statistics
 * accumulate in fp64 with a fixed reduction shape (results do not depend on timing).
 */
typedef struct fx_matrix_stats {
    int64_t nan_count, inf_count; /* run_sanity_checks raises when either is non-zero (src/feature_extraction.py:337-340) */
    double mean_abs_mean;         /* np.abs(emb.mean(axis=0)).mean()   (:345) */
    double mean_std;              /* emb.std(axis=0).mean(), ddof = 0  (:346) */
} fx_matrix_stats;

/* run_sanity_checks (src/feature_extraction.py:334-356) + the fit of StandardScaler (src/standardize_features.py:41-43):
 * column mean / population std / variance as fp64 device arrays [d] (each may be NULL) and the scalars in *stats
 * (host, may be NULL; when given the call synchronises `stream`). */
int fx_column_stats(fx_handle h, const float *emb_dev, int64_t n, int d, double *col_mean_dev, double *col_std_dev,
                    double *col_var_dev, fx_matrix_stats *stats, void *stream);

/* StandardScaler.transform as scikit-learn (>= 1.3; 1.9 installed) executes it on a float32 matrix: mean and scale
 * rounded to fp32, z = (x - mean) / scale in fp32 IEEE arithmetic, bit for bit.  scale_dev is the caller's (std with
 * near-constant columns replaced by 1, sklearn's _is_constant_feature rule). */
int fx_standardize(fx_handle h, const float *emb_dev, int64_t n, int d, const double *col_mean_dev, const double *col_scale_dev,
                   float *out_dev, void *stream);

/* nearest_neighbor_probe (src/feature_extraction.py:359-398): for each of the q (<= 32) query rows, the row with the
 * largest cosine similarity, the query itself excluded, first maximum on ties; fp32 arithmetic.  Host index / output
 * arrays; synchronises `stream`. */
int fx_neighbor_probe(fx_handle h, const float *emb_dev, int64_t n, int d, const int64_t *query_rows_host, int q,
                      int64_t *neighbor_rows_host, float *similarity_host, void *stream);

/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
uint64_t fx_launch_count(fx_handle h);
/* Bytes the host-buffer entry points (fx_embed_host*) have copied host -> device since creation.  A uniform batch is
 * copied as one 2-D copy of the source rows the crop can touch -- the transform never reads the rest (bench.py's
 * e2e.h2d_bytes_per_step is counted here, not assumed). */
uint64_t fx_h2d_bytes(fx_handle h);

/*
 * Per-launch timing of fx_forward for roofline reporting: with profiling enabled every launch of the
 * trunk is bracketed by CUDA events on the launching stream.  fx_profile_read waits for the last
 * fx_forward and returns milliseconds for slots 0..19 = the conv groups in fx_load_weights order
 * (slot 0 = fused stem incl. max-pool) and slot 20 = global average pool.  capacity >= 21.
 */
int fx_profile_enable(fx_handle h, int on);
int fx_profile_read(fx_handle h, float *ms, int capacity);

/* ---- test / inspection entry points (used by tests/ for per-layer parity) ---- */

/*
 * Fold, pack and run ONE conv+bn group (any 3x3 / 1x1 / strided shape the trunk kernels support,
 * or the 7x7 stem) on a caller supplied activation: in_dev is fp32 NHWC [n][hin][win][cin];
 * residual_dev (may be NULL) and out_dev are fp32 NHWC [n][ho][wo][cout].  In BF16 precision the
 * input/residual are rounded to bf16, the tcgen05 kernel runs, and its bf16 result is widened back
 * to fp32.  Synchronises `stream` before returning.
 */
int fx_debug_conv(fx_handle h, const fx_conv_bn *layer, int hin, int win, const float *in_dev,
                  const float *residual_dev, int n, int relu /* bit 0: ReLU; bit 1 (BF16 engines, cin % 64 == 0): write the
                  fp32 accumulator itself instead of its bf16 rounding, through the per-tap kernel */,
                  float *out_dev, void *stream);

/* The fused stem: conv1 7x7/s2 + folded bn1 + ReLU + 3x3/s2/p1 max-pool in ONE tcgen05 kernel
 * (torchvision/models/resnet.py:197-200,268-271).  in_dev fp32 NHWC [n][224][224][3] (normalised),
 * out_dev fp32 NHWC [n][56][56][64].  BF16 engines only.  Synchronises `stream`. */
int fx_debug_stem_pool(fx_handle h, const fx_conv_bn *layer, const float *in_dev, int n, float *out_dev, void *stream);

/* Copy the raw conv1 staging buffer of the first n staged images to out_dev (device pointer, `bytes` must be
 * n * per-image size: bf16 [115][116][16] space-to-depth = 426,880 B, fp32 [230][232][4] = 853,760 B).  Lets
 * tests compare what fx_preprocess hands the trunk with the oracle, padding included.  Synchronises `stream`. */
int fx_debug_staging(fx_handle h, int n, void *out_dev, size_t bytes, void *stream);

/* Copy out the folded parameters of loaded layer `layer` as the kernels see them (fp32,
 * [cout][kh][kw][cin] order, bf16-rounded in BF16 precision) -- host pointers, either may be NULL. */
int fx_debug_folded(fx_handle h, int layer, float *weight_host, float *bias_host);

/* One 4-D bf16 TMA box load (tensor map built from the arguments) -> raw shared-memory bytes
 * copied to out_dev.  Pins the TMA behaviours the conv kernel relies on (zero fill out of bounds,
 * element strides, 16-byte-strided overlapping windows, swizzle patterns). */
int fx_debug_tma_probe(fx_handle h, const void *base_dev, const uint64_t *dims, const uint64_t *strides_bytes,
                       const uint32_t *box, const uint32_t *elem_strides, int swizzle, const int *coords,
                       int bytes, uint8_t *out_dev);

/* One tcgen05.mma chain D[128][64] = A[shift .. shift+128) * B^T (fp32 out) where A is a 256-row,
 * kb_elems-wide (64/32/16 -> SWIZZLE_128B/64B/32B) K-major bf16 tile loaded by one TMA box and the A
 * descriptor starts `shift_rows` rows into it.  Pins the shifted-view behaviour the halo-tile conv
 * kernels rely on.  b_dev is [64][kb_elems] bf16. */
int fx_debug_umma_shift(fx_handle h, const void *a_dev, const void *b_dev, int kb_elems, int shift_rows,
                        int base_offset, float *out_dev);

/* tcgen05.mma issue-rate microbenchmark: every SM issues iters x (rowb/32) MMAs M128 x n_cols x K16 from
 * zeroed shared-memory tiles, the A view starting shift_rows (+ k*tap_stride_rows, k = 0..8) rows into
 * the tile.  out_dev[sm_count] <- SM cycles per MMA.  Evidence for DESIGN.md's operand-fetch floors. */
int fx_debug_mma_rate(fx_handle h, int n_cols, int rowb, int shift_rows, int tap_stride_rows, int iters, float *out_dev);

/* Host-side integer pieces of the preprocess (no GPU needed): torchvision's Resize(256) output
 * size, CenterCrop(224)'s round-half-even offset, Pillow's fixed-point coefficient table
 * (returns taps per output sample; with NULL arrays only returns that count). */
int fx_host_resized_size(int h, int w, int *oh, int *ow);
int fx_host_crop_offset(int size);
int fx_host_coeffs(int in_size, int out_size, int32_t *xmin, int32_t *count, int32_t *taps, int taps_capacity);

#ifdef __cplusplus
}
#endif
#endif /* FX_B200_H */
