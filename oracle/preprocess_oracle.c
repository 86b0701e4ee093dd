/*
 * preprocess_oracle.c -- CPU restatement of the reference's deterministic preprocess.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Path restated (reference = /root/reference):
 *   src/feature_extraction.py:200-207  Compose[Resize(256), CenterCrop(224), ToTensor, Normalize]
 *   src/feature_extraction.py:233-240  preprocess_image: the PIL image goes in as decoded (no .convert)
 * The arithmetic itself lives in two third-party packages that are NOT under /root/reference
 * (pyproject.toml:11,15 give lower bounds only; installed: torchvision 0.26.0, Pillow 12.2.0):
 *   torchvision/transforms/functional.py:353-384  _compute_resized_output_size (short side -> 256)
 *   torchvision/transforms/functional.py:470-477  identity shortcut, PIL dispatch
 *   torchvision/transforms/_functional_pil.py:242-253 -> PIL/Image.py:2328-2437  Image.resize(BILINEAR)
 *   Pillow src/libImaging/Resample.c (not on disk; published algorithm restated below:
 *       precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
 *       ImagingResampleVertical_8bpc -- 22-bit fixed point, horizontal pass first,
 *       uint8 intermediate)
 *   torchvision/transforms/functional.py:556-594  center_crop (Python round-half-even offsets)
 *   torchvision/transforms/functional.py:166-178  to_tensor   (u8 -> f32, /255)
 *   torchvision/transforms/_functional_tensor.py:916-928 normalize ((x-mean)/std in f32)
 *
 * Pinned by tests/test_oracle_cpu.py against the installed Pillow/torchvision and against the
 * golden vectors minted from the real reference module (tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; double arithmetic must not be fused).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FXO_PRECISION_BITS 22 /* 32 - 8 - 2, Resample.c */

static const float FXO_MEAN[3] = {0.485f, 0.456f, 0.406f}; /* src/feature_extraction.py:64 */
static const float FXO_STD[3] = {0.229f, 0.224f, 0.225f};  /* src/feature_extraction.py:65 */

/* Triangle filter, support 1.0 (Pillow bilinear_filter). */
static double fxo_tri(double x) {
    if (x < 0.0) x = -x;
    return x < 1.0 ? 1.0 - x : 0.0;
}

/* Max taps per output sample for one axis. */
int fxo_ksize(int in_size, int out_size) {
    double scale = (double)in_size / (double)out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(1.0 * fs) * 2 + 1;
}

/*
 * One axis of Pillow's coefficient table.  bounds[2*xx] = first input index, bounds[2*xx+1] = tap
 * count; kk[xx*ksize + i] = 22-bit fixed-point weight.  Returns ksize.
 */
int fxo_coeffs(int in_size, int out_size, int32_t *bounds, int32_t *kk) {
    double scale = (double)in_size / (double)out_size;
    double fs = scale < 1.0 ? 1.0 : scale; /* antialias only when shrinking */
    double support = 1.0 * fs;
    int ksize = (int)ceil(support) * 2 + 1;
    double ss = 1.0 / fs;
    double *w = (double *)malloc(sizeof(double) * (size_t)ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        int n = xmax - xmin;
        double ww = 0.0;
        for (int x = 0; x < n; ++x) {
            w[x] = fxo_tri((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < n; ++x) {
            if (ww != 0.0) w[x] /= ww;
        }
        int32_t *k = kk + (size_t)xx * ksize;
        for (int x = 0; x < n; ++x) {
            double v = w[x] * (double)(1 << FXO_PRECISION_BITS);
            k[x] = v < 0 ? (int32_t)(-0.5 + v) : (int32_t)(0.5 + v);
        }
        for (int x = n; x < ksize; ++x) k[x] = 0;
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = n;
    }
    free(w);
    return ksize;
}

static uint8_t fxo_clip8(int32_t acc) {
    int32_t v = acc >> FXO_PRECISION_BITS; /* arithmetic shift on a signed 32-bit accumulator */
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* torchvision _compute_resized_output_size for an int size (short side -> target). */
void fxo_resized_size(int h, int w, int target, int *oh, int *ow) {
    int short_side = w <= h ? w : h, long_side = w <= h ? h : w;
    int new_long = (int)((double)((long long)target * long_side) / (double)short_side);
    if (w <= h) { *ow = target; *oh = new_long; } else { *ow = new_long; *oh = target; }
}

/* Python int(round(x / 2.0)) for a non-negative integer numerator: round half to even. */
int fxo_crop_offset(int size, int crop) {
    int d = size - crop; /* >= 0 after Resize(256) */
    int q = d / 2;
    if (d % 2 == 0) return q;
    return (q % 2 == 0) ? q : q + 1; /* q + 0.5 -> even neighbour */
}

/* Pillow Image.resize((ow, oh), BILINEAR) on an HWC uint8 image; dst is [oh][ow][c]. */
int fxo_resize_bilinear_u8(const uint8_t *src, int h, int w, int c, uint8_t *dst, int oh, int ow) {
    const uint8_t *hsrc = src;
    uint8_t *tmp = NULL;
    if (ow != w) { /* horizontal pass first, rounded to uint8 */
        int ks = fxo_ksize(w, ow);
        int32_t *b = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)ow);
        int32_t *k = (int32_t *)malloc(sizeof(int32_t) * (size_t)ow * ks);
        fxo_coeffs(w, ow, b, k);
        tmp = (uint8_t *)malloc((size_t)h * ow * c);
        for (int y = 0; y < h; ++y)
            for (int xx = 0; xx < ow; ++xx) {
                int xmin = b[2 * xx], n = b[2 * xx + 1];
                for (int ch = 0; ch < c; ++ch) {
                    int32_t acc = 1 << (FXO_PRECISION_BITS - 1);
                    for (int i = 0; i < n; ++i)
                        acc += k[(size_t)xx * ks + i] * (int32_t)src[((size_t)y * w + xmin + i) * c + ch];
                    tmp[((size_t)y * ow + xx) * c + ch] = fxo_clip8(acc);
                }
            }
        free(b); free(k);
        hsrc = tmp;
    }
    if (oh != h) { /* vertical pass on the uint8 intermediate */
        int ks = fxo_ksize(h, oh);
        int32_t *b = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)oh);
        int32_t *k = (int32_t *)malloc(sizeof(int32_t) * (size_t)oh * ks);
        fxo_coeffs(h, oh, b, k);
        size_t row = (size_t)ow * c;
        for (int yy = 0; yy < oh; ++yy) {
            int ymin = b[2 * yy], n = b[2 * yy + 1];
            for (size_t x = 0; x < row; ++x) {
                int32_t acc = 1 << (FXO_PRECISION_BITS - 1);
                for (int i = 0; i < n; ++i)
                    acc += k[(size_t)yy * ks + i] * (int32_t)hsrc[(size_t)(ymin + i) * row + x];
                dst[(size_t)yy * row + x] = fxo_clip8(acc);
            }
        }
        free(b); free(k);
    } else {
        memcpy(dst, hsrc, (size_t)oh * ow * c);
    }
    free(tmp);
    return 0;
}

/* ToTensor + Normalize as a per-channel table of the 256 byte values (fp32, true division). */
void fxo_build_lut(float *lut /* [3][256] */) {
    for (int ch = 0; ch < 3; ++ch)
        for (int v = 0; v < 256; ++v) {
            volatile float x = (float)v / 255.0f;
            volatile float y = x - FXO_MEAN[ch];
            lut[ch * 256 + v] = y / FXO_STD[ch];
        }
}

/*
 * Resize(256) -> CenterCrop(224) on an HWC uint8 image; out_u8 is [224][224][c].
 * Returns 0, or -1 if the resized image is smaller than the crop (cannot happen after Resize(256)).
 */
int fxo_resize_crop_u8(const uint8_t *src, int h, int w, int c, uint8_t *out_u8) {
    int oh, ow;
    fxo_resized_size(h, w, 256, &oh, &ow);
    if (oh < 224 || ow < 224) return -1;
    uint8_t *rs;
    int owned = 0;
    if (oh == h && ow == w) {
        rs = (uint8_t *)src; /* torchvision returns the image untouched */
    } else {
        rs = (uint8_t *)malloc((size_t)oh * ow * c);
        owned = 1;
        fxo_resize_bilinear_u8(src, h, w, c, rs, oh, ow);
    }
    int top = fxo_crop_offset(oh, 224), left = fxo_crop_offset(ow, 224);
    for (int y = 0; y < 224; ++y)
        memcpy(out_u8 + (size_t)y * 224 * c, rs + ((size_t)(y + top) * ow + left) * c, (size_t)224 * c);
    if (owned) free(rs);
    return 0;
}

/*
 * Whole transform for one 3-channel HWC uint8 image -> fp32 CHW [3][224][224], exactly what
 * build_transform() returns for a mode-"RGB" PIL image.  c must be 3: the reference raises on
 * any other channel count, and so does the product.
 */
int fxo_preprocess_rgb(const uint8_t *src, int h, int w, float *out_chw) {
    static float lut[3 * 256];
    static int lut_ready = 0;
    if (!lut_ready) { fxo_build_lut(lut); lut_ready = 1; }
    uint8_t *crop = (uint8_t *)malloc((size_t)224 * 224 * 3);
    int rc = fxo_resize_crop_u8(src, h, w, 3, crop);
    if (rc == 0)
        for (int ch = 0; ch < 3; ++ch)
            for (int i = 0; i < 224 * 224; ++i) out_chw[(size_t)ch * 224 * 224 + i] = lut[ch * 256 + crop[(size_t)i * 3 + ch]];
    free(crop);
    return rc;
}

/*
 * Evaluation transform of the classifier-head paths (src/training/common.py:111-117):
 * Resize((224, 224)) -- each axis resampled independently, no crop -- then ToTensor + Normalize.
 * torchvision hands (w, h) = (224, 224) to Image.resize (tv:transforms/_functional_pil.py:242-253), whose
 * passes are skipped per axis when the size already matches.  fp32 CHW [3][224][224].
 */
int fxo_preprocess_square224_rgb(const uint8_t *src, int h, int w, float *out_chw) {
    static float lut[3 * 256];
    static int lut_ready = 0;
    if (!lut_ready) { fxo_build_lut(lut); lut_ready = 1; }
    uint8_t *rs = (uint8_t *)malloc((size_t)224 * 224 * 3);
    fxo_resize_bilinear_u8(src, h, w, 3, rs, 224, 224);
    for (int ch = 0; ch < 3; ++ch)
        for (int i = 0; i < 224 * 224; ++i) out_chw[(size_t)ch * 224 * 224 + i] = lut[ch * 256 + rs[(size_t)i * 3 + ch]];
    free(rs);
    return 0;
}
