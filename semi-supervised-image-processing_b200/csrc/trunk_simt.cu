// CUDA-core kernels of the trunk:
//   * simt_conv: fp32 implicit-GEMM convolution (+folded-BN bias, +residual, +ReLU) on NHWC
//     activations -- the FX_PRECISION_FP32 "tight tolerance" mode.  True fp32 products and
//     accumulation (no TF32), so it tracks the reference's CPU fp32 path
//     (torchvision/models/resnet.py:89-105,266-282 run by src/feature_extraction.py:290-291) to
//     ~1e-6 relative L2.
//   * maxpool 3x3/s2/p1 (resnet.py:200), global average pool (resnet.py:206) and layout helpers,
//     shared by both precisions.  All are HBM/L2-bound and vectorised 16 bytes per thread.
#include "fx_common.cuh"

namespace fx {

// ------------------------------------------------------------------------------------------
// fp32 implicit GEMM: M = n*ho*wo output pixels, N = cout, K = kh*kw*cin (cin innermost).
// 64x64 tile per 256-thread block, 4x4 outputs per thread, K chunks of 16.
// ------------------------------------------------------------------------------------------
constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtConvArgs {
    const float* in;
    const float* w;
    const float* bias;
    const float* residual;
    float* out;
    int n, hin, win, cin;  // physical input dims
    int ho, wo, cout;
    int kh, kw, stride, pad;
    int K;  // kh*kw*cin
    int relu;
};

__global__ void __launch_bounds__(256) simt_conv_kernel(const SimtConvArgs a) {
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long M = (long long)a.n * a.ho * a.wo;
    const long long m0 = (long long)blockIdx.x * SBM;
    const int n0 = blockIdx.y * SBN;

    // loader roles: one float4 of A and one of B per thread per K chunk
    const int lm = tid >> 2;        // 0..63: pixel (A) / cout (B) within the tile
    const int lk = (tid & 3) * 4;   // 0,4,8,12: k offset within the chunk
    const long long pm = m0 + lm;
    const bool pm_ok = pm < M;
    int img = 0, oh = 0, ow = 0;
    if (pm_ok) {
        img = (int)(pm / (a.ho * a.wo));
        const int rem = (int)(pm - (long long)img * a.ho * a.wo);
        oh = rem / a.wo;
        ow = rem - oh * a.wo;
    }
    const int ih0 = oh * a.stride - a.pad, iw0 = ow * a.stride - a.pad;
    const float* in_img = a.in + (size_t)img * a.hin * a.win * a.cin;
    const int bn = n0 + lm;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += SBK) {
        const int k = k0 + lk;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (k < a.K) {
            const int tap = k / a.cin, c = k - tap * a.cin;
            const int r = tap / a.kw, s = tap - r * a.kw;
            const int ih = ih0 + r, iw = iw0 + s;
            if (pm_ok && ih >= 0 && ih < a.hin && iw >= 0 && iw < a.win)
                av = *reinterpret_cast<const float4*>(in_img + ((size_t)ih * a.win + iw) * a.cin + c);
            if (bn < a.cout) bv = *reinterpret_cast<const float4*>(a.w + (size_t)bn * a.K + k);
        }
        As[lk + 0][lm] = av.x;
        As[lk + 1][lm] = av.y;
        As[lk + 2][lm] = av.z;
        As[lk + 3][lm] = av.w;
        Bs[lk + 0][lm] = bv.x;
        Bs[lk + 1][lm] = bv.y;
        Bs[lk + 2][lm] = bv.z;
        Bs[lk + 3][lm] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            const float4 x = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 y = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float xa[4] = {x.x, x.y, x.z, x.w};
            const float ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int nn = n0 + tx * 4;
    if (nn >= a.cout) return;
    const float4 b = *reinterpret_cast<const float4*>(a.bias + nn);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4 v = make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
        const size_t o = (size_t)m * a.cout + nn;
        if (a.residual) {
            const float4 r = *reinterpret_cast<const float4*>(a.residual + o);
            v.x += r.x;
            v.y += r.y;
            v.z += r.z;
            v.w += r.w;
        }
        if (a.relu) {
            v.x = fmaxf(v.x, 0.f);
            v.y = fmaxf(v.y, 0.f);
            v.z = fmaxf(v.z, 0.f);
            v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4*>(a.out + o) = v;
    }
}

int simt_conv(fx_engine* e, const PackedLayer& L, const float* in, int hin_phys, int win_phys, int cin_phys, int pad,
              const float* residual, float* out, int n, int relu, cudaStream_t stream) {
    SimtConvArgs a;
    a.in = in;
    a.w = L.w_f32;
    a.bias = L.bias;
    a.residual = residual;
    a.out = out;
    a.n = n;
    a.hin = hin_phys;
    a.win = win_phys;
    a.cin = cin_phys;
    a.ho = L.g.hout;
    a.wo = L.g.wout;
    a.cout = L.g.cout;
    a.kh = L.g.kh;
    a.kw = L.g.kw;
    a.stride = L.g.stride;
    a.pad = pad;
    a.K = L.g.kh * L.g.kw * cin_phys;
    a.relu = relu;
    const long long M = (long long)n * a.ho * a.wo;
    dim3 grid((unsigned)((M + SBM - 1) / SBM), (a.cout + SBN - 1) / SBN);
    simt_conv_kernel<<<grid, 256, 0, stream>>>(a);
    FX_LAUNCH_CHECK(e, "simt_conv_kernel");
    return FX_OK;
}

// ------------------------------------------------------------------------------------------
// maxpool 3x3 / stride 2 / pad 1 on NHWC; 16 bytes of channels per thread
// ------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void maxpool_kernel(const void* __restrict__ in_, void* __restrict__ out_, int n, int h, int w, int c) {
    constexpr int V = BF16 ? 8 : 4;  // channels per 16 bytes
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    const int cv = c / V;
    const size_t total = (size_t)n * ho * wo * cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % cv);
        size_t p = i / cv;
        const int x = (int)(p % wo);
        p /= wo;
        const int y = (int)(p % ho);
        const int img = (int)(p / ho);
        float m[V];
#pragma unroll
        for (int k = 0; k < V; ++k) m[k] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * y - 1 + dy;
            if (iy < 0 || iy >= h) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * x - 1 + dx;
                if (ix < 0 || ix >= w) continue;
                const size_t off = (((size_t)img * h + iy) * w + ix) * c + (size_t)cc * V;
                if (BF16) {
                    const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(in_) + off);
                    const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        m[2 * k] = fmaxf(m[2 * k], __uint_as_float(u[k] << 16));
                        m[2 * k + 1] = fmaxf(m[2 * k + 1], __uint_as_float(u[k] & 0xffff0000u));
                    }
                } else {
                    const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in_) + off);
                    m[0] = fmaxf(m[0], v.x);
                    m[1] = fmaxf(m[1], v.y);
                    m[2] = fmaxf(m[2], v.z);
                    m[3] = fmaxf(m[3], v.w);
                }
            }
        }
        const size_t oo = (((size_t)img * ho + y) * wo + x) * c + (size_t)cc * V;
        if (BF16) {
            uint4 v;
            unsigned* u = &v.x;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                u[k] = (__float_as_uint(m[2 * k]) >> 16) | (__float_as_uint(m[2 * k + 1]) & 0xffff0000u);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_) + oo) = v;
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + oo) = make_float4(m[0], m[1], m[2], m[3]);
        }
    }
}

int maxpool_3x3s2(fx_engine* e, const void* in, void* out, int n, int h, int w, int c, bool bf16, cudaStream_t stream) {
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    const size_t total = (size_t)n * ho * wo * (c / (bf16 ? 8 : 4));
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)e->sm_count * 32);
    if (bf16)
        maxpool_kernel<true><<<blocks, 256, 0, stream>>>(in, out, n, h, w, c);
    else
        maxpool_kernel<false><<<blocks, 256, 0, stream>>>(in, out, n, h, w, c);
    FX_LAUNCH_CHECK(e, "maxpool_kernel");
    return FX_OK;
}

// ------------------------------------------------------------------------------------------
// global average pool: [n][hw][c] -> fp32 [n][c]; fixed summation order (deterministic)
// ------------------------------------------------------------------------------------------
// HW > 0: the plane size is a compile-time constant (49 for ResNet-18 at 224 x 224), all HW loads of a thread are issued
// before the first add -- the runtime loop kept ~4 loads per thread in flight and ran at 2.4 TB/s on a 25.7 MB input.
// The additions happen in the same order (p ascending) either way: the results are bit-identical.
template <bool BF16, int HW>
__global__ void avgpool_kernel(const void* __restrict__ in_, float* __restrict__ out, int n, int hw, int c) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    const int img = i / c, ch = i - img * c;
    float s = 0.f;
    if (HW > 0) {
        float v[HW > 0 ? HW : 1];
#pragma unroll
        for (int p = 0; p < HW; ++p) {
            const size_t off = ((size_t)img * HW + p) * c + ch;
            v[p] = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in_)[off]) : reinterpret_cast<const float*>(in_)[off];
        }
#pragma unroll
        for (int p = 0; p < HW; ++p) s += v[p];
    } else {
        for (int p = 0; p < hw; ++p) {
            const size_t off = ((size_t)img * hw + p) * c + ch;
            s += BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in_)[off])
                      : reinterpret_cast<const float*>(in_)[off];
        }
    }
    out[i] = s / (float)hw;
}

const void* avgpool_kernel_ptr(bool in_is_bf16) {
    return in_is_bf16 ? reinterpret_cast<const void*>(avgpool_kernel<true, 49>) : reinterpret_cast<const void*>(avgpool_kernel<false, 49>);
}

int avgpool_7x7(fx_engine* e, const void* in, bool in_is_bf16, float* out, int n, int hw, int c, cudaStream_t stream) {
    const int total = n * c;
    const dim3 grid((total + 127) / 128), block(128);
    if (hw == 49) {
        if (in_is_bf16)
            FX_CUDA(e, launch_pdl(avgpool_kernel<true, 49>, grid, block, 0, stream, in, out, n, hw, c));
        else
            FX_CUDA(e, launch_pdl(avgpool_kernel<false, 49>, grid, block, 0, stream, in, out, n, hw, c));
    } else if (in_is_bf16) {
        FX_CUDA(e, launch_pdl(avgpool_kernel<true, 0>, grid, block, 0, stream, in, out, n, hw, c));
    } else {
        FX_CUDA(e, launch_pdl(avgpool_kernel<false, 0>, grid, block, 0, stream, in, out, n, hw, c));
    }
    FX_LAUNCH_CHECK(e, "avgpool_kernel");
    return FX_OK;
}

// ------------------------------------------------------------------------------------------
// classifier head: logits = emb @ W^T + b, probs = softmax(logits) -- fc of create_model (src/training/common.py:
// 299-304) + torch.softmax(outputs, dim=1) (src/training/semi_supervised.py:61, common.py:463, threshold_sweep.py:34).
// One warp per image; a lane sums its 16 products in k order, the lanes combine by a fixed xor tree: the result
// depends on the image alone.  exp through expf (not the fast intrinsic).
// ------------------------------------------------------------------------------------------
__global__ void head_kernel(const float* __restrict__ emb, const float* __restrict__ w, const float* __restrict__ b, int n, int classes,
                            float* __restrict__ logits, float* __restrict__ probs) {
    pdl_launch_dependents();
    pdl_wait();
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (img >= n) return;
    float x[kEmbed / 32];
#pragma unroll
    for (int j = 0; j < kEmbed / 32; ++j) x[j] = emb[(size_t)img * kEmbed + j * 32 + lane];
    float mine = 0.f, mx = -INFINITY;  // lane c keeps logit c (classes <= FX_MAX_CLASSES = 32)
    for (int c = 0; c < classes; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < kEmbed / 32; ++j) acc = fmaf(x[j], __ldg(w + (size_t)c * kEmbed + j * 32 + lane), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc += __ldg(b + c);
        if (lane == c) mine = acc;
        mx = fmaxf(mx, acc);
    }
    if (logits && lane < classes) logits[(size_t)img * classes + lane] = mine;
    if (!probs) return;
    const float ev = lane < classes ? expf(mine - mx) : 0.f;
    float sum = ev;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane < classes) probs[(size_t)img * classes + lane] = ev / sum;
}

int head_run(fx_engine* e, const float* emb, int n, float* logits, float* probs, cudaStream_t stream) {
    FX_CUDA(e, launch_pdl(head_kernel, dim3((n + 3) / 4), dim3(128), 0, stream, emb, static_cast<const float*>(e->head_w),
                          static_cast<const float*>(e->head_b), n, e->head_classes, logits, probs));
    FX_LAUNCH_CHECK(e, "head_kernel");
    return FX_OK;
}

// ------------------------------------------------------------------------------------------
// layout / precision helpers (debug entry points only)
// ------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __bfloat162float(in[i]);
}

int f32_to_bf16(fx_engine* e, const float* in, __nv_bfloat16* out, size_t count, cudaStream_t stream) {
    if (!count) return FX_OK;
    f32_to_bf16_kernel<<<(int)std::min<size_t>((count + 255) / 256, 4096), 256, 0, stream>>>(in, out, count);
    FX_LAUNCH_CHECK(e, "f32_to_bf16_kernel");
    return FX_OK;
}
int bf16_to_f32(fx_engine* e, const __nv_bfloat16* in, float* out, size_t count, cudaStream_t stream) {
    if (!count) return FX_OK;
    bf16_to_f32_kernel<<<(int)std::min<size_t>((count + 255) / 256, 4096), 256, 0, stream>>>(in, out, count);
    FX_LAUNCH_CHECK(e, "bf16_to_f32_kernel");
    return FX_OK;
}

// fp32 NHWC [n][224][224][3] -> conv1 staging layout (interior only; the pad stays zero)
template <bool BF16>
__global__ void pad_in0_kernel(const float* __restrict__ in, void* __restrict__ out, int n) {
    const size_t total = (size_t)n * kCrop * kCrop;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t img = i / (kCrop * kCrop);
        const int rem = (int)(i - img * kCrop * kCrop);
        const int y = rem / kCrop, x = rem - y * kCrop;
        const float c0 = in[3 * i], c1 = in[3 * i + 1], c2 = in[3 * i + 2];
        if (BF16) {
            const int py = y + kIn0Pad, px = x + kIn0Pad;
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) +
                               ((img * kS2dH + (py >> 1)) * kS2dW + (px >> 1)) * kS2dC + ((py & 1) * 2 + (px & 1)) * 3;
            o[0] = __float2bfloat16_rn(c0);
            o[1] = __float2bfloat16_rn(c1);
            o[2] = __float2bfloat16_rn(c2);
        } else {
            const size_t pix = (img * kIn0H + (y + kIn0Pad)) * kIn0W + (x + kIn0Pad);
            reinterpret_cast<float4*>(out)[pix] = make_float4(c0, c1, c2, 0.f);
        }
    }
}

int pad_nhwc3_to_in0(fx_engine* e, const float* in_nhwc3, void* in0, bool bf16, int n, cudaStream_t stream) {
    const size_t total = (size_t)n * kCrop * kCrop;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 4096);
    if (bf16)
        pad_in0_kernel<true><<<blocks, 256, 0, stream>>>(in_nhwc3, in0, n);
    else
        pad_in0_kernel<false><<<blocks, 256, 0, stream>>>(in_nhwc3, in0, n);
    FX_LAUNCH_CHECK(e, "pad_in0_kernel");
    return FX_OK;
}

}  // namespace fx
