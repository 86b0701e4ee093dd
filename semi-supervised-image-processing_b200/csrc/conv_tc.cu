// bf16 implicit-GEMM convolution on the Blackwell tensor cores (tcgen05 + TMEM), fed by TMA: the
// PER-TAP kernels.  They serve the Conv2d+BatchNorm2d(+ReLU)(+residual add) groups of the frozen
// ResNet-18 trunk (src/feature_extraction.py:290-291; torchvision/models/resnet.py:89-105,197-206,
// 266-282) that the halo-tile kernels of conv_flat.cu do not: the stride-2 3x3 convs, the 1x1/s2
// downsample convs and all of layer3 / layer4 (14x14 and 7x7 planes).  BN is folded into the
// weights/bias on the host (engine.cu).
//
// GEMM view: M = n*ho*wo output pixels, N = cout, K = (kh, kw, cin) with cin innermost.
//   * activations are NHWC bf16, so the K-slice of one filter tap (r, s) and 64 input channels of
//     a rectangular patch of output pixels is a 4-D box {64 c, Wt w, Ht h, Nt n} of the input
//     tensor: ONE tiled TMA load (cp.async.bulk.tensor.4d) brings it in as a 128-row x 128-byte
//     K-major SWIZZLE_128B operand tile.  Conv zero padding = TMA out-of-bounds zero fill;
//     stride-2 convs = TMA element strides.  No im2col buffer ever exists.
//   * weights are packed [cout][K] bf16; one 2-D TMA load per K-slice gives the B operand tile.
//   * one elected thread (elect.sync in warp-uniform code) issues tcgen05.mma into a double-buffered
//     fp32 accumulator in TMEM; 16 epilogue warps read it back with tcgen05.ld, add the folded-BN
//     bias (shared memory) and the residual (prefetched before the accumulator wait), apply ReLU, and
//     store bf16 NHWC (or fp32 for the layer feeding avgpool).
//   * warp-specialised, persistent over output tiles: warp 0 = TMA producer, warp 1 = MMA issuer,
//     warp 2 = TMEM allocator, warps 4-19 = epilogue; mbarrier rings smem(full/empty), tmem(full/empty).
//   * tc_conv_kernel: one CTA per 128 x BN tile (BN = 64).  tc2_conv_kernel: CTA PAIR (cta_group::2) per
//     256 x BN tile (BN = 128 / 256), half of every weight K-block per CTA -- the single-CTA form is bound
//     by the chip-wide L2->SM delivery (~12.4 TB/s measured); grouped launch of a block's 1x1/s2
//     downsample conv as extra tiles; split tail wave (half-N units) when the last wave is mostly empty.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tc_ptx.cuh"

namespace fx {

// ------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------
struct TcConvParams {
    int wt_log2, ht_log2, nt_log2;  // output tile = (1<<wt) x (1<<ht) pixels of (1<<nt) images = 128 rows
    int tiles_w, tiles_h, tiles_g;  // tiles along ow, oh, image groups
    int n_tiles_n;                  // cout / BN
    int total_tiles;
    int batch, ho, wo, cout;
    int kh, kw, cchunks;               // K loop: taps x (cin / BK)
    int cw_mul, ch_mul, pad_w, pad_h;  // A box start: w = ow0*cw_mul - pad_w + s, h = oh0*ch_mul - pad_h + r
    const float* bias;
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    float* out_f32;
    int relu;
    // grouped launch: `ds_tiles` extra tiles (after the main ones) compute the block's 1x1 / stride-2 downsample conv
    // (torchvision/models/resnet.py:239-243) on the SAME input and output tiling: one tap, no padding, no ReLU,
    // weights from map_b2.  It shares the launch (and fills the tail wave) of the block's 3x3 / stride-2 conv1.
    int ds_tiles;
    // CTA-pair kernel with resident weights (layer2.0): the downsample conv is FUSED into its conv1 unit instead -- its
    // input pixel (2 oh, 2 ow) is exactly the centre tap's A tile, so at that K-block a second MMA with the downsample
    // weights accumulates into a second TMEM accumulator; no extra units, no second pass over the input
    int ds_fused;
    const float* ds_bias;
    __nv_bfloat16* ds_out;
    int split_full, split_tail;  // CTA-pair kernel only: see the work-unit comment in tc2_conv_kernel
    // CTA-pair kernel, stride-2 layers: the input is addressed through two 5-D maps [cin][W/2][row parity][H/2][n], one
    // per COLUMN parity (base + cin elements), so that every tap is a plain unit-stride box of one parity plane.  A box
    // with elementStrides = 2 costs the TMA unit a walk over its whole bounding box (measured: ~1000 cycles per
    // 128-pixel box, four times the MMAs it feeds; the traffic itself was never the problem).
    int s2planes;
    int staged256;  // BN = 256 epilogue through per-warp staging blocks (FX_TC_STAGED256=0: direct per-lane stores / loads)
};

constexpr int kTcThreads = 128 + 16 * 32;  // 4 control warps + 16 epilogue warps (4 TMEM lane quarters x 4 column groups)

template <int BN, int BK, int STAGES>
struct TcSmem {
    static constexpr int kA = 128 * BK * 2;
    static constexpr int kB = BN * BK * 2;
    static constexpr int kBars = (2 * STAGES + 4) * 8;
    static constexpr int kBias = 2 * 512 * 4;  // folded-BN bias of every output channel (cout <= 512), main + downsample
    static constexpr int kTotal = 1024 /*align slack*/ + STAGES * (kA + kB) + kBars + 16 + kBias;
};

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_b2, const TcConvParams p) {
    using L = TcSmem<BN, BK, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = sbase, sB = sbase + STAGES * L::kA;
    const uint32_t bars = sB + STAGES * L::kB;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    const uint32_t tslot = tempty0 + 16;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));
    float* bias_s = reinterpret_cast<float*>(smem_raw + (tslot + 16 - smem_u32(smem_raw)));
    constexpr bool kBiasSmem = true;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 2 * BN;
    for (int i = threadIdx.x; i < p.cout; i += kTcThreads) {
        bias_s[i] = __ldg(p.bias + i);
        if (p.ds_tiles) bias_s[512 + i] = __ldg(p.ds_bias + i);
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, 512);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tslot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    pdl_wait();

    const int num_kb = p.kh * p.kw * p.cchunks;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int t0 = blockIdx.x; t0 < p.total_tiles + p.ds_tiles; t0 += gridDim.x) {
                const bool ds = t0 >= p.total_tiles;
                const int t = ds ? t0 - p.total_tiles : t0;
                const int n_tile = t % p.n_tiles_n;
                int m_tile = t / p.n_tiles_n;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int tg = m_tile / p.tiles_h;
                const int w0 = (tw << p.wt_log2) * p.cw_mul - (ds ? 0 : p.pad_w);
                const int h0 = (th << p.ht_log2) * p.ch_mul - (ds ? 0 : p.pad_h);
                const int n0 = tg << p.nt_log2;
                const int kh = ds ? 1 : p.kh, kw = ds ? 1 : p.kw;
                const CUtensorMap* mb = ds ? &map_b2 : &map_b;
                int kb = 0;
                for (int r = 0; r < kh; ++r)
                    for (int s = 0; s < kw; ++s)
                        for (int cc = 0; cc < p.cchunks; ++cc, ++kb) {
                            mbar_wait(empty0 + 8 * stage, phase ^ 1);
                            mbar_expect_tx(full0 + 8 * stage, L::kA + L::kB);
                            tma_load_4d(sA + stage * L::kA, &map_a, full0 + 8 * stage, cc * BK, w0 + s, h0 + r, n0);
                            tma_load_2d(sB + stage * L::kB, mb, full0 + 8 * stage, kb * BK, n_tile * BN);
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform schedule, one elected lane issues =====
        constexpr uint32_t idesc = make_idesc<BN>();
        constexpr uint64_t desc_hi = make_smem_desc<BK>(0) & 0xFFFFFFFF00000000ull;
        uint32_t stage = 0, phase = 0;
        int it = 0;
        for (int t0 = blockIdx.x; t0 < p.total_tiles + p.ds_tiles; t0 += gridDim.x, ++it) {
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            mbar_wait(tempty0 + 8 * as, aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            const int nkb = t0 >= p.total_tiles ? p.cchunks : num_kb;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_lo = (sA + stage * L::kA) >> 4, b_lo = (sB + stage * L::kB) >> 4;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(tmem_d, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (kb | k) != 0);
                    umma_commit(empty0 + 8 * stage);  // frees the smem slot once these MMAs retire
                    if (kb == nkb - 1) umma_commit(tfull0 + 8 * as);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+bias, +residual, ReLU) -> global =====
        const int q = warp & 3;           // TMEM lane quarter this warp may read
        const int cg = (warp - 4) >> 2;   // which quarter of the tile's BN columns this warp converts and stores
        const int row = q * 32 + lane;
        const int wl = row & ((1 << p.wt_log2) - 1);
        const int hl = (row >> p.wt_log2) & ((1 << p.ht_log2) - 1);
        const int nl = row >> (p.wt_log2 + p.ht_log2);
        int it = 0;
        for (int t0 = blockIdx.x; t0 < p.total_tiles + p.ds_tiles; t0 += gridDim.x, ++it) {
            const bool ds = t0 >= p.total_tiles;
            const int t = ds ? t0 - p.total_tiles : t0;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int n_tile = t % p.n_tiles_n;
            int m_tile = t / p.n_tiles_n;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int tg = m_tile / p.tiles_h;
            const int ow = (tw << p.wt_log2) + wl, oh = (th << p.ht_log2) + hl, img = (tg << p.nt_log2) + nl;
            const bool valid = ow < p.wo && oh < p.ho && img < p.batch;
            const size_t pix = ((size_t)img * p.ho + oh) * p.wo + ow;
            const size_t obase = pix * p.cout + (size_t)n_tile * BN;

            // the residual does not depend on the accumulator: fetch this thread's BN/4 channels while the MMAs
            // of the tile are still running, so the load latency is off the epilogue's critical path
            constexpr int kResVec = BN / 32;  // uint4 (8 bf16) per thread
            uint4 res[kResVec];
            const bool has_res = valid && !ds && p.residual != nullptr;
            if (has_res) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + cg * (BN / 4));
#pragma unroll
                for (int j = 0; j < kResVec; ++j) res[j] = __ldg(rp + j);
            }
            mbar_wait(tfull0 + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int ci = 0; ci < BN / 64; ++ci) {
                const int c0 = cg * (BN / 4) + ci * 16;
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
                    float f[16];
                    const float4* bp = kBiasSmem ? reinterpret_cast<const float4*>(bias_s + (ds ? 512 : 0) + n_tile * BN + c0)
                                                 : reinterpret_cast<const float4*>((ds ? p.ds_bias : p.bias) + n_tile * BN + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b = kBiasSmem ? bp[j] : __ldg(bp + j);
                        f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
                        f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
                        f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
                        f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
                    }
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint4 rv = res[ci * 2 + j];
                            const unsigned u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                f[8 * j + 2 * k] += __uint_as_float(u[k] << 16);
                                f[8 * j + 2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
                            }
                        }
                    }
                    if (p.relu && !ds) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (p.out_f32 && !ds) {
                        float4* op = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        uint4* op = reinterpret_cast<uint4*>((ds ? p.ds_out : p.out) + obase + c0);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint4 o;
                            unsigned* u = &o.x;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
                                u[k] = *reinterpret_cast<const unsigned*>(&h2);
                            }
                            op[j] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty0 + 8 * as);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the Cout >= 256 layers (layer3 / layer4).  Measured: the single-CTA
// kernel above is bound by L2->SM delivery (~12.4 TB/s chip-wide; a 128 x 256 tile needs 96 B/clk/SM at the
// tensor floor).  Here two CTAs of a cluster compute one 256 x 256 tile with ONE M=256 MMA stream issued by
// the even CTA: each CTA loads its own 128 A rows and only HALF of the weight K-block (128 of the 256 Cout
// rows), so the weight bytes per SM halve (64 B/clk/SM at the floor).  Barriers: the leader's `full`
// barrier counts the bytes of both CTAs' TMA loads; tcgen05.commit multicasts `empty` / `tmem full` to both.
// ------------------------------------------------------------------------------------------
// RESB (layer2.0 conv1 + downsample: 3x3 x 64 -> 128, ten K-blocks): this CTA's half of EVERY weight K-block stays
// resident in shared memory (80 KB), loaded once before the previous kernel has finished (PDL); the ring then carries
// activations only.  That kernel is bound by L2 -> SM delivery (24 KB per K-block and CTA against 256 MMA cycles);
// without the weight reloads it is 16 KB.
constexpr int kResKb = 10;
// BN = 256: five 32 KB stages; six leave no room for the epilogue's 16 x 2 KB staging blocks.  Measured per launch: direct
// epilogue 6 stages 47.6 us, 5 stages 49.9; staged epilogue 5 stages 45.6 .. 47.6 (stride-2 launches 43.4 -> 39.3); six stages
// bought by reading the bias through L1 instead of shared memory: slower again (the short units are paced by the epilogue).
constexpr int kTc2Stages256 = 5;
template <int BN, int BK, int STAGES, bool RESB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
tc2_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_b2,
                const __grid_constant__ CUtensorMap map_bh, const TcConvParams p) {
    constexpr int kA = 128 * BK * 2, kB = (BN / 2) * BK * 2;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = sbase, sB = sbase + STAGES * kA;  // RESB: sB = kResKb resident K-blocks, not a ring
    const uint32_t bars = sB + (RESB ? kResKb : STAGES) * kB;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
    const uint32_t tslot = tempty0 + 16, wbar = tslot + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));
    float* bias_s = reinterpret_cast<float*>(smem_raw + (tslot + 16 - smem_u32(smem_raw)));
    // RESB: 2 KB per epilogue warp (32 rows x 64 B) to turn the accumulator layout (one pixel per lane) into whole
    // 64-byte pieces of pixel rows before they go to global memory
    constexpr bool kBiasSmem = true;  // (BN = 256 without the copy, bias through L1, to make room for a sixth stage: measured slower)
    const uint32_t stage_out0 = (tslot + 16 + (kBiasSmem ? 2 * 512 * 4 : 0) + 127u) & ~127u;
    const bool fused = RESB && p.ds_fused;
    const bool staged256_on = p.staged256 != 0;
    if (kBiasSmem)
        for (int i = threadIdx.x; i < p.cout; i += kTcThreads) {
            bias_s[i] = __ldg(p.bias + i);
            if (p.ds_tiles || fused) bias_s[512 + i] = __ldg(p.ds_bias + i);
        }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    // RESB (BN = 128): room for the fused downsample accumulator beside each of the two main ones
    constexpr uint32_t kAccStride = RESB ? 2 * BN : BN;
    constexpr uint32_t kTmemCols = 2 * kAccStride;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full0 + 8 * i, 2);   // leader's arrive.expect_tx + the peer producer's arrive
            mbar_init(empty0 + 8 * i, 1);  // multicast commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);                       // multicast commit
            mbar_init(tempty0 + 8 * i, 2 * (kTcThreads - 128));  // every epilogue thread of both CTAs
        }
        if (RESB) mbar_init(wbar, 2);  // leader's arrive.expect_tx + the peer producer's arrive
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tslot, kTmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    const int num_kb = p.kh * p.kw * p.cchunks;
    if (RESB && warp == 0 && elect_one_sync()) {  // constant data: fetched before waiting for the previous kernel
        const int n_res = num_kb + ((p.ds_tiles || fused) ? p.cchunks : 0);
        if (leader) mbar_expect_tx(wbar, 2 * n_res * kB);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d_2cta(sB + kb * kB, &map_b, wbar, kb * BK, (int)rank * (BN / 2));
        if (p.ds_tiles || fused)
            for (int cc = 0; cc < p.cchunks; ++cc) tma_load_2d_2cta(sB + (num_kb + cc) * kB, &map_b2, wbar, cc * BK, (int)rank * (BN / 2));
        if (!leader) mbar_arrive_leader(wbar);
    }
    __syncwarp();
    pdl_wait();

    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_g;
    const int pair_tiles = ((m_tiles + 1) >> 1) * p.n_tiles_n;
    // Work units.  Normally one unit = one 256 x BN pair tile (+ the grouped downsample tiles).  With a SPLIT TAIL
    // (p.split_tail > 0; layer4 at batch 256 has 98 tiles for 74 pairs) the tiles of the last, mostly empty wave are
    // cut into two 256 x BN/2 halves so that the tail wave costs half a tile time: unit u < split_full is full tile
    // u, the others are (tile split_full + (u - split_full) / 2, half (u - split_full) & 1), N = BN/2 MMAs.
    const int n_units = p.split_tail ? p.split_full + 2 * p.split_tail : pair_tiles + p.ds_tiles;  // fused downsample: ds_tiles == 0

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 A rows + own half of the weight K-block =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int t0 = pair; t0 < n_units; t0 += n_pairs) {
                const bool ds = !p.split_tail && t0 >= pair_tiles;
                const int half = (p.split_tail && t0 >= p.split_full) ? ((t0 - p.split_full) & 1) : -1;
                const int t = ds ? t0 - pair_tiles : (half >= 0 ? p.split_full + ((t0 - p.split_full) >> 1) : t0);
                const int n_tile = t % p.n_tiles_n;
                int m_tile = (t / p.n_tiles_n) * 2 + (int)rank;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int tg = m_tile / p.tiles_h;  // >= tiles_g for the odd leftover: every load is out of bounds -> zeros
                const int w0 = (tw << p.wt_log2) * p.cw_mul - (ds ? 0 : p.pad_w);
                const int h0 = (th << p.ht_log2) * p.ch_mul - (ds ? 0 : p.pad_h);
                const int n0 = tg << p.nt_log2;
                const int kh = ds ? 1 : p.kh, kw = ds ? 1 : p.kw;
                const CUtensorMap* mb = ds ? &map_b2 : &map_b;
                int kb = 0;
                for (int r = 0; r < kh; ++r)
                    for (int s = 0; s < kw; ++s)
                        for (int cc = 0; cc < p.cchunks; ++cc, ++kb) {
                            mbar_wait(empty0 + 8 * stage, phase ^ 1);
                            if (leader) mbar_expect_tx(full0 + 8 * stage, RESB ? 2 * kA : (half >= 0 ? 2 * (kA + kB / 2) : 2 * (kA + kB)));
                            if (p.s2planes) {
                                // input pixel = 2 * output pixel + (tap - pad): parity plane (tap - pad) & 1, plane index
                                // shifted by floor((tap - pad) / 2); negative / too large indices are zero fill = padding
                                const int dw = ds ? 0 : s - p.pad_w, dh = ds ? 0 : r - p.pad_h;
                                const int pw = dw & 1, ph = dh & 1;
                                tma_load_5d_2cta(sA + stage * kA, pw ? &map_a1 : &map_a, full0 + 8 * stage, cc * BK,
                                                 (tw << p.wt_log2) + (dw - pw) / 2, ph, (th << p.ht_log2) + (dh - ph) / 2, n0);
                            } else {
                                tma_load_4d_2cta(sA + stage * kA, &map_a, full0 + 8 * stage, cc * BK, w0 + s, h0 + r, n0);
                            }
                            if (RESB) {
                            } else if (half >= 0)
                                tma_load_2d_2cta(sB + stage * kB, &map_bh, full0 + 8 * stage, kb * BK,
                                                 n_tile * BN + half * (BN / 2) + (int)rank * (BN / 4));
                            else
                                tma_load_2d_2cta(sB + stage * kB, mb, full0 + 8 * stage, kb * BK, n_tile * BN + (int)rank * (BN / 2));
                            if (!leader) mbar_arrive_leader(full0 + 8 * stage);
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (leader) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            constexpr uint64_t desc_hi = make_smem_desc<BK>(0) & 0xFFFFFFFF00000000ull;
            uint32_t stage = 0, phase = 0;
            int it = 0;
            if (RESB) {
                mbar_wait(wbar, 0);
                tc_fence_after();
            }
            constexpr uint32_t idesc_half = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 4) << 17) | ((uint32_t)(256 >> 4) << 24);
            for (int t0 = pair; t0 < n_units; t0 += n_pairs, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(tempty0 + 8 * as, aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * kAccStride;
                const bool ds_unit = !p.split_tail && t0 >= pair_tiles;
                const int nkb = ds_unit ? p.cchunks : num_kb;
                const uint32_t idesc_u = (p.split_tail && t0 >= p.split_full) ? idesc_half : idesc;
                const int centre0 = ((p.kh >> 1) * p.kw + (p.kw >> 1)) * p.cchunks;  // first K-block of the centre tap
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t a_lo = (sA + stage * kA) >> 4;
                        const uint32_t b_lo = (RESB ? sB + ((ds_unit ? num_kb : 0) + kb) * kB : sB + stage * kB) >> 4;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_2cta(tmem_d, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc_u, (kb | k) != 0);
                        if (RESB && fused && kb >= centre0 && kb < centre0 + p.cchunks) {
                            // the 1x1 / stride-2 downsample reads input pixel (2 oh, 2 ow): this very A tile
                            const int cc = kb - centre0;
                            const uint32_t bd_lo = (sB + (num_kb + cc) * kB) >> 4;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_bf16_2cta(tmem_d + BN, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(bd_lo + 2 * k), idesc, (cc | k) != 0);
                        }
                        umma_commit_2cta(empty0 + 8 * stage);
                        if (kb == nkb - 1) umma_commit_2cta(tfull0 + 8 * as);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own 128 accumulator rows -> (+bias, +residual, ReLU) -> global =====
        const int q = warp & 3;
        const int cg = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int wl = row & ((1 << p.wt_log2) - 1);
        const int hl = (row >> p.wt_log2) & ((1 << p.ht_log2) - 1);
        const int nl = row >> (p.wt_log2 + p.ht_log2);
        int it = 0;
        for (int t0 = pair; t0 < n_units; t0 += n_pairs, ++it) {
            const bool ds_unit = !p.split_tail && t0 >= pair_tiles;
            const int half = (p.split_tail && t0 >= p.split_full) ? ((t0 - p.split_full) & 1) : -1;
            const int t = ds_unit ? t0 - pair_tiles : (half >= 0 ? p.split_full + ((t0 - p.split_full) >> 1) : t0);
            // output channels of this unit handled by this warp: BN/4 (full tile) or BN/8 (half tile) per column group
            const int cpg = half >= 0 ? BN / 8 : BN / 4;
            const int ch0 = (half >= 0 ? half * (BN / 2) : 0) + cg * cpg;  // first channel within the tile's BN
            const int tc0 = cg * cpg;                                      // first accumulator column
            const int nci = cpg / 16;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int n_tile = t % p.n_tiles_n;
            int m_tile = (t / p.n_tiles_n) * 2 + (int)rank;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int tg = m_tile / p.tiles_h;
            const int ow = (tw << p.wt_log2) + wl, oh = (th << p.ht_log2) + hl, img = (tg << p.nt_log2) + nl;
            const bool valid = ow < p.wo && oh < p.ho && img < p.batch;
            const size_t pix = ((size_t)img * p.ho + oh) * p.wo + ow;
            const size_t obase = pix * p.cout + (size_t)n_tile * BN;

            if (BN == 256 && STAGES == kTc2Stages256 && !RESB && (nci & 1) == 0 && !(p.out_f32 && !ds_unit) && staged256_on) {
                // ===== staged epilogue (BN = 256): the accumulator layout is one pixel row per lane, so a direct 16-byte store (or
                // residual load) per lane touches 32 different lines per instruction -- 32 wavefronts on the shared-memory / L1
                // data pipe that the MMA operand fetch and the TMA writes already fill.  Rounds of 32 channels go through a 2 KB
                // block per warp instead (32 rows x 64 B, chunk c of row r at (c ^ (r >> 1)) & 3): residual in (8 rows x 64
                // contiguous bytes per instruction), result out the same way. =====
                const bool res_unit = !ds_unit && p.residual != nullptr;
                const uint32_t stg = stage_out0 + (uint32_t)(warp - 4) * 2048;
                const int mypix = valid ? (int)pix : -1;
                int prs[4];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) prs[i4] = __shfl_sync(0xffffffffu, mypix, i4 * 8 + (lane >> 2));
                const size_t cb = (size_t)n_tile * BN + ch0 + (lane & 3) * 8;  // this lane's 8 channels of a round's 32 in the store-out role
                uint4 rres[4];  // the residual of the coming round (round 1's loads fly while round 0 is computed)
                if (res_unit) {
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4)  // unpredicated, clamped: a predicated load is waited for at once
                        rres[i4] = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)max(prs[i4], 0) * p.cout + cb));
                }
                mbar_wait(tfull0 + 8 * as, aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + as * kAccStride + ((uint32_t)(q * 32) << 16);
                __nv_bfloat16* obase_p = (ds_unit ? p.ds_out : p.out);
#pragma unroll
                for (int rd = 0; rd < 2; ++rd) {
                    if (rd * 2 >= nci) break;
                    if (res_unit) {
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const int r = i4 * 8 + (lane >> 2);
                            sts128(stg + r * 64 + ((((lane & 3) ^ (r >> 1)) & 3) << 4), rres[i4]);
                        }
                        if (rd == 0 && nci > 2) {
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4)
                                rres[i4] = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)max(prs[i4], 0) * p.cout + cb + 32));
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int ci = rd * 2 + c2;
                        const int c0 = ch0 + ci * 16;
                        uint32_t v[16];
                        tmem_ld16(taddr + tc0 + ci * 16, v);
                        tmem_ld_wait();
                        if (ci == nci - 1) {
                            tc_fence_before();
                            mbar_arrive_leader(tempty0 + 8 * as);
                        }
                        if (valid) {
                            float f[16];
                            const float4* bp = reinterpret_cast<const float4*>(bias_s + (ds_unit ? 512 : 0) + n_tile * BN + c0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 b = bp[j];
                                f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
                                f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
                                f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
                                f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
                            }
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint32_t sa = stg + lane * 64 + ((((c2 * 2 + j) ^ (lane >> 1)) & 3) << 4);
                                if (res_unit) {
                                    const uint4 rv = lds128(sa);
                                    const unsigned u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        f[8 * j + 2 * k] += __uint_as_float(u[k] << 16);
                                        f[8 * j + 2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
                                    }
                                }
                                if (p.relu && !ds_unit) {
#pragma unroll
                                    for (int k = 0; k < 8; ++k) f[8 * j + k] = fmaxf(f[8 * j + k], 0.f);
                                }
                                uint4 o;
                                unsigned* u2 = &o.x;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
                                    u2[k] = *reinterpret_cast<const unsigned*>(&h2);
                                }
                                sts128(sa, o);
                            }
                        }
                    }
                    __syncwarp();
                    {
                        uint4 val[4];
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const int r = i4 * 8 + (lane >> 2);
                            val[i4] = lds128(stg + r * 64 + ((((lane & 3) ^ (r >> 1)) & 3) << 4));
                        }
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4)
                            if (prs[i4] >= 0) *reinterpret_cast<uint4*>(obase_p + (size_t)prs[i4] * p.cout + cb + rd * 32) = val[i4];
                    }
                    __syncwarp();
                }
                continue;
            }
            // fused downsample: pass 0 drains the conv1 accumulator, pass 1 the downsample one beside it
            const int n_pass = (RESB && fused) ? 2 : 1;
            for (int pass = 0; pass < n_pass; ++pass) {
            const bool ds = (RESB && fused) ? pass == 1 : ds_unit;
            // the residual does not depend on the accumulator: fetch this thread's BN/4 channels while the MMAs
            // of the tile are still running, so the load latency is off the epilogue's critical path
            constexpr int kResVec = BN / 32;  // uint4 (8 bf16) per thread
            uint4 res[kResVec];
            const bool has_res = valid && !ds && p.residual != nullptr;
            if (has_res) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + ch0);
#pragma unroll
                for (int j = 0; j < kResVec; ++j)
                    if (j < 2 * nci) res[j] = __ldg(rp + j);
            }
            if (pass == 0) {
                mbar_wait(tfull0 + 8 * as, aphase);
                tc_fence_after();
            }
            const uint32_t taddr = tmem_base + as * kAccStride + (pass ? BN : 0) + ((uint32_t)(q * 32) << 16);
            // Short-K units (nine K-blocks, or ONE for the grouped downsample) are paced by this epilogue, and a store in
            // which every lane writes 16 bytes of a different pixel row is 32 separate line requests.  RESB: stage the
            // warp's 32 rows x 64 B in shared memory (16-byte chunk c of row r at ((c ^ (r >> 1)) & 3), conflict-free
            // both ways) and store 8 rows x 64 contiguous bytes per instruction.
            const bool staged = RESB && BN == 128 && !(p.out_f32 && !ds);
            const uint32_t stg = stage_out0 + (uint32_t)(warp - 4) * 2048;
#pragma unroll
            for (int ci = 0; ci < BN / 64; ++ci) {
                if (ci >= nci) break;
                const int c0 = ch0 + ci * 16;
                uint32_t v[16];
                tmem_ld16(taddr + tc0 + ci * 16, v);
                tmem_ld_wait();
                if (valid) {
                    float f[16];
                    const float4* bp = kBiasSmem ? reinterpret_cast<const float4*>(bias_s + (ds ? 512 : 0) + n_tile * BN + c0)
                                                 : reinterpret_cast<const float4*>((ds ? p.ds_bias : p.bias) + n_tile * BN + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b = kBiasSmem ? bp[j] : __ldg(bp + j);
                        f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
                        f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
                        f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
                        f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
                    }
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint4 rv = res[ci * 2 + j];
                            const unsigned u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                f[8 * j + 2 * k] += __uint_as_float(u[k] << 16);
                                f[8 * j + 2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
                            }
                        }
                    }
                    if (p.relu && !ds) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (p.out_f32 && !ds) {
                        float4* op = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        uint4* op = reinterpret_cast<uint4*>((ds ? p.ds_out : p.out) + obase + c0);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint4 o;
                            unsigned* u = &o.x;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 2 * k], f[8 * j + 2 * k + 1]);
                                u[k] = *reinterpret_cast<const unsigned*>(&h2);
                            }
                            if (staged)
                                sts128(stg + lane * 64 + ((((ci * 2 + j) ^ (lane >> 1)) & 3) << 4), o);
                            else
                                op[j] = o;
                        }
                    }
                }
            }
            if (pass == n_pass - 1) {
                tc_fence_before();
                mbar_arrive_leader(tempty0 + 8 * as);
            }
            if (staged) {
                // the shared-memory pipe is busy feeding the MMAs, so every dependent trip through it costs hundreds of cycles:
                // the four shuffles go out together, then the four loads, then the four stores (two trips instead of eight)
                __syncwarp();
                const int mypix = valid ? (int)pix : -1;
                __nv_bfloat16* obuf = (ds ? p.ds_out : p.out) + (size_t)n_tile * BN + ch0 + (lane & 3) * 8;
                int pr[4];
                uint4 val[4];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) pr[i4] = __shfl_sync(0xffffffffu, mypix, i4 * 8 + (lane >> 2));
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const int r = i4 * 8 + (lane >> 2);
                    val[i4] = lds128(stg + r * 64 + ((((lane & 3) ^ (r >> 1)) & 3) << 4));
                }
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4)
                    if (pr[i4] >= 0) *reinterpret_cast<uint4*>(obuf + (size_t)pr[i4] * p.cout) = val[i4];
                __syncwarp();
            }
            }  // pass
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory / barriers stay alive until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// TMA probe (test entry point): one box load -> shared memory -> global, raw bytes.
// ------------------------------------------------------------------------------------------
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int bytes,
                                 uint8_t* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bar = sbase + ((bytes + 15) & ~15);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) sptr[i] = 0xCD;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        tma_load_4d(sbase, &map, bar, c0, c1, c2, c3);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sptr[i];
}

// ------------------------------------------------------------------------------------------
// UMMA descriptor probe (test entry point): D[128 x 64] = A[shift .. shift+128) x B^T where A is a
// [256 rows][KB] K-major swizzled tile written by ONE TMA box and the A descriptor simply starts
// `shift` rows into it (start address not aligned to the swizzle repeat).  Pins the behaviour the
// halo-tile convolution kernels rely on: the swizzle XOR is a function of the absolute shared
// memory address, so a row-shifted view of a TMA-written tile is a valid operand.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_shift_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int kb_elems,
                        int shift_rows, int base_offset, int layout_code, float* out) {
    const int ws_mode = base_offset >= 100;  // test hook: base_offset 100 = weight-stationary MMA with base offset 0
    if (ws_mode) base_offset = 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int rowbytes = kb_elems * 2;
    const uint32_t sA = sbase, sB = sbase + 256 * rowbytes;
    const uint32_t bar0 = sB + 64 * rowbytes, bar1 = bar0 + 8, tslot = bar1 + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tslot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar0, 320 * rowbytes);
        tma_load_2d(sA, &map_a, bar0, 0, 0);
        tma_load_2d(sB, &map_b, bar0, 0, 0);
        mbar_wait(bar0, 0);
        tc_fence_after();
        const uint64_t sbo = (uint64_t)((8 * rowbytes) >> 4);
        const uint64_t hi = (sbo << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) | ((uint64_t)layout_code << 61);
        const uint64_t da = (uint64_t)(((sA + shift_rows * rowbytes) >> 4) & 0x3FFF) | hi;
        const uint64_t db = (uint64_t)((sB >> 4) & 0x3FFF) | (sbo << 32) | (1ull << 46) | ((uint64_t)layout_code << 61);
        constexpr uint32_t idesc = make_idesc<64>();
        if (ws_mode) {  // weight-stationary form, every B slice latched in collector b0 and used once
            for (int k = 0; k < kb_elems / 16; ++k) umma_ws_b0_fill(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k != 0);
        } else {
            for (int k = 0; k < kb_elems / 16; ++k) umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k != 0);
        }
        umma_commit(bar1);
    }
    __syncwarp();
    mbar_wait(bar1, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
struct TcState {
    PFN_cuTensorMapEncodeTiled encode = nullptr;
};

int tc_encode_map(fx_engine* e, CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr, CUtensorMapSwizzle sw,
                      const char* what) {
    TcState* st = static_cast<TcState*>(e->tc_state);
    CUresult r = st->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                            strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(e, FX_ERR_CUDA, std::string("cuTensorMapEncodeTiled(") + what + ") failed with CUresult " + std::to_string((int)r));
    return FX_OK;
}

template <int BN, int BK, int STAGES>
static int launch_tc(fx_engine* e, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const TcConvParams& p,
                     cudaStream_t stream) {
    using L = TcSmem<BN, BK, STAGES>;
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(tc_conv_kernel<BN, BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        attr_done[e->device & 255] = true;
    }
    const int grid = std::min(p.total_tiles + p.ds_tiles, e->sm_count);
    FX_CUDA(e, launch_pdl(tc_conv_kernel<BN, BK, STAGES>, dim3(grid), dim3(kTcThreads), L::kTotal, stream, ma, mb, mb2, p));
    FX_LAUNCH_CHECK(e, "tc_conv_kernel");
    return FX_OK;
}

template <int BN, int BK, int STAGES, bool RESB = false>
static int launch_tc2(fx_engine* e, const CUtensorMap& ma, const CUtensorMap& ma1, const CUtensorMap& mb, const CUtensorMap& mb2,
                      const CUtensorMap& mbh, TcConvParams p, cudaStream_t stream) {
    constexpr int kBBytes = (BN / 2) * BK * 2;
    constexpr int kSmem = 1024 + STAGES * 128 * BK * 2 + (RESB ? kResKb : STAGES) * kBBytes + (2 * STAGES + 4) * 8 + 32 + 2 * 512 * 4 +
                          ((RESB || (BN == 256 && STAGES == kTc2Stages256)) ? 128 + 16 * 2048 : 0);
    static_assert(!RESB || BN == 128, "the fused downsample accumulator needs 4 * BN <= 512 TMEM columns");
    static_assert(kSmem <= 232448, "tc2_conv_kernel: shared memory");
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(tc2_conv_kernel<BN, BK, STAGES, RESB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_done[e->device & 255] = true;
    }
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_g;
    const int pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;
    if (p.ds_tiles) p.ds_tiles = pair_tiles;  // the downsample conv has the same tiling, counted in pair tiles here
    static const bool fuse_on = [] {
        const char* v = getenv("FX_TC_FUSEDS");
        return !(v && v[0] == '0');
    }();
    if (RESB && p.ds_tiles && fuse_on && (p.kh & 1) && (p.kw & 1)) {  // downsample as a second accumulator of the conv1 unit
        p.ds_fused = 1;
        p.ds_tiles = 0;
    }
    int pairs = std::max(1, std::min(pair_tiles + p.ds_tiles, e->sm_count / 2));
    // split the tail wave into half-N units when it would leave more than half of the pairs idle
    const int tail = pair_tiles % pairs;
    if (!RESB && !p.ds_tiles && pair_tiles > pairs && tail > 0 && 2 * tail <= pairs && !getenv("FX_DEBUG_NO_SPLIT")) {
        p.split_tail = tail;
        p.split_full = pair_tiles - tail;
    }
    if (const char* dbg = getenv("FX_DEBUG_TC_PAIRS")) pairs = std::max(1, std::min(pairs, atoi(dbg)));  // fabric experiments only
    FX_CUDA(e, launch_pdl(tc2_conv_kernel<BN, BK, STAGES, RESB>, dim3(2 * pairs), dim3(kTcThreads), kSmem, stream, ma, ma1, mb, mb2, mbh, p));  // cluster dims are a kernel attribute
    FX_LAUNCH_CHECK(e, "tc2_conv_kernel");
    return FX_OK;
}

int tc_init(fx_engine* e) {
    TcState* st = new TcState();
    e->tc_state = st;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FX_CUDA(e, cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
        return set_error(e, FX_ERR_CUDA, "driver does not export cuTensorMapEncodeTiled");
    st->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    return FX_OK;
}

void tc_free(fx_engine* e) {
    delete static_cast<TcState*>(e->tc_state);
    e->tc_state = nullptr;
}

// Pick the power-of-two output box (Wt, Ht, Nt), Wt*Ht*Nt = 128, that wastes the fewest rows.
void choose_tile(int n, int ho, int wo, int& wt, int& ht, int& nt) {
    long long best = -1;
    for (int a = 0; a <= 7; ++a)
        for (int b = 0; a + b <= 7; ++b) {
            const int c = 7 - a - b;
            const long long tiles = (long long)((wo + (1 << a) - 1) >> a) * ((ho + (1 << b) - 1) >> b) * ((n + (1 << c) - 1) >> c);
            // fewer tiles first; then wider rows (longer contiguous TMA runs)
            if (best < 0 || tiles < best || (tiles == best && a > wt)) {
                best = tiles;
                wt = a;
                ht = b;
                nt = c;
            }
        }
}

// ds / ds_out (optional): the block's 1x1 / stride-2 downsample conv, run as extra tiles of the same launch.
int tc_conv_packed(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual,
                   __nv_bfloat16* out, float* out_f32, int n, int relu, cudaStream_t stream, const PackedLayer* ds,
                   __nv_bfloat16* ds_out) {
    const LayerGeom& g = L.g;
    if (ds) {
        const LayerGeom& d = ds->g;
        if (!(d.kh == 1 && d.kw == 1 && d.pad == 0 && d.stride == g.stride && d.cin == g.cin && d.cout == g.cout && d.hin == g.hin &&
              d.win == g.win && d.hout == g.hout && d.wout == g.wout && ds_out && !residual && !out_f32))
            return set_error(e, FX_ERR_INVALID, "tc_conv: the grouped downsample conv must be the 1x1 twin of the strided conv");
    }
    TcConvParams p;
    std::memset(&p, 0, sizeof(p));
    p.batch = n;
    p.ho = g.hout;
    p.wo = g.wout;
    p.cout = g.cout;
    p.bias = L.bias;
    p.residual = residual;
    p.out = out;
    p.out_f32 = out_f32;
    p.relu = relu;
    if (g.cout % 64 != 0) return set_error(e, FX_ERR_UNSUPPORTED, "tc_conv: cout must be a multiple of 64");
    const int bn = g.cout % 256 == 0 ? 256 : (g.cout % 128 == 0 ? 128 : 64);
    p.n_tiles_n = g.cout / bn;
    choose_tile(n, g.hout, g.wout, p.wt_log2, p.ht_log2, p.nt_log2);
    p.tiles_w = (g.wout + (1 << p.wt_log2) - 1) >> p.wt_log2;
    p.tiles_h = (g.hout + (1 << p.ht_log2) - 1) >> p.ht_log2;
    p.tiles_g = (n + (1 << p.nt_log2) - 1) >> p.nt_log2;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_g * p.n_tiles_n;

    CUtensorMap ma, ma1, mb;
    if (g.cin % 64 != 0) return set_error(e, FX_ERR_UNSUPPORTED, "tc_conv: cin must be a multiple of 64 (the stem runs on the flat kernel)");
    static const bool planes_on = [] {
        const char* v = getenv("FX_TC_S2PLANES");
        return !(v && v[0] == '0');
    }();
    // stride-2 layers on the CTA-pair kernel: parity-plane maps (TcConvParams::s2planes)
    p.s2planes = planes_on && bn >= 128 && g.stride == 2 && g.hin % 2 == 0 && g.win % 2 == 0 && g.kh == 3 && g.kw == 3 && g.pad == 1;
    static const bool staged256 = [] {
        const char* v = getenv("FX_TC_STAGED256");
        return !(v && v[0] == '0');
    }();
    p.staged256 = staged256;
    if (p.s2planes) {
        const uint64_t rowb = (uint64_t)g.win * g.cin * 2;
        const uint64_t dims[5] = {(uint64_t)g.cin, (uint64_t)g.win / 2, 2, (uint64_t)g.hin / 2, (uint64_t)n};
        const uint64_t strides[4] = {(uint64_t)g.cin * 4, rowb, 2 * rowb, (uint64_t)g.hin * rowb};
        const uint32_t box[5] = {64, 1u << p.wt_log2, 1, 1u << p.ht_log2, 1u << p.nt_log2};
        const uint32_t estr[5] = {1, 1, 1, 1, 1};
        int rc = tc_encode_map(e, &ma, in, 5, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "conv A (even columns)");
        if (rc != FX_OK) return rc;
        rc = tc_encode_map(e, &ma1, in + g.cin, 5, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "conv A (odd columns)");
        if (rc != FX_OK) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)g.cin, (uint64_t)g.win, (uint64_t)g.hin, (uint64_t)n};
        const uint64_t strides[3] = {(uint64_t)g.cin * 2, (uint64_t)g.win * g.cin * 2, (uint64_t)g.hin * g.win * g.cin * 2};
        const uint32_t box[4] = {64, (uint32_t)g.stride << p.wt_log2, (uint32_t)g.stride << p.ht_log2, 1u << p.nt_log2};
        const uint32_t estr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
        int rc = FX_OK;
        if (!p.s2planes) {
            rc = tc_encode_map(e, &ma, in, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "conv A");
            if (rc != FX_OK) return rc;
            ma1 = ma;
        }
        const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
        const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
        const uint32_t bbox[2] = {64, (uint32_t)(bn >= 128 ? bn / 2 : bn)};  // the CTA-pair kernel loads half a K-block per CTA
        const uint32_t be[2] = {1, 1};
        rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "conv B");
        if (rc != FX_OK) return rc;
    }
    p.kh = g.kh;
    p.kw = g.kw;
    p.cchunks = g.cin / 64;
    p.cw_mul = g.stride;
    p.ch_mul = g.stride;
    p.pad_w = g.pad;
    p.pad_h = g.pad;
    CUtensorMap mb2 = mb;
    if (ds) {
        const uint64_t bd[2] = {(uint64_t)ds->k_bf16, (uint64_t)g.cout};
        const uint64_t bs[1] = {(uint64_t)ds->k_bf16 * 2};
        const uint32_t bbox[2] = {64, (uint32_t)(bn >= 128 ? bn / 2 : bn)};
        const uint32_t be[2] = {1, 1};
        int rc = tc_encode_map(e, &mb2, ds->w_bf16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "downsample B");
        if (rc != FX_OK) return rc;
        p.ds_tiles = p.total_tiles;
        p.ds_bias = ds->bias;
        p.ds_out = ds_out;
    }
    CUtensorMap mbh = mb;
    if (bn >= 128) {  // quarter-of-a-K-block box for the split-tail half tiles of the CTA-pair kernel
        const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
        const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
        const uint32_t bbox[2] = {64, (uint32_t)(bn / 4)};
        const uint32_t be[2] = {1, 1};
        int rc = tc_encode_map(e, &mbh, L.w_bf16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "conv B (half tile)");
        if (rc != FX_OK) return rc;
    }
    switch (bn) {
        case 64: return launch_tc<64, 64, 8>(e, ma, mb, mb2, p, stream);
        case 128: {
            // few K-blocks (layer2.0 conv1 + its downsample): weights resident in shared memory; FX_TC_RESB=0 streams them
            static const bool resb_on = [] {
                const char* v = getenv("FX_TC_RESB");
                return !(v && v[0] == '0');
            }();
            const int n_res = p.kh * p.kw * p.cchunks + (p.ds_tiles ? p.cchunks : 0);
            if (resb_on && p.n_tiles_n == 1 && n_res <= kResKb) return launch_tc2<128, 64, 6, true>(e, ma, ma1, mb, mb2, mbh, p, stream);
            return launch_tc2<128, 64, 8>(e, ma, ma1, mb, mb2, mbh, p, stream);
        }
        default:
            // (a 4-CTA-cluster variant with the weight K-blocks multicast to two pairs was measured on B200 in round 2:
            // byte-identical rows, 64 us instead of 47.5 us per layer3 / layer4 launch -- removed)
            // fp32 output (the last conv before the average pool) and FX_TC_STAGED256=0: direct epilogue, six stages
            if (p.out_f32 || !p.staged256) return launch_tc2<256, 64, 6>(e, ma, ma1, mb, mb2, mbh, p, stream);
            return launch_tc2<256, 64, kTc2Stages256>(e, ma, ma1, mb, mb2, mbh, p, stream);
    }
}

int tc_conv(fx_engine* e, int li, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out, float* out_f32,
            int n, int relu, cudaStream_t stream) {
    return tc_conv_packed(e, e->layers[li], in, residual, out, out_f32, n, relu, stream, nullptr, nullptr);
}

// Test hook behind fx_debug_tma_probe (engine.cu): encode an arbitrary 4-D bf16 map, load one box.
int tc_tma_probe(fx_engine* e, const void* base, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                 const uint32_t* estr, int swizzle, const int* coords, int bytes, uint8_t* out_dev, cudaStream_t stream) {
    CUtensorMap m;
    CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                            : CU_TENSOR_MAP_SWIZZLE_NONE;
    int rc = tc_encode_map(e, &m, base, 4, dims, strides, box, estr, sw, "probe");
    if (rc != FX_OK) return rc;
    const int smem = 1024 + ((bytes + 15) & ~15) + 16;
    FX_CUDA(e, cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    tma_probe_kernel<<<1, 128, smem, stream>>>(m, coords[0], coords[1], coords[2], coords[3], bytes, out_dev);
    FX_LAUNCH_CHECK(e, "tma_probe_kernel");
    return FX_OK;
}

// ------------------------------------------------------------------------------------------
// MMA issue-rate microbenchmark (inspection entry point): every CTA issues `iters` rounds of
// (rowb/32) back-to-back tcgen05.mma M128 x N x K16 from fixed shared-memory tiles whose A view starts
// `shift_rows` rows into the tile and moves by `tap_stride_rows` between rounds (9 positions, like
// filter taps).  Reports SM cycles per MMA -> the operand-fetch floor for each (N, swizzle, shift).
// ------------------------------------------------------------------------------------------
template <int ROWB, int NACC>
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int n_cols, int shift_rows, int tap_stride_rows, int iters, int ws_reuse, float* cycles_per_mma) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = sbase, sB = sbase + 96 * 1024;
    const uint32_t bar = sB + 64 * 1024, tslot = bar + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x * 16; i < 160 * 1024; i += blockDim.x * 16)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sbase + i), "r"(0u) : "memory");
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tslot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    if (warp == 0) {
        constexpr int KSTEPS = ROWB / 32;
        constexpr uint64_t hi = make_smem_desc_rowb<ROWB>(0) & 0xFFFFFFFF00000000ull;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = (sA >> 4) + (uint32_t)shift_rows * (ROWB / 16);
        const uint32_t astep = (uint32_t)tap_stride_rows * (ROWB / 16);
        const uint32_t b0 = sB >> 4;
        const uint32_t bstep = (uint32_t)((n_cols * ROWB) >> 4) * (n_cols * ROWB * 8 <= 64 * 1024 ? 1u : 0u);
        long long t0 = 0, t1 = 0;
        const int rounds = iters / 8;
        if (elect_one_sync()) {
            t0 = clock64();
            for (int it = 0; it < rounds; ++it) {
#pragma unroll
                for (int tap = 0; tap < 8; ++tap) {
                    if (ws_reuse == 0) {
                        const uint32_t d = tmem_base + (uint32_t)(tap % NACC) * (uint32_t)n_cols;
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            umma_bf16(d, hi | (uint64_t)(a0 + tap * astep + 2 * k), hi | (uint64_t)(b0 + tap * bstep + 2 * k), idesc, 1);
                    } else {
                        // weight-stationary: B slice k of a tap is latched in collector b<k> by the first
                        // accumulator and re-used by the other NACC-1 (same B, different A rows and D)
                        if constexpr (KSTEPS <= 4) {
#pragma unroll
                            for (int acc = 0; acc < NACC; ++acc) {
                                const uint32_t d = tmem_base + (uint32_t)acc * (uint32_t)n_cols;
                                const uint32_t a_lo = a0 + tap * astep + acc * 128 * (ROWB / 16);
                                if (acc == 0 || ws_reuse == 1) {
                                    umma_ws<0, true>(d, hi | (uint64_t)(a_lo), hi | (uint64_t)(b0 + tap * bstep), idesc, 1);
                                    if (KSTEPS > 1) umma_ws<1, true>(d, hi | (uint64_t)(a_lo + 2), hi | (uint64_t)(b0 + tap * bstep + 2), idesc, 1);
                                    if (KSTEPS > 2) umma_ws<2, true>(d, hi | (uint64_t)(a_lo + 4), hi | (uint64_t)(b0 + tap * bstep + 4), idesc, 1);
                                    if (KSTEPS > 2) umma_ws<3, true>(d, hi | (uint64_t)(a_lo + 6), hi | (uint64_t)(b0 + tap * bstep + 6), idesc, 1);
                                } else {
                                    umma_ws<0, false>(d, hi | (uint64_t)(a_lo), hi | (uint64_t)(b0 + tap * bstep), idesc, 1);
                                    if (KSTEPS > 1) umma_ws<1, false>(d, hi | (uint64_t)(a_lo + 2), hi | (uint64_t)(b0 + tap * bstep + 2), idesc, 1);
                                    if (KSTEPS > 2) umma_ws<2, false>(d, hi | (uint64_t)(a_lo + 4), hi | (uint64_t)(b0 + tap * bstep + 4), idesc, 1);
                                    if (KSTEPS > 2) umma_ws<3, false>(d, hi | (uint64_t)(a_lo + 6), hi | (uint64_t)(b0 + tap * bstep + 6), idesc, 1);
                                }
                            }
                        }
                    }
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
        if (elect_one_sync()) cycles_per_mma[blockIdx.x] = (float)(t1 - t0) / (float)(rounds * 8 * KSTEPS * (ws_reuse ? NACC : 1));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int ROWB>
static void launch_mma_rate(int nacc, int grid, int smem, cudaStream_t stream, int n_cols, int shift_rows, int ts, int iters, int ws, float* out) {
    cudaFuncSetAttribute(mma_rate_kernel<ROWB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(mma_rate_kernel<ROWB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(mma_rate_kernel<ROWB, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (nacc >= 4)
        mma_rate_kernel<ROWB, 4><<<grid, 128, smem, stream>>>(n_cols, shift_rows, ts, iters, ws, out);
    else if (nacc >= 2)
        mma_rate_kernel<ROWB, 2><<<grid, 128, smem, stream>>>(n_cols, shift_rows, ts, iters, ws, out);
    else
        mma_rate_kernel<ROWB, 1><<<grid, 128, smem, stream>>>(n_cols, shift_rows, ts, iters, ws, out);
}

// tap_stride_rows >= 1000 encodes "round-robin over (tap_stride_rows / 1000) accumulators".
int tc_mma_rate(fx_engine* e, int n_cols, int rowb, int shift_rows, int tap_stride_rows, int iters, float* out_dev, cudaStream_t stream) {
    // encoding of tap_stride_rows: + 1000 * accumulators + 100000 * ws (1: weight-stationary, refill every MMA; 2: fill once, re-use)
    const int ws = tap_stride_rows / 100000;
    tap_stride_rows %= 100000;
    const int nacc = std::max(1, tap_stride_rows / 1000);
    tap_stride_rows %= 1000;
    if ((n_cols != 64 && n_cols != 128 && n_cols != 256) || (rowb != 128 && rowb != 64 && rowb != 32))
        return set_error(e, FX_ERR_INVALID, "mma_rate: N must be 64/128/256 and rowb 128/64/32");
    if (shift_rows < 0 || tap_stride_rows < 0 || (shift_rows + 7 * tap_stride_rows + 128 * (ws ? nacc : 1)) * rowb > 96 * 1024 || nacc * n_cols > 512)
        return set_error(e, FX_ERR_INVALID, "mma_rate: views leave the shared-memory tile");
    const int smem = 1024 + 160 * 1024 + 64;
    if (rowb == 128)
        launch_mma_rate<128>(nacc, e->sm_count, smem, stream, n_cols, shift_rows, tap_stride_rows, iters, ws, out_dev);
    else if (rowb == 64)
        launch_mma_rate<64>(nacc, e->sm_count, smem, stream, n_cols, shift_rows, tap_stride_rows, iters, ws, out_dev);
    else
        launch_mma_rate<32>(nacc, e->sm_count, smem, stream, n_cols, shift_rows, tap_stride_rows, iters, ws, out_dev);
    FX_LAUNCH_CHECK(e, "mma_rate_kernel");
    return FX_OK;
}

// Test hook behind fx_debug_umma_shift (engine.cu).
int tc_umma_shift_probe(fx_engine* e, const void* a_dev, const void* b_dev, int kb_elems, int shift_rows, int base_offset,
                        float* out_dev, cudaStream_t stream) {
    if (kb_elems != 64 && kb_elems != 32 && kb_elems != 16) return set_error(e, FX_ERR_INVALID, "umma probe: K block must be 64, 32 or 16");
    if (shift_rows < 0 || shift_rows > 128) return set_error(e, FX_ERR_INVALID, "umma probe: shift must be in [0,128]");
    if (base_offset >= 100 && base_offset != 100) return set_error(e, FX_ERR_INVALID, "umma probe: base_offset 100 selects the weight-stationary form");
    const CUtensorMapSwizzle sw = kb_elems == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kb_elems == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const int layout_code = kb_elems == 64 ? 2 : kb_elems == 32 ? 4 : 6;
    CUtensorMap ma, mb;
    const uint64_t da[2] = {(uint64_t)kb_elems, 256}, db[2] = {(uint64_t)kb_elems, 64};
    const uint64_t st[1] = {(uint64_t)kb_elems * 2};
    const uint32_t ba[2] = {(uint32_t)kb_elems, 256}, bb[2] = {(uint32_t)kb_elems, 64}, es[2] = {1, 1};
    int rc = tc_encode_map(e, &ma, a_dev, 2, da, st, ba, es, sw, "umma probe A");
    if (rc != FX_OK) return rc;
    rc = tc_encode_map(e, &mb, b_dev, 2, db, st, bb, es, sw, "umma probe B");
    if (rc != FX_OK) return rc;
    const int smem = 1024 + 320 * kb_elems * 2 + 64;
    FX_CUDA(e, cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    umma_shift_probe_kernel<<<1, 128, smem, stream>>>(ma, mb, kb_elems, shift_rows, base_offset, layout_code, out_dev);
    FX_LAUNCH_CHECK(e, "umma_shift_probe_kernel");
    return FX_OK;
}

}  // namespace fx
