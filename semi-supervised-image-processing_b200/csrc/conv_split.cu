// Tight-tolerance mode on the tensor cores: fp32-accurate implicit-GEMM convolution from SPLIT fp16 operands.
//
// The reference runs the trunk in fp32 (src/feature_extraction.py:289-291; torchvision/models/resnet.py:89-105,266-282).
// The tensor cores take 16-bit operands, so every fp32 value x travels as two fp16 numbers, hi = fp16(x) and
// lo = fp16(x - hi): 22 significant bits.  A product a*w is then three MMAs into the same accumulator,
//     a_hi*w_hi + a_hi*w_lo + a_lo*w_hi            (a_lo*w_lo ~ 2^-22 |a w| is dropped),
// each exact in the fp32 accumulator's input (fp16 x fp16 fits fp32).  Folded weights are scaled by a power of two per
// layer so that their lo parts stay fp16-normal; the epilogue undoes it exactly.
//
// What decides the structure is the ACCUMULATION, measured on B200 (tools/probe_accum.py, profiles/r02_accum_probe.log):
// tcgen05's fp32 accumulator adds with truncation -- the error is a bias toward zero that grows linearly with the number
// of chained MMAs (relative 7.6e-10 * K for zero-mean data: 4.4e-6 at K = 4608, 1.3e-5 for all-positive operands), which
// alone would eat the 1e-5 budget of the tight mode.  So the chain is cut after every 64-wide K-block: the MMA warp
// accumulates ONE K-block (12 MMAs) into a TMEM buffer, hands it to the epilogue warps, and those add it to fp32
// accumulators in REGISTERS with round-to-nearest adds while the tensor core works on the next K-block into the other
// buffer.  (Same idea as accumulating outside the tensor core in Ootomo & Yokota's error-corrected GEMM.)
//
// Layout: activations NHWC, hi plane followed by lo plane (the two together are exactly the bytes of the fp32 tensor).
// GEMM view, TMA boxes, tile choice and zero-fill padding as conv_tc.cu; one CTA per 128 x BN tile, BN = 64 / 128.
// Warp roles: 0 = TMA producer (A_hi, A_lo, W_hi, W_lo per stage), 1 = MMA issuer, 2 = TMEM allocator, 4..11 = eight
// accumulate / epilogue warps (32 accumulator rows x BN/2 columns each).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>

#include "tc_ptx.cuh"

namespace fx {

void choose_tile(int n, int ho, int wo, int& wt, int& ht, int& nt);

struct SplitParams {
    int wt_log2, ht_log2, nt_log2;
    int tiles_w, tiles_h, tiles_g, n_tiles_n, total_tiles;
    int batch, ho, wo, cout;
    int kh, kw, cchunks;
    int cw_mul, ch_mul, pad_w, pad_h;
    const float* bias;
    float unscale;           // 2^-S: the folded weights were packed as w * 2^S
    const __half* res_hi;    // residual (split), may be null
    const __half* res_lo;
    __half* out_hi;          // split output ...
    __half* out_lo;
    float* out_f32;          // ... or plain fp32 (the layer feeding the average pool)
    int relu;
};

constexpr int kSplitThreads = 128 + 8 * 32;

__device__ __forceinline__ void tmem_ld32s(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// BK: channels per K-block (64: SWIZZLE_128B rows; 16 / 32 narrower rows are possible but starve on TMA box issue).
// GROUP: K-blocks accumulated in TMEM before the hand-over (1: a chain of BK / 16 k-steps).
// PAIR: two CTAs of a cluster (launch attribute) compute one 256 x BN tile with M = 256 MMAs issued by the even CTA
// (cta_group::2, protocol of tc2_conv_kernel): each CTA loads its own 128 activation rows (hi + lo) but only HALF of the
// weight K-block (hi + lo) -- 48 KB instead of 64 KB per K-block by TMA and 6 KB instead of 8 KB per MMA out of shared
// memory, on a kernel whose 3 MMAs per operand pair make the shared-memory pipe the bound.
template <int BN, int BK, int STAGES, int GROUP, bool PAIR>
__global__ void __launch_bounds__(kSplitThreads, 1)
split_conv_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                  const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, const SplitParams p) {
    constexpr int kA = 128 * BK * 2, kB = (PAIR ? BN / 2 : BN) * BK * 2, kStage = 2 * kA + 2 * kB;
    constexpr int KSTEPS = BK / 16;
    constexpr int HC = BN / 2;  // accumulator columns per epilogue warp
    // hand-over buffers: the whole TMEM (512 columns), so that the tensor core can run a short-K tile ahead while the
    // accumulate warps are busy with the previous tile's epilogue (residual, split, stores)
    constexpr int NBUF = 512 / BN;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = sbase + STAGES * kStage;
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 8 * NBUF;
    const uint32_t tslot = tempty0 + 8 * NBUF;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));
    float* bias_s = reinterpret_cast<float*>(smem_raw + (tslot + 16 - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < p.cout; i += kSplitThreads) bias_s[i] = __ldg(p.bias + i);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    // work units: PAIR: unit u = N tile u % n_tiles_n of the M tiles 2 * (u / n_tiles_n) + {0, 1} (one per CTA of the pair)
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_g;
    const int n_units = PAIR ? ((m_tiles + 1) >> 1) * p.n_tiles_n : p.total_tiles;
    const int first_unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_ah);
        tma_prefetch_desc(&map_al);
        tma_prefetch_desc(&map_bh);
        tma_prefetch_desc(&map_bl);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full0 + 8 * i, PAIR ? 2 : 1);  // PAIR: the leader's arrive.expect_tx + the peer producer's arrive
            mbar_init(empty0 + 8 * i, 1);             // (multicast) commit
        }
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(tfull0 + 8 * i, 1);                       // (multicast) commit
            mbar_init(tempty0 + 8 * i, PAIR ? 512 : 256);       // every accumulate thread (of both CTAs)
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR)
            tmem_alloc_2cta(tslot, 512);
        else
            tmem_alloc(tslot, 512);
    }
    tc_fence_before();
    if (PAIR)
        cluster_sync_all();
    else
        __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    pdl_wait();

    const int num_kb = p.kh * p.kw * p.cchunks;

    if (warp == 0) {
        // ===== TMA producer: hi and lo planes of the activation box and of the weight K-block =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int t = first_unit; t < n_units; t += unit_step) {
                const int n_tile = t % p.n_tiles_n;
                int m_tile = PAIR ? (t / p.n_tiles_n) * 2 + (int)rank : t / p.n_tiles_n;
                const int tw = m_tile % p.tiles_w;
                m_tile /= p.tiles_w;
                const int th = m_tile % p.tiles_h;
                const int tg = m_tile / p.tiles_h;  // PAIR: >= tiles_g for the odd leftover: every load is out of bounds -> zeros
                const int w0 = (tw << p.wt_log2) * p.cw_mul - p.pad_w;
                const int h0 = (th << p.ht_log2) * p.ch_mul - p.pad_h;
                const int n0 = tg << p.nt_log2;
                int kb = 0;
                for (int r = 0; r < p.kh; ++r)
                    for (int s = 0; s < p.kw; ++s)
                        for (int cc = 0; cc < p.cchunks; ++cc, ++kb) {
                            mbar_wait(empty0 + 8 * stage, phase ^ 1);
                            const uint32_t st = sbase + stage * kStage;
                            if (PAIR) {
                                if (leader) mbar_expect_tx(full0 + 8 * stage, 2 * kStage);
                                tma_load_4d_2cta(st, &map_ah, full0 + 8 * stage, cc * BK, w0 + s, h0 + r, n0);
                                tma_load_4d_2cta(st + kA, &map_al, full0 + 8 * stage, cc * BK, w0 + s, h0 + r, n0);
                                tma_load_2d_2cta(st + 2 * kA, &map_bh, full0 + 8 * stage, kb * BK, n_tile * BN + (int)rank * (BN / 2));
                                tma_load_2d_2cta(st + 2 * kA + kB, &map_bl, full0 + 8 * stage, kb * BK, n_tile * BN + (int)rank * (BN / 2));
                                if (!leader) mbar_arrive_leader(full0 + 8 * stage);
                            } else {
                                mbar_expect_tx(full0 + 8 * stage, kStage);
                                tma_load_4d(st, &map_ah, full0 + 8 * stage, cc * BK, w0 + s, h0 + r, n0);
                                tma_load_4d(st + kA, &map_al, full0 + 8 * stage, cc * BK, w0 + s, h0 + r, n0);
                                tma_load_2d(st + 2 * kA, &map_bh, full0 + 8 * stage, kb * BK, n_tile * BN);
                                tma_load_2d(st + 2 * kA + kB, &map_bl, full0 + 8 * stage, kb * BK, n_tile * BN);
                            }
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: GROUP K-blocks (KSTEPS k-steps x 3 split terms each) per TMEM buffer, then hand it over =====
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);  // f32 accumulate, fp16 x fp16
        constexpr uint64_t desc_hi = make_smem_desc_rowb<BK * 2>(0) & 0xFFFFFFFF00000000ull;
        uint32_t stage = 0, phase = 0, g = 0;
        for (int t = first_unit; leader && t < n_units; t += unit_step) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const uint32_t buf = g % NBUF, bphase = (g / NBUF) & 1;
                const int in_group = kb % GROUP;
                if (in_group == 0) mbar_wait(tempty0 + 8 * buf, bphase ^ 1);
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const bool hand_over = in_group == GROUP - 1 || kb == num_kb - 1;
                if (elect_one_sync()) {
                    const uint32_t st = sbase + stage * kStage;
                    const uint32_t ah = st >> 4, al = (st + kA) >> 4, bh = (st + 2 * kA) >> 4, bl = (st + 2 * kA + kB) >> 4;
                    const uint32_t d = tmem_base + buf * BN;
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                        if (PAIR) {
                            umma_bf16_2cta(d, desc_hi | (uint64_t)(ah + 2 * k), desc_hi | (uint64_t)(bh + 2 * k), idesc, (in_group | k) != 0);
                            umma_bf16_2cta(d, desc_hi | (uint64_t)(ah + 2 * k), desc_hi | (uint64_t)(bl + 2 * k), idesc, 1);
                            umma_bf16_2cta(d, desc_hi | (uint64_t)(al + 2 * k), desc_hi | (uint64_t)(bh + 2 * k), idesc, 1);
                        } else {
                            umma_bf16(d, desc_hi | (uint64_t)(ah + 2 * k), desc_hi | (uint64_t)(bh + 2 * k), idesc, (in_group | k) != 0);
                            umma_bf16(d, desc_hi | (uint64_t)(ah + 2 * k), desc_hi | (uint64_t)(bl + 2 * k), idesc, 1);
                            umma_bf16(d, desc_hi | (uint64_t)(al + 2 * k), desc_hi | (uint64_t)(bh + 2 * k), idesc, 1);
                        }
                    }
                    if (PAIR) {
                        umma_commit_2cta(empty0 + 8 * stage);
                        if (hand_over) umma_commit_2cta(tfull0 + 8 * buf);
                    } else {
                        umma_commit(empty0 + 8 * stage);
                        if (hand_over) umma_commit(tfull0 + 8 * buf);
                    }
                }
                __syncwarp();
                if (hand_over) ++g;
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== accumulate (round-to-nearest, in registers) + epilogue =====
        const int q = warp & 3;            // TMEM lane quarter
        const int half = (warp - 4) >> 2;  // which half of the tile's BN columns
        const int row = q * 32 + lane;
        const int wl = row & ((1 << p.wt_log2) - 1);
        const int hl = (row >> p.wt_log2) & ((1 << p.ht_log2) - 1);
        const int nl = row >> (p.wt_log2 + p.ht_log2);
        uint32_t g = 0;
        // BN = 64 (layer 1, the stem): the residual of the tile is fetched into registers BEFORE the K loop, so its latency
        // hides behind the tile's MMAs; with 64 columns per warp (BN = 128) the registers are not there and it is read at the end.
        constexpr bool kPrefetchRes = HC == 32;
        for (int t = first_unit; t < n_units; t += unit_step) {
            const int n_tile = t % p.n_tiles_n;
            int m_tile = PAIR ? (t / p.n_tiles_n) * 2 + (int)rank : t / p.n_tiles_n;
            const int tw = m_tile % p.tiles_w;
            m_tile /= p.tiles_w;
            const int th = m_tile % p.tiles_h;
            const int tg = m_tile / p.tiles_h;
            const int ow = (tw << p.wt_log2) + wl, oh = (th << p.ht_log2) + hl, img = (tg << p.nt_log2) + nl;
            const bool valid = ow < p.wo && oh < p.ho && img < p.batch;
            const size_t pix = ((size_t)img * p.ho + oh) * p.wo + ow;
            const int c0 = n_tile * BN + half * HC;
            const size_t obase = pix * p.cout + c0;
            uint4 pre_h[kPrefetchRes ? HC / 8 : 1], pre_l[kPrefetchRes ? HC / 8 : 1];
            if (kPrefetchRes && p.res_hi && valid) {
#pragma unroll
                for (int j = 0; j < HC / 8; ++j) {
                    pre_h[j] = __ldg(reinterpret_cast<const uint4*>(p.res_hi + obase) + j);
                    pre_l[j] = __ldg(reinterpret_cast<const uint4*>(p.res_lo + obase) + j);
                }
            }
            float acc[HC];
#pragma unroll
            for (int j = 0; j < HC; ++j) acc[j] = 0.f;
            for (int ho = 0; ho < (num_kb + GROUP - 1) / GROUP; ++ho, ++g) {
                const uint32_t buf = g % NBUF, bphase = (g / NBUF) & 1;
                mbar_wait(tfull0 + 8 * buf, bphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + buf * BN + half * HC + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int c = 0; c < HC; c += 32) {
                    uint32_t v[32];
                    tmem_ld32s(taddr + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                if (PAIR)
                    mbar_arrive_leader(tempty0 + 8 * buf);
                else
                    mbar_arrive(tempty0 + 8 * buf);
            }
            if (!valid) continue;
#pragma unroll
            for (int c = 0; c < HC; c += 8) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf(acc[c + j], p.unscale, bias_s[c0 + c + j]);
                if (p.res_hi) {
                    uint4 rh, rl;
                    if (kPrefetchRes) {
                        rh = pre_h[kPrefetchRes ? c / 8 : 0];
                        rl = pre_l[kPrefetchRes ? c / 8 : 0];
                    } else {
                        rh = __ldg(reinterpret_cast<const uint4*>(p.res_hi + obase + c));
                        rl = __ldg(reinterpret_cast<const uint4*>(p.res_lo + obase + c));
                    }
                    const __half2* h2 = reinterpret_cast<const __half2*>(&rh);
                    const __half2* l2 = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a = __half22float2(h2[j]), b = __half22float2(l2[j]);
                        f[2 * j] += a.x + b.x;  // hi + lo is the residual's fp32 value (exact: 22 bits)
                        f[2 * j + 1] += a.y + b.y;
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (p.out_f32) {
                    float4* op = reinterpret_cast<float4*>(p.out_f32 + obase + c);
                    op[0] = make_float4(f[0], f[1], f[2], f[3]);
                    op[1] = make_float4(f[4], f[5], f[6], f[7]);
                } else {
                    uint4 oh4, ol4;
                    __half2* ph = reinterpret_cast<__half2*>(&oh4);
                    __half2* pl = reinterpret_cast<__half2*>(&ol4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                        const float2 hf = __half22float2(h);
                        ph[j] = h;
                        pl[j] = __floats2half2_rn(f[2 * j] - hf.x, f[2 * j + 1] - hf.y);
                    }
                    *reinterpret_cast<uint4*>(p.out_hi + obase + c) = oh4;
                    *reinterpret_cast<uint4*>(p.out_lo + obase + c) = ol4;
                }
            }
        }
    }

    tc_fence_before();
    if (PAIR)
        cluster_sync_all();  // the peer's shared memory / barriers stay alive until both CTAs are done
    else
        __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR)
            tmem_dealloc_2cta(tmem_base, 512);
        else
            tmem_dealloc(tmem_base, 512);
    }
}

// ---- fp32 <-> split helpers ---------------------------------------------------------------------
__global__ void f32_to_split_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, size_t count4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(in)[i];
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        reinterpret_cast<__half2*>(hi)[2 * i] = h0;
        reinterpret_cast<__half2*>(hi)[2 * i + 1] = h1;
        reinterpret_cast<__half2*>(lo)[2 * i] = __floats2half2_rn(v.x - f0.x, v.y - f0.y);
        reinterpret_cast<__half2*>(lo)[2 * i + 1] = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    }
}

__global__ void split_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out, size_t count2) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count2; i += (size_t)gridDim.x * blockDim.x) {
        const float2 a = __half22float2(reinterpret_cast<const __half2*>(hi)[i]), b = __half22float2(reinterpret_cast<const __half2*>(lo)[i]);
        reinterpret_cast<float2*>(out)[i] = make_float2(a.x + b.x, a.y + b.y);
    }
}

int f32_to_split(fx_engine* e, const float* in, void* out_split, size_t count, cudaStream_t stream) {
    if (!count) return FX_OK;
    if (count % 4) return set_error(e, FX_ERR_INVALID, "f32_to_split: element count must be a multiple of 4");
    __half* hi = static_cast<__half*>(out_split);
    f32_to_split_kernel<<<(int)std::min<size_t>((count / 4 + 255) / 256, (size_t)e->sm_count * 16), 256, 0, stream>>>(in, hi, hi + count, count / 4);
    FX_LAUNCH_CHECK(e, "f32_to_split_kernel");
    return FX_OK;
}

int split_to_f32(fx_engine* e, const void* in_split, float* out, size_t count, cudaStream_t stream) {
    if (!count) return FX_OK;
    if (count % 2) return set_error(e, FX_ERR_INVALID, "split_to_f32: element count must be even");
    const __half* hi = static_cast<const __half*>(in_split);
    split_to_f32_kernel<<<(int)std::min<size_t>((count / 2 + 255) / 256, (size_t)e->sm_count * 16), 256, 0, stream>>>(hi, hi + count, out, count / 2);
    FX_LAUNCH_CHECK(e, "split_to_f32_kernel");
    return FX_OK;
}

// Pack the folded fp32 weights [cout][K] as scaled hi / lo fp16 planes; returns log2 of the scale.
int split_pack_weights(const std::vector<float>& w, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo) {
    float mx = 0.f;
    for (float v : w) mx = std::max(mx, std::fabs(v));
    int s = 0;
    if (mx > 0.f) {
        int ex;
        std::frexp(mx, &ex);  // mx = m * 2^ex, m in [0.5, 1)
        s = 12 - ex;          // largest |w| * 2^s in [2^11, 2^12): far from fp16's 65504, lo parts of weights 10^5 x smaller still normal
    }
    hi.resize(w.size());
    lo.resize(w.size());
    for (size_t i = 0; i < w.size(); ++i) {
        const float x = std::ldexp(w[i], s);
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        hi[i] = __half_as_ushort(h);
        lo[i] = __half_as_ushort(l);
    }
    return s;
}

template <int BN, int BK, int STAGES, int GROUP, bool PAIR = false>
static int launch_split(fx_engine* e, const CUtensorMap& mah, const CUtensorMap& mal, const CUtensorMap& mbh, const CUtensorMap& mbl,
                        const SplitParams& p, cudaStream_t stream) {
    constexpr int kSmem = 1024 + STAGES * (2 * 128 * BK * 2 + 2 * (PAIR ? BN / 2 : BN) * BK * 2) + (2 * STAGES + 2 * (512 / BN)) * 8 + 32 + 512 * 4;
    static_assert(kSmem <= 232448, "split_conv_kernel: shared memory");
    auto kernel = split_conv_kernel<BN, BK, STAGES, GROUP, PAIR>;
    static bool attr_done[256] = {};
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_done[e->device & 255] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
    if (PAIR) {
        const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_g;
        const int n_units = ((m_tiles + 1) / 2) * p.n_tiles_n;
        cfg.gridDim = dim3(2 * std::max(1, std::min(n_units, e->sm_count / 2)));
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    } else {
        cfg.gridDim = dim3(std::min(p.total_tiles, e->sm_count));
    }
    cfg.blockDim = dim3(kSplitThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = stream;
    cfg.attrs = attr;
    FX_CUDA(e, cudaLaunchKernelEx(&cfg, kernel, mah, mal, mbh, mbl, p));
    FX_LAUNCH_CHECK(e, "split_conv_kernel");
    return FX_OK;
}

// One conv + folded BN (+ residual, + ReLU) on split activations: `in` / `residual` / `out` are hi plane followed by lo
// plane; out_f32 (optional) receives plain fp32 instead.
int split_conv(fx_engine* e, const PackedLayer& L, const void* in, const void* residual, void* out, float* out_f32, int n, int relu,
               cudaStream_t stream) {
    const LayerGeom& g = L.g;
    if (!L.w_h16 || !L.w_l16) return set_error(e, FX_ERR_STATE, "split_conv: the layer has no split weight pack");
    if (g.cin % 64 != 0 || g.cout % 64 != 0) return set_error(e, FX_ERR_UNSUPPORTED, "split_conv: cin and cout must be multiples of 64");
    SplitParams p;
    std::memset(&p, 0, sizeof(p));
    const size_t in_count = (size_t)n * g.hin * g.win * g.cin, out_count = (size_t)n * g.hout * g.wout * g.cout;
    const __half* in_hi = static_cast<const __half*>(in);
    p.batch = n;
    p.ho = g.hout;
    p.wo = g.wout;
    p.cout = g.cout;
    p.bias = L.bias;
    p.unscale = std::ldexp(1.0f, -L.w_scale_log2);
    p.res_hi = static_cast<const __half*>(residual);
    p.res_lo = residual ? p.res_hi + out_count : nullptr;
    p.out_hi = static_cast<__half*>(out);
    p.out_lo = out ? p.out_hi + out_count : nullptr;
    p.out_f32 = out_f32;
    p.relu = relu;
    const int bn = g.cout % 128 == 0 ? 128 : 64;
    p.n_tiles_n = g.cout / bn;
    choose_tile(n, g.hout, g.wout, p.wt_log2, p.ht_log2, p.nt_log2);
    p.tiles_w = (g.wout + (1 << p.wt_log2) - 1) >> p.wt_log2;
    p.tiles_h = (g.hout + (1 << p.ht_log2) - 1) >> p.ht_log2;
    p.tiles_g = (n + (1 << p.nt_log2) - 1) >> p.nt_log2;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_g * p.n_tiles_n;
    p.kh = g.kh;
    p.kw = g.kw;
    p.cchunks = g.cin / 64;
    p.cw_mul = p.ch_mul = g.stride;
    p.pad_w = p.pad_h = g.pad;
    // 16-bit elements either way: the bf16 tensor-map type moves fp16 bytes unchanged (no arithmetic, zero fill is zero)
    CUtensorMap mah, mal, mbh, mbl;
    const uint64_t dims[4] = {(uint64_t)g.cin, (uint64_t)g.win, (uint64_t)g.hin, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)g.cin * 2, (uint64_t)g.win * g.cin * 2, (uint64_t)g.hin * g.win * g.cin * 2};
    const uint32_t box[4] = {64, (uint32_t)g.stride << p.wt_log2, (uint32_t)g.stride << p.ht_log2, 1u << p.nt_log2};
    const uint32_t estr[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    int rc = tc_encode_map(e, &mah, in_hi, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "split conv A hi");
    if (rc == FX_OK) rc = tc_encode_map(e, &mal, in_hi + in_count, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "split conv A lo");
    const uint64_t K = (uint64_t)g.kh * g.kw * g.cin;
    const uint64_t bd[2] = {K, (uint64_t)g.cout};
    const uint64_t bs[1] = {K * 2};
    // N = 128 layers run as CTA pairs (half of every weight K-block per CTA); FX_SPLIT_PAIR=0 keeps the single-CTA kernel
    static const bool pair_on = [] {
        const char* v = getenv("FX_SPLIT_PAIR");
        return !(v && v[0] == '0');
    }();
    const bool pair = bn == 128 && pair_on;
    const uint32_t bbox[2] = {64, (uint32_t)(pair ? bn / 2 : bn)};
    const uint32_t be[2] = {1, 1};
    if (rc == FX_OK) rc = tc_encode_map(e, &mbh, L.w_h16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "split conv W hi");
    if (rc == FX_OK) rc = tc_encode_map(e, &mbl, L.w_l16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "split conv W lo");
    if (rc != FX_OK) return rc;
    if (pair) return launch_split<128, 64, 4, 1, true>(e, mah, mal, mbh, mbl, p, stream);
    if (bn == 128) return launch_split<128, 64, 3, 1>(e, mah, mal, mbh, mbl, p, stream);
    return launch_split<64, 64, 4, 1>(e, mah, mal, mbh, mbl, p, stream);
}

// ---- the stem in the tight mode ------------------------------------------------------------------
// conv1 7x7 / stride 2 / pad 3 + bn1 + ReLU (torchvision/models/resnet.py:197-199, 268-270) as a 4x4 / stride-1 conv over the
// SPACE-TO-DEPTH form of the normalised crop (fx_common.cuh: [n][115][116][16], channel (dy*2+dx)*3+c), split fp16, with the
// four horizontal taps gathered into the channel dimension: four 64-wide K-blocks.  Output: split [n][112][112][64].

// fp32 staging tensor [n][230][232][4] (what fx_preprocess writes for FX_PRECISION_FP32) -> split planes of the
// X-GATHERED space-to-depth crop [n][115][112][64]: position (Y, X) holds the four s2d pixels (Y, X .. X+3), i.e. channel
// s*16 + (dy*2+dx)*3 + c = padded pixel (2Y+dy, 2(X+s)+dx), channel c.  With the four horizontal taps in the channel
// dimension a K-block is one 128-byte row per pixel (SWIZZLE_128B, the box shape every other layer uses) and the stem is
// a 4x1 conv: 4 TMA boxes per operand and tile instead of 16 boxes of 32-byte rows, which the TMA unit could not feed
// fast enough (measured: 1.08 ms per launch at batch 256 with per-tap 16-channel boxes).
constexpr int kXsW = 112, kXsC = 64;
__global__ void in0_to_split_xs2d_kernel(const float* __restrict__ in0, __half* __restrict__ hi, __half* __restrict__ lo, int n) {
    // one thread per (position, s): 16 channels = one 32-byte piece of each plane
    const size_t total = (size_t)n * kS2dH * kXsW * 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i & 3);
        const size_t pos = i >> 2;
        const size_t img = pos / (kS2dH * kXsW);
        const int rem = (int)(pos - img * kS2dH * kXsW);
        const int Y = rem / kXsW, X = rem - Y * kXsW + s;  // s2d pixel (Y, X), X <= 114
        float v[16];
#pragma unroll
        for (int j = 12; j < 16; ++j) v[j] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const float4 px = *reinterpret_cast<const float4*>(in0 + ((img * kIn0H + (2 * Y + dy)) * kIn0W + (2 * X + dx)) * kIn0C);
                v[(dy * 2 + dx) * 3 + 0] = px.x;
                v[(dy * 2 + dx) * 3 + 1] = px.y;
                v[(dy * 2 + dx) * 3 + 2] = px.z;
            }
        uint4 oh[2], ol[2];
        __half2* ph = reinterpret_cast<__half2*>(oh);
        __half2* pl = reinterpret_cast<__half2*>(ol);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            const float2 hf = __half22float2(h);
            ph[j] = h;
            pl[j] = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        }
        reinterpret_cast<uint4*>(hi)[2 * i] = oh[0];
        reinterpret_cast<uint4*>(hi)[2 * i + 1] = oh[1];
        reinterpret_cast<uint4*>(lo)[2 * i] = ol[0];
        reinterpret_cast<uint4*>(lo)[2 * i + 1] = ol[1];
    }
}

// 3x3 / stride-2 / pad-1 max-pool (resnet.py:200) on a split tensor: max over hi + lo, re-split.  8 channels per thread.
__global__ void maxpool_split_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, __half* __restrict__ out_hi,
                                     __half* __restrict__ out_lo, int n, int h, int w, int c) {
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1, cv = c / 8;
    const size_t total = (size_t)n * ho * wo * cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i % cv);
        size_t q = i / cv;
        const int x = (int)(q % wo);
        q /= wo;
        const int y = (int)(q % ho);
        const int img = (int)(q / ho);
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * y - 1 + dy;
            if (iy < 0 || iy >= h) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * x - 1 + dx;
                if (ix < 0 || ix >= w) continue;
                const size_t off = (((size_t)img * h + iy) * w + ix) * c + (size_t)cc * 8;
                const uint4 a = *reinterpret_cast<const uint4*>(in_hi + off), b = *reinterpret_cast<const uint4*>(in_lo + off);
                const __half2* ha = reinterpret_cast<const __half2*>(&a);
                const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 fa = __half22float2(ha[k]), fb = __half22float2(hb[k]);
                    m[2 * k] = fmaxf(m[2 * k], fa.x + fb.x);
                    m[2 * k + 1] = fmaxf(m[2 * k + 1], fa.y + fb.y);
                }
            }
        }
        uint4 oh, ol;
        __half2* ph = reinterpret_cast<__half2*>(&oh);
        __half2* pl = reinterpret_cast<__half2*>(&ol);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 hh = __floats2half2_rn(m[2 * k], m[2 * k + 1]);
            const float2 hf = __half22float2(hh);
            ph[k] = hh;
            pl[k] = __floats2half2_rn(m[2 * k] - hf.x, m[2 * k + 1] - hf.y);
        }
        const size_t oo = (((size_t)img * ho + y) * wo + x) * c + (size_t)cc * 8;
        *reinterpret_cast<uint4*>(out_hi + oo) = oh;
        *reinterpret_cast<uint4*>(out_lo + oo) = ol;
    }
}

// Space-to-depth pack of the folded stem weights [64][7][7][3] -> [64][4*4*16] (as engine.cu's bf16 stem pack), fp32.
void split_stem_s2d_weights(const std::vector<float>& host_w, int cout, std::vector<float>& out) {
    out.assign((size_t)cout * 16 * kS2dC, 0.f);
    for (int o = 0; o < cout; ++o)
        for (int r = 0; r < 7; ++r)
            for (int s = 0; s < 7; ++s)
                for (int i = 0; i < 3; ++i)
                    out[(size_t)o * 16 * kS2dC + ((r >> 1) * 4 + (s >> 1)) * kS2dC + ((r & 1) * 2 + (s & 1)) * 3 + i] =
                        host_w[((size_t)o * 49 + r * 7 + s) * 3 + i];
}

size_t split_stem_scratch_bytes(int n) { return (size_t)n * kS2dH * kXsW * kXsC * 2 * 2; }

// in0_f32: the fp32 staging tensor of the n staged images; xs2d_split: scratch for its x-gathered split form
// (split_stem_scratch_bytes); conv_out_split: [n][112][112][64] split; pooled_split: [n][56][56][64] split.
int split_stem(fx_engine* e, const PackedLayer& L, const float* in0_f32, void* xs2d_split, void* conv_out_split, void* pooled_split, int n,
               cudaStream_t stream) {
    if (!L.w_h16 || !L.w_l16 || L.g.cout != 64) return set_error(e, FX_ERR_STATE, "split_stem: the stem has no split weight pack");
    const size_t xs_count = (size_t)n * kS2dH * kXsW * kXsC, conv_count = (size_t)n * 112 * 112 * 64, pool_count = (size_t)n * 56 * 56 * 64;
    __half* s_hi = static_cast<__half*>(xs2d_split);
    {
        const size_t items = (size_t)n * kS2dH * kXsW * 4;
        in0_to_split_xs2d_kernel<<<(int)std::min<size_t>((items + 255) / 256, (size_t)e->sm_count * 32), 256, 0, stream>>>(in0_f32, s_hi, s_hi + xs_count, n);
        FX_LAUNCH_CHECK(e, "in0_to_split_xs2d_kernel");
    }
    SplitParams p;
    std::memset(&p, 0, sizeof(p));
    p.batch = n;
    p.ho = p.wo = 112;
    p.cout = 64;
    p.bias = L.bias;
    p.unscale = std::ldexp(1.0f, -L.w_scale_log2);
    p.out_hi = static_cast<__half*>(conv_out_split);
    p.out_lo = p.out_hi + conv_count;
    p.relu = 1;
    p.n_tiles_n = 1;
    choose_tile(n, 112, 112, p.wt_log2, p.ht_log2, p.nt_log2);
    p.tiles_w = (112 + (1 << p.wt_log2) - 1) >> p.wt_log2;
    p.tiles_h = (112 + (1 << p.ht_log2) - 1) >> p.ht_log2;
    p.tiles_g = (n + (1 << p.nt_log2) - 1) >> p.nt_log2;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_g;
    p.kh = 4;  // filter-row pairs; the four horizontal taps live in the 64 channels
    p.kw = 1;
    p.cchunks = 1;
    p.cw_mul = p.ch_mul = 1;
    CUtensorMap mah, mal, mbh, mbl;
    const uint64_t dims[4] = {(uint64_t)kXsC, (uint64_t)kXsW, (uint64_t)kS2dH, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)kXsC * 2, (uint64_t)kXsW * kXsC * 2, (uint64_t)kS2dH * kXsW * kXsC * 2};
    const uint32_t box[4] = {64, 1u << p.wt_log2, 1u << p.ht_log2, 1u << p.nt_log2};
    const uint32_t estr[4] = {1, 1, 1, 1};
    int rc = tc_encode_map(e, &mah, s_hi, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "split stem A hi");
    if (rc == FX_OK) rc = tc_encode_map(e, &mal, s_hi + xs_count, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "split stem A lo");
    const uint64_t bd[2] = {256, 64};  // [cout][(r, s, 16 channels)]: K-block r = the 64 weights of filter-row pair r
    const uint64_t bs[1] = {512};
    const uint32_t bbox[2] = {64, 64};
    const uint32_t be[2] = {1, 1};
    if (rc == FX_OK) rc = tc_encode_map(e, &mbh, L.w_h16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "split stem W hi");
    if (rc == FX_OK) rc = tc_encode_map(e, &mbl, L.w_l16, 2, bd, bs, bbox, be, CU_TENSOR_MAP_SWIZZLE_128B, "split stem W lo");
    if (rc != FX_OK) return rc;
    if ((rc = launch_split<64, 64, 4, 1>(e, mah, mal, mbh, mbl, p, stream)) != FX_OK) return rc;
    __half* o_hi = static_cast<__half*>(pooled_split);
    const size_t items = pool_count / 8;
    maxpool_split_kernel<<<(int)std::min<size_t>((items + 255) / 256, (size_t)e->sm_count * 32), 256, 0, stream>>>(p.out_hi, p.out_lo, o_hi,
                                                                                                                    o_hi + pool_count, n, 112, 112, 64);
    FX_LAUNCH_CHECK(e, "maxpool_split_kernel");
    return FX_OK;
}

}  // namespace fx
