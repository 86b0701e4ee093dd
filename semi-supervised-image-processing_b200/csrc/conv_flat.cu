// Weight-stationary, halo-tile implicit-GEMM convolution on tcgen05 / TMEM ("flat" kernel).
//
// Serves the stride-1 3x3 convolutions of layer1 / layer2 and (in space-to-depth form) the 7x7/s2
// stem of the frozen ResNet-18 the reference runs at src/feature_extraction.py:290-291
// (torchvision/models/resnet.py:89-105,197-206,266-282).  Why it exists: feeding each filter tap
// with its own TMA box (conv_tc.cu) re-reads every input pixel kh*kw times from L2 and re-reads the
// weights for every output tile; for the Cout=64/128 layers that traffic, not the tensor pipe, is
// the bound (profiles/r01_launches_v0.md).  Here
//   * a CTA owns one 64-wide slice of output channels for the whole launch and keeps that slice's
//     folded weights resident in shared memory (one TMA burst at start: 72 KB layer1, 144 KB
//     layer2, 32 KB stem);
//   * the input is streamed ONCE per tile: a TMA box brings R+KH-1 input rows, each P = W+KW-1
//     pixels wide (the conv zero padding is TMA out-of-bounds fill), for one 64-channel chunk.  In
//     shared memory that box is a flat list of pixels, one swizzled row of ROWB bytes each;
//   * output pixel (i, x) of the tile gets flat index m = i*P + x; filter tap (r, s) of that pixel
//     is smem row m + r*P + s.  So the A operand of tap (r, s) for 128 consecutive m is the SAME
//     tile viewed r*P + s rows further down: one UMMA descriptor whose start address is shifted.
//     (The swizzle XOR is a function of the absolute smem address, so a row-shifted view of a
//     TMA-written tile is a valid operand -- pinned by tests via fx_debug_umma_shift.)  No im2col,
//     no per-tap loads; x >= W positions are junk rows that are never stored (W/P efficiency);
//   * accumulators are 64-column TMEM slots in a ring of 8, so the epilogue of one 128-pixel
//     M-tile overlaps the MMAs of the next ones; 8 epilogue warps (+bias, +residual, ReLU, bf16);
//   * consecutive launches walk their tiles in opposite directions (TileWalk), the residual of the next M-tile is
//     fetched into registers while the current one is processed, the layer-1 bias is a constant-bank operand.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue.
#include <algorithm>
#include <cstring>

#include "tc_ptx.cuh"

namespace fx {

struct FlatParams {
    int P, W, H;           // smem row pitch (pixels), valid output width / height
    int R;                 // output rows computed per work tile
    int rstep, yfirst;     // tile t covers output rows [t*rstep + yfirst, +R)  (pooled stem: rstep 8, yfirst -1)
    int tiles_per_img, n_work;
    int chunks;            // 64-channel chunks of the input (1 for the stem)
    int ns;                // cout / 64 output-channel slices
    int x0, ypad;          // TMA box origin: x = x0, y = y0 - ypad
    int cout;
    int nstages, stage_bytes, box_bytes, w_bytes, slack_bytes;
    // The folded bias travels IN the kernel parameters (constant bank): the epilogue's FADDs take it as a constant operand.
    // Read from shared memory it cost 32 of the epilogue's 92 (153 with a residual) shared-memory wavefronts per warp and
    // M-tile (a broadcast LDS.128 is two wavefronts), on the data pipe the MMA operand fetch saturates.  Needs cout == 64.
    float bias_v[64];
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    int relu;
    int sched;             // tile walk: bit 0 = descending, bit 1 = strided over the CTAs (see TileWalk)
};

constexpr int kFlatThreads = 384;
constexpr int kFlatPoolThreads = 128 + 16 * 32;  // pooled stem: 8 epilogue warps + 8 pool warps
constexpr int kPoolRing = 6;                     // conv rows kept in shared memory for the fused max-pool
constexpr int kFlatSlots = 8;  // 8 x 64 fp32 columns = the whole TMEM
constexpr int kEpiBytes = 8 * 4096 + 256;      // epilogue staging (8 warps x 4 KB) + bias
constexpr int kPoolRingBytes = 6 * 112 * 128;  // pooled stem: six conv rows (112 px x 64 ch bf16)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// Read-once global data (no L1 allocation).  Deliberately unpredicated: behind a predicated load -- a `live ? load : 0`
// select, or a guarded in/out asm operand -- ptxas copies the value out of a scratch register at once, i.e. waits for
// the load right there, which defeats a prefetch.  Callers clamp the address instead.
__device__ __forceinline__ uint4 ldg_stream128(const void* gptr) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(gptr));
    return r;
}
// Which work items a CTA (or CTA pair) processes, and in which order.  The activations of the stem and of layer 1 at
// batch 256 (103 MB per tensor) do not fit in L2 next to the tensors written meanwhile, so a kernel that walks its
// input in the order the previous kernel wrote it finds nothing of it in L2 (LRU: the oldest lines went first).
// Consecutive launches therefore walk in OPPOSITE directions: what the previous kernel touched last is read first.
//   sched bit 1 clear: a contiguous run of items per CTA (item = begin + i);   set: items strided over the CTAs, so
//   that the whole grid sweeps the tensor front to back in time;   bit 0: the same items in descending order.
struct TileWalk {
    int begin, step, count;
    __device__ __forceinline__ TileWalk(int n_items, int cta, int n_cta, int sched) {
        if (sched & 2) {
            count = cta < n_items ? (n_items - cta + n_cta - 1) / n_cta : 0;
            begin = (sched & 1) ? n_items - 1 - cta : cta;
            step = (sched & 1) ? -n_cta : n_cta;
        } else {
            const int first = (int)((long long)n_items * cta / n_cta), last = (int)((long long)n_items * (cta + 1) / n_cta);
            count = last - first;
            begin = (sched & 1) ? last - 1 : first;
            step = (sched & 1) ? -1 : 1;
        }
    }
};

// POOL (stem only): the epilogue keeps the ReLU'd conv rows of the tile in a shared-memory ring and
// writes the 3x3 / stride-2 / pad-1 max-pooled rows (torchvision/models/resnet.py:200) instead of the
// conv output: work tile = 4 pooled rows = 9 conv rows (one recomputed), `out` is [n][H/2][W/2][64].
template <int ROWB, int KH, int KW, bool POOL>
__global__ void __launch_bounds__(POOL ? kFlatPoolThreads : kFlatThreads, 1)
flat_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const FlatParams p) {
    constexpr int TAPS = KH * KW;
    constexpr int KSTEPS = ROWB / 32;  // K=16 bf16 MMAs per smem row
    constexpr int WTILE = 64 * ROWB;   // one (tap, chunk) weight tile: 64 cout rows
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sW = sbase;
    const uint32_t sA = sbase + p.w_bytes;
    const uint32_t stage0 = sA + p.nstages * p.stage_bytes + p.slack_bytes;  // epilogue staging / pooled-stem conv-row ring
    const uint32_t bias0 = stage0 + (POOL ? kPoolRingBytes : 8 * 4096);      // 64 fp32 bias values of this slice
    const uint32_t bars = bias0 + 256;
    const uint32_t full0 = bars, empty0 = full0 + 8 * p.nstages, tfull0 = empty0 + 8 * p.nstages;
    const uint32_t tempty0 = tfull0 + 8 * kFlatSlots, wbar = tempty0 + 8 * kFlatSlots, tslot = wbar + 8;
    const uint32_t mdone0 = tslot + 8, pool_sync0 = mdone0 + 8 * kFlatSlots;  // pooled stem only
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const int nslice = blockIdx.x % p.ns;
    // contiguous run of work tiles per CTA (balances the short last tile of each image, keeps halo rows in L2)
    const int n_cta = gridDim.x / p.ns, cta = blockIdx.x / p.ns;
    const TileWalk walk(p.n_work, cta, n_cta, p.sched);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nstages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < kFlatSlots; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, 128);
            if (POOL) mbar_init(mdone0 + 8 * i, 128);
        }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tslot, 512);
    if (warp == 3) {
        if (POOL && lane == 0) *reinterpret_cast<int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw))) = 0;
        if (POOL) {  // the pooled stem keeps its bias in shared memory (constant operands measured 4 us slower there)
            float* bs = reinterpret_cast<float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
            bs[lane] = p.bias_v[lane];
            bs[lane + 32] = p.bias_v[lane + 32];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;

    // the resident weight slice is constant data: fetch it before waiting for the previous kernel (PDL)
    if (warp == 0 && elect_one_sync()) {
        mbar_expect_tx(wbar, TAPS * p.chunks * WTILE);
        for (int kb = 0; kb < TAPS * p.chunks; ++kb) tma_load_2d(sW + kb * WTILE, &map_b, wbar, kb * (ROWB / 2), nslice * 64);
    }
    __syncwarp();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer: one halo box per (tile, chunk) =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int wi = 0, w = walk.begin; wi < walk.count; ++wi, w += walk.step) {
                const int img = w / p.tiles_per_img;
                const int y0 = (w - img * p.tiles_per_img) * p.rstep + p.yfirst;
                for (int c = 0; c < p.chunks; ++c) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(full0 + 8 * stage, p.box_bytes);
                    tma_load_4d(sA + stage * p.stage_bytes, &map_a, full0 + 8 * stage, c * 64, p.x0, y0 - p.ypad, img);
                    if (++stage == (uint32_t)p.nstages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the (uniform) schedule, one elected lane issues =====
        constexpr uint32_t idesc = make_idesc<64>();
        constexpr uint64_t desc_hi = make_smem_desc_rowb<ROWB>(0) & 0xFFFFFFFF00000000ull;
        mbar_wait(wbar, 0);
        tc_fence_after();
        uint32_t stage = 0, phase = 0, g_base = 0;
        const uint32_t w_lo = sW >> 4;
        const uint32_t row_units = ROWB / 16;  // descriptor address units (16 B) per smem row
        for (int wi = 0, w = walk.begin; wi < walk.count; ++wi, w += walk.step) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * p.rstep + p.yfirst;
            const int rows_valid = min(p.R, p.H - y0);
            const int n_mt = ((rows_valid - 1) * p.P + p.W - 1) / 128 + 1;
            for (int c = 0; c < p.chunks; ++c) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t a_lo_stage = (sA + stage * p.stage_bytes) >> 4;
                // M-tiles are issued in PAIRS with the taps interleaved (tap-outer, M-tile-inner): back-to-back
                // MMAs into one accumulator run at 56-72 cycles, alternating accumulators at the 48-cycle
                // operand-fetch floor (profiles/r01_mma_ws_and_accumulator_probe.log).
                for (int mt = 0; mt < n_mt; mt += 2) {
                    const bool two = mt + 1 < n_mt;
                    const uint32_t g0 = g_base + mt, slot0 = g0 & (kFlatSlots - 1), use0 = g0 / kFlatSlots;
                    const uint32_t g1 = g0 + 1, slot1 = g1 & (kFlatSlots - 1), use1 = g1 / kFlatSlots;
                    if (c == 0) {
                        mbar_wait(tempty0 + 8 * slot0, (use0 & 1) ^ 1);
                        if (two) mbar_wait(tempty0 + 8 * slot1, (use1 & 1) ^ 1);
                        tc_fence_after();
                    }
                    if (elect_one_sync()) {
                        const uint32_t d0 = tmem_base + slot0 * 64, d1 = tmem_base + slot1 * 64;
                        const uint32_t a_lo_mt = a_lo_stage + (uint32_t)(mt * 128) * row_units;
                        const uint32_t b_lo_c = w_lo + (uint32_t)c * (WTILE / 16);
#pragma unroll
                        for (int tap = 0; tap < TAPS; ++tap) {
                            const int r = tap / KW, s = tap % KW;
                            const uint32_t a_lo = a_lo_mt + (uint32_t)(r * p.P + s) * row_units;
                            const uint32_t b_lo = b_lo_c + (uint32_t)(tap * p.chunks) * (WTILE / 16);
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                umma_bf16(d0, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (c | tap | k) != 0);
                            if (two) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k)
                                    umma_bf16(d1, desc_hi | (uint64_t)(a_lo + 128 * row_units + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc,
                                              (c | tap | k) != 0);
                            }
                        }
                        if (c == p.chunks - 1) {
                            umma_commit(tfull0 + 8 * slot0);
                            if (two) umma_commit(tfull0 + 8 * slot1);
                        }
                    }
                    __syncwarp();
                }
                if (elect_one_sync()) umma_commit(empty0 + 8 * stage);
                __syncwarp();
                if (++stage == (uint32_t)p.nstages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            g_base += n_mt;
        }
    } else if (warp >= 4 && warp < 12 && POOL) {
        // ===== pooled-stem epilogue, stage 1: TMEM -> (+bias, ReLU) -> bf16 conv rows in a smem ring =====
        // Two groups of four warps take alternate M-tiles and never synchronise with each other.  Conv
        // row number `gr` (counted over the whole CTA: 9 per work tile) lives in ring slot gr % kRing as
        // [x][64 ch] (16-byte chunk index XOR-swizzled by x).  Conv row -1 (first band of an image) is
        // stored as zeros: neutral for a max over post-ReLU values (the reference pads with -inf).
        // A row slot may be overwritten once the pool warps have released row gr - kRing (rows_released).
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const float* bias_s = reinterpret_cast<const float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        volatile int* rows_released = reinterpret_cast<volatile int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw)));
        const uint32_t row_bytes = (uint32_t)p.W * 128;
        const int n_mt = ((p.R - 1) * p.P + p.W - 1) / 128 + 1;
        uint32_t g = 0;
        int tile_idx = 0;
        for (int w = walk.begin; tile_idx < walk.count; w += walk.step, ++tile_idx) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * p.rstep + p.yfirst;
            for (int mt = 0; mt < n_mt; ++mt, ++g) {
                if ((int)(g & 1) != grp) continue;
                const uint32_t slot = g & (kFlatSlots - 1), use = g / kFlatSlots;
                const int m = mt * 128 + q * 32 + lane;
                const int i = m / p.P, x = m - i * p.P;
                const bool valid = x < p.W && i < p.R;
                const int gr = tile_idx * p.R + i;  // CTA-wide conv row number
                if (valid) {
                    while (*rows_released < gr - kPoolRing + 1) __nanosleep(64);
                }
                __syncwarp();
                mbar_wait(tfull0 + 8 * slot, use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + slot * 64 + ((uint32_t)(q * 32) << 16);
                const uint32_t srow = stage0 + (uint32_t)(gr % kPoolRing) * row_bytes + (uint32_t)x * 128;
                const bool keep = y0 + i >= 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(taddr + h * 32, v);
                    tmem_ld_wait();
                    if (h == 1) {
                        tc_fence_before();
                        mbar_arrive(tempty0 + 8 * slot);
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + h * 32 + j * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + h * 32 + j * 8 + 4);
                            const float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                                                __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                                                __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                                                __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
                            uint4 o;
                            unsigned* ou = &o.x;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(f[2 * k], 0.f), fmaxf(f[2 * k + 1], 0.f));
                                ou[k] = keep ? *reinterpret_cast<const unsigned*>(&h2) : 0u;
                            }
                            sts128(srow + (((h * 4 + j) ^ (x & 7)) << 4), o);
                        }
                    }
                }
                // this group's 128 threads have written their rows of M-tile g
                mbar_arrive(mdone0 + 8 * slot);
            }
        }
    } else if (warp >= 12 && POOL) {
        // ===== pooled-stem epilogue, stage 2: 3x3 / stride-2 / pad-1 max over the ring -> NHWC [n][56][56][64] =====
        // Eight pool warps walk the M-tiles in order (mdone barriers), emit pooled row j of a tile as soon as
        // the M-tile that completes conv row 2j+2 is in the ring, then release the rows nobody needs any more.
        const int te = threadIdx.x - 12 * 32;  // 0..255
        volatile int* rows_released = reinterpret_cast<volatile int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw)));
        const int Wc = p.W, Wp = p.W >> 1, Hp = p.H >> 1;
        const uint32_t row_bytes = (uint32_t)Wc * 128;
        const int n_mt = ((p.R - 1) * p.P + p.W - 1) / 128 + 1;
        uint32_t g = 0;
        int tile_idx = 0;
        for (int w = walk.begin; tile_idx < walk.count; w += walk.step, ++tile_idx) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * p.rstep + p.yfirst;
            int jnext = 0;
            for (int mt = 0; mt < n_mt; ++mt, ++g) {
                const uint32_t slot = g & (kFlatSlots - 1), use = g / kFlatSlots;
                mbar_wait(mdone0 + 8 * slot, use & 1);
                // pooled rows whose last conv row (2j+2) ends inside this M-tile
                while (2 * jnext + 2 < p.R && ((2 * jnext + 2) * p.P + Wc - 1) / 128 <= mt) {
                    const int j = jnext++;
                    const int prow = (y0 + 1) / 2 + j;
                    const int gr0 = tile_idx * p.R + 2 * j;
                    const uint32_t r0 = stage0 + (uint32_t)(gr0 % kPoolRing) * row_bytes;
                    const uint32_t r1 = stage0 + (uint32_t)((gr0 + 1) % kPoolRing) * row_bytes;
                    const uint32_t r2 = stage0 + (uint32_t)((gr0 + 2) % kPoolRing) * row_bytes;
                    for (int item = te; item < Wp * 8; item += 256) {
                        const int pw = item >> 3, ch = item & 7;
                        __nv_bfloat162 acc[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int xx = 2 * pw + dx;
                            if (xx >= 0) {
                                const uint32_t off = (uint32_t)xx * 128 + ((ch ^ (xx & 7)) << 4);
                                const uint4 a = lds128(r0 + off), bq = lds128(r1 + off), cq = lds128(r2 + off);
                                const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
                                const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&bq);
                                const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&cq);
#pragma unroll
                                for (int k = 0; k < 4; ++k) acc[k] = __hmax2(acc[k], __hmax2(ha[k], __hmax2(hb[k], hc[k])));
                            }
                        }
                        if (prow < Hp)
                            *reinterpret_cast<uint4*>(p.out + (((size_t)img * Hp + prow) * Wp + pw) * 64 + ch * 8) =
                                *reinterpret_cast<const uint4*>(acc);
                    }
                    // rows gr0 and gr0+1 are not needed by later pooled rows (row gr0+2 is: it is row 2(j+1))
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (te == 0) {
                        const bool last = 2 * (j + 1) + 2 >= p.R;  // last pooled row of the tile: its third row is free too
                        *rows_released = gr0 + (last ? 3 : 2);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+bias, +residual, ReLU) -> bf16 -> smem -> coalesced NHWC =====
        // Two groups of four warps take alternate M-tiles (g even / odd).  A warp owns 32 accumulator
        // rows x 64 channels = 4 KB, staged in a private swizzled smem block so that global traffic is
        // whole 128-byte pixel rows (8 lanes x 16 B) instead of one 16-byte piece per lane per line.
        // The residual of the warp's NEXT tile travels in registers until this block is free (see rres below).
        const int q = warp & 3;           // TMEM lane quarter this warp may read
        const int grp = (warp - 4) >> 2;  // which M-tiles (g & 1) this warp handles
        const uint32_t stg = stage0 + (uint32_t)(warp - 4) * 4096;
        const int cbase = nslice * 64;
        const int rr0 = lane >> 3, ch = lane & 7;  // store-out role: row (it*4 + rr0), 16-byte chunk ch
        const bool has_res = p.residual != nullptr;

        // this group's M-tile sequence
        struct Cursor {
            int w, wi, mt, n_mt, img, y0, rows_valid;
            uint32_t g;
        };
        auto tile_setup = [&](Cursor& c) {
            c.img = c.w / p.tiles_per_img;
            c.y0 = (c.w - c.img * p.tiles_per_img) * p.rstep + p.yfirst;
            c.rows_valid = min(p.R, p.H - c.y0);
            c.n_mt = ((c.rows_valid - 1) * p.P + p.W - 1) / 128 + 1;
        };
        auto advance = [&](Cursor& c) {  // next M-tile in schedule order; returns false at the end
            ++c.g;
            if (++c.mt == c.n_mt) {
                c.mt = 0;
                c.w += walk.step;
                if (++c.wi >= walk.count) return false;
                tile_setup(c);
            }
            return true;
        };
        auto pix_of = [&](const Cursor& c) {  // output pixel index of this lane's accumulator row, or -1
            const int m = c.mt * 128 + q * 32 + lane;
            const int i = m / p.P, x = m - i * p.P;
            return (x < p.W && i < c.rows_valid) ? ((c.img * p.H + c.y0 + i) * p.W + x) : -1;
        };
        // The residual of the group's NEXT M-tile is fetched into registers (coalesced: 8 lanes x 16 B per pixel row) BEFORE
        // the current tile is processed, and moved to the staging block once that tile's store-out is done: a whole tile
        // period hides the load.  (A cp.async into the staging block could only start after the store-out and was waited
        // for at once: the epilogue, not the tensor pipe, paced the convs with a residual -- 85 -> 76 us per launch.)
        uint4 rres[8];
        auto load_res = [&](int px) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int pr = __shfl_sync(0xffffffffu, px, it * 4 + rr0);
                rres[it] = ldg_stream128(p.residual + (size_t)max(pr, 0) * p.cout + cbase + ch * 8);  // junk rows read pixel 0, unused
            }
        };
        auto stash_res = [&]() {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rr0;
                sts128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4), rres[it]);
            }
        };

        Cursor cur;
        cur.w = walk.begin, cur.wi = 0, cur.mt = 0, cur.g = 0;
        bool live = walk.count > 0;
        if (live) {
            tile_setup(cur);
            if (grp == 1) live = advance(cur);
        }
        int pix = live ? pix_of(cur) : -1;
        if (live && has_res) load_res(pix);
        while (live) {
            Cursor nxt = cur;
            const bool nlive = advance(nxt) && advance(nxt);
            const int npix = nlive ? pix_of(nxt) : -1;
            if (has_res) {
                stash_res();
                if (nlive) load_res(npix);
            }
            const uint32_t slot = cur.g & (kFlatSlots - 1), use = cur.g / kFlatSlots;
            __syncwarp();
            mbar_wait(tfull0 + 8 * slot, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + slot * 64 + ((uint32_t)(q * 32) << 16);
            const uint32_t srow = stg + lane * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + h * 32, v);
                tmem_ld_wait();
                // issued here, while v[] occupies the registers the TMEM load needs: ptxas then loads straight into rres
                // (issued before the TMEM load it used scratch registers and copied -- waited -- at once)
                if (h == 1) {
                    tc_fence_before();
                    mbar_arrive(tempty0 + 8 * slot);  // accumulator is in registers: hand the slot back
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t sa = srow + (((h * 4 + j) ^ (lane & 7)) << 4);
                    const float* bv = p.bias_v + h * 32 + j * 8;  // compile-time offsets: constant-bank operands
                    const float4 b0 = make_float4(bv[0], bv[1], bv[2], bv[3]);
                    const float4 b1 = make_float4(bv[4], bv[5], bv[6], bv[7]);
                    float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                                  __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                                  __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                                  __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
                    if (has_res && pix >= 0) {
                        const uint4 rv = lds128(sa);
                        const unsigned u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            f[2 * k] += __uint_as_float(u[k] << 16);
                            f[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                    }
                    uint4 o;
                    unsigned* ou = &o.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                        ou[k] = *reinterpret_cast<const unsigned*>(&h2);
                    }
                    sts128(sa, o);
                }
            }
            __syncwarp();
            // store-out: 4 pixel rows (512 contiguous bytes when the pixels are neighbours) per instruction
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rr0;
                const int pr = __shfl_sync(0xffffffffu, pix, rr);
                if (pr >= 0) {
                    const uint4 val = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    *reinterpret_cast<uint4*>(p.out + (size_t)pr * p.cout + cbase + ch * 8) = val;
                }
            }
            __syncwarp();
            cur = nxt;
            live = nlive;
            pix = npix;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// stemw: the pooled stem with TWO output columns per accumulator row (N = 128).
//
// flat_conv_kernel<32, 4, 4, true> spends its time on the 128 B/clk shared-memory pipe: every M128 x N64 x K16 MMA pulls
// 4 KB of A and 2 KB of B through it (48 cycles for 32 cycles of math), 16 of them per 128 conv outputs, and the ring
// writes / pool reads of the fused max-pool travel the same pipe.  Here the staging tensor is viewed as
// PAIRS of s2d pixels -- [n][115][58][32 ch], 64-byte rows, SWIZZLE_64B -- and accumulator row (i, j') holds BOTH conv
// outputs of the pair, (i, 2j') in columns 0..63 and (i, 2j'+1) in columns 64..127.  Output 2j' reads s2d columns
// 2j'..2j'+3, output 2j'+1 reads 2j'+1..2j'+4: the union is five s2d columns t = 0..4 = pairs j', j'+1, j'+2 (first half),
// so one filter row is five K = 16 steps instead of 2 x 4, and the A fetch per conv output is 5/8 of what it was:
//   t = 1, 2, 3   N = 128   B rows = [tap t weights (output 2j') | tap t-1 weights (output 2j'+1)]
//   t = 0         N = 64    tap 0 into columns 0..63          (the very first step is N = 128 with a ZERO second half:
//   t = 4         N = 64    tap 3 into columns 64..127         it initialises both halves of the accumulator)
// With the 16 tap tiles stored in DESCENDING tap order, [tap t | tap t-1] is simply the 128-row window that starts at tap
// t: the resident weights stay 32 KB (+ one zero tile).  Every output still sums its 16 taps in the old order (adding an
// exact zero first changes nothing), so the pooled rows are bit-identical to flat_conv_kernel's (knob test, FX_STEMW=0).
// Row pitch: 57 pairs, not 58 -- the tap that would read pair 57 (s2d columns 114 / 115: conv padding and the dead column,
// all zero) reads pair 0 of the next row instead, first half = s2d column 0 = padded pixels 0 / 1 = zero as well.  A work
// tile (9 conv rows = 4 pooled rows) is then exactly 8 * 57 + 56 = 512 positions = 4 M-tiles, none wasted.
// Epilogue: all eight warps drain one M-tile (lane quarter x 32-channel half); a thread holds both outputs of its pair,
// so it stores max(even, odd) and the odd one -- pooled column w is max(odd[w-1], maxpair[w]) over three rows: two ring
// reads per row instead of three.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue, 12..19 = pool.
// ------------------------------------------------------------------------------------------
constexpr int kStemwThreads = 128 + 16 * 32;
constexpr int kStemwRing = 7;                 // conv rows kept for the pool
constexpr int kStemwSlots = 4;                // 4 x 128 fp32 columns = the whole TMEM
constexpr int kStemwP = 57, kStemwPairs = 56;  // smem row pitch in pairs / valid pairs per conv row
constexpr int kStemwR = 9;                    // conv rows per work tile (8t-1 .. 8t+7)
constexpr int kStemwBoxBytes = (kStemwR + 3) * kStemwP * 64;
constexpr int kStemwStage = (kStemwBoxBytes + 1023) & ~1023;
constexpr int kStemwRowBytes = kStemwPairs * 256;  // ring row: [pair][max(even, odd) | odd][64 ch] bf16
constexpr int kStemwWBytes = 17 * 2048;            // 16 tap tiles (64 cout x 16 ch) + the zero tile
static_assert(kStemwStage - kStemwBoxBytes >= 64, "the last M-tile's deepest tap reads one row past the box: it must be padding");
static_assert((kStemwR - 1) * kStemwP + kStemwPairs == 4 * 128, "a work tile is exactly four M-tiles");

struct StemwParams {
    int n_work, tiles_per_img, Hp, Wp;  // pooled output height / width
    float bias_v[64];
    __nv_bfloat16* out;
    int sched;
};

__global__ void __launch_bounds__(kStemwThreads, 1)
stemw_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const StemwParams p) {
    constexpr int P = kStemwP, R = kStemwR, NMT = 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sW = sbase;
    const uint32_t sA = sW + kStemwWBytes;
    const uint32_t ring0 = sA + 2 * kStemwStage;
    const uint32_t bias0 = ring0 + kStemwRing * kStemwRowBytes;
    const uint32_t bars = bias0 + 256;
    const uint32_t full0 = bars, empty0 = full0 + 16, tfull0 = empty0 + 16, tempty0 = tfull0 + 8 * kStemwSlots;
    const uint32_t mdone0 = tempty0 + 8 * kStemwSlots, wbar = mdone0 + 8 * kStemwSlots, tslot = wbar + 8, pool_sync0 = tslot + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const TileWalk walk(p.n_work, blockIdx.x, gridDim.x, p.sched);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < kStemwSlots; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, 256);
            mbar_init(mdone0 + 8 * i, 256);
        }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tslot, 512);
    if (warp == 3) {
        if (lane == 0) *reinterpret_cast<int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw))) = 0;
        float* bs = reinterpret_cast<float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        bs[lane] = p.bias_v[lane];
        bs[lane + 32] = p.bias_v[lane + 32];
    }
    {  // the zero tile behind tap (0, 0) and the padding behind each A stage: written once, read by the MMAs (async proxy)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < 2048 / 16; i += kStemwThreads) sts128(sW + 16 * 2048 + i * 16, z);
        constexpr int kPad16 = (kStemwStage - kStemwBoxBytes) / 16;
        for (int i = threadIdx.x; i < 2 * kPad16; i += kStemwThreads)
            sts128(sA + (i / kPad16) * kStemwStage + kStemwBoxBytes + (i % kPad16) * 16, z);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;

    // resident weights, tap tiles in descending order (constant data: fetched before waiting for the previous kernel)
    if (warp == 0 && elect_one_sync()) {
        mbar_expect_tx(wbar, 16 * 2048);
        for (int ty = 0; ty < 4; ++ty)
            for (int tx = 0; tx < 4; ++tx) tma_load_2d(sW + ((3 - ty) * 4 + (3 - tx)) * 2048, &map_b, wbar, (ty * 4 + tx) * 16, 0);
    }
    __syncwarp();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer: one 12-row box of pairs per work tile =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int wi = 0, w = walk.begin; wi < walk.count; ++wi, w += walk.step) {
                const int img = w / p.tiles_per_img;
                const int y0 = (w - img * p.tiles_per_img) * 8 - 1;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                mbar_expect_tx(full0 + 8 * stage, kStemwBoxBytes);
                tma_load_4d(sA + stage * kStemwStage, &map_a, full0 + 8 * stage, 0, 0, y0, img);
                if (++stage == 2) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc_w = make_idesc<128>(), idesc_n = make_idesc<64>();
        constexpr uint64_t adesc_hi = make_smem_desc_rowb<64>(0) & 0xFFFFFFFF00000000ull;
        constexpr uint64_t bdesc_hi = make_smem_desc_rowb<32>(0) & 0xFFFFFFFF00000000ull;
        mbar_wait(wbar, 0);
        tc_fence_after();
        uint32_t stage = 0, phase = 0, g_base = 0;
        const uint32_t w_lo = sW >> 4;
        for (int wi = 0; wi < walk.count; ++wi) {
            mbar_wait(full0 + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_lo_stage = (sA + stage * kStemwStage) >> 4;
            // two M-tiles at a time, K steps interleaved over the two accumulators (flat_conv_kernel)
            for (int mt = 0; mt < NMT; mt += 2) {
                const uint32_t g0 = g_base + mt, slot0 = g0 & (kStemwSlots - 1), use0 = g0 / kStemwSlots;
                const uint32_t g1 = g0 + 1, slot1 = g1 & (kStemwSlots - 1), use1 = g1 / kStemwSlots;
                mbar_wait(tempty0 + 8 * slot0, (use0 & 1) ^ 1);
                mbar_wait(tempty0 + 8 * slot1, (use1 & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d0 = tmem_base + slot0 * 128, d1 = tmem_base + slot1 * 128;
                    const uint32_t a_lo_mt = a_lo_stage + (uint32_t)(mt * 128) * 4;  // 64-byte rows = 4 descriptor units
#pragma unroll
                    for (int ty = 0; ty < 4; ++ty) {
#pragma unroll
                        for (int t = 0; t < 5; ++t) {
                            const uint32_t a_lo = a_lo_mt + (uint32_t)(ty * P + (t >> 1)) * 4 + (uint32_t)(t & 1) * 2;
                            const int tile = (3 - ty) * 4 + (3 - (t == 4 ? 3 : t));  // window start: tap t (tap 3 for the odd-only step)
                            const uint32_t b_lo = w_lo + (uint32_t)tile * (2048 / 16);
                            const bool wide = (t >= 1 && t <= 3) || (ty == 0 && t == 0);
                            const uint32_t idesc = wide ? idesc_w : idesc_n;
                            const uint32_t dcol = t == 4 ? 64u : 0u;
                            const uint32_t acc = (ty | t) != 0;
                            umma_bf16(d0 + dcol, adesc_hi | (uint64_t)a_lo, bdesc_hi | (uint64_t)b_lo, idesc, acc);
                            umma_bf16(d1 + dcol, adesc_hi | (uint64_t)(a_lo + 128 * 4), bdesc_hi | (uint64_t)b_lo, idesc, acc);
                        }
                    }
                    umma_commit(tfull0 + 8 * slot0);
                    umma_commit(tfull0 + 8 * slot1);
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(empty0 + 8 * stage);
            __syncwarp();
            if (++stage == 2) {
                stage = 0;
                phase ^= 1;
            }
            g_base += NMT;
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== epilogue: TMEM -> (+bias, ReLU) -> bf16 [max(even, odd) | odd] per pair into the conv-row ring =====
        const int q = warp & 3;
        const int chalf = (warp - 4) >> 2;  // which 32 of the 64 output channels
        const float* bias_s = reinterpret_cast<const float*>(smem_raw + (bias0 - smem_u32(smem_raw))) + chalf * 32;
        volatile int* rows_released = reinterpret_cast<volatile int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw)));
        uint32_t g = 0;
        int tile_idx = 0;
        for (int w = walk.begin; tile_idx < walk.count; w += walk.step, ++tile_idx) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * 8 - 1;
            for (int mt = 0; mt < NMT; ++mt, ++g) {
                const uint32_t slot = g & (kStemwSlots - 1), use = g / kStemwSlots;
                const int m = mt * 128 + q * 32 + lane;
                const int i = m / P, jp = m - i * P;
                const bool valid = jp < kStemwPairs;
                const int gr = tile_idx * R + i;  // CTA-wide conv row number
                if (valid) {
                    while (*rows_released < gr - kStemwRing + 1) __nanosleep(64);
                }
                __syncwarp();
                mbar_wait(tfull0 + 8 * slot, use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + slot * 128 + chalf * 32 + ((uint32_t)(q * 32) << 16);
                const uint32_t srow = ring0 + (uint32_t)(gr % kStemwRing) * kStemwRowBytes + (uint32_t)jp * 256;
                const bool keep = y0 + i >= 0;  // conv row -1 is stored as zeros: neutral for a max over post-ReLU values
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t ve[16], vo[16];
                    tmem_ld16(taddr + h * 16, ve);
                    tmem_ld16(taddr + 64 + h * 16, vo);
                    tmem_ld_wait();
                    if (h == 1) {
                        tc_fence_before();
                        mbar_arrive(tempty0 + 8 * slot);
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + h * 16 + j * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + h * 16 + j * 8 + 4);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            uint4 om, oo;
                            unsigned* um = &om.x;
                            unsigned* uo = &oo.x;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float e0 = fmaxf(__uint_as_float(ve[8 * j + 2 * k]) + bb[2 * k], 0.f);
                                const float e1 = fmaxf(__uint_as_float(ve[8 * j + 2 * k + 1]) + bb[2 * k + 1], 0.f);
                                const float o0 = fmaxf(__uint_as_float(vo[8 * j + 2 * k]) + bb[2 * k], 0.f);
                                const float o1 = fmaxf(__uint_as_float(vo[8 * j + 2 * k + 1]) + bb[2 * k + 1], 0.f);
                                const __nv_bfloat162 hm = __floats2bfloat162_rn(fmaxf(e0, o0), fmaxf(e1, o1));
                                const __nv_bfloat162 ho = __floats2bfloat162_rn(o0, o1);
                                um[k] = keep ? *reinterpret_cast<const unsigned*>(&hm) : 0u;
                                uo[k] = keep ? *reinterpret_cast<const unsigned*>(&ho) : 0u;
                            }
                            const uint32_t chunk = (uint32_t)((chalf * 4 + h * 2 + j) ^ (jp & 7)) << 4;
                            sts128(srow + chunk, om);
                            sts128(srow + 128 + chunk, oo);
                        }
                    }
                }
                mbar_arrive(mdone0 + 8 * slot);
            }
        }
    } else if (warp >= 12) {
        // ===== pool: 3x3 / stride-2 / pad-1 max over the ring -> NHWC [n][56][56][64] =====
        const int te = threadIdx.x - 12 * 32;  // 0..255
        volatile int* rows_released = reinterpret_cast<volatile int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw)));
        uint32_t g = 0;
        int tile_idx = 0;
        for (int w = walk.begin; tile_idx < walk.count; w += walk.step, ++tile_idx) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * 8 - 1;
            int jnext = 0;
            __nv_bfloat162 carry[2][4];
            for (int mt = 0; mt < NMT; ++mt, ++g) {
                const uint32_t slot = g & (kStemwSlots - 1), use = g / kStemwSlots;
                mbar_wait(mdone0 + 8 * slot, use & 1);
                // pooled rows whose last conv row (2j+2) ends inside this M-tile
                while (2 * jnext + 2 < R && ((2 * jnext + 2) * P + kStemwPairs - 1) / 128 <= mt) {
                    const int j = jnext++;
                    const int prow = (y0 + 1) / 2 + j;
                    const int gr0 = tile_idx * R + 2 * j;
                    const uint32_t r0 = ring0 + (uint32_t)(gr0 % kStemwRing) * kStemwRowBytes;
                    const uint32_t r1 = ring0 + (uint32_t)((gr0 + 1) % kStemwRing) * kStemwRowBytes;
                    const uint32_t r2 = ring0 + (uint32_t)((gr0 + 2) % kStemwRing) * kStemwRowBytes;
                    // a thread keeps its (column, channel chunk) items for the whole tile: the horizontal max of conv row 2j+2 is
                    // carried in registers to pooled row j+1, whose first row it is (two ring rows read per pooled row, not three)
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int item = te + it * 256;
                        if (item < p.Wp * 8) {
                            const int pw = item >> 3, ch = item & 7;
                            // conv columns 2pw-1, 2pw, 2pw+1 = odd of pair pw-1, max(even, odd) of pair pw
                            const uint32_t offm = (uint32_t)pw * 256 + ((uint32_t)(ch ^ (pw & 7)) << 4);
                            const uint32_t offo = (uint32_t)max(pw - 1, 0) * 256 + 128 + ((uint32_t)(ch ^ (max(pw - 1, 0) & 7)) << 4);
                            auto hrow = [&](uint32_t rb, __nv_bfloat162 (&h)[4]) {
                                const uint4 a = lds128(rb + offm);
                                const uint4 o = lds128(rb + (pw > 0 ? offo : offm));  // column -1 is padding: the pair's own max again
                                const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
                                const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
                                for (int k = 0; k < 4; ++k) h[k] = __hmax2(ha[k], ho[k]);
                            };
                            __nv_bfloat162 h0[4], h1[4], h2[4];
                            if (j == 0) {
                                hrow(r0, h0);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) h0[k] = carry[it][k];
                            }
                            hrow(r1, h1);
                            hrow(r2, h2);
                            __nv_bfloat162 acc[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                acc[k] = __hmax2(h0[k], __hmax2(h1[k], h2[k]));
                                carry[it][k] = h2[k];
                            }
                            if (prow < p.Hp)
                                *reinterpret_cast<uint4*>(p.out + (((size_t)img * p.Hp + prow) * p.Wp + pw) * 64 + ch * 8) =
                                    *reinterpret_cast<const uint4*>(acc);
                        }
                    }
                    // rows gr0 and gr0+1 are not needed by later pooled rows (row gr0+2 is: it is row 2(j+1))
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (te == 0) {
                        const bool last = 2 * (j + 1) + 2 >= R;  // last pooled row of the tile: its third row is free too
                        *rows_released = gr0 + (last ? 3 : 2);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// stemw2: stemw_conv_kernel as a CTA PAIR (cta_group::2): one M = 256 MMA per step covers the same M-tile of both CTAs' work
// tiles, and each CTA supplies HALF of the step's B operand -- an N = 128 step costs 4 KB of A + 2 KB of B per CTA instead of
// 4 + 4 (48 cycles of the shared-memory pipe for 64 cycles of math), an N = 64 step 4 + 1.  The leader's half is always the
// even output's weights, the peer's the odd output's, so no tile is stored twice: 33 KB per CTA.  Protocol as flat2_conv_kernel.
constexpr int kStemw2WBytes = 33 * 1024;
constexpr int kStemw2Stages = 3;                        // A stages (the half-size ring rows below pay for the third)
constexpr int kStemw2Ring = 8;                          // conv rows kept for the pool
constexpr int kStemw2RowBytes = kStemwPairs * 128;      // ring row: [pair][64 ch] bf16, one (pre-maxed) value per pair
__host__ __device__ constexpr uint32_t stemw2_w_off(int ty, int t) {  // byte offset of step (ty, t)'s half operand in issue order
    return (uint32_t)((ty == 0 ? 2 * t : 9 + (ty - 1) * 8 + (t == 0 ? 0 : 2 * t - 1)) * 1024);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kStemwThreads, 1)
stemw2_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_bn,
                   const StemwParams p) {
    constexpr int P = kStemwP, R = kStemwR, NMT = 4;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sW = sbase;
    const uint32_t sA = sW + kStemw2WBytes;
    const uint32_t ring0 = sA + kStemw2Stages * kStemwStage;                      // conv-row ring: [row][pair][64 ch] of hmax values (see the epilogue)
    const uint32_t side0 = ring0 + kStemw2Ring * kStemw2RowBytes;     // [row][2][64 ch]: odd outputs handed across warp boundaries
    const uint32_t bias0 = side0 + kStemw2Ring * 2 * 128;
    const uint32_t bars = bias0 + 256;
    const uint32_t full0 = bars, empty0 = full0 + 8 * kStemw2Stages, tfull0 = empty0 + 8 * kStemw2Stages, tempty0 = tfull0 + 8 * kStemwSlots;
    const uint32_t mdone0 = tempty0 + 8 * kStemwSlots, wbar = mdone0 + 8 * kStemwSlots, tslot = wbar + 8, pool_sync0 = tslot + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const TileWalk walk(p.n_work >> 1, blockIdx.x >> 1, gridDim.x >> 1, p.sched);  // unit u = work tiles 2u (leader) and 2u + 1 (peer); n_work is even

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_bn);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStemw2Stages; ++i) {
            mbar_init(full0 + 8 * i, 2);   // leader's arrive.expect_tx + the peer producer's arrive
            mbar_init(empty0 + 8 * i, 1);  // multicast commit
        }
        for (int i = 0; i < kStemwSlots; ++i) {
            mbar_init(tfull0 + 8 * i, 1);         // multicast commit
            mbar_init(tempty0 + 8 * i, 2 * 256);  // the eight epilogue warps of both CTAs
            mbar_init(mdone0 + 8 * i, 256);
        }
        mbar_init(wbar, 2);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tslot, 512);
    if (warp == 3) {
        if (lane == 0) *reinterpret_cast<int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw))) = 0;
        float* bs = reinterpret_cast<float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        bs[lane] = p.bias_v[lane];
        bs[lane + 32] = p.bias_v[lane + 32];
    }
    {  // the odd outputs' (= the peer's) half of the very first step and the padding behind each A stage: written once, read
       // by the MMAs (async proxy)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        if (!leader)
            for (int i = threadIdx.x; i < 2048 / 16; i += kStemwThreads) sts128(sW + i * 16, z);
        constexpr int kPad16 = (kStemwStage - kStemwBoxBytes) / 16;
        for (int i = threadIdx.x; i < kStemw2Stages * kPad16; i += kStemwThreads)
            sts128(sA + (i / kPad16) * kStemwStage + kStemwBoxBytes + (i % kPad16) * 16, z);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;

    // resident weights: this CTA's half of every step's B operand, in issue order (constant data: fetched before waiting for
    // the previous kernel).  N = 128 steps: all 64 output channels of ONE of the two outputs (leader: even output = tap t,
    // peer: odd output = tap t - 1); N = 64 steps: 32 of the 64 channels of the single output the step feeds.
    if (warp == 0 && elect_one_sync()) {
        if (leader) mbar_expect_tx(wbar, 2 * kStemw2WBytes - 2048);
        for (int ty = 0; ty < 4; ++ty)
            for (int t = 0; t < 5; ++t) {
                const bool wide = (t >= 1 && t <= 3) || (ty == 0 && t == 0);
                const uint32_t dst = sW + stemw2_w_off(ty, t);
                if (wide) {
                    const int tx = leader ? t : t - 1;
                    if (tx >= 0) tma_load_2d_2cta(dst, &map_b, wbar, (ty * 4 + tx) * 16, 0);
                } else {
                    tma_load_2d_2cta(dst, &map_bn, wbar, (ty * 4 + (t == 0 ? 0 : 3)) * 16, (int)rank * 32);
                }
            }
        if (!leader) mbar_arrive_leader(wbar);
    }
    __syncwarp();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer: one 12-row box of pairs per work tile =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int wi = 0, u = walk.begin; wi < walk.count; ++wi, u += walk.step) {
                const int w = 2 * u + (int)rank;
                const int img = w / p.tiles_per_img;
                const int y0 = (w - img * p.tiles_per_img) * 8 - 1;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (leader) mbar_expect_tx(full0 + 8 * stage, 2 * kStemwBoxBytes);
                tma_load_4d_2cta(sA + stage * kStemwStage, &map_a, full0 + 8 * stage, 0, 0, y0, img);
                if (!leader) mbar_arrive_leader(full0 + 8 * stage);
                if (++stage == kStemw2Stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ===== MMA issuer (leader CTA only): M = 256 = the same M-tile of both CTAs' work tiles =====
        constexpr uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        constexpr uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        constexpr uint64_t adesc_hi = make_smem_desc_rowb<64>(0) & 0xFFFFFFFF00000000ull;
        constexpr uint64_t bdesc_hi = make_smem_desc_rowb<32>(0) & 0xFFFFFFFF00000000ull;
        mbar_wait(wbar, 0);
        tc_fence_after();
        uint32_t stage = 0, phase = 0, g_base = 0;
        const uint32_t w_lo = sW >> 4;
        for (int wi = 0; wi < walk.count; ++wi) {
            mbar_wait(full0 + 8 * stage, phase);
            tc_fence_after();
            const uint32_t a_lo_stage = (sA + stage * kStemwStage) >> 4;
            // two M-tiles at a time, K steps interleaved over the two accumulators (flat_conv_kernel)
            for (int mt = 0; mt < NMT; mt += 2) {
                const uint32_t g0 = g_base + mt, slot0 = g0 & (kStemwSlots - 1), use0 = g0 / kStemwSlots;
                const uint32_t g1 = g0 + 1, slot1 = g1 & (kStemwSlots - 1), use1 = g1 / kStemwSlots;
                mbar_wait(tempty0 + 8 * slot0, (use0 & 1) ^ 1);
                mbar_wait(tempty0 + 8 * slot1, (use1 & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d0 = tmem_base + slot0 * 128, d1 = tmem_base + slot1 * 128;
                    const uint32_t a_lo_mt = a_lo_stage + (uint32_t)(mt * 128) * 4;  // 64-byte rows = 4 descriptor units
#pragma unroll
                    for (int ty = 0; ty < 4; ++ty) {
#pragma unroll
                        for (int t = 0; t < 5; ++t) {
                            const uint32_t a_lo = a_lo_mt + (uint32_t)(ty * P + (t >> 1)) * 4 + (uint32_t)(t & 1) * 2;
                            const uint32_t b_lo = w_lo + (uint32_t)stemw2_w_off(ty, t) / 16;
                            const bool wide = (t >= 1 && t <= 3) || (ty == 0 && t == 0);
                            const uint32_t idesc = wide ? idesc_w : idesc_n;
                            const uint32_t dcol = t == 4 ? 64u : 0u;
                            const uint32_t acc = (ty | t) != 0;
                            umma_bf16_2cta(d0 + dcol, adesc_hi | (uint64_t)a_lo, bdesc_hi | (uint64_t)b_lo, idesc, acc);
                            umma_bf16_2cta(d1 + dcol, adesc_hi | (uint64_t)(a_lo + 128 * 4), bdesc_hi | (uint64_t)b_lo, idesc, acc);
                        }
                    }
                    umma_commit_2cta(tfull0 + 8 * slot0);
                    umma_commit_2cta(tfull0 + 8 * slot1);
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit_2cta(empty0 + 8 * stage);
            __syncwarp();
            if (++stage == kStemw2Stages) {
                stage = 0;
                phase ^= 1;
            }
            g_base += NMT;
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== epilogue: TMEM -> (+bias, ReLU) -> bf16 [max(even, odd) | odd] per pair into the conv-row ring =====
        const int q = warp & 3;
        const int chalf = (warp - 4) >> 2;  // which 32 of the 64 output channels
        const float* bias_s = reinterpret_cast<const float*>(smem_raw + (bias0 - smem_u32(smem_raw))) + chalf * 32;
        volatile int* rows_released = reinterpret_cast<volatile int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw)));
        uint32_t g = 0;
        int tile_idx = 0;
        for (int u = walk.begin; tile_idx < walk.count; u += walk.step, ++tile_idx) {
            const int w = 2 * u + (int)rank;
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * 8 - 1;
            for (int mt = 0; mt < NMT; ++mt, ++g) {
                const uint32_t slot = g & (kStemwSlots - 1), use = g / kStemwSlots;
                const int m = mt * 128 + q * 32 + lane;
                const int i = m / P, jp = m - i * P;
                const bool valid = jp < kStemwPairs;
                const int gr = tile_idx * R + i;  // CTA-wide conv row number
                if (valid) {
                    while (*rows_released < gr - kStemw2Ring + 1) __nanosleep(64);
                }
                __syncwarp();
                mbar_wait(tfull0 + 8 * slot, use & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + slot * 128 + chalf * 32 + ((uint32_t)(q * 32) << 16);
                const uint32_t rslot = (uint32_t)(gr % kStemw2Ring);
                const uint32_t srow = ring0 + rslot * kStemw2RowBytes + (uint32_t)jp * 128;
                const bool keep = y0 + i >= 0;  // conv row -1 is stored as zeros: neutral for a max over post-ReLU values
                // The ring holds ONE value per pair: hmax[j'] = max(odd[j'-1], even[j'], odd[j']) = the three conv columns 2j'-1 .. 2j'+1
                // of pooled column j'.  odd[j'-1] comes from the neighbouring lane by shuffle; where that neighbour sits in another
                // warp (lane 0), the neighbour (its lane 31) leaves its odd value in a small side array and the pool adds it.
                const bool from_left = lane > 0 && jp > 0;
                const bool hand_over = lane == 31 && valid && jp + 1 < kStemwPairs;
                const uint32_t sside = side0 + (rslot * 2 + (uint32_t)(((m + 1) >> 5) - ((i * P) >> 5) - 1)) * 128;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t ve[16], vo[16];
                    tmem_ld16(taddr + h * 16, ve);
                    tmem_ld16(taddr + 64 + h * 16, vo);
                    tmem_ld_wait();
                    if (h == 1) {
                        tc_fence_before();
                        mbar_arrive_leader(tempty0 + 8 * slot);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + h * 16 + j * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + h * 16 + j * 8 + 4);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        uint4 om, oo;
                        unsigned* um = &om.x;
                        unsigned* uo = &oo.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float e0 = fmaxf(__uint_as_float(ve[8 * j + 2 * k]) + bb[2 * k], 0.f);
                            const float e1 = fmaxf(__uint_as_float(ve[8 * j + 2 * k + 1]) + bb[2 * k + 1], 0.f);
                            const float o0 = fmaxf(__uint_as_float(vo[8 * j + 2 * k]) + bb[2 * k], 0.f);
                            const float o1 = fmaxf(__uint_as_float(vo[8 * j + 2 * k + 1]) + bb[2 * k + 1], 0.f);
                            const __nv_bfloat162 ho = __floats2bfloat162_rn(o0, o1);
                            const unsigned odd = (keep && valid) ? *reinterpret_cast<const unsigned*>(&ho) : 0u;
                            const unsigned left = __shfl_up_sync(0xffffffffu, odd, 1);
                            __nv_bfloat162 hm = __floats2bfloat162_rn(fmaxf(e0, o0), fmaxf(e1, o1));
                            if (from_left) hm = __hmax2(hm, *reinterpret_cast<const __nv_bfloat162*>(&left));
                            um[k] = keep ? *reinterpret_cast<const unsigned*>(&hm) : 0u;
                            uo[k] = odd;
                        }
                        const uint32_t cidx = (uint32_t)(chalf * 4 + h * 2 + j);
                        if (valid) sts128(srow + ((cidx ^ (uint32_t)(jp & 7)) << 4), om);
                        if (hand_over) sts128(sside + (cidx << 4), oo);
                    }
                }
                mbar_arrive(mdone0 + 8 * slot);
            }
        }
    } else if (warp >= 12) {
        // ===== pool: 3x3 / stride-2 / pad-1 max over the ring -> NHWC [n][56][56][64] =====
        const int te = threadIdx.x - 12 * 32;  // 0..255
        volatile int* rows_released = reinterpret_cast<volatile int*>(smem_raw + (pool_sync0 - smem_u32(smem_raw)));
        uint32_t g = 0;
        int tile_idx = 0;
        for (int u = walk.begin; tile_idx < walk.count; u += walk.step, ++tile_idx) {
            const int w = 2 * u + (int)rank;
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * 8 - 1;
            int jnext = 0;
            __nv_bfloat162 carry[2][4];
            for (int mt = 0; mt < NMT; ++mt, ++g) {
                const uint32_t slot = g & (kStemwSlots - 1), use = g / kStemwSlots;
                mbar_wait(mdone0 + 8 * slot, use & 1);
                // pooled rows whose last conv row (2j+2) ends inside this M-tile
                while (2 * jnext + 2 < R && ((2 * jnext + 2) * P + kStemwPairs - 1) / 128 <= mt) {
                    const int j = jnext++;
                    const int prow = (y0 + 1) / 2 + j;
                    const int gr0 = tile_idx * R + 2 * j;
                    const uint32_t s0 = (uint32_t)(gr0 % kStemw2Ring), s1 = (uint32_t)((gr0 + 1) % kStemw2Ring), s2 = (uint32_t)((gr0 + 2) % kStemw2Ring);
                    // a thread keeps its (column, channel chunk) items for the whole tile: the value of conv row 2j+2 is carried in
                    // registers to pooled row j+1, whose first row it is (two ring rows read per pooled row, not three)
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int item = te + it * 256;
                        if (item < p.Wp * 8) {
                            const int pw = item >> 3, ch = item & 7;
                            const uint32_t offm = (uint32_t)pw * 128 + ((uint32_t)(ch ^ (pw & 7)) << 4);
                            auto hrow = [&](uint32_t rs, int irow, __nv_bfloat162 (&h)[4]) {
                                const uint4 a = lds128(ring0 + rs * kStemw2RowBytes + offm);
                                const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
                                for (int k = 0; k < 4; ++k) h[k] = ha[k];
                                const int pos = irow * P + pw;  // position of the pair inside the work tile: lane 0 of an epilogue warp?
                                if (pw > 0 && (pos & 31) == 0) {
                                    const uint4 o = lds128(side0 + (rs * 2 + (uint32_t)((pos >> 5) - ((irow * P) >> 5) - 1)) * 128 + ((uint32_t)ch << 4));
                                    const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) h[k] = __hmax2(h[k], ho[k]);
                                }
                            };
                            __nv_bfloat162 h0[4], h1[4], h2[4];
                            if (j == 0) {
                                hrow(s0, 2 * j, h0);
                            } else {
#pragma unroll
                                for (int k = 0; k < 4; ++k) h0[k] = carry[it][k];
                            }
                            hrow(s1, 2 * j + 1, h1);
                            hrow(s2, 2 * j + 2, h2);
                            __nv_bfloat162 acc[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                acc[k] = __hmax2(h0[k], __hmax2(h1[k], h2[k]));
                                carry[it][k] = h2[k];
                            }
                            if (prow < p.Hp)
                                *reinterpret_cast<uint4*>(p.out + (((size_t)img * p.Hp + prow) * p.Wp + pw) * 64 + ch * 8) =
                                    *reinterpret_cast<const uint4*>(acc);
                        }
                    }
                    // rows gr0 and gr0+1 are not needed by later pooled rows (row gr0+2 is: it is row 2(j+1))
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    if (te == 0) {
                        const bool last = 2 * (j + 1) + 2 >= R;  // last pooled row of the tile: its third row is free too
                        *rows_released = gr0 + (last ? 3 : 2);
                    }
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory / barriers stay alive until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// flat2: the layer-1 kernel (3x3 / stride 1, Cin = Cout = 64) as a CTA PAIR (cta_group::2).
//
// Why: flat_conv_kernel's N = 64 MMAs are bound by the 128 B/clk shared-memory pipe -- every M128 x N64 x K16 MMA
// fetches 4 KB of A and 2 KB of B.  With one M = 256 MMA per CTA pair each CTA supplies its own 128 A
// rows (its own work tile) but only HALF of B (32 of the 64 output channels): 5 KB per MMA instead of 6, and only
// half of the folded weights (36 KB instead of 72 KB) resident per CTA, which buys a third A stage.
// Protocol (as tc2_conv_kernel): the even CTA of the pair issues every MMA; its `full` / weight barriers count the TMA
// bytes of both CTAs; `tcgen05.commit` multicasts `empty` / accumulator-ready to both; the epilogue warps of both CTAs
// arrive on the leader's accumulator-free barriers.  Unit u of a pair = work tiles 2u (leader) and 2u + 1 (peer).
// Warp roles per CTA: 0 = TMA producer, 1 = MMA issuer (leader only), 2 = TMEM allocator, 4..11 = epilogue (bias: kernel parameters).
// ------------------------------------------------------------------------------------------
constexpr int kFlat2Threads = 384;

struct Flat2Params {
    int P, W, H, R, tiles_per_img, n_work, batch, cout;
    int nstages, stage_bytes, box_bytes, slack_bytes;
    float bias_v[64];  // folded bias as constant operands (FlatParams)
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    int relu;
    int sched;  // tile walk over the pair's units (TileWalk)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFlat2Threads, 1)
flat2_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Flat2Params p) {
    constexpr int TAPS = 9, KW = 3, ROWB = 128, KSTEPS = 4;
    constexpr int WTILE = 32 * ROWB;  // one tap's weights of THIS CTA's 32 output channels
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sW = sbase;
    const uint32_t sA = sbase + TAPS * WTILE;
    const uint32_t stage0 = sA + p.nstages * p.stage_bytes + p.slack_bytes;  // epilogue staging: 8 warps x 4 KB
    const uint32_t bias0 = stage0 + 8 * 4096;
    const uint32_t bars = bias0 + 256;
    const uint32_t full0 = bars, empty0 = full0 + 8 * p.nstages, tfull0 = empty0 + 8 * p.nstages;
    const uint32_t tempty0 = tfull0 + 8 * kFlatSlots, wbar = tempty0 + 8 * kFlatSlots, tslot = wbar + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_units = (p.n_work + 1) >> 1;
    const TileWalk walk(n_units, pair, n_pairs, p.sched);
    const int n_mt = ((p.R - 1) * p.P + p.W - 1) / 128 + 1;  // every tile is issued at full height (rows past H are zero fill)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nstages; ++i) {
            mbar_init(full0 + 8 * i, 2);   // leader's arrive.expect_tx + the peer producer's arrive
            mbar_init(empty0 + 8 * i, 1);  // multicast commit
        }
        for (int i = 0; i < kFlatSlots; ++i) {
            mbar_init(tfull0 + 8 * i, 1);     // multicast commit
            mbar_init(tempty0 + 8 * i, 256);  // one epilogue group (128 threads) of each CTA
        }
        mbar_init(wbar, 2);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tslot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;

    // this CTA's half of the folded weights is constant data: fetch it before waiting for the previous kernel (PDL)
    if (warp == 0 && elect_one_sync()) {
        if (leader) mbar_expect_tx(wbar, 2 * TAPS * WTILE);
        for (int tap = 0; tap < TAPS; ++tap) tma_load_2d_2cta(sW + tap * WTILE, &map_b, wbar, tap * 64, (int)rank * 32);
        if (!leader) mbar_arrive_leader(wbar);
    }
    __syncwarp();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): the halo box of this CTA's tile of the unit =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int ui = 0, u = walk.begin; ui < walk.count; ++ui, u += walk.step) {
                const int w = 2 * u + (int)rank;
                const int img = w < p.n_work ? w / p.tiles_per_img : p.batch;  // the odd leftover: out of bounds -> zeros
                const int y0 = w < p.n_work ? (w - img * p.tiles_per_img) * p.R : 0;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (leader) mbar_expect_tx(full0 + 8 * stage, 2 * p.box_bytes);
                tma_load_4d_2cta(sA + stage * p.stage_bytes, &map_a, full0 + 8 * stage, 0, -1, y0 - 1, img);
                if (!leader) mbar_arrive_leader(full0 + 8 * stage);
                if (++stage == (uint32_t)p.nstages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): M = 256 (both CTAs' tiles), N = 64 =====
        if (leader) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            constexpr uint64_t desc_hi = make_smem_desc_rowb<ROWB>(0) & 0xFFFFFFFF00000000ull;
            mbar_wait(wbar, 0);
            tc_fence_after();
            uint32_t stage = 0, phase = 0, g_base = 0;
            const uint32_t w_lo = sW >> 4;
            const uint32_t row_units = ROWB / 16;
            for (int ui = 0, u = walk.begin; ui < walk.count; ++ui, u += walk.step) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                const uint32_t a_lo_stage = (sA + stage * p.stage_bytes) >> 4;
                for (int mt = 0; mt < n_mt; mt += 2) {
                    const bool two = mt + 1 < n_mt;
                    const uint32_t g0 = g_base + mt, slot0 = g0 & (kFlatSlots - 1), use0 = g0 / kFlatSlots;
                    const uint32_t g1 = g0 + 1, slot1 = g1 & (kFlatSlots - 1), use1 = g1 / kFlatSlots;
                    mbar_wait(tempty0 + 8 * slot0, (use0 & 1) ^ 1);
                    if (two) mbar_wait(tempty0 + 8 * slot1, (use1 & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t d0 = tmem_base + slot0 * 64, d1 = tmem_base + slot1 * 64;
                        const uint32_t a_lo_mt = a_lo_stage + (uint32_t)(mt * 128) * row_units;
#pragma unroll
                        for (int tap = 0; tap < TAPS; ++tap) {
                            const uint32_t a_lo = a_lo_mt + (uint32_t)((tap / KW) * p.P + (tap % KW)) * row_units;
                            const uint32_t b_lo = w_lo + (uint32_t)tap * (WTILE / 16);
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                umma_bf16_2cta(d0, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (tap | k) != 0);
                            if (two) {
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k)
                                    umma_bf16_2cta(d1, desc_hi | (uint64_t)(a_lo + 128 * row_units + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k),
                                                   idesc, (tap | k) != 0);
                            }
                        }
                        umma_commit_2cta(tfull0 + 8 * slot0);
                        if (two) umma_commit_2cta(tfull0 + 8 * slot1);
                    }
                    __syncwarp();
                }
                if (elect_one_sync()) umma_commit_2cta(empty0 + 8 * stage);
                __syncwarp();
                if (++stage == (uint32_t)p.nstages) {
                    stage = 0;
                    phase ^= 1;
                }
                g_base += n_mt;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): as flat_conv_kernel, on this CTA's tile of every unit =====
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const uint32_t stg = stage0 + (uint32_t)(warp - 4) * 4096;
        const int rr0 = lane >> 3, ch = lane & 7;
        const bool has_res = p.residual != nullptr;

        struct Cursor {
            int u, ui, mt, img, y0, rows_valid;
            uint32_t g;
        };
        auto tile_setup = [&](Cursor& c) {
            const int w = 2 * c.u + (int)rank;
            if (w < p.n_work) {
                c.img = w / p.tiles_per_img;
                c.y0 = (w - c.img * p.tiles_per_img) * p.R;
                c.rows_valid = min(p.R, p.H - c.y0);
            } else {
                c.rows_valid = 0;  // the odd leftover: nothing to store
            }
        };
        auto advance = [&](Cursor& c) {
            ++c.g;
            if (++c.mt == n_mt) {
                c.mt = 0;
                c.u += walk.step;
                if (++c.ui >= walk.count) return false;
                tile_setup(c);
            }
            return true;
        };
        auto pix_of = [&](const Cursor& c) {
            const int m = c.mt * 128 + q * 32 + lane;
            const int i = m / p.P, x = m - i * p.P;
            return (x < p.W && i < c.rows_valid) ? ((c.img * p.H + c.y0 + i) * p.W + x) : -1;
        };
        // residual of the group's next M-tile: registers first, staging block after this tile's store-out (flat_conv_kernel)
        uint4 rres[8];
        auto load_res = [&](int px) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int pr = __shfl_sync(0xffffffffu, px, it * 4 + rr0);
                rres[it] = ldg_stream128(p.residual + (size_t)max(pr, 0) * p.cout + ch * 8);  // junk rows read pixel 0, unused
            }
        };
        auto stash_res = [&]() {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rr0;
                sts128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4), rres[it]);
            }
        };

        Cursor cur;
        cur.u = walk.begin, cur.ui = 0, cur.mt = 0, cur.img = 0, cur.y0 = 0, cur.rows_valid = 0, cur.g = 0;
        bool live = walk.count > 0;
        if (live) {
            tile_setup(cur);
            if (grp == 1) live = advance(cur);
        }
        int pix = live ? pix_of(cur) : -1;
        if (live && has_res) load_res(pix);
        while (live) {
            Cursor nxt = cur;
            const bool nlive = advance(nxt) && advance(nxt);
            const int npix = nlive ? pix_of(nxt) : -1;
            if (has_res) {
                stash_res();
                if (nlive) load_res(npix);
            }
            const uint32_t slot = cur.g & (kFlatSlots - 1), use = cur.g / kFlatSlots;
            __syncwarp();
            mbar_wait(tfull0 + 8 * slot, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + slot * 64 + ((uint32_t)(q * 32) << 16);
            const uint32_t srow = stg + lane * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + h * 32, v);
                tmem_ld_wait();
                if (h == 1) {
                    tc_fence_before();
                    mbar_arrive_leader(tempty0 + 8 * slot);  // accumulator is in registers: hand the slot back to the leader
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t sa = srow + (((h * 4 + j) ^ (lane & 7)) << 4);
                    const float* bv = p.bias_v + h * 32 + j * 8;  // compile-time offsets: constant-bank operands
                    const float4 b0 = make_float4(bv[0], bv[1], bv[2], bv[3]);
                    const float4 b1 = make_float4(bv[4], bv[5], bv[6], bv[7]);
                    float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                                  __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                                  __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                                  __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
                    if (has_res && pix >= 0) {
                        const uint4 rv = lds128(sa);
                        const unsigned uu[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            f[2 * k] += __uint_as_float(uu[k] << 16);
                            f[2 * k + 1] += __uint_as_float(uu[k] & 0xffff0000u);
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                    }
                    uint4 o;
                    unsigned* ou = &o.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                        ou[k] = *reinterpret_cast<const unsigned*>(&h2);
                    }
                    sts128(sa, o);
                }
            }
            __syncwarp();
            {  // shuffles, loads, stores in three batches: every dependent trip through the (busy) shared-memory pipe is slow
                int prs[8];
                uint4 vals[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) prs[it] = __shfl_sync(0xffffffffu, pix, it * 4 + rr0);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rr0;
                    vals[it] = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
                }
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    if (prs[it] >= 0) *reinterpret_cast<uint4*>(p.out + (size_t)prs[it] * p.cout + ch * 8) = vals[it];
            }
            __syncwarp();
            cur = nxt;
            live = nlive;
            pix = npix;
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory / barriers stay alive until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// flat2w: layer 1 with TWO output columns per accumulator row (the stemw idea on flat2's CTA pair).
//
// flat2_conv_kernel is bound by the shared-memory pipe: per M256 x N64 x K16 MMA each CTA fetches 4 KB of A and 1 KB of B
// (40 cycles for 32 cycles of math), nine taps x four K-steps per M-tile.  Here accumulator row (i, j') holds the outputs
// (i, 2j') in columns 0..63 and (i, 2j'+1) in columns 64..127.  Together they read input columns 2j'-1 .. 2j'+2: four
// column steps t per filter row instead of 2 x 3, so the A bytes per output fall by a third:
//   t = 1, 2   N = 128   even output: tap dx = t,  odd output: tap dx = t-1   (each CTA supplies one of the two halves of B)
//   t = 0      N = 64    even output only, tap dx = 0 -> columns 0..63        (filter row 0: N = 128 with a ZERO odd half, so
//   t = 3      N = 64    odd output only,  tap dx = 2 -> columns 64..127       that the first MMA initialises all 128 columns)
// The input is staged as two column-parity planes (E: columns 0, 2, ..; O: columns -1, 1, ..; one 5-D map over
// [64][W/2][parity][H][n], unit-stride boxes), each a flat list of 128-byte rows with pitch P = W/2 + 1; input column
// 2j' + t - 1 of filter row dy is plane (t odd ? E : O) viewed dy * P + (t >> 1) rows further down.  A work tile is R = 4
// output rows = 3 * 29 + 28 = 115 positions = ONE M-tile per CTA; all eight epilogue warps drain it (lane quarter x even /
// odd output).  Every output sums its nine taps in flat2's order: rows are bit-identical (knob test, FX_FLAT2W=0).
// ------------------------------------------------------------------------------------------
constexpr int kF2wSlots = 4;          // 4 x 128 fp32 columns
constexpr int kF2wUnit = 32 * 128;    // weight storage unit: 32 output channels x 64 input channels (one box)
constexpr int kF2wUnits = 19;         // per CTA: filter row 0: 2 + 2 + 2 + 1, rows 1 and 2: 1 + 2 + 2 + 1

struct Flat2wParams {
    int P, Wp, W, H, R, tiles_per_img, n_work, batch, cout;
    int nstages, stage_bytes, plane_bytes, box_bytes;
    float bias_v[64];
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    int relu;
    int sched;
};

__device__ __forceinline__ constexpr int f2w_unit_of(int dy, int t) {
    return dy == 0 ? 2 * t : 7 + (dy - 1) * 6 + (t == 0 ? 0 : 2 * t - 1);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFlat2Threads, 1)
flat2w_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Flat2wParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sW = sbase;
    const uint32_t sA = sW + kF2wUnits * kF2wUnit;
    const uint32_t stage0 = sA + p.nstages * p.stage_bytes;  // epilogue staging: 8 warps x 4 KB (the junk rows of the last A stage over-read into it)
    const uint32_t bars = stage0 + 8 * 4096;
    const uint32_t full0 = bars, empty0 = full0 + 8 * p.nstages, tfull0 = empty0 + 8 * p.nstages;
    const uint32_t tempty0 = tfull0 + 8 * kF2wSlots, wbar = tempty0 + 8 * kF2wSlots, tslot = wbar + 8;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_units = (p.n_work + 1) >> 1;
    const TileWalk walk(n_units, pair, n_pairs, p.sched);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.nstages; ++i) {
            mbar_init(full0 + 8 * i, 2);   // leader's arrive.expect_tx + the peer producer's arrive
            mbar_init(empty0 + 8 * i, 1);  // multicast commit
        }
        for (int i = 0; i < kF2wSlots; ++i) {
            mbar_init(tfull0 + 8 * i, 1);        // multicast commit
            mbar_init(tempty0 + 8 * i, 2 * 256);  // the eight epilogue warps of both CTAs
        }
        mbar_init(wbar, 2);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tslot, 512);
    if (!leader) {  // the odd outputs' half of the very first column step is zero (it only initialises the accumulator)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < 2 * kF2wUnit / 16; i += kFlat2Threads) sts128(sW + i * 16, z);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;

    // this CTA's halves of the folded weights are constant data: fetched before waiting for the previous kernel (PDL)
    if (warp == 0 && elect_one_sync()) {
        if (leader) mbar_expect_tx(wbar, (2 * kF2wUnits - 2) * kF2wUnit);
        for (int dy = 0; dy < 3; ++dy)
            for (int t = 0; t < 4; ++t) {
                const bool wide = t == 1 || t == 2 || (dy == 0 && t == 0);
                const uint32_t dst = sW + f2w_unit_of(dy, t) * kF2wUnit;
                if (wide) {
                    const int dx = leader ? t : t - 1;  // even outputs: tap t, odd outputs: tap t - 1
                    if (dx >= 0) {
                        tma_load_2d_2cta(dst, &map_b, wbar, (dy * 3 + dx) * 64, 0);
                        tma_load_2d_2cta(dst + kF2wUnit, &map_b, wbar, (dy * 3 + dx) * 64, 32);
                    }
                } else {
                    tma_load_2d_2cta(dst, &map_b, wbar, (dy * 3 + (t == 0 ? 0 : 2)) * 64, (int)rank * 32);
                }
            }
        if (!leader) mbar_arrive_leader(wbar);
    }
    __syncwarp();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): the two parity planes of this CTA's tile =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int ui = 0, u = walk.begin; ui < walk.count; ++ui, u += walk.step) {
                const int w = 2 * u + (int)rank;
                const int img = w < p.n_work ? w / p.tiles_per_img : p.batch;  // the odd leftover: out of bounds -> zeros
                const int y0 = w < p.n_work ? (w - img * p.tiles_per_img) * p.R : 0;
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (leader) mbar_expect_tx(full0 + 8 * stage, 4 * p.box_bytes);
                const uint32_t dst = sA + stage * p.stage_bytes;
                tma_load_5d_2cta(dst, &map_a, full0 + 8 * stage, 0, 0, 0, y0 - 1, img);                   // E: columns 0, 2, .., W
                tma_load_5d_2cta(dst + p.plane_bytes, &map_a, full0 + 8 * stage, 0, -1, 1, y0 - 1, img);  // O: columns -1, 1, .., W-1
                if (!leader) mbar_arrive_leader(full0 + 8 * stage);
                if (++stage == (uint32_t)p.nstages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): M = 256 (both CTAs' tiles) =====
        if (leader) {
            constexpr uint32_t idesc_w = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            constexpr uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            constexpr uint64_t desc_hi = make_smem_desc_rowb<128>(0) & 0xFFFFFFFF00000000ull;
            mbar_wait(wbar, 0);
            tc_fence_after();
            uint32_t stage = 0, phase = 0;
            const uint32_t w_lo = sW >> 4;
            for (int ui = 0; ui < walk.count; ++ui) {
                const uint32_t slot = ui & (kF2wSlots - 1), use = ui / kF2wSlots;
                mbar_wait(full0 + 8 * stage, phase);
                mbar_wait(tempty0 + 8 * slot, (use & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t d = tmem_base + slot * 128;
                    const uint32_t a_e = (sA + stage * p.stage_bytes) >> 4, a_o = a_e + ((uint32_t)p.plane_bytes >> 4);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const bool wide = t == 1 || t == 2 || (dy == 0 && t == 0);
                            const uint32_t a_lo = ((t & 1) ? a_e : a_o) + (uint32_t)(dy * p.P + (t >> 1)) * 8;
                            const uint32_t b_lo = w_lo + (uint32_t)f2w_unit_of(dy, t) * (kF2wUnit / 16);
                            const uint32_t dd = d + ((!wide && t == 3) ? 64u : 0u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_2cta(dd, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), wide ? idesc_w : idesc_n,
                                               (dy | t | k) != 0);
                        }
                    }
                    umma_commit_2cta(empty0 + 8 * stage);
                    umma_commit_2cta(tfull0 + 8 * slot);
                }
                __syncwarp();
                if (++stage == (uint32_t)p.nstages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): lane quarter q x (even | odd) output of the pair =====
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const uint32_t stg = stage0 + (uint32_t)(warp - 4) * 4096;
        const int rr0 = lane >> 3, ch = lane & 7;
        const bool has_res = p.residual != nullptr;
        const int m = q * 32 + lane;
        const int mi = m / p.P, mj = m - mi * p.P;
        auto pix_of = [&](int u) {
            const int w = 2 * u + (int)rank;
            if (w >= p.n_work) return -1;
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * p.R;
            return (mj < p.Wp && mi < min(p.R, p.H - y0)) ? ((img * p.H + y0 + mi) * p.W + 2 * mj + half) : -1;
        };
        uint4 rres[8];
        auto load_res = [&](int px) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int pr = __shfl_sync(0xffffffffu, px, it * 4 + rr0);
                rres[it] = ldg_stream128(p.residual + (size_t)max(pr, 0) * p.cout + ch * 8);  // junk rows read pixel 0, unused
            }
        };
        auto stash_res = [&]() {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rr0;
                sts128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4), rres[it]);
            }
        };
        int pix = walk.count > 0 ? pix_of(walk.begin) : -1;
        if (walk.count > 0 && has_res) load_res(pix);
        for (int ui = 0, u = walk.begin; ui < walk.count; ++ui, u += walk.step) {
            const bool nlive = ui + 1 < walk.count;
            const int npix = nlive ? pix_of(u + walk.step) : -1;
            if (has_res) {
                stash_res();
                if (nlive) load_res(npix);
            }
            const uint32_t slot = ui & (kF2wSlots - 1), use = ui / kF2wSlots;
            __syncwarp();
            mbar_wait(tfull0 + 8 * slot, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + slot * 128 + half * 64 + ((uint32_t)(q * 32) << 16);
            const uint32_t srow = stg + lane * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tmem_ld32(taddr + h * 32, v);
                tmem_ld_wait();
                if (h == 1) {
                    tc_fence_before();
                    mbar_arrive_leader(tempty0 + 8 * slot);  // accumulator is in registers: hand the slot back to the leader
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t sa = srow + (((h * 4 + j) ^ (lane & 7)) << 4);
                    const float* bv = p.bias_v + h * 32 + j * 8;  // compile-time offsets: constant-bank operands
                    const float4 b0 = make_float4(bv[0], bv[1], bv[2], bv[3]);
                    const float4 b1 = make_float4(bv[4], bv[5], bv[6], bv[7]);
                    float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                                  __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                                  __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                                  __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
                    if (has_res && pix >= 0) {
                        const uint4 rv = lds128(sa);
                        const unsigned uu[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            f[2 * k] += __uint_as_float(uu[k] << 16);
                            f[2 * k + 1] += __uint_as_float(uu[k] & 0xffff0000u);
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                    }
                    uint4 o;
                    unsigned* ou = &o.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                        ou[k] = *reinterpret_cast<const unsigned*>(&h2);
                    }
                    sts128(sa, o);
                }
            }
            __syncwarp();
            {
                int prs[8];
                uint4 vals[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) prs[it] = __shfl_sync(0xffffffffu, pix, it * 4 + rr0);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rr0;
                    vals[it] = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
                }
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    if (prs[it] >= 0) *reinterpret_cast<uint4*>(p.out + (size_t)prs[it] * p.cout + ch * 8) = vals[it];
            }
            __syncwarp();
            pix = npix;
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory / barriers stay alive until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// flat128: the same halo-tile / shifted-view A operand, but 128 output channels per MMA (N=128 issues at
// the tensor floor, 64 cycles, where two N=64 MMAs cost 96) and the weights STREAMED per (tap, chunk)
// K-block through a ring, because a 128-channel slice of a 3x3x128 filter bank (288 KB) does not fit in
// shared memory.  A work tile is R rows = two 128-pixel M-tiles that share every weight K-block; the two
// accumulators (2 x 128 TMEM columns) are double-buffered across tiles.  Used for layer2's 3x3/s1 convs.
// Warp roles: 0 = A (activation) producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = B (weight) producer,
// 4..19 = epilogue (16 warps: 4 TMEM lane quarters x 2 M-tiles x 2 halves of the 128 channels; the
// epilogue of an N=128 tile is as long as its MMAs for one warp, so it is spread wide).
// ------------------------------------------------------------------------------------------
constexpr int kFlat128Threads = 128 + 16 * 32;  // 4 control warps + 16 epilogue warps

struct Flat128Params {
    int P, W, H, R, tiles_per_img, n_work, chunks, ns, cout, batch;
    int a_stages, a_stage_bytes, a_box_bytes, b_stages, slack_bytes;
    const float* bias;
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    int relu;
};

__global__ void __launch_bounds__(kFlat128Threads, 1)
flat128_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Flat128Params p) {
    constexpr int TAPS = 9, KW = 3, ROWB = 128, KSTEPS = 4;
    constexpr int BTILE = 128 * ROWB;  // one weight K-block: 128 cout rows x 64 k
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sB = sbase;
    const uint32_t sA = sB + p.b_stages * BTILE;
    const uint32_t stage0 = sA + p.a_stages * p.a_stage_bytes + p.slack_bytes;  // 16 epilogue warps x 4 KB
    const uint32_t bias0 = stage0 + 16 * 4096;                                 // 128 fp32
    const uint32_t bars = bias0 + 512;
    const uint32_t afull0 = bars, aempty0 = afull0 + 8 * p.a_stages;
    const uint32_t bfull0 = aempty0 + 8 * p.a_stages, bempty0 = bfull0 + 8 * p.b_stages;
    const uint32_t tfull0 = bempty0 + 8 * p.b_stages, tempty0 = tfull0 + 16, tslot = tempty0 + 16;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const int nslice = blockIdx.x % p.ns;
    // contiguous run of work tiles per CTA (balances the short last tile of each image, keeps halo rows in L2)
    const int n_cta = gridDim.x / p.ns, cta = blockIdx.x / p.ns;
    const int w_first = (int)((long long)p.n_work * cta / n_cta), w_last = (int)((long long)p.n_work * (cta + 1) / n_cta);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.a_stages; ++i) {
            mbar_init(afull0 + 8 * i, 1);
            mbar_init(aempty0 + 8 * i, 1);
        }
        for (int i = 0; i < p.b_stages; ++i) {
            mbar_init(bfull0 + 8 * i, 1);
            mbar_init(bempty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, 512);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tslot, 512);
    if (warp == 3) {
        float* bs = reinterpret_cast<float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        for (int i = lane; i < 128; i += 32) bs[i] = __ldg(p.bias + nslice * 128 + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===== A producer: one halo box per (tile, chunk) =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int w = w_first; w < w_last; ++w) {
                const int img = w / p.tiles_per_img;
                const int y0 = (w - img * p.tiles_per_img) * p.R;
                for (int c = 0; c < p.chunks; ++c) {
                    mbar_wait(aempty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(afull0 + 8 * stage, p.a_box_bytes);
                    tma_load_4d(sA + stage * p.a_stage_bytes, &map_a, afull0 + 8 * stage, c * 64, -1, y0 - 1, img);
                    if (++stage == (uint32_t)p.a_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===== B producer: the weight K-blocks in the order the MMA warp consumes them =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int w = w_first; w < w_last; ++w)
                for (int c = 0; c < p.chunks; ++c)
                    for (int tap = 0; tap < TAPS; ++tap) {
                        mbar_wait(bempty0 + 8 * stage, phase ^ 1);
                        mbar_expect_tx(bfull0 + 8 * stage, BTILE);
                        tma_load_2d(sB + stage * BTILE, &map_b, bfull0 + 8 * stage, (tap * p.chunks + c) * 64, nslice * 128);
                        if (++stage == (uint32_t)p.b_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc<128>();
        constexpr uint64_t desc_hi = make_smem_desc_rowb<ROWB>(0) & 0xFFFFFFFF00000000ull;
        uint32_t astage = 0, aphase = 0, bstage = 0, bphase = 0;
        int it = 0;
        for (int w = w_first; w < w_last; ++w, ++it) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * p.R;
            const int rows_valid = min(p.R, p.H - y0);
            const int n_mt = ((rows_valid - 1) * p.P + p.W - 1) / 128 + 1;
            const uint32_t set = it & 1, sphase = (it >> 1) & 1;
            mbar_wait(tempty0 + 8 * set, sphase ^ 1);
            tc_fence_after();
            for (int c = 0; c < p.chunks; ++c) {
                mbar_wait(afull0 + 8 * astage, aphase);
                tc_fence_after();
                const uint32_t a_lo_stage = (sA + astage * p.a_stage_bytes) >> 4;
                const uint32_t d0 = tmem_base + set * 256;
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(bfull0 + 8 * bstage, bphase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t b_lo = (sB + bstage * BTILE) >> 4;
                        const uint32_t a_lo = a_lo_stage + (uint32_t)((tap / KW) * p.P + (tap % KW)) * (ROWB / 16);
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            umma_bf16(d0, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (c | tap | k) != 0);
                        if (n_mt > 1) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                umma_bf16(d0 + 128, desc_hi | (uint64_t)(a_lo + 128 * (ROWB / 16) + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc,
                                          (c | tap | k) != 0);
                        }
                        umma_commit(bempty0 + 8 * bstage);
                    }
                    __syncwarp();
                    if (++bstage == (uint32_t)p.b_stages) {
                        bstage = 0;
                        bphase ^= 1;
                    }
                }
                if (elect_one_sync()) {
                    umma_commit(aempty0 + 8 * astage);
                    if (c == p.chunks - 1) umma_commit(tfull0 + 8 * set);
                }
                __syncwarp();
                if (++astage == (uint32_t)p.a_stages) {
                    astage = 0;
                    aphase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;
        const int mt = ((warp - 4) >> 2) & 1;  // which M-tile of the work tile this warp handles
        const int pass = (warp - 4) >> 3;      // which 64 of the 128 output channels
        const uint32_t stg = stage0 + (uint32_t)(warp - 4) * 4096;
        const float* bias_s = reinterpret_cast<const float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        const int cbase = nslice * 128;
        const int rr0 = lane >> 3, ch = lane & 7;
        const bool has_res = p.residual != nullptr;
        int it = 0;
        for (int w = w_first; w < w_last; ++w, ++it) {
            const int img = w / p.tiles_per_img;
            const int y0 = (w - img * p.tiles_per_img) * p.R;
            const int rows_valid = min(p.R, p.H - y0);
            const int n_mt = ((rows_valid - 1) * p.P + p.W - 1) / 128 + 1;
            const uint32_t set = it & 1, sphase = (it >> 1) & 1;
            const int m = mt * 128 + q * 32 + lane;
            const int i = m / p.P, x = m - i * p.P;
            const int pix = (mt < n_mt && x < p.W && i < rows_valid) ? ((img * p.H + y0 + i) * p.W + x) : -1;
            // the residual does not depend on the accumulator: its cp.async runs under the wait for the tile's MMAs (the
            // staging block is free: the previous tile's store-out is complete)
            if (has_res) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int rr = t * 4 + rr0;
                    const int pr = __shfl_sync(0xffffffffu, pix, rr);
                    if (pr >= 0)
                        cp_async16(stg + rr * 128 + ((ch ^ (rr & 7)) << 4), p.residual + (size_t)pr * p.cout + cbase + pass * 64 + ch * 8);
                }
                cp_async_commit();
            }
            mbar_wait(tfull0 + 8 * set, sphase);
            tc_fence_after();
            if (mt >= n_mt) {  // nothing of ours in this (short) tile
                tc_fence_before();
                mbar_arrive(tempty0 + 8 * set);
                continue;
            }
            const uint32_t taddr = tmem_base + set * 256 + mt * 128 + ((uint32_t)(q * 32) << 16);
            const uint32_t srow = stg + lane * 128;
            {
                if (has_res) {
                    cp_async_wait_all();
                    __syncwarp();
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(taddr + pass * 64 + h * 32, v);
                    tmem_ld_wait();
                    if (h == 1) {
                        tc_fence_before();
                        mbar_arrive(tempty0 + 8 * set);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t sa = srow + (((h * 4 + j) ^ (lane & 7)) << 4);
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + pass * 64 + h * 32 + j * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + pass * 64 + h * 32 + j * 8 + 4);
                        float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                                      __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                                      __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                                      __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
                        if (has_res && pix >= 0) {
                            const uint4 rv = lds128(sa);
                            const unsigned u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                f[2 * k] += __uint_as_float(u[k] << 16);
                                f[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                        }
                        uint4 o;
                        unsigned* ou = &o.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                            ou[k] = *reinterpret_cast<const unsigned*>(&h2);
                        }
                        sts128(sa, o);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int rr = t * 4 + rr0;
                    const int pr = __shfl_sync(0xffffffffu, pix, rr);
                    if (pr >= 0) {
                        const uint4 val = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
                        *reinterpret_cast<uint4*>(p.out + (size_t)pr * p.cout + cbase + pass * 64 + ch * 8) = val;
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// flat128x2: flat128_conv_kernel as a CTA PAIR (cta_group::2), as flat2_conv_kernel is to flat_conv_kernel: one
// M = 256 x N = 128 MMA per pair; each CTA brings its own work tile (A) and HALF of every weight K-block (64 of the
// 128 output channels of the slice).  Per MMA and CTA 4 KB of A + 2 KB of B leave the shared-memory pipe (the math
// takes 64 cycles, the single-CTA form's 8 KB took 64 cycles of fetch), and the streamed weights -- the larger part
// of this kernel's L2 -> SM traffic, 288 KB per work tile -- are halved.
// ------------------------------------------------------------------------------------------

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFlat128Threads, 1)
flat128x2_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Flat128Params p) {
    constexpr int TAPS = 9, KW = 3, ROWB = 128, KSTEPS = 4;
    constexpr int BTILE = 64 * ROWB;  // THIS CTA's half of a weight K-block: 64 of the 128 cout rows x 64 k
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sB = sbase;
    const uint32_t sA = sB + p.b_stages * BTILE;
    const uint32_t stage0 = sA + p.a_stages * p.a_stage_bytes + p.slack_bytes;  // 16 epilogue warps x 4 KB
    const uint32_t bias0 = stage0 + 16 * 4096;                                 // 128 fp32
    const uint32_t bars = bias0 + 512;
    const uint32_t afull0 = bars, aempty0 = afull0 + 8 * p.a_stages;
    const uint32_t bfull0 = aempty0 + 8 * p.a_stages, bempty0 = bfull0 + 8 * p.b_stages;
    const uint32_t tfull0 = bempty0 + 8 * p.b_stages, tempty0 = tfull0 + 16, tslot = tempty0 + 16;
    uint32_t* tslot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tslot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;
    const int nslice = pair % p.ns;
    // contiguous run of UNITS per pair; unit u = (image pair j = u / tiles_per_img, tile t = u % tiles_per_img): the leader
    // CTA works on tile t of image 2j, the peer on tile t of image 2j + 1 -- both tiles have the same height, so the
    // short last tile of an image (one M-tile instead of two) stays short for the whole M = 256 MMA
    const int n_pair = (gridDim.x >> 1) / p.ns, pidx = pair / p.ns;
    const int n_units = ((p.batch + 1) >> 1) * p.tiles_per_img;
    const int w_first = (int)((long long)n_units * pidx / n_pair), w_last = (int)((long long)n_units * (pidx + 1) / n_pair);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.a_stages; ++i) {
            mbar_init(afull0 + 8 * i, 2);   // leader's arrive.expect_tx + the peer producer's arrive
            mbar_init(aempty0 + 8 * i, 1);  // multicast commit
        }
        for (int i = 0; i < p.b_stages; ++i) {
            mbar_init(bfull0 + 8 * i, 2);
            mbar_init(bempty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);        // multicast commit
            mbar_init(tempty0 + 8 * i, 2 * 512);  // the 16 epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tslot, 512);
    if (warp == 3) {
        float* bs = reinterpret_cast<float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        for (int i = lane; i < 128; i += 32) bs[i] = __ldg(p.bias + nslice * 128 + i);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tslot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===== A producer: one halo box per (tile, chunk) =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int u = w_first; u < w_last; ++u) {
                const int j = u / p.tiles_per_img;
                const int img = 2 * j + (int)rank;  // == batch for the odd leftover: out of bounds -> zeros
                const int y0 = (u - j * p.tiles_per_img) * p.R;
                for (int c = 0; c < p.chunks; ++c) {
                    mbar_wait(aempty0 + 8 * stage, phase ^ 1);
                    if (leader) mbar_expect_tx(afull0 + 8 * stage, 2 * p.a_box_bytes);
                    tma_load_4d_2cta(sA + stage * p.a_stage_bytes, &map_a, afull0 + 8 * stage, c * 64, -1, y0 - 1, img);
                    if (!leader) mbar_arrive_leader(afull0 + 8 * stage);
                    if (++stage == (uint32_t)p.a_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===== B producer: the weight K-blocks in the order the MMA warp consumes them =====
        if (elect_one_sync()) {
            uint32_t stage = 0, phase = 0;
            for (int u = w_first; u < w_last; ++u)
                for (int c = 0; c < p.chunks; ++c)
                    for (int tap = 0; tap < TAPS; ++tap) {
                        mbar_wait(bempty0 + 8 * stage, phase ^ 1);
                        if (leader) mbar_expect_tx(bfull0 + 8 * stage, 2 * BTILE);
                        tma_load_2d_2cta(sB + stage * BTILE, &map_b, bfull0 + 8 * stage, (tap * p.chunks + c) * 64, nslice * 128 + (int)rank * 64);
                        if (!leader) mbar_arrive_leader(bfull0 + 8 * stage);
                        if (++stage == (uint32_t)p.b_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
        }
    } else if (warp == 1 && leader) {
        // ===== MMA issuer (leader CTA only): M = 256 (both CTAs' tiles), N = 128 =====
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        constexpr uint64_t desc_hi = make_smem_desc_rowb<ROWB>(0) & 0xFFFFFFFF00000000ull;
        uint32_t astage = 0, aphase = 0, bstage = 0, bphase = 0;
        int it = 0;
        for (int u = w_first; u < w_last; ++u, ++it) {
            const int y0 = (u % p.tiles_per_img) * p.R;
            const int rows_valid = min(p.R, p.H - y0);
            const int n_mt = ((rows_valid - 1) * p.P + p.W - 1) / 128 + 1;
            const uint32_t set = it & 1, sphase = (it >> 1) & 1;
            mbar_wait(tempty0 + 8 * set, sphase ^ 1);
            tc_fence_after();
            for (int c = 0; c < p.chunks; ++c) {
                mbar_wait(afull0 + 8 * astage, aphase);
                tc_fence_after();
                const uint32_t a_lo_stage = (sA + astage * p.a_stage_bytes) >> 4;
                const uint32_t d0 = tmem_base + set * 256;
#pragma unroll
                for (int tap = 0; tap < TAPS; ++tap) {
                    mbar_wait(bfull0 + 8 * bstage, bphase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t b_lo = (sB + bstage * BTILE) >> 4;
                        const uint32_t a_lo = a_lo_stage + (uint32_t)((tap / KW) * p.P + (tap % KW)) * (ROWB / 16);
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            umma_bf16_2cta(d0, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (c | tap | k) != 0);
                        if (n_mt > 1) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k)
                                umma_bf16_2cta(d0 + 128, desc_hi | (uint64_t)(a_lo + 128 * (ROWB / 16) + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc,
                                               (c | tap | k) != 0);
                        }
                        umma_commit_2cta(bempty0 + 8 * bstage);
                    }
                    __syncwarp();
                    if (++bstage == (uint32_t)p.b_stages) {
                        bstage = 0;
                        bphase ^= 1;
                    }
                }
                if (elect_one_sync()) {
                    umma_commit_2cta(aempty0 + 8 * astage);
                    if (c == p.chunks - 1) umma_commit_2cta(tfull0 + 8 * set);
                }
                __syncwarp();
                if (++astage == (uint32_t)p.a_stages) {
                    astage = 0;
                    aphase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;
        const int mt = ((warp - 4) >> 2) & 1;  // which M-tile of the work tile this warp handles
        const int pass = (warp - 4) >> 3;      // which 64 of the 128 output channels
        const uint32_t stg = stage0 + (uint32_t)(warp - 4) * 4096;
        const float* bias_s = reinterpret_cast<const float*>(smem_raw + (bias0 - smem_u32(smem_raw)));
        const int cbase = nslice * 128;
        const int rr0 = lane >> 3, ch = lane & 7;
        const bool has_res = p.residual != nullptr;
        int it = 0;
        for (int u = w_first; u < w_last; ++u, ++it) {
            const int j = u / p.tiles_per_img;
            const int img = 2 * j + (int)rank;
            const int y0 = (u - j * p.tiles_per_img) * p.R;
            const int rows_valid = min(p.R, p.H - y0);
            const int n_mt = ((rows_valid - 1) * p.P + p.W - 1) / 128 + 1;
            const bool real = img < p.batch;  // the odd leftover image stores nothing
            const uint32_t set = it & 1, sphase = (it >> 1) & 1;
            const int m = mt * 128 + q * 32 + lane;
            const int i = m / p.P, x = m - i * p.P;
            const int pix = (real && mt < n_mt && x < p.W && i < rows_valid) ? ((img * p.H + y0 + i) * p.W + x) : -1;
            // the residual does not depend on the accumulator: its cp.async runs under the wait for the tile's MMAs (the
            // staging block is free: the previous tile's store-out is complete)
            int prs[8];  // pixel index of the 8 rows this lane moves (residual in, result out), -1 = junk row
#pragma unroll
            for (int t = 0; t < 8; ++t) prs[t] = __shfl_sync(0xffffffffu, pix, t * 4 + rr0);
            if (has_res) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int rr = t * 4 + rr0;
                    if (prs[t] >= 0)
                        cp_async16(stg + rr * 128 + ((ch ^ (rr & 7)) << 4), p.residual + (size_t)prs[t] * p.cout + cbase + pass * 64 + ch * 8);
                }
                cp_async_commit();
            }
            mbar_wait(tfull0 + 8 * set, sphase);
            tc_fence_after();
            if (mt >= n_mt) {  // nothing of ours in this (short) tile
                tc_fence_before();
                mbar_arrive_leader(tempty0 + 8 * set);
                continue;
            }
            const uint32_t taddr = tmem_base + set * 256 + mt * 128 + ((uint32_t)(q * 32) << 16);
            const uint32_t srow = stg + lane * 128;
            {
                if (has_res) {
                    cp_async_wait_all();
                    __syncwarp();
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(taddr + pass * 64 + h * 32, v);
                    tmem_ld_wait();
                    if (h == 1) {
                        tc_fence_before();
                        mbar_arrive_leader(tempty0 + 8 * set);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t sa = srow + (((h * 4 + j) ^ (lane & 7)) << 4);
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + pass * 64 + h * 32 + j * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + pass * 64 + h * 32 + j * 8 + 4);
                        float f[8] = {__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                                      __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                                      __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                                      __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
                        if (has_res && pix >= 0) {
                            const uint4 rv = lds128(sa);
                            const unsigned u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                f[2 * k] += __uint_as_float(u[k] << 16);
                                f[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                        }
                        uint4 o;
                        unsigned* ou = &o.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                            ou[k] = *reinterpret_cast<const unsigned*>(&h2);
                        }
                        sts128(sa, o);
                    }
                }
                __syncwarp();
                {
                    uint4 vals[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int rr = t * 4 + rr0;
                        vals[t] = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    }
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (prs[t] >= 0) *reinterpret_cast<uint4*>(p.out + (size_t)prs[t] * p.cout + cbase + pass * 64 + ch * 8) = vals[t];
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory / barriers stay alive until both CTAs are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
constexpr int kSmemMax = 227 * 1024;


// Shared-memory plan for a given R; returns the total dynamic smem bytes, or 0 if it cannot work.
template <int ROWB, int KH, int KW>
static int plan_flat(FlatParams& p, bool pool = false) {
    const int stage_rows = (p.R + KH - 1) * p.P;
    p.box_bytes = stage_rows * ROWB;
    p.stage_bytes = (p.box_bytes + 1023) & ~1023;
    p.w_bytes = (KH * KW * p.chunks * 64 * ROWB + 1023) & ~1023;
    const int n_mt_max = ((p.R - 1) * p.P + p.W - 1) / 128 + 1;
    if (n_mt_max > kFlatSlots && p.chunks > 1) return 0;  // multi-chunk tiles keep all their accumulators live
    const int reach_rows = n_mt_max * 128 + (KH - 1) * p.P + (KW - 1);  // rows a (junk) view may touch
    p.slack_bytes = (std::max(0, reach_rows * ROWB - p.stage_bytes) + 1023) & ~1023;
    const int bar_bytes = 8 * (2 * 8 + 3 * kFlatSlots + 2) + 64;
    const int fixed = 1024 + p.w_bytes + p.slack_bytes + (pool ? kPoolRingBytes + 256 : kEpiBytes) + bar_bytes;
    p.nstages = std::min(p.chunks > 1 ? 4 : 6, (kSmemMax - fixed) / p.stage_bytes);
    if (p.nstages < 2) return 0;
    return fixed + p.nstages * p.stage_bytes;
}

template <int ROWB, int KH, int KW, bool POOL>
static int launch_flat(fx_engine* e, const CUtensorMap& ma, const CUtensorMap& mb, FlatParams& p, int smem, cudaStream_t stream) {
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(flat_conv_kernel<ROWB, KH, KW, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        attr_done[e->device & 255] = true;
    }
    int grid = std::min(e->sm_count, p.n_work * p.ns);
    grid -= grid % p.ns;
    if (grid < p.ns) grid = p.ns;
    FX_CUDA(e, launch_pdl(flat_conv_kernel<ROWB, KH, KW, POOL>, dim3(grid), dim3(POOL ? kFlatPoolThreads : kFlatThreads), smem, stream, ma, mb, p));
    FX_LAUNCH_CHECK(e, "flat_conv_kernel");
    return FX_OK;
}

// The pooled stem with two conv outputs per accumulator row: FX_STEMW = 2 (default) the CTA-pair kernel stemw2_conv_kernel,
// 1 the single-CTA stemw_conv_kernel, 0 flat_conv_kernel<32, 4, 4, true>.
static int stemw_mode() {
    static const int mode = [] {
        const char* v = getenv("FX_STEMW");
        return v ? atoi(v) : 2;
    }();
    return mode;
}

static int stemw_conv(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, __nv_bfloat16* out, int n, cudaStream_t stream, int sched) {
    const LayerGeom& g = L.g;
    static_assert(kS2dW % 2 == 0 && kS2dW / 2 >= kStemwP && kS2dC == 16, "stemw: the staging tensor is read as pairs of s2d pixels");
    if (g.cout != 64 || L.host_b.size() != 64 || g.hout != 112 || g.wout != 2 * kStemwPairs || L.k_bf16 != 256)
        return set_error(e, FX_ERR_UNSUPPORTED, "stemw_conv: not the 7x7 / stride-2 stem on a 224 x 224 crop");
    StemwParams p;
    std::memset(&p, 0, sizeof(p));
    std::memcpy(p.bias_v, L.host_b.data(), sizeof(p.bias_v));
    p.out = out;
    p.sched = sched;
    p.tiles_per_img = g.hout / 8;
    p.n_work = n * p.tiles_per_img;
    p.Hp = g.hout / 2;
    p.Wp = g.wout / 2;
    CUtensorMap ma, mb;
    const uint32_t ones[4] = {1, 1, 1, 1};
    const uint64_t dims[4] = {2 * kS2dC, kS2dW / 2, (uint64_t)kS2dH, (uint64_t)n};
    const uint64_t strides[3] = {2 * kS2dC * 2, (uint64_t)kS2dW * kS2dC * 2, (uint64_t)kS2dH * kS2dW * kS2dC * 2};
    const uint32_t box[4] = {2 * kS2dC, kStemwP, kStemwR + 3, 1};
    int rc = tc_encode_map(e, &ma, in, 4, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_64B, "stemw A");
    if (rc != FX_OK) return rc;
    const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
    const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
    const uint32_t bbox[2] = {16, 64};
    rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, ones, CU_TENSOR_MAP_SWIZZLE_32B, "stemw B");
    if (rc != FX_OK) return rc;
    constexpr int kSmem = 1024 + kStemwWBytes + 2 * kStemwStage + kStemwRing * kStemwRowBytes + 256 + 8 * (4 + 3 * kStemwSlots + 3) + 64;
    constexpr int kSmem2 = 1024 + kStemw2WBytes + kStemw2Stages * kStemwStage + kStemw2Ring * (kStemw2RowBytes + 256) + 256 +
                           8 * (2 * kStemw2Stages + 3 * kStemwSlots + 3) + 64;
    static_assert(kSmem2 <= kSmemMax, "stemw2_conv_kernel: shared memory");
    static_assert(kSmem <= kSmemMax, "stemw_conv_kernel: shared memory");
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(stemw_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        FX_CUDA(e, cudaFuncSetAttribute(stemw2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
        attr_done[e->device & 255] = true;
    }
    if (stemw_mode() >= 2 && p.n_work % 2 == 0) {  // CTA pairs: unit = two work tiles
        CUtensorMap mbn;
        const uint32_t nbox[2] = {16, 32};
        rc = tc_encode_map(e, &mbn, L.w_bf16, 2, bd, bs, nbox, ones, CU_TENSOR_MAP_SWIZZLE_32B, "stemw B (32 channels)");
        if (rc != FX_OK) return rc;
        const int pairs = std::max(1, std::min(e->sm_count / 2, p.n_work / 2));
        FX_CUDA(e, launch_pdl(stemw2_conv_kernel, dim3(2 * pairs), dim3(kStemwThreads), kSmem2, stream, ma, mb, mbn, p));  // cluster dims are a kernel attribute
        FX_LAUNCH_CHECK(e, "stemw2_conv_kernel");
        return FX_OK;
    }
    const int grid = std::max(1, std::min(e->sm_count, p.n_work));
    FX_CUDA(e, launch_pdl(stemw_conv_kernel, dim3(grid), dim3(kStemwThreads), kSmem, stream, ma, mb, p));
    FX_LAUNCH_CHECK(e, "stemw_conv_kernel");
    return FX_OK;
}

// flat128 on CTA pairs (flat128x2_conv_kernel); FX_FLAT128X2=0 keeps the single-CTA kernel.
static bool flat128x2_enabled() {
    static const bool on = [] {
        const char* v = getenv("FX_FLAT128X2");
        return !(v && v[0] == '0');
    }();
    return on;
}

static int flat128_conv(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                        int n, int relu, cudaStream_t stream) {
    const LayerGeom& g = L.g;
    Flat128Params p;
    std::memset(&p, 0, sizeof(p));
    p.bias = L.bias;
    p.residual = residual;
    p.out = out;
    p.relu = relu;
    p.cout = g.cout;
    p.ns = g.cout / 128;
    p.W = g.wout;
    p.H = g.hout;
    p.P = g.win + 2;
    p.chunks = g.cin / 64;
    p.R = std::max(1, std::min(g.hout, 256 / p.P));  // two 128-pixel M-tiles per work tile
    p.tiles_per_img = (p.H + p.R - 1) / p.R;
    p.n_work = n * p.tiles_per_img;
    p.a_box_bytes = (p.R + 2) * p.P * 128;
    p.a_stage_bytes = (p.a_box_bytes + 1023) & ~1023;
    const int reach_rows = 256 + 2 * p.P + 2;
    p.slack_bytes = (std::max(0, reach_rows * 128 - p.a_stage_bytes) + 1023) & ~1023;
    p.a_stages = 2;  // chunk granularity: the next tile's chunk c loads as soon as this tile's chunk c is consumed
    const int fixed = 1024 + p.a_stages * p.a_stage_bytes + p.slack_bytes + 16 * 4096 + 512 + 8 * (2 * 3 + 2 * 8 + 4) + 64;
    const bool pair = flat128x2_enabled();
    const int btile = pair ? 64 * 128 : 128 * 128;  // the CTA-pair kernel streams half a weight K-block per CTA
    p.batch = n;
    p.b_stages = std::min(pair ? 10 : 6, (kSmemMax - fixed) / btile);
    if (p.b_stages < 2) return set_error(e, FX_ERR_UNSUPPORTED, "flat128_conv: tile does not fit in shared memory");
    const int smem = fixed + p.b_stages * btile;
    CUtensorMap ma, mb;
    const uint32_t ones[4] = {1, 1, 1, 1};
    const uint64_t dims[4] = {(uint64_t)g.cin, (uint64_t)g.win, (uint64_t)g.hin, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)g.cin * 2, (uint64_t)g.win * g.cin * 2, (uint64_t)g.hin * g.win * g.cin * 2};
    const uint32_t box[4] = {64, (uint32_t)p.P, (uint32_t)(p.R + 2), 1};
    int rc = tc_encode_map(e, &ma, in, 4, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat128 A");
    if (rc != FX_OK) return rc;
    const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
    const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
    const uint32_t bbox[2] = {64, (uint32_t)(pair ? 64 : 128)};
    rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat128 B");
    if (rc != FX_OK) return rc;
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(flat128_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        FX_CUDA(e, cudaFuncSetAttribute(flat128x2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        attr_done[e->device & 255] = true;
    }
    if (pair) {
        const int n_units = ((n + 1) / 2) * p.tiles_per_img;
        int pairs = std::min(e->sm_count / 2, n_units * p.ns);
        pairs -= pairs % p.ns;
        if (pairs < p.ns) pairs = p.ns;
        FX_CUDA(e, launch_pdl(flat128x2_conv_kernel, dim3(2 * pairs), dim3(kFlat128Threads), smem, stream, ma, mb, p));  // cluster dims are a kernel attribute
        FX_LAUNCH_CHECK(e, "flat128x2_conv_kernel");
        return FX_OK;
    }
    int grid = std::min(e->sm_count, p.n_work * p.ns);
    grid -= grid % p.ns;
    if (grid < p.ns) grid = p.ns;
    FX_CUDA(e, launch_pdl(flat128_conv_kernel, dim3(grid), dim3(kFlat128Threads), smem, stream, ma, mb, p));
    FX_LAUNCH_CHECK(e, "flat128_conv_kernel");
    return FX_OK;
}

// Layer 1 on CTA pairs (flat2_conv_kernel): measured 78 -> 62 us per launch for the convs WITHOUT a residual (batch 256).
// With the residual prefetched into registers and the bias in the constant bank the convs WITH a residual are faster on
// it as well (74.3 -> 72.7 us; with the old cp.async residual they were HBM-latency-bound either way, 85 vs 87 us).
// FX_FLAT2 = 0 disables it, = 1 keeps the residual convs on flat_conv_kernel (measurement knobs).
static int flat2_mode() {
    static const int mode = [] {
        const char* v = getenv("FX_FLAT2");
        return v ? atoi(v) : 2;
    }();
    return mode;
}

static int flat2_conv(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                      int n, int relu, cudaStream_t stream, int sched) {
    const LayerGeom& g = L.g;
    Flat2Params p;
    std::memset(&p, 0, sizeof(p));
    p.sched = sched;
    if (g.cout != 64 || L.host_b.size() != 64) return set_error(e, FX_ERR_UNSUPPORTED, "flat2_conv: cout must be 64");
    std::memcpy(p.bias_v, L.host_b.data(), sizeof(p.bias_v));
    p.residual = residual;
    p.out = out;
    p.relu = relu;
    p.cout = g.cout;
    p.batch = n;
    p.W = g.wout;
    p.H = g.hout;
    p.P = g.win + 2;
    p.R = std::max(1, std::min(g.hout, 256 / p.P));
    p.tiles_per_img = (p.H + p.R - 1) / p.R;
    p.n_work = n * p.tiles_per_img;
    p.box_bytes = (p.R + 2) * p.P * 128;
    p.stage_bytes = (p.box_bytes + 1023) & ~1023;
    const int n_mt = ((p.R - 1) * p.P + p.W - 1) / 128 + 1;
    const int reach_rows = n_mt * 128 + 2 * p.P + 2;
    p.slack_bytes = (std::max(0, reach_rows * 128 - p.stage_bytes) + 1023) & ~1023;
    const int bar_bytes = 8 * (2 * 8 + 2 * kFlatSlots + 2) + 64;
    const int fixed = 1024 + 9 * 32 * 128 + p.slack_bytes + kEpiBytes + bar_bytes;
    p.nstages = std::min(4, (kSmemMax - fixed) / p.stage_bytes);
    if (p.nstages < 2 || n_mt > kFlatSlots) return set_error(e, FX_ERR_UNSUPPORTED, "flat2_conv: tile does not fit");
    const int smem = fixed + p.nstages * p.stage_bytes;
    CUtensorMap ma, mb;
    const uint32_t ones[4] = {1, 1, 1, 1};
    const uint64_t dims[4] = {(uint64_t)g.cin, (uint64_t)g.win, (uint64_t)g.hin, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)g.cin * 2, (uint64_t)g.win * g.cin * 2, (uint64_t)g.hin * g.win * g.cin * 2};
    const uint32_t box[4] = {64, (uint32_t)p.P, (uint32_t)(p.R + 2), 1};
    int rc = tc_encode_map(e, &ma, in, 4, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat2 A");
    if (rc != FX_OK) return rc;
    const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
    const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
    const uint32_t bbox[2] = {64, 32};  // one tap's K-block of HALF of the output channels
    rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat2 B");
    if (rc != FX_OK) return rc;
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(flat2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        attr_done[e->device & 255] = true;
    }
    const int n_units = (p.n_work + 1) / 2;
    const int pairs = std::max(1, std::min(e->sm_count / 2, n_units));
    FX_CUDA(e, launch_pdl(flat2_conv_kernel, dim3(2 * pairs), dim3(kFlat2Threads), smem, stream, ma, mb, p));  // cluster dims are a kernel attribute
    FX_LAUNCH_CHECK(e, "flat2_conv_kernel");
    return FX_OK;
}

// layer 1 on flat2w_conv_kernel (two outputs per accumulator row); FX_FLAT2W=0 keeps flat2_conv_kernel.
static bool flat2w_supported(const LayerGeom& g, bool has_residual) {
    static const bool on = [] {
        const char* v = getenv("FX_FLAT2W");
        return !(v && v[0] == '0');
    }();
    return on && flat2_mode() >= (has_residual ? 2 : 1) && g.kh == 3 && g.kw == 3 && g.stride == 1 && g.pad == 1 && g.cin == 64 && g.cout == 64 && g.win % 2 == 0 &&
           g.win >= 8 && g.win / 2 + 1 <= 128 - g.win / 2 && g.hin == g.hout && g.win == g.wout;
}

static int flat2w_conv(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                       int n, int relu, cudaStream_t stream, int sched) {
    const LayerGeom& g = L.g;
    Flat2wParams p;
    std::memset(&p, 0, sizeof(p));
    p.sched = sched;
    if (L.host_b.size() != 64) return set_error(e, FX_ERR_UNSUPPORTED, "flat2w_conv: cout must be 64");
    std::memcpy(p.bias_v, L.host_b.data(), sizeof(p.bias_v));
    p.residual = residual;
    p.out = out;
    p.relu = relu;
    p.cout = g.cout;
    p.batch = n;
    p.W = g.wout;
    p.H = g.hout;
    p.Wp = g.wout / 2;
    p.P = p.Wp + 1;
    p.R = std::max(1, std::min(g.hout, (128 - p.Wp) / p.P + 1));  // (R - 1) * P + Wp <= 128: one M-tile per work tile
    p.tiles_per_img = (p.H + p.R - 1) / p.R;
    p.n_work = n * p.tiles_per_img;
    p.box_bytes = (p.R + 2) * p.P * 128;
    p.plane_bytes = (p.box_bytes + 1023) & ~1023;
    p.stage_bytes = 2 * p.plane_bytes;
    const int bar_bytes = 8 * (2 * 4 + 2 * kF2wSlots + 2) + 64;
    const int fixed = 1024 + kF2wUnits * kF2wUnit + 8 * 4096 + bar_bytes;
    p.nstages = std::min(4, (kSmemMax - fixed) / p.stage_bytes);
    // the junk rows of a plane read up to 127 + 2P + 1 rows in: the last plane's over-read must stay inside the staging block
    if (p.nstages < 2 || (128 + 2 * p.P + 2) * 128 > p.plane_bytes + 8 * 4096)
        return set_error(e, FX_ERR_UNSUPPORTED, "flat2w_conv: tile does not fit");
    const int smem = fixed + p.nstages * p.stage_bytes;
    CUtensorMap ma, mb;
    const uint32_t ones[5] = {1, 1, 1, 1, 1};
    const uint64_t rowb = (uint64_t)g.win * g.cin * 2;
    const uint64_t dims[5] = {(uint64_t)g.cin, (uint64_t)g.win / 2, 2, (uint64_t)g.hin, (uint64_t)n};
    const uint64_t strides[4] = {(uint64_t)g.cin * 4, (uint64_t)g.cin * 2, rowb, (uint64_t)g.hin * rowb};
    const uint32_t box[5] = {64, (uint32_t)p.P, 1, (uint32_t)(p.R + 2), 1};
    int rc = tc_encode_map(e, &ma, in, 5, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat2w A (column-parity planes)");
    if (rc != FX_OK) return rc;
    const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
    const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
    const uint32_t bbox[2] = {64, 32};  // one tap's K-block of 32 output channels
    rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat2w B");
    if (rc != FX_OK) return rc;
    static bool attr_done[256] = {};  // per device ordinal
    if (!attr_done[e->device & 255]) {
        FX_CUDA(e, cudaFuncSetAttribute(flat2w_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        attr_done[e->device & 255] = true;
    }
    const int n_units = (p.n_work + 1) / 2;
    const int pairs = std::max(1, std::min(e->sm_count / 2, n_units));
    FX_CUDA(e, launch_pdl(flat2w_conv_kernel, dim3(2 * pairs), dim3(kFlat2Threads), smem, stream, ma, mb, p));  // cluster dims are a kernel attribute
    FX_LAUNCH_CHECK(e, "flat2w_conv_kernel");
    return FX_OK;
}

static bool flat2_supported(const LayerGeom& g, bool has_residual) {
    return flat2_mode() >= (has_residual ? 2 : 1) && g.kh == 3 && g.kw == 3 && g.stride == 1 && g.pad == 1 && g.cin == 64 && g.cout == 64 && g.win + 2 <= 128 && g.win >= 8;
}

static bool flat128_supported(const LayerGeom& g) {
    return g.kh == 3 && g.kw == 3 && g.stride == 1 && g.pad == 1 && g.cin % 64 == 0 && g.cout % 128 == 0 && g.win + 2 <= 64 &&
           g.win >= 16;
}

// Can this conv+bn group run on the flat kernel?
bool flat_supported(const LayerGeom& g) {
    if (flat128_supported(g)) return true;
    if (g.cin == 3) return g.kh == 7 && g.kw == 7 && g.stride == 2 && g.pad == 3 && g.hin == kCrop && g.win == kCrop && g.cout == 64;
    return g.kh == 3 && g.kw == 3 && g.stride == 1 && g.pad == 1 && g.cin % 64 == 0 && g.cin <= 128 && g.cout == 64 &&
           g.win + 2 <= 128 && g.win >= 8;
}

int flat_conv(fx_engine* e, const PackedLayer& L, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out, int n,
              int relu, bool pool, cudaStream_t stream, int sched) {
    const LayerGeom& g = L.g;
    FlatParams p;
    std::memset(&p, 0, sizeof(p));
    p.sched = sched;
    if (!flat128_supported(g)) {  // stem and layer 1: the bias is part of the kernel parameters
        if (g.cout != 64 || L.host_b.size() != 64) return set_error(e, FX_ERR_UNSUPPORTED, "flat_conv: cout must be 64");
        std::memcpy(p.bias_v, L.host_b.data(), sizeof(p.bias_v));
    }
    p.residual = residual;
    p.out = out;
    p.relu = relu;
    p.cout = g.cout;
    p.ns = g.cout / 64;
    p.W = g.wout;
    p.H = g.hout;
    CUtensorMap ma, mb;
    const uint32_t ones[4] = {1, 1, 1, 1};
    if (g.cin == 3) {
        // stem in space-to-depth form: a 4x4 stride-1 conv over [115][116][16] (see engine.cu pack)
        p.P = kS2dW;
        p.chunks = 1;
        p.x0 = 0;
        p.ypad = 0;
        if (pool) {
            if (!relu || residual) return set_error(e, FX_ERR_INVALID, "flat_conv: the pooled stem is conv+bn+relu+maxpool");
            if (stemw_mode() >= 1) return stemw_conv(e, L, in, out, n, stream, sched);
            p.R = 9;  // 4 pooled rows need conv rows 8t-1 .. 8t+7
            p.rstep = 8;
            p.yfirst = -1;
            p.tiles_per_img = p.H / 8;
        } else {
            p.R = 8;
            p.rstep = 8;
            p.yfirst = 0;
            p.tiles_per_img = (p.H + p.R - 1) / p.R;
        }
        const int smem = plan_flat<32, 4, 4>(p, pool);
        if (!smem) return set_error(e, FX_ERR_UNSUPPORTED, "flat_conv: stem tile does not fit in shared memory");
        const uint64_t dims[4] = {(uint64_t)kS2dC, (uint64_t)kS2dW, (uint64_t)kS2dH, (uint64_t)n};
        const uint64_t strides[3] = {(uint64_t)kS2dC * 2, (uint64_t)kS2dW * kS2dC * 2, (uint64_t)kS2dH * kS2dW * kS2dC * 2};
        const uint32_t box[4] = {(uint32_t)kS2dC, (uint32_t)p.P, (uint32_t)(p.R + 3), 1};
        int rc = tc_encode_map(e, &ma, in, 4, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_32B, "stem A");
        if (rc != FX_OK) return rc;
        const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
        const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
        const uint32_t bbox[2] = {16, 64};
        rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, ones, CU_TENSOR_MAP_SWIZZLE_32B, "stem B");
        if (rc != FX_OK) return rc;
        p.n_work = n * p.tiles_per_img;
        return pool ? launch_flat<32, 4, 4, true>(e, ma, mb, p, smem, stream) : launch_flat<32, 4, 4, false>(e, ma, mb, p, smem, stream);
    }
    if (pool) return set_error(e, FX_ERR_INVALID, "flat_conv: only the stem has a fused max-pool");
    if (flat128_supported(g)) return flat128_conv(e, L, in, residual, out, n, relu, stream);
    if (flat2w_supported(g, residual != nullptr)) return flat2w_conv(e, L, in, residual, out, n, relu, stream, sched);
    if (flat2_supported(g, residual != nullptr)) return flat2_conv(e, L, in, residual, out, n, relu, stream, sched);
    p.P = g.win + 2;
    p.chunks = g.cin / 64;
    p.x0 = -1;
    p.ypad = 1;
    // tallest tile (fewest halo re-reads) of at most two 128-pixel M-tiles that fits next to the weights
    int smem = 0;
    for (p.R = std::max(1, std::min(g.hout, 256 / p.P)); p.R >= 1; --p.R)
        if ((smem = plan_flat<128, 3, 3>(p)) != 0) break;
    if (!smem) return set_error(e, FX_ERR_UNSUPPORTED, "flat_conv: layer does not fit in shared memory");
    p.rstep = p.R;
    p.yfirst = 0;
    p.tiles_per_img = (p.H + p.R - 1) / p.R;
    p.n_work = n * p.tiles_per_img;
    const uint64_t dims[4] = {(uint64_t)g.cin, (uint64_t)g.win, (uint64_t)g.hin, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)g.cin * 2, (uint64_t)g.win * g.cin * 2, (uint64_t)g.hin * g.win * g.cin * 2};
    const uint32_t box[4] = {64, (uint32_t)p.P, (uint32_t)(p.R + 2), 1};
    int rc = tc_encode_map(e, &ma, in, 4, dims, strides, box, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat A");
    if (rc != FX_OK) return rc;
    const uint64_t bd[2] = {(uint64_t)L.k_bf16, (uint64_t)g.cout};
    const uint64_t bs[1] = {(uint64_t)L.k_bf16 * 2};
    const uint32_t bbox[2] = {64, 64};
    rc = tc_encode_map(e, &mb, L.w_bf16, 2, bd, bs, bbox, ones, CU_TENSOR_MAP_SWIZZLE_128B, "flat B");
    if (rc != FX_OK) return rc;
    return launch_flat<128, 3, 3, false>(e, ma, mb, p, smem, stream);
}

}  // namespace fx
