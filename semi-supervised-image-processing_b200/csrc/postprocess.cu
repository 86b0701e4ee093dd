// On-device post-processing of the gathered [n][d] fp32 embedding matrix.
//
// Reference lines replaced (all numpy / scikit-learn on the host, multi-GB scans at N = 1M):
//   run_sanity_checks       src/feature_extraction.py:334-356   NaN / Inf guard, mean(|column mean|), mean(column std)
//   nearest_neighbor_probe  src/feature_extraction.py:359-398   cosine nearest neighbour of 8 sampled rows
//   StandardScaler fit      src/standardize_features.py:41-43   column mean / scale, z = (x - mean) / scale
//
// All three are HBM-bound single passes over the matrix (2 KB per image).  Determinism: every reduction has a
// fixed shape -- a block owns a contiguous run of rows and sums them in row order, block partials are combined in
// block order -- so results do not depend on timing, and column statistics accumulate in fp64 (as scikit-learn's
// StandardScaler does; numpy's fp32 row-by-row accumulation in run_sanity_checks drifts by ~n * 2^-24).
#include <cfloat>
#include <cmath>

#include "fx_common.cuh"

namespace fx {

constexpr int kStatRowsPerBlock = 256;
constexpr int kStatThreads = 256;

// pass 1: per block, per column: sum over the block's rows (fp64) + NaN / Inf counts
__global__ void colsum_kernel(const float* __restrict__ x, long long n, int d, double* __restrict__ part, unsigned long long* __restrict__ bad) {
    const long long r0 = (long long)blockIdx.x * kStatRowsPerBlock, r1 = min(n, r0 + kStatRowsPerBlock);
    unsigned long long nan_c = 0, inf_c = 0;
    for (int c = threadIdx.x; c < d; c += kStatThreads) {
        double s = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const float v = x[r * d + c];
            nan_c += isnan(v);
            inf_c += isinf(v);
            s += (double)v;
        }
        part[(size_t)blockIdx.x * d + c] = s;
    }
    if (nan_c) atomicAdd(bad, nan_c);  // integer counts: order-independent
    if (inf_c) atomicAdd(bad + 1, inf_c);
}

// combine block partials in block order -> column mean (fp64)
__global__ void colmean_kernel(const double* __restrict__ part, int blocks, int d, long long n, double* __restrict__ mean) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += part[(size_t)b * d + c];
    mean[c] = s / (double)n;
}

// pass 2: per block, per column: sum of squared deviations from the column mean (two-pass variance)
__global__ void colsq_kernel(const float* __restrict__ x, long long n, int d, const double* __restrict__ mean, double* __restrict__ part) {
    const long long r0 = (long long)blockIdx.x * kStatRowsPerBlock, r1 = min(n, r0 + kStatRowsPerBlock);
    for (int c = threadIdx.x; c < d; c += kStatThreads) {
        const double m = mean[c];
        double s = 0.0;
        for (long long r = r0; r < r1; ++r) {
            const double t = (double)x[r * d + c] - m;
            s += t * t;
        }
        part[(size_t)blockIdx.x * d + c] = s;
    }
}

// column std (population, ddof = 0) + the two scalars of run_sanity_checks
__global__ void colstd_kernel(const double* __restrict__ part, int blocks, int d, long long n, const double* __restrict__ mean,
                              double* __restrict__ stdv, double* __restrict__ var_out, double* __restrict__ scalars) {
    __shared__ double s_abs[256], s_std[256];
    double a = 0.0, sd = 0.0;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < blocks; ++b) s += part[(size_t)b * d + c];
        const double var = s / (double)n;
        if (var_out) var_out[c] = var;
        const double st = sqrt(var);
        if (stdv) stdv[c] = st;
        a += fabs(mean[c]);
        sd += st;
    }
    s_abs[threadIdx.x] = a;
    s_std[threadIdx.x] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, ts = 0.0;
        for (int i = 0; i < (int)blockDim.x; ++i) {  // fixed order
            ta += s_abs[i];
            ts += s_std[i];
        }
        scalars[0] = ta / d;
        scalars[1] = ts / d;
    }
}

// StandardScaler.transform as scikit-learn >= 1.3 executes it on a float32 matrix (sklearn/preprocessing/_data.py:
// `X -= astype(mean_, X.dtype); X /= astype(scale_, X.dtype)`): mean and scale rounded to fp32, then fp32 IEEE
// subtraction and division -- reproduced bit for bit.
__global__ void standardize_kernel(const float* __restrict__ x, long long total, int d, const double* __restrict__ mean,
                                   const double* __restrict__ scale, float* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % d);
        const float t = __fsub_rn(x[i], (float)mean[c]);
        out[i] = __fdiv_rn(t, (float)scale[c]);
    }
}

// max(||row||, 1e-12) in fp32 (src/feature_extraction.py:380-381); one warp per row.  Rows are then DIVIDED by it, as
// the reference does (:382), not multiplied by a reciprocal.
__global__ void rownorm_kernel(const float* __restrict__ x, long long n, int d, float* __restrict__ inv) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float v = x[r * d + k];
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) inv[r] = fmaxf(sqrtf(s), 1e-12f);
}

// Cosine similarity of every row against q (<= 32) query rows; a block owns kStatRowsPerBlock rows (one warp per row
// at a time) and keeps, per query, the best (similarity, row) it saw -- ties go to the smaller row, as np.argmax.
constexpr int kProbeMaxQ = 32;
__global__ void probe_kernel(const float* __restrict__ x, long long n, int d, const float* __restrict__ inv, const long long* __restrict__ qidx,
                             int q, float* __restrict__ best_sim, long long* __restrict__ best_row) {
    extern __shared__ float s_q[];  // [q][d] normalised query rows
    __shared__ float s_best[8][kProbeMaxQ];
    __shared__ long long s_row[8][kProbeMaxQ];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < q * d; i += blockDim.x) {
        const long long qr = qidx[i / d];
        s_q[i] = x[qr * d + (i % d)] / inv[qr];
    }
    __syncthreads();
    float my_best = -INFINITY;  // lane j tracks query j
    long long my_row = -1;
    const long long r0 = (long long)blockIdx.x * kStatRowsPerBlock, r1 = min(n, r0 + kStatRowsPerBlock);
    for (long long r = r0 + warp; r < r1; r += nwarp) {
        const float iv = inv[r];
        for (int j = 0; j < q; ++j) {
            float s = 0.f;
            for (int k = lane; k < d; k += 32) s = fmaf(x[r * d + k] / iv, s_q[j * d + k], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == j && r != qidx[j] && s > my_best) {  // rows ascend within a warp: strict > keeps the first maximum
                my_best = s;
                my_row = r;
            }
        }
    }
    if (lane < q) {
        s_best[warp][lane] = my_best;
        s_row[warp][lane] = my_row;
    }
    __syncthreads();
    if (warp == 0 && lane < q) {
        float b = -INFINITY;
        long long br = -1;
        for (int w = 0; w < nwarp; ++w) {
            const float v = s_best[w][lane];
            const long long vr = s_row[w][lane];
            if (vr >= 0 && (v > b || (v == b && vr < br) || br < 0)) {
                b = v;
                br = vr;
            }
        }
        best_sim[(size_t)blockIdx.x * q + lane] = b;
        best_row[(size_t)blockIdx.x * q + lane] = br;
    }
}

static int scratch_reserve(fx_engine* e, size_t bytes) {
    if (bytes <= e->post_cap) return FX_OK;
    cudaFree(e->post_scratch);
    e->post_scratch = nullptr;
    e->post_cap = 0;
    cudaError_t a = cudaMalloc(&e->post_scratch, bytes);
    if (a != cudaSuccess) return set_error(e, FX_ERR_NOMEM, std::string("cudaMalloc(post-processing scratch): ") + cudaGetErrorString(a));
    e->post_cap = bytes;
    return FX_OK;
}

int post_column_stats(fx_engine* e, const float* x, long long n, int d, double* mean_dev, double* std_dev, double* var_dev,
                      fx_matrix_stats* out, cudaStream_t stream) {
    const int blocks = (int)((n + kStatRowsPerBlock - 1) / kStatRowsPerBlock);
    // scratch: partials [blocks][d] f64 | mean [d] | scalars [2] f64 | bad [2] u64
    const size_t part_b = sizeof(double) * (size_t)blocks * d, need = part_b + sizeof(double) * (d + 2) + 16;
    int rc = scratch_reserve(e, need);
    if (rc != FX_OK) return rc;
    double* part = static_cast<double*>(e->post_scratch);
    double* mean = mean_dev ? mean_dev : part + (size_t)blocks * d;
    double* scalars = part + (size_t)blocks * d + d;
    unsigned long long* bad = reinterpret_cast<unsigned long long*>(scalars + 2);
    FX_CUDA(e, cudaMemsetAsync(bad, 0, 16, stream));
    colsum_kernel<<<blocks, kStatThreads, 0, stream>>>(x, n, d, part, bad);
    FX_LAUNCH_CHECK(e, "colsum_kernel");
    colmean_kernel<<<(d + 255) / 256, 256, 0, stream>>>(part, blocks, d, n, mean);
    FX_LAUNCH_CHECK(e, "colmean_kernel");
    colsq_kernel<<<blocks, kStatThreads, 0, stream>>>(x, n, d, mean, part);
    FX_LAUNCH_CHECK(e, "colsq_kernel");
    colstd_kernel<<<1, 256, 0, stream>>>(part, blocks, d, n, mean, std_dev, var_dev, scalars);
    FX_LAUNCH_CHECK(e, "colstd_kernel");
    if (out) {
        double sc[2];
        unsigned long long b[2];
        FX_CUDA(e, cudaMemcpyAsync(sc, scalars, sizeof(sc), cudaMemcpyDeviceToHost, stream));
        FX_CUDA(e, cudaMemcpyAsync(b, bad, sizeof(b), cudaMemcpyDeviceToHost, stream));
        FX_CUDA(e, cudaStreamSynchronize(stream));
        out->nan_count = (int64_t)b[0];
        out->inf_count = (int64_t)b[1];
        out->mean_abs_mean = sc[0];
        out->mean_std = sc[1];
    }
    return FX_OK;
}

int post_standardize(fx_engine* e, const float* x, long long n, int d, const double* mean_dev, const double* scale_dev, float* out,
                     cudaStream_t stream) {
    const long long total = n * d;
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)e->sm_count * 32);
    standardize_kernel<<<blocks, 256, 0, stream>>>(x, total, d, mean_dev, scale_dev, out);
    FX_LAUNCH_CHECK(e, "standardize_kernel");
    return FX_OK;
}

int post_neighbor_probe(fx_engine* e, const float* x, long long n, int d, const int64_t* qidx_host, int q, int64_t* nbr_host,
                        float* sim_host, cudaStream_t stream) {
    const int blocks = (int)((n + kStatRowsPerBlock - 1) / kStatRowsPerBlock);
    // scratch: inv [n] f32 | qidx [q] i64 | best_sim [blocks][q] f32 | best_row [blocks][q] i64
    const size_t inv_b = (sizeof(float) * (size_t)n + 15) & ~(size_t)15, q_b = sizeof(long long) * kProbeMaxQ;
    const size_t sim_b = (sizeof(float) * (size_t)blocks * q + 15) & ~(size_t)15, row_b = sizeof(long long) * (size_t)blocks * q;
    int rc = scratch_reserve(e, inv_b + q_b + sim_b + row_b);
    if (rc != FX_OK) return rc;
    uint8_t* base = static_cast<uint8_t*>(e->post_scratch);
    float* inv = reinterpret_cast<float*>(base);
    long long* qidx = reinterpret_cast<long long*>(base + inv_b);
    float* best_sim = reinterpret_cast<float*>(base + inv_b + q_b);
    long long* best_row = reinterpret_cast<long long*>(base + inv_b + q_b + sim_b);
    FX_CUDA(e, cudaMemcpyAsync(qidx, qidx_host, sizeof(long long) * q, cudaMemcpyHostToDevice, stream));
    rownorm_kernel<<<(int)((n + 7) / 8), 256, 0, stream>>>(x, n, d, inv);
    FX_LAUNCH_CHECK(e, "rownorm_kernel");
    const size_t smem = sizeof(float) * (size_t)q * d;
    static bool attr_done[16] = {};
    if (!attr_done[e->device & 15]) {
        FX_CUDA(e, cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done[e->device & 15] = true;
    }
    probe_kernel<<<blocks, 256, smem, stream>>>(x, n, d, inv, qidx, q, best_sim, best_row);
    FX_LAUNCH_CHECK(e, "probe_kernel");
    std::vector<float> hs((size_t)blocks * q);
    std::vector<long long> hr((size_t)blocks * q);
    FX_CUDA(e, cudaMemcpyAsync(hs.data(), best_sim, sizeof(float) * hs.size(), cudaMemcpyDeviceToHost, stream));
    FX_CUDA(e, cudaMemcpyAsync(hr.data(), best_row, sizeof(long long) * hr.size(), cudaMemcpyDeviceToHost, stream));
    FX_CUDA(e, cudaStreamSynchronize(stream));
    for (int j = 0; j < q; ++j) {  // blocks own ascending row ranges: strict > keeps the first maximum (np.argmax)
        float b = -INFINITY;
        long long br = -1;
        for (int k = 0; k < blocks; ++k) {
            const long long r = hr[(size_t)k * q + j];
            if (r >= 0 && (br < 0 || hs[(size_t)k * q + j] > b)) {
                b = hs[(size_t)k * q + j];
                br = r;
            }
        }
        nbr_host[j] = br;
        sim_host[j] = b;
    }
    return FX_OK;
}

}  // namespace fx
