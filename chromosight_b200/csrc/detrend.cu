// K0a: distance-law detrending of a CSR contact map.
//   preprocessing.py:129-197  distance_law (smooth=False, fun=nanmean)
//   preprocessing.py:256-310  detrend
// Segmented sum / count by diagonal (col - row) with a per-block shared-memory
// histogram flushed by atomics, then an element-wise divide.
#include "common.cuh"

namespace cs {

constexpr int kLawSmemDiags = 4096;

__global__ void diag_accumulate(const int64_t *__restrict__ indptr,
                                const int32_t *__restrict__ indices,
                                const double *__restrict__ data, int n,
                                const uint8_t *__restrict__ detect, int n_diags, double *gsum,
                                unsigned long long *gcnt, int use_smem) {
    extern __shared__ unsigned char sm[];
    double *ssum = reinterpret_cast<double *>(sm);
    unsigned int *scnt = reinterpret_cast<unsigned int *>(ssum + (use_smem ? n_diags : 0));
    if (use_smem) {
        for (int i = threadIdx.x; i < n_diags; i += blockDim.x) {
            ssum[i] = 0.0;
            scnt[i] = 0u;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
        if (detect && !detect[r]) continue;  // pre:180-185: both bins must be detectable
        const int64_t b = indptr[r], e = indptr[r + 1];
        for (int64_t k = b + lane; k < e; k += 32) {
            const int c = indices[k];
            const int d = c - r;
            if (d < 0 || d >= n_diags) continue;  // upper diagonals 0..max_dist only
            if (detect && !detect[c]) continue;
            const double v = data[k];
            if (!(v > 0.0)) continue;  // pre:187 keeps strictly positive pixels (drops NaN)
            if (use_smem) {
                atomicAdd(&ssum[d], v);
                atomicAdd(&scnt[d], 1u);
            } else {
                atomicAdd(&gsum[d], v);
                atomicAdd(&gcnt[d], 1ull);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_diags; i += blockDim.x) {
            if (scnt[i]) {
                atomicAdd(&gsum[i], ssum[i]);
                atomicAdd(&gcnt[i], (unsigned long long)scnt[i]);
            }
        }
    }
}

__global__ void law_finalize(const double *gsum, const unsigned long long *gcnt, int n_diags,
                             int n, double *law) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double v = 0.0;  // pre:289: NaN (empty diagonal) -> 0; beyond max_dist the law is 0
        if (i < n_diags && gcnt[i] > 0) v = gsum[i] / (double)gcnt[i];
        law[i] = v;
    }
}

__global__ void detrend_rows(const int64_t *__restrict__ indptr,
                             const int32_t *__restrict__ indices, const double *__restrict__ in,
                             double *out, int n_rows, const double *__restrict__ law, int n_law,
                             double max_val) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += gridDim.x * wpb) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        for (int64_t k = b + lane; k < e; k += 32) {
            int d = indices[k] - r;
            d = d < 0 ? -d : d;
            const double y = d < n_law ? law[d] : 0.0;
            double v = in[k] / y;  // pre:304 (x/0 -> inf, 0/0 -> NaN as in numpy)
            if (max_val >= 0.0 && v >= max_val) v = 1.0;  // pre:308-309
            out[k] = v;
        }
    }
}

}  // namespace cs

using namespace cs;

extern "C" int cs_distance_law(const int64_t *d_indptr, const int32_t *d_indices,
                               const double *d_data, int32_t n, const uint8_t *d_detect,
                               int32_t n_diags, double *d_sum, int64_t *d_cnt, double *d_law,
                               void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(d_indptr && d_indices && d_data && d_sum && d_cnt && d_law && n > 0 && n_diags > 0,
               "cs_distance_law: bad arguments");
    if (n_diags > n) n_diags = n;
    CS_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)n_diags * sizeof(double), st));
    CS_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)n_diags * sizeof(int64_t), st));
    const int use_smem = n_diags <= kLawSmemDiags;
    const size_t smem = use_smem ? (size_t)n_diags * (sizeof(double) + sizeof(unsigned int)) : 0;
    int grid = (n + 7) / 8;
    if (grid > 148 * 4) grid = 148 * 4;
    if (smem > 48 * 1024)
        CS_CUDA(cudaFuncSetAttribute(diag_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    diag_accumulate<<<grid, 256, smem, st>>>(d_indptr, d_indices, d_data, n, d_detect, n_diags,
                                             d_sum, (unsigned long long *)d_cnt, use_smem);
    CS_LAUNCHED();
    law_finalize<<<(n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256, 256, 0, st>>>(
        d_sum, (const unsigned long long *)d_cnt, n_diags, n, d_law);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_detrend_apply(const int64_t *d_indptr, const int32_t *d_indices,
                                const double *d_data_in, double *d_data_out, int32_t n_rows,
                                const double *d_law, int32_t n_law, double max_val, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(d_indptr && d_indices && d_data_in && d_data_out && d_law && n_rows > 0,
               "cs_detrend_apply: bad arguments");
    int grid = (n_rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    detrend_rows<<<grid, 256, 0, st>>>(d_indptr, d_indices, d_data_in, d_data_out, n_rows, d_law,
                                       n_law, max_val);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
