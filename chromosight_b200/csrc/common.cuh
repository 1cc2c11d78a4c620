// Shared helpers of the chromosight_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/chromosight_b200.h"

namespace cs {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define CS_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                          cudaGetErrorString(_e));                                 \
            return CS_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define CS_REQUIRE(cond, ...)                                                      \
    do {                                                                           \
        if (!(cond)) {                                                             \
            cs::set_error(__VA_ARGS__);                                            \
            return CS_ERR_INVALID;                                                 \
        }                                                                          \
    } while (0)

// row-ranged building blocks shared by the one-shot and the pipelined host paths
int fill_begin(const cs_layout *L, float *d_img, int32_t n_rows, int32_t n_cols, int32_t mask_mode,
               int32_t sym_upper, int32_t max_dist, int32_t frame_mk, int32_t frame_nk,
               int32_t *d_err, cudaStream_t st);
int fill_rows(const cs_layout *L, float *d_img, const int64_t *d_sig_indptr,
              const int32_t *d_sig_indices, const double *d_sig_data, int32_t n_rows, int32_t r0,
              int32_t r1, int32_t row_off, int32_t col_off, int32_t mask_mode,
              const int64_t *d_mask_indptr, const int32_t *d_mask_indices, const cs_geo_mask *geo,
              int32_t sym_upper, int32_t max_dist, int32_t frame_mk, int32_t frame_nk,
              int32_t *d_err, cudaStream_t st);
// rows [r0, r1) of the score image: per-row counts + chunk-local scan (the slab total lands in
// d_total), then the offsets (+ base) added, then the CSR entries written
// (scratch_off: the range's part of the scan scratch, scan_slot(r0, k) for the k-th range)
int scores_count_rows(const cs_layout *Lo, const float *d_out, int32_t dmin, int32_t dmax,
                      int64_t *d_indptr, int32_t r0, int32_t r1, int64_t *d_total, cudaStream_t st,
                      int32_t scratch_off = 0);
int scores_finish_rows(const cs_layout *Lo, int64_t *d_indptr, int32_t r0, int32_t r1,
                       int64_t base, cudaStream_t st, int32_t scratch_off = 0);
int32_t scan_slot(int32_t r0, int32_t k);
int scores_emit_rows(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                     int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                     const int64_t *d_indptr,
                     int32_t r0, int32_t r1, int32_t *d_indices, double *d_data, double *d_log10p,
                     cudaStream_t st);

// the narrow wire format of the host path (scores.cu) and its expansion on the host
// (host_expand.cpp): rows [r0, r1) of a band result, entries indptr[r0]..indptr[r1]
int scores_emit_rows_narrow(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                            int32_t nmiss_bytes, int32_t n_window, const int64_t *d_indptr,
                            int32_t r0, int32_t r1, float *d_score, float *d_log10p,
                            uint8_t *d_off, cudaStream_t st, int64_t limit = INT64_MAX);
// score / log10p / off: wire arrays indexed like the CSR entries; indptr: final row pointers
// (host); writes data, logp (may be null), indices, indices2 (may be null) for the entries of
// rows [r0, r1); col = row + dlo + off.  Uses up to `threads` host threads.
void expand_rows(const float *score, const float *log10p, const uint8_t *off,
                 const int64_t *indptr, int32_t r0, int32_t r1, int32_t dlo, double *data,
                 double *logp, int32_t *indices, int32_t *indices2, int threads);
int expand_threads_default();
int pixels_lex_sorted(const int64_t *bin1, const int64_t *bin2, int64_t n, int threads);
int64_t pixels_inter_index(const int64_t *bin1, const int64_t *bin2, int64_t n, const int16_t *bin_chrom,
                           int32_t C, int64_t *order, int64_t *starts, int threads);
int64_t band_csr_from_pixels(const int64_t *bin1, const int64_t *bin2, const void *count, int count_dtype,
                             int64_t n_pix, const double *weight, int64_t s, int64_t e, int64_t max_diag,
                             int64_t *indptr, int32_t *indices, double *data, int threads);
}  // namespace cs
#include <functional>
namespace cs {
// fn(0..n-1) on one of the library's persistent host worker pools (host_expand.cpp; 0: widening
// of results, 1: staging copies of uploads); the caller runs share 0
void parallel_for(int n, const std::function<void(int)> &fn, int which);
int upload_threads_default();
void copy_stream(void *dst, const void *src, size_t n);

// exact float64 recomputation, from the CSR signal, of the scores at / near a Pearson threshold
// (pearson.cu): makes the candidate set of pick_foci independent of float32 rounding
struct RefineArgs {
    const cs_kernel_desc *K;
    const cs_pearson_opts *opts;  // mask mode, frame geometry, missing_tol, nobs_full, nmiss_bytes
    const int64_t *d_indptr;      // signal CSR (matrix coordinates)
    const int32_t *d_indices;
    const double *d_data;
    int rows, cols, pr, pc;       // matrix shape and its offset in the framed image
    const int64_t *d_m_indptr;    // pixel mask CSR (mask mode 1)
    const int32_t *d_m_indices;
    int trim_lo, trim_hi;         // diagonals of the pixel mask kept by the frame
    const cs_layout *Lo;          // score image (matrix coordinates)
    float *d_out;
    void *d_nmiss;
    double threshold;
    int dmin, dmax;
    int2 *d_list;                 // scratch: cap pixels + one counter
    long long cap;
    unsigned long long *d_count;
};
int exact_refine(const RefineArgs &R, cudaStream_t st, long long *n_refined);
int exact_refine_enqueue(const RefineArgs &R, cudaStream_t st);
// emit_candidates without the read-back of the count (scores.cu)
int scores_candidates_enqueue(const cs_layout *Lo, const float *d_out, const void *d_nmiss, int32_t nmiss_bytes,
                              int32_t n_window, int32_t dmin, int32_t dmax, float threshold, cs_candidate *d_cand,
                              int64_t cap, int64_t *d_count, cudaStream_t st);

// ... and from the refinement list (device-side length) instead of a pass over the band
int scores_candidates_from_list(const cs_layout *Lo, const float *d_out, const void *d_nmiss, int32_t nmiss_bytes,
                                int32_t n_window, int32_t dmin, int32_t dmax, float threshold, const int2 *d_list,
                                const unsigned long long *d_n, long long list_cap, cs_candidate *d_cand, int64_t cap,
                                int64_t *d_count, cudaStream_t st);

#define CS_LAUNCHED() (cs::g_launches.fetch_add(1, std::memory_order_relaxed))

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Pixel (Y, X) -> element offset in an image with layout L.
__host__ __device__ __forceinline__ long long img_index(int pitch, int dlo, int Y, int X) {
    return (long long)Y * pitch + (X - dlo);
}

__device__ __forceinline__ float quiet_nan_f() { return __int_as_float(0x7fc00000); }

// The missing-count plane written by the Pearson kernel (uint8 or uint16, NULL = none):
// number of observations of the window at element i (det:1110-1121).
struct NmissPlane {
    const void *p;
    int bytes, n_window;
};
__device__ __forceinline__ float nobs_at(const NmissPlane &M, long long i) {
    if (!M.p) return (float)M.n_window;
    const int nm = M.bytes == 2 ? (int)((const unsigned short *)M.p)[i]
                                : (int)((const unsigned char *)M.p)[i];
    return (float)(M.n_window - nm);
}

// stats.py:74-81: log10 of the two-sided p-value of a Pearson coefficient through Fisher's z,
// log10(2 Phi(-|z|)) = log10(erfc(|z| / sqrt 2)), z = atanh(r) sqrt(n - 3).
// Evaluated as log10(erfcx(a)) - a^2 log10(e) (a = |z| / sqrt 2), which never underflows, in
// float32: the relative error (~3e-7) is far below what the 1e-5 tolerance on r itself does
// to log10 p.  scipy's ndtr underflows to exactly 0 once a^2 > log(DBL_MAX): -inf there.
__device__ __forceinline__ double log10_pval(float r, float n_obs) {
    const float z = fabsf(atanhf(r)) * sqrtf(n_obs - 3.f);
    if (z != z) return (double)z;
    const float a = z * 0.70710678118654752440f;
    const float a2 = a * a;
    if (a2 > 7.09782712893383996843e2f) return -INFINITY;
    return (double)(log10f(erfcxf(a)) - a2 * 0.43429448190325182765f);
}

}  // namespace cs
