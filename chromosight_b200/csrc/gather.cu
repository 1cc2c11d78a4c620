// K3: per-coordinate window gather and validation (detection.py:18-155,
// validate_patterns) and score / p-value lookup at given coordinates
// (detection.py:134, 337-339) -- the per-pattern Python loop of the reference
// (~200 us per coordinate) as one warp per coordinate.
//
// The reference validates on the zero-padded matrix (det:291-298) whose k
// sub-diagonals were set to NaN (det:300-310); here the padding and the NaN
// diagonals are arithmetic on the coordinates of the unpadded CSR matrix.
#include "common.cuh"

namespace cs {

struct GatherParams {
    int rows, cols;          // unpadded matrix
    int wh, ww;              // window (kernel) shape
    int pad_r, pad_c;        // zero padding added above / left of the matrix
    int det_r, det_c;        // shift applied to the detectable-bin ids
    int big_k;               // NaN sub-diagonals 1..big_k of the padded matrix (0 = none)
    double zero_tol, missing_tol;
};

__device__ __forceinline__ double csr_value(const int64_t *__restrict__ indptr,
                                            const int32_t *__restrict__ indices,
                                            const double *__restrict__ data, int y, int x) {
    int64_t lo = indptr[y], hi = indptr[y + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int c = indices[mid];
        if (c < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < indptr[y + 1] && indices[lo] == x) ? data[lo] : 0.0;
}

// value of padded pixel (Y, X) as validate_patterns sees it
__device__ __forceinline__ double window_pixel(const GatherParams &G,
                                               const int64_t *__restrict__ indptr,
                                               const int32_t *__restrict__ indices,
                                               const double *__restrict__ data,
                                               const uint8_t *__restrict__ vrow,
                                               const uint8_t *__restrict__ vcol, int Y, int X) {
    // missing bins: ids not in the (shifted) detectable lists, padding included
    const int ry = Y - G.det_r, cx = X - G.det_c;
    const bool mrow = ry < 0 || ry >= G.rows || (vrow && !vrow[ry]);
    const bool mcol = cx < 0 || cx >= G.cols || (vcol && !vcol[cx]);
    if (mrow || mcol) return __longlong_as_double(0x7ff8000000000000ll);
    if (G.big_k > 0 && Y - X >= 1 && Y - X <= G.big_k)
        return __longlong_as_double(0x7ff8000000000000ll);
    const int y = Y - G.pad_r, x = X - G.pad_c;
    if (y < 0 || y >= G.rows || x < 0 || x >= G.cols) return 0.0;
    return csr_value(indptr, indices, data, y, x);
}

// one warp per coordinate; coords are PADDED coordinates (row, col)
__global__ void gather_windows(GatherParams G, const int64_t *__restrict__ indptr,
                               const int32_t *__restrict__ indices,
                               const double *__restrict__ data,
                               const uint8_t *__restrict__ vrow,
                               const uint8_t *__restrict__ vcol,
                               const int32_t *__restrict__ coords, long long P,
                               double *__restrict__ windows, uint8_t *__restrict__ valid) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int npx = G.wh * G.ww;
    const int H = G.rows + 2 * G.pad_r, W = G.cols + 2 * G.pad_c;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (long long p = warp; p < P; p += nwarps) {
        const int p1 = coords[2 * p], p2 = coords[2 * p + 1];
        const int half_h = G.wh / 2 + 1, half_w = G.ww / 2 + 1;
        const int high = p1 - half_h + 1, low = p1 + half_h;
        const int left = p2 - half_w + 1, right = p2 + half_w;
        double *out = windows + p * npx;
        bool ok = high >= 0 && low < H && left >= 0 && right < W;  // det:97-102
        if (ok) {
            int nzero = 0, nmiss = 0;
            for (int idx = lane; idx < npx; idx += 32) {
                const int i = idx / G.ww, j = idx - i * G.ww;
                const double v = window_pixel(G, indptr, indices, data, vrow, vcol, high + i, left + j);
                out[idx] = v;
                nzero += (v == 0.0);
                nmiss += !(fabs(v) <= 1.79769313486231570815e308);  // ~isfinite
            }
            for (int o = 16; o > 0; o >>= 1) {
                nzero += __shfl_xor_sync(0xffffffffu, nzero, o);
                nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
            }
            const double prop_undetected = (double)nmiss / (double)npx;
            const double prop_zero = (double)nzero / (double)(npx - nmiss);  // 0/0 -> NaN -> invalid
            ok = (prop_undetected < G.missing_tol) && (prop_zero < G.zero_tol);
        }
        if (!ok)
            for (int idx = lane; idx < npx; idx += 32) out[idx] = qnan;
        if (lane == 0) valid[p] = ok ? 1 : 0;
    }
}

struct LookupView {
    int rows, cols, dlo, dhi, pitch, dense, dmin, dmax;
};

// scores (and log10 p-values) of the score image at UNPADDED coordinates; pixels outside
// the image or the stored band read as 0 (absent from the sparse map).  dmin..dmax applies to
// the score only: the reference looks the score up in the diag-trimmed map (det:270, 134) but
// the p-value in the untrimmed one (det:337-339)
__global__ void lookup_scores(LookupView S, const float *__restrict__ sc,
                              NmissPlane nobs,
                              const int32_t *__restrict__ coords, long long P,
                              double *__restrict__ score, double *__restrict__ log10p) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P;
         p += (long long)gridDim.x * blockDim.x) {
        const int y = coords[2 * p], x = coords[2 * p + 1];
        double v = 0.0, lp = 0.0;
        const int d = x - y;
        bool in = y >= 0 && y < S.rows && x >= 0 && x < S.cols;
        if (in && !S.dense) in = d >= S.dlo && d <= S.dhi;
        if (in) {
            const long long i = (long long)y * S.pitch + (x - (S.dense ? 0 : S.dlo));
            const float vf = sc[i];
            if (d >= S.dmin && d <= S.dmax) v = (double)vf;
            if (vf != 0.f) lp = log10_pval(vf, nobs_at(nobs, i));
        }
        score[p] = v;
        if (log10p) log10p[p] = lp;
    }
}

}  // namespace cs

using namespace cs;

extern "C" int cs_window_gather(const cs_gather_args *a, const int32_t *d_coords, int64_t n_coords,
                                double *d_windows, uint8_t *d_valid, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(a && a->d_indptr && a->d_indices && a->d_data && d_coords && d_windows && d_valid,
               "cs_window_gather: null argument");
    CS_REQUIRE(a->rows > 0 && a->cols > 0 && a->win_h > 0 && a->win_w > 0,
               "cs_window_gather: bad shapes");
    if (n_coords <= 0) return CS_OK;
    GatherParams G;
    G.rows = a->rows;
    G.cols = a->cols;
    G.wh = a->win_h;
    G.ww = a->win_w;
    G.pad_r = a->pad_rows;
    G.pad_c = a->pad_cols;
    G.det_r = a->det_shift_row;
    G.det_c = a->det_shift_col;
    G.big_k = a->nan_subdiag;
    G.zero_tol = a->zero_tol;
    G.missing_tol = a->missing_tol;
    const int threads = 256;
    long long blocks = (n_coords * 32 + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    gather_windows<<<(int)blocks, threads, 0, st>>>(G, a->d_indptr, a->d_indices, a->d_data,
                                                   a->d_valid_row, a->d_valid_col, d_coords,
                                                   (long long)n_coords, d_windows, d_valid);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_scores_lookup(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                                int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                                const int32_t *d_coords, int64_t n_coords, double *d_score,
                                double *d_log10p, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(Lo && d_out && d_coords && d_score, "cs_scores_lookup: null argument");
    if (n_coords <= 0) return CS_OK;
    LookupView S;
    S.rows = Lo->rows;
    S.cols = Lo->cols;
    S.dlo = Lo->dlo;
    S.dhi = Lo->dhi;
    S.pitch = Lo->pitch;
    S.dense = Lo->dense;
    S.dmin = dmin;
    S.dmax = dmax;
    long long blocks = (n_coords + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    lookup_scores<<<(int)blocks, 256, 0, st>>>(S, d_out, NmissPlane{d_nmiss, nmiss_bytes, n_window}, d_coords,
                                               (long long)n_coords, d_score, d_log10p);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
