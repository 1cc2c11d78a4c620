// K0b: framing + densification of a CSR signal (and its missing mask) into the
// float32 band / dense image consumed by the Pearson kernel.
//   detection.py:979-991      zero frame of (mk-1, nk-1) pixels around the signal
//   preprocessing.py:404-498  frame_missing_mask (margins + sub-diagonals)
//   preprocessing.py:501-532  check_missing_mask (signal must be 0 under the mask)
// A pixel mask (mask_mode 1) is stored as NaN sentinels in the image itself; the geometric
// mask of make_missing_mask / frame_missing_mask (mask_mode 2) as a constant fill value under
// the missing rows, columns and strip (the Pearson kernel gets the geometry itself).  Either
// way the Pearson kernel reads ONE array (4 B / pixel).
#include <stdarg.h>
#include "common.cuh"

namespace cs {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct FillParams {
    int rows, cols, dlo, dhi, pitch, dense;
    int row_off, col_off;
};

__device__ __forceinline__ bool stored(const FillParams &F, int Y, int X) {
    if (Y < 0 || Y >= F.rows || X < 0 || X >= F.cols) return false;
    if (F.dense) return true;
    const int d = X - Y;
    return d >= F.dlo && d <= F.dhi;
}
__device__ __forceinline__ long long at(const FillParams &F, int Y, int X) {
    return (long long)Y * F.pitch + (X - (F.dense ? 0 : F.dlo));
}

// one warp per CSR row
__global__ void scatter_signal(FillParams F, const int64_t *__restrict__ indptr,
                               const int32_t *__restrict__ indices,
                               const double *__restrict__ data, int r0, int r1, int ignore_below,
                               float *img, int *err) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = r0 + blockIdx.x * wpb + (threadIdx.x >> 5); r < r1; r += gridDim.x * wpb) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        const int Y = r + F.row_off;
        for (int64_t k = b + lane; k < e; k += 32) {
            const int X = indices[k] + F.col_off;
            const double v = data[k];
            if (stored(F, Y, X)) {
                // duplicates of a non-canonical CSR would need atomics; scipy sums them
                // before we get here (host side calls sum_duplicates)
                img[at(F, Y, X)] = (float)v;
            } else if (v != 0.0 && !(ignore_below && !F.dense && X - Y < F.dlo)) {
                // (sym_upper: pixels below the stored band cannot reach a kept score, det:1098)
                atomicAdd(err + 1, 1);
            }
        }
    }
}

// user mask pixels -> NaN; counts signal pixels that are non-zero under the mask
__global__ void scatter_mask(FillParams F, const int64_t *__restrict__ indptr,
                             const int32_t *__restrict__ indices, int r0, int r1, int trim_lo,
                             int trim_hi, int do_trim, float *img, int *err) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = r0 + blockIdx.x * wpb + (threadIdx.x >> 5); r < r1; r += gridDim.x * wpb) {
        const int64_t b = indptr[r], e = indptr[r + 1];
        const int Y = r + F.row_off;
        for (int64_t k = b + lane; k < e; k += 32) {
            const int c = indices[k];
            const int X = c + F.col_off;
            if (!stored(F, Y, X)) continue;
            const long long i = at(F, Y, X);
            const float v = img[i];
            if (v != 0.f && v == v) atomicAdd(err, 1);  // pre:516-523
            // pre:452-454: the mask is diag-trimmed before framing
            if (do_trim && (c - r < trim_lo || c - r > trim_hi)) continue;
            img[i] = quiet_nan_f();
        }
    }
}

// NaN over a rectangle of the framed image (margins of frame_missing_mask)
__global__ void nan_rect(FillParams F, int y0, int y1, int x0, int x1, float *img) {
    const long long w = x1 - x0;
    const long long n = (long long)(y1 - y0) * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int Y = y0 + (int)(i / w), X = x0 + (int)(i % w);
        if (stored(F, Y, X)) img[at(F, Y, X)] = quiet_nan_f();
    }
}

// pre:483-497: big_k diagonals below the main diagonal of the framed image
__global__ void nan_subdiag(FillParams F, int big_k, int Y0, int Y1, float *img, int *err) {
    const long long n = (long long)(Y1 - Y0) * big_k;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int Y = Y0 + (int)(i / big_k), X = Y - 1 - (int)(i % big_k);
        if (!stored(F, Y, X)) continue;
        const long long k = at(F, Y, X);
        const float v = img[k];
        if (v != 0.f && v == v) atomicAdd(err, 1);
        img[k] = quiet_nan_f();
    }
}

// geometric mask (mask_mode 2): one warp per image row writes the fill value under the missing
// strip, the row's missing pixels (the whole flagged band of a missing row, the missing columns
// of any other row) and the frame's margins; signal found under the mask is counted (pre:516-523)
struct GeoFill {
    const uint32_t *rbits, *cbits;
    int mlo, mhi, my0, my1, mx0, mx1, margin_mode, top_x1, right_y0, sdlo, sdhi;
    float fill;
};

__device__ __forceinline__ void fill_span(const FillParams &F, int Y, int xa, int xb, float fill,
                                          bool check, int lane, float *img, int *err) {
    // columns [xa, xb] of row Y that are stored
    if (xa < 0) xa = 0;
    if (xb > F.cols - 1) xb = F.cols - 1;
    if (!F.dense) {
        xa = max(xa, Y + F.dlo);
        xb = min(xb, Y + F.dhi);
    }
    for (int X = xa + lane; X <= xb; X += 32) {
        const long long i = at(F, Y, X);
        if (check) {
            const float v = img[i];
            if (v != 0.f) atomicAdd(err, 1);
        }
        img[i] = fill;
    }
}

__global__ void geo_fill(FillParams F, GeoFill Gm, int Y0, int Y1, float *img, int *err) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int Y = Y0 + blockIdx.x * wpb + (threadIdx.x >> 5); Y < Y1; Y += gridDim.x * wpb) {
        const bool inrow = Y >= Gm.my0 && Y < Gm.my1;
        // strip first (the margins written below may overlap it and are not checked)
        if (Gm.sdhi >= Gm.sdlo) fill_span(F, Y, Y + Gm.sdlo, Y + Gm.sdhi, Gm.fill, true, lane, img, err);
        if (inrow) {
            long long lo = (long long)Y + Gm.mlo, hi = (long long)Y + Gm.mhi;
            int xlo = (int)(lo < Gm.mx0 ? Gm.mx0 : lo), xhi = (int)(hi > Gm.mx1 - 1 ? Gm.mx1 - 1 : hi);
            if ((Gm.rbits[Y >> 5] >> (Y & 31)) & 1u) {
                fill_span(F, Y, xlo, xhi, Gm.fill, true, lane, img, err);
            } else {
                if (!F.dense) {
                    xlo = max(xlo, Y + F.dlo);
                    xhi = min(xhi, Y + F.dhi);
                }
                for (int X = xlo + lane; X <= xhi; X += 32)
                    if ((Gm.cbits[X >> 5] >> (X & 31)) & 1u) {
                        const long long i = at(F, Y, X);
                        const float v = img[i];
                        if (v != 0.f) atomicAdd(err, 1);
                        img[i] = Gm.fill;
                    }
            }
        }
        if (Gm.margin_mode == 2) {
            if (!inrow) {
                fill_span(F, Y, 0, F.cols - 1, Gm.fill, false, lane, img, err);
            } else {
                fill_span(F, Y, 0, Gm.mx0 - 1, Gm.fill, false, lane, img, err);
                fill_span(F, Y, Gm.mx1, F.cols - 1, Gm.fill, false, lane, img, err);
            }
        } else if (Gm.margin_mode == 1) {
            if (Y < Gm.my0) fill_span(F, Y, 0, Gm.top_x1 - 1, Gm.fill, false, lane, img, err);
            if (Y >= Gm.right_y0) fill_span(F, Y, Gm.mx1, F.cols - 1, Gm.fill, false, lane, img, err);
        }
    }
}

}  // namespace cs

using namespace cs;

extern "C" int cs_version(void) { return CS_ABI_VERSION; }
extern "C" const char *cs_last_error(void) { return cs::g_err; }
extern "C" int64_t cs_launch_count(void) { return (int64_t)cs::g_launches.load(); }

extern "C" int cs_layout_band_padded(cs_layout *L, int32_t rows, int32_t cols, int32_t dlo, int32_t dhi,
                                     int32_t gap) {
    CS_REQUIRE(L && rows > 0 && cols > 0 && dhi >= dlo && gap >= 0, "cs_layout_band: bad arguments");
    if (dlo < -(rows - 1)) dlo = -(rows - 1);
    if (dhi > cols - 1) dhi = cols - 1;
    CS_REQUIRE(dhi >= dlo, "cs_layout_band: empty band");
    L->rows = rows;
    L->cols = cols;
    L->dlo = dlo;
    L->dhi = dhi;
    L->dense = 0;
    // a row holds pitch + 1 elements: the dhi - dlo + 1 stored diagonals, then >= gap zeros
    int pitch = round_up(dhi - dlo + gap, 4);
    if (pitch < 4) pitch = 4;
    L->pitch = pitch;
    int64_t n = (int64_t)(rows - 1) * pitch + ((int64_t)cols - dlo);
    int64_t n2 = (int64_t)rows * (pitch + 1);
    if (n2 > n) n = n2;
    L->n_elems = (n + 3 + 64) / 4 * 4;
    return CS_OK;
}

extern "C" int cs_layout_band(cs_layout *L, int32_t rows, int32_t cols, int32_t dlo, int32_t dhi) {
    return cs_layout_band_padded(L, rows, cols, dlo, dhi, 0);
}

extern "C" int cs_layout_dense(cs_layout *L, int32_t rows, int32_t cols) {
    CS_REQUIRE(L && rows > 0 && cols > 0, "cs_layout_dense: bad arguments");
    L->rows = rows;
    L->cols = cols;
    L->dlo = 0;
    L->dhi = 0;
    L->dense = 1;
    L->pitch = round_up(cols, 4);
    L->n_elems = (int64_t)rows * L->pitch + 64;
    return CS_OK;
}

namespace cs {

static FillParams make_fill(const cs_layout *L, int row_off, int col_off) {
    FillParams F;
    F.rows = L->rows;
    F.cols = L->cols;
    F.dlo = L->dlo;
    F.dhi = L->dhi;
    F.pitch = L->pitch;
    F.dense = L->dense;
    F.row_off = row_off;
    F.col_off = col_off;
    return F;
}

// Zero the image and the error counters, write the margin rectangles of
// frame_missing_mask (pre:461-480); they never overlap the signal.
int fill_begin(const cs_layout *L, float *d_img, int32_t n_rows, int32_t n_cols, int32_t mask_mode,
               int32_t sym_upper, int32_t max_dist, int32_t frame_mk, int32_t frame_nk,
               int32_t *d_err, cudaStream_t st) {
    FillParams F = make_fill(L, 0, 0);
    CS_CUDA(cudaMemsetAsync(d_img, 0, (size_t)L->n_elems * sizeof(float), st));
    CS_CUDA(cudaMemsetAsync(d_err, 0, 2 * sizeof(int32_t), st));
    if (mask_mode == 1 && frame_mk > 0) {
        const int threads = 256;
        const bool banded = sym_upper && max_dist >= 0;
        const int H = L->rows, W = L->cols, mk = frame_mk, nk = frame_nk;
        const int ns = n_cols;
        struct R {
            int y0, y1, x0, x1;
        } rc[4];
        int nr = 0;
        if (banded) {
            const int max_m = max_dist + mk, max_n = max_dist + nk;
            const int tn = max_n < ns ? max_n : ns;
            rc[nr++] = {0, mk - 1, nk - 1, nk - 1 + tn};                                 // pre:461-463
            rc[nr++] = {H - (max_m + 1) > 0 ? H - (max_m + 1) : 0, H, W - (nk - 1), W};  // pre:475
            rc[nr++] = {0, mk - 1, 0, nk - 1};                                           // pre:477
        } else {
            rc[nr++] = {0, mk - 1, 0, W};
            rc[nr++] = {H - (mk - 1), H, 0, W};
            rc[nr++] = {0, H, 0, nk - 1};
            rc[nr++] = {0, H, W - (nk - 1), W};
        }
        for (int i = 0; i < nr; ++i) {
            const long long n = (long long)(rc[i].y1 - rc[i].y0) * (rc[i].x1 - rc[i].x0);
            if (n <= 0) continue;
            int g = (int)((n + threads - 1) / threads);
            if (g > 148 * 8) g = 148 * 8;
            nan_rect<<<g, threads, 0, st>>>(F, rc[i].y0, rc[i].y1, rc[i].x0, rc[i].x1, d_img);
            CS_LAUNCHED();
        }
    }
    CS_CUDA(cudaGetLastError());
    (void)n_rows;
    return CS_OK;
}

// Signal rows [r0, r1): scatter the signal, then the user mask, then the sub-diagonal mask
// of the image rows they occupy (the first / last call also covers the top / bottom frame).
int fill_rows(const cs_layout *L, float *d_img, const int64_t *d_sig_indptr,
              const int32_t *d_sig_indices, const double *d_sig_data, int32_t n_rows,
              int32_t r0, int32_t r1, int32_t row_off, int32_t col_off, int32_t mask_mode,
              const int64_t *d_mask_indptr, const int32_t *d_mask_indices, const cs_geo_mask *geo,
              int32_t sym_upper, int32_t max_dist, int32_t frame_mk, int32_t frame_nk,
              int32_t *d_err, cudaStream_t st) {
    if (r1 <= r0) return CS_OK;
    FillParams F = make_fill(L, row_off, col_off);
    const int threads = 256;
    const int wpb = threads / 32;
    int grid = (r1 - r0 + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    if (d_sig_indptr) {
        scatter_signal<<<grid, threads, 0, st>>>(F, d_sig_indptr, d_sig_indices, d_sig_data, r0, r1,
                                                 sym_upper ? 1 : 0, d_img, d_err);
        CS_LAUNCHED();
    }
    if (mask_mode == 1) {
        const bool framed = frame_mk > 0;
        const bool banded = sym_upper && max_dist >= 0;
        const int big_k = frame_mk > frame_nk ? frame_mk : frame_nk;
        scatter_mask<<<grid, threads, 0, st>>>(F, d_mask_indptr, d_mask_indices, r0, r1, 0,
                                               max_dist + big_k, (framed && banded) ? 1 : 0, d_img,
                                               d_err);
        CS_LAUNCHED();
        if (framed && sym_upper) {
            const int Y0 = r0 == 0 ? 0 : r0 + row_off;
            const int Y1 = r1 == n_rows ? L->rows : r1 + row_off;
            const long long n = (long long)(Y1 - Y0) * big_k;
            int g = (int)((n + threads - 1) / threads);
            if (g > 148 * 16) g = 148 * 16;
            if (g > 0) {
                nan_subdiag<<<g, threads, 0, st>>>(F, big_k, Y0, Y1, d_img, d_err);
                CS_LAUNCHED();
            }
        }
    }
    if (mask_mode == 2) {
        GeoFill Gm;
        Gm.rbits = (const uint32_t *)geo->d_row_bits;
        Gm.cbits = (const uint32_t *)geo->d_col_bits;
        Gm.mlo = geo->mask_dlo, Gm.mhi = geo->mask_dhi;
        Gm.my0 = geo->mat_y0, Gm.my1 = geo->mat_y1, Gm.mx0 = geo->mat_x0, Gm.mx1 = geo->mat_x1;
        Gm.margin_mode = geo->margin_mode;
        Gm.top_x1 = geo->top_x1, Gm.right_y0 = geo->right_y0;
        Gm.sdlo = geo->strip_dlo, Gm.sdhi = geo->strip_dhi;
        Gm.fill = geo->fill_value;
        // image rows of these signal rows; the first / last call also covers the frame
        const int Y0 = r0 == 0 ? 0 : r0 + row_off;
        const int Y1 = r1 == n_rows ? L->rows : r1 + row_off;
        int g = (Y1 - Y0 + wpb - 1) / wpb;
        if (g > 148 * 64) g = 148 * 64;  // (a row is a few dependent accesses: latency hides in warps)
        if (g > 0) {
            geo_fill<<<g, threads, 0, st>>>(F, Gm, Y0, Y1, d_img, d_err);
            CS_LAUNCHED();
        }
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // namespace cs

extern "C" int cs_image_fill_f32(const cs_layout *L, float *d_img, const int64_t *d_sig_indptr,
                                 const int32_t *d_sig_indices, const double *d_sig_data,
                                 int32_t n_rows, int32_t n_cols, int32_t row_off, int32_t col_off,
                                 int32_t mask_mode, const int64_t *d_mask_indptr,
                                 const int32_t *d_mask_indices, const cs_geo_mask *geo,
                                 int32_t sym_upper, int32_t max_dist, int32_t frame_mk,
                                 int32_t frame_nk, int32_t *d_err, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(L && d_img && d_err, "cs_image_fill_f32: null argument");
    CS_REQUIRE(n_rows + row_off <= L->rows && n_cols + col_off <= L->cols,
               "signal does not fit in the image");
    if (mask_mode == 1) CS_REQUIRE(d_mask_indptr && d_mask_indices, "mask arrays missing");
    if (mask_mode == 2)
        CS_REQUIRE(geo && geo->d_row_bits && geo->d_col_bits, "geometric mask: bit vectors missing");
    int rc = fill_begin(L, d_img, n_rows, n_cols, mask_mode, sym_upper, max_dist, frame_mk, frame_nk,
                        d_err, st);
    if (rc) return rc;
    return fill_rows(L, d_img, d_sig_indptr, d_sig_indices, d_sig_data, n_rows, 0, n_rows, row_off,
                     col_off, mask_mode, d_mask_indptr, d_mask_indices, geo, sym_upper, max_dist,
                     frame_mk, frame_nk, d_err, st);
}
