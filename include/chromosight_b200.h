/*
 * chromosight_b200 -- C ABI of the B200 (sm_100a) hot path of koszullab/chromosight.
 *
 * The reference is pure Python (no FFI exists in it); each entry point below
 * names the reference function it stands in for (file:line relative to the
 * reference tree, v1.6.3 @ ecb32c5).  All sizes are in elements, all pointers
 * are plain C pointers; `d_` prefixed arguments are DEVICE pointers owned by
 * the caller, everything else is HOST memory.  Every function returns 0 on
 * success or a negative cs_status; cs_last_error() gives the message of the
 * last failure on the calling thread.  `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).
 */
#ifndef CHROMOSIGHT_B200_H
#define CHROMOSIGHT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_ABI_VERSION 3

typedef enum cs_status {
    CS_OK = 0,
    CS_ERR_INVALID = -1,      /* bad argument (the Python layer raises ValueError) */
    CS_ERR_CUDA = -2,         /* CUDA runtime / driver failure */
    CS_ERR_NOMEM = -3,
    CS_ERR_MASKED_SIGNAL = -4 /* signal non-zero under the missing mask (pre:501-532) */
} cs_status;

int cs_version(void);
const char *cs_last_error(void);
/* Number of kernels launched by this library since process start (bench.py's gpu_launches). */
int64_t cs_launch_count(void);

/* ------------------------------------------------------------------------
 * Band layout.  A (framed) image of rows x cols pixels of which only the
 * diagonals dlo <= X - Y <= dhi are stored.  Pixel (Y, X) lives at element
 *      Y * pitch + (X - dlo)          with pitch % 4 == 0,
 * i.e. the classic skewed band of pitch + 1 diagonals per row, addressed so
 * that a TMA tensor map of row stride `pitch` presents it as an ordinary
 * matrix.  A dense image is the same formula with dlo = 0 and
 * pitch = roundup4(cols); set dense = 1.
 * ------------------------------------------------------------------------ */
typedef struct cs_layout {
    int32_t rows, cols;
    int32_t dlo, dhi; /* stored diagonals (ignored when dense) */
    int32_t pitch;
    int32_t dense;
    int64_t n_elems; /* allocation size in elements (floats) */
} cs_layout;

int cs_layout_band(cs_layout *L, int32_t rows, int32_t cols, int32_t dlo, int32_t dhi);
/* The same band with at least `gap` zero elements between the stored diagonals of consecutive
 * rows (pitch = roundup4(dhi - dlo + gap)).  A tile box of the Pearson kernel that sticks out
 * of the band by at most `gap` columns then reads zeros instead of the neighbouring rows'
 * pixels, and the kernel skips its alias fix-up pass; cs_pearson_plan reports the gap its
 * tiling needs.  The fill functions never write the gap (cs_image_fill_f32 zeroes it). */
int cs_layout_band_padded(cs_layout *L, int32_t rows, int32_t cols, int32_t dlo, int32_t dhi,
                          int32_t gap);
int cs_layout_dense(cs_layout *L, int32_t rows, int32_t cols);

/* ------------------------------------------------------------------------
 * Missing-pixel masks.  Two representations reach the device:
 *   mask_mode 1  an arbitrary pixel mask (CSR pattern).  Missing pixels are written
 *                into the image as NaN sentinels.
 *   mask_mode 2  the mask preprocessing.make_missing_mask (pre:535-633) builds from the
 *                detectable bins and frame_missing_mask (pre:404-498) frames: a pixel
 *                inside the matrix is missing iff its row or its column is a missing
 *                bin and it lies on diagonals mask_dlo..mask_dhi; plus the frame's
 *                margins and the strip of sub-diagonals (pre:483-497).  It is given by
 *                two bit vectors and a few integers (cs_geo_mask, IMAGE coordinates);
 *                missing pixels hold `fill_value` in the image (any constant is
 *                algebraically exact; one close to the signal's mean keeps the float32
 *                sums of heavily masked windows well conditioned; NaN for kernels wider
 *                than 31 columns, which take the one-warp-per-window kernel).
 * ------------------------------------------------------------------------ */
typedef struct cs_geo_mask {
    /* device bit vectors, bit i of word i / 32 = image row (column) i is a missing bin
     * inside the matrix; at least 2 zero words before word 0 and after the last word */
    const void *d_row_bits;
    const void *d_col_bits;
    int32_t mask_dlo, mask_dhi;   /* flagged diagonals X - Y (image coordinates) */
    int32_t mat_y0, mat_y1;       /* the matrix inside the frame: rows mat_y0 <= Y < mat_y1 */
    int32_t mat_x0, mat_x1;       /* and columns mat_x0 <= X < mat_x1 */
    int32_t margin_mode;          /* 0 no frame, 1 banded frame (pre:461-477), 2 all four margins */
    int32_t top_x1;               /* banded frame: top margin missing for X < top_x1 */
    int32_t right_y0;             /* banded frame: right margin missing for Y >= right_y0 */
    int32_t strip_dlo, strip_dhi; /* diagonals entirely missing (pre:483-497); empty if dhi < dlo */
    float fill_value;
} cs_geo_mask;

/* ------------------------------------------------------------------------
 * Framing + densification: detection.py:979-991 (zero frame around the
 * signal), preprocessing.py:404-498 (frame_missing_mask) and
 * preprocessing.py:501-532 (check_missing_mask).
 *
 * Scatters the CSR signal (n_rows x n_cols, float64 values, int32 column
 * indices, int64 row pointers) into the float32 image `d_img` (layout L) at
 * offset (row_off, col_off), writes the missing mask (mode 1: user mask pixels
 * from a CSR pattern, then the geometric frame of frame_missing_mask when
 * frame_mk > 0, as NaN sentinels; mode 2: `geo`), and counts signal pixels that
 * are non-zero under the mask into *d_err (int32[2], device).
 *   mask_mode: 0 no mask, 1 pixel mask (CSR), 2 geometric mask (geo).
 *   sym_upper / max_dist (-1 = None): as in frame_missing_mask.
 *   frame_mk, frame_nk: kernel shape when the image is framed (full=True), 0 otherwise.
 * ------------------------------------------------------------------------ */
int cs_image_fill_f32(const cs_layout *L, float *d_img,
                      const int64_t *d_sig_indptr, const int32_t *d_sig_indices,
                      const double *d_sig_data, int32_t n_rows, int32_t n_cols,
                      int32_t row_off, int32_t col_off,
                      int32_t mask_mode, const int64_t *d_mask_indptr,
                      const int32_t *d_mask_indices, const cs_geo_mask *geo,
                      int32_t sym_upper, int32_t max_dist,
                      int32_t frame_mk, int32_t frame_nk,
                      int32_t *d_err, void *stream);

/* ------------------------------------------------------------------------
 * Pearson map: detection.py:917-1131 (_normxcorr2_sparse; the no-mask branch
 * det:1002-1020 and the masked branch det:1021-1092) for every window centred
 * on an output pixel.  detection.py:595-723 (xcorr2) when raw_xcorr = 1.
 *
 * Output pixel set (image coordinates): oy0 <= Y < oy1, ox0 <= X < ox1 and
 * odlo <= X - Y <= odhi.  Results go to the float32 image `d_out` with layout
 * Lout, whose pixel (Y - out_row_shift, X - out_col_shift) receives the score of
 * window (Y, X) (the shifts undo the frame, det:1124-1129).  `d_nmiss` (uint8 or
 * uint16 per opts->nmiss_bytes, same layout, may be NULL) receives the number of
 * missing pixels of each window when nobs_full: the number of observations of
 * det:1110-1116 is kh*kw - nmiss.  The caller ZEROES the plane first: only windows with a
 * count are written.
 * ------------------------------------------------------------------------ */
typedef struct cs_kernel_desc {
    int32_t kh, kw;          /* kernel shape (mk, nk), both odd */
    const double *k_corr;    /* kh*kw, kernel correlated with the signal (truncated when tsvd) */
    const double *k_mask;    /* kh*kw, kernel correlated with the mask (det:1035-1040) */
    const double *k2_mask;   /* kh*kw, squared kernel correlated with the mask (det:1041-1046) */
    double k_sum, k2_sum;    /* sums of the ORIGINAL kernel and of its square (det:1024-1027) */
    double k_mean, k_std;    /* mean / std of the original kernel (det:1003-1004) */
} cs_kernel_desc;

typedef struct cs_pearson_opts {
    int32_t mask_mode;       /* 0 no mask, 1 NaN sentinels in the image, 2 geometric (geo) */
    double missing_tol;      /* det:1069-1072 */
    double xcorr_threshold;  /* 1e-4, det:595 */
    int32_t raw_xcorr;       /* 1: write thresholded raw cross-correlation instead of Pearson */
    int32_t nobs_full;       /* 1: n_obs = present pixels (full=True & mask), det:1110 */
    int32_t tile_rows;       /* 0 = auto */
    int32_t out_row_shift;   /* score of window (Y, X) is written to pixel           */
    int32_t out_col_shift;   /* (Y - out_row_shift, X - out_col_shift) of the output */
    int32_t nmiss_bytes;     /* element size of d_nmiss: 1 or 2 */
    cs_geo_mask geo;         /* mask_mode 2 */
} cs_pearson_opts;

int cs_pearson_f32(const cs_layout *Limg, const float *d_img,
                   const cs_kernel_desc *K, const cs_pearson_opts *opts,
                   int32_t oy0, int32_t oy1, int32_t ox0, int32_t ox1,
                   int32_t odlo, int32_t odhi,
                   const cs_layout *Lout, float *d_out, void *d_nmiss,
                   void *stream);
/* Height (output rows) of the tiles cs_pearson_f32 would use for this call.  A caller that
 * splits one region into row ranges (the slab pipeline of cs_normxcorr2_host) cuts at
 * multiples of it and passes it back as opts->tile_rows, so that every tile -- and with it
 * every float32 rounding -- is the same as in a single launch. */
int cs_pearson_tile_rows(const cs_layout *Limg, const cs_kernel_desc *K,
                         const cs_pearson_opts *opts,
                         int32_t oy0, int32_t oy1, int32_t ox0, int32_t ox1,
                         int32_t odlo, int32_t odhi, int32_t *tile_rows);
/* The same plan, plus the gap (cs_layout_band_padded) a banded image needs for the kernel to
 * skip its alias fix-up: 0 when the image is dense or the traversal does not follow the band
 * (the kernel then keeps the fix-up; results are identical either way). */
int cs_pearson_plan(const cs_layout *Limg, const cs_kernel_desc *K,
                    const cs_pearson_opts *opts,
                    int32_t oy0, int32_t oy1, int32_t ox0, int32_t ox1,
                    int32_t odlo, int32_t odhi, int32_t *tile_rows, int32_t *band_gap);

/* ------------------------------------------------------------------------
 * Score map -> CSR (what normxcorr2 returns, det:1098-1131): non-zero scores
 * as float64 CSR, plus log10 p-values (stats.py:43-81) at the same pattern.
 * Two calls: count (fills d_indptr[0..rows], returns nnz in *nnz_host after
 * synchronising the stream) then emit.
 * ------------------------------------------------------------------------ */
/* d_indptr must hold rows + 1 + cs_scan_scratch(rows) elements (scan workspace behind
 * the row pointers). */
int64_t cs_scan_scratch(int32_t rows);
int cs_scores_count(const cs_layout *Lout, const float *d_out,
                    int32_t dmin, int32_t dmax, /* keep only dmin <= col-row <= dmax */
                    int64_t *d_indptr, int64_t *nnz_host, void *stream);
/* d_nmiss / nmiss_bytes / n_window: the missing-count plane cs_pearson_f32 wrote (NULL: every
 * window has n_window = kh*kw observations). */
int cs_scores_emit(const cs_layout *Lout, const float *d_out, const void *d_nmiss,
                   int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                   const int64_t *d_indptr, int32_t *d_indices, double *d_data,
                   double *d_log10p /* may be NULL */, void *stream);

/* Candidate pixels (score >= threshold) as (row, col, score, log10p) records:
 * the thresholding of pick_foci (det:417-421) fused with the p-value lookup
 * (det:337-339). Returns the number found in *n_host (capped at cap). */
typedef struct cs_candidate {
    int32_t row, col;
    float score;
    float log10p;
} cs_candidate;
int cs_scores_candidates(const cs_layout *Lout, const float *d_out, const void *d_nmiss,
                         int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                         float threshold,
                         cs_candidate *d_cand, int64_t cap, int64_t *d_count,
                         int64_t *n_host, void *stream);

/* pick_foci on the device (det:387-456 with label_foci det:459-554 and filter_foci
 * det:557-592): pixels with score >= threshold on diagonals dmin..dmax form 4-connected foci;
 * every focus of at least min_size pixels yields one record.  Foci are numbered by their first
 * pixel in row-major order (records come unordered: sort by (first_row, first_col)); (row, col)
 * is the focus' highest score, the first in row-major order among equals, as np.argmax picks it.
 * d_work: cs_foci_work_bytes(L) bytes of device scratch. */
typedef struct cs_focus {
    int32_t first_row, first_col;
    int32_t row, col;
    float score;
    int32_t size;
} cs_focus;
int64_t cs_foci_work_bytes(const cs_layout *Lout);
int cs_scores_foci(const cs_layout *Lout, const float *d_out, int32_t dmin, int32_t dmax,
                   double threshold, int32_t min_size, void *d_work, cs_focus *d_foci,
                   int64_t cap, int64_t *d_count, int64_t *n_host, void *stream);

/* ------------------------------------------------------------------------
 * Window gather + validation: detection.py:18-155 (validate_patterns) on the
 * zero-padded, sub-diagonal-NaN matrix that pattern_detector builds
 * (det:291-310), one warp per coordinate.  `d_coords` holds (row, col) pairs in
 * PADDED coordinates.  Windows that fall outside the matrix or fail the
 * zero / missing tolerances are filled with NaN and flagged 0 in d_valid.
 * cs_scores_lookup: conv_mat[p1, p2] (det:134) and the log10 p-value (det:337-339)
 * at UNPADDED coordinates of the score image; absent pixels read as 0.
 * ------------------------------------------------------------------------ */
typedef struct cs_gather_args {
    int32_t rows, cols;            /* unpadded matrix */
    const int64_t *d_indptr;       /* device CSR of the unpadded matrix (sorted indices) */
    const int32_t *d_indices;
    const double *d_data;
    const uint8_t *d_valid_row;    /* device, 1 = detectable bin (NULL = all) */
    const uint8_t *d_valid_col;
    int32_t win_h, win_w;          /* kernel shape */
    int32_t pad_rows, pad_cols;    /* zero padding above / left of the matrix (det:291-294) */
    int32_t det_shift_row, det_shift_col; /* shift of the detectable ids (det:295-296) */
    int32_t nan_subdiag;           /* big_k of det:300-310, 0 for inter maps */
    double zero_tol, missing_tol;
} cs_gather_args;
int cs_window_gather(const cs_gather_args *a, const int32_t *d_coords, int64_t n_coords,
                     double *d_windows /* n_coords * win_h * win_w */, uint8_t *d_valid,
                     void *stream);
int cs_scores_lookup(const cs_layout *Lout, const float *d_out, const void *d_nmiss,
                     int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                     const int32_t *d_coords,
                     int64_t n_coords, double *d_score, double *d_log10p /* may be NULL */,
                     void *stream);

/* ------------------------------------------------------------------------
 * Distance-law detrending: preprocessing.py:129-197 (distance_law, smooth=False,
 * fun=nanmean) and preprocessing.py:256-310 (detrend).
 * cs_distance_law: per upper diagonal d <= max_dist, mean of the strictly
 * positive pixels whose bins are both detectable (d_detect: uint8[n], 1 = ok).
 * d_sum/d_cnt: double[n_diags] / int64[n_diags] workspaces (zeroed inside),
 * d_law: double[n] output (0 beyond n_diags and where no pixel qualifies,
 * as after pre:289).
 * cs_detrend_apply: data[i] /= law[|row-col|]; values >= max_val -> 1
 * (max_val < 0 disables), pre:303-309.
 * ------------------------------------------------------------------------ */
int cs_distance_law(const int64_t *d_indptr, const int32_t *d_indices, const double *d_data,
                    int32_t n, const uint8_t *d_detect, int32_t n_diags,
                    double *d_sum, int64_t *d_cnt, double *d_law, void *stream);
int cs_detrend_apply(const int64_t *d_indptr, const int32_t *d_indices, const double *d_data_in,
                     double *d_data_out, int32_t n_rows, const double *d_law, int32_t n_law,
                     double max_val, void *stream);

/* ------------------------------------------------------------------------
 * Narrow result format of the host path.  Band results of at most 256 diagonals leave the
 * device as float32 score, float32 log10 p and the column as a uint8 offset from the row's
 * first stored diagonal (9 B per stored score over PCIe instead of 20); this HOST function
 * widens rows [r0, r1) of it into the float64 / int32 arrays of the CSR matrices
 * (det:1098-1131) with up to `threads` threads (0 = default).  log10p / logp / indices2 may
 * be NULL.  Exposed for tests; cs_normxcorr2_host and cs_session_download call it.
 * ------------------------------------------------------------------------ */
int cs_expand_rows(const float *score, const float *log10p, const uint8_t *off,
                   const int64_t *indptr, int32_t r0, int32_t r1, int32_t dlo,
                   double *data, double *logp, int32_t *indices, int32_t *indices2,
                   int32_t threads);

/* ------------------------------------------------------------------------
 * HOST function: the upper band of one chromosome straight out of a .cool pixel table
 * (cooler's pixels/{bin1_id, bin2_id, count}, sorted by bin1 then bin2; what
 * `clr.matrix(sparse=True, balance=True)[s:e, s:e]` + sp.triu + diag_trim keep of it,
 * cm:527-624).  `bin1` / `bin2` / `count` point at the first pixel with bin1 >= s, n_pix =
 * pixels with bin1 < e; count_dtype 0 int32, 1 int64, 2 float64; `weight` = bins/weight of the
 * whole genome (NULL: raw counts): value = count * w[bin1] * w[bin2], pixels whose value is
 * not finite (masked bins) are dropped.  Writes canonical CSR rows (indptr[e - s + 1],
 * indices / data with room for n_pix entries) and returns the number of entries, < 0 on
 * error.  One fused, multi-threaded pass instead of ~15 numpy passes.
 * ------------------------------------------------------------------------ */
/* HOST helpers of the same reader.  cs_pixels_lex_sorted: 1 when the table is sorted by (bin1,
 * bin2) without duplicates (cooler's invariant; what makes row slices canonical CSR rows), 0 if
 * not, < 0 on error.  cs_pixels_inter_index: the inter-chromosomal pixels grouped by (chromosome
 * of bin1, chromosome of bin2) by a stable counting sort: pixel indices to order[n_pix], block
 * offsets to starts[n_chroms^2 + 1] (block c1 * n_chroms + c2), bin_chrom[b] = chromosome of
 * bin b; returns the number of inter pixels.  The sub-matrices of cm:235-322 then cost one
 * slice each instead of one scan of the table each. */
int cs_pixels_lex_sorted(const int64_t *bin1, const int64_t *bin2, int64_t n_pix);
int64_t cs_pixels_inter_index(const int64_t *bin1, const int64_t *bin2, int64_t n_pix,
                              const int16_t *bin_chrom, int32_t n_chroms, int64_t *order,
                              int64_t *starts);
int64_t cs_band_csr_from_pixels(const int64_t *bin1, const int64_t *bin2, const void *count,
                                int32_t count_dtype, int64_t n_pix, const double *weight,
                                int64_t s, int64_t e, int64_t max_diag, int64_t *indptr,
                                int32_t *indices, double *data, int32_t threads);

/* ------------------------------------------------------------------------
 * Host-buffer, whole-call entry point: what chromosight.utils.detection.
 * normxcorr2 (det:807-914) does for a sparse signal, from host CSR arrays to
 * host CSR arrays, including host<->device copies through pinned staging.
 * The result buffers are owned by the library until cs_result_free().
 * ------------------------------------------------------------------------ */
typedef struct cs_csr_result {
    int64_t nnz;
    int32_t rows, cols;
    int64_t *indptr;  /* rows + 1 */
    int32_t *indices; /* nnz */
    double *data;     /* nnz */
    double *log10p;   /* nnz or NULL */
    int64_t *p_indptr;  /* second copy of indptr / indices for the p-value matrix, so   */
    int32_t *p_indices; /* that the two returned matrices share no storage (NULL if !pval) */
    double ms_h2d, ms_kernels, ms_d2h; /* device-side timings of the call */
    int64_t n_windows;
    int64_t h2d_bytes, d2h_bytes; /* bytes copied host<->device by the call */
} cs_csr_result;

typedef struct cs_normxcorr2_args {
    int32_t rows, cols;
    const int64_t *indptr;
    const int32_t *indices;
    const double *data;
    int32_t has_mask;          /* 0 none, 1 pixel mask (CSR pattern), 2 geometric (below) */
    const int64_t *mask_indptr;
    const int32_t *mask_indices;
    /* has_mask == 2: the mask make_missing_mask(shape, valid_rows, valid_cols, max_dist,
     * sym_upper) would build (pre:535-633), given by its ingredients: uint8[rows] / uint8[cols]
     * (1 = missing bin) and the diagonals col - row on which missing bins flag their pixels
     * (INT32_MIN / INT32_MAX = unbounded).  The frame of frame_missing_mask is added by the
     * library when full. */
    const uint8_t *miss_row;
    const uint8_t *miss_col;
    int32_t mask_dlo, mask_dhi;
    int32_t sym_upper;
    int32_t max_dist; /* -1 = None */
    int32_t full;
    int32_t pval;
    int32_t trim_to_max_dist; /* extension: drop scores beyond max_dist (det:270) */
    int32_t sig_dmin, sig_dmax; /* diagonal extent of the stored signal (col-row); sig_dmin ==
                                 * INT32_MIN: measured by the library (indices sorted per row) */
    cs_kernel_desc kernel;
    double missing_tol;
    int32_t device;
    int32_t raw_xcorr;       /* 1: xcorr2 (det:595-723) instead of normxcorr2 */
    double xcorr_threshold;  /* threshold of xcorr2 when raw_xcorr (normxcorr2 always uses 1e-4) */
    /* 1: `indices` and `data` are DEVICE arrays on `device` (e.g. the output of
     * cs_detrend_apply, never brought to the host); `indptr` stays a host array and sig_dmin /
     * sig_dmax must be given.  Session calls only (cs_session_upload, _upload_run_scores):
     * ContactMap.create_mat -> pattern_detector (cm:607-624, det:253-263) without a host
     * round trip of the detrended sub-matrix. */
    int32_t device_payload;
    /* Score only the output rows out_row0 <= row < out_row1 (matrix rows; out_row1 <= out_row0:
     * all rows); the other rows of the result are empty.  A rank of the row-slab path
     * (rowslab.py) scores its owned rows only, not the halo it received. */
    int32_t out_row0, out_row1;
} cs_normxcorr2_args;

int cs_normxcorr2_host(const cs_normxcorr2_args *a, cs_csr_result *res);
void cs_result_free(cs_csr_result *res);

/* The same call split in phases, so that a caller can keep the inputs resident in
 * HBM and re-run the device part (bench.py's `value` leg; iterated detection,
 * cli:730-792, re-runs the same sub-matrix with a new kernel):
 *   upload   : plan + host CSR -> HBM through pinned staging
 *   run      : K0b fill -> K1 Pearson -> K2 CSR compaction, device only
 *   download : CSR result -> pinned host buffers (free with cs_result_free)
 *   candidates: pixels with score >= threshold into a caller-owned DEVICE buffer. */
typedef struct cs_session cs_session;
typedef struct cs_run_stats {
    double ms_fill, ms_pearson, ms_compact, ms_total; /* CUDA-event times of the run */
    int64_t n_windows, nnz, launches, h2d_bytes, d2h_bytes;
} cs_run_stats;
int cs_session_create(int32_t device, cs_session **out);
void cs_session_destroy(cs_session *s);
/* Launch on the caller's stream (e.g. PyTorch's current stream) instead of the library's
 * own; NULL restores the library stream.  SURVEY 8b: one CUDA stream per call. */
int cs_session_set_stream(cs_session *s, void *stream);
int cs_session_upload(cs_session *s, const cs_normxcorr2_args *a);
int cs_session_run(cs_session *s, cs_run_stats *stats);
/* The same run without waiting for the device (a re-run of the same upload; the first run of an
 * upload, which sizes the result arrays, is synchronous anyway): the caller keeps enqueueing --
 * cs_session_candidates right behind it, one synchronisation for both -- and the run's checks
 * (signal under the mask, capacity of the result arrays) are made by the next call that
 * synchronises.  cs_session_wait waits and returns the statistics of the last run. */
int cs_session_run_enqueue(cs_session *s);
int cs_session_wait(cs_session *s, cs_run_stats *stats);
/* fill -> Pearson only: the scores stay a float32 image in HBM, which is all that
 * cs_session_candidates / _foci / _validate read (pattern_detector, det:265-345, needs
 * p-values at the foci only); cs_session_download compacts on demand. */
int cs_session_run_scores(cs_session *s, cs_run_stats *stats);
/* cs_session_upload + cs_session_run_scores in one call; large inputs are cut into row slabs so
 * that the staging / DMA of slab s+1 overlaps the kernels of slab s (what pattern_detector
 * issues per sub-matrix, det:253-263). */
int cs_session_upload_run_scores(cs_session *s, const cs_normxcorr2_args *a, cs_run_stats *stats);
int cs_session_candidates(cs_session *s, float threshold, int32_t dmin, int32_t dmax,
                          cs_candidate *d_cand, int64_t cap, int64_t *d_count, int64_t *n_host);
int cs_session_download(cs_session *s, cs_csr_result *res);
/* cs_scores_foci on the scores of the last run: at most `cap` records into host_foci, sorted by
 * first pixel (the reference's focus order); *n_host = number of foci found. */
int cs_session_foci(cs_session *s, double threshold, int32_t dmin, int32_t dmax, int32_t min_size,
                    cs_focus *host_foci, int64_t cap, int64_t *n_host);
/* validate_patterns on the session's matrix and last scores: host coordinates in
 * (UNPADDED, n_coords x 2), host results out.  `full`/`inter` select the padding and NaN
 * sub-diagonals of det:291-310; host_valid_row/col are uint8[rows]/[cols] (1 = detectable). */
int cs_session_validate(cs_session *s, const int32_t *host_coords, int64_t n_coords,
                        const uint8_t *host_valid_row, const uint8_t *host_valid_col,
                        int32_t inter, double zero_tol, double missing_tol, int32_t score_dmax,
                        double *host_windows, uint8_t *host_valid, double *host_score,
                        double *host_log10p);

#ifdef __cplusplus
}
#endif
#endif /* CHROMOSIGHT_B200_H */
