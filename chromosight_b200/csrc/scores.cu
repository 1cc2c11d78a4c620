// K2: score image -> what the reference hands back.
//   detection.py:1098-1131   non-zero scores as CSR float64 + log10 p-values
//   stats.py:43-81           corr_to_pval (Fisher z, two-sided, log10)
//   detection.py:417-421     pick_foci's thresholding, fused as a candidate list
#include "common.cuh"

namespace cs {

struct ScoreView {
    int rows, cols, dlo, dhi, pitch, dense;
    int dmin, dmax;  // keep dmin <= col - row <= dmax
};

__device__ __forceinline__ void row_range(const ScoreView &S, int y, int &x0, int &x1) {
    long long lo = (long long)y + S.dmin, hi = (long long)y + S.dmax;
    if (!S.dense) {
        if (lo < (long long)y + S.dlo) lo = (long long)y + S.dlo;
        if (hi > (long long)y + S.dhi) hi = (long long)y + S.dhi;
    }
    if (lo < 0) lo = 0;
    if (hi > S.cols - 1) hi = S.cols - 1;
    x0 = (int)lo;
    x1 = (int)hi + 1;  // exclusive; may be <= x0
}
__device__ __forceinline__ long long sidx(const ScoreView &S, int y, int x) {
    return (long long)y * S.pitch + (x - (S.dense ? 0 : S.dlo));
}

__global__ void count_rows(ScoreView S, const float *__restrict__ sc, int64_t *counts, int r0,
                           int r1) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = r0 + blockIdx.x * wpb + (threadIdx.x >> 5); y < r1; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        int c = 0;
        // eight independent loads per lane and pass (a row of <= 256 scores in one pass)
        for (int xb = x0 + lane; xb < x1; xb += 256) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (xb + 32 * k < x1) ? sc[sidx(S, y, xb + 32 * k)] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) c += (v[k] != 0.f);
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) counts[y + 1] = c;
    }
    if (r0 == 0 && blockIdx.x == 0 && threadIdx.x == 0) counts[0] = 0;
}

// In-place inclusive scan of a[1..n] (a[0] = 0), three launches: every block scans a chunk
// of kScanChunk elements and records its total, one block scans the totals, every block adds
// its offset.
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanChunk = kScanThreads * kScanPer;

__device__ __forceinline__ long long block_exclusive(long long v, long long *wsum, long long &total) {
    // exclusive prefix of v over the block (blockDim.x <= 1024); total = block sum
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        long long w = (lane < (int)(blockDim.x >> 5)) ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    total = wsum[(blockDim.x >> 5) - 1];
    const long long pre = (wid > 0 ? wsum[wid - 1] : 0) + inc - v;
    __syncthreads();
    return pre;
}

__global__ void scan_chunks(int64_t *a, int n, int64_t *totals) {
    __shared__ long long wsum[32];
    const long long base = 1 + (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanPer;
    long long v[kScanPer], s = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        v[k] = (base + k <= n) ? a[base + k] : 0;
        s += v[k];
    }
    long long total;
    long long run = block_exclusive(s, wsum, total);
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        run += v[k];
        if (base + k <= n) a[base + k] = run;
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

// exclusive scan of the chunk totals by one block
__global__ void scan_totals(int64_t *totals, int m, int64_t *grand) {
    __shared__ long long wsum[32];
    long long carry = 0;
    for (int base = 0; base < m; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = i < m ? totals[i] : 0;
        long long total;
        const long long pre = block_exclusive(v, wsum, total);
        if (i < m) totals[i] = carry + pre;
        carry += total;
    }
    if (threadIdx.x == 0 && grand) {
        *grand = carry;  // may be pinned host memory
        __threadfence_system();
    }
}

__global__ void scan_add(int64_t *a, int n, const int64_t *totals, long long base_off) {
    const long long off = totals[blockIdx.x] + base_off;
    if (off == 0) return;
    const long long base = 1 + (long long)blockIdx.x * kScanChunk + (long long)threadIdx.x * kScanPer;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)
        if (base + k <= n) a[base + k] += off;
}

__global__ void emit_rows(ScoreView S, const float *__restrict__ sc,
                          NmissPlane nobs,
                          const int64_t *__restrict__ indptr, int32_t *indices, double *data,
                          double *log10p, int r0, int r1) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = r0 + blockIdx.x * wpb + (threadIdx.x >> 5); y < r1; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        int64_t pos = indptr[y];
        for (int xb = x0; xb < x1; xb += 32) {
            const int x = xb + lane;
            float v = 0.f;
            long long i = 0;
            if (x < x1) {
                i = sidx(S, y, x);
                v = sc[i];
            }
            const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
            if (v != 0.f) {
                const int64_t o = pos + __popc(m & ((1u << lane) - 1u));
                indices[o] = x;
                data[o] = (double)v;
                if (log10p) {
                    const float n = nobs_at(nobs, i);
                    log10p[o] = log10_pval(v, n);
                }
            }
            pos += __popc(m);
        }
    }
}

// The same rows in the narrow wire format of the host path: float32 score, float32 log10 p
// and the column as a uint8 offset from the first stored diagonal of the row (bands of at
// most 256 diagonals): 9 bytes per stored score instead of 20.  The float64 CSR arrays scipy
// gets are expanded from these on the host (host_expand.cpp), losslessly: the scores are
// float32 on the device anyway and the p-values are evaluated in float32.
__global__ void emit_rows_narrow(ScoreView S, const float *__restrict__ sc, NmissPlane nobs,
                                 const int64_t *__restrict__ indptr, float *score, float *log10p,
                                 unsigned char *off, int r0, int r1, long long limit) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = r0 + blockIdx.x * wpb + (threadIdx.x >> 5); y < r1; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        int64_t pos = indptr[y];
        const int xbase = y + S.dlo;  // column of offset 0
        for (int xb = x0; xb < x1; xb += 32) {
            const int x = xb + lane;
            float v = 0.f;
            long long i = 0;
            if (x < x1) {
                i = sidx(S, y, x);
                v = sc[i];
            }
            const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
            const int64_t o = pos + __popc(m & ((1u << lane) - 1u));
            if (v != 0.f && o < limit) {  // (limit: capacity of the arrays when sized from an earlier run)
                score[o] = v;
                off[o] = (unsigned char)(x - xbase);
                if (log10p) log10p[o] = (float)log10_pval(v, nobs_at(nobs, i));
            }
            pos += __popc(m);
        }
    }
}

__global__ void emit_candidates(ScoreView S, const float *__restrict__ sc,
                                NmissPlane nobs,
                                float threshold, cs_candidate *out, long long cap,
                                unsigned long long *count) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < S.rows; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        for (int xb = x0; xb < x1; xb += 32) {
            const int x = xb + lane;
            float v = 0.f;
            long long i = 0;
            bool hit = false;
            if (x < x1) {
                i = sidx(S, y, x);
                v = sc[i];
                hit = (v != 0.f) && (v >= threshold);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0) continue;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const unsigned long long o = base + __popc(m & ((1u << lane) - 1u));
                if ((long long)o < cap) {
                    cs_candidate c;
                    c.row = y;
                    c.col = x;
                    c.score = v;
                    const float n = nobs_at(nobs, i);
                    c.log10p = (float)log10_pval(v, n);
                    out[o] = c;
                }
            }
        }
    }
}

// The same records from a list of pixels (the refinement list of exact_refine_enqueue: every
// pixel >= threshold - 2e-5 on the scanned diagonals) instead of a second pass over the band.
// The list length lives on the device; an overflowed list (> list_cap) emits nothing.
__global__ void emit_candidates_list(ScoreView S, const float *__restrict__ sc, NmissPlane nobs, float threshold,
                                     const int2 *__restrict__ list, const unsigned long long *n_dev,
                                     long long list_cap, cs_candidate *out, long long cap,
                                     unsigned long long *count) {
    const unsigned long long nd = *n_dev;
    if (nd > (unsigned long long)list_cap) return;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < (long long)nd;
         k += (long long)gridDim.x * blockDim.x) {
        const int2 p = list[k];
        const long long i = sidx(S, p.x, p.y);
        const float v = sc[i];
        if (v != 0.f && v >= threshold) {
            const unsigned long long o = atomicAdd(count, 1ull);
            if ((long long)o < cap) {
                cs_candidate c;
                c.row = p.x;
                c.col = p.y;
                c.score = v;
                c.log10p = (float)log10_pval(v, nobs_at(nobs, i));
                out[o] = c;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// pick_foci on the device (detection.py:387-456, label_foci 459-554, filter_foci 557-592):
// pixels with score >= threshold form 4-connected foci; foci of at least min_size pixels
// yield one record: the focus' first pixel in row-major order (which numbers the foci) and
// its best pixel (highest score, first in row-major order among equals).
// Labels live in a uint32 image with the score image's layout; the label of a focus is the
// linear index of its first pixel (union by smaller index), so no renumbering pass is needed.
// ---------------------------------------------------------------------------
constexpr unsigned kNoLabel = 0xffffffffu;

__device__ __forceinline__ bool is_candidate(float v, double thr) {
    return v != 0.f && (double)v >= thr;  // the reference compares float64 values
}

__device__ __forceinline__ unsigned uf_find(const unsigned *lab, unsigned i) {
    unsigned p = ((const volatile unsigned *)lab)[i];
    while (p != i) {
        i = p;
        p = ((const volatile unsigned *)lab)[i];
    }
    return i;
}

__device__ __forceinline__ void uf_union(unsigned *lab, unsigned a, unsigned b) {
    for (;;) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        if (a > b) {
            const unsigned t = a;
            a = b;
            b = t;
        }
        // hang the larger root under the smaller one
        const unsigned old = atomicMin(&lab[b], a);
        if (old == b) return;
        b = old;
    }
}

// pass 1: label = own index for candidates, none elsewhere; statistics cleared
__global__ void foci_init(ScoreView S, const float *__restrict__ sc, double thr, unsigned *lab,
                          unsigned *cnt, unsigned long long *best) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < S.rows; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        for (int x = x0 + lane; x < x1; x += 32) {
            const long long i = sidx(S, y, x);
            const bool c = is_candidate(sc[i], thr);
            lab[i] = c ? (unsigned)i : kNoLabel;
            if (c) {
                cnt[i] = 0u;
                best[i] = 0ull;
            }
        }
    }
}

// pass 2: unions with the left and the upper neighbour (4-connectivity, det:508-540)
__global__ void foci_merge(ScoreView S, unsigned *lab) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < S.rows; y += gridDim.x * wpb) {
        int x0, x1, u0 = 0, u1 = 0;
        row_range(S, y, x0, x1);
        if (y > 0) row_range(S, y - 1, u0, u1);
        for (int x = x0 + lane; x < x1; x += 32) {
            const long long i = sidx(S, y, x);
            if (lab[i] == kNoLabel) continue;
            if (x > x0 && lab[i - 1] != kNoLabel) uf_union(lab, (unsigned)i, (unsigned)(i - 1));
            if (y > 0 && x >= u0 && x < u1) {
                const long long j = sidx(S, y - 1, x);
                if (lab[j] != kNoLabel) uf_union(lab, (unsigned)i, (unsigned)j);
            }
        }
    }
}

// monotone map float -> uint32 (larger float, larger key)
__device__ __forceinline__ unsigned float_key(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// pass 3: every candidate adds itself to its root: size, and the best (score, first index)
__global__ void foci_reduce(ScoreView S, const float *__restrict__ sc, unsigned *lab,
                            unsigned *cnt, unsigned long long *best) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < S.rows; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        for (int x = x0 + lane; x < x1; x += 32) {
            const long long i = sidx(S, y, x);
            if (lab[i] == kNoLabel) continue;
            const unsigned r = uf_find(lab, (unsigned)i);
            lab[i] = r;  // flatten
            atomicAdd(&cnt[r], 1u);
            // larger score wins; among equal scores the smaller index
            const unsigned long long key =
                ((unsigned long long)float_key(sc[i]) << 32) | (unsigned long long)(~(unsigned)i);
            atomicMax(&best[r], key);
        }
    }
}

// pass 4: one record per root with enough pixels
__global__ void foci_emit(ScoreView S, const float *__restrict__ sc, const unsigned *lab,
                          const unsigned *cnt, const unsigned long long *best, int min_size,
                          cs_focus *out, long long cap, unsigned long long *count) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int xoff = S.dense ? 0 : S.dlo;
    for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < S.rows; y += gridDim.x * wpb) {
        int x0, x1;
        row_range(S, y, x0, x1);
        for (int x = x0 + lane; x < x1; x += 32) {
            const long long i = sidx(S, y, x);
            if (lab[i] != (unsigned)i || (int)cnt[i] < min_size) continue;
            const unsigned long long o = atomicAdd(count, 1ull);
            if ((long long)o >= cap) continue;
            const unsigned bi = ~(unsigned)(best[i] & 0xffffffffull);
            cs_focus f;
            f.first_row = y;
            f.first_col = x;
            if (S.dense) {
                f.row = (int)(bi / (unsigned)S.pitch);
                f.col = (int)(bi - (unsigned)f.row * (unsigned)S.pitch);
            } else {
                // band: index = row * (pitch + 1) + (col - row - dlo)
                f.row = (int)(bi / (unsigned)(S.pitch + 1));
                f.col = f.row + (int)(bi - (unsigned)f.row * (unsigned)(S.pitch + 1)) + xoff;
            }
            f.score = sc[bi];
            f.size = (int)cnt[i];
            out[o] = f;
        }
    }
}

static ScoreView make_view(const cs_layout *L, int dmin, int dmax) {
    ScoreView S;
    S.rows = L->rows;
    S.cols = L->cols;
    S.dlo = L->dlo;
    S.dhi = L->dhi;
    S.pitch = L->pitch;
    S.dense = L->dense;
    S.dmin = dmin;
    S.dmax = dmax;
    return S;
}

static int64_t *scan_scratch_of(const cs_layout *Lo, int64_t *d_indptr) {
    // chunk totals live behind the row pointers (d_indptr holds rows + 1 + cs_scan_scratch(rows))
    return d_indptr + (size_t)Lo->rows + 1;
}

int scores_count_rows(const cs_layout *Lo, const float *d_out, int32_t dmin, int32_t dmax,
                      int64_t *d_indptr, int32_t r0, int32_t r1, int64_t *d_total, cudaStream_t st,
                      int32_t scratch_off) {
    if (r1 <= r0) return CS_OK;
    ScoreView S = make_view(Lo, dmin, dmax);
    const int n = r1 - r0;
    int grid = (n + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    count_rows<<<grid, 256, 0, st>>>(S, d_out, d_indptr, r0, r1);
    CS_LAUNCHED();
    const int nchunk = (n + kScanChunk - 1) / kScanChunk;
    int64_t *totals = scan_scratch_of(Lo, d_indptr) + scratch_off;
    scan_chunks<<<nchunk, kScanThreads, 0, st>>>(d_indptr + r0, n, totals);
    CS_LAUNCHED();
    scan_totals<<<1, 1024, 0, st>>>(totals, nchunk, d_total);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int scores_finish_rows(const cs_layout *Lo, int64_t *d_indptr, int32_t r0, int32_t r1, int64_t base,
                       cudaStream_t st, int32_t scratch_off) {
    if (r1 <= r0) return CS_OK;
    const int n = r1 - r0;
    const int nchunk = (n + kScanChunk - 1) / kScanChunk;
    scan_add<<<nchunk, kScanThreads, 0, st>>>(d_indptr + r0, n,
                                              scan_scratch_of(Lo, d_indptr) + scratch_off,
                                              (long long)base);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int scores_emit_rows(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                     int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                     const int64_t *d_indptr,
                     int32_t r0, int32_t r1, int32_t *d_indices, double *d_data, double *d_log10p,
                     cudaStream_t st) {
    if (r1 <= r0) return CS_OK;
    ScoreView S = make_view(Lo, dmin, dmax);
    int grid = (r1 - r0 + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    emit_rows<<<grid, 256, 0, st>>>(S, d_out, NmissPlane{d_nmiss, nmiss_bytes, n_window}, d_indptr, d_indices, d_data,
                                    d_log10p, r0, r1);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // namespace cs

using namespace cs;

// Row ranges scanned independently (the slabs of the host pipeline) own disjoint parts of
// the scratch: a range starting at row r0, the k-th of a call, uses cs_scan_slot(r0, k).
int cs::scores_emit_rows_narrow(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                                int32_t nmiss_bytes, int32_t n_window, const int64_t *d_indptr,
                                int32_t r0, int32_t r1, float *d_score, float *d_log10p,
                                uint8_t *d_off, cudaStream_t st, int64_t limit) {
    if (r1 <= r0) return CS_OK;
    CS_REQUIRE(!Lo->dense && Lo->dhi - Lo->dlo < 256, "narrow result format needs a band of <= 256 diagonals");
    ScoreView S = make_view(Lo, -(1 << 30), 1 << 30);
    int grid = (r1 - r0 + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    emit_rows_narrow<<<grid, 256, 0, st>>>(S, d_out, NmissPlane{d_nmiss, nmiss_bytes, n_window},
                                           d_indptr, d_score, d_log10p, d_off, r0, r1, (long long)limit);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int64_t cs_scan_scratch(int32_t rows) {
    return (int64_t)((rows + cs::kScanChunk - 1) / cs::kScanChunk) + 2 + 256;
}
int32_t cs::scan_slot(int32_t r0, int32_t k) { return r0 / cs::kScanChunk + k; }

extern "C" int cs_scores_count(const cs_layout *Lo, const float *d_out, int32_t dmin, int32_t dmax,
                               int64_t *d_indptr, int64_t *nnz_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(Lo && d_out && d_indptr && nnz_host, "cs_scores_count: null argument");
    int rc = scores_count_rows(Lo, d_out, dmin, dmax, d_indptr, 0, Lo->rows, nullptr, st);
    if (rc) return rc;
    if ((rc = scores_finish_rows(Lo, d_indptr, 0, Lo->rows, 0, st))) return rc;
    CS_CUDA(cudaMemcpyAsync(nnz_host, d_indptr + Lo->rows, sizeof(int64_t), cudaMemcpyDeviceToHost,
                            st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

extern "C" int cs_scores_emit(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                              int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                              const int64_t *d_indptr, int32_t *d_indices, double *d_data,
                              double *d_log10p, void *stream) {
    CS_REQUIRE(Lo && d_out && d_indptr && d_indices && d_data, "cs_scores_emit: null argument");
    return scores_emit_rows(Lo, d_out, d_nmiss, nmiss_bytes, n_window, dmin, dmax, d_indptr, 0, Lo->rows,
                            d_indices, d_data, d_log10p, (cudaStream_t)stream);
}

int cs::scores_candidates_enqueue(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                                  int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                                  float threshold, cs_candidate *d_cand, int64_t cap, int64_t *d_count,
                                  cudaStream_t st) {
    ScoreView S = make_view(Lo, dmin, dmax);
    CS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    int grid = (S.rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    emit_candidates<<<grid, 256, 0, st>>>(S, d_out, NmissPlane{d_nmiss, nmiss_bytes, n_window}, threshold, d_cand,
                                          (long long)cap, (unsigned long long *)d_count);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int cs::scores_candidates_from_list(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                                    int32_t nmiss_bytes, int32_t n_window, int32_t dmin, int32_t dmax,
                                    float threshold, const int2 *d_list, const unsigned long long *d_n,
                                    long long list_cap, cs_candidate *d_cand, int64_t cap, int64_t *d_count,
                                    cudaStream_t st) {
    ScoreView S = make_view(Lo, dmin, dmax);
    CS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    emit_candidates_list<<<148 * 4, 256, 0, st>>>(S, d_out, NmissPlane{d_nmiss, nmiss_bytes, n_window}, threshold,
                                                  d_list, d_n, list_cap, d_cand, (long long)cap,
                                                  (unsigned long long *)d_count);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_scores_candidates(const cs_layout *Lo, const float *d_out, const void *d_nmiss,
                                    int32_t nmiss_bytes, int32_t n_window, int32_t dmin,
                                    int32_t dmax, float threshold,
                                    cs_candidate *d_cand, int64_t cap, int64_t *d_count,
                                    int64_t *n_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(Lo && d_out && d_cand && d_count && n_host, "cs_scores_candidates: null argument");
    int rc = scores_candidates_enqueue(Lo, d_out, d_nmiss, nmiss_bytes, n_window, dmin, dmax, threshold, d_cand,
                                       cap, d_count, st);
    if (rc) return rc;
    CS_CUDA(cudaMemcpyAsync(n_host, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}

extern "C" int64_t cs_foci_work_bytes(const cs_layout *L) {
    if (!L) return 0;
    // label (4 B) + size (4 B) + best (8 B) per image element
    return (int64_t)L->n_elems * 16 + 256;
}

extern "C" int cs_scores_foci(const cs_layout *L, const float *d_scores, int32_t dmin, int32_t dmax,
                              double threshold, int32_t min_size, void *d_work, cs_focus *d_foci,
                              int64_t cap, int64_t *d_count, int64_t *n_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(L && d_scores && d_work && d_foci && d_count && n_host,
               "cs_scores_foci: null argument");
    CS_REQUIRE(L->n_elems < (1ll << 32) - 1, "score image too large for 32-bit focus labels");
    if (!L->dense) CS_REQUIRE(L->pitch >= L->dhi - L->dlo, "band layout pitch too small");
    ScoreView S = make_view(L, dmin, dmax);
    unsigned char *w = (unsigned char *)d_work;
    unsigned long long *best = (unsigned long long *)w;              // 8-byte aligned first
    unsigned *lab = (unsigned *)(w + (size_t)L->n_elems * 8);
    unsigned *cnt = lab + L->n_elems;
    CS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    int grid = (S.rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    foci_init<<<grid, 256, 0, st>>>(S, d_scores, threshold, lab, cnt, best);
    CS_LAUNCHED();
    foci_merge<<<grid, 256, 0, st>>>(S, lab);
    CS_LAUNCHED();
    foci_reduce<<<grid, 256, 0, st>>>(S, d_scores, lab, cnt, best);
    CS_LAUNCHED();
    foci_emit<<<grid, 256, 0, st>>>(S, d_scores, lab, cnt, best, min_size, d_foci, cap,
                                    (unsigned long long *)d_count);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaMemcpyAsync(n_host, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}
