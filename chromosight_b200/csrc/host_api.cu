// Host-buffer entry point: chromosight.utils.detection.normxcorr2 (det:807-914)
// for a sparse signal, host CSR in -> host CSR out.
//
//   host CSR --memcpy--> pinned staging --cudaMemcpyAsync--> HBM
//        K0b image fill -> K1 Pearson tiles -> K2 CSR compaction (+ p-values)
//   HBM --cudaMemcpyAsync--> pinned result buffers (pooled, handed to the caller)
//
// Device and pinned buffers are cached per device and only grow, so steady-state
// calls do no allocation.
#include <mutex>
#include <vector>
#include "common.cuh"

namespace cs {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return CS_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CS_CUDA(cudaMalloc(&p, want));
        cap = want;
        return CS_OK;
    }
};

struct PinBlock {
    void *p;
    size_t cap;
    bool busy;
};

constexpr size_t kStageBytes = 32u << 20;

struct HostCtx {
    int device = -1;
    cudaStream_t st = nullptr;
    DevBuf sig_indptr, sig_indices, sig_data, m_indptr, m_indices, img, out, nobs, r_indptr,
        r_indices, r_data, r_p, err;
    void *stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<PinBlock> pool;
    std::mutex mu;
};

static std::mutex g_ctx_mu;
static std::vector<HostCtx *> g_ctx;

static int get_ctx(int device, HostCtx **out) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (HostCtx *c : g_ctx)
        if (c->device == device) {
            *out = c;
            return CS_OK;
        }
    CS_CUDA(cudaSetDevice(device));
    HostCtx *c = new HostCtx();
    c->device = device;
    CS_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CS_CUDA(cudaHostAlloc(&c->stage[i], kStageBytes, cudaHostAllocDefault));
        CS_CUDA(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 4; ++i) CS_CUDA(cudaEventCreate(&c->ev[i]));
    g_ctx.push_back(c);
    *out = c;
    return CS_OK;
}

// pageable host -> device through the two pinned staging buffers (CPU memcpy of chunk
// i+1 overlaps the DMA of chunk i)
static int h2d_staged(HostCtx *c, void *dst, const void *src, size_t bytes) {
    size_t done = 0;
    int k = 0;
    while (done < bytes) {
        const size_t n = bytes - done < kStageBytes ? bytes - done : kStageBytes;
        CS_CUDA(cudaEventSynchronize(c->stage_ev[k]));
        memcpy(c->stage[k], (const char *)src + done, n);
        CS_CUDA(cudaMemcpyAsync((char *)dst + done, c->stage[k], n, cudaMemcpyHostToDevice, c->st));
        CS_CUDA(cudaEventRecord(c->stage_ev[k], c->st));
        done += n;
        k ^= 1;
    }
    return CS_OK;
}

static int pin_alloc(HostCtx *c, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    PinBlock *best = nullptr;
    for (PinBlock &b : c->pool)
        if (!b.busy && b.cap >= bytes && (!best || b.cap < best->cap)) best = &b;
    if (best && best->cap <= 2 * bytes + (1u << 20)) {
        best->busy = true;
        *out = best->p;
        return CS_OK;
    }
    // drop idle blocks that are too small before growing
    for (size_t i = 0; i < c->pool.size();) {
        if (!c->pool[i].busy && c->pool[i].cap < bytes) {
            cudaFreeHost(c->pool[i].p);
            c->pool.erase(c->pool.begin() + i);
        } else
            ++i;
    }
    void *p = nullptr;
    CS_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    c->pool.push_back({p, bytes, true});
    *out = p;
    return CS_OK;
}

static void pin_release(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (HostCtx *c : g_ctx)
        for (PinBlock &b : c->pool)
            if (b.p == p) {
                b.busy = false;
                return;
            }
}

}  // namespace cs

using namespace cs;

extern "C" void cs_result_free(cs_csr_result *r) {
    if (!r) return;
    pin_release(r->indptr);
    pin_release(r->indices);
    pin_release(r->data);
    pin_release(r->log10p);
    r->indptr = nullptr;
    r->indices = nullptr;
    r->data = nullptr;
    r->log10p = nullptr;
}

extern "C" int cs_normxcorr2_host(const cs_normxcorr2_args *a, cs_csr_result *res) {
    CS_REQUIRE(a && res, "cs_normxcorr2_host: null argument");
    CS_REQUIRE(a->rows > 0 && a->cols > 0 && a->indptr && a->indices && a->data,
               "cs_normxcorr2_host: bad signal");
    memset(res, 0, sizeof(*res));
    HostCtx *c = nullptr;
    int rc = get_ctx(a->device, &c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(a->device));
    cudaStream_t st = c->st;
    const cs_kernel_desc &K = a->kernel;
    const int mk = K.kh, nk = K.kw;
    const int kh = (mk - 1) / 2, kw = (nk - 1) / 2;
    const int pr = a->full ? mk - 1 : 0, pc = a->full ? nk - 1 : 0;
    const int H = a->rows + 2 * pr, W = a->cols + 2 * pc;
    int oy0, oy1, ox0, ox1;
    if (a->full) {
        oy0 = pr, oy1 = pr + a->rows, ox0 = pc, ox1 = pc + a->cols;
    } else {
        oy0 = kh, oy1 = a->rows - kh, ox0 = kw, ox1 = a->cols - kw;
    }
    res->rows = a->rows;
    res->cols = a->cols;
    const int64_t nnz_in = a->indptr[a->rows];
    const bool empty = (oy1 <= oy0) || (ox1 <= ox0) || nnz_in == 0 || a->sig_dmax < a->sig_dmin;

    // diagonal ranges in image coordinates
    const int sh = pc - pr;
    long long od_lo = (long long)a->sig_dmin + sh - (kh + kw);
    long long od_hi = (long long)a->sig_dmax + sh + (kh + kw);
    if (a->sym_upper && od_lo < sh) od_lo = sh;  // det:1098-1099 (triu of the cropped map)
    if (a->trim_to_max_dist && a->max_dist >= 0 && od_hi > (long long)a->max_dist + sh)
        od_hi = (long long)a->max_dist + sh;
    if (!empty) {
        const long long dmin_poss = (long long)ox0 - (oy1 - 1), dmax_poss = (long long)(ox1 - 1) - oy0;
        if (od_lo < dmin_poss) od_lo = dmin_poss;
        if (od_hi > dmax_poss) od_hi = dmax_poss;
    }
    if (empty || od_hi < od_lo) {
        void *ip = nullptr;
        rc = pin_alloc(c, (size_t)(a->rows + 1) * sizeof(int64_t), &ip);
        if (rc) return rc;
        memset(ip, 0, (size_t)(a->rows + 1) * sizeof(int64_t));
        res->indptr = (int64_t *)ip;
        return CS_OK;
    }
    const int id_lo = (int)(od_lo - (kh + kw)), id_hi = (int)(od_hi + (kh + kw));

    cs_layout Li, Lo;
    const bool band_img = (long long)(id_hi - id_lo + 1) * 2 < (long long)W;
    if (band_img)
        rc = cs_layout_band(&Li, H, W, id_lo, id_hi);
    else
        rc = cs_layout_dense(&Li, H, W);
    if (rc) return rc;
    // scores live in original coordinates: diagonal = image diagonal - sh
    const bool band_out = (od_hi - od_lo + 1) * 2 < (long long)a->cols;
    if (band_out)
        rc = cs_layout_band(&Lo, a->rows, a->cols, (int)(od_lo - sh), (int)(od_hi - sh));
    else
        rc = cs_layout_dense(&Lo, a->rows, a->cols);
    if (rc) return rc;

    // ---- device buffers ---------------------------------------------------------
    const size_t n_ip = (size_t)a->rows + 1;
    if ((rc = c->sig_indptr.ensure(n_ip * sizeof(int64_t)))) return rc;
    if ((rc = c->sig_indices.ensure((size_t)nnz_in * sizeof(int32_t)))) return rc;
    if ((rc = c->sig_data.ensure((size_t)nnz_in * sizeof(double)))) return rc;
    int64_t nnz_m = 0;
    if (a->has_mask) {
        CS_REQUIRE(a->mask_indptr && a->mask_indices, "mask arrays missing");
        nnz_m = a->mask_indptr[a->rows];
        if ((rc = c->m_indptr.ensure(n_ip * sizeof(int64_t)))) return rc;
        if ((rc = c->m_indices.ensure((size_t)(nnz_m > 0 ? nnz_m : 1) * sizeof(int32_t)))) return rc;
    }
    if ((rc = c->img.ensure((size_t)Li.n_elems * sizeof(float)))) return rc;
    if ((rc = c->out.ensure((size_t)Lo.n_elems * sizeof(float)))) return rc;
    const bool want_nobs = a->pval && a->has_mask && a->full;
    if (want_nobs)
        if ((rc = c->nobs.ensure((size_t)Lo.n_elems * sizeof(uint16_t)))) return rc;
    if ((rc = c->r_indptr.ensure(n_ip * sizeof(int64_t)))) return rc;
    if ((rc = c->err.ensure(64))) return rc;

    // ---- H2D ----------------------------------------------------------------------
    CS_CUDA(cudaEventRecord(c->ev[0], st));
    if ((rc = h2d_staged(c, c->sig_indptr.p, a->indptr, n_ip * sizeof(int64_t)))) return rc;
    if ((rc = h2d_staged(c, c->sig_indices.p, a->indices, (size_t)nnz_in * sizeof(int32_t)))) return rc;
    if ((rc = h2d_staged(c, c->sig_data.p, a->data, (size_t)nnz_in * sizeof(double)))) return rc;
    if (a->has_mask) {
        if ((rc = h2d_staged(c, c->m_indptr.p, a->mask_indptr, n_ip * sizeof(int64_t)))) return rc;
        if (nnz_m > 0)
            if ((rc = h2d_staged(c, c->m_indices.p, a->mask_indices, (size_t)nnz_m * sizeof(int32_t))))
                return rc;
    }
    CS_CUDA(cudaEventRecord(c->ev[1], st));

    // ---- kernels --------------------------------------------------------------------
    rc = cs_image_fill_f32(&Li, (float *)c->img.p, (const int64_t *)c->sig_indptr.p,
                           (const int32_t *)c->sig_indices.p, (const double *)c->sig_data.p,
                           a->rows, a->cols, pr, pc, a->has_mask ? 1 : 0,
                           (const int64_t *)c->m_indptr.p, (const int32_t *)c->m_indices.p,
                           a->sym_upper, a->max_dist, a->full ? mk : 0, a->full ? nk : 0,
                           (int32_t *)c->err.p, st);
    if (rc) return rc;
    // scores outside the computed set must read as 0
    CS_CUDA(cudaMemsetAsync(c->out.p, 0, (size_t)Lo.n_elems * sizeof(float), st));
    cs_pearson_opts po;
    memset(&po, 0, sizeof(po));
    po.has_mask = a->has_mask;
    po.missing_tol = a->missing_tol;
    po.xcorr_threshold = a->raw_xcorr ? a->xcorr_threshold : 1e-4;
    po.raw_xcorr = a->raw_xcorr;
    po.nobs_full = want_nobs ? 1 : 0;
    po.out_row_shift = pr;
    po.out_col_shift = pc;
    rc = cs_pearson_f32(&Li, (const float *)c->img.p, &K, &po, oy0, oy1, ox0, ox1, (int)od_lo,
                        (int)od_hi, &Lo, (float *)c->out.p, want_nobs ? (uint16_t *)c->nobs.p : nullptr,
                        st);
    if (rc) return rc;
    // windows evaluated (the metric's unit)
    {
        long long nw = 0;
        for (int Y = oy0; Y < oy1; ++Y) {
            long long lo = (long long)Y + od_lo, hi = (long long)Y + od_hi;
            if (lo < ox0) lo = ox0;
            if (hi > ox1 - 1) hi = ox1 - 1;
            if (hi >= lo) nw += hi - lo + 1;
        }
        res->n_windows = nw;
    }
    int64_t nnz = 0;
    rc = cs_scores_count(&Lo, (const float *)c->out.p, -(1 << 30), (1 << 30),
                         (int64_t *)c->r_indptr.p, &nnz, st);
    if (rc) return rc;
    int32_t herr[2] = {0, 0};
    CS_CUDA(cudaMemcpyAsync(herr, c->err.p, sizeof(herr), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    if (herr[0] > 0 && a->has_mask) {
        set_error("There are %d non-zero elements reported as missing.", herr[0]);
        return CS_ERR_MASKED_SIGNAL;
    }
    if (herr[1] > 0) {
        set_error("internal: %d signal pixels fell outside the stored band", herr[1]);
        return CS_ERR_INVALID;
    }
    if ((rc = c->r_indices.ensure((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t)))) return rc;
    if ((rc = c->r_data.ensure((size_t)(nnz > 0 ? nnz : 1) * sizeof(double)))) return rc;
    if (a->pval)
        if ((rc = c->r_p.ensure((size_t)(nnz > 0 ? nnz : 1) * sizeof(double)))) return rc;
    if (nnz > 0) {
        rc = cs_scores_emit(&Lo, (const float *)c->out.p, want_nobs ? (const uint16_t *)c->nobs.p : nullptr,
                            mk * nk, -(1 << 30), (1 << 30), (const int64_t *)c->r_indptr.p,
                            (int32_t *)c->r_indices.p, (double *)c->r_data.p,
                            a->pval ? (double *)c->r_p.p : nullptr, st);
        if (rc) return rc;
    }
    CS_CUDA(cudaEventRecord(c->ev[2], st));

    // ---- D2H into pooled pinned buffers ------------------------------------------------
    void *h_ip = nullptr, *h_ix = nullptr, *h_d = nullptr, *h_p = nullptr;
    if ((rc = pin_alloc(c, n_ip * sizeof(int64_t), &h_ip))) return rc;
    if ((rc = pin_alloc(c, (size_t)nnz * sizeof(int32_t), &h_ix))) return rc;
    if ((rc = pin_alloc(c, (size_t)nnz * sizeof(double), &h_d))) return rc;
    if (a->pval)
        if ((rc = pin_alloc(c, (size_t)nnz * sizeof(double), &h_p))) return rc;
    CS_CUDA(cudaMemcpyAsync(h_ip, c->r_indptr.p, n_ip * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (nnz > 0) {
        CS_CUDA(cudaMemcpyAsync(h_ix, c->r_indices.p, (size_t)nnz * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaMemcpyAsync(h_d, c->r_data.p, (size_t)nnz * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
        if (a->pval)
            CS_CUDA(cudaMemcpyAsync(h_p, c->r_p.p, (size_t)nnz * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
    }
    CS_CUDA(cudaEventRecord(c->ev[3], st));
    CS_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    res->ms_h2d = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
    res->ms_kernels = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
    res->ms_d2h = ms;
    res->nnz = nnz;
    res->indptr = (int64_t *)h_ip;
    res->indices = (int32_t *)h_ix;
    res->data = (double *)h_d;
    res->log10p = (double *)h_p;
    return CS_OK;
}
