// Host-buffer entry point: chromosight.utils.detection.normxcorr2 (det:807-914)
// for a sparse signal, host CSR in -> host CSR out.
//
//   host CSR --memcpy--> pinned staging --cudaMemcpyAsync--> HBM
//        K0b image fill -> K1 Pearson tiles -> K2 CSR compaction (+ p-values)
//   HBM --cudaMemcpyAsync--> pinned result buffers (pooled, handed to the caller)
//
// Device and pinned buffers are cached per device and only grow, so steady-state
// calls do no allocation.
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
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

constexpr int kStageBufs = 4;  // staging buffers in rotation: host copies run ahead of the DMA
constexpr size_t kStageBytes = 16u << 20;
static const size_t kStageChunk = [] {  // one DMA per chunk
    const char *e = getenv("CS_STAGE_CHUNK_MB");
    size_t mb = (e && atoi(e) > 0) ? (size_t)atoi(e) : 16;
    if (mb > 16) mb = 16;
    return mb << 20;
}();

struct HostCtx {
    int device = -1;
    cudaStream_t st = nullptr, st_copy = nullptr, st_emit = nullptr, st_up = nullptr;
    void *stage[kStageBufs] = {nullptr};
    cudaEvent_t stage_ev[kStageBufs] = {nullptr};
    int stage_next = 0;  // rotation continues across calls of h2d_staged
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<PinBlock> pool;
    std::mutex mu;       // device work of one call phase
    std::mutex pool_mu;  // the pinned-block pool (alloc and release come from any thread)
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
    CS_CUDA(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
    CS_CUDA(cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking));
    {
        // result compaction of finished slabs overtakes the Pearson tiles of later ones
        int lo_pri = 0, hi_pri = 0;
        CS_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CS_CUDA(cudaStreamCreateWithPriority(&c->st_emit, cudaStreamNonBlocking, hi_pri));
    }
    for (int i = 0; i < kStageBufs; ++i) {
        CS_CUDA(cudaHostAlloc(&c->stage[i], kStageBytes, cudaHostAllocDefault));
        CS_CUDA(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 4; ++i) CS_CUDA(cudaEventCreate(&c->ev[i]));
    g_ctx.push_back(c);
    *out = c;
    return CS_OK;
}

// memcpy split over the upload pool's threads (up to 4; CS_COPY_THREADS overrides): the staging
// copy of pageable inputs is on the device's critical path.
// (g_wide_copy: a scores-only call has no download and no widening competing for the host's
// memory system and cores -- its staging copies may use the widening pool's threads as well)
static std::atomic<int> g_wide_copy{0};
static void memcpy_mt(void *dst, const void *src, size_t n) {
    const size_t kMin = 1u << 20;
    const bool wide = g_wide_copy.load(std::memory_order_relaxed) > 0;
    const int maxt = wide ? std::max(expand_threads_default(), upload_threads_default()) : upload_threads_default();
    int nt = (int)(n / kMin);
    if (nt > maxt) nt = maxt;
    static const bool plain = getenv("CS_STAGE_PLAIN_MEMCPY") != nullptr;
    auto cp = [](void *d, const void *s_, size_t m) {
        if (plain) memcpy(d, s_, m);
        else copy_stream(d, s_, m);
    };
    if (nt <= 1) {
        cp(dst, src, n);
        return;
    }
    const size_t per = ((n / nt) + 4095) & ~(size_t)4095;
    parallel_for(nt, [&](int t) {
        const size_t o = per * (size_t)t;
        if (o >= n) return;
        const size_t len = (o + per > n) ? n - o : per;
        cp((char *)dst + o, (const char *)src + o, len);
    }, wide ? 0 : 1);
}

// true when `p` lies in page-locked host memory CUDA knows about (cudaHostAlloc /
// cudaHostRegister, e.g. a torch pinned tensor): such buffers are DMA'd directly
static bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// pageable host -> device through the two pinned staging buffers (CPU memcpy of chunk
// i+1 overlaps the DMA of chunk i)
struct StageTimes {  // CS_TRACE: where the staging thread's time goes (ms)
    double wait = 0, copy = 0, api = 0;
};
static thread_local StageTimes *g_stage_times = nullptr;
static inline double ms_now() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename Poll>
static int h2d_staged(HostCtx *c, cudaStream_t st, void *dst, const void *src, size_t bytes,
                      Poll poll) {
    size_t done = 0;
    int k = c->stage_next;
    if (bytes >= (1u << 16) && is_pinned(src) && is_pinned((const char *)src + bytes - 1)) {
        if (int prc = poll()) return prc;
        CS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return CS_OK;
    }
    while (done < bytes) {
        const size_t n = bytes - done < kStageChunk ? bytes - done : kStageChunk;
        if (int prc = poll()) return prc;
        const double t0 = g_stage_times ? ms_now() : 0.0;
        CS_CUDA(cudaEventSynchronize(c->stage_ev[k]));
        const double t1 = g_stage_times ? ms_now() : 0.0;
        memcpy_mt(c->stage[k], (const char *)src + done, n);
        const double t2 = g_stage_times ? ms_now() : 0.0;
        CS_CUDA(cudaMemcpyAsync((char *)dst + done, c->stage[k], n, cudaMemcpyHostToDevice, st));
        CS_CUDA(cudaEventRecord(c->stage_ev[k], st));
        if (g_stage_times) {
            const double t3 = ms_now();
            g_stage_times->wait += t1 - t0;
            g_stage_times->copy += t2 - t1;
            g_stage_times->api += t3 - t2;
        }
        done += n;
        k = (k + 1) % kStageBufs;
    }
    c->stage_next = k;
    return CS_OK;
}
static int h2d_staged(HostCtx *c, cudaStream_t st, void *dst, const void *src, size_t bytes) {
    return h2d_staged(c, st, dst, src, bytes, [] { return 0; });
}

static int pin_alloc(HostCtx *c, size_t bytes, void **out) {
    std::lock_guard<std::mutex> plk(c->pool_mu);
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

// caller holds no lock on g_ctx_mu (pin_release takes it); used from scope guards
static void pin_release(void *p);
static void pin_release_unlocked(void *p) { pin_release(p); }

static void pin_release(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (HostCtx *c : g_ctx) {
        std::lock_guard<std::mutex> plk(c->pool_mu);
        for (PinBlock &b : c->pool)
            if (b.p == p) {
                b.busy = false;
                return;
            }
    }
}

}  // namespace cs

using namespace cs;

extern "C" int cs_expand_rows(const float *score, const float *log10p, const uint8_t *off,
                              const int64_t *indptr, int32_t r0, int32_t r1, int32_t dlo,
                              double *data, double *logp, int32_t *indices, int32_t *indices2,
                              int32_t threads) {
    CS_REQUIRE(score && off && indptr && data && indices && r1 >= r0, "cs_expand_rows: bad arguments");
    CS_REQUIRE((log10p != nullptr) == (logp != nullptr), "cs_expand_rows: log10p and logp go together");
    expand_rows(score, log10p, off, indptr, r0, r1, dlo, data, logp, indices, indices2,
                threads > 0 ? threads : expand_threads_default());
    return CS_OK;
}

extern "C" int cs_pixels_lex_sorted(const int64_t *bin1, const int64_t *bin2, int64_t n_pix) {
    if (!(bin1 && bin2) || n_pix < 0) {
        set_error("cs_pixels_lex_sorted: bad arguments");
        return CS_ERR_INVALID;
    }
    return pixels_lex_sorted(bin1, bin2, n_pix, expand_threads_default());
}

extern "C" int64_t cs_pixels_inter_index(const int64_t *bin1, const int64_t *bin2, int64_t n_pix,
                                         const int16_t *bin_chrom, int32_t n_chroms, int64_t *order,
                                         int64_t *starts) {
    if (!(bin1 && bin2 && bin_chrom && order && starts) || n_pix < 0 || n_chroms < 1 || n_chroms > 4096) {
        set_error("cs_pixels_inter_index: bad arguments");
        return CS_ERR_INVALID;
    }
    return pixels_inter_index(bin1, bin2, n_pix, bin_chrom, n_chroms, order, starts, expand_threads_default());
}

extern "C" int64_t cs_band_csr_from_pixels(const int64_t *bin1, const int64_t *bin2, const void *count,
                                           int32_t count_dtype, int64_t n_pix, const double *weight,
                                           int64_t s, int64_t e, int64_t max_diag, int64_t *indptr,
                                           int32_t *indices, double *data, int32_t threads) {
    if (!(bin1 && bin2 && count && indptr && indices && data && e > s && n_pix >= 0 && count_dtype >= 0 &&
          count_dtype <= 2)) {
        set_error("cs_band_csr_from_pixels: bad arguments");
        return CS_ERR_INVALID;
    }
    return band_csr_from_pixels(bin1, bin2, count, count_dtype, n_pix, weight, s, e, max_diag, indptr, indices,
                                data, threads > 0 ? threads : expand_threads_default());
}

extern "C" void cs_result_free(cs_csr_result *r) {
    if (!r) return;
    pin_release(r->indptr);
    pin_release(r->indices);
    pin_release(r->data);
    pin_release(r->log10p);
    pin_release(r->p_indptr);
    pin_release(r->p_indices);
    r->p_indptr = nullptr;
    r->p_indices = nullptr;
    r->indptr = nullptr;
    r->indices = nullptr;
    r->data = nullptr;
    r->log10p = nullptr;
}


// ---------------------------------------------------------------------------
// session: plan + device-resident inputs of one normxcorr2 call
// ---------------------------------------------------------------------------
struct cs_session {
    HostCtx *c = nullptr;
    cudaStream_t user_st = nullptr;
    bool use_user_st = false;
    cudaStream_t stream() const { return use_user_st ? user_st : c->st; }
    std::mutex call_mu;  // one host-to-host call (upload, run, download) at a time
    // device-resident inputs, images and results of this session
    DevBuf sig_indptr, sig_indices, sig_data, m_indptr, m_indices, img, out, nobs, r_indptr,
        r_indices, r_data, r_p, err, g_coords, g_win, g_flag, g_score, g_p, g_vrow, g_vcol, f_work,
        f_rec, geo_bits, w_score, w_logp, w_off, x_list;
    long long n_refined = 0;
    bool narrow = false;   // result held in the narrow wire format (band of <= 256 diagonals)
    cs_geo_mask geo;       // has_mask == 2: the mask's geometry in image coordinates
    int pearson_mask = 0;  // mask mode handed to the Pearson kernel (0 / 1 NaN sentinels / 2 geometric)
    int nmiss_bytes = 1;   // element size of the missing-count plane
    int trim_lo = 0, trim_hi = 0;  // diagonals of a pixel mask the frame keeps (pre:452-454)
    double refined_thr = 0.0;      // threshold the scores were last refined for
    bool refined = false;
    bool uploaded = false, ran = false, empty = false, compacted = false;
    cs_normxcorr2_args a;
    std::vector<double> k_corr, k_mask, k2_mask;
    cs_layout Li, Lo;
    int oy0 = 0, oy1 = 0, ox0 = 0, ox1 = 0, od_lo = 0, od_hi = 0, pr = 0, pc = 0;
    bool want_nobs = false;
    int64_t nnz_in = 0, nnz_m = 0, nnz_out = 0, n_windows = 0;
    bool planes_zeroed = false;  // score / count planes zeroed since the upload (a run overwrites every window)
    int64_t nnz_hint = -1;   // non-zero scores of the last compaction of this upload (sizes the next one)
    int64_t nnz_async = 0;   // read-back target of a compaction enqueued without synchronisation
    bool pending = false;    // a run was enqueued without synchronisation (cs_session_run_enqueue)
    int32_t herr_async[2] = {0, 0};
    long long l0_pending = 0;
    size_t h2d_bytes = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

extern "C" int cs_session_create(int32_t device, cs_session **out) {
    CS_REQUIRE(out, "cs_session_create: null argument");
    HostCtx *c = nullptr;
    int rc = get_ctx(device, &c);
    if (rc) return rc;
    cs_session *s = new cs_session();
    s->c = c;
    for (int i = 0; i < 6; ++i) CS_CUDA(cudaEventCreate(&s->ev[i]));
    *out = s;
    return CS_OK;
}

extern "C" void cs_session_destroy(cs_session *s) {
    if (!s) return;
    for (int i = 0; i < 6; ++i)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    DevBuf *bufs[] = {&s->sig_indptr, &s->sig_indices, &s->sig_data, &s->m_indptr, &s->m_indices,
                      &s->img,        &s->out,         &s->nobs,     &s->r_indptr, &s->r_indices,
                      &s->r_data,     &s->r_p,         &s->err,      &s->g_coords, &s->g_win,
                      &s->g_flag,     &s->g_score,     &s->g_p,      &s->g_vrow,   &s->g_vcol,
                      &s->f_work,     &s->f_rec,       &s->geo_bits,
                      &s->w_score,    &s->w_logp,      &s->w_off,
                      &s->x_list};
    cudaSetDevice(s->c->device);
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    delete s;
}

extern "C" int cs_session_set_stream(cs_session *s, void *stream) {
    CS_REQUIRE(s, "cs_session_set_stream: null session");
    s->user_st = (cudaStream_t)stream;
    s->use_user_st = stream != nullptr;
    return CS_OK;
}

// Plan the call and copy its inputs to the device (through pinned staging).
static void session_pearson_opts(const cs_session *s, cs_pearson_opts *po);

static int session_upload_impl(cs_session *s, const cs_normxcorr2_args *a, bool skip_payload) {
    CS_REQUIRE(s && a, "cs_session_upload: null argument");
    CS_REQUIRE(a->rows > 0 && a->cols > 0 && a->indptr && a->indices && a->data,
               "cs_session_upload: bad signal");
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    if (s->pending) {  // a run nobody waited for: let it drain, its results are replaced
        cudaStreamSynchronize(s->stream());
        s->pending = false;
    }
    cudaStream_t st = s->stream();
    s->uploaded = s->ran = false;
    s->nnz_hint = -1;
    s->planes_zeroed = false;
    s->a = *a;
    const cs_kernel_desc &K = a->kernel;
    CS_REQUIRE(K.kh >= 1 && K.kw >= 1 && K.k_corr, "cs_session_upload: bad kernel");
    const int nk2 = K.kh * K.kw;
    s->k_corr.assign(K.k_corr, K.k_corr + nk2);
    s->k_mask.assign(K.k_mask ? K.k_mask : K.k_corr, (K.k_mask ? K.k_mask : K.k_corr) + nk2);
    if (K.k2_mask)
        s->k2_mask.assign(K.k2_mask, K.k2_mask + nk2);
    else {
        s->k2_mask.resize(nk2);
        for (int i = 0; i < nk2; ++i) s->k2_mask[i] = K.k_corr[i] * K.k_corr[i];
    }
    s->a.kernel.k_corr = s->k_corr.data();
    s->a.kernel.k_mask = s->k_mask.data();
    s->a.kernel.k2_mask = s->k2_mask.data();
    s->a.indptr = nullptr;  // host arrays are not kept
    s->a.indices = nullptr;
    s->a.data = nullptr;
    s->a.mask_indptr = nullptr;
    s->a.mask_indices = nullptr;
    s->a.miss_row = nullptr;
    s->a.miss_col = nullptr;

    const int mk = K.kh, nk = K.kw;
    const int kh = (mk - 1) / 2, kw = (nk - 1) / 2;
    s->pr = a->full ? mk - 1 : 0;
    s->pc = a->full ? nk - 1 : 0;
    const int pr = s->pr, pc = s->pc;
    const int H = a->rows + 2 * pr, W = a->cols + 2 * pc;
    if (a->full) {
        s->oy0 = pr, s->oy1 = pr + a->rows, s->ox0 = pc, s->ox1 = pc + a->cols;
    } else {
        s->oy0 = kh, s->oy1 = a->rows - kh, s->ox0 = kw, s->ox1 = a->cols - kw;
    }
    if (a->out_row1 > a->out_row0) {
        // a row range of the result only (first matrix row = image row pr in full mode, 0 else)
        const int base = a->full ? pr : 0;
        if (base + a->out_row0 > s->oy0) s->oy0 = base + a->out_row0;
        if (base + a->out_row1 < s->oy1) s->oy1 = base + a->out_row1;
    }
    s->nnz_in = a->indptr[a->rows];
    int sig_dmin = a->sig_dmin, sig_dmax = a->sig_dmax;
    if (a->device_payload)
        CS_REQUIRE(a->sig_dmin != INT32_MIN && a->device == c->device,
                   "device payload: sig_dmin / sig_dmax must be given, arrays on the session's device");
    if (a->sig_dmin == INT32_MIN) {
        // diagonal extent from the first and last stored column of every row (two cache
        // misses per row: split over a few threads)
        const int nt = a->rows >= (1 << 16) ? 8 : 1;
        std::vector<int> lo_t(nt, INT32_MAX), hi_t(nt, INT32_MIN);
        auto scan = [&](int t) {
            const int r0 = (int)((long long)a->rows * t / nt), r1 = (int)((long long)a->rows * (t + 1) / nt);
            int lo = INT32_MAX, hi = INT32_MIN;
            for (int r = r0; r < r1; ++r) {
                const int64_t e0 = a->indptr[r], e1 = a->indptr[r + 1];
                if (e1 <= e0) continue;
                const int l = a->indices[e0] - r, h = a->indices[e1 - 1] - r;
                if (l < lo) lo = l;
                if (h > hi) hi = h;
            }
            lo_t[t] = lo;
            hi_t[t] = hi;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(scan, t);
        scan(0);
        for (auto &t : th) t.join();
        int dmin = INT32_MAX, dmax = INT32_MIN;
        for (int t = 0; t < nt; ++t) {
            if (lo_t[t] < dmin) dmin = lo_t[t];
            if (hi_t[t] > dmax) dmax = hi_t[t];
        }
        if (dmin > dmax) dmin = 0, dmax = -1;
        s->a.sig_dmin = sig_dmin = dmin;
        s->a.sig_dmax = sig_dmax = dmax;
    }
    s->empty = (s->oy1 <= s->oy0) || (s->ox1 <= s->ox0) || s->nnz_in == 0 ||
               sig_dmax < sig_dmin;
    // diagonal ranges in image coordinates
    const int sh = pc - pr;
    long long od_lo = (long long)sig_dmin + sh - (kh + kw);
    long long od_hi = (long long)sig_dmax + sh + (kh + kw);
    // det:1098-1099: sp.triu of the FRAMED map (before the crop of det:1124-1129) keeps the
    // image diagonals X - Y >= 0; for non-square kernels that is matrix diagonal -(nk - mk)
    if (a->sym_upper && od_lo < 0) od_lo = 0;
    if (a->trim_to_max_dist && a->max_dist >= 0 && od_hi > (long long)a->max_dist + sh)
        od_hi = (long long)a->max_dist + sh;
    if (!s->empty) {
        const long long dmin_poss = (long long)s->ox0 - (s->oy1 - 1);
        const long long dmax_poss = (long long)(s->ox1 - 1) - s->oy0;
        if (od_lo < dmin_poss) od_lo = dmin_poss;
        if (od_hi > dmax_poss) od_hi = dmax_poss;
    }
    if (od_hi < od_lo) s->empty = true;
    s->n_windows = 0;
    s->nnz_out = 0;
    s->h2d_bytes = 0;
    if (s->empty) {
        s->uploaded = true;
        return CS_OK;
    }
    s->od_lo = (int)od_lo;
    s->od_hi = (int)od_hi;
    const int id_lo = (int)(od_lo - (kh + kw)), id_hi = (int)(od_hi + (kh + kw));
    int rc;
    const bool band_img = (long long)(id_hi - id_lo + 1) * 2 < (long long)W;
    rc = band_img ? cs_layout_band(&s->Li, H, W, id_lo, id_hi) : cs_layout_dense(&s->Li, H, W);
    if (rc) return rc;
    // scores live in original coordinates: diagonal = image diagonal - sh
    const bool band_out = (od_hi - od_lo + 1) * 2 < (long long)a->cols;
    rc = band_out ? cs_layout_band(&s->Lo, a->rows, a->cols, (int)(od_lo - sh), (int)(od_hi - sh))
                  : cs_layout_dense(&s->Lo, a->rows, a->cols);
    if (rc) return rc;
    // band results of at most 256 diagonals travel as float32 score / float32 log10 p / uint8
    // diagonal offset and are widened on the host (host_expand.cpp)
    s->narrow = band_out && (od_hi - od_lo) < 256 && !getenv("CS_WIDE_RESULT");
    for (int Y = s->oy0; Y < s->oy1; ++Y) {
        long long lo = (long long)Y + od_lo, hi = (long long)Y + od_hi;
        if (lo < s->ox0) lo = s->ox0;
        if (hi > s->ox1 - 1) hi = s->ox1 - 1;
        if (hi >= lo) s->n_windows += hi - lo + 1;
    }

    // ---- device buffers ---------------------------------------------------------
    const size_t n_ip = (size_t)a->rows + 1;
    if ((rc = s->sig_indptr.ensure(n_ip * sizeof(int64_t)))) return rc;
    if ((rc = s->sig_indices.ensure((size_t)s->nnz_in * sizeof(int32_t)))) return rc;
    if ((rc = s->sig_data.ensure((size_t)s->nnz_in * sizeof(double)))) return rc;
    s->nnz_m = 0;
    CS_REQUIRE(a->has_mask >= 0 && a->has_mask <= 2, "cs_session_upload: bad has_mask");
    memset(&s->geo, 0, sizeof(s->geo));
    const bool wide = K.kh > 31 || K.kw > 31;
    s->pearson_mask = a->has_mask == 2 ? (wide ? 1 : 2) : a->has_mask;
    {
        // missing counts that can go with a non-zero score: N - max(min_present, 1) at most
        const int N = K.kh * K.kw;
        const int min_present = (int)((1.0 - a->missing_tol) * (double)N);
        s->nmiss_bytes = (N - (min_present > 1 ? min_present : 1) <= 255) ? 1 : 2;
    }
    size_t geo_h2d = 0;
    s->trim_lo = -(1 << 29), s->trim_hi = 1 << 29;
    if (a->has_mask) {
        // the frame of frame_missing_mask (pre:404-498) in image coordinates, for both mask forms
        cs_geo_mask &g = s->geo;
        const bool banded = a->sym_upper && a->max_dist >= 0;
        const int big_k = mk > nk ? mk : nk;
        long long lo = a->has_mask == 2 ? (long long)a->mask_dlo : -(1ll << 40);
        long long hi = a->has_mask == 2 ? (long long)a->mask_dhi : (1ll << 40);
        if (a->full && banded) {
            // pre:452-454: the mask is diag-trimmed to max_dist + max(mk, nk) before framing
            if (lo < 0) lo = 0;
            if (hi > (long long)a->max_dist + big_k) hi = (long long)a->max_dist + big_k;
            s->trim_lo = 0, s->trim_hi = a->max_dist + big_k;
        }
        const long long lim = 1ll << 29;
        lo += pc - pr, hi += pc - pr;
        g.mask_dlo = (int)(lo < -lim ? -lim : (lo > lim ? lim : lo));
        g.mask_dhi = (int)(hi < -lim ? -lim : (hi > lim ? lim : hi));
        g.mat_y0 = pr, g.mat_y1 = pr + a->rows, g.mat_x0 = pc, g.mat_x1 = pc + a->cols;
        g.margin_mode = a->full ? (banded ? 1 : 2) : 0;
        if (banded) {
            const int max_n = a->max_dist + nk, max_m = a->max_dist + mk;
            g.top_x1 = pc + (max_n < a->cols ? max_n : a->cols);   // pre:461-463, 477
            g.right_y0 = H - (max_m + 1) > 0 ? H - (max_m + 1) : 0;  // pre:475
        }
        g.strip_dlo = 0, g.strip_dhi = -1;
        if (a->full && a->sym_upper) {
            g.strip_dlo = -big_k, g.strip_dhi = -1;  // pre:483-497
            // the strip is missing whatever the bins say, and no window of the kept upper
            // triangle reaches below it: the two descriptions stay disjoint
            if (g.mask_dlo < 0) g.mask_dlo = 0;
        }
        // The value under the missing pixels: any constant is exact; one near the mean of the
        // data keeps the float32 sums of masked windows conditioned.  Detrended intra maps have
        // mean 1 on every diagonal; sparse maps (inter-chromosomal: a stored pixel per ~10^4) have
        // mean ~0.  The wide kernel wants NaN sentinels.
        const double area = (double)a->rows * (double)((long long)sig_dmax - sig_dmin + 1 < a->cols
                                                           ? (long long)sig_dmax - sig_dmin + 1
                                                           : a->cols);
        const bool sparse_map = area > 0 && (double)s->nnz_in < 0.2 * area;
        g.fill_value = wide ? __builtin_nanf("") : (sparse_map ? 0.0f : 1.0f);
    }
    if (a->has_mask == 2) {
        CS_REQUIRE(a->miss_row && a->miss_col, "geometric mask: missing-bin vectors missing");
        // bit vectors over image rows / columns, 4 zero words of padding on both sides
        const int nwr = (H + 31) / 32 + 8, nwc = (W + 31) / 32 + 8;
        std::vector<uint32_t> hb((size_t)nwr + nwc, 0u);
        uint32_t *hr = hb.data() + 4, *hc = hb.data() + nwr + 4;
        for (int r = 0; r < a->rows; ++r)
            if (a->miss_row[r]) hr[(r + pr) >> 5] |= 1u << ((r + pr) & 31);
        for (int cidx = 0; cidx < a->cols; ++cidx)
            if (a->miss_col[cidx]) hc[(cidx + pc) >> 5] |= 1u << ((cidx + pc) & 31);
        if ((rc = s->geo_bits.ensure(hb.size() * sizeof(uint32_t)))) return rc;
        // pageable source: staged by the runtime before the call returns
        CS_CUDA(cudaMemcpyAsync(s->geo_bits.p, hb.data(), hb.size() * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, st));
        geo_h2d = hb.size() * sizeof(uint32_t);
        s->geo.d_row_bits = (const uint32_t *)s->geo_bits.p + 4;
        s->geo.d_col_bits = (const uint32_t *)s->geo_bits.p + nwr + 4;
    }
    if (a->has_mask == 1) {
        CS_REQUIRE(a->mask_indptr && a->mask_indices, "mask arrays missing");
        s->nnz_m = a->mask_indptr[a->rows];
        if ((rc = s->m_indptr.ensure(n_ip * sizeof(int64_t)))) return rc;
        if ((rc = s->m_indices.ensure((size_t)(s->nnz_m > 0 ? s->nnz_m : 1) * sizeof(int32_t))))
            return rc;
    }
    s->want_nobs = a->pval && a->has_mask && a->full && !a->raw_xcorr;
    if (band_img) {
        // zeros between the rows' bands as wide as the tile boxes stick out of the band: the
        // Pearson kernel then needs no alias fix-up (cs_layout_band_padded)
        cs_pearson_opts po;
        session_pearson_opts(s, &po);
        int32_t TRp = 32, gap = 0;
        static const bool nopad = getenv("CS_NO_BAND_GAP") != nullptr;
        if (!nopad && cs_pearson_plan(&s->Li, &s->a.kernel, &po, s->oy0, s->oy1, s->ox0, s->ox1, s->od_lo,
                                      s->od_hi, &TRp, &gap) == CS_OK &&
            gap > 0 && gap <= 160)
            if ((rc = cs_layout_band_padded(&s->Li, H, W, id_lo, id_hi, gap))) return rc;
    }
    if ((rc = s->img.ensure((size_t)s->Li.n_elems * sizeof(float)))) return rc;
    if ((rc = s->out.ensure((size_t)s->Lo.n_elems * sizeof(float)))) return rc;
    if (s->want_nobs)
        if ((rc = s->nobs.ensure((size_t)s->Lo.n_elems * (size_t)s->nmiss_bytes))) return rc;
    if ((rc = s->r_indptr.ensure((n_ip + (size_t)cs_scan_scratch(a->rows)) * sizeof(int64_t))))
        return rc;
    if ((rc = s->err.ensure(64))) return rc;

    // ---- H2D ----------------------------------------------------------------------
    CS_CUDA(cudaEventRecord(s->ev[0], st));
    if ((rc = h2d_staged(c, st, s->sig_indptr.p, a->indptr, n_ip * sizeof(int64_t)))) return rc;
    if (a->device_payload) {
        // the CSR entries are already in HBM (device-side preprocessing): device-to-device
        CS_REQUIRE(!skip_payload, "device payload cannot be slab-pipelined");
        if (s->nnz_in > 0) {
            CS_CUDA(cudaMemcpyAsync(s->sig_indices.p, a->indices, (size_t)s->nnz_in * sizeof(int32_t),
                                    cudaMemcpyDeviceToDevice, st));
            CS_CUDA(cudaMemcpyAsync(s->sig_data.p, a->data, (size_t)s->nnz_in * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
        }
    } else if (!skip_payload) {
        if ((rc = h2d_staged(c, st, s->sig_indices.p, a->indices,
                             (size_t)s->nnz_in * sizeof(int32_t))))
            return rc;
        if ((rc = h2d_staged(c, st, s->sig_data.p, a->data, (size_t)s->nnz_in * sizeof(double))))
            return rc;
    }
    s->h2d_bytes = n_ip * sizeof(int64_t) +
                   (a->device_payload ? 0 : (size_t)s->nnz_in * (sizeof(int32_t) + sizeof(double)));
    s->h2d_bytes += geo_h2d;
    if (a->has_mask == 1) {
        if ((rc = h2d_staged(c, st, s->m_indptr.p, a->mask_indptr, n_ip * sizeof(int64_t)))) return rc;
        if (s->nnz_m > 0 && !skip_payload)
            if ((rc = h2d_staged(c, st, s->m_indices.p, a->mask_indices,
                                 (size_t)s->nnz_m * sizeof(int32_t))))
                return rc;
        s->h2d_bytes += n_ip * sizeof(int64_t) + (size_t)s->nnz_m * sizeof(int32_t);
    }
    CS_CUDA(cudaEventRecord(s->ev[1], st));
    s->uploaded = true;
    return CS_OK;
}

// Pearson options of a session's call
static void session_pearson_opts(const cs_session *s, cs_pearson_opts *po) {
    const cs_normxcorr2_args &a = s->a;
    memset(po, 0, sizeof(*po));
    po->mask_mode = s->pearson_mask;
    po->missing_tol = a.missing_tol;
    po->xcorr_threshold = a.raw_xcorr ? a.xcorr_threshold : 1e-4;
    po->raw_xcorr = a.raw_xcorr;
    po->nobs_full = s->want_nobs ? 1 : 0;
    po->out_row_shift = s->pr;
    po->out_col_shift = s->pc;
    po->nmiss_bytes = s->nmiss_bytes;
    po->geo = s->geo;
}

extern "C" int cs_session_upload(cs_session *s, const cs_normxcorr2_args *a) {
    return session_upload_impl(s, a, false);
}

// fill -> Pearson -> CSR compaction, all on the device, inputs already resident.
static int session_compact(cs_session *s, cudaStream_t st, int64_t *nnz_out, bool *deferred = nullptr);

static int session_finish_pending(cs_session *s, cudaStream_t st);
// an enqueued run is awaited and checked before anything reads its results
static int session_settle(cs_session *s) {
    if (!s->pending) return CS_OK;
    cudaStream_t st = s->stream();
    CS_CUDA(cudaStreamSynchronize(st));
    return session_finish_pending(s, st);
}
static void session_fill_stats(cs_session *s, cs_run_stats *stats, long long l0);

static int session_run_impl(cs_session *s, cs_run_stats *stats, bool compact, bool enqueue_only = false) {
    CS_REQUIRE(s && s->uploaded, "cs_session_run: nothing uploaded");
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = s->stream();
    if (s->pending) {
        CS_CUDA(cudaStreamSynchronize(st));
        if (int prc = session_finish_pending(s, st)) return prc;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    s->ran = false;
    s->refined = false;
    if (s->empty) {
        s->nnz_out = 0;
        s->ran = true;
        return CS_OK;
    }
    const cs_normxcorr2_args &a = s->a;
    const cs_kernel_desc &K = a.kernel;
    const long long l0 = g_launches.load();
    CS_CUDA(cudaEventRecord(s->ev[2], st));
    int rc = cs_image_fill_f32(&s->Li, (float *)s->img.p, (const int64_t *)s->sig_indptr.p,
                               (const int32_t *)s->sig_indices.p, (const double *)s->sig_data.p,
                               a.rows, a.cols, s->pr, s->pc, a.has_mask,
                               (const int64_t *)s->m_indptr.p, (const int32_t *)s->m_indices.p,
                               &s->geo, a.sym_upper, a.max_dist, a.full ? K.kh : 0,
                               a.full ? K.kw : 0, (int32_t *)s->err.p, st);
    if (rc) return rc;
    // scores outside the computed set must read as 0.  The kernel writes every window of the
    // region on every run, and missing counts for the same windows of the same mask: the planes
    // are zeroed once per upload, not per run.
    if (!s->planes_zeroed) {
        CS_CUDA(cudaMemsetAsync(s->out.p, 0, (size_t)s->Lo.n_elems * sizeof(float), st));
        if (s->want_nobs)  // missing counts: only windows that have one are written by the kernel
            CS_CUDA(cudaMemsetAsync(s->nobs.p, 0, (size_t)s->Lo.n_elems * (size_t)s->nmiss_bytes, st));
        s->planes_zeroed = true;
    }
    cs_pearson_opts po;
    session_pearson_opts(s, &po);
    CS_CUDA(cudaEventRecord(s->ev[3], st));
    rc = cs_pearson_f32(&s->Li, (const float *)s->img.p, &K, &po, s->oy0, s->oy1, s->ox0, s->ox1,
                        s->od_lo, s->od_hi, &s->Lo, (float *)s->out.p,
                        s->want_nobs ? s->nobs.p : nullptr, st);
    if (rc) return rc;
    CS_CUDA(cudaEventRecord(s->ev[4], st));
    int64_t nnz = 0;
    // (a session member, not a local: an enqueue-only run returns before the copy lands)
    int32_t *herr = s->herr_async;
    herr[0] = herr[1] = 0;
    CS_CUDA(cudaMemcpyAsync(herr, s->err.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    s->compacted = false;
    s->l0_pending = l0;
    bool deferred = false;
    if (compact) {
        if ((rc = session_compact(s, st, &nnz, &deferred))) return rc;
        if (deferred) CS_CUDA(cudaEventRecord(s->ev[5], st));
    }
    if (enqueue_only && deferred) {
        // nothing is awaited: the error counters and the count are checked by the next call that
        // synchronises (cs_session_candidates, cs_session_wait, ...)
        s->pending = true;
        s->nnz_out = s->nnz_hint;
        s->ran = true;
        return CS_OK;
    }
    // (the only synchronisation of a re-run; the first compaction of an upload has one more)
    if (!compact || deferred) CS_CUDA(cudaStreamSynchronize(st));
    if (deferred) {
        if (s->nnz_async > s->nnz_hint) {
            // more scores than the arrays were sized for: compact again, counting first
            if ((rc = session_compact(s, st, &nnz, nullptr))) return rc;
            deferred = false;
        } else {
            nnz = s->nnz_async;
            s->nnz_out = s->nnz_hint = nnz;
            s->compacted = true;
        }
    }
    if (herr[0] > 0 && a.has_mask) {
        set_error("There are %d non-zero elements reported as missing.", herr[0]);
        return CS_ERR_MASKED_SIGNAL;
    }
    // with trim_to_max_dist the band is deliberately narrower than the signal
    if (herr[1] > 0 && !a.trim_to_max_dist) {
        set_error("internal: %d signal pixels fell outside the stored band", herr[1]);
        return CS_ERR_INVALID;
    }
    if (!deferred) {
        CS_CUDA(cudaEventRecord(s->ev[5], st));
        CS_CUDA(cudaStreamSynchronize(st));
    }
    s->nnz_out = nnz;
    s->ran = true;
    if (stats) session_fill_stats(s, stats, l0);
    return CS_OK;
}

static void session_fill_stats(cs_session *s, cs_run_stats *stats, long long l0) {
    const cs_normxcorr2_args &a = s->a;
    const int64_t nnz = s->nnz_out;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]);
    stats->ms_fill = ms;
    cudaEventElapsedTime(&ms, s->ev[3], s->ev[4]);
    stats->ms_pearson = ms;
    cudaEventElapsedTime(&ms, s->ev[4], s->ev[5]);
    stats->ms_compact = ms;
    cudaEventElapsedTime(&ms, s->ev[2], s->ev[5]);
    stats->ms_total = ms;
    stats->n_windows = s->n_windows;
    stats->nnz = nnz;
    stats->launches = g_launches.load() - l0;
    stats->h2d_bytes = (int64_t)s->h2d_bytes;
    const size_t per_nz = s->narrow ? (sizeof(float) + 1 + (a.pval ? sizeof(float) : 0))
                                    : (sizeof(int32_t) + sizeof(double) + (a.pval ? sizeof(double) : 0));
    stats->d2h_bytes = (int64_t)(((size_t)a.rows + 1) * sizeof(int64_t) + (size_t)nnz * per_nz);
}

// The checks a run enqueued without synchronisation left open (the stream has been synchronised
// by the caller): error counters of the fill, capacity of the compacted arrays.
static int session_finish_pending(cs_session *s, cudaStream_t st) {
    if (!s->pending) return CS_OK;
    s->pending = false;
    const cs_normxcorr2_args &a = s->a;
    if (s->herr_async[0] > 0 && a.has_mask) {
        s->ran = false;
        set_error("There are %d non-zero elements reported as missing.", s->herr_async[0]);
        return CS_ERR_MASKED_SIGNAL;
    }
    if (s->herr_async[1] > 0 && !a.trim_to_max_dist) {
        s->ran = false;
        set_error("internal: %d signal pixels fell outside the stored band", s->herr_async[1]);
        return CS_ERR_INVALID;
    }
    if (s->nnz_async > s->nnz_hint) {
        int64_t nnz = 0;
        if (int rc = session_compact(s, st, &nnz, nullptr)) return rc;
    } else {
        s->nnz_out = s->nnz_hint = s->nnz_async;
        s->compacted = true;
    }
    return CS_OK;
}

// cs_session_run without waiting for the device (a re-run of the same upload; the first run of
// an upload, which sizes the result arrays, runs synchronously): the caller goes on enqueueing
// (cs_session_candidates) and the run's checks are made at the next synchronisation.
extern "C" int cs_session_run_enqueue(cs_session *s) {
    return session_run_impl(s, nullptr, true, true);
}

// Wait for an enqueued run and report its statistics (also valid after a synchronous run).
extern "C" int cs_session_wait(cs_session *s, cs_run_stats *stats) {
    CS_REQUIRE(s && s->uploaded, "cs_session_wait: nothing uploaded");
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = s->stream();
    CS_CUDA(cudaStreamSynchronize(st));
    if (int rc = session_finish_pending(s, st)) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        if (!s->empty && s->ran) session_fill_stats(s, stats, s->l0_pending);
    }
    return CS_OK;
}

extern "C" int cs_session_run(cs_session *s, cs_run_stats *stats) {
    return session_run_impl(s, stats, true);
}

// fill -> Pearson only: the scores stay an image in HBM (candidates, foci, validate and lookups
// read it); the CSR compaction and the p-values of every stored score are left to a later
// cs_session_download, which pattern_detector never needs (det:337-339 reads p at the foci)
extern "C" int cs_session_run_scores(cs_session *s, cs_run_stats *stats) {
    return session_run_impl(s, stats, false);
}

// K2 of the last run: non-zero scores -> CSR (narrow wire format for bands of <= 256 diagonals)
static int session_compact(cs_session *s, cudaStream_t st, int64_t *nnz_out, bool *deferred) {
    const cs_normxcorr2_args &a = s->a;
    const cs_kernel_desc &K = a.kernel;
    int64_t nnz = 0;
    int rc;
    if (deferred) *deferred = false;
    if (deferred && s->narrow && s->nnz_hint > 0 && s->w_score.cap >= (size_t)s->nnz_hint * sizeof(float) &&
        s->w_off.cap >= (size_t)s->nnz_hint && (!a.pval || s->w_logp.cap >= (size_t)s->nnz_hint * sizeof(float))) {
        // a re-run of the same upload: the arrays are sized from the last compaction, so counting,
        // scan and emission are enqueued back to back and the count is read with the run's own
        // final synchronisation (the caller checks it against the capacity)
        if ((rc = scores_count_rows(&s->Lo, (const float *)s->out.p, -(1 << 30), (1 << 30),
                                    (int64_t *)s->r_indptr.p, 0, s->Lo.rows, nullptr, st)))
            return rc;
        if ((rc = scores_finish_rows(&s->Lo, (int64_t *)s->r_indptr.p, 0, s->Lo.rows, 0, st))) return rc;
        if ((rc = scores_emit_rows_narrow(&s->Lo, (const float *)s->out.p, s->want_nobs ? s->nobs.p : nullptr,
                                          s->nmiss_bytes, K.kh * K.kw, (const int64_t *)s->r_indptr.p, 0,
                                          a.rows, (float *)s->w_score.p,
                                          a.pval ? (float *)s->w_logp.p : nullptr, (uint8_t *)s->w_off.p, st,
                                          s->nnz_hint)))
            return rc;
        CS_CUDA(cudaMemcpyAsync(&s->nnz_async, (const int64_t *)s->r_indptr.p + s->Lo.rows, sizeof(int64_t),
                                cudaMemcpyDeviceToHost, st));
        *deferred = true;
        *nnz_out = s->nnz_hint;
        return CS_OK;
    }
    rc = cs_scores_count(&s->Lo, (const float *)s->out.p, -(1 << 30), (1 << 30),
                         (int64_t *)s->r_indptr.p, &nnz, st);
    if (rc) return rc;
    const size_t nz1 = (size_t)(nnz > 0 ? nnz : 1);
    if (s->narrow) {
        if ((rc = s->w_score.ensure(nz1 * sizeof(float)))) return rc;
        if ((rc = s->w_off.ensure(nz1))) return rc;
        if (a.pval)
            if ((rc = s->w_logp.ensure(nz1 * sizeof(float)))) return rc;
        if (nnz > 0) {
            rc = scores_emit_rows_narrow(&s->Lo, (const float *)s->out.p,
                                         s->want_nobs ? s->nobs.p : nullptr, s->nmiss_bytes,
                                         K.kh * K.kw, (const int64_t *)s->r_indptr.p, 0, a.rows,
                                         (float *)s->w_score.p,
                                         a.pval ? (float *)s->w_logp.p : nullptr,
                                         (uint8_t *)s->w_off.p, st);
            if (rc) return rc;
        }
    } else {
        if ((rc = s->r_indices.ensure(nz1 * sizeof(int32_t)))) return rc;
        if ((rc = s->r_data.ensure(nz1 * sizeof(double)))) return rc;
        if (a.pval)
            if ((rc = s->r_p.ensure(nz1 * sizeof(double)))) return rc;
        if (nnz > 0) {
            rc = cs_scores_emit(&s->Lo, (const float *)s->out.p,
                                s->want_nobs ? s->nobs.p : nullptr, s->nmiss_bytes, K.kh * K.kw,
                                -(1 << 30), (1 << 30), (const int64_t *)s->r_indptr.p,
                                (int32_t *)s->r_indices.p, (double *)s->r_data.p,
                                a.pval ? (double *)s->r_p.p : nullptr, st);
            if (rc) return rc;
        }
    }
    s->nnz_out = nnz;
    s->nnz_hint = nnz;
    s->compacted = true;
    *nnz_out = nnz;
    return CS_OK;
}

// Exact scores at and near `threshold` (see exact_refine): run once per (run, threshold)
// before the thresholding of pick_foci, so that candidates and foci do not depend on float32
// rounding.  CS_NO_REFINE=1 skips it (timing experiments).
static const long long kRefineCap = 1 << 23;  // 8 M pixels (64 MB of scratch)

static bool refine_off(const cs_session *s) {
    static const bool off = getenv("CS_NO_REFINE") != nullptr;
    return off || s->a.raw_xcorr;
}

static int refine_args(cs_session *s, double threshold, int32_t dmin, int32_t dmax, cs_pearson_opts *po,
                       RefineArgs *Rp) {
    const cs_normxcorr2_args &a = s->a;
    session_pearson_opts(s, po);
    po->mask_mode = a.has_mask;  // the predicate of the exact path, whatever the image holds
    const long long cap = kRefineCap;
    int rc;
    if ((rc = s->x_list.ensure((size_t)cap * sizeof(int2) + 64))) return rc;
    RefineArgs &R = *Rp;
    memset(&R, 0, sizeof(R));
    R.K = &a.kernel;
    R.opts = po;
    R.d_indptr = (const int64_t *)s->sig_indptr.p;
    R.d_indices = (const int32_t *)s->sig_indices.p;
    R.d_data = (const double *)s->sig_data.p;
    R.rows = a.rows, R.cols = a.cols, R.pr = s->pr, R.pc = s->pc;
    R.d_m_indptr = (const int64_t *)s->m_indptr.p;
    R.d_m_indices = (const int32_t *)s->m_indices.p;
    R.trim_lo = s->trim_lo, R.trim_hi = s->trim_hi;
    R.Lo = &s->Lo;
    R.d_out = (float *)s->out.p;
    R.d_nmiss = s->want_nobs ? s->nobs.p : nullptr;
    R.threshold = threshold;
    R.dmin = dmin, R.dmax = dmax;
    R.d_list = (int2 *)s->x_list.p;
    R.cap = cap;
    R.d_count = (unsigned long long *)((char *)s->x_list.p + (size_t)cap * sizeof(int2));
    return CS_OK;
}

static void refine_done(cs_session *s, double threshold, long long n) {
    s->refined = true;
    s->refined_thr = threshold;
    s->n_refined = n;
    // the CSR result (if any) was compacted from the unrefined image
    if (n > 0) s->compacted = false;
}

static int session_refine(cs_session *s, double threshold, int32_t dmin, int32_t dmax, cudaStream_t st) {
    if (s->refined && s->refined_thr == threshold) return CS_OK;
    if (refine_off(s)) return CS_OK;
    cs_pearson_opts po;
    RefineArgs R;
    int rc = refine_args(s, threshold, dmin, dmax, &po, &R);
    if (rc) return rc;
    long long n = 0;
    if ((rc = exact_refine(R, st, &n))) return rc;
    refine_done(s, threshold, n);
    return CS_OK;
}

// Candidate pixels (score >= threshold) of the last run, written to a caller-owned
// device buffer (e.g. the send buffer of the final all-gather).
extern "C" int cs_session_candidates(cs_session *s, float threshold, int32_t dmin, int32_t dmax,
                                     cs_candidate *d_cand, int64_t cap, int64_t *d_count,
                                     int64_t *n_host) {
    CS_REQUIRE(s && s->ran && d_cand && d_count && n_host, "cs_session_candidates: bad arguments");
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    if (s->empty) {
        *n_host = 0;
        return CS_OK;
    }
    cudaStream_t st = s->stream();
    const void *nb = s->want_nobs ? s->nobs.p : nullptr;
    const int nwin = s->a.kernel.kh * s->a.kernel.kw;
    if (!(s->refined && s->refined_thr == (double)threshold) && !refine_off(s)) {
        // refinement and thresholding back to back, ONE synchronisation: the length of the
        // refinement list stays on the device
        cs_pearson_opts po;
        RefineArgs R;
        int rc = refine_args(s, (double)threshold, dmin, dmax, &po, &R);
        if (rc) return rc;
        rc = exact_refine_enqueue(R, st);
        if (rc < 0) return rc;
        if (rc == 0) {
            // every candidate is on the refinement list (score >= threshold - 2e-5 on these
            // diagonals): thresholding reads the list, not the band a second time
            if ((rc = scores_candidates_from_list(&s->Lo, (const float *)s->out.p, nb, s->nmiss_bytes, nwin, dmin,
                                                  dmax, threshold, R.d_list, R.d_count, R.cap, d_cand, cap,
                                                  d_count, st)))
                return rc;
            unsigned long long n_ref = 0;
            CS_CUDA(cudaMemcpyAsync(n_host, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            CS_CUDA(cudaMemcpyAsync(&n_ref, R.d_count, sizeof(n_ref), cudaMemcpyDeviceToHost, st));
            CS_CUDA(cudaStreamSynchronize(st));
            if (int prc = session_finish_pending(s, st)) return prc;
            if ((long long)n_ref <= R.cap) {
                refine_done(s, (double)threshold, (long long)n_ref);
                return CS_OK;
            }
            // the list overflowed (nothing was redone): the two-pass form below
        }
    }
    if (s->pending) {
        CS_CUDA(cudaStreamSynchronize(st));
        if (int prc = session_finish_pending(s, st)) return prc;
    }
    if (int rrc = session_refine(s, (double)threshold, dmin, dmax, st)) return rrc;
    return cs_scores_candidates(&s->Lo, (const float *)s->out.p, nb, s->nmiss_bytes, nwin, dmin, dmax,
                                threshold, d_cand, cap, d_count, n_host, st);
}

// pick_foci (det:387-456) on the scores of the last run; records sorted by first pixel.
extern "C" int cs_session_foci(cs_session *s, double threshold, int32_t dmin, int32_t dmax,
                               int32_t min_size, cs_focus *host_foci, int64_t cap,
                               int64_t *n_host) {
    CS_REQUIRE(s && s->ran && n_host && (cap == 0 || host_foci), "cs_session_foci: bad arguments");
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    if (int src = session_settle(s)) return src;
    *n_host = 0;
    if (s->empty) return CS_OK;
    cudaStream_t st = s->stream();
    int rc;
    if ((rc = s->f_work.ensure((size_t)cs_foci_work_bytes(&s->Lo)))) return rc;
    int64_t dcap = cap > 0 ? cap : 1;
    if ((rc = s->f_rec.ensure((size_t)dcap * sizeof(cs_focus) + 64))) return rc;
    int64_t *d_count = (int64_t *)((char *)s->f_rec.p + (size_t)dcap * sizeof(cs_focus));
    d_count = (int64_t *)(((uintptr_t)d_count + 7) & ~(uintptr_t)7);
    if ((rc = session_refine(s, threshold, dmin, dmax, st))) return rc;
    int64_t n = 0;
    rc = cs_scores_foci(&s->Lo, (const float *)s->out.p, dmin, dmax, threshold, min_size,
                        s->f_work.p, (cs_focus *)s->f_rec.p, dcap, d_count, &n, st);
    if (rc) return rc;
    *n_host = n;
    const int64_t m = n < cap ? n : cap;
    if (m > 0) {
        CS_CUDA(cudaMemcpyAsync(host_foci, s->f_rec.p, (size_t)m * sizeof(cs_focus),
                                cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
        std::sort(host_foci, host_foci + m, [](const cs_focus &a, const cs_focus &b) {
            return a.first_row != b.first_row ? a.first_row < b.first_row : a.first_col < b.first_col;
        });
    }
    return CS_OK;
}

// D2H of the CSR result of the last run into pooled pinned buffers.
extern "C" int cs_session_download(cs_session *s, cs_csr_result *res) {
    CS_REQUIRE(s && res && s->ran, "cs_session_download: run first");
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    if (int src = session_settle(s)) return src;
    cudaStream_t st = s->stream();
    memset(res, 0, sizeof(*res));
    const cs_normxcorr2_args &a = s->a;
    res->rows = a.rows;
    res->cols = a.cols;
    res->n_windows = s->n_windows;
    const size_t n_ip = (size_t)a.rows + 1;
    int rc;
    void *h_ip = nullptr, *h_ix = nullptr, *h_d = nullptr, *h_p = nullptr, *h_ip2 = nullptr,
         *h_ix2 = nullptr;
    if ((rc = pin_alloc(c, n_ip * sizeof(int64_t), &h_ip))) return rc;
    res->indptr = (int64_t *)h_ip;
    if (a.pval) {
        if ((rc = pin_alloc(c, n_ip * sizeof(int64_t), &h_ip2))) return rc;
        res->p_indptr = (int64_t *)h_ip2;
    }
    if (s->empty) {
        memset(h_ip, 0, n_ip * sizeof(int64_t));
        if (h_ip2) memset(h_ip2, 0, n_ip * sizeof(int64_t));
        return CS_OK;
    }
    if (!s->compacted) {  // the last run kept the scores as an image only
        int64_t nz = 0;
        if ((rc = session_compact(s, st, &nz))) return rc;
    }
    const int64_t nnz = s->nnz_out;
    if ((rc = pin_alloc(c, (size_t)nnz * sizeof(int32_t), &h_ix))) return rc;
    if ((rc = pin_alloc(c, (size_t)nnz * sizeof(double), &h_d))) return rc;
    if (a.pval) {
        if ((rc = pin_alloc(c, (size_t)nnz * sizeof(double), &h_p))) return rc;
        if ((rc = pin_alloc(c, (size_t)nnz * sizeof(int32_t), &h_ix2))) return rc;
    }
    if (s->narrow) {
        // wire arrays to pinned scratch, then widened by host threads
        void *h_ws = nullptr, *h_wp = nullptr, *h_wo = nullptr;
        if ((rc = pin_alloc(c, (size_t)nnz * sizeof(float), &h_ws))) return rc;
        if ((rc = pin_alloc(c, (size_t)nnz, &h_wo))) return rc;
        if (a.pval)
            if ((rc = pin_alloc(c, (size_t)nnz * sizeof(float), &h_wp))) return rc;
        struct WGuard {
            void *a, *b, *c;
            ~WGuard() {
                pin_release_unlocked(a);
                pin_release_unlocked(b);
                pin_release_unlocked(c);
            }
        } wguard{h_ws, h_wp, h_wo};
        CS_CUDA(cudaEventRecord(s->ev[0], st));
        CS_CUDA(cudaMemcpyAsync(h_ip, s->r_indptr.p, n_ip * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        if (nnz > 0) {
            CS_CUDA(cudaMemcpyAsync(h_ws, s->w_score.p, (size_t)nnz * sizeof(float), cudaMemcpyDeviceToHost, st));
            CS_CUDA(cudaMemcpyAsync(h_wo, s->w_off.p, (size_t)nnz, cudaMemcpyDeviceToHost, st));
            if (a.pval)
                CS_CUDA(cudaMemcpyAsync(h_wp, s->w_logp.p, (size_t)nnz * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        CS_CUDA(cudaEventRecord(s->ev[1], st));
        CS_CUDA(cudaStreamSynchronize(st));
        if (nnz > 0)
            expand_rows((const float *)h_ws, (const float *)h_wp, (const uint8_t *)h_wo,
                        (const int64_t *)h_ip, 0, a.rows, s->Lo.dlo, (double *)h_d, (double *)h_p,
                        (int32_t *)h_ix, (int32_t *)h_ix2, expand_threads_default());
        if (a.pval) memcpy(h_ip2, h_ip, n_ip * sizeof(int64_t));
        float msn = 0.f;
        cudaEventElapsedTime(&msn, s->ev[0], s->ev[1]);
        res->ms_d2h = msn;
        res->nnz = nnz;
        res->indices = (int32_t *)h_ix;
        res->data = (double *)h_d;
        res->log10p = (double *)h_p;
        res->p_indices = (int32_t *)h_ix2;
        res->d2h_bytes = (int64_t)(n_ip * sizeof(int64_t) +
                                   (size_t)nnz * (sizeof(float) + 1 + (a.pval ? sizeof(float) : 0)));
        return CS_OK;
    }
    CS_CUDA(cudaEventRecord(s->ev[0], st));
    CS_CUDA(cudaMemcpyAsync(h_ip, s->r_indptr.p, n_ip * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (nnz > 0) {
        CS_CUDA(cudaMemcpyAsync(h_ix, s->r_indices.p, (size_t)nnz * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaMemcpyAsync(h_d, s->r_data.p, (size_t)nnz * sizeof(double),
                                cudaMemcpyDeviceToHost, st));
        if (a.pval) {
            CS_CUDA(cudaMemcpyAsync(h_p, s->r_p.p, (size_t)nnz * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
            CS_CUDA(cudaMemcpyAsync(h_ix2, s->r_indices.p, (size_t)nnz * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, st));
        }
    }
    if (a.pval)
        CS_CUDA(cudaMemcpyAsync(h_ip2, s->r_indptr.p, n_ip * sizeof(int64_t),
                                cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaEventRecord(s->ev[1], st));
    CS_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
    res->ms_d2h = ms;
    res->nnz = nnz;
    res->indices = (int32_t *)h_ix;
    res->data = (double *)h_d;
    res->log10p = (double *)h_p;
    res->p_indices = (int32_t *)h_ix2;
    res->d2h_bytes = (int64_t)(n_ip * sizeof(int64_t) * (a.pval ? 2 : 1) +
                               (size_t)nnz * (sizeof(int32_t) * (a.pval ? 2 : 1) + sizeof(double) *
                                                                                      (a.pval ? 2 : 1)));
    return CS_OK;
}

// The uploads of a slab-pipelined call run ahead on a thread of their own (the staging copies
// of pageable inputs are host work): slab k needs the signal (and pixel mask) rows below
// need_of[k]; `done` counts the slabs whose rows are enqueued on the upload stream, ev[k] is
// recorded behind them.
struct SlabUploader {
    std::atomic<int> done{0};
    std::atomic<int> rc{0};
    std::atomic<bool> stop{false};
    std::thread th;
    ~SlabUploader() {
        stop.store(true);
        if (th.joinable()) th.join();
    }
    void start(HostCtx *c, cudaStream_t sth, const cs_normxcorr2_args *a, cs_session *s,
               const std::vector<int> &need_of, cudaEvent_t *evs, bool utrace) {
        const int dev = c->device, ns = (int)need_of.size();
        const bool pixmask = a->has_mask == 1;
        const int *needs = need_of.data();
        int32_t *d_ix = (int32_t *)s->sig_indices.p, *d_mix = (int32_t *)s->m_indices.p;
        double *d_dat = (double *)s->sig_data.p;
        SlabUploader *up = this;
        th = std::thread([=] {
            cudaSetDevice(dev);
            StageTimes stt;
            if (utrace) g_stage_times = &stt;
            const double tu0 = ms_now();
            struct Report {
                bool on;
                StageTimes *t;
                double t0;
                ~Report() {
                    if (on)
                        fprintf(stderr, "uploader: %.2f ms total; waiting for a staging buffer %.2f, host copies %.2f, "
                                        "CUDA calls %.2f\n", ms_now() - t0, t->wait, t->copy, t->api);
                }
            } report{utrace, &stt, tu0};
            int from = 0;
            for (int k = 0; k < ns && !up->stop.load(); ++k) {
                const int to = needs[k];
                int r = CS_OK;
                if (to > from) {
                    const int64_t e0 = a->indptr[from], e1 = a->indptr[to];
                    if (e1 > e0) {
                        r = h2d_staged(c, sth, d_ix + e0, a->indices + e0, (size_t)(e1 - e0) * sizeof(int32_t));
                        if (!r)
                            r = h2d_staged(c, sth, d_dat + e0, a->data + e0, (size_t)(e1 - e0) * sizeof(double));
                    }
                    if (!r && pixmask) {
                        const int64_t m0 = a->mask_indptr[from], m1 = a->mask_indptr[to];
                        if (m1 > m0)
                            r = h2d_staged(c, sth, d_mix + m0, a->mask_indices + m0,
                                           (size_t)(m1 - m0) * sizeof(int32_t));
                    }
                    if (!r && cudaEventRecord(evs[k], sth) != cudaSuccess) r = CS_ERR_CUDA;
                    from = to;
                }
                if (r) {
                    up->rc.store(r);
                    return;
                }
                up->done.store(k + 1, std::memory_order_release);
            }
        });
    }
};

// slab boundaries in image rows (multiples of the tile height; with `ramp` the first slabs are
// small so that the download -- the longest leg -- starts early) and the signal rows each needs
static void plan_slabs(const cs_session *s, int TRp, int nslab, bool ramp, std::vector<int> &Yb,
                       std::vector<int> &need_of) {
    Yb.assign(nslab + 1, 0);
    std::vector<double> w(nslab, 1.0);
    if (ramp && nslab >= 6) w[0] = 0.15, w[1] = 0.3, w[2] = 0.6;
    double tot_w = 0.0, acc_w = 0.0;
    for (double v : w) tot_w += v;
    const int R = s->oy1 - s->oy0;
    Yb[0] = s->oy0;
    for (int i = 1; i <= nslab; ++i) {
        acc_w += w[i - 1];
        long long y = (long long)((double)R * acc_w / tot_w);
        y = (y + TRp / 2) / TRp * TRp;
        Yb[i] = s->oy0 + (int)(y > R ? R : y);
        if (Yb[i] < Yb[i - 1]) Yb[i] = Yb[i - 1];
    }
    Yb[0] = s->oy0;
    Yb[nslab] = s->oy1;
    const int kh = (s->a.kernel.kh - 1) / 2;
    need_of.assign(nslab, 0);
    int prev = 0;
    for (int k = 0; k < nslab; ++k) {
        int need = Yb[k + 1] + kh - s->pr;  // signal rows the slab's windows read
        if (need > s->a.rows || k == nslab - 1) need = s->a.rows;
        if (need < prev) need = prev;
        need_of[k] = prev = need;
    }
}

// Large inputs, scores only: upload, fill and Pearson tiles cut into row slabs so that the
// staging / DMA of slab s+1 overlaps the kernels of slab s; the scores stay an image in HBM
// (what pattern_detector reads through foci / validate).
static int session_upload_run_pipelined(cs_session *s, const cs_normxcorr2_args *a, int nslab,
                                        cs_run_stats *stats) {
    int rc = session_upload_impl(s, a, true);
    if (rc) return rc;
    if (stats) memset(stats, 0, sizeof(*stats));
    s->ran = false;
    s->refined = false;
    s->compacted = false;
    if (s->empty) {
        s->nnz_out = 0;
        s->ran = true;
        return CS_OK;
    }
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = s->stream(), st_h = c->st_up;
    const cs_normxcorr2_args &A = s->a;
    const cs_kernel_desc &K = A.kernel;
    const long long l0 = g_launches.load();
    std::vector<cudaEvent_t> ev_up(nslab);
    for (int i = 0; i < nslab; ++i) CS_CUDA(cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming));
    struct EvGuard {
        std::vector<cudaEvent_t> &a;
        ~EvGuard() {
            for (auto e : a) cudaEventDestroy(e);
        }
    } evguard{ev_up};
    CS_CUDA(cudaEventRecord(s->ev[2], st));
    rc = fill_begin(&s->Li, (float *)s->img.p, A.rows, A.cols, A.has_mask, A.sym_upper, A.max_dist,
                    A.full ? K.kh : 0, A.full ? K.kw : 0, (int32_t *)s->err.p, st);
    if (rc) return rc;
    CS_CUDA(cudaMemsetAsync(s->out.p, 0, (size_t)s->Lo.n_elems * sizeof(float), st));
    if (s->want_nobs)  // missing counts: only windows that have one are written by the kernel
        CS_CUDA(cudaMemsetAsync(s->nobs.p, 0, (size_t)s->Lo.n_elems * (size_t)s->nmiss_bytes, st));
    cs_pearson_opts po;
    session_pearson_opts(s, &po);
    int32_t TRp = 32;
    if ((rc = cs_pearson_tile_rows(&s->Li, &K, &po, s->oy0, s->oy1, s->ox0, s->ox1, s->od_lo,
                                   s->od_hi, &TRp)))
        return rc;
    po.tile_rows = TRp;
    std::vector<int> Yb, need_of;
    plan_slabs(s, TRp, nslab, false, Yb, need_of);
    // the upload stream must not run ahead of the row pointers (uploaded on `st` by the plan)
    CS_CUDA(cudaEventRecord(s->ev[3], st));
    CS_CUDA(cudaStreamWaitEvent(st_h, s->ev[3], 0));
    struct WideCopy {  // for the duration of this call
        WideCopy() { g_wide_copy.fetch_add(1); }
        ~WideCopy() { g_wide_copy.fetch_sub(1); }
    } wide_copy;
    SlabUploader upl;
    upl.start(c, st_h, a, s, need_of, ev_up.data(), false);
    int up_end = 0;
    for (int k = 0; k < nslab; ++k) {
        while (upl.done.load(std::memory_order_acquire) <= k) {
            if (upl.rc.load()) {
                set_error("upload of the input failed");
                return upl.rc.load();
            }
            std::this_thread::sleep_for(std::chrono::microseconds(10));
        }
        const int need = need_of[k];
        if (need > up_end) {
            CS_CUDA(cudaStreamWaitEvent(st, ev_up[k], 0));
            rc = fill_rows(&s->Li, (float *)s->img.p, (const int64_t *)s->sig_indptr.p,
                           (const int32_t *)s->sig_indices.p, (const double *)s->sig_data.p, A.rows,
                           up_end, need, s->pr, s->pc, A.has_mask, (const int64_t *)s->m_indptr.p,
                           (const int32_t *)s->m_indices.p, &s->geo, A.sym_upper, A.max_dist,
                           A.full ? K.kh : 0, A.full ? K.kw : 0, (int32_t *)s->err.p, st);
            if (rc) return rc;
            up_end = need;
        }
        if (Yb[k + 1] > Yb[k]) {
            rc = cs_pearson_f32(&s->Li, (const float *)s->img.p, &K, &po, Yb[k], Yb[k + 1], s->ox0,
                                s->ox1, s->od_lo, s->od_hi, &s->Lo, (float *)s->out.p,
                                s->want_nobs ? s->nobs.p : nullptr, st);
            if (rc) return rc;
        }
    }
    int32_t herr[2] = {0, 0};
    CS_CUDA(cudaMemcpyAsync(herr, s->err.p, sizeof(herr), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaEventRecord(s->ev[5], st));
    CS_CUDA(cudaStreamSynchronize(st));
    if (herr[0] > 0 && A.has_mask) {
        set_error("There are %d non-zero elements reported as missing.", herr[0]);
        return CS_ERR_MASKED_SIGNAL;
    }
    if (herr[1] > 0 && !A.trim_to_max_dist) {
        set_error("internal: %d signal pixels fell outside the stored band", herr[1]);
        return CS_ERR_INVALID;
    }
    s->nnz_out = 0;
    s->ran = true;
    if (stats) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->ev[2], s->ev[5]);
        stats->ms_total = ms;
        stats->n_windows = s->n_windows;
        stats->launches = g_launches.load() - l0;
        stats->h2d_bytes = (int64_t)s->h2d_bytes;
    }
    return CS_OK;
}

// Upload + scores in one call: slab-pipelined for large inputs, else cs_session_upload followed
// by cs_session_run_scores.
extern "C" int cs_session_upload_run_scores(cs_session *s, const cs_normxcorr2_args *a,
                                            cs_run_stats *stats) {
    CS_REQUIRE(s && a && a->indptr, "cs_session_upload_run_scores: null argument");
    long long rows_out = a->full ? a->rows : a->rows - (a->kernel.kh - 1);
    const int64_t nnz_in = a->indptr[a->rows];
    int nslab = 8;
    if (const char *e = getenv("CS_PIPELINE_SLABS")) nslab = atoi(e);
    if (nslab > 1 && nnz_in >= (4 << 20) && rows_out >= 64 * nslab && !a->device_payload)
        return session_upload_run_pipelined(s, a, nslab, stats);
    int rc = cs_session_upload(s, a);
    if (rc) return rc;
    return cs_session_run_scores(s, stats);
}

// Large calls: the same work as upload + run + download, cut into row slabs so that the
// upload of slab s+1, the kernels of slab s and the download of slab s-1 overlap (PCIe is
// full duplex; the CSR result is 5x the input).  Per slab, on the compute stream:
//   H2D of the CSR rows it adds -> fill of those image rows -> Pearson tiles of its output
//   rows -> per-row counts + scan; once its non-zero count is known on the host: row
//   pointers, CSR entries + p-values; then, on the copy stream, its D2H.
static int normxcorr2_pipelined(cs_session *s, const cs_normxcorr2_args *a, cs_csr_result *res,
                                int nslab) {
    const auto t_enter = std::chrono::steady_clock::now();
    int rc = session_upload_impl(s, a, true);
    if (rc) return rc;
    if (s->empty) return 1;  // nothing to pipeline: the caller finishes through run + download
    const double ms_plan = std::chrono::duration<double, std::milli>(
                               std::chrono::steady_clock::now() - t_enter).count();
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = s->stream(), st_d = c->st_copy, st_e = c->st_emit, st_h = c->st_up;
    const cs_normxcorr2_args &A = s->a;
    const cs_kernel_desc &K = A.kernel;
    const int kh = (K.kh - 1) / 2;
    const size_t n_ip = (size_t)A.rows + 1;
    memset(res, 0, sizeof(*res));
    res->rows = A.rows;
    res->cols = A.cols;
    res->n_windows = s->n_windows;
    // worst-case result buffers: every window non-zero
    const size_t cap = (size_t)s->n_windows;
    const bool narrow = s->narrow;
    void *h_ip = nullptr, *h_ix = nullptr, *h_d = nullptr, *h_p = nullptr, *h_ip2 = nullptr,
         *h_ix2 = nullptr, *h_tot = nullptr, *h_ws = nullptr, *h_wp = nullptr, *h_wo = nullptr;
    if (narrow) {
        if ((rc = s->w_score.ensure(cap * sizeof(float)))) return rc;
        if ((rc = s->w_off.ensure(cap))) return rc;
        if (A.pval)
            if ((rc = s->w_logp.ensure(cap * sizeof(float)))) return rc;
        // pinned landing buffers of the wire arrays (scratch of this call)
        if ((rc = pin_alloc(c, cap * sizeof(float), &h_ws))) return rc;
        if ((rc = pin_alloc(c, cap, &h_wo))) return rc;
        if (A.pval)
            if ((rc = pin_alloc(c, cap * sizeof(float), &h_wp))) return rc;
    } else {
        if ((rc = s->r_indices.ensure(cap * sizeof(int32_t)))) return rc;
        if ((rc = s->r_data.ensure(cap * sizeof(double)))) return rc;
        if (A.pval)
            if ((rc = s->r_p.ensure(cap * sizeof(double)))) return rc;
    }
    struct WireGuard {
        void *a, *b, *c;
        ~WireGuard() {
            pin_release_unlocked(a);
            pin_release_unlocked(b);
            pin_release_unlocked(c);
        }
    } wireguard{h_ws, h_wp, h_wo};
    if ((rc = pin_alloc(c, n_ip * sizeof(int64_t), &h_ip))) return rc;
    res->indptr = (int64_t *)h_ip;
    if ((rc = pin_alloc(c, cap * sizeof(int32_t), &h_ix))) return rc;
    res->indices = (int32_t *)h_ix;
    if ((rc = pin_alloc(c, cap * sizeof(double), &h_d))) return rc;
    res->data = (double *)h_d;
    if (A.pval) {
        if ((rc = pin_alloc(c, cap * sizeof(double), &h_p))) return rc;
        res->log10p = (double *)h_p;
        if ((rc = pin_alloc(c, n_ip * sizeof(int64_t), &h_ip2))) return rc;
        res->p_indptr = (int64_t *)h_ip2;
        if ((rc = pin_alloc(c, cap * sizeof(int32_t), &h_ix2))) return rc;
        res->p_indices = (int32_t *)h_ix2;
    }
    if ((rc = pin_alloc(c, (size_t)(nslab + 1) * sizeof(int64_t), &h_tot))) return rc;
    struct Guard {  // the totals block goes back to the pool whatever happens
        void *p;
        ~Guard() { pin_release_unlocked(p); }
    } guard{h_tot};
    int64_t *tot = (int64_t *)h_tot;
    std::vector<cudaEvent_t> ev_tot(nslab), ev_emit(nslab), ev_up(nslab);
    for (int i = 0; i < nslab; ++i) {
        CS_CUDA(cudaEventCreateWithFlags(&ev_tot[i], cudaEventDisableTiming));
        CS_CUDA(cudaEventCreateWithFlags(&ev_emit[i], cudaEventDisableTiming));
        CS_CUDA(cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming));
    }
    struct EvGuard {
        std::vector<cudaEvent_t> &a, &b, &c;
        ~EvGuard() {
            for (auto e : a) cudaEventDestroy(e);
            for (auto e : b) cudaEventDestroy(e);
            for (auto e : c) cudaEventDestroy(e);
        }
    } evguard{ev_tot, ev_emit, ev_up};

    // A helper thread follows the downloads slab by slab.
    // Narrow format: it widens the wire arrays of each slab into the float64 / int32 arrays of
    // the two result matrices (expand_rows, a few threads with streaming stores).
    // Wide format: the p-value matrix owns a second copy of the column indices (the two
    // returned matrices share no storage); it is made on the host instead of crossing PCIe
    // twice (single-rank processes only: with several ranks per box the host copies compete
    // for memory bandwidth; CS_HOST_INDEX_COPY overrides).
    bool host_ix2 = A.pval && !narrow;
    if (const char *w = getenv("LOCAL_WORLD_SIZE"))
        if (atoi(w) > 1) host_ix2 = false;
    if (const char *e = getenv("CS_HOST_INDEX_COPY")) host_ix2 = A.pval && !narrow && atoi(e) != 0;
    struct Follower {
        std::vector<cudaEvent_t> ev;          // slab's download has landed
        std::vector<int64_t> off, cnt;
        std::vector<int> r0, r1;
        std::atomic<int> n_enq{0};
        std::atomic<bool> stop{false};
        std::thread th;
        ~Follower() {
            stop.store(true);
            if (th.joinable()) th.join();
            for (auto e : ev) cudaEventDestroy(e);
        }
    } fol;
    if (host_ix2 || narrow) {
        fol.ev.resize(nslab);
        fol.off.assign(nslab, 0);
        fol.cnt.assign(nslab, 0);
        fol.r0.assign(nslab, 0);
        fol.r1.assign(nslab, 0);
        for (int i = 0; i < nslab; ++i)
            CS_CUDA(cudaEventCreateWithFlags(&fol.ev[i], cudaEventDisableTiming));
        const int dev = c->device;
        Follower *fp = &fol;
        const int ns = nslab;
        if (narrow) {
            const float *ws = (const float *)h_ws, *wp = (const float *)h_wp;
            const uint8_t *wo = (const uint8_t *)h_wo;
            const int64_t *ip = (const int64_t *)h_ip;
            double *dd = (double *)h_d, *dp = (double *)h_p;
            int32_t *ix = (int32_t *)h_ix, *ix2 = (int32_t *)h_ix2;
            const int dlo = s->Lo.dlo, nthreads = expand_threads_default();
            fol.th = std::thread([=] {
                cudaSetDevice(dev);
                for (int k = 0; k < ns; ++k) {
                    while (fp->n_enq.load(std::memory_order_acquire) <= k) {
                        if (fp->stop.load()) return;
                        std::this_thread::sleep_for(std::chrono::microseconds(20));
                    }
                    if (fp->cnt[k] == 0) continue;
                    if (cudaEventSynchronize(fp->ev[k]) != cudaSuccess) return;
                    expand_rows(ws, wp, wo, ip, fp->r0[k], fp->r1[k], dlo, dd, dp, ix, ix2, nthreads);
                }
            });
        } else {
            int32_t *src = (int32_t *)h_ix, *dst = (int32_t *)h_ix2;
            fol.th = std::thread([fp, src, dst, dev, ns] {
                cudaSetDevice(dev);
                for (int k = 0; k < ns; ++k) {
                    while (fp->n_enq.load(std::memory_order_acquire) <= k) {
                        if (fp->stop.load()) return;
                        std::this_thread::sleep_for(std::chrono::microseconds(20));
                    }
                    if (fp->cnt[k] == 0) continue;
                    if (cudaEventSynchronize(fp->ev[k]) != cudaSuccess) return;
                    // a few threads: one core does not keep up with the DMA it follows
                    memcpy_mt(dst + fp->off[k], src + fp->off[k], (size_t)fp->cnt[k] * sizeof(int32_t));
                }
            });
        }
    }

    // CS_TRACE=1: per-slab timeline on stderr (host clock and device events, ms from the start)
    const bool trace = getenv("CS_TRACE") != nullptr;
    auto now = [] {
        return std::chrono::duration<double, std::milli>(
                   std::chrono::steady_clock::now().time_since_epoch())
            .count();
    };
    const double t_begin = now();
    std::vector<double> th0(nslab, 0), th1(nslab, 0), tf0(nslab, 0), tf1(nslab, 0);
    std::vector<cudaEvent_t> tev(trace ? 4 * nslab : 0);
    for (auto &e : tev) CS_CUDA(cudaEventCreate(&e));
    struct TevGuard {
        std::vector<cudaEvent_t> &v;
        ~TevGuard() {
            for (auto e : v) cudaEventDestroy(e);
        }
    } tevguard{tev};

    CS_CUDA(cudaEventRecord(s->ev[2], st));
    rc = fill_begin(&s->Li, (float *)s->img.p, A.rows, A.cols, A.has_mask, A.sym_upper,
                    A.max_dist, A.full ? K.kh : 0, A.full ? K.kw : 0, (int32_t *)s->err.p, st);
    if (rc) return rc;
    CS_CUDA(cudaMemsetAsync(s->out.p, 0, (size_t)s->Lo.n_elems * sizeof(float), st));
    if (s->want_nobs)  // missing counts: only windows that have one are written by the kernel
        CS_CUDA(cudaMemsetAsync(s->nobs.p, 0, (size_t)s->Lo.n_elems * (size_t)s->nmiss_bytes, st));
    cs_pearson_opts po;
    session_pearson_opts(s, &po);
    const void *nb = s->want_nobs ? s->nobs.p : nullptr;
    int32_t TRp = 32;
    if ((rc = cs_pearson_tile_rows(&s->Li, &K, &po, s->oy0, s->oy1, s->ox0, s->ox1, s->od_lo,
                                   s->od_hi, &TRp)))
        return rc;
    po.tile_rows = TRp;
    std::vector<int> Yb, need_of;
    plan_slabs(s, TRp, nslab, true, Yb, need_of);
    auto csr_row = [&](int i) { return i == 0 ? 0 : (i == nslab ? A.rows : Yb[i] - s->pr); };
    int up_end = 0;
    int64_t base = 0;
    auto finalize = [&](int k) -> int {
        if (trace) tf0[k] = now() - t_begin;
        CS_CUDA(cudaEventSynchronize(ev_tot[k]));
        if (trace) tf1[k] = now() - t_begin;
        const int64_t nnz_k = tot[k];
        const int cr0 = csr_row(k), cr1 = csr_row(k + 1);
        // on the high-priority stream: ordered after the slab's counts, not after the tiles of
        // the slabs enqueued since
        CS_CUDA(cudaStreamWaitEvent(st_e, ev_tot[k], 0));
        int r = scores_finish_rows(&s->Lo, (int64_t *)s->r_indptr.p, cr0, cr1, base, st_e,
                                   scan_slot(cr0, k));
        if (r) return r;
        if (nnz_k > 0) {
            if (narrow)
                r = scores_emit_rows_narrow(&s->Lo, (const float *)s->out.p, nb, s->nmiss_bytes,
                                            K.kh * K.kw, (const int64_t *)s->r_indptr.p, cr0, cr1,
                                            (float *)s->w_score.p,
                                            A.pval ? (float *)s->w_logp.p : nullptr,
                                            (uint8_t *)s->w_off.p, st_e);
            else
                r = scores_emit_rows(&s->Lo, (const float *)s->out.p, nb, s->nmiss_bytes, K.kh * K.kw,
                                     -(1 << 30), 1 << 30, (const int64_t *)s->r_indptr.p, cr0, cr1,
                                     (int32_t *)s->r_indices.p, (double *)s->r_data.p,
                                     A.pval ? (double *)s->r_p.p : nullptr, st_e);
            if (r) return r;
        }
        CS_CUDA(cudaEventRecord(ev_emit[k], st_e));
        CS_CUDA(cudaStreamWaitEvent(st_d, ev_emit[k], 0));
        if (trace) CS_CUDA(cudaEventRecord(tev[4 * k + 2], st_d));
        const size_t o = (size_t)base, n = (size_t)nnz_k;
        if (narrow) {
            // the slab's final row pointers (the expansion needs them), then its wire arrays
            CS_CUDA(cudaMemcpyAsync((int64_t *)h_ip + cr0, (int64_t *)s->r_indptr.p + cr0,
                                    (size_t)(cr1 - cr0 + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                    st_d));
            if (nnz_k > 0) {
                CS_CUDA(cudaMemcpyAsync((float *)h_ws + o, (float *)s->w_score.p + o, n * sizeof(float),
                                        cudaMemcpyDeviceToHost, st_d));
                CS_CUDA(cudaMemcpyAsync((uint8_t *)h_wo + o, (uint8_t *)s->w_off.p + o, n,
                                        cudaMemcpyDeviceToHost, st_d));
                if (A.pval)
                    CS_CUDA(cudaMemcpyAsync((float *)h_wp + o, (float *)s->w_logp.p + o,
                                            n * sizeof(float), cudaMemcpyDeviceToHost, st_d));
                CS_CUDA(cudaEventRecord(fol.ev[k], st_d));
            }
            fol.off[k] = (int64_t)o;
            fol.cnt[k] = (int64_t)n;
            fol.r0[k] = cr0;
            fol.r1[k] = cr1;
        } else if (nnz_k > 0) {
            CS_CUDA(cudaMemcpyAsync((int32_t *)h_ix + o, (int32_t *)s->r_indices.p + o,
                                    n * sizeof(int32_t), cudaMemcpyDeviceToHost, st_d));
            if (host_ix2) {
                CS_CUDA(cudaEventRecord(fol.ev[k], st_d));
                fol.off[k] = (int64_t)o;
                fol.cnt[k] = (int64_t)n;
            }
            CS_CUDA(cudaMemcpyAsync((double *)h_d + o, (double *)s->r_data.p + o, n * sizeof(double),
                                    cudaMemcpyDeviceToHost, st_d));
            if (A.pval) {
                CS_CUDA(cudaMemcpyAsync((double *)h_p + o, (double *)s->r_p.p + o,
                                        n * sizeof(double), cudaMemcpyDeviceToHost, st_d));
                if (!host_ix2)
                    CS_CUDA(cudaMemcpyAsync((int32_t *)h_ix2 + o, (int32_t *)s->r_indices.p + o,
                                            n * sizeof(int32_t), cudaMemcpyDeviceToHost, st_d));
            }
        }
        if (host_ix2 || narrow) fol.n_enq.store(k + 1, std::memory_order_release);
        if (trace) CS_CUDA(cudaEventRecord(tev[4 * k + 3], st_d));
        base += nnz_k;
        return CS_OK;
    };
    // slabs whose totals are known are finalized (row pointers, CSR entries, download) as soon
    // as possible: between the staging chunks of later slabs and after every enqueue
    int n_counted = 0, n_final = 0;
    auto poll = [&]() -> int {
        while (n_final < n_counted && cudaEventQuery(ev_tot[n_final]) == cudaSuccess) {
            if (int r = finalize(n_final)) return r;
            ++n_final;
        }
        return CS_OK;
    };
    // the uploads run ahead on a thread of their own; the enqueueing thread only waits for the
    // rows a slab needs and keeps finalizing finished slabs meanwhile
    SlabUploader upl;
    st_h = c->st_up;  // always a stream of its own
    upl.start(c, st_h, a, s, need_of, ev_up.data(), trace);
    for (int k = 0; k < nslab; ++k) {
        const int Y0 = Yb[k], Y1 = Yb[k + 1];
        if (trace) {
            th0[k] = now() - t_begin;
            CS_CUDA(cudaEventRecord(tev[4 * k], st));
        }
        const int need = need_of[k];
        // wait for the slab's rows to be on their way, finalizing finished slabs meanwhile
        while (upl.done.load(std::memory_order_acquire) <= k) {
            if (upl.rc.load()) {
                set_error("upload of the input failed");
                return upl.rc.load();
            }
            if ((rc = poll())) return rc;
            std::this_thread::sleep_for(std::chrono::microseconds(10));
        }
        if (need > up_end) {
            CS_CUDA(cudaStreamWaitEvent(st, ev_up[k], 0));
            rc = fill_rows(&s->Li, (float *)s->img.p, (const int64_t *)s->sig_indptr.p,
                           (const int32_t *)s->sig_indices.p, (const double *)s->sig_data.p, A.rows,
                           up_end, need, s->pr, s->pc, A.has_mask,
                           (const int64_t *)s->m_indptr.p, (const int32_t *)s->m_indices.p,
                           &s->geo, A.sym_upper, A.max_dist, A.full ? K.kh : 0,
                           A.full ? K.kw : 0, (int32_t *)s->err.p, st);
            if (rc) return rc;
            up_end = need;
        }
        if (Y1 > Y0) {
            rc = cs_pearson_f32(&s->Li, (const float *)s->img.p, &K, &po, Y0, Y1, s->ox0, s->ox1,
                                s->od_lo, s->od_hi, &s->Lo, (float *)s->out.p,
                                s->want_nobs ? s->nobs.p : nullptr, st);
            if (rc) return rc;
        }
        // the slab total is written straight into pinned host memory (UVA): a D2H copy on this
        // stream would queue behind the bulk downloads of the copy stream
        tot[k] = 0;
        rc = scores_count_rows(&s->Lo, (const float *)s->out.p, -(1 << 30), 1 << 30,
                               (int64_t *)s->r_indptr.p, csr_row(k), csr_row(k + 1), tot + k, st,
                               scan_slot(csr_row(k), k));
        if (rc) return rc;
        CS_CUDA(cudaEventRecord(ev_tot[k], st));
        n_counted = k + 1;
        if (trace) {
            CS_CUDA(cudaEventRecord(tev[4 * k + 1], st));
            th1[k] = now() - t_begin;
        }
        if ((rc = poll())) return rc;
    }
    for (; n_final < nslab; ++n_final)
        if ((rc = finalize(n_final))) return rc;
    int32_t *herr = (int32_t *)(tot + nslab);  // last slot of the pinned totals block
    CS_CUDA(cudaMemcpyAsync(herr, s->err.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    // row pointers: final once every slab has been finalized (the copy stream follows them
    // all); the narrow path has downloaded them slab by slab
    if (!narrow) {
        CS_CUDA(cudaMemcpyAsync(h_ip, s->r_indptr.p, n_ip * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                st_d));
        if (A.pval)
            CS_CUDA(cudaMemcpyAsync(h_ip2, s->r_indptr.p, n_ip * sizeof(int64_t),
                                    cudaMemcpyDeviceToHost, st_d));
    }
    // the span ends when the last download has landed
    CS_CUDA(cudaEventRecord(ev_emit[0], st_d));
    CS_CUDA(cudaStreamWaitEvent(st, ev_emit[0], 0));
    CS_CUDA(cudaEventRecord(s->ev[5], st));
    CS_CUDA(cudaStreamSynchronize(st));
    CS_CUDA(cudaStreamSynchronize(st_d));
    if (narrow && A.pval) memcpy(h_ip2, h_ip, n_ip * sizeof(int64_t));
    if (fol.th.joinable()) fol.th.join();  // the last slab's expansion / index copy
    if (trace) {
        fprintf(stderr, "slab  host:enq0  enq1  fin_wait0 fin_wait1 | dev:compute0 compute1 d2h0 d2h1 (ms)\n");
        for (int k = 0; k < nslab; ++k) {
            float c0 = 0, c1 = 0, d0 = 0, d1 = 0;
            cudaEventElapsedTime(&c0, s->ev[2], tev[4 * k]);
            cudaEventElapsedTime(&c1, s->ev[2], tev[4 * k + 1]);
            cudaEventElapsedTime(&d0, s->ev[2], tev[4 * k + 2]);
            cudaEventElapsedTime(&d1, s->ev[2], tev[4 * k + 3]);
            fprintf(stderr, "%4d  %8.2f %8.2f %8.2f %8.2f | %8.2f %8.2f %8.2f %8.2f  nnz %lld\n", k,
                    th0[k], th1[k], tf0[k], tf1[k], c0, c1, d0, d1, (long long)tot[k]);
        }
        fprintf(stderr, "host total %.2f ms (plan + row pointers before it: %.2f ms)\n",
                now() - t_begin, ms_plan);
    }
    if (herr[0] > 0 && A.has_mask) {
        set_error("There are %d non-zero elements reported as missing.", herr[0]);
        return CS_ERR_MASKED_SIGNAL;
    }
    if (herr[1] > 0 && !A.trim_to_max_dist) {
        set_error("internal: %d signal pixels fell outside the stored band", herr[1]);
        return CS_ERR_INVALID;
    }
    s->nnz_out = base;
    s->ran = true;
    s->refined = false;
    s->compacted = false;  // the pipeline's CSR arrays are not the resident session's
    res->nnz = base;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev[2], s->ev[5]);
    res->ms_kernels = ms;  // upload, kernels and most of the download overlap inside this span
    res->h2d_bytes = (int64_t)s->h2d_bytes;
    if (narrow)
        res->d2h_bytes = (int64_t)(n_ip * sizeof(int64_t) +
                                   (size_t)base * (sizeof(float) + 1 + (A.pval ? sizeof(float) : 0)));
    else
        res->d2h_bytes = (int64_t)(n_ip * sizeof(int64_t) * (A.pval ? 2 : 1) +
                                   (size_t)base * (sizeof(int32_t) * ((A.pval && !host_ix2) ? 2 : 1) +
                                                   sizeof(double) * (A.pval ? 2 : 1)));
    return CS_OK;
}

// One-shot: upload + run + download on a per-device cached session.
extern "C" int cs_normxcorr2_host(const cs_normxcorr2_args *a, cs_csr_result *res) {
    CS_REQUIRE(a && res, "cs_normxcorr2_host: null argument");
    CS_REQUIRE(!a->device_payload, "cs_normxcorr2_host takes host arrays (device payload: use a session)");
    static std::mutex mu;
    static std::vector<cs_session *> cache;
    cs_session *s = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (cs_session *x : cache)
            if (x->c->device == a->device) s = x;
        if (!s) {
            int rc = cs_session_create(a->device, &s);
            if (rc) return rc;
            cache.push_back(s);
        }
    }
    // the cached session holds the inputs and results of ONE call: concurrent callers on the
    // same device take turns for the whole upload / run / download sequence
    std::lock_guard<std::mutex> call_lk(s->call_mu);
    // big calls go through the slab pipeline
    bool uploaded = false;
    {
        long long rows_out = a->full ? a->rows : a->rows - (a->kernel.kh - 1);
        const int64_t nnz_in = a->indptr ? a->indptr[a->rows] : 0;
        int nslab = 12;
        if (const char *e = getenv("CS_PIPELINE_SLABS")) nslab = atoi(e);
        if (nslab > 1 && nnz_in >= (4 << 20) && rows_out >= 64 * nslab) {
            int rc = normxcorr2_pipelined(s, a, res, nslab);
            if (rc != 1) {
                if (rc) cs_result_free(res);
                return rc;
            }
            uploaded = true;
        }
    }
    int rc = uploaded ? CS_OK : cs_session_upload(s, a);
    if (rc) return rc;
    float ms_h2d = 0.f;
    cs_run_stats stt;
    rc = cs_session_run(s, &stt);
    if (rc) return rc;
    if (!s->empty) cudaEventElapsedTime(&ms_h2d, s->ev[0], s->ev[1]);
    rc = cs_session_download(s, res);
    if (rc) return rc;
    res->ms_h2d = ms_h2d;
    res->ms_kernels = stt.ms_total;
    res->h2d_bytes = stt.h2d_bytes;
    return CS_OK;
}

// validate_patterns (det:18-155) + score / p-value lookup on the session's matrix.
extern "C" int cs_session_validate(cs_session *s, const int32_t *host_coords, int64_t n_coords,
                                   const uint8_t *host_valid_row, const uint8_t *host_valid_col,
                                   int32_t inter, double zero_tol, double missing_tol,
                                   int32_t score_dmax, double *host_windows, uint8_t *host_valid,
                                   double *host_score, double *host_log10p) {
    CS_REQUIRE(s && s->uploaded && s->ran, "cs_session_validate: upload and run first");
    CS_REQUIRE(n_coords >= 0 && (n_coords == 0 || (host_coords && host_windows && host_valid)),
               "cs_session_validate: null argument");
    if (n_coords == 0) return CS_OK;
    HostCtx *c = s->c;
    std::lock_guard<std::mutex> lk(c->mu);
    CS_CUDA(cudaSetDevice(c->device));
    if (int src = session_settle(s)) return src;
    cudaStream_t st = s->stream();
    const cs_normxcorr2_args &a = s->a;
    const int km = a.kernel.kh, kn = a.kernel.kw;
    const int kh = (km - 1) / 2, kw = (kn - 1) / 2;
    const size_t npx = (size_t)km * kn;
    int rc;
    if ((rc = s->g_coords.ensure((size_t)n_coords * 4 * sizeof(int32_t)))) return rc;
    if ((rc = s->g_win.ensure((size_t)n_coords * npx * sizeof(double)))) return rc;
    if ((rc = s->g_flag.ensure((size_t)n_coords))) return rc;
    if ((rc = s->g_score.ensure((size_t)n_coords * sizeof(double)))) return rc;
    if ((rc = s->g_p.ensure((size_t)n_coords * sizeof(double)))) return rc;
    if ((rc = s->g_vrow.ensure((size_t)a.rows))) return rc;
    if ((rc = s->g_vcol.ensure((size_t)a.cols))) return rc;
    // padded coordinates (det:291-298; zero_pad_sparse(mat, kh, kw) pads kw rows and kh columns,
    // the coordinates are shifted by (kh, kw) -- identical for square kernels) and the
    // coordinates at which the padded correlation map is read (det:134)
    std::vector<int32_t> hc((size_t)n_coords * 4);
    const int sr = a.full ? kh : 0, sc = a.full ? kw : 0;
    const int pad_r = a.full ? kw : 0, pad_c = a.full ? kh : 0;
    for (int64_t i = 0; i < n_coords; ++i) {
        const int32_t c1 = host_coords[2 * i], c2 = host_coords[2 * i + 1];
        hc[2 * i] = c1 + sr;
        hc[2 * i + 1] = c2 + sc;
        hc[(size_t)2 * n_coords + 2 * i] = c1 + sr - pad_r;
        hc[(size_t)2 * n_coords + 2 * i + 1] = c2 + sc - pad_c;
    }
    CS_CUDA(cudaMemcpyAsync(s->g_coords.p, hc.data(), hc.size() * sizeof(int32_t),
                            cudaMemcpyHostToDevice, st));
    if (host_valid_row)
        CS_CUDA(cudaMemcpyAsync(s->g_vrow.p, host_valid_row, (size_t)a.rows, cudaMemcpyHostToDevice, st));
    if (host_valid_col)
        CS_CUDA(cudaMemcpyAsync(s->g_vcol.p, host_valid_col, (size_t)a.cols, cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaStreamSynchronize(st));  // hc is about to go out of scope
    cs_gather_args g;
    memset(&g, 0, sizeof(g));
    g.rows = a.rows;
    g.cols = a.cols;
    g.d_indptr = (const int64_t *)s->sig_indptr.p;
    g.d_indices = (const int32_t *)s->sig_indices.p;
    g.d_data = (const double *)s->sig_data.p;
    g.d_valid_row = host_valid_row ? (const uint8_t *)s->g_vrow.p : nullptr;
    g.d_valid_col = host_valid_col ? (const uint8_t *)s->g_vcol.p : nullptr;
    g.win_h = km;
    g.win_w = kn;
    g.pad_rows = pad_r;
    g.pad_cols = pad_c;
    g.det_shift_row = sr;
    g.det_shift_col = sc;
    g.nan_subdiag = inter ? 0 : (km > kn ? km : kn);
    g.zero_tol = zero_tol;
    g.missing_tol = missing_tol;
    const int32_t *d_pad = (const int32_t *)s->g_coords.p;
    const int32_t *d_conv = d_pad + (size_t)2 * n_coords;
    rc = cs_window_gather(&g, d_pad, n_coords, (double *)s->g_win.p, (uint8_t *)s->g_flag.p, st);
    if (rc) return rc;
    CS_CUDA(cudaMemcpyAsync(host_windows, s->g_win.p, (size_t)n_coords * npx * sizeof(double),
                            cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaMemcpyAsync(host_valid, s->g_flag.p, (size_t)n_coords, cudaMemcpyDeviceToHost, st));
    if (host_score || host_log10p) {
        if (s->empty) {
            CS_CUDA(cudaStreamSynchronize(st));
            for (int64_t i = 0; i < n_coords; ++i) {
                if (host_score) host_score[i] = 0.0;
                if (host_log10p) host_log10p[i] = 0.0;
            }
            return CS_OK;
        }
        const void *nb = s->want_nobs ? s->nobs.p : nullptr;
        // scores of the trimmed map (det:270) at the padded-map coordinates
        rc = cs_scores_lookup(&s->Lo, (const float *)s->out.p, nb, s->nmiss_bytes, km * kn, inter ? -(1 << 30) : 0,
                              inter ? (1 << 30) : score_dmax, d_conv, n_coords,
                              (double *)s->g_score.p, nullptr, st);
        if (rc) return rc;
        if (host_score)
            CS_CUDA(cudaMemcpyAsync(host_score, s->g_score.p, (size_t)n_coords * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
        if (host_log10p) {
            // p-values of the untrimmed map at the final coordinates (det:337-339): upload them
            // over the padded ones (already consumed)
            CS_CUDA(cudaStreamSynchronize(st));
            CS_CUDA(cudaMemcpyAsync(s->g_coords.p, host_coords, (size_t)n_coords * 2 * sizeof(int32_t),
                                    cudaMemcpyHostToDevice, st));
            rc = cs_scores_lookup(&s->Lo, (const float *)s->out.p, nb, s->nmiss_bytes, km * kn, -(1 << 30), 1 << 30,
                                  d_pad, n_coords, (double *)s->g_score.p, (double *)s->g_p.p, st);
            if (rc) return rc;
            CS_CUDA(cudaMemcpyAsync(host_log10p, s->g_p.p, (size_t)n_coords * sizeof(double),
                                    cudaMemcpyDeviceToHost, st));
        }
    }
    CS_CUDA(cudaStreamSynchronize(st));
    return CS_OK;
}
