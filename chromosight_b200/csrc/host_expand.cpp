// Host side of the narrow result format: float32 scores / log10 p-values and uint8 diagonal
// offsets (9 B per stored score over PCIe) -> the float64 data arrays and int32 column indices
// of the two scipy CSR matrices normxcorr2 returns (det:1098-1131), 24 B per stored score.
// The expansion runs on a few threads with non-temporal AVX-512 stores where the CPU has them
// (measured on the B200 boxes, scripts/ubench/expand_bw.c: 8 threads write 120 GB/s, a
// 46 M-entry result in 9 ms -- the PCIe copy of the wide format takes 18 ms).
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace cs {

// A small persistent worker pool: spawning threads per slab costs more than the slab's work
// (a new thread needs a good part of a millisecond before it runs).  Callers from several
// threads share it; parallel_for runs one share on the calling thread.
class WorkerPool {
public:
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return (int)th_.size(); }
    void parallel_for(int n, const std::function<void(int)> &fn) {
        if (n <= 1) {
            if (n == 1) fn(0);
            return;
        }
        // completion state shared with the tasks (kept alive by them: the last worker may still
        // be inside notify when the caller wakes up)
        struct Done {
            std::mutex mu;
            std::condition_variable cv;
            int left;
        };
        auto done = std::make_shared<Done>();
        done->left = n - 1;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int i = 1; i < n; ++i)
                q_.emplace_back([done, &fn, i] {
                    fn(i);
                    std::lock_guard<std::mutex> dl(done->mu);
                    if (--done->left == 0) done->cv.notify_one();
                });
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> dl(done->mu);
        done->cv.wait(dl, [&] { return done->left == 0; });
    }

private:
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
        }
    }
    std::vector<std::thread> th_;
    std::deque<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

int expand_threads_default();
int upload_threads_default();

// two pools, so that the staging copies of the upload (on the critical path of the device)
// never queue behind the widening of results (leaked on purpose: worker threads must not be
// joined from a static destructor at exit)
static WorkerPool &pool(int which) {
    static WorkerPool *p[2] = {new WorkerPool(std::max(expand_threads_default(), 2) - 1),
                               new WorkerPool(std::max(upload_threads_default(), 2) - 1)};
    return *p[which];
}

void parallel_for(int n, const std::function<void(int)> &fn, int which) {
    pool(which).parallel_for(n, fn);
}

static void expand_scalar(const float *score, const float *log10p, const uint8_t *off,
                          const int64_t *indptr, int32_t r0, int32_t r1, int32_t dlo,
                          double *data, double *logp, int32_t *indices, int32_t *indices2) {
    for (int64_t k = indptr[r0]; k < indptr[r1]; ++k) data[k] = (double)score[k];
    if (logp)
        for (int64_t k = indptr[r0]; k < indptr[r1]; ++k) logp[k] = (double)log10p[k];
    for (int32_t r = r0; r < r1; ++r) {
        const int32_t base = r + dlo;
        for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
            const int32_t c = base + (int32_t)off[k];
            indices[k] = c;
            if (indices2) indices2[k] = c;
        }
    }
}

#if defined(__x86_64__)
#define CS_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq")))
// float32 -> float64 over entries [k0, k1) with streaming stores (64-byte aligned destinations;
// the unaligned head and the tail go through plain stores)
CS_AVX512 static void widen_avx512(const float *src, double *dst, int64_t k0, int64_t k1) {
    int64_t k = k0;
    while (k < k1 && ((uintptr_t)(dst + k) & 63)) {
        dst[k] = (double)src[k];
        ++k;
    }
    for (; k + 16 <= k1; k += 16) {
        const __m512 v = _mm512_loadu_ps(src + k);
        _mm512_stream_pd(dst + k, _mm512_cvtps_pd(_mm512_castps512_ps256(v)));
        _mm512_stream_pd(dst + k + 8, _mm512_cvtps_pd(_mm512_extractf32x8_ps(v, 1)));
    }
    for (; k < k1; ++k) dst[k] = (double)src[k];
}

CS_AVX512 static void expand_avx512(const float *score, const float *log10p, const uint8_t *off,
                          const int64_t *indptr, int32_t r0, int32_t r1, int32_t dlo,
                          double *data, double *logp, int32_t *indices, int32_t *indices2) {
    const int64_t k0 = indptr[r0], k1 = indptr[r1];
    widen_avx512(score, data, k0, k1);
    if (logp) widen_avx512(log10p, logp, k0, k1);
    // indices: per row (col = row + dlo + offset).  A masked head up to the next 64-byte
    // boundary of the destination, streaming stores over the aligned middle, a masked tail.
    const bool same_align = indices2 == nullptr || ((((uintptr_t)indices) ^ ((uintptr_t)indices2)) & 63) == 0;
    for (int32_t r = r0; r < r1; ++r) {
        const __m512i base = _mm512_set1_epi32(r + dlo);
        int64_t k = indptr[r];
        const int64_t e = indptr[r + 1];
        if (k >= e) continue;
        int head = (int)((16 - (((uintptr_t)(indices + k) >> 2) & 15)) & 15);
        if (head > e - k) head = (int)(e - k);
        if (head) {
            const __mmask16 m = (__mmask16)((1u << head) - 1u);
            const __m512i c = _mm512_add_epi32(
                _mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(m, (const void *)(off + k))), base);
            _mm512_mask_storeu_epi32((void *)(indices + k), m, c);
            if (indices2) _mm512_mask_storeu_epi32((void *)(indices2 + k), m, c);
            k += head;
        }
        for (; k + 16 <= e; k += 16) {
            const __m512i c = _mm512_add_epi32(
                _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(off + k))), base);
            _mm512_stream_si512((__m512i *)(indices + k), c);
            if (indices2) {
                if (same_align) _mm512_stream_si512((__m512i *)(indices2 + k), c);
                else _mm512_storeu_si512((void *)(indices2 + k), c);
            }
        }
        if (k < e) {  // the row's tail, one masked vector
            const __mmask16 m = (__mmask16)((1u << (int)(e - k)) - 1u);
            const __m512i c = _mm512_add_epi32(
                _mm512_cvtepu8_epi32(_mm_maskz_loadu_epi8(m, (const void *)(off + k))), base);
            _mm512_mask_storeu_epi32((void *)(indices + k), m, c);
            if (indices2) _mm512_mask_storeu_epi32((void *)(indices2 + k), m, c);
        }
    }
    _mm_sfence();
}
#endif

int expand_threads_default() {
    static const int n = [] {
        if (const char *e = getenv("CS_EXPAND_THREADS"))
            if (atoi(e) > 0) return atoi(e);
        int hw = (int)std::thread::hardware_concurrency();
        if (hw <= 0) hw = 4;
        int ranks = 1;  // the ranks of one box share its cores
        if (const char *w = getenv("LOCAL_WORLD_SIZE"))
            if (atoi(w) > 1) ranks = atoi(w);
        int t = hw / (2 * ranks);
        return std::max(1, std::min(t, 8));
    }();
    return n;
}

#if defined(__x86_64__)
CS_AVX512 static void copy_stream_avx512(void *dst, const void *src, size_t n) {
    char *d = (char *)dst;
    const char *s = (const char *)src;
    size_t k = 0;
    while (k < n && ((uintptr_t)(d + k) & 63)) {
        d[k] = s[k];
        ++k;
    }
    for (; k + 256 <= n; k += 256) {
        const __m512i a = _mm512_loadu_si512((const void *)(s + k));
        const __m512i b = _mm512_loadu_si512((const void *)(s + k + 64));
        const __m512i c = _mm512_loadu_si512((const void *)(s + k + 128));
        const __m512i e = _mm512_loadu_si512((const void *)(s + k + 192));
        _mm512_stream_si512((__m512i *)(d + k), a);
        _mm512_stream_si512((__m512i *)(d + k + 64), b);
        _mm512_stream_si512((__m512i *)(d + k + 128), c);
        _mm512_stream_si512((__m512i *)(d + k + 192), e);
    }
    for (; k < n; ++k) d[k] = s[k];
    _mm_sfence();
}
#endif

// memcpy whose stores bypass the cache: the destination (a pinned staging buffer) is read next
// by the DMA engine, not by a CPU
void copy_stream(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    if (has512 && n >= 4096) {
        copy_stream_avx512(dst, src, n);
        return;
    }
#endif
    memcpy(dst, src, n);
}

int upload_threads_default() {
    static const int n = [] {
        if (const char *e = getenv("CS_COPY_THREADS"))
            if (atoi(e) > 0) return atoi(e);
        int hw = (int)std::thread::hardware_concurrency();
        if (hw <= 0) hw = 4;
        int ranks = 1;
        if (const char *w = getenv("LOCAL_WORLD_SIZE"))
            if (atoi(w) > 1) ranks = atoi(w);
        return std::max(1, std::min(hw / (4 * ranks), 4));
    }();
    return n;
}

void expand_rows(const float *score, const float *log10p, const uint8_t *off,
                 const int64_t *indptr, int32_t r0, int32_t r1, int32_t dlo, double *data,
                 double *logp, int32_t *indices, int32_t *indices2, int threads) {
    if (r1 <= r0 || indptr[r1] <= indptr[r0]) return;
#if defined(__x86_64__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                               __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq");
#else
    static const bool has512 = false;
#endif
    auto run = [&](int32_t a, int32_t b) {
#if defined(__x86_64__)
        if (has512) {
            expand_avx512(score, log10p, off, indptr, a, b, dlo, data, logp, indices, indices2);
            return;
        }
#endif
        expand_scalar(score, log10p, off, indptr, a, b, dlo, data, logp, indices, indices2);
    };
    const int64_t n = indptr[r1] - indptr[r0];
    int nt = (int)std::min<int64_t>(threads, n / (1 << 18) + 1);
    if (nt <= 1) {
        run(r0, r1);
        return;
    }
    // row ranges of about equal entry counts
    std::vector<int32_t> cut(nt + 1, r1);
    cut[0] = r0;
    for (int t = 1; t < nt; ++t) {
        const int64_t want = indptr[r0] + n * t / nt;
        cut[t] = (int32_t)(std::lower_bound(indptr + r0, indptr + r1, want) - indptr);
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    parallel_for(nt, [&](int t) {
        if (cut[t + 1] > cut[t]) run(cut[t], cut[t + 1]);
    }, 0);
}

// Upper band of one chromosome straight out of a .cool pixel table (sorted by bin1, then bin2):
// the pixels with s <= bin1 < e (n_pix of them, given from the first one on), bin2 < e and
// bin2 - bin1 <= max_diag become canonical CSR rows; with `weight` the values are balanced
// (count * w[bin1] * w[bin2]) and pixels on masked bins (non-finite value) are dropped.
// Two parallel passes over contiguous row ranges: count per row, then write.
template <typename T>
static int64_t band_csr_impl(const int64_t *bin1, const int64_t *bin2, const T *count, int64_t n_pix,
                             const double *weight, int64_t s, int64_t e, int64_t max_diag,
                             int64_t *indptr, int32_t *indices, double *data, int threads) {
    const int64_t nrows = e - s;
    if (nrows <= 0) return 0;
    int nt = (int)std::min<int64_t>(std::max(threads, 1), std::max<int64_t>(1, n_pix / (1 << 16)));
    std::vector<int64_t> pcut(nt + 1, n_pix), rcut(nt + 1, nrows);
    pcut[0] = 0;
    rcut[0] = 0;
    for (int t = 1; t < nt; ++t) {
        // cut between rows: first pixel of the row that holds pixel n_pix * t / nt
        int64_t p = n_pix * t / nt;
        const int64_t row = bin1[p];
        p = std::lower_bound(bin1, bin1 + n_pix, row) - bin1;
        pcut[t] = std::max(p, pcut[t - 1]);
        rcut[t] = std::max(row - s, rcut[t - 1]);
    }
    auto keep = [&](int64_t k, double &v) {
        const int64_t b1 = bin1[k], b2 = bin2[k];
        if (b2 >= e || b2 - b1 > max_diag || b2 < b1) return false;
        v = (double)count[k];
        if (weight) {
            v = v * weight[b1] * weight[b2];
            if (!std::isfinite(v)) return false;
        }
        return true;
    };
    indptr[0] = 0;
    parallel_for(nt, [&](int t) {
        for (int64_t r = rcut[t]; r < rcut[t + 1]; ++r) indptr[r + 1] = 0;
        for (int64_t k = pcut[t]; k < pcut[t + 1]; ++k) {
            double v;
            if (keep(k, v)) ++indptr[bin1[k] - s + 1];
        }
    }, 0);
    for (int64_t r = 0; r < nrows; ++r) indptr[r + 1] += indptr[r];
    parallel_for(nt, [&](int t) {
        int64_t row = -1, o = 0;
        for (int64_t k = pcut[t]; k < pcut[t + 1]; ++k) {
            double v;
            if (!keep(k, v)) continue;
            const int64_t r = bin1[k] - s;
            if (r != row) {
                row = r;
                o = indptr[r];
            }
            indices[o] = (int32_t)(bin2[k] - s);
            data[o] = v;
            ++o;
        }
    }, 0);
    return indptr[nrows];
}

int64_t band_csr_from_pixels(const int64_t *bin1, const int64_t *bin2, const void *count, int count_dtype,
                             int64_t n_pix, const double *weight, int64_t s, int64_t e, int64_t max_diag,
                             int64_t *indptr, int32_t *indices, double *data, int threads) {
    switch (count_dtype) {
        case 0: return band_csr_impl(bin1, bin2, (const int32_t *)count, n_pix, weight, s, e, max_diag, indptr, indices, data, threads);
        case 1: return band_csr_impl(bin1, bin2, (const int64_t *)count, n_pix, weight, s, e, max_diag, indptr, indices, data, threads);
        case 2: return band_csr_impl(bin1, bin2, (const double *)count, n_pix, weight, s, e, max_diag, indptr, indices, data, threads);
        default: return -1;
    }
}

// 1 when the pixel table is sorted by (bin1, bin2) without duplicates (cooler's invariant)
int pixels_lex_sorted(const int64_t *bin1, const int64_t *bin2, int64_t n, int threads) {
    if (n < 2) return 1;
    const int nt = (int)std::min<int64_t>(std::max(threads, 1), std::max<int64_t>(1, n / (1 << 18)));
    std::vector<int> okv(nt, 1);
    parallel_for(nt, [&](int t) {
        const int64_t a = 1 + (n - 1) * t / nt, b = 1 + (n - 1) * (t + 1) / nt;
        int ok = 1;
        for (int64_t k = a; k < b; ++k) {
            const int64_t d1 = bin1[k] - bin1[k - 1];
            ok &= (d1 > 0) | ((d1 == 0) & (bin2[k] > bin2[k - 1]));
        }
        okv[t] = ok;
    }, 0);
    for (int v : okv)
        if (!v) return 0;
    return 1;
}

// The inter-chromosomal pixels grouped by (chromosome of bin1, chromosome of bin2): a stable
// counting sort by block key c1 * C + c2.  bin_chrom: chromosome of every bin.  Writes the pixel
// indices, block after block and in table order inside a block, to `order` (room for n) and the
// block offsets to starts[C * C + 1]; returns the number of inter pixels.
int64_t pixels_inter_index(const int64_t *bin1, const int64_t *bin2, int64_t n, const int16_t *bin_chrom,
                           int32_t C, int64_t *order, int64_t *starts, int threads) {
    const int nb = C * C;
    const int nt = (int)std::min<int64_t>(std::max(threads, 1), std::max<int64_t>(1, n / (1 << 18)));
    std::vector<std::vector<int64_t>> hist(nt, std::vector<int64_t>(nb, 0));
    parallel_for(nt, [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        std::vector<int64_t> &h = hist[t];
        for (int64_t k = a; k < b; ++k) {
            const int c1 = bin_chrom[bin1[k]], c2 = bin_chrom[bin2[k]];
            if (c1 != c2) ++h[c1 * C + c2];
        }
    }, 0);
    // offsets: block-major, then thread (keeps table order inside a block)
    int64_t run = 0;
    for (int b = 0; b < nb; ++b) {
        starts[b] = run;
        for (int t = 0; t < nt; ++t) {
            const int64_t c = hist[t][b];
            hist[t][b] = run;
            run += c;
        }
    }
    starts[nb] = run;
    parallel_for(nt, [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        std::vector<int64_t> &h = hist[t];
        for (int64_t k = a; k < b; ++k) {
            const int c1 = bin_chrom[bin1[k]], c2 = bin_chrom[bin2[k]];
            if (c1 != c2) order[h[c1 * C + c2]++] = k;
        }
    }, 0);
    return run;
}

}  // namespace cs
