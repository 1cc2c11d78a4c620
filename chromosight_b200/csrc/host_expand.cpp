// Host side of the narrow result format: float32 scores / log10 p-values and uint8 diagonal
// offsets (9 B per stored score over PCIe) -> the float64 data arrays and int32 column indices
// of the two scipy CSR matrices normxcorr2 returns (det:1098-1131), 24 B per stored score.
// The expansion runs on a few threads with non-temporal AVX-512 stores where the CPU has them
// (measured on the B200 boxes, scripts/ubench/expand_bw.c: 8 threads write 120 GB/s, a
// 46 M-entry result in 9 ms -- the PCIe copy of the wide format takes 18 ms).
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <thread>
#include <vector>

namespace cs {

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
    // indices: per row (col = row + dlo + offset)
    for (int32_t r = r0; r < r1; ++r) {
        const __m512i base = _mm512_set1_epi32(r + dlo);
        int64_t k = indptr[r];
        const int64_t e = indptr[r + 1];
        for (; k + 16 <= e; k += 16) {
            const __m512i c = _mm512_add_epi32(
                _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(off + k))), base);
            _mm512_storeu_si512((void *)(indices + k), c);
            if (indices2) _mm512_storeu_si512((void *)(indices2 + k), c);
        }
        for (; k < e; ++k) {
            const int32_t c = r + dlo + (int32_t)off[k];
            indices[k] = c;
            if (indices2) indices2[k] = c;
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
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t)
        if (cut[t + 1] > cut[t]) th.emplace_back(run, cut[t], cut[t + 1]);
    run(cut[0], cut[1]);
    for (auto &x : th) x.join();
}

}  // namespace cs
