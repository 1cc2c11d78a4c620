// Host micro-benchmark: how fast can T threads expand the narrow wire format of the result
// (float32 score, float32 log10 p, uint8 diagonal offset) into what scipy wants (float64 data
// x2, int32 indices x2)?  Decides whether the D2H of cs_normxcorr2_host ships 9 B or 20 B per
// stored score.   gcc -O3 -march=native -pthread expand_bw.c -o expand_bw
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static size_t N;
static float *sc, *lp;
static uint8_t *off;
static double *d1, *d2;
static int32_t *i1, *i2;
static int mode;  // 0 plain stores, 1 non-temporal

typedef struct { size_t a, b; } range_t;

static void *work(void *arg) {
    range_t *r = (range_t *)arg;
    size_t k = r->a, e = r->b;
#ifdef __AVX512F__
    if (mode == 1) {
        for (; k + 16 <= e; k += 16) {
            __m512 s = _mm512_loadu_ps(sc + k), p = _mm512_loadu_ps(lp + k);
            _mm512_stream_pd(d1 + k, _mm512_cvtps_pd(_mm512_castps512_ps256(s)));
            _mm512_stream_pd(d1 + k + 8, _mm512_cvtps_pd(_mm512_extractf32x8_ps(s, 1)));
            _mm512_stream_pd(d2 + k, _mm512_cvtps_pd(_mm512_castps512_ps256(p)));
            _mm512_stream_pd(d2 + k + 8, _mm512_cvtps_pd(_mm512_extractf32x8_ps(p, 1)));
            __m512i o = _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(off + k)));
            o = _mm512_add_epi32(o, _mm512_set1_epi32((int)(k >> 8)));
            _mm512_stream_si512((__m512i *)(i1 + k), o);
            _mm512_stream_si512((__m512i *)(i2 + k), o);
        }
        _mm_sfence();
    }
#endif
    for (; k < e; ++k) {
        d1[k] = sc[k];
        d2[k] = lp[k];
        int32_t v = (int32_t)off[k] + (int32_t)(k >> 8);
        i1[k] = v;
        i2[k] = v;
    }
    return NULL;
}

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv) {
    N = (argc > 1 ? atol(argv[1]) : 46000000);
    N &= ~(size_t)63;
    sc = aligned_alloc(64, N * 4); lp = aligned_alloc(64, N * 4); off = aligned_alloc(64, N);
    d1 = aligned_alloc(64, N * 8); d2 = aligned_alloc(64, N * 8);
    i1 = aligned_alloc(64, N * 4); i2 = aligned_alloc(64, N * 4);
    memset(sc, 1, N * 4); memset(lp, 2, N * 4); memset(off, 3, N);
    memset(d1, 0, N * 8); memset(d2, 0, N * 8); memset(i1, 0, N * 4); memset(i2, 0, N * 4);
    int tl[] = {1, 2, 4, 6, 8, 12, 16};
    for (mode = 0; mode < 2; ++mode)
        for (unsigned ti = 0; ti < sizeof(tl) / sizeof(tl[0]); ++ti) {
            int T = tl[ti];
            double best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                pthread_t th[64];
                range_t rg[64];
                double t0 = now();
                for (int t = 0; t < T; ++t) {
                    rg[t].a = (N / T * t) & ~(size_t)63;
                    rg[t].b = t == T - 1 ? N : ((N / T * (t + 1)) & ~(size_t)63);
                    pthread_create(&th[t], NULL, work, &rg[t]);
                }
                for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
                double dt = now() - t0;
                if (dt < best) best = dt;
            }
            printf("%s stores, %2d threads: %.2f ms for %zu entries (%.1f GB/s written, %.1f GB/s read)\n",
                   mode ? "non-temporal" : "plain", T, best * 1e3, N, N * 24e-9 / best, N * 9e-9 / best);
        }
    // plain memcpy of the same volume for scale
    double t0 = now();
    memcpy(d1, d2, N * 8);
    printf("memcpy 1 thread: %.1f GB/s\n", N * 8e-9 / (now() - t0));
    return 0;
}
