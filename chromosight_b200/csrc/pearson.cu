// K1: fused sliding-window Pearson correlation of a small dense kernel against a
// banded / dense float32 image (detection.py:917-1131 of the reference).
//
// One CTA = one output tile of TR rows that follows the band (a parallelogram in
// matrix coordinates), possibly one of several column chunks.  The input tile
// ((TR+KH-1) rows x IC columns, matrix coordinates) is fetched with ONE TMA box
// load out of the skewed band (the row stride of the tensor map is `pitch`, see
// cs_layout); the kernel tables arrive by a bulk copy on the same mbarrier.  Then,
// all in shared memory:
//   fix-up   one warp per tile row: out-of-band aliases and the declared strip -> 0, NaN
//            sentinels -> bit array + value 0 (per-row keep words, branch-free);
//   per thread, a block of RU x RT = 2 x 8 windows (footprint 18 x 24 pixels for 17 x 17):
//   pivot    mean of the middle footprint row; all sums are taken on S - pivot (exact
//            algebra, keeps float32 accurate);
//   FMA pass per input row one LDS.128 sweep of the 24-pixel segment, then per window 9
//            packed fma.rn.f32x2 whose two halves collect alternate taps (aligned register
//            pairs: no shuffles, no duplicated taps) -- no mask work in this loop;
//   sums     packed column sums of S - pivot and its square over KH rows, then KW-wide
//            sliding sums along the row;
//   epilogue per window: masked kernel sums and the missing count from the bit array
//            (missing pixels of a footprint grouped into rectangles, summed through 2-D
//            prefix tables of K and K^2 in float64, exact), the reference's formulas in
//            float64, one float32 score (+ uint16 observation count); ill-conditioned
//            windows are redone in float64 from the tile by the whole warp.
// Kernels wider than 31 columns take pearson_wide (one warp per window, float64).
// No tensor cores: this is a CUDA-core stencil (BASELINE.json north_star).
#include <stdlib.h>
#include <vector>
#include "common.cuh"

namespace cs {

constexpr int RU = 2;  // window rows per thread
constexpr int RT = 8;  // window columns per thread
// column shift of row group g in a banded (skewed) traversal: the band moves right by one
// column per row; blocks stay 16-byte aligned in the tile
__host__ __device__ __forceinline__ int skew_shift(int g) { return (RU * g) & ~3; }
constexpr int kSkewSlack = (RU % 4) ? 3 : 0;  // columns lost to that rounding

struct PearsonParams {
    // image
    int rows, cols, dlo, dhi, dense;
    // output pixel set (image coordinates)
    int oy0, oy1, ox0, ox1, odlo, odhi;
    // tiling
    int TR, G, NBc, nchunks, skew;
    int IC, IR, NW;
    // kernel geometry
    int KH, KW, KWP2, N;
    // output image
    float *out;
    unsigned short *nobs;
    int out_pitch, out_dlo, osy, osx;
    // tables (device): float part then double part, copied to shared memory by every CTA
    const float *ftab;
    const double *dtab;
    int n_ftab, n_dtab, tab_bytes;
    unsigned long long *cnt;  // CS_DEBUG_COUNT: statistics of the mask code (experiments)
    int dbg;  // CS_DEBUG_SKIP bit mask (timing experiments only): see the kernel
    double q, sumKp, sumKp2, ksum, k2sum, kmean, kstd, thr, invN, vK0;
    int min_present, kmean_zero, has_mask, raw_xcorr, nobs_full;
    int sdlo, sdhi, st_base, st_n;  // declared-missing diagonal strip and its tables
    // shared memory carve-up (bytes)
    int off_bits, off_K, off_D, off_stat, off_grp, off_bar;
};

// ---------------------------------------------------------------- PTX helpers  // [sec:packedops]
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x,
                                            int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// 1-D bulk copy global -> shared memory, completion counted on the mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// two float32 FMAs per instruction (FFMA2): acc.{lo,hi} += a.{lo,hi} * b.{lo,hi}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a,
                                     unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
// acc.{lo,hi} += a.{lo,hi}
__device__ __forceinline__ void acc2(unsigned long long &acc, unsigned long long a) {
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__device__ __forceinline__ double thr0(double v, double t) { return fabs(v) < t ? 0.0 : v; }  // [sec:scorefn]

// Squared error amplification above which a window is recomputed in float64: the
// float32 sums carry ~1e-6 relative error on well-conditioned windows, and the score
// error grows like amp = f * rms(S') * rms(K') / (sigma_S * sigma_K).
constexpr double kAmpLimit2 = 9.0;
// Same for the column sums, which are computed in float64 and rounded once to float32
// (relative error 6e-8 each): var(S) loses at most 1.8e-7 * amp^2.
constexpr double kAmpLimitV = 25.0;

// The reference's formulas (det:1002-1020 no mask, det:1021-1092 masked) from the
// window sums of the shifted signal S' = S - p and the shifted kernel K' = K_corr - q.
//   h1, h2    : sum S', sum S'^2 over the N window pixels (missing pixels count as S = 0)
//   h2_loc    : sum (S' - block pivot)^2, the quantity the float32 product sum was formed on
//   s3        : sum S' * K'
//   sKm, sKm2 : sums of the mask kernels (K and K^2) over the missing pixels
// `redo` is set when the window is too ill-conditioned for the float32 sums.
template <bool MASK>
__device__ __forceinline__ float score_from_sums(const PearsonParams &P, double p, double h1,
                                                 double h2, double h2_loc, int nmiss, double s3,
                                                 double sKm, double sKm2, int &nobs, bool &redo) {
    const double invN = P.invN;
    const double m1 = h1 * invN;
    double A1 = m1 + p;
    double A2 = fma(h2, invN, fma(2.0 * p, m1, p * p));
    double A3 = fma(s3 + fma(P.q, h1, p * P.sumKp), invN, p * P.q);
    nobs = P.N;
    redo = false;
    if (P.raw_xcorr) return (float)thr0(A3 * (double)P.N, P.thr);
    A1 = thr0(A1, P.thr);
    A2 = thr0(A2, P.thr);
    A3 = thr0(A3, P.thr);
    double cov, den2, f2 = 1.0;
    bool ok = true;
    if (!MASK) {
        const double vS = fma(-A1, A1, A2);
        cov = fma(-A1, P.kmean, A3);
        den2 = vS * P.vK0;
        ok = vS >= 0.0;
    } else {
        // one branch-free path: with nmiss == 0 these are the unmasked formulas
        // (f = 1, mK = kernel mean, m2K = mean of K^2)
        const int npres = P.N - nmiss;
        // 1 / npres: float32 reciprocal + one Newton step (exact to ~1e-15)
        double inv = (double)__frcp_rn((float)npres);
        inv = inv * fma(-(double)npres, inv, 2.0);
        const double f = (double)P.N * inv;
        f2 = f * f;
        sKm = thr0(sKm, P.thr);
        sKm2 = thr0(sKm2, P.thr);
        const double mK = (P.ksum - sKm) * inv;
        const double m2K = (P.k2sum - sKm2) * inv;
        const double mS = A1 * f;
        const double vS = fma(A2, f, -mS * mS);
        cov = fma(-A1, mK, A3) * f;
        den2 = vS * fma(-mK, mK, m2K);
        ok = nmiss == 0 || ((npres > 0) && (npres >= P.min_present) && !P.kmean_zero);
        if (P.nobs_full && nmiss != 0 && npres != 0) nobs = npres;
    }
    // det:1066,1088-1091: denom = sqrt(den2); |denom| < 1e-10 or NaN -> 0
    float r = 0.f;
    if (ok && den2 >= 1e-20 && den2 < 1e300) {
        // float32 product sums: conditioned by the spread around the block pivot (h2_loc);
        // float32-rounded column sums: by the spread around the tile pivot (h2)
        const double c1 = P.sumKp2 * invN * invN * f2;
        redo = (h2_loc * c1 > kAmpLimit2 * den2) || (h2 * c1 > kAmpLimitV * den2);
        r = (float)cov * rsqrtf((float)den2);
        if (!(fabsf(r) <= 3.0e38f)) r = 0.f;
        r = fminf(1.f, fmaxf(-1.f, r));
    } else if (ok && A2 != 0.0 && den2 < 1e-20) {
        // a non-zero window whose variance vanished: exactly flat, or flat up to the
        // float32 rounding of the column sums -- let the float64 pass decide
        redo = true;
    }
    return r;
}

// 64 bits of the (linear, one bit per tile pixel) missing-pixel bit array from bit `pos` on  // [sec:maskfn]
__device__ __forceinline__ unsigned long long row_bits(const uint32_t *bits, int pos) {
    const int w = pos >> 5, sh = pos & 31;
    const uint32_t a = bits[w], b = bits[w + 1], c = bits[w + 2];
    const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return ((unsigned long long)hi << 32) | lo;
}

// Missing pixels of a window given as kernel rows [i0, i1) x the set bits of wb: add the
// sums of the two mask kernels over them.  IK / IK2 are 2-D prefix tables,
// IK[i][j] = sum of K[i' < i][j' < j], row pitch KW1.
__device__ __forceinline__ void add_rects(unsigned wb, int i0, int i1, const double *IK,
                                          const double *IK2, int KW1, double &sKm, double &sKm2) {
    const double *t0 = IK + i0 * KW1, *t1 = IK + i1 * KW1;
    const double *u0 = IK2 + i0 * KW1, *u1 = IK2 + i1 * KW1;
    while (wb) {
        const int a = __ffs(wb) - 1;
        const int b = a + __ffs(~(wb >> a)) - 1;  // first zero above a ends the run (< 32 bits)
        sKm += (t1[b] - t1[a]) - (t0[b] - t0[a]);
        sKm2 += (u1[b] - u1[a]) - (u0[b] - u0[a]);
        wb = (b >= 32) ? 0u : (wb >> b) << b;
    }
}

constexpr int kMaxGroups = 4;  // rectangles of missing pixels per footprint kept in shared memory  // [sec:kernelsetup]

// ---------------------------------------------------------------- the kernel
template <int KW, bool MASK>
__global__ void __launch_bounds__(256, 2)
pearson_tiles(const __grid_constant__ CUtensorMap tmap, const PearsonParams P) {
    constexpr int kw = (KW - 1) / 2;
    constexpr int kwa = (kw + 3) / 4 * 4;
    constexpr int off = kwa - kw;
    constexpr int XW = RT + KW - 1;           // pixels of one footprint row
    constexpr int NQ = (off + XW + 3) / 4;    // float4 loads per footprint row
    constexpr int KWP2 = (KW + 1 + 3) / 4 * 4;  // floats of one padded tap row
    constexpr int NP = (KW + 1) / 2;            // tap pairs per kernel row
    static_assert(((off + RT - 1) >> 1) + NP <= 2 * NQ, "tap pairs run past the loaded segment");
    constexpr unsigned KWMASK = (KW == 32) ? 0xffffffffu : ((1u << KW) - 1u);

    extern __shared__ __align__(1024) unsigned char smem[];
    float *__restrict__ tile = reinterpret_cast<float *>(smem);
    uint32_t *__restrict__ bits = reinterpret_cast<uint32_t *>(smem + P.off_bits);
    // taps K' = K_corr - q, two rows of KWP2 floats per kernel row: the row padded with
    // zeros (pairs (k0,k1),(k2,k3),... for windows at an even tile column) and the same row
    // shifted right by one (pairs (0,k0),(k1,k2),... for windows at an odd tile column)
    const float *__restrict__ Ktab = reinterpret_cast<const float *>(smem + P.off_K);
    const double *__restrict__ Dt = reinterpret_cast<const double *>(smem + P.off_D);
    float *__restrict__ statS = reinterpret_cast<float *>(smem + P.off_stat);
    unsigned long long *__restrict__ grpS =
        reinterpret_cast<unsigned long long *>(smem + P.off_grp);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P.off_bar);

    const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
    const int rb = blockIdx.x / P.nchunks;
    const int ch = blockIdx.x - rb * P.nchunks;
    const int KH = P.KH;
    const int kh = (KH - 1) / 2;
    const int Y0 = P.oy0 + rb * P.TR;
    // aligned X' (= X - dlo) of the first block of row group 0
    int xb;
    {
        const int v = P.skew ? (Y0 + P.odlo - P.dlo) : (P.ox0 - P.dlo);
        xb = (v >= 0 ? v / 4 : -((-v + 3) / 4)) * 4;
    }
    xb += RT * ch * P.NBc;
    const int TXp = xb - kwa;  // X' of tile column 0
    const int TY = Y0 - kh;    // image row of tile row 0
    const int IC = P.IC, IR = P.IR, NW = P.NW;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        fence_proxy_async();
        // the tile (2-D tensor box) and the kernel tables (one linear block: float part, then
        // float64 part) arrive through the same barrier
        mbar_expect_tx(bar, (uint32_t)(IC * IR * sizeof(float)) + (uint32_t)P.tab_bytes);
        tma_load_2d(tile, &tmap, bar, TXp, TY);
        bulk_load_1d(smem + P.off_K, P.ftab, (uint32_t)P.tab_bytes, bar);
    }
    if (MASK)
        for (int i = tid; i < NW; i += nthr) bits[i] = 0u;
    __syncthreads();
    mbar_wait(bar, 0);

    // ---- phase A: fix-up of the tile, four pixels per thread ---------------------- [sec:A1]
    // out-of-band aliases and declared-missing strip -> 0, NaN sentinels -> bit array and 0
    // (missing pixels count as S = 0, det:1050-1060)
    int anynz = 0;
    {
        const int ICq4 = IC >> 2;
        for (int iy = (P.dbg & 16) ? IR : (tid >> 5); iy < IR; iy += nthr >> 5) {
            // in-band tile columns of this row: [clo, chi)
            const int dbase = (TXp + P.dlo) - (TY + iy);  // diagonal of tile column 0
            int clo = 0, chi = IC, slo = 0, shi = 0;
            if (!P.dense) {
                clo = max(P.dlo - dbase, 0);
                chi = min(P.dhi - dbase + 1, IC);
            }
            if (MASK) {
                slo = P.sdlo - dbase;
                shi = P.sdhi - dbase + 1;  // declared-missing strip: counted analytically
            }
            // keep-bits of the row, 32 columns per lane (lanes 0..7): in the band, off the strip
            unsigned rowword;
            {
                const int base = 32 * lane;
                const int a = min(max(clo - base, 0), 32), b = min(max(chi - base, 0), 32);
                rowword = (unsigned)(((1ull << b) - 1ull) & ~((1ull << a) - 1ull));
                if (MASK) {
                    const int sa = min(max(slo - base, 0), 32), sb = min(max(shi - base, 0), 32);
                    if (sb > sa) rowword &= ~(unsigned)(((1ull << sb) - 1ull) & ~((1ull << sa) - 1ull));
                }
            }
            for (int cb = 0; cb < ICq4; cb += 32) {
                const int c4 = cb + lane;
                const int c0 = 4 * c4;
                const unsigned kwd = __shfl_sync(0xffffffffu, rowword, (c0 >> 5) & 31);
                if (c4 >= ICq4) continue;
                float4 *ptr = reinterpret_cast<float4 *>(tile + iy * IC) + c4;
                const float4 v = *ptr;
                const unsigned keep = (kwd >> (c0 & 31)) & 0xfu;
                // NaN sentinels among the kept pixels -> bit array; dropped and missing pixels -> 0
                unsigned live = keep;
                if (v.x != v.x) live &= ~1u;
                if (v.y != v.y) live &= ~2u;
                if (v.z != v.z) live &= ~4u;
                if (v.w != v.w) live &= ~8u;
                const unsigned nb = keep & ~live;
                const float4 w = make_float4((live & 1u) ? v.x : 0.f, (live & 2u) ? v.y : 0.f,
                                             (live & 4u) ? v.z : 0.f, (live & 8u) ? v.w : 0.f);
                anynz |= (__float_as_uint(w.x) | __float_as_uint(w.y) | __float_as_uint(w.z) |
                          __float_as_uint(w.w)) != 0u;
                if (live != 0xfu) *ptr = w;
                if (MASK && nb) atomicOr(&bits[(iy * IC + c0) >> 5], nb << ((iy * IC + c0) & 31));
            }
        }
    }
    const int tnz = __syncthreads_or(anynz);

    if (!tnz) {
        // all-zero signal: every score of the tile is 0 (variance 0 -> det:1088-1091)
        for (int item = tid; item < P.G * P.NBc; item += nthr) {
            const int g = item / P.NBc, m = item - g * P.NBc;
            const int Xp0 = xb + skew_shift(g) * P.skew + RT * m;
            for (int u = 0; u < RU; ++u) {
                const int Y = Y0 + RU * g + u;
                if (Y >= P.oy1) continue;
                for (int t = 0; t < RT; ++t) {
                    const int X = Xp0 + t + P.dlo;
                    const int d = X - Y;
                    if (X < P.ox0 || X >= P.ox1 || d < P.odlo || d > P.odhi) continue;
                    const long long oi =
                        (long long)(Y - P.osy) * P.out_pitch + ((X - P.osx) - P.out_dlo);
                    P.out[oi] = 0.f;
                    if (P.nobs) P.nobs[oi] = (unsigned short)P.N;
                }
            }
        }
        return;
    }

    // double tables (mask branch): 2-D prefix sums of the mask kernels, their column sums,
    // and the sums over the declared-missing strip per output diagonal
    constexpr int KW1 = KW + 1;
    const double *IK = Dt;                              // [KH + 1][KW + 1]
    const double *IK2 = IK + (KH + 1) * KW1;            // [KH + 1][KW + 1]
    const double *Kcol = IK2 + (KH + 1) * KW1;          // [KW]
    const double *K2col = Kcol + KW;                    // [KW]
    const double *StK = K2col + KW;                     // [st_n]
    const double *StK2 = StK + P.st_n;                  // [st_n]
    const double *StC = StK2 + P.st_n;                  // [st_n]

    // ---- blocks of RU x RT windows, one per thread -------------------------------------
    // Lanes 0-3 / 4-7 of every quarter-warp take blocks of row groups g / g + 2: in a banded
    // traversal their tile columns differ by 4 (mod 8), which makes the 16-byte loads of a
    // quarter-warp hit disjoint banks.
    const int NQd = (P.NBc + 3) >> 2;
    const int nitems = 8 * NQd * 2 * ((P.G + 3) >> 2);
    // every lane runs every loop (work is predicated): the epilogue contains warp-wide steps
    for (int base = (P.dbg & 256) ? nitems : 0; base < nitems; base += nthr) {
        int g, m;
        bool live;
        {
            const int idx = base + tid;
            // a warp = 4 column blocks x 8 row groups: a missing row or column then touches
            // most lanes of the warps it crosses (less divergence in the mask code)
            const int npair = 2 * ((P.G + 3) >> 2);
            const int rest = idx >> 3, pr = rest % npair, M = rest / npair;
            g = (pr >> 1) * 4 + (pr & 1) + 2 * ((idx >> 2) & 1);
            m = 4 * M + (idx & 3);
            live = idx < nitems && g < P.G && m < P.NBc;
            if (!live) g = m = 0;
        }
        const int Xp0 = xb + skew_shift(g) * P.skew + RT * m;  // X' of output column t = 0
        const int cxa = skew_shift(g) * P.skew + RT * m;       // aligned tile column of x[0]
        const int Yg = Y0 + RU * g;
        // blocks without any valid output pixel are skipped
        bool any = false;
        if (live) {
            for (int u = 0; u < RU; ++u) {
                const int Y = Yg + u;
                if (Y >= P.oy1) continue;
                const int Xlo = max(P.ox0, Y + P.odlo), Xhi = min(P.ox1 - 1, Y + P.odhi);
                const int Xa = Xp0 + P.dlo;
                if (Xa + RT - 1 >= Xlo && Xa <= Xhi) any = true;
            }
        }
        // per-window results of the two passes, parked in shared memory: the epilogue runs as
        // compact loops.  [k], [NWB + k], [2 NWB + k] with k = u * RT + t.
        constexpr int NWB = RU * RT;
        float *mystat = statS + tid;
        float pl = 0.f;
        if (any) {                                                           // [sec:pivot]
            // block pivot: mean of the middle footprint row.  Sums and products are formed on
            // S - pl (any pivot is algebraically exact; a close one keeps float32 accurate).
            const float4 *rp4 = reinterpret_cast<const float4 *>(
                tile + (RU * g + (KH + RU - 1) / 2) * IC + cxa);
            float sacc = 0.f;
#pragma unroll
            for (int qd = 0; qd < NQ; ++qd) {
                const float4 v = rp4[qd];
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (4 * qd + e >= off && 4 * qd + e < off + XW) sacc += vv[e];
            }
            pl = sacc * (1.0f / (float)XW);
        }
        if (any && !(P.dbg & 8)) {                                           // [sec:main]
            const unsigned long long npl2 = pack2(-pl, -pl);
            // one packed accumulator per window: .lo and .hi collect alternate taps
            unsigned long long acc[RU][RT];
#pragma unroll
            for (int u = 0; u < RU; ++u)
#pragma unroll
                for (int t = 0; t < RT; ++t) acc[u][t] = 0ull;

            const int nrow = KH + RU - 1;
#pragma unroll 1
            for (int iy = 0; iy < nrow; ++iy) {
                const ulonglong2 *rp =
                    reinterpret_cast<const ulonglong2 *>(tile + (RU * g + iy) * IC + cxa);
                // xe[q] = (x[2q], x[2q+1]) - block pivot
                unsigned long long xe[2 * NQ];
#pragma unroll
                for (int qd = 0; qd < NQ; ++qd) {
                    const ulonglong2 v = rp[qd];
                    xe[2 * qd] = add2(v.x, npl2);
                    xe[2 * qd + 1] = add2(v.y, npl2);
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int i = iy - u;
                    if (i < 0 || i >= KH) continue;  // uniform across the block
                    const ulonglong2 *kp =
                        reinterpret_cast<const ulonglong2 *>(Ktab + i * 2 * KWP2);
                    unsigned long long ke[KWP2 / 2], ko[KWP2 / 2];
#pragma unroll
                    for (int qd = 0; qd < KWP2 / 4; ++qd) {
                        const ulonglong2 a = kp[qd], b = kp[KWP2 / 4 + qd];
                        ke[2 * qd] = a.x;
                        ke[2 * qd + 1] = a.y;
                        ko[2 * qd] = b.x;
                        ko[2 * qd + 1] = b.y;
                    }
#pragma unroll
                    for (int t = 0; t < RT; ++t) {
                        const int c = off + t;  // compile-time after unrolling
#pragma unroll
                        for (int mm = 0; mm < NP; ++mm)
                            fma2(acc[u][t], xe[(c >> 1) + mm], (c & 1) ? ko[mm] : ke[mm]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RU; ++u)
#pragma unroll
                for (int t = 0; t < RT; ++t) {
                    float lo, hi;
                    unpack2(acc[u][t], lo, hi);
                    mystat[(u * RT + t) * nthr] = lo + hi;
                }
        }
        if (any && !(P.dbg & 4)) {                                           // [sec:sums]
            // window sums of (S - pl) and (S - pl)^2: column sums over KH rows in packed
            // registers, then KW-wide sliding sums along the row
            const unsigned long long npl2 = pack2(-pl, -pl);
            unsigned long long cs[2 * NQ], cq[2 * NQ];
#pragma unroll
            for (int q = 0; q < 2 * NQ; ++q) cs[q] = cq[q] = 0ull;
#pragma unroll 1
            for (int iy = 0; iy < KH; ++iy) {
                const ulonglong2 *rp =
                    reinterpret_cast<const ulonglong2 *>(tile + (RU * g + iy) * IC + cxa);
#pragma unroll
                for (int qd = 0; qd < NQ; ++qd) {
                    const ulonglong2 v = rp[qd];
                    const unsigned long long a = add2(v.x, npl2), b = add2(v.y, npl2);
                    acc2(cs[2 * qd], a);
                    acc2(cs[2 * qd + 1], b);
                    fma2(cq[2 * qd], a, a);
                    fma2(cq[2 * qd + 1], b, b);
                }
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                if (u > 0) {
                    // rows [u, u + KH): row u + KH - 1 enters, row u - 1 leaves
                    const ulonglong2 *rin = reinterpret_cast<const ulonglong2 *>(
                        tile + (RU * g + u + KH - 1) * IC + cxa);
                    const ulonglong2 *rout =
                        reinterpret_cast<const ulonglong2 *>(tile + (RU * g + u - 1) * IC + cxa);
                    const unsigned long long pl2 = pack2(pl, pl);
#pragma unroll
                    for (int qd = 0; qd < NQ; ++qd) {
                        const ulonglong2 vi = rin[qd], vo = rout[qd];
                        const unsigned long long a = add2(vi.x, npl2), b = add2(vi.y, npl2);
                        const unsigned long long c = add2(vo.x, npl2), d = add2(vo.y, npl2);
                        const unsigned long long nc = sub2(pl2, vo.x), nd = sub2(pl2, vo.y);
                        cs[2 * qd] = add2(add2(cs[2 * qd], a), nc);
                        cs[2 * qd + 1] = add2(add2(cs[2 * qd + 1], b), nd);
                        fma2(cq[2 * qd], a, a);
                        fma2(cq[2 * qd + 1], b, b);
                        fma2(cq[2 * qd], nc, c);  // - (x - pl)^2
                        fma2(cq[2 * qd + 1], nd, d);
                    }
                }
                float col[4 * NQ], cqq[4 * NQ];
#pragma unroll
                for (int q = 0; q < 2 * NQ; ++q) {
                    unpack2(cs[q], col[2 * q], col[2 * q + 1]);
                    unpack2(cq[q], cqq[2 * q], cqq[2 * q + 1]);
                }
                float g1 = 0.f, g2 = 0.f;
#pragma unroll
                for (int e = off; e < off + KW; ++e) {
                    g1 += col[e];
                    g2 += cqq[e];
                }
#pragma unroll
                for (int t = 0; t < RT; ++t) {
                    if (t > 0) {
                        g1 += col[off + t + KW - 1] - col[off + t - 1];
                        g2 += cqq[off + t + KW - 1] - cqq[off + t - 1];
                    }
                    mystat[(NWB + u * RT + t) * nthr] = g1;
                    mystat[(2 * NWB + u * RT + t) * nthr] = g2;
                }
            }
        }

        // footprint mask summary.  colfull: columns missing on every footprint row.  The other
        // missing pixels: consecutive rows with the same pattern form one rectangle group
        // (a missing row, the visible part of a missing column at the edge of the mask band,
        // a frame margin); footprints with more groups than fit fall back to row-by-row.
        const int fc0 = cxa + off;  // tile column of footprint column 0
        unsigned long long colfull = 0ull, rowsel = 0ull, bor = 0ull;                // [sec:masksum]
        int ng = 0;
        if (MASK && any && !(P.dbg & 1)) {
            constexpr unsigned long long FWMASK = (XW >= 64) ? ~0ull : ((1ull << XW) - 1ull);
            unsigned long long band = FWMASK;
            const int fr = KH + RU - 1;
            for (int r = 0; r < fr; ++r) {
                const unsigned long long b = row_bits(bits, (RU * g + r) * IC + fc0) & FWMASK;
                band &= b;
                bor |= b;
            }
            if (bor != 0ull) {
                colfull = band;
                if (bor & ~band) {
                    unsigned long long prev = 0ull;
                    int start = 0;
                    for (int r = 0; r <= fr; ++r) {
                        unsigned long long b = 0ull;
                        if (r < fr) b = row_bits(bits, (RU * g + r) * IC + fc0) & FWMASK & ~band;
                        if (b) rowsel |= 1ull << r;
                        if (b != prev) {
                            if (prev) {
                                if (ng < kMaxGroups)  // pattern (< 40 bits) | first row | end row
                                    grpS[ng * nthr + tid] = prev | ((unsigned long long)start << 40) |
                                                            ((unsigned long long)r << 48);
                                ++ng;
                            }
                            start = r;
                            prev = b;
                        }
                    }
                }
            }
        }

        // valid windows of the block, bit k = u * RT + t
        unsigned okb = 0u;
        if (any) {
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int Y = Yg + u;
                // columns of the row: [max(ox0, Y + odlo), min(ox1 - 1, Y + odhi)] as t range
                const int tlo = max(max(P.ox0, Y + P.odlo) - (Xp0 + P.dlo), 0);
                const int thi = min(min(P.ox1 - 1, Y + P.odhi) - (Xp0 + P.dlo), RT - 1);
                if (Y < P.oy1 && thi >= tlo)
                    okb |= (((2u << thi) - 1u) & ~((1u << tlo) - 1u)) << (u * RT);
            }
        }
        if (P.cnt && MASK && any) {
            constexpr unsigned long long FWM2 = (XW >= 64) ? ~0ull : ((1ull << XW) - 1ull);
            atomicAdd(&P.cnt[0], 1ull);                       // blocks
            if (bor) atomicAdd(&P.cnt[1], 1ull);              // blocks with any missing pixel
            if (ng > 0) atomicAdd(&P.cnt[2], 1ull);           // blocks with groups
            if (ng > kMaxGroups) atomicAdd(&P.cnt[3], 1ull);  // blocks on the row-by-row path
            atomicAdd(&P.cnt[4], (unsigned long long)ng);     // groups
            for (int k = 0; k < min(ng, kMaxGroups); ++k) {
                const unsigned long long gp = grpS[k * nthr + tid];
                if ((gp & ((1ull << 40) - 1ull)) == (FWM2 & ~colfull)) atomicAdd(&P.cnt[5], 1ull);  // full rows
                atomicAdd(&P.cnt[6], (unsigned long long)((int)(gp >> 48) - (int)((gp >> 40) & 255)));  // rows in groups
            }
            if (colfull) atomicAdd(&P.cnt[7], 1ull);          // blocks with full columns
        }
        if (P.dbg & 512) bor = 0ull;  // timing experiment: summary computed, window sums skipped
#pragma unroll 1
        for (int u = 0; u < RU; ++u) {
            const int Y = Yg + u;
#pragma unroll 1
            for (int t = 0; t < RT; ++t) {
                const int X = Xp0 + t + P.dlo;
                const int d = X - Y;
                const bool wok = (okb >> (u * RT + t)) & 1u;
                int nmiss = 0;
                double sKm = 0.0, sKm2 = 0.0;
                if (MASK && wok) {                                           // [sec:strip]
                    // the declared-missing diagonal strip: a function of the window's diagonal
                    const unsigned sd = (unsigned)(d - P.st_base);
                    if (sd < (unsigned)P.st_n) {
                        nmiss = (int)StC[sd];
                        sKm = StK[sd];
                        sKm2 = StK2[sd];
                    }
                }
                if (MASK && wok && ((unsigned)(bor >> t) & KWMASK)) {        // [sec:rects]
                    // columns missing over the whole footprint: whole kernel columns
                    const unsigned cbw = (P.dbg & 1024) ? 0u : ((unsigned)(colfull >> t) & KWMASK);
                    nmiss += KH * __popc(cbw);
                    for (unsigned c = cbw; c; c &= c - 1) {
                        const int j = __ffs(c) - 1;
                        sKm += Kcol[j];
                        sKm2 += K2col[j];
                    }
                    if (P.dbg & 2048) {
                    } else if (ng <= kMaxGroups) {
                        // remaining missing pixels as rectangles: rows [i0, i1) x runs of taps
                        for (int k = 0; k < ng; ++k) {
                            const unsigned long long gp = grpS[k * nthr + tid];
                            const int i0 = max((int)((gp >> 40) & 255) - u, 0);
                            const int i1 = min((int)(gp >> 48) - u, KH);
                            const unsigned wb = (unsigned)(gp >> t) & KWMASK;
                            if (i0 < i1 && wb) {
                                nmiss += (i1 - i0) * __popc(wb);
                                add_rects(wb, i0, i1, IK, IK2, KW1, sKm, sKm2);
                            }
                        }
                    } else {
                        // row by row
                        const unsigned KHMASK = (KH >= 32) ? 0xffffffffu : ((1u << KH) - 1u);
                        for (unsigned rs = (unsigned)(rowsel >> u) & KHMASK; rs; rs &= rs - 1) {
                            const int i = __ffs(rs) - 1;
                            const unsigned wb =
                                (unsigned)(row_bits(bits, (RU * g + u + i) * IC + fc0) >> t) &
                                KWMASK & ~cbw;
                            nmiss += __popc(wb);
                            add_rects(wb, i, i + 1, IK, IK2, KW1, sKm, sKm2);
                        }
                    }
                }
                int nobs = P.N;                                              // [sec:score]
                bool redo = false;
                float r = 0.f;
                if (wok) {
                    const int k = u * RT + t;
                    const double s3 = (double)mystat[k * nthr];
                    const double g1 = (double)mystat[(NWB + k) * nthr];
                    const double g2 = (double)mystat[(2 * NWB + k) * nthr];
                    if (P.dbg & 2)
                        r = (float)(s3 + g1 + g2);
                    else
                        r = score_from_sums<MASK>(P, (double)pl, g1, g2, g2, nmiss, s3, sKm, sKm2,
                                                  nobs, redo);
                    if (P.dbg & 32) redo = false;
                }
                // ill-conditioned windows (flat signal or mostly missing): the warp redoes the
                // window sums in float64 from the tile, 32 pixels at a time      [sec:redo]
                const int woff = (RU * g + u) * IC + fc0 + t;
                for (unsigned todo = __ballot_sync(0xffffffffu, redo); todo; todo &= todo - 1) {
                    const int src = __ffs(todo) - 1;
                    const float *wp = tile + __shfl_sync(0xffffffffu, woff, src);
                    double h1 = 0.0, h2 = 0.0, s3 = 0.0;
                    for (int idx = lane; idx < KH * KW; idx += 32) {
                        const int i = idx / KW, j = idx - i * KW;
                        const double sv = (double)wp[i * IC + j];
                        h1 += sv;
                        h2 = fma(sv, sv, h2);
                        s3 = fma(sv, (double)Ktab[i * 2 * KWP2 + j], s3);
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        h1 += __shfl_xor_sync(0xffffffffu, h1, o);
                        h2 += __shfl_xor_sync(0xffffffffu, h2, o);
                        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
                    }
                    if (lane == src) {
                        bool again;
                        r = score_from_sums<MASK>(P, 0.0, h1, h2, h2, nmiss, s3, sKm, sKm2, nobs,
                                                  again);
                    }
                }
                if (wok) {
                    // score and observation count replace s3 and g1 in the window's slots
                    mystat[(u * RT + t) * nthr] = r;
                    mystat[(NWB + u * RT + t) * nthr] = __int_as_float(nobs);
                }
            }
        }
        // ---- scores (and observation counts) to the output band, 16 bytes at a time where
        // the row of eight windows is complete and aligned                     [sec:store]
        if (!(P.dbg & 64)) {
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const unsigned ob = (okb >> (u * RT)) & ((1u << RT) - 1u);
                if (!ob) continue;
                const int Y = Yg + u;
                const long long oi0 =
                    (long long)(Y - P.osy) * P.out_pitch + ((Xp0 + P.dlo - P.osx) - P.out_dlo);
                float rr[RT];
                int nn[RT];
#pragma unroll
                for (int t = 0; t < RT; ++t) {
                    rr[t] = mystat[(u * RT + t) * nthr];
                    nn[t] = __float_as_int(mystat[(NWB + u * RT + t) * nthr]);
                }
                if (ob == ((1u << RT) - 1u) && (oi0 & 3) == 0) {
#pragma unroll
                    for (int t = 0; t < RT; t += 4)
                        *reinterpret_cast<float4 *>(P.out + oi0 + t) =
                            make_float4(rr[t], rr[t + 1], rr[t + 2], rr[t + 3]);
                    if (P.nobs) {
#pragma unroll
                        for (int t = 0; t < RT; t += 4)
                            *reinterpret_cast<uint2 *>(P.nobs + oi0 + t) = make_uint2(
                                (unsigned)nn[t] | ((unsigned)nn[t + 1] << 16),
                                (unsigned)nn[t + 2] | ((unsigned)nn[t + 3] << 16));
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < RT; ++t)
                        if ((ob >> t) & 1u) {
                            P.out[oi0 + t] = rr[t];
                            if (P.nobs) P.nobs[oi0 + t] = (unsigned short)nn[t];
                        }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- wide kernels
// One warp per window, any kernel size: the path for kernels wider than the tiled kernel's
// 31 columns (the 81 x 81 centromere preset scans one or two diagonals, so the window count is
// small and the footprint large).  The six window sums of det:1002-1092 are accumulated in
// float64 straight from the image in HBM / L2 (NaN sentinel = missing pixel, pixels off the
// stored band = 0), then the same formulas as everywhere else.
struct WideParams {
    const float *img;
    int pitch, dlo, dhi, dense;
    const double *kc, *km, *k2m;  // K_corr, K_mask, K2_mask [KH * KW]
};

template <bool MASK>
__global__ void __launch_bounds__(256)
pearson_wide(const PearsonParams P, const WideParams W) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int KH = P.KH, KW = P.KW, kh = (KH - 1) / 2, kw = (KW - 1) / 2;
    for (int Y = P.oy0 + blockIdx.x; Y < P.oy1; Y += gridDim.x) {
        const int Xlo = max(P.ox0, Y + P.odlo), Xhi = min(P.ox1 - 1, Y + P.odhi);
        for (int X = Xlo + blockIdx.y * wpb + (threadIdx.x >> 5); X <= Xhi; X += gridDim.y * wpb) {
            double h1 = 0.0, h2 = 0.0, s3 = 0.0, sKm = 0.0, sKm2 = 0.0;
            int nmiss = 0;
            for (int idx = lane; idx < KH * KW; idx += 32) {
                const int i = idx / KW, j = idx - i * KW;
                const int Yp = Y - kh + i, Xp = X - kw + j;
                float v = 0.f;
                if (W.dense)
                    v = W.img[(long long)Yp * W.pitch + Xp];
                else {
                    const int d = Xp - Yp;
                    if (d >= W.dlo && d <= W.dhi) v = W.img[(long long)Yp * W.pitch + (Xp - W.dlo)];
                }
                if (!(v == v)) {  // missing pixel: counts as S = 0
                    if (MASK) {
                        ++nmiss;
                        sKm += W.km[idx];
                        sKm2 += W.k2m[idx];
                    }
                    continue;
                }
                const double sv = (double)v;
                h1 += sv;
                h2 = fma(sv, sv, h2);
                s3 = fma(sv, W.kc[idx], s3);
            }
            for (int o = 16; o > 0; o >>= 1) {
                h1 += __shfl_xor_sync(0xffffffffu, h1, o);
                h2 += __shfl_xor_sync(0xffffffffu, h2, o);
                s3 += __shfl_xor_sync(0xffffffffu, s3, o);
                if (MASK) {
                    sKm += __shfl_xor_sync(0xffffffffu, sKm, o);
                    sKm2 += __shfl_xor_sync(0xffffffffu, sKm2, o);
                    nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
                }
            }
            if (lane == 0) {
                int nobs;
                bool redo;
                // raw sums: pivot 0 and (P.q = 0) the unshifted kernel
                const float r = score_from_sums<MASK>(P, 0.0, h1, h2, h2, nmiss, s3, sKm, sKm2, nobs,
                                                      redo);
                const long long oi =
                    (long long)(Y - P.osy) * P.out_pitch + ((X - P.osx) - P.out_dlo);
                P.out[oi] = r;
                if (P.nobs) P.nobs[oi] = (unsigned short)nobs;
            }
        }
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

template <int KW, bool MASK>
static int launch_kw(const CUtensorMap &tmap, const PearsonParams &P, int grid, int threads,
                     size_t smem, cudaStream_t st) {
    auto kern = pearson_tiles<KW, MASK>;
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, st>>>(tmap, P);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

template <bool MASK>
static int launch_mask(int KW, const CUtensorMap &tmap, const PearsonParams &P, int grid,
                       int threads, size_t smem, cudaStream_t st) {
    switch (KW) {
#define CS_CASE(n) \
    case n:        \
        return launch_kw<n, MASK>(tmap, P, grid, threads, smem, st);
        CS_CASE(3) CS_CASE(5) CS_CASE(7) CS_CASE(9) CS_CASE(11) CS_CASE(13) CS_CASE(15) CS_CASE(17)
        CS_CASE(19) CS_CASE(21) CS_CASE(23) CS_CASE(25) CS_CASE(27) CS_CASE(29) CS_CASE(31)
#undef CS_CASE
        default:
            set_error("kernel width %d not supported (odd widths 3..31)", KW);
            return CS_ERR_INVALID;
    }
}

// device scratch holding the kernel tables of one launch; a small ring so that
// back-to-back launches do not race on it.
struct KtabRing {
    void *buf[8] = {nullptr};
    size_t cap[8] = {0};
    int next = 0;
};
static thread_local KtabRing g_ring;

}  // namespace cs

using namespace cs;

// kernels wider than 31: the one-warp-per-window kernel
static int pearson_wide_launch(const cs_layout *Li, const float *d_img, const cs_kernel_desc *K,
                               const cs_pearson_opts *opts, int32_t oy0, int32_t oy1, int32_t ox0,
                               int32_t ox1, int32_t odlo, int32_t odhi, const cs_layout *Lo,
                               float *d_out, uint16_t *d_nobs, cudaStream_t st) {
    PearsonParams P;
    memset(&P, 0, sizeof(P));
    const int nk = K->kh * K->kw;
    P.oy0 = oy0, P.oy1 = oy1, P.ox0 = ox0, P.ox1 = ox1;
    const int dmin_poss = ox0 - (oy1 - 1), dmax_poss = (ox1 - 1) - oy0;
    P.odlo = odlo < dmin_poss ? dmin_poss : odlo;
    P.odhi = odhi > dmax_poss ? dmax_poss : odhi;
    CS_REQUIRE(P.odhi >= P.odlo, "empty output diagonal range");
    P.KH = K->kh, P.KW = K->kw, P.N = nk;
    P.osy = opts->out_row_shift, P.osx = opts->out_col_shift;
    P.out = d_out, P.nobs = d_nobs;
    P.out_pitch = Lo->pitch;
    P.out_dlo = Lo->dense ? 0 : Lo->dlo;
    {
        bool ok = oy0 - P.osy >= 0 && ox0 - P.osx >= 0 && Lo->rows >= oy1 - P.osy &&
                  Lo->cols >= ox1 - P.osx;
        const int sh = P.osx - P.osy;
        if (ok && !Lo->dense) ok = Lo->dlo <= P.odlo - sh && Lo->dhi >= P.odhi - sh;
        if (!ok) {
            set_error("output image does not cover scores on diagonals [%d,%d]", P.odlo - sh,
                      P.odhi - sh);
            return CS_ERR_INVALID;
        }
    }
    if (opts->has_mask) CS_REQUIRE(K->k_mask && K->k2_mask, "mask kernels missing");
    double sumK = 0.0;
    for (int i = 0; i < nk; ++i) sumK += K->k_corr[i];
    P.q = 0.0;
    P.sumKp = sumK;
    P.sumKp2 = 0.0;
    P.ksum = K->k_sum, P.k2sum = K->k2_sum, P.kmean = K->k_mean, P.kstd = K->k_std;
    P.thr = opts->xcorr_threshold;
    P.invN = 1.0 / (double)nk;
    P.vK0 = opts->has_mask ? (K->k2_sum / (double)nk - K->k_mean * K->k_mean) : K->k_std * K->k_std;
    P.min_present = (int)((1.0 - opts->missing_tol) * (double)nk);
    P.kmean_zero = (K->k_mean == 0.0);
    P.has_mask = opts->has_mask;
    P.raw_xcorr = opts->raw_xcorr;
    P.nobs_full = opts->nobs_full;
    // tables: K_corr, K_mask, K2_mask as float64
    const size_t tbytes = (size_t)3 * nk * sizeof(double);
    std::vector<double> h((size_t)3 * nk, 0.0);
    for (int i = 0; i < nk; ++i) {
        h[i] = K->k_corr[i];
        if (opts->has_mask) {
            h[nk + i] = K->k_mask[i];
            h[2 * nk + i] = K->k2_mask[i];
        }
    }
    KtabRing &ring = g_ring;
    const int slot = ring.next;
    ring.next = (ring.next + 1) % 8;
    if (ring.cap[slot] < tbytes) {
        if (ring.buf[slot]) cudaFree(ring.buf[slot]);
        ring.buf[slot] = nullptr;
        ring.cap[slot] = 0;
        CS_CUDA(cudaMalloc(&ring.buf[slot], tbytes));
        ring.cap[slot] = tbytes;
    }
    CS_CUDA(cudaMemcpyAsync(ring.buf[slot], h.data(), tbytes, cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaStreamSynchronize(st));  // `h` is pageable and local
    WideParams Wp;
    Wp.img = d_img;
    Wp.pitch = Li->pitch;
    Wp.dense = Li->dense;
    Wp.dlo = Li->dense ? 0 : Li->dlo;
    Wp.dhi = Li->dense ? 0 : Li->dhi;
    Wp.kc = (const double *)ring.buf[slot];
    Wp.km = Wp.kc + nk;
    Wp.k2m = Wp.km + nk;
    const int nrows = oy1 - oy0;
    const int Wo = P.odhi - P.odlo + 1, ncols = ox1 - ox0;
    const int per_row = Wo < ncols ? Wo : ncols;
    dim3 grid((unsigned)(nrows < (1 << 20) ? nrows : (1 << 20)), (unsigned)((per_row + 7) / 8 > 64 ? 64 : (per_row + 7) / 8));
    if (opts->has_mask)
        pearson_wide<true><<<grid, 256, 0, st>>>(P, Wp);
    else
        pearson_wide<false><<<grid, 256, 0, st>>>(P, Wp);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// plan_only: stop after the tiling has been chosen and report the tile height
static int pearson_impl(const cs_layout *Li, const float *d_img, const cs_kernel_desc *K,
                        const cs_pearson_opts *opts, int32_t oy0, int32_t oy1, int32_t ox0,
                        int32_t ox1, int32_t odlo, int32_t odhi, const cs_layout *Lo, float *d_out,
                        uint16_t *d_nobs, void *stream, bool plan_only, int32_t *tile_rows_out) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(Li && K && opts && (plan_only || (d_img && Lo && d_out)),
               "cs_pearson_f32: null argument");
    CS_REQUIRE(K->kh >= 1 && K->kw >= 3 && (K->kh & 1) && (K->kw & 1),
               "kernel shape must be odd (got %dx%d)", K->kh, K->kw);
    CS_REQUIRE(K->kh <= 255 && K->kw <= 255, "kernel %dx%d too large (255x255 at most)", K->kh,
               K->kw);
    CS_REQUIRE(oy1 > oy0 && ox1 > ox0, "empty output region");
    const int kh = (K->kh - 1) / 2, kw = (K->kw - 1) / 2;
    CS_REQUIRE(oy0 - kh >= 0 && oy1 + kh <= Li->rows && ox0 - kw >= 0 && ox1 + kw <= Li->cols,
               "output region needs windows outside the image");
    if (K->kh > 31 || K->kw > 31) {
        if (plan_only) {
            if (tile_rows_out) *tile_rows_out = 32;
            return CS_OK;
        }
        return pearson_wide_launch(Li, d_img, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, Lo, d_out,
                                   d_nobs, st);
    }
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return CS_ERR_CUDA;
    }

    PearsonParams P;
    memset(&P, 0, sizeof(P));
    P.rows = Li->rows;
    P.cols = Li->cols;
    P.dense = Li->dense;
    P.dlo = Li->dense ? 0 : Li->dlo;
    P.dhi = Li->dense ? 0 : Li->dhi;
    P.oy0 = oy0;
    P.oy1 = oy1;
    P.ox0 = ox0;
    P.ox1 = ox1;
    // clip the output diagonal range to what the region can contain
    int dmin_poss = ox0 - (oy1 - 1), dmax_poss = (ox1 - 1) - oy0;
    if (odlo < dmin_poss) odlo = dmin_poss;
    if (odhi > dmax_poss) odhi = dmax_poss;
    CS_REQUIRE(odhi >= odlo, "empty output diagonal range");
    P.odlo = odlo;
    P.odhi = odhi;
    P.KH = K->kh;
    P.KW = K->kw;
    P.KWP2 = round_up(K->kw + 1, 4);
    P.N = K->kh * K->kw;
    const int kwa = round_up(kw, 4);
    const int nrows_out = oy1 - oy0;

    // ---- tables -------------------------------------------------------------------
    // float: K' = K_corr - q as two padded rows per kernel row (see Ktab), [KH][2][KWP2]
    // double (mask): row prefix sums of K_mask and K2_mask [KH][KW+1] each, column sums [KW] each
    const int nk = K->kh * K->kw;
    double qd = 0.0;
    for (int i = 0; i < nk; ++i) qd += K->k_corr[i];
    qd /= nk;
    const float qf = (float)qd;
    P.n_ftab = 2 * K->kh * P.KWP2;
    // declared-missing strip: tables over the output diagonals whose windows touch it
    P.sdlo = 0;
    P.sdhi = -1;
    P.st_base = 0;
    P.st_n = 0;
    if (opts->has_mask && opts->strip_dhi >= opts->strip_dlo) {
        P.sdlo = opts->strip_dlo;
        P.sdhi = opts->strip_dhi;
        P.st_base = P.sdlo - (kh + kw);
        P.st_n = (P.sdhi + (kh + kw)) - P.st_base + 1;
    }
    P.n_dtab = opts->has_mask ? (2 * (K->kh + 1) * (K->kw + 1) + 2 * K->kw + 3 * P.st_n) : 0;
    if (opts->has_mask) CS_REQUIRE(K->k_mask && K->k2_mask, "mask kernels missing");
    const size_t fbytes = (size_t)round_up(P.n_ftab, 4) * sizeof(float);
    const size_t kbytes = (fbytes + (size_t)P.n_dtab * sizeof(double) + 15) / 16 * 16;
    P.tab_bytes = (int)kbytes;
    unsigned char *hk = (unsigned char *)malloc(kbytes);
    if (!hk) return CS_ERR_NOMEM;
    memset(hk, 0, kbytes);
    float *hf = (float *)hk;
    double *hd = (double *)(hk + fbytes);
    double sumKp = 0.0, sumKp2 = 0.0;
    for (int i = 0; i < K->kh; ++i)
        for (int j = 0; j < K->kw; ++j) {
            const float v = (float)(K->k_corr[i * K->kw + j] - (double)qf);
            hf[(2 * i) * P.KWP2 + j] = v;          // (k0,k1),(k2,k3),...
            hf[(2 * i + 1) * P.KWP2 + j + 1] = v;  // (0,k0),(k1,k2),...
            sumKp += (double)v;
            sumKp2 += (double)v * (double)v;
        }
    if (opts->has_mask) {
        const int kw1 = K->kw + 1;
        double *ik = hd, *ik2 = ik + (K->kh + 1) * kw1;
        double *kc = ik2 + (K->kh + 1) * kw1, *k2c = kc + K->kw;
        double *stk = k2c + K->kw, *stk2 = stk + P.st_n, *stc = stk2 + P.st_n;
        // 2-D prefix tables: ik[i][j] = sum of k_mask[i' < i][j' < j]
        for (int i = 0; i < K->kh; ++i) {
            double a = 0.0, b = 0.0;
            for (int j = 0; j < K->kw; ++j) {
                a += K->k_mask[i * K->kw + j];
                b += K->k2_mask[i * K->kw + j];
                ik[(i + 1) * kw1 + j + 1] = ik[i * kw1 + j + 1] + a;
                ik2[(i + 1) * kw1 + j + 1] = ik2[i * kw1 + j + 1] + b;
                kc[j] += K->k_mask[i * K->kw + j];
                k2c[j] += K->k2_mask[i * K->kw + j];
            }
        }
        // window centred on diagonal d: tap (i, j) sits on diagonal d + (j - kw) - (i - kh)
        for (int sd = 0; sd < P.st_n; ++sd) {
            const int d = P.st_base + sd;
            for (int i = 0; i < K->kh; ++i)
                for (int j = 0; j < K->kw; ++j) {
                    const int dt = d + (j - kw) - (i - kh);
                    if (dt >= P.sdlo && dt <= P.sdhi) {
                        stk[sd] += K->k_mask[i * K->kw + j];
                        stk2[sd] += K->k2_mask[i * K->kw + j];
                        stc[sd] += 1.0;
                    }
                }
        }
    }

    // ---- tiling ---------------------------------------------------------------
    const int Wo = odhi - odlo + 1;  // output diagonals
    const int ncols_out = ox1 - ox0;
    // banded traversal when the output band is narrow relative to the region
    P.skew = (Wo + 16 < ncols_out) ? 1 : 0;
    int TR = opts->tile_rows > 0 ? round_up(opts->tile_rows, RU) : 32;
    if (const char *e = getenv("CS_TILE_ROWS"))  // tuning knob for experiments
        if (atoi(e) > 0) TR = round_up(atoi(e), RU);
    if (TR > round_up(nrows_out, RU)) TR = round_up(nrows_out, RU);
    size_t smem = 0;
    int NBc = 0, nchunks = 0, IC = 0, IR = 0, NW = 0, threads = 0;
    for (;; TR -= RU) {
        if (TR < RU) {
            free(hk);
            set_error("kernel %dx%d does not fit in shared memory", K->kh, K->kw);
            return CS_ERR_INVALID;
        }
        const int G = TR / RU;
        const int span = P.skew ? (Wo + RU - 1 + 3 + kSkewSlack) : (ncols_out + 3);
        const int nblk_total = (span + RT - 1) / RT;
        const int nb_max = (256 - 2 * kwa - skew_shift(G - 1) * P.skew) / RT;
        if (nb_max < 1) continue;
        // aim at <= 256 items per tile and a box of at most 256 columns
        int nb_want = 256 / G;
        if (nb_want > nb_max) nb_want = nb_max;
        if (nb_want < 1) nb_want = 1;
        nchunks = (nblk_total + nb_want - 1) / nb_want;
        NBc = (nblk_total + nchunks - 1) / nchunks;
        IC = RT * NBc + skew_shift(G - 1) * P.skew + 2 * kwa;
        IR = TR + K->kh - 1;
        if (IR > 256 || IC > 256) continue;
        threads = round_up(8 * ((NBc + 3) / 4) * 2 * ((G + 3) / 4), 32);  // padded item count
        if (threads > 256) threads = 256;
        if (threads < 64) threads = 64;
        NW = (IC * IR + 31) / 32 + 4;  // words of the linear bit array
        size_t o = (size_t)IC * IR * sizeof(float);
        o = (o + 15) / 16 * 16;
        P.off_bits = (int)o;
        if (opts->has_mask) o += (size_t)NW * sizeof(uint32_t);
        o = (o + 15) / 16 * 16;
        P.off_K = (int)o;
        o += fbytes;
        P.off_D = (int)o;
        o += (size_t)P.n_dtab * sizeof(double);
        o = (o + 15) / 16 * 16;
        P.off_stat = (int)o;
        o += (size_t)3 * RU * RT * threads * sizeof(float);
        o = (o + 15) / 16 * 16;
        P.off_grp = (int)o;
        if (opts->has_mask) o += (size_t)kMaxGroups * threads * sizeof(unsigned long long);
        P.off_bar = (int)o;
        o += 16;
        smem = o;
        if (smem <= 113 * 1024 || (TR == RU && smem <= 227 * 1024)) break;
    }
    if (plan_only) {
        free(hk);
        if (tile_rows_out) *tile_rows_out = TR;
        return CS_OK;
    }
    P.TR = TR;
    P.G = TR / RU;
    P.NBc = NBc;
    P.nchunks = nchunks;
    P.IC = IC;
    P.IR = IR;
    P.NW = NW;
    const int nrb = (nrows_out + TR - 1) / TR;
    const long long grid_ll = (long long)nrb * nchunks;
    if (grid_ll >= (1ll << 31)) {
        free(hk);
        set_error("grid too large");
        return CS_ERR_INVALID;
    }

    // ---- output -----------------------------------------------------------------
    P.osy = opts->out_row_shift;
    P.osx = opts->out_col_shift;
    P.out = d_out;
    P.nobs = d_nobs;
    P.out_pitch = Lo->pitch;
    P.out_dlo = Lo->dense ? 0 : Lo->dlo;
    {
        bool ok = oy0 - P.osy >= 0 && ox0 - P.osx >= 0 && Lo->rows >= oy1 - P.osy &&
                  Lo->cols >= ox1 - P.osx;
        const int sh = P.osx - P.osy;
        // output pixel (y, x) = (Y - osy, X - osx); its diagonal is d - (osx - osy)
        if (ok && !Lo->dense) ok = Lo->dlo <= odlo - sh && Lo->dhi >= odhi - sh;
        if (!ok) {
            free(hk);
            set_error("output image does not cover scores on diagonals [%d,%d]", odlo - sh,
                      odhi - sh);
            return CS_ERR_INVALID;
        }
    }

    KtabRing &ring = g_ring;
    const int slot = ring.next;
    ring.next = (ring.next + 1) % 8;
    if (ring.cap[slot] < kbytes) {
        if (ring.buf[slot]) cudaFree(ring.buf[slot]);
        ring.buf[slot] = nullptr;
        ring.cap[slot] = 0;
        if (cudaMalloc(&ring.buf[slot], kbytes) != cudaSuccess) {
            free(hk);
            set_error("cudaMalloc of the kernel tables failed");
            return CS_ERR_NOMEM;
        }
        ring.cap[slot] = kbytes;
    }
    // pageable source: the copy is staged by the runtime before the call returns
    cudaError_t ce = cudaMemcpyAsync(ring.buf[slot], hk, kbytes, cudaMemcpyHostToDevice, st);
    free(hk);
    CS_CUDA(ce);
    P.ftab = (const float *)ring.buf[slot];
    P.dtab = (const double *)((const unsigned char *)ring.buf[slot] + fbytes);
    P.q = (double)qf;
    P.sumKp = sumKp;
    P.sumKp2 = sumKp2;
    P.ksum = K->k_sum;
    P.k2sum = K->k2_sum;
    P.kmean = K->k_mean;
    P.kstd = K->k_std;
    P.thr = opts->xcorr_threshold;
    P.invN = 1.0 / (double)P.N;
    P.vK0 = opts->has_mask ? (K->k2_sum / (double)P.N - K->k_mean * K->k_mean)
                           : K->k_std * K->k_std;
    P.min_present = (int)((1.0 - opts->missing_tol) * (double)P.N);
    P.kmean_zero = (K->k_mean == 0.0);
    P.has_mask = opts->has_mask;
    P.dbg = getenv("CS_DEBUG_SKIP") ? atoi(getenv("CS_DEBUG_SKIP")) : 0;
    P.cnt = nullptr;
    static unsigned long long *g_cnt = nullptr;
    if (getenv("CS_DEBUG_COUNT")) {
        if (!g_cnt) cudaMalloc(&g_cnt, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(g_cnt, 0, 16 * sizeof(unsigned long long), st);
        P.cnt = g_cnt;
    }
    P.raw_xcorr = opts->raw_xcorr;
    P.nobs_full = opts->nobs_full;

    // ---- tensor map -----------------------------------------------------------------
    CUtensorMap tmap;
    {
        // dimension 0 = X' = X - dlo, dimension 1 = image row
        cuuint64_t dims[2] = {(cuuint64_t)(Li->cols - P.dlo), (cuuint64_t)Li->rows};
        cuuint64_t strides[1] = {(cuuint64_t)Li->pitch * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)IC, (cuuint32_t)IR};
        cuuint32_t estr[2] = {1, 1};
        CS_REQUIRE(Li->pitch % 4 == 0, "image pitch must be a multiple of 4");
        CS_REQUIRE(((uintptr_t)d_img & 15) == 0, "image must be 16-byte aligned");
        CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)d_img, dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu pitch %d box %d x %d",
                      (int)cr, (unsigned long long)dims[0], (unsigned long long)dims[1], Li->pitch,
                      IC, IR);
            return CS_ERR_CUDA;
        }
    }
    int lrc = opts->has_mask ? launch_mask<true>(K->kw, tmap, P, (int)grid_ll, threads, smem, st)
                             : launch_mask<false>(K->kw, tmap, P, (int)grid_ll, threads, smem, st);
    if (P.cnt && lrc == CS_OK) {
        unsigned long long h[16];
        cudaMemcpyAsync(h, P.cnt, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr,
                "mask stats: blocks %llu, with missing %llu, with groups %llu, row-by-row %llu, groups "
                "%llu (full rows %llu, rows in groups %llu), blocks with full columns %llu\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
    return lrc;
}

extern "C" int cs_pearson_f32(const cs_layout *Li, const float *d_img, const cs_kernel_desc *K,
                              const cs_pearson_opts *opts, int32_t oy0, int32_t oy1, int32_t ox0,
                              int32_t ox1, int32_t odlo, int32_t odhi, const cs_layout *Lo,
                              float *d_out, uint16_t *d_nobs, void *stream) {
    return pearson_impl(Li, d_img, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, Lo, d_out, d_nobs,
                        stream, false, nullptr);
}

extern "C" int cs_pearson_tile_rows(const cs_layout *Li, const cs_kernel_desc *K,
                                    const cs_pearson_opts *opts, int32_t oy0, int32_t oy1,
                                    int32_t ox0, int32_t ox1, int32_t odlo, int32_t odhi,
                                    int32_t *tile_rows) {
    CS_REQUIRE(tile_rows, "cs_pearson_tile_rows: null argument");
    return pearson_impl(Li, nullptr, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, nullptr, nullptr,
                        nullptr, nullptr, true, tile_rows);
}
