// K1: fused sliding-window Pearson correlation of a small dense kernel against a
// banded / dense float32 image (detection.py:917-1131 of the reference).
//
// One CTA = one output tile of TR rows that follows the band (a parallelogram in
// matrix coordinates), possibly one of several column chunks.  The input tile
// ((TR+KH-1) rows x IC columns, matrix coordinates) is fetched with ONE TMA box
// load out of the skewed band (the row stride of the tensor map is `pitch`, see
// cs_layout); the kernel tables arrive by a bulk copy on the same mbarrier.  Then,
// all in shared memory / registers:
//   fix-up   out-of-band aliases of the box (two triangles) -> 0; with a pixel mask
//            (MODE_BITS) also NaN sentinels -> bit array + value 0;
//   per thread, a block of RU x RT = 2 x 8 windows (footprint 18 x 24 pixels for 17 x 17):
//   pivot    mean of the middle footprint row; all sums are taken on S - pivot (exact
//            algebra, keeps float32 accurate);
//   FMA pass per input row one LDS.128 sweep of the 24-pixel segment, then per window 9
//            packed fma.rn.f32x2 whose two halves collect alternate taps (aligned register
//            pairs: no shuffles, no duplicated taps) -- no mask work in this loop;
//   sums     packed column sums of S - pivot and its square over KH rows, then KW-wide
//            sliding sums along the row;
//   mask     (MODE_GEO: the mask of make_missing_mask / frame_missing_mask, given as bit
//            vectors of missing rows and columns + a diagonal band + a missing strip)
//            missing count and the two mask-kernel sums of every window from row / column
//            prefix tables of the centred kernel: one table look-up per missing row or
//            column of the footprint, none for the 70 % of blocks without any;
//   score    the reference's formulas in centred float32 algebra, eight windows at a time
//            (independent chains), float32 score + uint8 missing count, 16-byte stores;
//   exact    windows the float32 path cannot decide (ill-conditioned, a 1e-4 threshold of
//            xcorr2 within rounding distance, footprint on the frame's margins or -- with
//            a pixel mask -- on a missing pixel) are redone by the whole warp in float64
//            from the tile with the per-pixel mask predicate.
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

enum { MODE_NOMASK = 0, MODE_BITS = 1, MODE_GEO = 2 };  // = cs_pearson_opts.mask_mode

struct PearsonParams {
    // image
    int rows, cols, dlo, dhi, dense;
    // output pixel set (image coordinates)
    int oy0, oy1, ox0, ox1, odlo, odhi;
    // tiling
    int TR, G, NBc, nchunks, skew, nrb;
    int IC, IR, NW;
    int fixup;  // banded image: the box aliases neighbouring rows' pixels outside the band
    // kernel geometry
    int KH, KW, KWP2, N;
    // output image
    float *out;
    void *nmiss;  // uint8 / uint16 plane (nmiss16) of missing counts, may be null
    int nmiss16;
    int out_pitch, out_dlo, osy, osx;
    // tables (device): one linear block copied to shared memory by every CTA
    const unsigned char *tab;
    int tab_bytes;
    const double *dK;  // global, float64: K_corr, K_mask, K2_mask [N each] (exact path)
    // float32 constants of the centred algebra (K' = K - qf, S' = S - pivot)
    float qf, sumKp, sumKp2, ksump, k2sump, delta, invN, thr, fillc, escale, den_floor;
    // float64 constants of the exact path
    double ksum, k2sum, kmean, kstd, thr_d, invN_d, vK0, sumKc_d;
    int min_present, kmean_zero, raw_xcorr, nobs_full;
    // geometric mask (image coordinates)
    const uint32_t *rbits, *cbits;  // missing image rows / columns inside the matrix
    int mlo, mhi;                   // diagonals on which a missing bin flags its pixels
    int my0, my1, mx0, mx1;         // the matrix inside the frame
    int margin_mode;                // 0 no frame, 1 banded frame (pre:461-477), 2 all four margins
    int top_x1, right_y0;           // banded frame: top margin columns < top_x1, right margin rows >= right_y0
    int sdlo, sdhi, st_base, st_n;  // missing diagonal strip (pre:483-497) and its tables
    // shared memory carve-up (bytes)
    int off_bits, off_K, off_PR, off_PC, off_ST, off_rb, off_cb, off_bar;
#ifdef CS_ABLATE
    unsigned long long *cnt;  // statistics (windows on the exact path, ...)
    int dbg;                  // phase switches for timing experiments (results are wrong)
#endif
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
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ double thr0(double v, double t) { return fabs(v) < t ? 0.0 : v; }  // [sec:scorefn]

// Squared error amplification above which a window leaves the float32 path: the
// float32 sums carry ~1e-6 relative error on well-conditioned windows, and the score
// error grows like amp = rms(S') * rms(K') / (sigma_S * sigma_K).
constexpr float kAmpLimit2 = 8.0f;
// ... and the same for the tap products, which are accumulated on the uncentred pixels: their
// rounding error scales with rms(S) rms(K') / (sigma_S sigma_K) / N, two orders of magnitude
// below the one of the variance, hence the much wider limit on the squared ratio.
constexpr float kAmpLimitQ = 1000.0f;

// The reference's formulas (det:1002-1020 no mask, det:1021-1092 masked) in float64 from the
// raw window sums: h1, h2 = sum S, sum S^2 (missing pixels count as S = 0), s3 = sum S * K_corr,
// nmiss, sKm / sKm2 = sums of the mask kernels (K and K^2) over the missing pixels.
template <bool MASK>
__device__ __forceinline__ double exact_score_core(const PearsonParams &P, double h1, double h2,
                                                   double s3, int nmiss, double sKm, double sKm2,
                                                   int &nmiss_out) {
    nmiss_out = 0;
    if (P.raw_xcorr) return thr0(s3, P.thr_d);
    const double A1 = thr0(h1 * P.invN_d, P.thr_d);
    const double A2 = thr0(h2 * P.invN_d, P.thr_d);
    const double A3 = thr0(s3 * P.invN_d, P.thr_d);
    double cov, den2;
    bool ok = true;
    if (!MASK || nmiss == 0) {
        const double vS = fma(-A1, A1, A2);
        cov = fma(-A1, P.kmean, A3);
        den2 = vS * P.vK0;
        ok = vS >= 0.0;
    } else {
        const int npres = P.N - nmiss;
        const double f = (double)P.N / (double)npres;
        sKm = thr0(sKm, P.thr_d);
        sKm2 = thr0(sKm2, P.thr_d);
        const double mK = (P.ksum - sKm) / (double)npres;
        const double m2K = (P.k2sum - sKm2) / (double)npres;
        const double mS = A1 * f;
        const double vS = fma(A2, f, -mS * mS);
        cov = fma(-A1, mK, A3) * f;
        den2 = vS * fma(-mK, mK, m2K);
        ok = (npres > 0) && (npres >= P.min_present) && !P.kmean_zero;
        if (P.nobs_full && npres != 0) nmiss_out = nmiss;
    }
    // det:1066,1088-1091: denom = sqrt(den2); |denom| < 1e-10 or NaN -> 0
    double r = 0.0;
    if (ok && den2 >= 1e-20 && den2 < 1e300) {
        r = fmin(1.0, fmax(-1.0, cov / sqrt(den2)));
        if (!(r == r)) r = 0.0;
    }
    return r;
}
template <bool MASK>
__device__ __forceinline__ float exact_score(const PearsonParams &P, double h1, double h2,
                                             double s3, int nmiss, double sKm, double sKm2,
                                             int &nmiss_out) {
    return (float)exact_score_core<MASK>(P, h1, h2, s3, nmiss, sKm, sKm2, nmiss_out);
}
// `nb` (<= 64) bits of a bit array (one bit per position, 32 per word) from bit `pos` on
__device__ __forceinline__ unsigned long long bits64(const uint32_t *bits, int pos) {  // [sec:maskfn]
    const int w = pos >> 5, sh = pos & 31;
    const uint32_t a = bits[w], b = bits[w + 1], c = bits[w + 2];
    const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t bits32(const uint32_t *bits, int pos) {
    const int w = pos >> 5, sh = pos & 31;
    return __funnelshift_r(bits[w], bits[w + 1], sh);
}

// per-pixel missing predicate of the geometric mask (image coordinates), SURVEY 3.3
__device__ __forceinline__ bool geo_missing(const PearsonParams &P, int Y, int X, bool rbit,
                                            bool cbit) {
    const int d = X - Y;
    if (d >= P.sdlo && d <= P.sdhi) return true;  // pre:483-497
    const bool inside = Y >= P.my0 && Y < P.my1 && X >= P.mx0 && X < P.mx1;
    if (inside) return (rbit || cbit) && d >= P.mlo && d <= P.mhi;
    if (P.margin_mode == 2) return true;  // pre:464-466, 479-480
    if (P.margin_mode == 1)
        return (Y < P.my0 && X < P.top_x1) ||      // pre:461-463, 477
               (X >= P.mx1 && Y >= P.right_y0);    // pre:475
    return false;
}

// ---------------------------------------------------------------- the kernel
template <int KW, int MODE>
__global__ void __launch_bounds__(256, 2)
pearson_tiles(const __grid_constant__ CUtensorMap tmap, const PearsonParams P) {
    constexpr bool MASK = MODE != MODE_NOMASK;
    constexpr int kw = (KW - 1) / 2;
    constexpr int kwa = (kw + 3) / 4 * 4;
    constexpr int off = kwa - kw;
    constexpr int XW = RT + KW - 1;           // pixels of one footprint row
    constexpr int NQ = (off + XW + 3) / 4;    // float4 loads per footprint row
    constexpr int KWP2 = (KW + 1 + 3) / 4 * 4;  // floats of one padded tap row
    constexpr int NP = (KW + 1) / 2;            // tap pairs per kernel row
    static_assert(((off + RT - 1) >> 1) + NP <= 2 * NQ, "tap pairs run past the loaded segment");
    constexpr unsigned KWMASK = (KW == 32) ? 0xffffffffu : ((1u << KW) - 1u);
    constexpr int KW1 = KW + 1;

    extern __shared__ __align__(1024) unsigned char smem[];
    float *__restrict__ tile = reinterpret_cast<float *>(smem);
    uint32_t *__restrict__ bits = reinterpret_cast<uint32_t *>(smem + P.off_bits);
    // taps K' = K_corr - q, two rows of KWP2 floats per kernel row: the row padded with
    // zeros (pairs (k0,k1),(k2,k3),... for windows at an even tile column) and the same row
    // shifted right by one (pairs (0,k0),(k1,k2),... for windows at an odd tile column)
    const float *__restrict__ Ktab = reinterpret_cast<const float *>(smem + P.off_K);
    // prefix tables of the centred mask kernels (K' = K_mask - q, K2' = K2_mask - 2 q K_mask + q^2)
    // as (K', K2') pairs: PR[i][j] = sum over j' < j of row i, PC[j][i] = sum over i' < i of column j
    const float2 *__restrict__ PR = reinterpret_cast<const float2 *>(smem + P.off_PR);
    const float2 *__restrict__ PC = reinterpret_cast<const float2 *>(smem + P.off_PC);
    // the missing strip per window diagonal: (count, sum K', sum K2', -)
    const float4 *__restrict__ ST = reinterpret_cast<const float4 *>(smem + P.off_ST);
    uint32_t *__restrict__ rbw = reinterpret_cast<uint32_t *>(smem + P.off_rb);
    uint32_t *__restrict__ cbw = reinterpret_cast<uint32_t *>(smem + P.off_cb);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P.off_bar);

    const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
    // row tiles are taken alternately from both ends of the region: the tiles on the frame's
    // margins (many exact-path windows) start first instead of forming the tail of the grid
    const int rbi = blockIdx.x / P.nchunks;
    const int ch = blockIdx.x - rbi * P.nchunks;
    const int rb = (rbi & 1) ? (P.nrb - 1 - (rbi >> 1)) : (rbi >> 1);
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
    const int TX = TXp + P.dlo;  // image column of tile column 0
    const int IC = P.IC, IR = P.IR, NW = P.NW;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        fence_proxy_async();
        // the tile (2-D tensor box) and the kernel tables (one linear block) arrive through
        // the same barrier
        mbar_expect_tx(bar, (uint32_t)(IC * IR * sizeof(float)) + (uint32_t)P.tab_bytes);
        tma_load_2d(tile, &tmap, bar, TXp, TY);
        bulk_load_1d(smem + P.off_K, P.tab, (uint32_t)P.tab_bytes, bar);
    }
    if (MODE == MODE_BITS)
        for (int i = tid; i < NW; i += nthr) bits[i] = 0u;
    if (MODE == MODE_GEO) {
        // missing-row / missing-column bits of the tile: bit k of the array = tile row / column k
        // (the global vectors carry zero words on both sides)
        const int nrw = (IR + 31) / 32 + 3, ncw = (IC + 31) / 32 + 4;
        for (int i = tid; i < nrw + ncw; i += nthr) {
            const bool isr = i < nrw;
            const int w = isr ? i : i - nrw;
            const int pos = (isr ? TY : TX) + 32 * w;  // image row / column of the word's bit 0
            const int lim = isr ? P.rows : P.cols;
            uint32_t v = 0u;
            if (pos + 32 > 0 && pos < lim) {
                const uint32_t *src = isr ? P.rbits : P.cbits;
                const int gw = pos >> 5, sh = pos & 31;  // arithmetic shift: floor for pos < 0
                v = __funnelshift_r(src[gw], src[gw + 1], sh);
            }
            (isr ? rbw : cbw)[w] = v;
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // ---- phase A: fix-up of the tile ------------------------------------------------ [sec:A1]
    if (MODE == MODE_BITS) {
        // out-of-band aliases -> 0, NaN sentinels -> bit array and 0, four pixels per thread
        const int ICq4 = IC >> 2;
        for (int iy = tid >> 5; iy < IR; iy += nthr >> 5) {
            const int dbase = TX - (TY + iy);  // diagonal of tile column 0
            int clo = 0, chi = IC;
            if (!P.dense) {
                clo = max(P.dlo - dbase, 0);
                chi = min(P.dhi - dbase + 1, IC);
            }
            unsigned rowword;  // keep-bits of the row, 32 columns per lane (lanes 0..7)
            {
                const int base = 32 * lane;
                const int a = min(max(clo - base, 0), 32), b = min(max(chi - base, 0), 32);
                rowword = (unsigned)(((1ull << b) - 1ull) & ~((1ull << a) - 1ull));
            }
            for (int cb = 0; cb < ICq4; cb += 32) {
                const int c4 = cb + lane;
                const int c0 = 4 * c4;
                const unsigned kwd = __shfl_sync(0xffffffffu, rowword, (c0 >> 5) & 31);
                if (c4 >= ICq4) continue;
                float4 *ptr = reinterpret_cast<float4 *>(tile + iy * IC) + c4;
                const float4 v = *ptr;
                const unsigned keep = (kwd >> (c0 & 31)) & 0xfu;
                unsigned live = keep;
                if (v.x != v.x) live &= ~1u;
                if (v.y != v.y) live &= ~2u;
                if (v.z != v.z) live &= ~4u;
                if (v.w != v.w) live &= ~8u;
                const unsigned nb = keep & ~live;
                if (live != 0xfu)
                    *ptr = make_float4((live & 1u) ? v.x : 0.f, (live & 2u) ? v.y : 0.f,
                                       (live & 4u) ? v.z : 0.f, (live & 8u) ? v.w : 0.f);
                if (nb) atomicOr(&bits[(iy * IC + c0) >> 5], nb << ((iy * IC + c0) & 31));
            }
        }
        __syncthreads();
    } else if (P.fixup) {
        // only the aliases of the box outside the stored band: two triangles of at most
        // IR + 3 columns at the ends of the rows; one warp per tile row (not needed when the
        // band is stored with a gap of zeros that wide, cs_layout_band_padded)
#ifdef CS_ABLATE
        if (!(P.dbg & 16))
#endif
        for (int iy = tid >> 5; iy < IR; iy += nthr >> 5) {
            const int dbase = TX - (TY + iy);
            const int clo = min(max(P.dlo - dbase, 0), IC);
            const int chi = max(min(P.dhi - dbase + 1, IC), 0);
            for (int c = lane; c < clo; c += 32) tile[iy * IC + c] = 0.f;
            for (int c = chi + lane; c < IC; c += 32) tile[iy * IC + c] = 0.f;
        }
        __syncthreads();
    }

    // ---- blocks of RU x RT windows, one per thread -------------------------------------
    // Lanes 0-3 / 4-7 of every quarter-warp take blocks of row groups g / g + 2: in a banded
    // traversal their tile columns differ by 4 (mod 8), which makes the 16-byte loads of a
    // quarter-warp hit disjoint banks.
    const int NQd = (P.NBc + 3) >> 2;
    const int nitems = 8 * NQd * 2 * ((P.G + 3) >> 2);
    const float invN = P.invN;
    // every lane runs every loop (work is predicated): the exact path is warp-wide
    for (int base = 0; base < nitems; base += nthr) {
        int g, m;
        bool live;
        {
            const int idx = base + tid;
            // a warp = 4 column blocks x 8 row groups
            const int npair = 2 * ((P.G + 3) >> 2);
            const int rest = idx >> 3, pr = rest % npair, M = rest / npair;
            g = (pr >> 1) * 4 + (pr & 1) + 2 * ((idx >> 2) & 1);
            m = 4 * M + (idx & 3);
            live = idx < nitems && g < P.G && m < P.NBc;
            if (!live) g = m = 0;
        }
        const int cxa = skew_shift(g) * P.skew + RT * m;  // aligned tile column of x[0]
        const int X0 = TX + cxa + kwa;                    // image column of window t = 0
        const int Yg = Y0 + RU * g;
        const int fc0 = cxa + off;                        // tile column of footprint column 0
        // valid windows of the block, bit k = u * RT + t
        unsigned okb = 0u;
        if (live) {
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int Y = Yg + u;
                const int tlo = max(max(P.ox0, Y + P.odlo) - X0, 0);
                const int thi = min(min(P.ox1 - 1, Y + P.odhi) - X0, RT - 1);
                if (Y < P.oy1 && thi >= tlo)
                    okb |= (((2u << thi) - 1u) & ~((1u << tlo) - 1u)) << (u * RT);
            }
        }
        const bool any = okb != 0u;
        unsigned slow = 0u;  // windows for the exact path
        float pl = 0.f;
        // sum S' K' of the windows of row u = 0 (s3a) and u = 1 (s3b)
        static_assert(RU == 2, "the epilogue is written for two window rows per thread");
        float s3a[RT], s3b[RT];
#pragma unroll
        for (int t = 0; t < RT; ++t) s3a[t] = s3b[t] = 0.f;
        if (any) {                                                           // [sec:pivot]
            // block pivot: mean of the middle footprint row (of three rows -- a quarter, the middle
            // and three quarters down -- for kernels of at most 9 columns, whose single row of
            // <= 16 pixels is a noisy estimate: fewer windows look ill-conditioned).  The sums of S
            // and S^2 are formed on S - pl (any pivot is exact; a close one keeps float32 accurate).
            constexpr int NPR = (KW <= 9) ? 3 : 1;
            const int fr_ = KH + RU - 1;
            float sacc = 0.f;
#pragma unroll
            for (int pr_ = 0; pr_ < NPR; ++pr_) {
                const int prow = NPR == 1 ? fr_ / 2 : (fr_ * (pr_ + 1)) / 4;
                const float4 *rp4 = reinterpret_cast<const float4 *>(tile + (RU * g + prow) * IC + cxa);
#pragma unroll
                for (int qd = 0; qd < NQ; ++qd) {
                    const float4 v = rp4[qd];
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * qd + e >= off && 4 * qd + e < off + XW) sacc += vv[e];
                }
            }
            pl = sacc * (1.0f / (float)(NPR * XW));
        }
#ifdef CS_ABLATE
        if (any && !(P.dbg & 8))
#else
        if (any)
#endif
        {                                                                    // [sec:main]
            // one packed accumulator per window: .lo and .hi collect alternate taps.  The products
            // are taken on the pixels as stored: K' sums to ~0, so the pivot enters as one
            // correction per window (- pl * sum K') and costs no arithmetic in this loop; the
            // uncentred products only matter for windows with rms(S) >> sigma(S), which the
            // conditioning test of the score sends to the exact path (kAmpLimitQ).
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
                // xe[q] = (x[2q], x[2q+1])
                unsigned long long xe[2 * NQ];
#pragma unroll
                for (int qd = 0; qd < NQ; ++qd) {
                    const ulonglong2 v = rp[qd];
                    xe[2 * qd] = v.x;
                    xe[2 * qd + 1] = v.y;
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
            const float corr = -pl * P.sumKp;  // sum (S - pl) K' = sum S K' - pl sum K'
#pragma unroll
            for (int t = 0; t < RT; ++t) {
                float lo, hi;
                unpack2(acc[0][t], lo, hi);
                s3a[t] = (lo + hi) + corr;
                unpack2(acc[1][t], lo, hi);
                s3b[t] = (lo + hi) + corr;
            }
        }

        // ---- sums, mask sums and scores, one row of RT windows at a time ---------------
        const unsigned long long npl2 = pack2(-pl, -pl);
        unsigned long long cs[2 * NQ], cq[2 * NQ];
#pragma unroll
        for (int q = 0; q < 2 * NQ; ++q) cs[q] = cq[q] = 0ull;
        if (any) {                                                           // [sec:sums]
            // column sums of (S - pl) and (S - pl)^2 over the KH rows of window row u = 0
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
        }
        // mask geometry of the block                                        [sec:masksum]
        uint32_t Rfp = 0u;             // missing rows of the footprint (bit = footprint row)
        unsigned long long Cfp = 0ull; // missing columns of the footprint
        unsigned long long bor0 = 0ull, bor1 = 0ull;  // MODE_BITS: columns with a missing pixel, per window row
        bool stripz = false;
        const int d00 = X0 - Yg;       // diagonal of window (u = 0, t = 0)
        if (MODE == MODE_GEO && any) {
            constexpr unsigned long long FWMASK = (XW >= 64) ? ~0ull : ((1ull << XW) - 1ull);
            const int fr = KH + RU - 1;
            Rfp = bits32(rbw, RU * g) & ((fr >= 32) ? 0xffffffffu : ((1u << fr) - 1u));
            Cfp = bits64(cbw, fc0) & FWMASK;
            // windows whose diagonal touches the strip tables
            stripz = (d00 + RT - 1 >= P.st_base) && (d00 - (RU - 1) < P.st_base + P.st_n);
#ifdef CS_ABLATE
            if (P.dbg & 1) Rfp = 0u, Cfp = 0ull, stripz = false;
#endif
        }
        if (MODE == MODE_BITS && any) {
            constexpr unsigned long long FWMASK = (XW >= 64) ? ~0ull : ((1ull << XW) - 1ull);
            unsigned long long mid = 0ull;
            const int fr = KH + RU - 1;
#pragma unroll 1
            for (int r = 0; r < fr; ++r) {
                const unsigned long long b = bits64(bits, (RU * g + r) * IC + fc0) & FWMASK;
                if (r == 0) bor0 = b;
                else if (r == fr - 1) bor1 = b;
                else mid |= b;
            }
            // window row 0 = footprint rows 0..KH-1, window row 1 = rows 1..KH
            bor0 |= mid;
            bor1 |= mid;
        }

#pragma unroll 1
        for (int u = 0; u < RU; ++u) {
            const int Y = Yg + u;
            if (u > 0 && any) {
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
            const unsigned ob = (okb >> (u * RT)) & ((1u << RT) - 1u);
            if (ob) {
                // ---- missing count and mask-kernel sums of the RT windows of this row
                int nm[RT];
                float sK[RT], sK2[RT];
#pragma unroll
                for (int t = 0; t < RT; ++t) {
                    nm[t] = 0;
                    sK[t] = sK2[t] = 0.f;
                }
                unsigned slowu = 0u;
                if (MODE == MODE_GEO) {
                    const uint32_t KHMASK = (KH >= 32) ? 0xffffffffu : ((1u << KH) - 1u);
                    const uint32_t Rs = (Rfp >> u) & KHMASK;
                    const int Yw = Y - kh;  // image row of window row 0
                    if (stripz) {                                            // [sec:strip]
#pragma unroll
                        for (int t = 0; t < RT; ++t) {
                            const unsigned sd = (unsigned)(d00 + t - u - P.st_base);
                            if (sd < (unsigned)P.st_n) {
                                const float4 s = ST[sd];
                                nm[t] = (int)s.x;
                                sK[t] = s.y;
                                sK2[t] = s.z;
                            }
                        }
                    }
                    // missing rows: the taps of kernel row i inside the row's flagged span   [sec:rects]
                    for (uint32_t rs = Rs; rs; rs &= rs - 1) {
                        const int i = __ffs(rs) - 1;
                        const int Yr = Yw + i;
                        const int xlo = max(Yr + P.mlo, P.mx0) - (X0 - kw);
                        const int xhi1 = min(Yr + P.mhi, P.mx1 - 1) + 1 - (X0 - kw);
                        const float2 *pr = PR + i * KW1;
#pragma unroll
                        for (int t = 0; t < RT; ++t) {
                            const int jlo = min(max(xlo - t, 0), KW);
                            const int jhi = min(max(xhi1 - t, 0), KW);
                            if (jhi > jlo) {
                                const float2 p1 = pr[jhi], p0 = pr[jlo];
                                nm[t] += jhi - jlo;
                                sK[t] += p1.x - p0.x;
                                sK2[t] += p1.y - p0.y;
                            }
                        }
                    }
                    // missing columns: the taps of kernel column j inside the column's flagged span,
                    // minus the pixels on missing rows (counted above)
                    for (unsigned long long cc = Cfp; cc; cc &= cc - 1) {
                        const int c = __ffsll((long long)cc) - 1;
                        const int Xc = X0 - kw + c;
                        const int ilo = min(max(max(Xc - P.mhi, P.my0) - Yw, 0), KH);
                        const int ihi = min(max(min(Xc - P.mlo, P.my1 - 1) + 1 - Yw, 0), KH);
                        if (ihi <= ilo) continue;
                        const uint32_t rsel =
                            Rs & (((ihi >= 32) ? 0xffffffffu : ((1u << ihi) - 1u)) & ~((1u << ilo) - 1u));
                        const int nc = (ihi - ilo) - __popc(rsel);
#pragma unroll
                        for (int t = 0; t < RT; ++t) {
                            const int j = c - t;
                            if (j >= 0 && j < KW) {
                                const float2 p1 = PC[j * (KH + 1) + ihi], p0 = PC[j * (KH + 1) + ilo];
                                float xk = p1.x - p0.x, xk2 = p1.y - p0.y;
                                for (uint32_t rr2 = rsel; rr2; rr2 &= rr2 - 1) {
                                    const float2 *q = PR + (__ffs(rr2) - 1) * KW1 + j;
                                    xk -= q[1].x - q[0].x;
                                    xk2 -= q[1].y - q[0].y;
                                }
                                nm[t] += nc;
                                sK[t] += xk;
                                sK2[t] += xk2;
                            }
                        }
                    }
                    // windows on a masked part of the frame: exact path (per-pixel predicate).
                    // All four margins when margin_mode = 2; with a banded frame only the top
                    // margin and the lower part of the right margin are masked (pre:461-477).
                    if (P.margin_mode != 0) {
                        const bool top = Yw < P.my0;
                        const bool bottom = P.margin_mode == 2 && Yw + KH > P.my1;
                        if (top || bottom) slowu = (1u << RT) - 1u;
                        const int tlo = P.mx0 - (X0 - kw);          // t < tlo: left margin
                        const int thi = P.mx1 - KW - (X0 - kw);     // t > thi: right margin
                        if (P.margin_mode == 2 && tlo > 0)
                            slowu |= (tlo >= RT) ? ((1u << RT) - 1u) : ((1u << tlo) - 1u);
                        if (thi < RT - 1 && (P.margin_mode == 2 || Yw + KH > P.right_y0))
                            slowu |= (thi < 0) ? ((1u << RT) - 1u) : (((1u << RT) - 1u) & ~((2u << thi) - 1u));
                    }
                }
                if (MODE == MODE_BITS) {
                    // windows with a missing pixel: exact path
                    const unsigned long long bu = u ? bor1 : bor0;
#pragma unroll
                    for (int t = 0; t < RT; ++t)
                        if ((unsigned)(bu >> t) & KWMASK) slowu |= 1u << t;
                }

                // ---- window sums by sliding along the row, scores            [sec:score]
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
                float rr[RT];
                int nn[RT];
                const float cm = MASK ? ((MODE == MODE_GEO ? P.fillc : 0.f) - pl) : 0.f;  // S' of a missing pixel
                const bool raw = P.raw_xcorr != 0;
                // per-block constants of the per-window formulas
                const float ncm2 = -cm * cm, pl2 = pl + pl;
                const float e0 = P.escale * fmaf(pl, pl, 1.f), e1 = P.escale * invN;
                const float thrN = P.thr * (float)P.N, Nf = (float)P.N;
                // windows with a missing pixel need min_present pixels and a kernel of non-zero mean
                const int minp = P.kmean_zero ? 0x7fffffff : max(P.min_present, 1);
                const int nmsel = (MASK && P.nobs_full) ? -1 : 0;
#pragma unroll
                for (int t = 0; t < RT; ++t) {
                    if (t > 0) {
                        g1 += col[off + t + KW - 1] - col[off + t - 1];
                        g2 += cqq[off + t + KW - 1] - cqq[off + t - 1];
                    }
                    // sums over the present pixels (centred): P1 = sum S', P2 = sum S'^2, Q3 = sum S' K'
                    const float nf = (float)nm[t];
                    const int npres = P.N - nm[t];
                    const float np = (float)npres;
                    float P1 = g1, P2 = g2, Q3 = s3a[t];
                    if (MASK) {
                        P1 = fmaf(-cm, nf, g1);
                        P2 = fmaf(ncm2, nf, g2);
                        Q3 = fmaf(-cm, sK[t], Q3);
                    }
                    // the three raw correlations xcorr2 thresholds at 1e-4 (det:716), over the
                    // whole window (missing pixels are zeros), times N
                    const float kp = MASK ? (P.sumKp - sK[t]) : P.sumKp;  // sum K' over present
                    const float U1 = fmaf(np, pl, P1);                    // sum S
                    const float U2 = fmaf(pl, P1 + U1, P2);               // sum S^2 = P2 + pl (2 P1 + n pl)
                    const float U3 = fmaf(P.qf, U1, fmaf(pl, kp, Q3));    // sum S K_corr
                    const float E = fmaf(e1, g2, e0);                     // their float32 error scale
                    const float zoneN = fmaf(P.thr, 1e-3f, E) * Nf;
                    const float a = MASK ? (P.ksump - sK[t]) : P.ksump;
                    const float b = MASK ? (P.k2sump - sK2[t]) : P.k2sump;
                    const float invn = MASK ? rcp_fast(np) : invN;
                    const float tt = P1 * invn;
                    const float C = fmaf(pl, P.delta, fmaf(-tt, a, Q3));   // n cov
                    const float VS = fmaf(-P1, tt, P2);                    // n var S
                    const float VK = fmaf(-a * invn, a, b);                // n var K
                    const float den2 = VS * VK;
                    bool ok = true;
                    if (MASK) ok = nm[t] == 0 || npres >= minp;
                    // det:1088-1091: |sqrt(var S var K)| < 1e-10 -> 0
                    const bool dok = den2 > P.den_floor * np * np;
                    float r = fminf(1.f, fmaxf(-1.f, C * rsqrt_fast(den2)));
                    // a mean of squares thresholded to 0 makes the variance <= 0: score 0
                    const bool zero2 = fabsf(U2) < thrN - zoneN;
                    bool hard = !zero2 && ok &&
                                (!dok ||                                       // flat window
                                 fminf(fminf(fabsf(U1), fabsf(U2)), fabsf(U3)) < thrN + zoneN ||  // a threshold within rounding distance
                                 g2 * P.sumKp2 > kAmpLimit2 * den2 ||          // ill-conditioned float32 sums
                                 fabsf(U2) * P.sumKp2 > kAmpLimitQ * den2);    // ... or tap products
                    if (zero2 || !ok || !dok) r = 0.f;
                    int nmo = (r != 0.f) ? (nm[t] & nmsel) : 0;
                    if (raw) {
                        // xcorr2: the thresholded raw correlation itself
                        const float au = fabsf(U3);
                        r = au < P.thr ? 0.f : U3;
                        hard = fabsf(au - P.thr) <= fmaf(P.thr, 1e-3f, E * Nf);
                        nmo = 0;
                    }
#ifdef CS_ABLATE
                    if (P.dbg & 32) hard = false;
#endif
                    if (hard) slowu |= 1u << t;
#ifdef CS_ABLATE
                    if (P.dbg & 2) r = Q3 + g1 + g2;
#endif
                    rr[t] = r;
                    nn[t] = nmo;
                }
                slow |= (slowu & ob) << (u * RT);

                // ---- scores (and missing counts) to the output band, 16 bytes at a time where
                // the row of eight windows is complete and aligned             [sec:store]
                const long long oi0 =
                    (long long)(Y - P.osy) * P.out_pitch + ((X0 - P.osx) - P.out_dlo);
                // the missing-count plane is zeroed by the caller: only rows with a count are written
                int anyn = 0;
#pragma unroll
                for (int t = 0; t < RT; ++t) anyn |= nn[t];
                const bool wn = P.nmiss != nullptr && anyn != 0;
                if (ob == ((1u << RT) - 1u) && (oi0 & 3) == 0) {
#pragma unroll
                    for (int t = 0; t < RT; t += 4)
                        *reinterpret_cast<float4 *>(P.out + oi0 + t) =
                            make_float4(rr[t], rr[t + 1], rr[t + 2], rr[t + 3]);
                    if (wn) {
                        if (P.nmiss16) {
#pragma unroll
                            for (int t = 0; t < RT; t += 4)
                                *reinterpret_cast<uint2 *>((unsigned short *)P.nmiss + oi0 + t) = make_uint2(
                                    (unsigned)nn[t] | ((unsigned)nn[t + 1] << 16),
                                    (unsigned)nn[t + 2] | ((unsigned)nn[t + 3] << 16));
                        } else {
#pragma unroll
                            for (int t = 0; t < RT; t += 4)
                                *reinterpret_cast<unsigned *>((unsigned char *)P.nmiss + oi0 + t) =
                                    (unsigned)nn[t] | ((unsigned)nn[t + 1] << 8) |
                                    ((unsigned)nn[t + 2] << 16) | ((unsigned)nn[t + 3] << 24);
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < RT; ++t)
                        if ((ob >> t) & 1u) {
                            P.out[oi0 + t] = rr[t];
                            if (wn) {
                                if (P.nmiss16) ((unsigned short *)P.nmiss)[oi0 + t] = (unsigned short)nn[t];
                                else ((unsigned char *)P.nmiss)[oi0 + t] = (unsigned char)nn[t];
                            }
                        }
                }
            }
            // the next row's products
#pragma unroll
            for (int t = 0; t < RT; ++t) s3a[t] = s3b[t];
        }

        // ---- exact path: the warp redoes the flagged windows in float64 from the tile,
        // 32 pixels at a time, with the per-pixel mask predicate                [sec:redo]
#ifdef CS_ABLATE
        if (P.cnt && slow) atomicAdd(&P.cnt[0], (unsigned long long)__popc(slow));
#endif
        for (unsigned todo = __ballot_sync(0xffffffffu, slow != 0u); todo; todo &= todo - 1) {
            const int src = __ffs(todo) - 1;
            unsigned sl = __shfl_sync(0xffffffffu, slow, src);
            const int sYg = __shfl_sync(0xffffffffu, Yg, src);
            const int sX0 = __shfl_sync(0xffffffffu, X0, src);
            const int sg = __shfl_sync(0xffffffffu, g, src);
            const int sfc0 = __shfl_sync(0xffffffffu, fc0, src);
            for (; sl; sl &= sl - 1) {
                const int k = __ffs(sl) - 1;
                const int u = k / RT, t = k - u * RT;
                const int Y = sYg + u, X = sX0 + t;
                const int trow = RU * sg + u, tcol = sfc0 + t;  // tile position of the window's corner
                double h1 = 0.0, h2 = 0.0, q3 = 0.0, sKm = 0.0, sKm2 = 0.0;
                int nmiss = 0;
#pragma unroll 2
                for (int idx = lane; idx < KH * KW; idx += 32) {
                    const int i = idx / KW, j = idx - i * KW;
                    bool miss = false;
                    if (MODE == MODE_GEO) {
                        const bool rbit = (rbw[(trow + i) >> 5] >> ((trow + i) & 31)) & 1u;
                        const bool cbit = (cbw[(tcol + j) >> 5] >> ((tcol + j) & 31)) & 1u;
                        miss = geo_missing(P, Y - kh + i, X - kw + j, rbit, cbit);
                    }
                    if (MODE == MODE_BITS) {
                        const int pos = (trow + i) * IC + tcol + j;
                        miss = (bits[pos >> 5] >> (pos & 31)) & 1u;
                    }
                    const double kc = __ldg(P.dK + idx);
                    const double sv = miss ? 0.0 : (double)tile[(trow + i) * IC + tcol + j];
                    h1 += sv;
                    h2 = fma(sv, sv, h2);
                    q3 = fma(sv, kc, q3);
                    if (MASK) {
                        const double km = __ldg(P.dK + P.N + idx), k2m = __ldg(P.dK + 2 * P.N + idx);
                        nmiss += miss ? 1 : 0;
                        sKm += miss ? km : 0.0;
                        sKm2 += miss ? k2m : 0.0;
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
                    h2 += __shfl_xor_sync(0xffffffffu, h2, o);
                    q3 += __shfl_xor_sync(0xffffffffu, q3, o);
                    if (MASK) {
                        sKm += __shfl_xor_sync(0xffffffffu, sKm, o);
                        sKm2 += __shfl_xor_sync(0xffffffffu, sKm2, o);
                        nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
                    }
                }
                if (lane == 0) {
                    int nmo;
                    const float r = exact_score<MASK>(P, h1, h2, q3, nmiss, sKm, sKm2, nmo);
                    const long long oi =
                        (long long)(Y - P.osy) * P.out_pitch + ((X - P.osx) - P.out_dlo);
                    P.out[oi] = r;
                    if (P.nmiss) {
                        if (P.nmiss16) ((unsigned short *)P.nmiss)[oi] = (unsigned short)nmo;
                        else ((unsigned char *)P.nmiss)[oi] = (unsigned char)nmo;
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
                        sKm += P.dK[P.N + idx];
                        sKm2 += P.dK[2 * P.N + idx];
                    }
                    continue;
                }
                const double sv = (double)v;
                h1 += sv;
                h2 = fma(sv, sv, h2);
                s3 = fma(sv, P.dK[idx], s3);
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
                int nmo;
                const float r = exact_score<MASK>(P, h1, h2, s3, nmiss, sKm, sKm2, nmo);
                const long long oi =
                    (long long)(Y - P.osy) * P.out_pitch + ((X - P.osx) - P.out_dlo);
                P.out[oi] = r;
                if (P.nmiss) {
                    if (P.nmiss16) ((unsigned short *)P.nmiss)[oi] = (unsigned short)nmo;
                    else ((unsigned char *)P.nmiss)[oi] = (unsigned char)nmo;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- exact scores from the CSR
// Scores that decide something (pixels at or near the Pearson threshold of pick_foci,
// det:417-421) are recomputed from the float64 CSR signal with the per-pixel mask predicate
// and the reference's float64 formulas, so that the candidate set -- and with it the foci --
// does not depend on float32 rounding.  One warp per listed pixel.
struct ExactArgs {
    // signal (matrix coordinates) and its place in the framed image
    const int64_t *indptr;
    const int32_t *indices;
    const double *data;
    int rows, cols, pr, pc;
    // pixel mask (mask_mode 1): CSR pattern, diagonals kept by the frame's trim
    const int64_t *m_indptr;
    const int32_t *m_indices;
    int trim_lo, trim_hi;
    int mask_mode;
    // score image (matrix coordinates)
    float *out;
    void *nmiss;
    int nmiss16, out_pitch, out_dlo, out_dense;
    double thr_pearson;
};

__device__ __forceinline__ bool csr_find(const int64_t *indptr, const int32_t *indices, int r, int c,
                                         int64_t &pos) {
    int64_t lo = indptr[r], hi = indptr[r + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (indices[mid] < c) lo = mid + 1;
        else hi = mid;
    }
    pos = lo;
    return lo < indptr[r + 1] && indices[lo] == c;
}

__global__ void collect_near(ExactArgs A, int rows, int cols, int dlo, int dhi, int dmin, int dmax,
                             float lo, float hi, int2 *list, long long cap,
                             unsigned long long *count) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int y = blockIdx.x * wpb + (threadIdx.x >> 5); y < rows; y += gridDim.x * wpb) {
        long long a = (long long)y + dmin, b = (long long)y + dmax;
        if (!A.out_dense) {
            a = a < (long long)y + dlo ? (long long)y + dlo : a;
            b = b > (long long)y + dhi ? (long long)y + dhi : b;
        }
        if (a < 0) a = 0;
        if (b > cols - 1) b = cols - 1;
        // eight independent loads per lane and pass (a row of <= 256 scores in one pass)
        const float *row = A.out + (long long)y * A.out_pitch - A.out_dlo;
        for (long long xb = a + lane; xb <= b; xb += 256) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (xb + 32 * k <= b) ? row[xb + 32 * k] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (v[k] != 0.f && v[k] >= lo && v[k] <= hi) {
                    const unsigned long long o = atomicAdd(count, 1ull);
                    if ((long long)o < cap) list[o] = make_int2(y, (int)(xb + 32 * k));
                }
        }
    }
}

template <int MASKMODE>
__global__ void exact_windows(const PearsonParams P, const ExactArgs A, const int2 *list, long long n) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int KH = P.KH, KW = P.KW, kh = (KH - 1) / 2, kw = (KW - 1) / 2;
    for (long long w = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); w < n;
         w += (long long)gridDim.x * wpb) {
        const int y = list[w].x, x = list[w].y;
        const int Y = y + A.pr, X = x + A.pc;  // image coordinates of the window centre
        double h1 = 0.0, h2 = 0.0, q3 = 0.0, sKm = 0.0, sKm2 = 0.0;
        int nmiss = 0;
        for (int idx = lane; idx < KH * KW; idx += 32) {
            const int i = idx / KW, j = idx - i * KW;
            const int Yp = Y - kh + i, Xp = X - kw + j;
            const int r = Yp - A.pr, c = Xp - A.pc;
            const bool inside = r >= 0 && r < A.rows && c >= 0 && c < A.cols;
            bool miss = false;
            if (MASKMODE == MODE_GEO) {
                bool rbit = false, cbit = false;
                if (inside) {
                    rbit = (P.rbits[Yp >> 5] >> (Yp & 31)) & 1u;
                    cbit = (P.cbits[Xp >> 5] >> (Xp & 31)) & 1u;
                }
                miss = geo_missing(P, Yp, Xp, rbit, cbit);
            }
            if (MASKMODE == MODE_BITS) {
                bool bit = false;
                if (inside && c - r >= A.trim_lo && c - r <= A.trim_hi) {
                    int64_t pos;
                    bit = csr_find(A.m_indptr, A.m_indices, r, c, pos);
                }
                // the frame (margins, strip) of frame_missing_mask around the pixel mask
                miss = geo_missing(P, Yp, Xp, bit, false);
            }
            if (miss) {
                ++nmiss;
                sKm += __ldg(P.dK + P.N + idx);
                sKm2 += __ldg(P.dK + 2 * P.N + idx);
                continue;
            }
            if (!inside) continue;
            int64_t pos;
            if (csr_find(A.indptr, A.indices, r, c, pos)) {
                const double sv = A.data[pos];
                h1 += sv;
                h2 = fma(sv, sv, h2);
                q3 = fma(sv, __ldg(P.dK + idx), q3);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            h1 += __shfl_xor_sync(0xffffffffu, h1, o);
            h2 += __shfl_xor_sync(0xffffffffu, h2, o);
            q3 += __shfl_xor_sync(0xffffffffu, q3, o);
            sKm += __shfl_xor_sync(0xffffffffu, sKm, o);
            sKm2 += __shfl_xor_sync(0xffffffffu, sKm2, o);
            nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
        }
        if (lane == 0) {
            // the double itself (not its float32 rounding) is compared with the threshold
            int nmo;
            const double rx = exact_score_core<MASKMODE != MODE_NOMASK>(P, h1, h2, q3, nmiss, sKm, sKm2, nmo);
            float r = (float)rx;
            // keep the side of the threshold the double value is on (pick_foci compares float64)
            if (rx >= A.thr_pearson && (double)r < A.thr_pearson) r = nextafterf(r, 2.f);
            if (rx < A.thr_pearson && (double)r >= A.thr_pearson) r = nextafterf(r, -2.f);
            const long long oi = (long long)y * A.out_pitch + (x - A.out_dlo);
            A.out[oi] = r;
            if (A.nmiss) {
                if (A.nmiss16) ((unsigned short *)A.nmiss)[oi] = (unsigned short)nmo;
                else ((unsigned char *)A.nmiss)[oi] = (unsigned char)nmo;
            }
        }
    }
}

// The same recomputation, one LANE per window row (kernels of at most 32 rows; 32 / KH windows
// per warp): the lane finds the first stored column of its row inside the window with one binary
// search and walks the few entries that follow, instead of one binary search per tap.  The
// mask sums need no memory at all (geometry).  Used for the geometric mask and without mask.
template <int MASKMODE>
__global__ void exact_windows_rows(const PearsonParams P, const ExactArgs A, const int2 *list, long long n,
                                   const unsigned long long *n_dev, long long cap) {
    constexpr bool MASK = MASKMODE != MODE_NOMASK;
    if (n_dev) {
        // the list length stays on the device (no host round trip between collecting and
        // redoing); an overflowing list is left to the host's fallback
        const unsigned long long nd = *n_dev;
        if (nd > (unsigned long long)cap) return;
        n = (long long)nd;
    }
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int KH = P.KH, KW = P.KW, kh = (KH - 1) / 2, kw = (KW - 1) / 2;
    const int wpw = 32 / KH;                    // windows per warp
    const int sub = lane / KH, i = lane - sub * KH;
    const long long nwarps = (long long)gridDim.x * wpb;
    for (long long w0 = (blockIdx.x * (long long)wpb + (threadIdx.x >> 5)) * wpw; w0 < n; w0 += nwarps * wpw) {
        const long long w = w0 + sub;
        const bool act = sub < wpw && w < n;
        int y = 0, x = 0;
        if (act) {
            const int2 p = list[w];
            y = p.x, x = p.y;
        }
        const int Y = y + A.pr, X = x + A.pc;   // image coordinates of the window centre
        const int Yp = Y - kh + i;              // image row of this lane
        const int r = Yp - A.pr;                // matrix row
        const int c0 = X - kw - A.pc;           // matrix column of tap j = 0
        const bool rin = act && r >= 0 && r < A.rows;
        double h1 = 0.0, h2 = 0.0, q3 = 0.0, sKm = 0.0, sKm2 = 0.0;
        int nmiss = 0;
        bool rbit = false;
        if (MASK && rin) rbit = (P.rbits[Yp >> 5] >> (Yp & 31)) & 1u;
        if (MASK && act) {
            // missing taps of this window row (pure geometry)
            for (int j = 0; j < KW; ++j) {
                const int Xp = X - kw + j, c = c0 + j;
                bool cbit = false;
                if (rin && c >= 0 && c < A.cols) cbit = (P.cbits[Xp >> 5] >> (Xp & 31)) & 1u;
                const bool inside = rin && c >= 0 && c < A.cols;
                if (geo_missing(P, Yp, Xp, inside && rbit, cbit)) {
                    ++nmiss;
                    sKm += __ldg(P.dK + P.N + i * KW + j);
                    sKm2 += __ldg(P.dK + 2 * P.N + i * KW + j);
                }
            }
        }
        if (rin) {
            const int ca = c0 < 0 ? 0 : c0;
            const int cb = (c0 + KW - 1 > A.cols - 1) ? A.cols - 1 : c0 + KW - 1;
            int64_t lo = A.indptr[r];
            const int64_t end = A.indptr[r + 1];
            int64_t hi = end;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (A.indices[mid] < ca) lo = mid + 1;
                else hi = mid;
            }
            for (int64_t pos = lo; pos < end; ++pos) {
                const int c = A.indices[pos];
                if (c > cb) break;
                const int j = c - c0;
                if (MASK) {
                    const int Xp = c + A.pc;
                    const bool cbit = (P.cbits[Xp >> 5] >> (Xp & 31)) & 1u;
                    if (geo_missing(P, Yp, Xp, rbit, cbit)) continue;
                }
                const double sv = A.data[pos];
                h1 += sv;
                h2 = fma(sv, sv, h2);
                q3 = fma(sv, __ldg(P.dK + i * KW + j), q3);
            }
        }
        // sum over the KH lanes of the window (segments never cross: lane + o stays inside)
        for (int o = 16; o > 0; o >>= 1) {
            const double a1 = __shfl_down_sync(0xffffffffu, h1, o), a2 = __shfl_down_sync(0xffffffffu, h2, o);
            const double a3 = __shfl_down_sync(0xffffffffu, q3, o);
            const bool take = i + o < KH;
            if (take) h1 += a1, h2 += a2, q3 += a3;
            if (MASK) {
                const double b1 = __shfl_down_sync(0xffffffffu, sKm, o), b2 = __shfl_down_sync(0xffffffffu, sKm2, o);
                const int b3 = __shfl_down_sync(0xffffffffu, nmiss, o);
                if (take) sKm += b1, sKm2 += b2, nmiss += b3;
            }
        }
        if (act && i == 0) {
            int nmo;
            const double rx = exact_score_core<MASK>(P, h1, h2, q3, nmiss, sKm, sKm2, nmo);
            float rr = (float)rx;
            // keep the side of the threshold the double value is on (pick_foci compares float64)
            if (rx >= A.thr_pearson && (double)rr < A.thr_pearson) rr = nextafterf(rr, 2.f);
            if (rx < A.thr_pearson && (double)rr >= A.thr_pearson) rr = nextafterf(rr, -2.f);
            const long long oi = (long long)y * A.out_pitch + (x - A.out_dlo);
            A.out[oi] = rr;
            if (A.nmiss) {
                if (A.nmiss16) ((unsigned short *)A.nmiss)[oi] = (unsigned short)nmo;
                else ((unsigned char *)A.nmiss)[oi] = (unsigned char)nmo;
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

template <int KW, int MODE>
static int launch_kw(const CUtensorMap &tmap, const PearsonParams &P, int grid, int threads,
                     size_t smem, cudaStream_t st) {
    auto kern = pearson_tiles<KW, MODE>;
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, st>>>(tmap, P);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

template <int MODE>
static int launch_mode(int KW, const CUtensorMap &tmap, const PearsonParams &P, int grid,
                       int threads, size_t smem, cudaStream_t st) {
    switch (KW) {
#define CS_CASE(n) \
    case n:        \
        return launch_kw<n, MODE>(tmap, P, grid, threads, smem, st);
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
    int get(size_t bytes, void **out) {
        const int slot = next;
        next = (next + 1) % 8;
        if (cap[slot] < bytes) {
            if (buf[slot]) cudaFree(buf[slot]);
            buf[slot] = nullptr;
            cap[slot] = 0;
            if (cudaMalloc(&buf[slot], bytes) != cudaSuccess) {
                set_error("cudaMalloc of the kernel tables failed");
                return CS_ERR_NOMEM;
            }
            cap[slot] = bytes;
        }
        *out = buf[slot];
        return CS_OK;
    }
};
static thread_local KtabRing g_ring;

static inline size_t align16(size_t v) { return (v + 15) / 16 * 16; }

}  // namespace cs

using namespace cs;

// constants shared by the tiled and the wide kernel; returns the mask mode
static int common_params(PearsonParams &P, const cs_kernel_desc *K, const cs_pearson_opts *opts) {
    const int nk = K->kh * K->kw;
    P.KH = K->kh, P.KW = K->kw, P.N = nk;
    P.osy = opts->out_row_shift, P.osx = opts->out_col_shift;
    P.ksum = K->k_sum, P.k2sum = K->k2_sum, P.kmean = K->k_mean, P.kstd = K->k_std;
    P.thr_d = opts->xcorr_threshold;
    P.thr = (float)opts->xcorr_threshold;
    P.invN_d = 1.0 / (double)nk;
    P.invN = (float)P.invN_d;
    P.vK0 = opts->mask_mode ? (K->k2_sum / (double)nk - K->k_mean * K->k_mean) : K->k_std * K->k_std;
    P.min_present = (int)((1.0 - opts->missing_tol) * (double)nk);
    P.kmean_zero = (K->k_mean == 0.0);
    P.raw_xcorr = opts->raw_xcorr;
    P.nobs_full = opts->nobs_full;
    P.den_floor = 1e-20f;
    return opts->mask_mode;
}

static int check_output(PearsonParams &P, const cs_layout *Lo, int oy0, int oy1, int ox0, int ox1) {
    P.out_pitch = Lo->pitch;
    P.out_dlo = Lo->dense ? 0 : Lo->dlo;
    bool ok = oy0 - P.osy >= 0 && ox0 - P.osx >= 0 && Lo->rows >= oy1 - P.osy &&
              Lo->cols >= ox1 - P.osx;
    const int sh = P.osx - P.osy;
    // output pixel (y, x) = (Y - osy, X - osx); its diagonal is d - (osx - osy)
    if (ok && !Lo->dense) ok = Lo->dlo <= P.odlo - sh && Lo->dhi >= P.odhi - sh;
    if (!ok) {
        set_error("output image does not cover scores on diagonals [%d,%d]", P.odlo - sh,
                  P.odhi - sh);
        return CS_ERR_INVALID;
    }
    return CS_OK;
}

// float64 copies of K_corr, K_mask, K2_mask behind `extra` bytes of other tables
static int upload_tables(const cs_kernel_desc *K, bool has_mask, const unsigned char *ftab,
                         size_t fbytes, cudaStream_t st, const unsigned char **d_ftab,
                         const double **d_dK) {
    const int nk = K->kh * K->kw;
    const size_t dbytes = (size_t)3 * nk * sizeof(double);
    const size_t total = align16(fbytes) + dbytes;
    std::vector<unsigned char> h(total, 0);
    if (fbytes) memcpy(h.data(), ftab, fbytes);
    double *hd = (double *)(h.data() + align16(fbytes));
    for (int i = 0; i < nk; ++i) {
        hd[i] = K->k_corr[i];
        hd[nk + i] = has_mask ? K->k_mask[i] : 0.0;
        hd[2 * nk + i] = has_mask ? K->k2_mask[i] : 0.0;
    }
    void *buf = nullptr;
    int rc = g_ring.get(total, &buf);
    if (rc) return rc;
    // pageable source: the copy is staged by the runtime before the call returns
    CS_CUDA(cudaMemcpyAsync(buf, h.data(), total, cudaMemcpyHostToDevice, st));
    *d_ftab = (const unsigned char *)buf;
    *d_dK = (const double *)((const unsigned char *)buf + align16(fbytes));
    return CS_OK;
}

// kernels wider than 31: the one-warp-per-window kernel
static int pearson_wide_launch(const cs_layout *Li, const float *d_img, const cs_kernel_desc *K,
                               const cs_pearson_opts *opts, int32_t oy0, int32_t oy1, int32_t ox0,
                               int32_t ox1, int32_t odlo, int32_t odhi, const cs_layout *Lo,
                               float *d_out, void *d_nmiss, cudaStream_t st) {
    PearsonParams P;
    memset(&P, 0, sizeof(P));
    P.oy0 = oy0, P.oy1 = oy1, P.ox0 = ox0, P.ox1 = ox1;
    const int dmin_poss = ox0 - (oy1 - 1), dmax_poss = (ox1 - 1) - oy0;
    P.odlo = odlo < dmin_poss ? dmin_poss : odlo;
    P.odhi = odhi > dmax_poss ? dmax_poss : odhi;
    CS_REQUIRE(P.odhi >= P.odlo, "empty output diagonal range");
    const int mode = common_params(P, K, opts);
    CS_REQUIRE(mode != 2, "kernels wider than 31 columns need the mask as NaN sentinels");
    P.out = d_out, P.nmiss = d_nmiss, P.nmiss16 = opts->nmiss_bytes == 2;
    int rc = check_output(P, Lo, oy0, oy1, ox0, ox1);
    if (rc) return rc;
    if (mode) CS_REQUIRE(K->k_mask && K->k2_mask, "mask kernels missing");
    const unsigned char *d_f = nullptr;
    if ((rc = upload_tables(K, mode != 0, nullptr, 0, st, &d_f, &P.dK))) return rc;
    WideParams Wp;
    Wp.img = d_img;
    Wp.pitch = Li->pitch;
    Wp.dense = Li->dense;
    Wp.dlo = Li->dense ? 0 : Li->dlo;
    Wp.dhi = Li->dense ? 0 : Li->dhi;
    const int nrows = oy1 - oy0;
    const int Wo = P.odhi - P.odlo + 1, ncols = ox1 - ox0;
    const int per_row = Wo < ncols ? Wo : ncols;
    dim3 grid((unsigned)(nrows < (1 << 20) ? nrows : (1 << 20)), (unsigned)((per_row + 7) / 8 > 64 ? 64 : (per_row + 7) / 8));
    if (mode)
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
                        void *d_nmiss, void *stream, bool plan_only, int32_t *tile_rows_out,
                        int32_t *band_gap_out = nullptr) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(Li && K && opts && (plan_only || (d_img && Lo && d_out)),
               "cs_pearson_f32: null argument");
    CS_REQUIRE(K->kh >= 1 && K->kw >= 3 && (K->kh & 1) && (K->kw & 1),
               "kernel shape must be odd (got %dx%d)", K->kh, K->kw);
    CS_REQUIRE(K->kh <= 255 && K->kw <= 255, "kernel %dx%d too large (255x255 at most)", K->kh,
               K->kw);
    CS_REQUIRE(oy1 > oy0 && ox1 > ox0, "empty output region");
    CS_REQUIRE(opts->mask_mode >= 0 && opts->mask_mode <= 2, "bad mask mode");
    const int kh = (K->kh - 1) / 2, kw = (K->kw - 1) / 2;
    CS_REQUIRE(oy0 - kh >= 0 && oy1 + kh <= Li->rows && ox0 - kw >= 0 && ox1 + kw <= Li->cols,
               "output region needs windows outside the image");
    if (K->kh > 31 || K->kw > 31) {
        if (plan_only) {
            if (tile_rows_out) *tile_rows_out = 32;
            if (band_gap_out) *band_gap_out = 0;
            return CS_OK;
        }
        return pearson_wide_launch(Li, d_img, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, Lo, d_out,
                                   d_nmiss, st);
    }
    PFN_encodeTiled enc = plan_only ? nullptr : get_encode();
    if (!enc && !plan_only) {
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
    const int mode = common_params(P, K, opts);
    P.KWP2 = round_up(K->kw + 1, 4);
    const int kwa = round_up(kw, 4);
    const int nrows_out = oy1 - oy0;
    const int nk = K->kh * K->kw;
    if (mode) {
        CS_REQUIRE(K->k_mask && K->k2_mask, "mask kernels missing");
        // the centred algebra of the float32 path takes one kernel for the signal and the mask
        for (int i = 0; i < nk; ++i)
            CS_REQUIRE(K->k_mask[i] == K->k_corr[i], "k_mask must equal k_corr");
    }

    // ---- tables -------------------------------------------------------------------
    // float: K' = K_corr - q as two padded rows per kernel row (see Ktab), [KH][2][KWP2];
    // mask modes: row / column prefix tables of (K', K2') pairs, strip table per diagonal
    double qd = 0.0;
    for (int i = 0; i < nk; ++i) qd += K->k_corr[i];
    const double sumKc = qd;
    qd /= nk;
    const float qf = (float)qd;
    const double q = (double)qf;
    const int n_ktab = 2 * K->kh * P.KWP2;
    P.sdlo = 0;
    P.sdhi = -1;
    P.st_base = 0;
    P.st_n = 0;
    if (mode == MODE_GEO) {
        P.sdlo = opts->geo.strip_dlo;
        P.sdhi = opts->geo.strip_dhi;
        if (P.sdhi >= P.sdlo) {
            P.st_base = P.sdlo - (kh + kw);
            P.st_n = (P.sdhi + (kh + kw)) - P.st_base + 1;
        } else {
            P.sdlo = 0, P.sdhi = -1;
        }
    }
    const int kw1 = K->kw + 1, kh1 = K->kh + 1;
    const size_t b_ktab = align16((size_t)n_ktab * sizeof(float));
    const size_t b_pr = mode == MODE_GEO ? align16((size_t)K->kh * kw1 * sizeof(float2)) : 0;
    const size_t b_pc = mode == MODE_GEO ? align16((size_t)K->kw * kh1 * sizeof(float2)) : 0;
    const size_t b_st = mode == MODE_GEO ? align16((size_t)P.st_n * sizeof(float4)) : 0;
    const size_t fbytes = b_ktab + b_pr + b_pc + b_st;
    P.tab_bytes = (int)fbytes;
    std::vector<unsigned char> hk(fbytes, 0);
    float *hf = (float *)hk.data();
    double sumKp = 0.0, sumKp2 = 0.0, sumAbsKp = 0.0;
    for (int i = 0; i < K->kh; ++i)
        for (int j = 0; j < K->kw; ++j) {
            const float v = (float)(K->k_corr[i * K->kw + j] - q);
            hf[(2 * i) * P.KWP2 + j] = v;          // (k0,k1),(k2,k3),...
            hf[(2 * i + 1) * P.KWP2 + j + 1] = v;  // (0,k0),(k1,k2),...
            sumKp += (double)v;
            sumKp2 += (double)v * (double)v;
            sumAbsKp += fabs((double)v);
        }
    if (mode == MODE_GEO) {
        // centred mask kernels, consistent with the float32 taps: K' as rounded above,
        // K2' = K2_mask - 2 q K_mask + q^2
        auto kp = [&](int i, int j) { return (double)hf[(2 * i) * P.KWP2 + j]; };
        auto k2p = [&](int i, int j) {
            const double km = K->k_mask[i * K->kw + j];
            return K->k2_mask[i * K->kw + j] - 2.0 * q * km + q * q;
        };
        float2 *pr = (float2 *)(hk.data() + b_ktab);
        float2 *pc = (float2 *)(hk.data() + b_ktab + b_pr);
        float4 *stt = (float4 *)(hk.data() + b_ktab + b_pr + b_pc);
        for (int i = 0; i < K->kh; ++i) {
            double a = 0.0, b = 0.0;
            pr[i * kw1] = make_float2(0.f, 0.f);
            for (int j = 0; j < K->kw; ++j) {
                a += kp(i, j);
                b += k2p(i, j);
                pr[i * kw1 + j + 1] = make_float2((float)a, (float)b);
            }
        }
        for (int j = 0; j < K->kw; ++j) {
            double a = 0.0, b = 0.0;
            pc[j * kh1] = make_float2(0.f, 0.f);
            for (int i = 0; i < K->kh; ++i) {
                a += kp(i, j);
                b += k2p(i, j);
                pc[j * kh1 + i + 1] = make_float2((float)a, (float)b);
            }
        }
        // window centred on diagonal d: tap (i, j) sits on diagonal d + (j - kw) - (i - kh)
        for (int sd = 0; sd < P.st_n; ++sd) {
            const int d = P.st_base + sd;
            double a = 0.0, b = 0.0, c = 0.0;
            for (int i = 0; i < K->kh; ++i)
                for (int j = 0; j < K->kw; ++j) {
                    const int dt = d + (j - kw) - (i - kh);
                    if (dt >= P.sdlo && dt <= P.sdhi) {
                        a += kp(i, j);
                        b += k2p(i, j);
                        c += 1.0;
                    }
                }
            stt[sd] = make_float4((float)c, (float)a, (float)b, 0.f);
        }
    }

    // ---- tiling ---------------------------------------------------------------
    const int Wo = odhi - odlo + 1;  // output diagonals
    const int ncols_out = ox1 - ox0;
    // banded traversal when the output band is narrow relative to the region
    P.skew = (Wo + 16 < ncols_out) ? 1 : 0;
    int TR = opts->tile_rows > 0 ? round_up(opts->tile_rows, RU) : 32;
#ifdef CS_ABLATE
    if (const char *e = getenv("CS_TILE_ROWS"))  // tuning knob for experiments
        if (atoi(e) > 0) TR = round_up(atoi(e), RU);
#endif
    if (TR > round_up(nrows_out, RU)) TR = round_up(nrows_out, RU);
    size_t smem = 0;
    int NBc = 0, nchunks = 0, IC = 0, IR = 0, NW = 0, threads = 0;
    for (;; TR -= RU) {
        if (TR < RU) {
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
        size_t o = align16((size_t)IC * IR * sizeof(float));
        P.off_bits = (int)o;
        if (mode == MODE_BITS) o += align16((size_t)NW * sizeof(uint32_t));
        P.off_K = (int)o;
        P.off_PR = (int)(o + b_ktab);
        P.off_PC = (int)(o + b_ktab + b_pr);
        P.off_ST = (int)(o + b_ktab + b_pr + b_pc);
        o += fbytes;
        P.off_rb = (int)o;
        if (mode == MODE_GEO) o += align16((size_t)((IR + 31) / 32 + 3) * sizeof(uint32_t));
        P.off_cb = (int)o;
        if (mode == MODE_GEO) o += align16((size_t)((IC + 31) / 32 + 4) * sizeof(uint32_t));
        P.off_bar = (int)o;
        o += 16;
        smem = o;
        if (smem <= 113 * 1024 || (TR == RU && smem <= 227 * 1024)) break;
    }
    // Width of the two triangles of a tile box that lie outside the stored band (a banded
    // traversal keeps them bounded): left, at the last box row of chunk 0; right, at the first
    // box row of the last chunk.  With at least that many zeros between the rows' bands
    // (cs_layout_band_padded) the box reads zeros there and the fix-up pass is skipped.
    int gap_need = 0;
    if (!Li->dense && P.skew) {
        const int Wimg = P.dhi - P.dlo + 1;
        const int left = IR - 1 - kh + kwa + 3 - (odlo - P.dlo);
        const int right = IC - Wimg + kh + (odlo - P.dlo) - kwa + RT * (nchunks - 1) * NBc;
        gap_need = left > right ? left : right;
        if (gap_need < 0) gap_need = 0;
    }
    P.fixup = Li->dense ? 0 : 1;
    if (!Li->dense && P.skew && Li->pitch - (P.dhi - P.dlo) >= gap_need) P.fixup = 0;
    if (plan_only) {
        if (tile_rows_out) *tile_rows_out = TR;
        if (band_gap_out) *band_gap_out = gap_need;
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
    P.nrb = nrb;
    const long long grid_ll = (long long)nrb * nchunks;
    if (grid_ll >= (1ll << 31)) {
        set_error("grid too large");
        return CS_ERR_INVALID;
    }

    // ---- output -----------------------------------------------------------------
    P.out = d_out;
    P.nmiss = d_nmiss;
    P.nmiss16 = opts->nmiss_bytes == 2;
    if (d_nmiss) CS_REQUIRE(opts->nmiss_bytes == 1 || opts->nmiss_bytes == 2, "nmiss_bytes must be 1 or 2");
    if (d_nmiss && opts->nmiss_bytes == 1)
        CS_REQUIRE(P.N - (P.min_present > 1 ? P.min_present : 1) <= 255,
                   "missing counts of this kernel need a 16-bit plane");
    int rc = check_output(P, Lo, oy0, oy1, ox0, ox1);
    if (rc) return rc;

    if ((rc = upload_tables(K, mode != 0, hk.data(), fbytes, st, &P.tab, &P.dK))) return rc;
    P.qf = qf;
    P.sumKp = (float)sumKp;
    P.sumKp2 = (float)sumKp2;
    P.sumKc_d = sumKc;
    P.ksump = (float)(K->k_sum - (double)nk * q);
    P.k2sump = (float)(K->k2_sum - 2.0 * q * K->k_sum + (double)nk * q * q);
    P.delta = (float)(sumKc - K->k_sum);
    P.fillc = opts->geo.fill_value;
    {
        // error scale of the three raw correlations in float32 (see the kernel): 4e-7 x the
        // kernel's magnitude
        const double kappa = sqrt(sumKp2 / nk) + fabs(q) + fabs(sumKp) / nk + sumAbsKp / nk;
        P.escale = (float)(4e-7 * (1.0 + kappa));
    }
    if (mode == MODE_GEO) {
        CS_REQUIRE(opts->geo.d_row_bits && opts->geo.d_col_bits, "geometric mask: bit vectors missing");
        P.rbits = (const uint32_t *)opts->geo.d_row_bits;
        P.cbits = (const uint32_t *)opts->geo.d_col_bits;
        P.mlo = opts->geo.mask_dlo;
        P.mhi = opts->geo.mask_dhi;
        P.my0 = opts->geo.mat_y0, P.my1 = opts->geo.mat_y1, P.mx0 = opts->geo.mat_x0, P.mx1 = opts->geo.mat_x1;
        P.margin_mode = opts->geo.margin_mode;
        P.top_x1 = opts->geo.top_x1;
        P.right_y0 = opts->geo.right_y0;
    }
#ifdef CS_ABLATE
    P.dbg = getenv("CS_DEBUG_SKIP") ? atoi(getenv("CS_DEBUG_SKIP")) : 0;
    P.cnt = nullptr;
    static unsigned long long *g_cnt = nullptr;
    if (getenv("CS_DEBUG_COUNT")) {
        if (!g_cnt) cudaMalloc(&g_cnt, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(g_cnt, 0, 16 * sizeof(unsigned long long), st);
        P.cnt = g_cnt;
    }
#endif

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
    int lrc;
    if (mode == MODE_GEO)
        lrc = launch_mode<MODE_GEO>(K->kw, tmap, P, (int)grid_ll, threads, smem, st);
    else if (mode == MODE_BITS)
        lrc = launch_mode<MODE_BITS>(K->kw, tmap, P, (int)grid_ll, threads, smem, st);
    else
        lrc = launch_mode<MODE_NOMASK>(K->kw, tmap, P, (int)grid_ll, threads, smem, st);
#ifdef CS_ABLATE
    if (P.cnt && lrc == CS_OK) {
        unsigned long long h[16];
        cudaMemcpyAsync(h, P.cnt, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "pearson stats: windows on the exact path %llu\n", h[0]);
    }
#endif
    return lrc;
}

// Pixels of the score image with a score >= threshold - 2e-5 on diagonals dmin..dmax (all
// candidates of pick_foci plus the borderline ones below the threshold; if they overflow the
// scratch list, only the borderline band of +-2e-5) are recomputed exactly from the CSR.
static int refine_setup(const RefineArgs &R, cudaStream_t st, PearsonParams &P, ExactArgs &A, int &mode) {
    const cs_kernel_desc *K = R.K;
    memset(&P, 0, sizeof(P));
    mode = common_params(P, K, R.opts);
    if (mode) CS_REQUIRE(K->k_mask && K->k2_mask, "mask kernels missing");
    P.mlo = -(1 << 29), P.mhi = 1 << 29;
    P.sdlo = 0, P.sdhi = -1;
    if (mode) {
        const cs_geo_mask &g = R.opts->geo;
        P.rbits = (const uint32_t *)g.d_row_bits;
        P.cbits = (const uint32_t *)g.d_col_bits;
        if (mode == MODE_GEO) P.mlo = g.mask_dlo, P.mhi = g.mask_dhi;
        P.my0 = g.mat_y0, P.my1 = g.mat_y1, P.mx0 = g.mat_x0, P.mx1 = g.mat_x1;
        P.margin_mode = g.margin_mode;
        P.top_x1 = g.top_x1, P.right_y0 = g.right_y0;
        if (g.strip_dhi >= g.strip_dlo) P.sdlo = g.strip_dlo, P.sdhi = g.strip_dhi;
    }
    const unsigned char *d_f = nullptr;
    int rc = upload_tables(K, mode != 0, nullptr, 0, st, &d_f, &P.dK);
    if (rc) return rc;
    memset(&A, 0, sizeof(A));
    A.indptr = R.d_indptr, A.indices = R.d_indices, A.data = R.d_data;
    A.rows = R.rows, A.cols = R.cols, A.pr = R.pr, A.pc = R.pc;
    A.m_indptr = R.d_m_indptr, A.m_indices = R.d_m_indices;
    A.trim_lo = R.trim_lo, A.trim_hi = R.trim_hi;
    A.mask_mode = mode;
    A.out = R.d_out;
    A.nmiss = R.d_nmiss;
    A.nmiss16 = R.opts->nmiss_bytes == 2;
    A.out_pitch = R.Lo->pitch;
    A.out_dlo = R.Lo->dense ? 0 : R.Lo->dlo;
    A.out_dense = R.Lo->dense;
    A.thr_pearson = R.threshold;
    return CS_OK;
}

// Enqueue-only variant for the geometric mask / no mask and kernels of at most 32 rows: every
// pixel >= threshold - 2e-5 is listed and redone with the list length kept on the device.
// Returns 1 when this form does not apply (pixel mask, tall kernel); the caller then reads
// *R.d_count with its next synchronisation: a count above R.cap means nothing was redone and
// exact_refine has to run.
int cs::exact_refine_enqueue(const RefineArgs &R, cudaStream_t st) {
    if (R.K->kh > 32 || R.opts->mask_mode == MODE_BITS) return 1;
    PearsonParams P;
    ExactArgs A;
    int mode;
    int rc = refine_setup(R, st, P, A, mode);
    if (rc) return rc;
    const float eps = 2e-5f;
    int grid = (R.Lo->rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    CS_CUDA(cudaMemsetAsync(R.d_count, 0, sizeof(unsigned long long), st));
    collect_near<<<grid, 256, 0, st>>>(A, R.Lo->rows, R.Lo->cols, R.Lo->dlo, R.Lo->dhi, R.dmin, R.dmax,
                                       (float)R.threshold - eps, 3.0e38f, R.d_list, R.cap, R.d_count);
    CS_LAUNCHED();
    const int blocks = 148 * 8;
    if (mode == MODE_GEO)
        exact_windows_rows<MODE_GEO><<<blocks, 256, 0, st>>>(P, A, R.d_list, 0, R.d_count, R.cap);
    else
        exact_windows_rows<MODE_NOMASK><<<blocks, 256, 0, st>>>(P, A, R.d_list, 0, R.d_count, R.cap);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int cs::exact_refine(const RefineArgs &R, cudaStream_t st, long long *n_refined) {
    *n_refined = 0;
    const cs_kernel_desc *K = R.K;
    PearsonParams P;
    ExactArgs A;
    int mode;
    int rc = refine_setup(R, st, P, A, mode);
    if (rc) return rc;
    const float eps = 2e-5f;
    int grid = (R.Lo->rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    unsigned long long n = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const float lo = (float)R.threshold - eps;
        const float hi = pass == 0 ? 3.0e38f : (float)R.threshold + eps;
        CS_CUDA(cudaMemsetAsync(R.d_count, 0, sizeof(unsigned long long), st));
        collect_near<<<grid, 256, 0, st>>>(A, R.Lo->rows, R.Lo->cols, R.Lo->dlo, R.Lo->dhi, R.dmin,
                                           R.dmax, lo, hi, R.d_list, R.cap, R.d_count);
        CS_LAUNCHED();
        CS_CUDA(cudaMemcpyAsync(&n, R.d_count, sizeof(n), cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
        if ((long long)n <= R.cap) break;
    }
    if ((long long)n > R.cap || n == 0) return CS_OK;  // nothing to do / too many to redo
    long long blocks = ((long long)n + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    static const bool by_tap = getenv("CS_EXACT_BY_TAP") != nullptr;  // the one-search-per-tap kernel
    if (K->kh <= 32 && mode != MODE_BITS && !by_tap) {
        const int wpw = 32 / K->kh;
        blocks = ((long long)n + 8 * wpw - 1) / (8 * wpw);
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (mode == MODE_GEO)
            exact_windows_rows<MODE_GEO><<<(int)blocks, 256, 0, st>>>(P, A, R.d_list, (long long)n, nullptr, 0);
        else
            exact_windows_rows<MODE_NOMASK><<<(int)blocks, 256, 0, st>>>(P, A, R.d_list, (long long)n, nullptr, 0);
    } else if (mode == MODE_GEO)
        exact_windows<MODE_GEO><<<(int)blocks, 256, 0, st>>>(P, A, R.d_list, (long long)n);
    else if (mode == MODE_BITS)
        exact_windows<MODE_BITS><<<(int)blocks, 256, 0, st>>>(P, A, R.d_list, (long long)n);
    else
        exact_windows<MODE_NOMASK><<<(int)blocks, 256, 0, st>>>(P, A, R.d_list, (long long)n);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    *n_refined = (long long)n;
    return CS_OK;
}

extern "C" int cs_pearson_f32(const cs_layout *Li, const float *d_img, const cs_kernel_desc *K,
                              const cs_pearson_opts *opts, int32_t oy0, int32_t oy1, int32_t ox0,
                              int32_t ox1, int32_t odlo, int32_t odhi, const cs_layout *Lo,
                              float *d_out, void *d_nmiss, void *stream) {
    return pearson_impl(Li, d_img, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, Lo, d_out, d_nmiss,
                        stream, false, nullptr);
}

extern "C" int cs_pearson_plan(const cs_layout *Li, const cs_kernel_desc *K, const cs_pearson_opts *opts,
                               int32_t oy0, int32_t oy1, int32_t ox0, int32_t ox1, int32_t odlo,
                               int32_t odhi, int32_t *tile_rows, int32_t *band_gap) {
    CS_REQUIRE(tile_rows && band_gap, "cs_pearson_plan: null argument");
    return pearson_impl(Li, nullptr, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, nullptr, nullptr,
                        nullptr, nullptr, true, tile_rows, band_gap);
}

extern "C" int cs_pearson_tile_rows(const cs_layout *Li, const cs_kernel_desc *K,
                                    const cs_pearson_opts *opts, int32_t oy0, int32_t oy1,
                                    int32_t ox0, int32_t ox1, int32_t odlo, int32_t odhi,
                                    int32_t *tile_rows) {
    CS_REQUIRE(tile_rows, "cs_pearson_tile_rows: null argument");
    return pearson_impl(Li, nullptr, K, opts, oy0, oy1, ox0, ox1, odlo, odhi, nullptr, nullptr,
                        nullptr, nullptr, true, tile_rows);
}
