// K1: fused sliding-window Pearson correlation of a small dense kernel against a
// banded / dense float32 image (detection.py:917-1131 of the reference).
//
// One CTA = one output tile of TR rows that follows the band (a parallelogram in
// matrix coordinates).  The input tile (TR+KH-1 rows x IC columns, matrix
// coordinates) is fetched with ONE TMA box load out of the skewed band (the row
// stride of the tensor map is pitch, see cs_layout), fixed up in shared memory
// (out-of-band aliases -> 0, NaN sentinels = missing), shifted by a tile pivot,
// then:
//   phase A  per-column sliding sums over KH rows in float64 (sum S', sum S'^2,
//            missing count) -> V in shared memory;
//   main     each thread owns a 4x4 block of windows: for every input row it
//            loads its row segment once (LDS.128) and feeds 4 x 4 x KW FFMAs;
//   epilogue horizontal KW-sums of V per window, the reference's formulas in
//            float64, one float32 score per window.
// No tensor cores: this is a CUDA-core stencil (BASELINE.json north_star).
#include "common.cuh"

namespace cs {

constexpr int kThreads = 256;

struct PearsonParams {
    // image
    int rows, cols, dlo, dhi, dense;
    // output pixel set (image coordinates)
    int oy0, oy1, ox0, ox1, odlo, odhi;
    // tiling
    int TR, G, NBc, nchunks, skew;
    int IC, IR, ICq;
    // kernel geometry
    int KH, KW, KWp, N;
    // output image
    float *out;
    unsigned short *nobs;
    int out_pitch, out_dlo, osy, osx;
    // kernel matrices [nmat][KH][KWp] float (device)
    const float *kmat;
    double q, sumKp, ksum, k2sum, kmean, kstd, thr;
    double sumKp2;   // sum K'^2
    double qm, qm2;  // pivots of the two mask kernels (their sums are accumulated centred)
    int min_present, kmean_zero, has_mask, raw_xcorr, nobs_full;
    // shared memory carve-up (bytes)
    int off_V, off_Vm, off_K, off_red, off_bar;
};

// ---------------------------------------------------------------- PTX helpers
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

__device__ __forceinline__ double thr0(double v, double t) { return fabs(v) < t ? 0.0 : v; }

// one kernel row (padded to a multiple of 4 floats, 16-byte aligned) -> registers;
// every lane reads the same address: a shared-memory broadcast
template <int KWQ>
__device__ __forceinline__ void load_krow(float *kk, const float *src) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
#pragma unroll
    for (int qd = 0; qd < KWQ; ++qd) {
        const float4 v = s4[qd];
        kk[4 * qd + 0] = v.x;
        kk[4 * qd + 1] = v.y;
        kk[4 * qd + 2] = v.z;
        kk[4 * qd + 3] = v.w;
    }
}

// Squared error amplification above which a window is recomputed in float64: the
// float32 sums carry ~1e-6 relative error on well-conditioned windows, and the score
// error grows like amp = f * rms(S') * rms(K') / (sigma_S * sigma_K).
constexpr double kAmpLimit2 = 9.0;

// The reference's formulas (det:1002-1020 no mask, det:1021-1092 masked) from the
// window sums of the shifted signal S' = S - p and shifted kernels.
//   h1, h2   : sum S', sum S'^2 over the N window pixels (missing pixels count as S = 0)
//   s3       : sum S' * K'            (K' = K_corr - q)
//   sKm,sKm2 : sums of the centred mask kernels over the missing pixels
template <bool MASK>
__device__ __forceinline__ float score_from_sums(const PearsonParams &P, double p, double h1,
                                                 double h2, int nmiss, double s3, double sKm_c,
                                                 double sKm2_c, int &nobs, double &amp2) {
    const double dN = (double)P.N;
    const double invN = 1.0 / dN;
    const double m1 = h1 * invN;
    double A1 = m1 + p;
    double A2 = h2 * invN + 2.0 * p * m1 + p * p;
    double A3 = (s3 + P.q * h1 + p * P.sumKp) * invN + p * P.q;
    nobs = P.N;
    amp2 = 0.0;
    if (P.raw_xcorr) return (float)thr0(A3 * dN, P.thr);
    A1 = thr0(A1, P.thr);
    A2 = thr0(A2, P.thr);
    A3 = thr0(A3, P.thr);
    double cov, den2, f = 1.0;
    bool ok = true;
    if (!MASK) {
        const double vS = A2 - A1 * A1;
        cov = A3 - A1 * P.kmean;
        den2 = vS * P.kstd * P.kstd;
        ok = vS >= 0.0;
    } else if (nmiss == 0) {
        const double vS = A2 - A1 * A1;
        const double k2mean = P.k2sum * invN;
        cov = A3 - A1 * P.kmean;
        den2 = vS * (k2mean - P.kmean * P.kmean);
    } else {
        const int npres = P.N - nmiss;
        f = dN / (double)npres;
        // mask kernels are stored centred (K - qm): add the pivot back
        const double sKm = thr0(sKm_c + P.qm * nmiss, P.thr);
        const double sKm2 = thr0(sKm2_c + P.qm2 * nmiss, P.thr);
        const double mK = (P.ksum - sKm) / (double)npres;
        const double m2K = (P.k2sum - sKm2) / (double)npres;
        const double mS = A1 * f;
        const double vS = A2 * f - mS * mS;
        cov = (A3 - A1 * mK) * f;
        den2 = vS * (m2K - mK * mK);
        ok = (npres > 0) && (npres >= P.min_present) && !P.kmean_zero;
        if (P.nobs_full && npres != 0) nobs = npres;
    }
    // det:1066,1088-1091: denom = sqrt(den2); |denom| < 1e-10 or NaN -> 0
    float r = 0.f;
    if (ok && den2 >= 1e-20 && den2 < 1e300) {
        amp2 = (h2 * P.sumKp2 * invN * invN) * f * f / den2;
        r = (float)cov * rsqrtf((float)den2);
        if (!(fabsf(r) <= 3.0e38f)) r = 0.f;
        r = fminf(1.f, fmaxf(-1.f, r));
    }
    return r;
}

// ---------------------------------------------------------------- the kernel
template <int KW, bool MASK>
__global__ void __launch_bounds__(kThreads, 2)
pearson_tiles(const __grid_constant__ CUtensorMap tmap, const PearsonParams P) {
    constexpr int kw = (KW - 1) / 2;
    constexpr int kwa = (kw + 3) / 4 * 4;
    constexpr int off = kwa - kw;
    constexpr int NQ = (2 * kwa + 4) / 4;
    constexpr int KWQ = (KW + 3) / 4;

    extern __shared__ __align__(1024) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    double2 *V = reinterpret_cast<double2 *>(smem + P.off_V);
    float *Vm = reinterpret_cast<float *>(smem + P.off_Vm);
    float *Ks = reinterpret_cast<float *>(smem + P.off_K);
    float *red = reinterpret_cast<float *>(smem + P.off_red);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + P.off_bar);

    const int tid = threadIdx.x;
    const int rb = blockIdx.x / P.nchunks;
    const int ch = blockIdx.x - rb * P.nchunks;
    const int kh = (P.KH - 1) / 2;
    const int Y0 = P.oy0 + rb * P.TR;
    // aligned X' (= X - dlo) of the first block of row group 0
    int xb;
    if (P.skew) {
        int v = Y0 + P.odlo - P.dlo;
        xb = (v >= 0 ? v / 4 : -((-v + 3) / 4)) * 4;
    } else {
        int v = P.ox0 - P.dlo;
        xb = (v >= 0 ? v / 4 : -((-v + 3) / 4)) * 4;
    }
    xb += 4 * ch * P.NBc;
    const int TXp = xb - kwa;  // X' of tile column 0
    const int TY = Y0 - kh;    // image row of tile row 0
    const int IC = P.IC, IR = P.IR;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        fence_proxy_async();
        mbar_expect_tx(bar, (uint32_t)(IC * IR * sizeof(float)));
        tma_load_2d(tile, &tmap, bar, TXp, TY);
    }
    // kernel matrices -> shared memory while the tile is in flight
    {
        const int nk = (MASK ? 3 : 1) * P.KH * P.KWp;
        for (int i = tid; i < nk; i += kThreads) Ks[i] = P.kmat[i];
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // ---- pass 1: band fix-up, tile statistics ---------------------------------
    float lsum = 0.f;
    int lnz = 0;
    for (int iy = tid / 64; iy < IR; iy += kThreads / 64) {
        const int Y = TY + iy;
        float *row = tile + iy * IC;
        for (int ix = tid % 64; ix < IC; ix += 64) {
            const int X = TXp + ix + P.dlo;
            const int d = X - Y;
            float v = row[ix];
            const bool inside = P.dense ? true : (d >= P.dlo && d <= P.dhi);
            if (!inside) {
                v = 0.f;
                row[ix] = 0.f;
            }
            if (v == v) {
                lsum += v;
                lnz |= (v != 0.f);
            } else if (!MASK) {
                row[ix] = 0.f;  // no-mask mode never sees sentinels; be safe
            }
        }
    }
    // block reduction of (sum, any non-zero)
    for (int o = 16; o > 0; o >>= 1) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        lnz |= __shfl_xor_sync(0xffffffffu, lnz, o);
    }
    if ((tid & 31) == 0) {
        red[tid >> 5] = lsum;
        red[8 + (tid >> 5)] = __int_as_float(lnz);
    }
    __syncthreads();
    float tsum = 0.f;
    int tnz = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        tsum += red[w];
        tnz |= __float_as_int(red[8 + w]);
    }
    const float pv = tsum / (float)(IC * IR);  // tile pivot
    const double p = (double)pv;

    // item decode ---------------------------------------------------------------
    const int nitems = P.G * P.NBc;

    if (!tnz) {
        // all-zero signal: every score of the tile is 0 (variance 0 -> det:1088-1091)
        for (int item = tid; item < nitems; item += kThreads) {
            const int g = item / P.NBc, m = item - g * P.NBc;
            const int Xp0 = xb + 4 * g * P.skew + 4 * m;
            for (int u = 0; u < 4; ++u) {
                const int Y = Y0 + 4 * g + u;
                if (Y >= P.oy1) continue;
                for (int t = 0; t < 4; ++t) {
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

    // ---- pass 2: shift by the pivot (missing pixels keep their NaN) -----------
    for (int i = tid; i < IC * IR; i += kThreads) {
        float v = tile[i];
        tile[i] = v - pv;  // NaN stays NaN
    }
    __syncthreads();

    // ---- phase A: vertical sliding sums in float64 ------------------------------
    const int Vpitch = 4 * P.ICq;  // entries per output row
    for (int ix = tid; ix < IC; ix += kThreads) {
        const int slot = (ix & 3) * P.ICq + (ix >> 2);
        double r1 = 0.0, r2 = 0.0;
        float rm = 0.f;
        for (int iy = 0; iy < IR; ++iy) {
            float v = tile[iy * IC + ix];
            if (MASK) {
                const bool miss = !(v == v);
                rm += miss ? 1.f : 0.f;
                v = miss ? -pv : v;
            }
            const double a = (double)v;
            r1 += a;
            r2 = fma(a, a, r2);
            const int yo = iy - (P.KH - 1);
            if (yo >= 0) {
                V[yo * Vpitch + slot] = make_double2(r1, r2);
                if (MASK) Vm[yo * Vpitch + slot] = rm;
                float w = tile[yo * IC + ix];
                if (MASK) {
                    const bool miss = !(w == w);
                    rm -= miss ? 1.f : 0.f;
                    w = miss ? -pv : w;
                }
                const double b = (double)w;
                r1 -= b;
                r2 = fma(-b, b, r2);
            }
        }
    }
    __syncthreads();

    // ---- main loop + epilogue ---------------------------------------------------
    const float *Kc = Ks;
    const float *Km = Ks + P.KH * P.KWp;
    const float *Km2 = Ks + 2 * P.KH * P.KWp;
    for (int item = tid; item < nitems; item += kThreads) {
        const int g = item / P.NBc, m = item - g * P.NBc;
        const int Xp0 = xb + 4 * g * P.skew + 4 * m;  // X' of output column t=0
        const int cxa = 4 * g * P.skew + 4 * m;       // aligned tile column of x[0]
        const int Yg = Y0 + 4 * g;
        // skip blocks without any valid output pixel
        {
            bool any = false;
            for (int u = 0; u < 4; ++u) {
                const int Y = Yg + u;
                if (Y >= P.oy1) continue;
                const int Xlo = max(P.ox0, Y + P.odlo), Xhi = min(P.ox1 - 1, Y + P.odhi);
                const int Xa = Xp0 + P.dlo;
                if (Xa + 3 >= Xlo && Xa <= Xhi) any = true;
            }
            if (!any) continue;
        }

        float acc[4][4], accm[4][4], accm2[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[u][t] = accm[u][t] = accm2[u][t] = 0.f;

        const int nrow = P.KH + 3;
        for (int iy = 0; iy < nrow; ++iy) {
            const float4 *rp = reinterpret_cast<const float4 *>(tile + (4 * g + iy) * IC + cxa);
            float x[4 * NQ];
            float mk[MASK ? 4 * NQ : 1];
#pragma unroll
            for (int qd = 0; qd < NQ; ++qd) {
                const float4 v = rp[qd];
                x[4 * qd + 0] = v.x;
                x[4 * qd + 1] = v.y;
                x[4 * qd + 2] = v.z;
                x[4 * qd + 3] = v.w;
            }
            if (MASK) {
#pragma unroll
                for (int e = 0; e < 4 * NQ; ++e) {
                    const bool miss = !(x[e] == x[e]);
                    mk[e] = miss ? 1.f : 0.f;
                    x[e] = miss ? -pv : x[e];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = iy - u;
                if (i < 0 || i >= P.KH) continue;
                float kk[4 * KWQ];
                load_krow<KWQ>(kk, Kc + i * P.KWp);
#pragma unroll
                for (int j = 0; j < KW; ++j) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[u][t] = fmaf(x[off + t + j], kk[j], acc[u][t]);
                }
                if (MASK) {
                    load_krow<KWQ>(kk, Km + i * P.KWp);
#pragma unroll
                    for (int j = 0; j < KW; ++j) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            accm[u][t] = fmaf(mk[off + t + j], kk[j], accm[u][t]);
                    }
                    load_krow<KWQ>(kk, Km2 + i * P.KWp);
#pragma unroll
                    for (int j = 0; j < KW; ++j) {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            accm2[u][t] = fmaf(mk[off + t + j], kk[j], accm2[u][t]);
                    }
                }
            }
        }

        // epilogue: one output row at a time
        const int qbase = cxa >> 2;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int Y = Yg + u;
            if (Y >= P.oy1) continue;
            const double2 *Vr = V + (4 * g + u) * Vpitch;
            const float *Vmr = Vm + (4 * g + u) * Vpitch;
            double h1 = 0.0, h2 = 0.0;
            float hm = 0.f;
#pragma unroll
            for (int e = off; e < off + KW; ++e) {
                const int s = (e & 3) * P.ICq + qbase + (e >> 2);
                const double2 v = Vr[s];
                h1 += v.x;
                h2 += v.y;
                if (MASK) hm += Vmr[s];
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (t > 0) {
                    const int e0 = off + t - 1, e1 = off + t - 1 + KW;
                    const int s0 = (e0 & 3) * P.ICq + qbase + (e0 >> 2);
                    const int s1 = (e1 & 3) * P.ICq + qbase + (e1 >> 2);
                    const double2 a = Vr[s0], b = Vr[s1];
                    h1 += b.x - a.x;
                    h2 += b.y - a.y;
                    if (MASK) hm += Vmr[s1] - Vmr[s0];
                }
                const int X = Xp0 + t + P.dlo;
                const int d = X - Y;
                if (X < P.ox0 || X >= P.ox1 || d < P.odlo || d > P.odhi) continue;
                int nmiss = 0;
                if (MASK) nmiss = (int)(hm + 0.5f);
                int nobs = P.N;
                double amp2 = 0.0;
                float r = score_from_sums<MASK>(P, p, h1, h2, nmiss, (double)acc[u][t],
                                                (double)accm[u][t], (double)accm2[u][t], nobs, amp2);
                if (amp2 > kAmpLimit2) {
                    // ill-conditioned window (flat signal or mostly missing): the float32
                    // accumulators are not accurate enough, redo the three sums in float64
                    double s3 = 0.0, sm = 0.0, sm2 = 0.0;
                    const float *wp = tile + (4 * g + u) * IC + cxa + off + t;
                    for (int i = 0; i < P.KH; ++i) {
                        const float *wr = wp + i * IC;
                        const float *kr = Kc + i * P.KWp;
                        for (int j = 0; j < KW; ++j) {
                            float sv = wr[j];
                            if (MASK && !(sv == sv)) {
                                sv = -pv;
                                sm += (double)Km[i * P.KWp + j];
                                sm2 += (double)Km2[i * P.KWp + j];
                            }
                            s3 = fma((double)sv, (double)kr[j], s3);
                        }
                    }
                    r = score_from_sums<MASK>(P, p, h1, h2, nmiss, s3, sm, sm2, nobs, amp2);
                }
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
static int launch_kw(const CUtensorMap &tmap, const PearsonParams &P, int grid, size_t smem,
                     cudaStream_t st) {
    auto kern = pearson_tiles<KW, MASK>;
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(tmap, P);
    CS_LAUNCHED();
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

template <bool MASK>
static int launch_mask(int KW, const CUtensorMap &tmap, const PearsonParams &P, int grid,
                       size_t smem, cudaStream_t st) {
    switch (KW) {
#define CS_CASE(n) \
    case n:        \
        return launch_kw<n, MASK>(tmap, P, grid, smem, st);
        CS_CASE(3) CS_CASE(5) CS_CASE(7) CS_CASE(9) CS_CASE(11) CS_CASE(13) CS_CASE(15) CS_CASE(17)
        CS_CASE(19) CS_CASE(21) CS_CASE(23) CS_CASE(25) CS_CASE(27) CS_CASE(29) CS_CASE(31)
#undef CS_CASE
        default:
            set_error("kernel width %d not supported (odd widths 3..31)", KW);
            return CS_ERR_INVALID;
    }
}

// device scratch holding the float kernel matrices of one launch; kept per stream-agnostic
// small ring so that back-to-back launches do not race on it.
struct KmatRing {
    float *buf[8] = {nullptr};
    size_t cap[8] = {0};
    int next = 0;
};
static thread_local KmatRing g_ring;

}  // namespace cs

using namespace cs;

extern "C" int cs_pearson_f32(const cs_layout *Li, const float *d_img, const cs_kernel_desc *K,
                              const cs_pearson_opts *opts, int32_t oy0, int32_t oy1, int32_t ox0,
                              int32_t ox1, int32_t odlo, int32_t odhi, const cs_layout *Lo,
                              float *d_out, uint16_t *d_nobs, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CS_REQUIRE(Li && d_img && K && opts && Lo && d_out, "cs_pearson_f32: null argument");
    CS_REQUIRE(K->kh >= 1 && K->kw >= 3 && (K->kh & 1) && (K->kw & 1),
               "kernel shape must be odd (got %dx%d)", K->kh, K->kw);
    CS_REQUIRE(K->kh * K->kw < 65535, "kernel too large");
    CS_REQUIRE(oy1 > oy0 && ox1 > ox0, "empty output region");
    const int kh = (K->kh - 1) / 2, kw = (K->kw - 1) / 2;
    CS_REQUIRE(oy0 - kh >= 0 && oy1 + kh <= Li->rows && ox0 - kw >= 0 && ox1 + kw <= Li->cols,
               "output region needs windows outside the image");
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
    if (!Li->dense) {
        // every pixel read by a window must be stored or be a true zero: windows of
        // outputs on diagonal d read diagonals d-(kw+kh) .. d+(kw+kh); pixels outside
        // the stored band count as zeros, which is what the caller asserts.
    }
    P.KH = K->kh;
    P.KW = K->kw;
    P.KWp = round_up(K->kw, 4);
    P.N = K->kh * K->kw;
    const int kwa = round_up(kw, 4);

    // ---- tiling ---------------------------------------------------------------
    const int Wo = odhi - odlo + 1;  // output diagonals
    const int ncols_out = ox1 - ox0;
    // banded traversal when the output band is narrow relative to the region
    P.skew = (Wo + 8 < ncols_out) ? 1 : 0;
    int TR = opts->tile_rows > 0 ? round_up(opts->tile_rows, 4) : 16;
    const int nmat = opts->has_mask ? 3 : 1;
    size_t smem = 0;
    int NBc = 0, nchunks = 0, IC = 0, IR = 0, ICq = 0;
    for (;; TR -= 4) {
        CS_REQUIRE(TR >= 4, "kernel %dx%d does not fit in shared memory", K->kh, K->kw);
        const int G = TR / 4;
        const int span = P.skew ? (Wo + 6) : (ncols_out + 3);
        const int nblk_total = (span + 3) / 4;
        const int nb_max = (256 - 2 * kwa - 4 * (G - 1) * P.skew) / 4;
        if (nb_max < 1) continue;
        // aim at ~kThreads items per tile
        int nb_want = kThreads / G;
        if (nb_want > nb_max) nb_want = nb_max;
        nchunks = (nblk_total + nb_want - 1) / nb_want;
        NBc = (nblk_total + nchunks - 1) / nchunks;
        IC = 4 * NBc + 4 * (G - 1) * P.skew + 2 * kwa;
        IR = TR + K->kh - 1;
        if (IR > 256) continue;
        ICq = (IC / 4) | 1;
        size_t o = (size_t)IC * IR * sizeof(float);
        o = (o + 15) / 16 * 16;
        P.off_V = (int)o;
        o += (size_t)TR * 4 * ICq * sizeof(double2);
        P.off_Vm = (int)o;
        if (opts->has_mask) o += (size_t)TR * 4 * ICq * sizeof(float);
        o = (o + 15) / 16 * 16;
        P.off_K = (int)o;
        o += (size_t)nmat * K->kh * P.KWp * sizeof(float);
        o = (o + 15) / 16 * 16;
        P.off_red = (int)o;
        o += 64 * sizeof(float);
        P.off_bar = (int)o;
        o += 16;
        smem = o;
        if (smem <= 112 * 1024 || (TR == 4 && smem <= 227 * 1024)) break;
    }
    P.TR = TR;
    P.G = TR / 4;
    P.NBc = NBc;
    P.nchunks = nchunks;
    P.IC = IC;
    P.IR = IR;
    P.ICq = ICq;
    const int nrb = (oy1 - oy0 + TR - 1) / TR;
    const long long grid_ll = (long long)nrb * nchunks;
    CS_REQUIRE(grid_ll < (1ll << 31), "grid too large");

    // ---- output -----------------------------------------------------------------
    P.osy = opts->out_row_shift;
    P.osx = opts->out_col_shift;
    CS_REQUIRE(oy0 - P.osy >= 0 && ox0 - P.osx >= 0 && Lo->rows >= oy1 - P.osy &&
                   Lo->cols >= ox1 - P.osx,
               "output image too small");
    P.out = d_out;
    P.nobs = d_nobs;
    P.out_pitch = Lo->pitch;
    P.out_dlo = Lo->dense ? 0 : Lo->dlo;
    if (!Lo->dense) {
        // output pixel (y, x) = (Y - osy, X - osx); its diagonal is d - (osx - osy)
        const int sh = P.osx - P.osy;
        CS_REQUIRE(Lo->dlo <= odlo - sh && Lo->dhi >= odhi - sh,
                   "output band [%d,%d] does not cover scores on diagonals [%d,%d]", Lo->dlo,
                   Lo->dhi, odlo - sh, odhi - sh);
    }

    // ---- kernel matrices ----------------------------------------------------------
    const int nk = K->kh * K->kw;
    double qd = 0.0;
    for (int i = 0; i < nk; ++i) qd += K->k_corr[i];
    qd /= nk;
    const float qf = (float)qd;
    float qmf = 0.f, qm2f = 0.f;
    if (opts->has_mask) {
        CS_REQUIRE(K->k_mask && K->k2_mask, "mask kernels missing");
        double a = 0.0, b = 0.0;
        for (int i = 0; i < nk; ++i) {
            a += K->k_mask[i];
            b += K->k2_mask[i];
        }
        qmf = (float)(a / nk);
        qm2f = (float)(b / nk);
    }
    const size_t kbytes = (size_t)nmat * K->kh * P.KWp * sizeof(float);
    float *hk = (float *)malloc(kbytes);
    if (!hk) return CS_ERR_NOMEM;
    memset(hk, 0, kbytes);
    double sumKp = 0.0, sumKp2 = 0.0;
    for (int i = 0; i < K->kh; ++i)
        for (int j = 0; j < K->kw; ++j) {
            const float v = (float)(K->k_corr[i * K->kw + j] - (double)qf);
            hk[i * P.KWp + j] = v;
            sumKp += (double)v;
            sumKp2 += (double)v * (double)v;
            if (opts->has_mask) {
                hk[(K->kh + i) * P.KWp + j] = (float)(K->k_mask[i * K->kw + j] - (double)qmf);
                hk[(2 * K->kh + i) * P.KWp + j] = (float)(K->k2_mask[i * K->kw + j] - (double)qm2f);
            }
        }
    KmatRing &ring = g_ring;
    const int slot = ring.next;
    ring.next = (ring.next + 1) % 8;
    if (ring.cap[slot] < kbytes) {
        if (ring.buf[slot]) cudaFree(ring.buf[slot]);
        ring.buf[slot] = nullptr;
        ring.cap[slot] = 0;
        CS_CUDA(cudaMalloc(&ring.buf[slot], kbytes));
        ring.cap[slot] = kbytes;
    }
    // pageable source: the copy is staged by the runtime before the call returns
    cudaError_t ce = cudaMemcpyAsync(ring.buf[slot], hk, kbytes, cudaMemcpyHostToDevice, st);
    free(hk);
    CS_CUDA(ce);
    P.kmat = ring.buf[slot];
    P.q = (double)qf;
    P.sumKp = sumKp;
    P.sumKp2 = sumKp2;
    P.qm = (double)qmf;
    P.qm2 = (double)qm2f;
    P.ksum = K->k_sum;
    P.k2sum = K->k2_sum;
    P.kmean = K->k_mean;
    P.kstd = K->k_std;
    P.thr = opts->xcorr_threshold;
    P.min_present = (int)((1.0 - opts->missing_tol) * (double)P.N);
    P.kmean_zero = (K->k_mean == 0.0);
    P.has_mask = opts->has_mask;
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
    if (opts->has_mask)
        return launch_mask<true>(K->kw, tmap, P, (int)grid_ll, smem, st);
    return launch_mask<false>(K->kw, tmap, P, (int)grid_ll, smem, st);
}
