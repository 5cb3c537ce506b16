// CUDA-core (SIMT) tap-GEMM kernels: the fp32 parity path and the path for channel counts the
// tcgen05 kernels do not take (Cin = 1 stem, num_class = 3 head, cv11 = 2..8 cardinal branches).
// Same descriptors and the same epilogue as the tcgen05 path (tbi_common.cuh).
#include "tbi_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// --------------------------------------------------------------------------------------------
// forward / data-gradient tap-GEMM
// --------------------------------------------------------------------------------------------
template <typename T, int TBN>
__global__ void __launch_bounds__(NT) tapgemm_simt_kernel(const __grid_constant__ tbi_tapgemm d) {
    // TBN = output channels per CTA tile (64 or 16); threads: 16 (co) x 16 (pixel) -> 4x(TBN/16) micro tile
    constexpr int CPT = TBN / 16;                       // couts per thread
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][TBN + 4];

    const int tid = threadIdx.x;
    const int g = blockIdx.z;
    const long long M = (long long)d.n * d.gh * d.gw;
    const long long p0 = (long long)blockIdx.x * BM;
    const int co_t0 = blockIdx.y * TBN;                 // within group
    const int Kg = d.ntaps * d.cin_g;

    // A-load role: pixel a_p = tid/4, channels (tid%4)*4 .. +3 of the 16-chunk
    const int a_p = tid >> 2, a_c = (tid & 3) * 4;
    long long ap = p0 + a_p;
    const bool a_valid = ap < M;
    int an = 0, agy = 0, agx = 0;
    if (a_valid) { agx = (int)(ap % d.gw); long long t = ap / d.gw; agy = (int)(t % d.gh); an = (int)(t / d.gh); }
    // B-load role: co = tid / (BK/4) ...: TBN*BK elements, NT threads
    const int tx = tid & 15, ty = tid >> 4;

    float acc[4][CPT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

    const T* s0 = (const T*)d.src[0].ptr;
    const T* s1 = (const T*)d.src[1].ptr;
    const T* wp = (const T*)d.w;
    const int c0 = d.src[0].c;
    const int cin_base = g * d.cin_g;                   // groups>1 only with a single source

    for (int tap = 0; tap < d.ntaps; ++tap) {
        const int iy = agy * d.in_stride + d.dy[tap];
        const int ix = agx * d.in_stride + d.dx[tap];
        const bool pix_ok = a_valid && iy >= 0 && iy < d.src[0].h && ix >= 0 && ix < d.src[0].w;
        for (int ci0 = 0; ci0 < d.cin_g; ci0 += BK) {
            // ---- A tile
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = ci0 + a_c + j;
                float v = 0.f;
                if (pix_ok && ci < d.cin_g) {
                    const int cg = cin_base + ci;
                    if (cg < c0) v = ldf(s0 + view_off(d.src[0], an, iy, ix, cg));
                    else         v = ldf(s1 + view_off(d.src[1], an, iy, ix, cg - c0));
                }
                As[a_c + j][a_p] = v;
            }
            // ---- B tile: Bs[k][co] = w[(g*cout_g + co_t0 + co)][tap*cin_g + ci0 + k]
            for (int e = tid; e < TBN * BK; e += NT) {
                const int co = e / BK, k = e % BK;
                const int cog = co_t0 + co, ci = ci0 + k;
                float v = 0.f;
                if (cog < d.cout_g && ci < d.cin_g)
                    v = ldf(wp + (size_t)(g * d.cout_g + cog) * Kg + (size_t)tap * d.cin_g + ci);
                Bs[k][co] = v;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[4], b[CPT];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < CPT; ++j) b[j] = Bs[k][tx * CPT + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long p = p0 + ty * 4 + i;
        if (p >= M) continue;
        const int gx = (int)(p % d.gw); long long t = p / d.gw; const int gy = (int)(t % d.gh); const int n = (int)(t / d.gh);
        const int oy = gy * d.epi.out_stride + d.epi.out_off_y, ox = gx * d.epi.out_stride + d.epi.out_off_x;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int cog = co_t0 + tx * CPT + j;
            if (cog < d.cout_g) epilogue_store<T>(d.epi, n, oy, ox, g * d.cout_g + cog, acc[i][j]);
        }
    }
}

// --------------------------------------------------------------------------------------------
// weight-gradient tap-GEMM: tile 64 ci x 64 co for one tap, K = a chunk of pixels (split-K, atomics)
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) tapwgrad_simt_kernel(const __grid_constant__ tbi_tapwgrad d, int pix_per_cta,
                                                          int ci_tiles, int co_tiles) {
    __shared__ float As[BK][BM + 4];     // [pixel][ci]
    __shared__ float Bs[BK][BN + 4];     // [pixel][co]
    __shared__ float bsum[BN];

    const int tid = threadIdx.x;
    int by = blockIdx.y;
    const int co_t = by % co_tiles; by /= co_tiles;
    const int ci_t = by % ci_tiles; by /= ci_tiles;
    const int tap = by;
    const int g = blockIdx.z;
    const long long M = (long long)d.n * d.gh * d.gw;
    const long long pbeg = (long long)blockIdx.x * pix_per_cta;
    const long long pend = min(M, pbeg + (long long)pix_per_cta);
    const int ci_t0 = ci_t * BM, co_t0 = co_t * BN;

    const int tx = tid & 15, ty = tid >> 4;             // ty -> ci micro row, tx -> co micro col
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* a0 = (const T*)d.a_src[0].ptr;
    const T* a1 = (const T*)d.a_src[1].ptr;
    const T* bp = (const T*)d.b_src.ptr;
    const int c0 = d.a_src[0].c;
    const bool do_bias = d.dbias != nullptr && tap == 0 && ci_t == 0;
    float bias_acc = 0.f;                               // thread tid<64 sums column tid when do_bias

    // load role: pixel row lp = tid/16 (0..15), channels lc = (tid%16)*4..+3
    const int lp = tid >> 4, lc = (tid & 15) * 4;

    for (long long pc = pbeg; pc < pend; pc += BK) {
        const long long p = pc + lp;
        bool ok = p < pend;
        int n = 0, gy = 0, gx = 0;
        if (ok) { gx = (int)(p % d.gw); long long t = p / d.gw; gy = (int)(t % d.gh); n = (int)(t / d.gh); }
        const int ay = gy * d.a_stride + d.a_dy[tap], ax = gx * d.a_stride + d.a_dx[tap];
        const int byy = gy * d.b_stride + d.b_dy[tap], bxx = gx * d.b_stride + d.b_dx[tap];
        const bool a_ok = ok && ay >= 0 && ay < d.a_src[0].h && ax >= 0 && ax < d.a_src[0].w;
        const bool b_ok = ok && byy >= 0 && byy < d.b_src.h && bxx >= 0 && bxx < d.b_src.w;
        // a tap contributes only where BOTH operands are in range; zero either side otherwise
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci_t0 + lc + j;
            float v = 0.f;
            if (a_ok && b_ok && ci < d.cin_g) {
                const int cg = g * d.cin_g + ci;
                if (cg < c0) v = ldf(a0 + view_off(d.a_src[0], n, ay, ax, cg));
                else         v = ldf(a1 + view_off(d.a_src[1], n, ay, ax, cg - c0));
            }
            As[lp][lc + j] = v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co_t0 + lc + j;
            float v = 0.f;
            if (b_ok && co < d.cout_g) v = ldf(bp + view_off(d.b_src, n, byy, bxx, g * d.cout_g + co));
            Bs[lp][lc + j] = v;
        }
        __syncthreads();
        if (do_bias && tid < BN) {
#pragma unroll
            for (int k = 0; k < BK; ++k) bias_acc += Bs[k][tid];
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    (void)bsum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci_t0 + ty * 4 + i;
        if (ci >= d.cin_g) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co_t0 + tx * 4 + j;
            if (co >= d.cout_g) continue;
            const int cog = g * d.cout_g + co;
            atomicAdd(d.dw + (size_t)tap * d.tap_stride + (size_t)ci * d.ci_stride + (size_t)cog * d.co_stride, acc[i][j]);
        }
    }
    if (do_bias && tid < BN) {
        const int co = co_t0 + tid;
        if (co < d.cout_g) atomicAdd(d.dbias + g * d.cout_g + co, bias_acc);
    }
}

}  // namespace

int tbi_tapgemm_simt(const tbi_tapgemm* dd, cudaStream_t s) {
    if (dd->nphase > 1) {                                  // one launch per parity phase on this path
        for (int p = 0; p < dd->nphase; ++p) {
            tbi_tapgemm q = *dd;
            q.nphase = 0;
            for (int t = 0; t < dd->ntaps; ++t) { q.dy[t] = dd->ph_dy[p][t]; q.dx[t] = dd->ph_dx[p][t]; }
            q.epi.out_stride = 2; q.epi.out_off_y = dd->ph_off_y[p]; q.epi.out_off_x = dd->ph_off_x[p];
            q.w = (const char*)dd->w + (size_t)p * dd->cout_g * dd->groups * dd->ntaps * dd->cin_g * tbi_dtype_size(dd->dtype);
            int rc = tbi_tapgemm_simt(&q, s);
            if (rc) return rc;
        }
        return TBI_OK;
    }
    const tbi_tapgemm* d = dd;
    TBI_CHECK(d->ntaps >= 1 && d->ntaps <= TBI_MAX_TAPS, TBI_ERR_BAD_SHAPE, "tapgemm: ntaps=%d", d->ntaps);
    TBI_CHECK(d->groups >= 1 && (d->groups == 1 || d->src[1].ptr == nullptr), TBI_ERR_UNSUPPORTED,
              "tapgemm: groups>1 needs a single source");
    const int ctot = d->src[0].c + (d->src[1].ptr ? d->src[1].c : 0);
    TBI_CHECK(ctot == d->cin_g * d->groups, TBI_ERR_BAD_SHAPE, "tapgemm: source channels %d != groups*cin_g %d",
              ctot, d->cin_g * d->groups);
    const long long M = (long long)d->n * d->gh * d->gw;
    TBI_CHECK(M > 0 && d->cout_g > 0 && d->cin_g > 0, TBI_ERR_BAD_SHAPE, "tapgemm: empty problem");
    const bool narrow = d->cout_g <= 16;
    const int tbn = narrow ? 16 : 64;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->cout_g + tbn - 1) / tbn), (unsigned)d->groups);
    if (d->dtype == TBI_F32) {
        if (narrow) tapgemm_simt_kernel<float, 16><<<grid, NT, 0, s>>>(*d);
        else        tapgemm_simt_kernel<float, 64><<<grid, NT, 0, s>>>(*d);
    } else if (d->dtype == TBI_BF16) {
        if (narrow) tapgemm_simt_kernel<__nv_bfloat16, 16><<<grid, NT, 0, s>>>(*d);
        else        tapgemm_simt_kernel<__nv_bfloat16, 64><<<grid, NT, 0, s>>>(*d);
    } else {
        return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapgemm: dtype %d", d->dtype);
    }
    TBI_CUDA_LAUNCH_CHECK("tapgemm_simt");
    return TBI_OK;
}

int tbi_tapwgrad_simt(const tbi_tapwgrad* d, cudaStream_t s) {
    TBI_CHECK(d->ntaps >= 1 && d->ntaps <= TBI_MAX_TAPS, TBI_ERR_BAD_SHAPE, "tapwgrad: ntaps=%d", d->ntaps);
    TBI_CHECK(d->groups >= 1 && (d->groups == 1 || d->a_src[1].ptr == nullptr), TBI_ERR_UNSUPPORTED,
              "tapwgrad: groups>1 needs a single source");
    const int ctot = d->a_src[0].c + (d->a_src[1].ptr ? d->a_src[1].c : 0);
    TBI_CHECK(ctot == d->cin_g * d->groups, TBI_ERR_BAD_SHAPE, "tapwgrad: source channels %d != groups*cin_g %d",
              ctot, d->cin_g * d->groups);
    TBI_CHECK(d->b_src.c == d->cout_g * d->groups || (d->groups == 1 && d->b_src.c > d->cout_g), TBI_ERR_BAD_SHAPE, "tapwgrad: dz channels %d != groups*cout_g %d",
              d->b_src.c, d->cout_g * d->groups);
    const long long M = (long long)d->n * d->gh * d->gw;
    TBI_CHECK(M > 0, TBI_ERR_BAD_SHAPE, "tapwgrad: empty problem");
    const int ci_tiles = (d->cin_g + BM - 1) / BM, co_tiles = (d->cout_g + BN - 1) / BN;
    const long long tiles = (long long)d->ntaps * ci_tiles * co_tiles * d->groups;
    // aim for ~4 waves of CTAs; a CTA handles at least 256 pixels
    long long want = (4LL * tbi_sm_count() * 4 + tiles - 1) / tiles;
    long long chunks = (M + 255) / 256;
    if (want < chunks) chunks = want;
    if (chunks < 1) chunks = 1;
    long long ppc = (M + chunks - 1) / chunks;
    ppc = (ppc + BK - 1) / BK * BK;
    chunks = (M + ppc - 1) / ppc;
    dim3 grid((unsigned)chunks, (unsigned)(d->ntaps * ci_tiles * co_tiles), (unsigned)d->groups);
    if (d->dtype == TBI_F32) tapwgrad_simt_kernel<float><<<grid, NT, 0, s>>>(*d, (int)ppc, ci_tiles, co_tiles);
    else if (d->dtype == TBI_BF16) tapwgrad_simt_kernel<__nv_bfloat16><<<grid, NT, 0, s>>>(*d, (int)ppc, ci_tiles, co_tiles);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapwgrad: dtype %d", d->dtype);
    TBI_CUDA_LAUNCH_CHECK("tapwgrad_simt");
    return TBI_OK;
}
