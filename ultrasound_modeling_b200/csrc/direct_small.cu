// Direct (CUDA-core) kernels for convolutions whose input has a handful of channels (the 1 -> 16 stem
// conv, TBI_ResNest.py:83).  K = taps*cin <= 36 is far below one UMMA K-step worth of useful work and the
// layer is purely HBM-bound (writes 32 B per pixel), so: one thread per pixel, weights in shared memory,
// 16-byte stores; the weight gradient is a per-thread register reduction + warp shuffles + one atomic
// per (block, element).
#include "tbi_common.cuh"
#include <stdlib.h>

namespace {

constexpr int MAXK = 36;

template <typename T, int V> struct alignas(sizeof(T) * V) PackD { T v[V]; };

// exp for v <= 0: one FMUL + MUFU (same helper as the tensor-core epilogue)
__device__ __forceinline__ float exp_neg_fast_d(float v) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v * 1.4426950408889634f));
    return y;
}
template <int ACT> __device__ __forceinline__ float act_ct(float v) {
    if (ACT == TBI_ACT_ELU)   return v > 0.f ? v : exp_neg_fast_d(v) - 1.f;
    if (ACT == TBI_ACT_LRELU) return v > 0.f ? v : 0.3f * v;
    if (ACT == TBI_ACT_RELU)  return fmaxf(v, 0.f);
    return v;
}

// weights in shared memory as [k][COUT] so one LDS.128 feeds four FMAs; activation is a template parameter
template <typename T, int COUT, int ACT>
__global__ void __launch_bounds__(256) smallcin_fwd_kernel(const __grid_constant__ tbi_tapgemm d) {
    __shared__ __align__(16) float ws[MAXK * COUT];
    __shared__ float bs[COUT];
    const int cin = d.cin_g, Kg = d.ntaps * cin;
    for (int i = threadIdx.x; i < COUT * Kg; i += blockDim.x) { const int c = i / Kg, k = i % Kg; ws[k * COUT + c] = ldf((const T*)d.w + i); }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = d.epi.bias ? d.epi.bias[i] : 0.f;
    __syncthreads();
    const tbi_epilogue& e = d.epi;
    const bool fast = e.drop_keep == nullptr && e.residual.ptr == nullptr && e.dact == TBI_ACT_NONE && e.split_c == 0 && !e.out_f32 &&
                      e.out.cstride == COUT && e.out.coff == 0 && e.out_stride == 1 && (((uintptr_t)e.out.ptr) & 15) == 0;
    const long long M = (long long)d.n * d.gh * d.gw;
    const T* src = (const T*)d.src[0].ptr;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < M; p += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(p % d.gw); long long t = p / d.gw; const int gy = (int)(t % d.gh); const int n = (int)(t / d.gh);
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
        for (int tap = 0; tap < d.ntaps; ++tap) {
            const int iy = gy + d.dy[tap], ix = gx + d.dx[tap];
            if (iy < 0 || iy >= d.src[0].h || ix < 0 || ix >= d.src[0].w) continue;
            const T* px = src + view_off(d.src[0], n, iy, ix, 0);
            for (int ci = 0; ci < cin; ++ci) {
                const float a = ldf(px + ci);
                const float4* wk = reinterpret_cast<const float4*>(ws + (tap * cin + ci) * COUT);
#pragma unroll
                for (int c4 = 0; c4 < COUT / 4; ++c4) {
                    const float4 w4 = wk[c4];
                    acc[4 * c4] = fmaf(a, w4.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(a, w4.y, acc[4 * c4 + 1]);
                    acc[4 * c4 + 2] = fmaf(a, w4.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(a, w4.w, acc[4 * c4 + 3]);
                }
            }
        }
        if (fast) {
            constexpr int V = 16 / sizeof(T);
            T* o = (T*)e.out.ptr + (size_t)p * COUT;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += V) {
                PackD<T, V> q;
#pragma unroll
                for (int j = 0; j < V; ++j) stf(&q.v[j], act_ct<ACT>(acc[c0 + j] + bs[c0 + j]));
                *reinterpret_cast<PackD<T, V>*>(o + c0) = q;
            }
        } else {
            const int oy = gy * e.out_stride + e.out_off_y, ox = gx * e.out_stride + e.out_off_x;
#pragma unroll
            for (int c = 0; c < COUT; ++c) epilogue_store<T>(e, n, oy, ox, c, acc[c]);
        }
    }
}

// one pass over dz: a thread accumulates all NT taps x COUT outputs for its pixels (CIN = 1), then warp shuffles,
// a block reduction in shared memory and one atomic per (block, element)
template <typename T, int COUT, int NT>
__global__ void __launch_bounds__(256) smallcin_wgrad_kernel(const __grid_constant__ tbi_tapwgrad d) {
    __shared__ float red[8][NT * COUT + COUT];
    float acc[NT][COUT];
    float bacc[COUT];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[t][c] = 0.f;
#pragma unroll
    for (int c = 0; c < COUT; ++c) bacc[c] = 0.f;
    const long long M = (long long)d.n * d.gh * d.gw;
    const T* a = (const T*)d.a_src[0].ptr;
    const T* b = (const T*)d.b_src.ptr;
    constexpr int V = 16 / sizeof(T);
    const bool vec = d.b_src.cstride % V == 0 && d.b_src.coff % V == 0 && (((uintptr_t)d.b_src.ptr) & 15) == 0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < M; p += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(p % d.gw); long long t = p / d.gw; const int gy = (int)(t % d.gh); const int n = (int)(t / d.gh);
        float g[COUT];
        const T* bp = b + view_off(d.b_src, n, gy, gx, 0);
        if (vec) {
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += V) {
                const PackD<T, V> q = *reinterpret_cast<const PackD<T, V>*>(bp + c0);
#pragma unroll
                for (int j = 0; j < V; ++j) g[c0 + j] = ldf(&q.v[j]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < COUT; ++c) g[c] = ldf(bp + c);
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) bacc[c] += g[c];
#pragma unroll
        for (int tp = 0; tp < NT; ++tp) {
            const int ay = gy + d.a_dy[tp], ax = gx + d.a_dx[tp];
            float x = 0.f;
            if (ay >= 0 && ay < d.a_src[0].h && ax >= 0 && ax < d.a_src[0].w) x = ldf(a + view_off(d.a_src[0], n, ay, ax, 0));
#pragma unroll
            for (int c = 0; c < COUT; ++c) acc[tp][c] = fmaf(x, g[c], acc[tp][c]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int tp = 0; tp < NT; ++tp)
#pragma unroll
        for (int c = 0; c < COUT; ++c) { const float v = warp_sum(acc[tp][c]); if (lane == 0) red[warp][tp * COUT + c] = v; }
#pragma unroll
    for (int c = 0; c < COUT; ++c) { const float v = warp_sum(bacc[c]); if (lane == 0) red[warp][NT * COUT + c] = v; }
    __syncthreads();
    for (int i = threadIdx.x; i < NT * COUT + COUT; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][i];
        if (i < NT * COUT) {
            const int tp = i / COUT, co = i % COUT;
            atomicAdd(d.dw + (size_t)tp * d.tap_stride + (size_t)co * d.co_stride, v);
        } else if (d.dbias) {
            atomicAdd(d.dbias + (i - NT * COUT), v);
        }
    }
}


// Row-walking form of the above for the standard 3x3 / pad 1 pattern on dense tensors: a thread owns SEG consecutive pixels
// of one image row and slides a 3x3 register window of x along it, so a pixel costs 3 scalar loads of x + 32 bytes of dz +
// the 144 FMAs -- no per-tap address arithmetic or bounds checks (the per-pixel form above spends ~300 instructions per
// pixel, most of them addressing).  The next pixel's loads are issued before the current pixel's FMAs.
template <int SEG>
__global__ void __launch_bounds__(256) smallcin_wgrad_rows_kernel(const __grid_constant__ tbi_tapwgrad d) {
    constexpr int COUT = 16, NT = 9;
    typedef __nv_bfloat16 T;
    __shared__ float red[8][NT * COUT + COUT];
    float acc[NT][COUT];
    float bacc[COUT];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[t][c] = 0.f;
#pragma unroll
    for (int c = 0; c < COUT; ++c) bacc[c] = 0.f;
    const int W = d.gw, H = d.gh, segs_x = W / SEG;
    const long long nseg = (long long)d.n * H * segs_x;
    const T* a = (const T*)d.a_src[0].ptr;
    const T* b = (const T*)d.b_src.ptr;
    for (long long sgi = blockIdx.x * (long long)blockDim.x + threadIdx.x; sgi < nseg; sgi += (long long)gridDim.x * blockDim.x) {
        const int xs = (int)(sgi % segs_x) * SEG; long long t = sgi / segs_x; const int y = (int)(t % H); const int n = (int)(t / H);
        const T* ar[3]; bool av[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) { const int yy = y + r - 1; av[r] = yy >= 0 && yy < H; ar[r] = a + ((size_t)n * H + (av[r] ? yy : y)) * W; }
        const uint4* bp = reinterpret_cast<const uint4*>(b + (((size_t)n * H + y) * W + xs) * COUT);
        float win[3][3];                                   // [row][column x-1, x, x+1]
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            win[r][0] = (av[r] && xs > 0) ? ldf(ar[r] + xs - 1) : 0.f;
            win[r][1] = av[r] ? ldf(ar[r] + xs) : 0.f;
        }
        uint4 g0 = bp[0], g1 = bp[1];
        float nx[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) nx[r] = (av[r] && xs + 1 < W) ? ldf(ar[r] + xs + 1) : 0.f;
#pragma unroll 2
        for (int i = 0; i < SEG; ++i) {
            const uint4 c0 = g0, c1 = g1;
#pragma unroll
            for (int r = 0; r < 3; ++r) win[r][2] = nx[r];
            if (i + 1 < SEG) {                             // prefetch pixel i+1
                g0 = bp[2 * (i + 1)]; g1 = bp[2 * (i + 1) + 1];
                const int xn = xs + i + 2;
#pragma unroll
                for (int r = 0; r < 3; ++r) nx[r] = (av[r] && xn < W) ? ldf(ar[r] + xn) : 0.f;
            }
            float g[COUT];
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&c0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&c1);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float2 f0 = __bfloat1622float2(h0[q]), f1 = __bfloat1622float2(h1[q]); g[2 * q] = f0.x; g[2 * q + 1] = f0.y; g[8 + 2 * q] = f1.x; g[8 + 2 * q + 1] = f1.y; }
#pragma unroll
            for (int c = 0; c < COUT; ++c) bacc[c] += g[c];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cx = 0; cx < 3; ++cx) {
                    const float x = win[r][cx];
#pragma unroll
                    for (int c = 0; c < COUT; ++c) acc[r * 3 + cx][c] = fmaf(x, g[c], acc[r * 3 + cx][c]);
                }
#pragma unroll
            for (int r = 0; r < 3; ++r) { win[r][0] = win[r][1]; win[r][1] = win[r][2]; }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int tp = 0; tp < NT; ++tp)
#pragma unroll
        for (int c = 0; c < COUT; ++c) { const float v = warp_sum(acc[tp][c]); if (lane == 0) red[warp][tp * COUT + c] = v; }
#pragma unroll
    for (int c = 0; c < COUT; ++c) { const float v = warp_sum(bacc[c]); if (lane == 0) red[warp][NT * COUT + c] = v; }
    __syncthreads();
    for (int i = threadIdx.x; i < NT * COUT + COUT; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][i];
        if (i < NT * COUT) {
            const int tp = i / COUT, co = i % COUT;
            atomicAdd(d.dw + (size_t)tp * d.tap_stride + (size_t)co * d.co_stride, v);
        } else if (d.dbias) {
            atomicAdd(d.dbias + (i - NT * COUT), v);
        }
    }
}

}  // namespace

bool tbi_tapgemm_direct_supported(const tbi_tapgemm* d) {
    return d->groups == 1 && d->src[1].ptr == nullptr && d->in_stride == 1 && d->nphase <= 1 && d->cout_g == 16 &&
           d->ntaps * d->cin_g <= MAXK && d->src[0].c == d->cin_g && (d->dtype == TBI_F32 || d->dtype == TBI_BF16);
}

int tbi_tapgemm_direct(const tbi_tapgemm* d, cudaStream_t s) {
    TBI_CHECK(tbi_tapgemm_direct_supported(d), TBI_ERR_UNSUPPORTED, "direct small-cin conv: unsupported shape");
    const long long M = (long long)d->n * d->gh * d->gw;
    long long blocks = (M + 255) / 256;
    const long long cap = (long long)tbi_sm_count() * 16;
    if (blocks > cap) blocks = cap;
#define TBI_SMALLCIN(ACTV) do { if (d->dtype == TBI_F32) smallcin_fwd_kernel<float, 16, ACTV><<<(unsigned)blocks, 256, 0, s>>>(*d); \
                                else smallcin_fwd_kernel<__nv_bfloat16, 16, ACTV><<<(unsigned)blocks, 256, 0, s>>>(*d); } while (0)
    switch (d->epi.act) {
        case TBI_ACT_ELU:   TBI_SMALLCIN(TBI_ACT_ELU); break;
        case TBI_ACT_RELU:  TBI_SMALLCIN(TBI_ACT_RELU); break;
        case TBI_ACT_LRELU: TBI_SMALLCIN(TBI_ACT_LRELU); break;
        default:            TBI_SMALLCIN(TBI_ACT_NONE); break;
    }
#undef TBI_SMALLCIN
    TBI_CUDA_LAUNCH_CHECK("smallcin_fwd");
    return TBI_OK;
}

bool tbi_tapwgrad_direct_supported(const tbi_tapwgrad* d) {
    if (!(d->groups == 1 && d->a_src[1].ptr == nullptr && d->a_stride == 1 && d->b_stride == 1 && d->cout_g == 16 && d->cin_g == 1 &&
          d->a_src[0].c == 1 && d->b_src.c == 16 && d->ntaps == 9 && (d->dtype == TBI_F32 || d->dtype == TBI_BF16))) return false;
    for (int t = 0; t < d->ntaps; ++t) if (d->b_dy[t] != 0 || d->b_dx[t] != 0) return false;
    return true;
}

int tbi_tapwgrad_direct(const tbi_tapwgrad* d, cudaStream_t s) {
    TBI_CHECK(tbi_tapwgrad_direct_supported(d), TBI_ERR_UNSUPPORTED, "direct small-cin wgrad: unsupported shape");
    const long long M = (long long)d->n * d->gh * d->gw;
    long long blocks = (M + 256 * 8 - 1) / (256 * 8);
    const long long cap = (long long)tbi_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    {   // row-walking kernel: bf16, dense 1-channel x and 16-channel dz, standard 3x3 taps, width a multiple of 32
        static const bool off = getenv("TBI_SMALLCIN_NO_ROWS") != nullptr;
        bool std_taps = d->ntaps == 9;
        for (int t = 0; t < 9 && std_taps; ++t) std_taps = d->a_dy[t] == t / 3 - 1 && d->a_dx[t] == t % 3 - 1;
        const tbi_view& A = d->a_src[0]; const tbi_view& B = d->b_src;
        if (!off && d->dtype == TBI_BF16 && std_taps && A.cstride == 1 && A.coff == 0 && A.h == d->gh && A.w == d->gw &&
            B.cstride == 16 && B.coff == 0 && B.h == d->gh && B.w == d->gw && d->gw % 32 == 0 && ((uintptr_t)B.ptr & 15) == 0) {
            const long long nseg = (long long)d->n * d->gh * (d->gw / 32);
            long long nb = (nseg + 255) / 256;
            const long long capb = (long long)tbi_sm_count() * 4;
            if (nb > capb) nb = capb;
            smallcin_wgrad_rows_kernel<32><<<(unsigned)nb, 256, 0, s>>>(*d);
            TBI_CUDA_LAUNCH_CHECK("smallcin_wgrad_rows");
            return TBI_OK;
        }
    }
    if (d->dtype == TBI_F32) smallcin_wgrad_kernel<float, 16, 9><<<(unsigned)blocks, 256, 0, s>>>(*d);
    else smallcin_wgrad_kernel<__nv_bfloat16, 16, 9><<<(unsigned)blocks, 256, 0, s>>>(*d);
    TBI_CUDA_LAUNCH_CHECK("smallcin_wgrad");
    return TBI_OK;
}
