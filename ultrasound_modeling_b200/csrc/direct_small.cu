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

// Row-strip form of the forward kernel for the layer that matters (TBI_ResNest.py:83: 1 -> 16 channels, 3x3, dense bf16
// tensors): a thread owns SEG consecutive pixels of one image row, keeps the 3 x (SEG+2) input window in registers and
// reads each weight vector from shared memory once per strip instead of once per pixel; no per-tap address arithmetic,
// bounds checks only at the strip ends, one 32-byte store (a whole sector) per pixel.  The per-pixel kernel above executed
// ~890 instructions per pixel (ncu: 117 M warp instructions, issue slots 72 % busy, 149 us for a layer whose HBM floor is
// 21 us); this one ~250.
template <int SEG, int ACT>
__global__ void __launch_bounds__(256) smallcin_fwd_rows_kernel(const __grid_constant__ tbi_tapgemm d) {
    constexpr int COUT = 16;
    typedef __nv_bfloat16 T;
    __shared__ __align__(16) float ws[9 * COUT];
    __shared__ __align__(16) float bs[COUT];
    for (int i = threadIdx.x; i < COUT * 9; i += blockDim.x) { const int c = i / 9, k = i % 9; ws[k * COUT + c] = ldf((const T*)d.w + i); }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = d.epi.bias ? d.epi.bias[i] : 0.f;
    __syncthreads();
    const int W = d.gw, H = d.gh, segs_x = W / SEG;
    const long long nseg = (long long)d.n * H * segs_x;
    const T* src = (const T*)d.src[0].ptr;
    T* out = (T*)d.epi.out.ptr;
    for (long long sgi = blockIdx.x * (long long)blockDim.x + threadIdx.x; sgi < nseg; sgi += (long long)gridDim.x * blockDim.x) {
        const int xs = (int)(sgi % segs_x) * SEG; long long t = sgi / segs_x; const int y = (int)(t % H); const int n = (int)(t / H);
        float win[3][SEG + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int yy = y + r - 1;
            const bool rv = yy >= 0 && yy < H;
            const T* row = src + ((size_t)n * H + (rv ? yy : y)) * W + xs;
            win[r][0] = (rv && xs > 0) ? ldf(row - 1) : 0.f;
#pragma unroll
            for (int i = 0; i < SEG; ++i) win[r][i + 1] = rv ? ldf(row + i) : 0.f;
            win[r][SEG + 1] = (rv && xs + SEG < W) ? ldf(row + SEG) : 0.f;
        }
        float acc[SEG][COUT];
#pragma unroll
        for (int i = 0; i < SEG; ++i)
#pragma unroll
            for (int c = 0; c < COUT; ++c) acc[i][c] = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cx = 0; cx < 3; ++cx) {
                const float4* wk = reinterpret_cast<const float4*>(ws + (r * 3 + cx) * COUT);
#pragma unroll
                for (int c4 = 0; c4 < COUT / 4; ++c4) {
                    const float4 w4 = wk[c4];
#pragma unroll
                    for (int i = 0; i < SEG; ++i) {
                        const float a = win[r][i + cx];
                        acc[i][4 * c4] = fmaf(a, w4.x, acc[i][4 * c4]); acc[i][4 * c4 + 1] = fmaf(a, w4.y, acc[i][4 * c4 + 1]);
                        acc[i][4 * c4 + 2] = fmaf(a, w4.z, acc[i][4 * c4 + 2]); acc[i][4 * c4 + 3] = fmaf(a, w4.w, acc[i][4 * c4 + 3]);
                    }
                }
            }
        T* o = out + (((size_t)n * H + y) * W + xs) * COUT;
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
            uint32_t q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // bias after the taps, like the per-pixel kernel: the two forms are bit-identical
                const __nv_bfloat162 h = __floats2bfloat162_rn(act_ct<ACT>(acc[i][2 * j] + bs[2 * j]), act_ct<ACT>(acc[i][2 * j + 1] + bs[2 * j + 1]));
                q[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + (size_t)i * COUT), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]),
                         "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
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

// Column-walking form: the 32 lanes of a warp are 16 ADJACENT pixels of one image row x the two halves of the 16 output
// channels, and a thread walks SEG rows DOWN its column.  Every load is then coalesced (a warp reads 512 contiguous bytes of dz
// and 32 contiguous bytes of x per step), where the row-walking kernel above had each lane on its own 1 KB-strided row segment
// (ncu: 5x sector over-fetch in L1TEX, L1TEX 69 % busy, 254 registers, 96 us for a layer whose HBM floor is 22 us).  Half the
// channels per thread halves the accumulator registers (72 + 8), so two blocks fit an SM.
template <int SEG>
__global__ void __launch_bounds__(256, 2) smallcin_wgrad_cols_kernel(const __grid_constant__ tbi_tapwgrad d) {
    constexpr int COUT = 16, CT = 8, NT = 9, DEPTH = 8;
    typedef __nv_bfloat16 T;
    __shared__ float red[8][NT * COUT + COUT];
    __shared__ uint4 ring[DEPTH][256];
    float acc[NT][CT];
    float bacc[CT];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[t][c] = 0.f;
#pragma unroll
    for (int c = 0; c < CT; ++c) bacc[c] = 0.f;
    const int W = d.gw, H = d.gh, segs_y = H / SEG;
    const long long nthr = (long long)d.n * segs_y * W * 2;
    const T* a = (const T*)d.a_src[0].ptr;
    const T* b = (const T*)d.b_src.ptr;
    for (long long ti = blockIdx.x * (long long)blockDim.x + threadIdx.x; ti < nthr; ti += (long long)gridDim.x * blockDim.x) {
        const int half = (int)(ti & 1);
        long long t = ti >> 1;
        const int x = (int)(t % W); t /= W;
        const int ys = (int)(t % segs_y) * SEG; const int n = (int)(t / segs_y);
        const T* ap = a + (size_t)n * H * W;                       // image n of x
        const uint4* bp = reinterpret_cast<const uint4*>(b + (((size_t)n * H + ys) * W + x) * COUT + half * CT);
        const size_t bstep = (size_t)W * COUT * sizeof(T) / sizeof(uint4);
        const bool xl = x > 0, xr = x + 1 < W;
        auto load_row = [&](int yy, float (&v)[3]) {
            const bool rv = yy >= 0 && yy < H;
            const T* row = ap + (size_t)(rv ? yy : 0) * W + x;
            v[0] = (rv && xl) ? ldf(row - 1) : 0.f; v[1] = rv ? ldf(row) : 0.f; v[2] = (rv && xr) ? ldf(row + 1) : 0.f;
        };
        float win[3][3];                                   // [row y-1, y, y+1][column x-1, x, x+1]
        load_row(ys - 1, win[0]); load_row(ys, win[1]);
        // dz rows arrive through a per-thread ring of DEPTH 16-byte slots in shared memory filled by cp.async: each thread copies
        // exactly the bytes it consumes itself, so no barrier is needed, and DEPTH-1 rows per thread (64 KB per SM) are in
        // flight without holding registers (with one register-prefetched row the kernel ran at 1.4 TB/s: latency-bound)
        const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(&ring[0][threadIdx.x]);
#pragma unroll
        for (int k = 0; k < DEPTH - 1; ++k) {
            if (k < SEG) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot0 + (uint32_t)(k * 256 * 16)), "l"(bp + (size_t)k * bstep) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // x: every lane loads only ITS pixel of the incoming window row, XQ rows ahead (a queue in registers: the loads are L2
        // hits with ~1 000 cycles of latency, one step is ~500), and takes the two neighbours from the adjacent pixels' lanes;
        // the first / last pixel of the warp's 16 load the pixel outside the warp as well
        constexpr int XQ = 4;
        const int pl = (int)(threadIdx.x & 31) >> 1;       // pixel within the warp's 16 (== x & 15)
        const bool e_lo = pl == 0, e_hi = pl == 15;
        // the queue holds the RAW 16-bit values: converting at load time (ldf) makes the load's consumer the next instruction and
        // the thread waits for every load where it is issued (ncu: 52 % of all stall samples on that one shift)
        const unsigned short* au = reinterpret_cast<const unsigned short*>(ap);
        auto own = [&](int yy) -> uint32_t { return (yy >= 0 && yy < H) ? (uint32_t)au[(size_t)yy * W + x] : 0u; };
        auto edge = [&](int yy) -> uint32_t {
            const bool rv = yy >= 0 && yy < H;
            if (e_lo) return (rv && xl) ? (uint32_t)au[(size_t)yy * W + x - 1] : 0u;
            if (e_hi) return (rv && xr) ? (uint32_t)au[(size_t)yy * W + x + 1] : 0u;
            return 0u;
        };
        uint32_t xq[XQ], eq[XQ];
#pragma unroll
        for (int j = 0; j < XQ; ++j) { xq[j] = own(ys + 1 + j); eq[j] = edge(ys + 1 + j); }
#pragma unroll 1
        for (int i0 = 0; i0 < SEG; i0 += XQ)
#pragma unroll
        for (int j = 0; j < XQ; ++j) {
            const int i = i0 + j;
            {
                const int k = i + DEPTH - 1;
                if (k < SEG) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot0 + (uint32_t)((k % DEPTH) * 256 * 16)), "l"(bp + (size_t)k * bstep) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            {
                const float c = __uint_as_float(xq[j] << 16), e = __uint_as_float(eq[j] << 16);
                const float l = __shfl_up_sync(0xffffffffu, c, 2), r = __shfl_down_sync(0xffffffffu, c, 2);
                win[2][0] = e_lo ? e : l; win[2][1] = c; win[2][2] = e_hi ? e : r;
                xq[j] = own(ys + i + 1 + XQ); eq[j] = edge(ys + i + 1 + XQ);
            }
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
            const uint4 c0 = ring[i % DEPTH][threadIdx.x];
            float g[CT];
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float2 f0 = __bfloat1622float2(h0[q]); g[2 * q] = f0.x; g[2 * q + 1] = f0.y; }
#pragma unroll
            for (int c = 0; c < CT; ++c) bacc[c] += g[c];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cx = 0; cx < 3; ++cx) {
                    const float xv = win[r][cx];
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc[r * 3 + cx][c] = fmaf(xv, g[c], acc[r * 3 + cx][c]);
                }
#pragma unroll
            for (int cx = 0; cx < 3; ++cx) { win[0][cx] = win[1][cx]; win[1][cx] = win[2][cx]; }
        }
    }
    // lanes of equal parity hold the same channel half: butterfly over offsets 16..2, lanes 0 / 1 end up with the two halves
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto half_sum = [](float v) { for (int o = 16; o > 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
#pragma unroll
    for (int tp = 0; tp < NT; ++tp)
#pragma unroll
        for (int c = 0; c < CT; ++c) { const float v = half_sum(acc[tp][c]); if (lane < 2) red[warp][tp * COUT + lane * CT + c] = v; }
#pragma unroll
    for (int c = 0; c < CT; ++c) { const float v = half_sum(bacc[c]); if (lane < 2) red[warp][NT * COUT + lane * CT + c] = v; }
    __syncthreads();
    constexpr int TOT = NT * COUT + COUT;
    for (int k = threadIdx.x; k < TOT; k += blockDim.x) {
        const int i = (int)((k + blockIdx.x * 37u) % (unsigned)TOT);       // blocks start at different words (they finish together)
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
    {   // row-strip kernel: bf16, one dense input channel, 16 dense output channels, standard 3x3 taps, plain epilogue
        static const bool off = getenv("TBI_SMALLCIN_NO_ROWS") != nullptr;
        const tbi_epilogue& e = d->epi;
        const tbi_view& A = d->src[0];
        bool ok = !off && d->dtype == TBI_BF16 && d->cin_g == 1 && d->ntaps == 9 && A.cstride == 1 && A.coff == 0 && A.h == d->gh && A.w == d->gw &&
                  d->gw % 4 == 0 && e.drop_keep == nullptr && e.residual.ptr == nullptr && e.dact == TBI_ACT_NONE && e.split_c == 0 && !e.out_f32 &&
                  e.out.cstride == 16 && e.out.coff == 0 && e.out.c == 16 && e.out.h == d->gh && e.out.w == d->gw && e.out_stride == 1 &&
                  e.out_off_y == 0 && e.out_off_x == 0 && (((uintptr_t)e.out.ptr) & 31) == 0 && (e.act == TBI_ACT_ELU || e.act == TBI_ACT_NONE);
        for (int t = 0; t < 9 && ok; ++t) ok = d->dy[t] == t / 3 - 1 && d->dx[t] == t % 3 - 1;
        if (ok) {
            const long long nseg = M / 4;
            long long nb = (nseg + 255) / 256;
            const long long capb = (long long)tbi_sm_count() * 16;
            if (nb > capb) nb = capb;
            if (e.act == TBI_ACT_ELU) smallcin_fwd_rows_kernel<4, TBI_ACT_ELU><<<(unsigned)nb, 256, 0, s>>>(*d);
            else smallcin_fwd_rows_kernel<4, TBI_ACT_NONE><<<(unsigned)nb, 256, 0, s>>>(*d);
            TBI_CUDA_LAUNCH_CHECK("smallcin_fwd_rows");
            return TBI_OK;
        }
    }
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
            static const bool rows = getenv("TBI_SMALLCIN_WGRAD_ROWS") != nullptr;
            if (!rows && d->gh % 32 == 0 && d->gw % 16 == 0) {
                const long long nthr = (long long)d->n * (d->gh / 32) * d->gw * 2;
                long long nb = (nthr + 255) / 256;
                // two resident blocks per SM (128 registers, 37 KB shared memory): more blocks add no parallelism, only more final
                // atomics on the same 160 words
                const long long capb = (long long)tbi_sm_count() * 2;
                if (nb > capb) nb = capb;
                smallcin_wgrad_cols_kernel<32><<<(unsigned)nb, 256, 0, s>>>(*d);
                TBI_CUDA_LAUNCH_CHECK("smallcin_wgrad_cols");
                return TBI_OK;
            }
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
