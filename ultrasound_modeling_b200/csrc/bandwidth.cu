// HBM-bound kernels of the TBI_ResNest path: pooling, split-attention tail, softmax+loss,
// activation backward, bias/BN parameter gradients, weight packing, Adam.
// All NHWC, 16-byte vector access along channels when the views allow it, fp32 math.
#include "tbi_common.cuh"
#include <initializer_list>

namespace {

template <typename T, int V> struct alignas(sizeof(T) * V) Pack { T v[V]; };

template <typename T, int V> __device__ __forceinline__ void ld_pack(const T* p, float (&f)[V]) {
    Pack<T, V> q = *reinterpret_cast<const Pack<T, V>*>(p);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = ldf(&q.v[i]);
}
template <typename T, int V> __device__ __forceinline__ void st_pack(T* p, const float (&f)[V]) {
    Pack<T, V> q;
#pragma unroll
    for (int i = 0; i < V; ++i) stf(&q.v[i], f[i]);
    *reinterpret_cast<Pack<T, V>*>(p) = q;
}

inline bool view_vec_ok(const tbi_view* v, int V, int esz) {
    if (!v || !v->ptr) return true;
    return v->c % V == 0 && v->cstride % V == 0 && v->coff % V == 0 && ((uintptr_t)v->ptr % (size_t)(V * esz)) == 0;
}
inline int pick_vec(int dtype, std::initializer_list<const tbi_view*> views) {
    const int V = dtype == TBI_F32 ? 4 : 8, esz = tbi_dtype_size(dtype);
    for (auto v : views) if (!view_vec_ok(v, V, esz)) return 1;
    return V;
}
inline unsigned grid_for(long long work, int threads, int max_waves = 32) {
    long long b = (work + threads - 1) / threads;
    long long cap = (long long)tbi_sm_count() * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---------------------------------------------------------------------------------------------
// average pooling 2x2
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void avgpool_fwd_kernel(int n, int ho, int wo, tbi_view x, tbi_view y) {
    const int cv = y.c / V;
    const long long total = (long long)n * ho * wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * V; long long t = i / cv;
        const int ox = (int)(t % wo); t /= wo; const int oy = (int)(t % ho); const int b = (int)(t / ho);
        float a[V], s[V];
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                ld_pack<T, V>((const T*)x.ptr + view_off(x, b, 2 * oy + dy, 2 * ox + dx, c), a);
#pragma unroll
                for (int k = 0; k < V; ++k) s[k] += a[k];
            }
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] *= 0.25f;
        st_pack<T, V>((T*)y.ptr + view_off(y, b, oy, ox, c), s);
    }
}

template <typename T, int V>
__global__ void avgpool_bwd_kernel(int n, int ho, int wo, tbi_view dy, tbi_view dx, int accumulate, int dact, tbi_view ref) {
    const int cv = dy.c / V;
    const long long total = (long long)n * ho * wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * V; long long t = i / cv;
        const int ox = (int)(t % wo); t /= wo; const int oy = (int)(t % ho); const int b = (int)(t / ho);
        float g[V];
        ld_pack<T, V>((const T*)dy.ptr + view_off(dy, b, oy, ox, c), g);
#pragma unroll
        for (int k = 0; k < V; ++k) g[k] *= 0.25f;
#pragma unroll
        for (int ay = 0; ay < 2; ++ay)
#pragma unroll
            for (int ax = 0; ax < 2; ++ax) {
                float o[V];
#pragma unroll
                for (int k = 0; k < V; ++k) o[k] = g[k];
                T* dst = (T*)dx.ptr + view_off(dx, b, 2 * oy + ay, 2 * ox + ax, c);
                if (dact != TBI_ACT_NONE) {
                    float r[V];
                    ld_pack<T, V>((const T*)ref.ptr + view_off(ref, b, 2 * oy + ay, 2 * ox + ax, c), r);
#pragma unroll
                    for (int k = 0; k < V; ++k) o[k] *= act_grad_from_out(dact, r[k]);
                }
                if (accumulate) {
                    float old[V];
                    ld_pack<T, V>(dst, old);
#pragma unroll
                    for (int k = 0; k < V; ++k) o[k] += old[k];
                }
                st_pack<T, V>(dst, o);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// activation backward, column sum
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void act_bwd_kernel(long long npix, int act, tbi_view dy, tbi_view yr, const uint8_t* keep, tbi_view dz) {
    const int cv = dz.c / V;
    const long long total = npix * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * V; const long long p = i / cv;
        float g[V], r[V];
        ld_pack<T, V>((const T*)dy.ptr + (size_t)p * dy.cstride + dy.coff + c, g);
        ld_pack<T, V>((const T*)yr.ptr + (size_t)p * yr.cstride + yr.coff + c, r);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            g[k] *= act_grad_from_out(act, r[k]);
            if (keep) g[k] *= (float)keep[(size_t)p * dz.c + c + k];
        }
        st_pack<T, V>((T*)dz.ptr + (size_t)p * dz.cstride + dz.coff + c, g);
    }
}

// G[n,i,j,(ky*k+kx)*cout+co] = dz[n,2i-pad+ky,2j-pad+kx,co]; one thread per (pixel, tap)
template <typename T>
__global__ void convt_gather_kernel(int n, int h, int w, int k, int pad, int cout, tbi_view dz, tbi_view g) {
    const long long total = (long long)n * h * w * k * k;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(i % (k * k)); long long t = i / (k * k);
        const int x = (int)(t % w); t /= w; const int y = (int)(t % h); const int b = (int)(t / h);
        const int sy = 2 * y - pad + tap / k, sx = 2 * x - pad + tap % k;
        T* dst = (T*)g.ptr + view_off(g, b, y, x, tap * cout);
        if (sy >= 0 && sy < dz.h && sx >= 0 && sx < dz.w) {
            const T* src = (const T*)dz.ptr + view_off(dz, b, sy, sx, 0);
            for (int c = 0; c < cout; ++c) dst[c] = src[c];
        } else {
            for (int c = 0; c < cout; ++c) stf(dst + c, 0.f);
        }
    }
}

// bf16, k = 4, cout <= 4, dz records of >= 4 channels (8-byte aligned), G records of exactly 64 channels at a 128-byte pitch:
// one thread per G pixel gathers its 16 source pixels (8 bytes each) and writes the 128-byte record with eight 16-byte stores
// (channels 16*cout .. 63 are written as zeros, which the contract allows: the consumers' weights are zero there).
template <int COUT>
__global__ void __launch_bounds__(256) convt_gather4_bf16_kernel(int n, int h, int w, tbi_view dz, tbi_view g) {
    const long long total = (long long)n * h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w); long long t = i / w; const int y = (int)(t % h); const int b = (int)(t / h);
        __align__(16) __nv_bfloat16 rec[64];
#pragma unroll
        for (int q = 0; q < 64; ++q) rec[q] = __float2bfloat16_rn(0.f);
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) {
            const int sy = 2 * y - 1 + tap / 4, sx = 2 * x - 1 + tap % 4;
            if (sy >= 0 && sy < dz.h && sx >= 0 && sx < dz.w) {
                const uint2 v = *reinterpret_cast<const uint2*>((const __nv_bfloat16*)dz.ptr + view_off(dz, b, sy, sx, 0));
                const __nv_bfloat16* vv = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
                for (int c = 0; c < COUT; ++c) rec[tap * COUT + c] = vv[c];
            }
        }
        // four 32-byte stores (the record is 128-byte aligned): whole sectors, where eight 16-byte stores wrote every sector in two halves
        char* dst = reinterpret_cast<char*>((__nv_bfloat16*)g.ptr + view_off(g, b, y, x, 0));
        const uint4* r4 = reinterpret_cast<const uint4*>(rec);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 a = r4[2 * q], c = r4[2 * q + 1];
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 32 * q), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                         "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w) : "memory");
        }
    }
}

// The inverse direction (forward of a k = 4, stride-2 transposed conv computed as a plain GEMM per INPUT pixel):
// Y[n,i,j,(ky*4+kx)*COUT+co] = sum_ci X[n,i,j,ci] * W[ky][kx][co][ci]  (records of y.cstride channels, fp32 or bf16), and
// out[n,oy,ox,co] = bias[co] + sum over the 2 x 2 (ky,kx) with 2i - 1 + ky = oy, 2j - 1 + kx = ox of Y[n,i,j,..].
// One thread per OUTPUT pixel; the four records it reads are shared with its neighbours (L1/L2).
template <typename TY, int COUT>
__global__ void __launch_bounds__(256) convt_scatter4_kernel(int n, int h, int w, tbi_view y, const float* __restrict__ bias, tbi_view out) {
    const int H = 2 * h, W = 2 * w;
    const long long total = (long long)n * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % W); long long t = i / W; const int oy = (int)(t % H); const int b = (int)(t / H);
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = bias ? bias[c] : 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ky = ((oy + 1) & 1) + 2 * a, iy = (oy + 1 - ky) >> 1;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kx = ((ox + 1) & 1) + 2 * e, ix = (ox + 1 - kx) >> 1;
                if (ix < 0 || ix >= w) continue;
                const TY* src = (const TY*)y.ptr + view_off(y, b, iy, ix, (ky * 4 + kx) * COUT);
#pragma unroll
                for (int c = 0; c < COUT; ++c) acc[c] += ldf(src + c);
            }
        }
        float* dst = (float*)out.ptr + view_off(out, b, oy, ox, 0);
#pragma unroll
        for (int c = 0; c < COUT; ++c) dst[c] = acc[c];
    }
}

template <typename T, int V>
__global__ void accumulate_kernel(long long npix, tbi_view src, tbi_view dst) {
    const int cv = dst.c / V;
    const long long total = npix * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * V; const long long p = i / cv;
        float a[V], b[V];
        ld_pack<T, V>((const T*)src.ptr + (size_t)p * src.cstride + src.coff + c, a);
        T* d = (T*)dst.ptr + (size_t)p * dst.cstride + dst.coff + c;
        ld_pack<T, V>(d, b);
#pragma unroll
        for (int k = 0; k < V; ++k) b[k] += a[k];
        st_pack<T, V>(d, b);
    }
}

// out[c] += sum_p x[p,c].  block = 256 threads = (256/CL) pixel lanes x CL channel lanes
template <typename T>
__global__ void colsum_kernel(long long npix, tbi_view x, float* out, int pix_per_block) {
    extern __shared__ float sm[];
    const int C = x.c;
    const int cl = min(C, (int)blockDim.x);                 // channel lanes
    const int pl = blockDim.x / cl;                          // pixel lanes
    const int lane_c = threadIdx.x % cl, lane_p = threadIdx.x / cl;
    const long long pbeg = (long long)blockIdx.x * pix_per_block;
    const long long pend = min(npix, pbeg + pix_per_block);
    for (int c0 = 0; c0 < C; c0 += cl) {
        const int c = c0 + lane_c;
        float s = 0.f;
        if (c < C && lane_p < pl)
            for (long long p = pbeg + lane_p; p < pend; p += pl) s += ldf((const T*)x.ptr + (size_t)p * x.cstride + x.coff + c);
        sm[threadIdx.x] = s;
        __syncthreads();
        if (lane_p == 0 && c < C) {
            float tot = 0.f;
            for (int q = 0; q < pl; ++q) tot += sm[q * cl + lane_c];
            atomicAdd(out + c, tot);
        }
        __syncthreads();
    }
}

// vectorised column sum: a thread owns V consecutive channels and strides over pixels
template <typename T, int V>
__global__ void colsum_vec_kernel(long long npix, tbi_view x, float* out, int pix_per_block) {
    extern __shared__ float sm[];                            // [pl][cl*V]
    const int cv = x.c / V;
    const int cl = min(cv, (int)blockDim.x), pl = blockDim.x / cl;
    const int lane_c = threadIdx.x % cl, lane_p = threadIdx.x / cl;
    const long long pbeg = (long long)blockIdx.x * pix_per_block;
    const long long pend = min(npix, pbeg + pix_per_block);
    for (int cv0 = 0; cv0 < cv; cv0 += cl) {
        const int ch = (cv0 + lane_c) * V;
        float s[V];
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] = 0.f;
        if (ch < x.c && lane_p < pl) {
            // U independent 16-byte loads in flight per thread (a single dependent load per iteration left the kernel at ~35 % of HBM peak)
            constexpr int U = 8;
            const T* base = (const T*)x.ptr + x.coff + ch + (size_t)(pbeg + lane_p) * x.cstride;
            const size_t step = (size_t)pl * x.cstride;
            const int cnt = (int)((pend - pbeg - lane_p + pl - 1) / pl);          // pixels this thread visits
            int i = 0;
            for (; i + U <= cnt; i += U) {
                float a[U][V];
#pragma unroll
                for (int u = 0; u < U; ++u) ld_pack<T, V>(base + (size_t)(i + u) * step, a[u]);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < V; ++k) s[k] += a[u][k];
            }
            for (; i < cnt; ++i) {
                float a[V];
                ld_pack<T, V>(base + (size_t)i * step, a);
#pragma unroll
                for (int k = 0; k < V; ++k) s[k] += a[k];
            }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) sm[(size_t)threadIdx.x * V + k] = s[k];
        __syncthreads();
        if (lane_p == 0 && ch < x.c) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float tot = 0.f;
                for (int q = 0; q < pl; ++q) tot += sm[(size_t)(q * cl + lane_c) * V + k];
                atomicAdd(out + ch + k, tot);
            }
        }
        __syncthreads();
    }
}

// Dense tensors (cstride == c, c*sizeof(T) a multiple of 16 B, total thread count a multiple of the chunks per pixel):
// the tensor is one flat array of 16-byte chunks and a thread striding by the total thread count always lands on the
// same channel chunk, so it accumulates V channels in registers over perfectly coalesced loads, U of them in flight.
template <typename T, int V>
__global__ void __launch_bounds__(256) colsum_flat_kernel(long long nchunks, int cv, const T* __restrict__ x, float* out) {
    __shared__ float sm[256][V + 1];
    constexpr int U = 8;
    const long long T_all = (long long)gridDim.x * blockDim.x;
    const long long g0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float s[V];
#pragma unroll
    for (int k = 0; k < V; ++k) s[k] = 0.f;
    long long j = g0;
    for (; j + (U - 1) * T_all < nchunks; j += U * T_all) {
        Pack<T, V> q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = *reinterpret_cast<const Pack<T, V>*>(x + (size_t)(j + u * T_all) * V);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < V; ++k) s[k] += ldf(&q[u].v[k]);
    }
    for (; j < nchunks; j += T_all) {
        const Pack<T, V> q = *reinterpret_cast<const Pack<T, V>*>(x + (size_t)j * V);
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] += ldf(&q.v[k]);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) sm[threadIdx.x][k] = s[k];
    __syncthreads();
    // threads with equal (tid % cv) hold the same channel chunk: one thread per (chunk, k) sums its 256/cv rows
    const int chunk0 = (int)(((long long)blockIdx.x * blockDim.x) % cv);          // chunk id of thread 0 (T_all % cv == 0 => fixed per block)
    for (int e = threadIdx.x; e < cv * V; e += blockDim.x) {
        const int lc = e / V, k = e % V;                                            // lc = tid % cv of the owning threads
        float tot = 0.f;
        for (int t = lc; t < (int)blockDim.x; t += cv) tot += sm[t][k];
        atomicAdd(out + ((chunk0 + lc) % cv) * V + k, tot);
    }
}

// ---------------------------------------------------------------------------------------------
// split attention
// ---------------------------------------------------------------------------------------------
// raw[n][ch] += sum over a pixel chunk of  u[n,p,ch] (* dv[n,p,k*c+cc] when MUL).  ch over K*R*c.
template <typename T, int V, bool MUL>
__global__ void splitatt_reduce_kernel(int hw, int R, int c, tbi_view u, tbi_view dv, float* raw, int pix_per_block) {
    extern __shared__ float sm[];                            // [pl][cl*V]
    const int C = u.c, cv = C / V;
    const int cl = min(cv, (int)blockDim.x), pl = blockDim.x / cl;
    const int lane_c = threadIdx.x % cl, lane_p = threadIdx.x / cl;
    const int n = blockIdx.y;
    const int pbeg = blockIdx.x * pix_per_block, pend = min(hw, pbeg + pix_per_block);
    const T* ub = (const T*)u.ptr + (size_t)n * hw * u.cstride + u.coff;
    const T* db = MUL ? (const T*)dv.ptr + (size_t)n * hw * dv.cstride + dv.coff : nullptr;
    for (int cv0 = 0; cv0 < cv; cv0 += cl) {
        const int ch = (cv0 + lane_c) * V;
        float s[V];
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] = 0.f;
        if (ch < C && lane_p < pl) {
            const int kk = ch / (R * c), cc = ch % c;        // cardinal index, channel within cvkk
#pragma unroll 4
            for (int p = pbeg + lane_p; p < pend; p += pl) {
                float a[V];
                ld_pack<T, V>(ub + (size_t)p * u.cstride + ch, a);
                if (MUL) {
                    float g[V];
                    ld_pack<T, V>(db + (size_t)p * dv.cstride + kk * c + cc, g);
#pragma unroll
                    for (int k = 0; k < V; ++k) s[k] = fmaf(a[k], g[k], s[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < V; ++k) s[k] += a[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) sm[(size_t)threadIdx.x * V + k] = s[k];
        __syncthreads();
        if (lane_p == 0 && ch < C) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float tot = 0.f;
                for (int q = 0; q < pl; ++q) tot += sm[(size_t)(q * cl + lane_c) * V + k];
                atomicAdd(raw + (size_t)n * C + ch + k, tot);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { float t = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, t) : v + t; }
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

// one block per (n,k): gap -> dense1 -> BN -> act -> dense2 x R -> softmax_c / sigmoid
// STAGE: the cardinal's FC weights are first copied to shared memory with one coalesced cooperative load, so the
// dependent dot products below never wait on global memory
template <bool STAGE>
__global__ void splitatt_fc_kernel(tbi_splitatt p) {
    extern __shared__ float sm[];
    const int c = p.c, c2 = p.c / 2, R = p.radix, K = p.kpaths;
    float* g = sm;                 // [c]
    float* h1 = sm + c;            // [c2]
    float* red = h1 + c2;          // [32]
    float* sw1 = red + 32;         // [c][c2]      (STAGE)
    float* sw2 = sw1 + c * c2;     // [R][c2][c]   (STAGE)
    const int n = blockIdx.x, k = blockIdx.y;
    if (STAGE) {
        const float* gw1 = p.w1 + (size_t)k * c * c2; const float* gw2 = p.w2 + (size_t)k * R * c2 * c;
        for (int i = threadIdx.x; i < c * c2; i += blockDim.x) sw1[i] = __ldg(gw1 + i);
        for (int i = threadIdx.x; i < R * c2 * c; i += blockDim.x) sw2[i] = __ldg(gw2 + i);
    }
    float* att = p.att + ((size_t)n * K + k) * R * c;
    const float inv_hw = 1.f / (float)(p.h * p.w);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += att[r * c + ch];     // raw sums were accumulated here
        s *= inv_hw;
        g[ch] = s;
        p.gap[((size_t)n * K + k) * c + ch] = s;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        float q = p.b1[k * c2 + j];
        const float* w1 = STAGE ? sw1 + j : p.w1 + (size_t)k * c * c2 + j;
#pragma unroll 8
        for (int ch = 0; ch < c; ++ch) q = fmaf(g[ch], w1[(size_t)ch * c2], q);
        const float sc = p.gamma[k * c2 + j] * rsqrtf(p.var[k * c2 + j] + p.bn_eps);
        q = (q - p.mean[k * c2 + j]) * sc + p.beta[k * c2 + j];
        q = act_apply(p.act, q);
        h1[j] = q;
        p.h1[((size_t)n * K + k) * c2 + j] = q;
    }
    __syncthreads();
    for (int r = 0; r < R; ++r) {
        const float* w2 = STAGE ? sw2 + (size_t)r * c2 * c : p.w2 + ((size_t)k * R + r) * c2 * c;
        const float* b2 = p.b2 + ((size_t)k * R + r) * c;
        // each thread owns channels ch = tid, tid+bd, ... ; z kept in att then normalised
        float lmax = -INFINITY;
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            float z = b2[ch];
#pragma unroll 8
            for (int j = 0; j < c2; ++j) z = fmaf(h1[j], w2[(size_t)j * c + ch], z);
            att[r * c + ch] = z;
            lmax = fmaxf(lmax, z);
        }
        if (R == 1) {
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) att[ch] = 1.f / (1.f + expf(-att[ch]));
        } else {
            const float m = block_reduce(lmax, red, true);
            float lsum = 0.f;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) { float e = expf(att[r * c + ch] - m); att[r * c + ch] = e; lsum += e; }
            const float tot = block_reduce(lsum, red, false);
            const float inv = 1.f / tot;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) att[r * c + ch] *= inv;
        }
        __syncthreads();
    }
}

// v[n,p,k*c+cc] = sum_r u[n,p,(k*R+r)*c+cc] * att[n,k,r,cc]
template <typename T, int V>
__global__ void splitatt_combine_kernel(int N, int hw, int K, int R, int c, tbi_view u, tbi_view v, const float* att) {
    const int cv = (K * c) / V;
    const long long total = (long long)N * hw * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % cv) * V; const long long pg = i / cv;     // pg = n*hw + p
        const int n = (int)(pg / hw);
        const int k = co / c, cc = co % c;
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) o[q] = 0.f;
        const float* a = att + ((size_t)n * K + k) * R * c + cc;
        const T* up = (const T*)u.ptr + (size_t)pg * u.cstride + u.coff + (size_t)k * R * c + cc;
        for (int r = 0; r < R; ++r) {
            float x[V];
            ld_pack<T, V>(up + (size_t)r * c, x);
#pragma unroll
            for (int q = 0; q < V; ++q) o[q] = fmaf(x[q], __ldg(a + r * c + q), o[q]);
        }
        st_pack<T, V>((T*)v.ptr + (size_t)pg * v.cstride + v.coff + co, o);
    }
}

// du[n,p,(k*R+r)*c+cc] = (dv[n,p,k*c+cc]*att[n,k,r,cc] + dgap[n,k,cc]/hw) * act'(u[...])
template <typename T, int V>
__global__ void splitatt_du_kernel(int N, int hw, int K, int R, int c, int act, tbi_view u, tbi_view dv, tbi_view du,
                                   const float* att, const float* dgap) {
    const int cv = (K * c) / V;
    const long long total = (long long)N * hw * cv;
    const float inv_hw = 1.f / (float)hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % cv) * V; const long long pg = i / cv;
        const int n = (int)(pg / hw);
        const int k = co / c, cc = co % c;
        float g[V];
        ld_pack<T, V>((const T*)dv.ptr + (size_t)pg * dv.cstride + dv.coff + co, g);
        const float* a = att + ((size_t)n * K + k) * R * c + cc;
        const float* dg = dgap + ((size_t)n * K + k) * c + cc;
        const size_t uo = (size_t)k * R * c + cc;
        for (int r = 0; r < R; ++r) {
            float x[V], o[V];
            ld_pack<T, V>((const T*)u.ptr + (size_t)pg * u.cstride + u.coff + uo + (size_t)r * c, x);
#pragma unroll
            for (int q = 0; q < V; ++q)
                o[q] = (g[q] * __ldg(a + r * c + q) + __ldg(dg + q) * inv_hw) * act_grad_from_out(act, x[q]);
            st_pack<T, V>((T*)du.ptr + (size_t)pg * du.cstride + du.coff + uo + (size_t)r * c, o);
        }
    }
}

// per (n,k): da (in scratch dz region) -> dz ; dbn, xhat, dgap
// scratch layout: dz [n][K][R][c] | dgap [n][K][c] | dbn [n][K][c2] | xhat [n][K][c2]
__global__ void splitatt_fc_bwd_kernel(tbi_splitatt p, float* scratch) {
    extern __shared__ float sm[];
    const int c = p.c, c2 = p.c / 2, R = p.radix, K = p.kpaths, N = p.n;
    float* dq = sm + c2;           // [c2]  (sm[0..c2) unused)
    float* red = dq + c2;          // [32]
    const int n = blockIdx.x, k = blockIdx.y;
    float* dz = scratch + ((size_t)n * K + k) * R * c;
    float* dgap = scratch + (size_t)N * K * R * c + ((size_t)n * K + k) * c;
    float* dbn = scratch + (size_t)N * K * R * c + (size_t)N * K * c + ((size_t)n * K + k) * c2;
    float* xhat = dbn + (size_t)N * K * c2;
    const float* att = p.att + ((size_t)n * K + k) * R * c;
    const float* g = p.gap + ((size_t)n * K + k) * c;
    const float* h1 = p.h1 + ((size_t)n * K + k) * c2;
    for (int r = 0; r < R; ++r) {
        if (R == 1) {
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) { const float a = att[ch]; dz[ch] = a * (1.f - a) * dz[ch]; }
        } else {
            float l = 0.f;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) l += att[r * c + ch] * dz[r * c + ch];
            const float dot = block_reduce(l, red, false);
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) dz[r * c + ch] = att[r * c + ch] * (dz[r * c + ch] - dot);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int j = wid; j < c2; j += nwarp) {                  // a warp per output: lanes stride the contiguous channel axis
        float s = 0.f;
        for (int r = 0; r < R; ++r) {
            const float* w2 = p.w2 + (((size_t)k * R + r) * c2 + j) * c;
            for (int ch = lane; ch < c; ch += 32) s = fmaf(dz[r * c + ch], __ldg(w2 + ch), s);
        }
        s = warp_sum(s);
        if (lane == 0) dq[j] = s;                            // parked in dq, consumed just below
    }
    __syncthreads();
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        const float s = dq[j];
        // recompute pre-BN q for xhat
        float q = p.b1[k * c2 + j];
        const float* w1 = p.w1 + (size_t)k * c * c2 + j;
#pragma unroll 8
        for (int ch = 0; ch < c; ++ch) q = fmaf(g[ch], __ldg(w1 + (size_t)ch * c2), q);
        const float istd = rsqrtf(p.var[k * c2 + j] + p.bn_eps);
        const float xh = (q - p.mean[k * c2 + j]) * istd;
        const float d = s * act_grad_from_out(p.act, h1[j]);
        dbn[j] = d;
        xhat[j] = xh;
        dq[j] = d * p.gamma[k * c2 + j] * istd;
    }
    __syncthreads();
    for (int ch = wid; ch < c; ch += nwarp) {
        float s = 0.f;
        const float* w1 = p.w1 + ((size_t)k * c + ch) * c2;
        for (int j = lane; j < c2; j += 32) s = fmaf(dq[j], __ldg(w1 + j), s);
        s = warp_sum(s);
        if (lane == 0) dgap[ch] = s;
    }
}

// parameter gradients: reduce over n.  grid.y selects the tensor; threads over its elements.
__global__ void splitatt_param_grad_kernel(tbi_splitatt p, const float* scratch, float* dw1, float* db1, float* dgamma,
                                           float* dbeta, float* dw2, float* db2) {
    const int c = p.c, c2 = p.c / 2, R = p.radix, K = p.kpaths, N = p.n;
    const float* dz = scratch;
    const float* dbn = scratch + (size_t)N * K * R * c + (size_t)N * K * c;
    const float* xhat = dbn + (size_t)N * K * c2;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (blockIdx.y == 0) {                                   // dw2 [K][R][c2][c]
        if (i >= (long long)K * R * c2 * c) return;
        const int ch = (int)(i % c); long long t = i / c; const int j = (int)(t % c2); t /= c2; const int r = (int)(t % R); const int k = (int)(t / R);
        float s = 0.f;
#pragma unroll 8
        for (int n = 0; n < N; ++n) s = fmaf(p.h1[((size_t)n * K + k) * c2 + j], dz[(((size_t)n * K + k) * R + r) * c + ch], s);
        dw2[i] += s;
    } else if (blockIdx.y == 1) {                            // db2 [K][R][c]
        if (i >= (long long)K * R * c) return;
        const int ch = (int)(i % c); long long t = i / c; const int r = (int)(t % R); const int k = (int)(t / R);
        float s = 0.f;
#pragma unroll 8
        for (int n = 0; n < N; ++n) s += dz[(((size_t)n * K + k) * R + r) * c + ch];
        db2[i] += s;
    } else if (blockIdx.y == 2) {                            // dw1 [K][c][c2]
        if (i >= (long long)K * c * c2) return;
        const int j = (int)(i % c2); long long t = i / c2; const int ch = (int)(t % c); const int k = (int)(t / c);
        const float sc = p.gamma[k * c2 + j] * rsqrtf(p.var[k * c2 + j] + p.bn_eps);
        float s = 0.f;
#pragma unroll 8
        for (int n = 0; n < N; ++n) s = fmaf(p.gap[((size_t)n * K + k) * c + ch], dbn[((size_t)n * K + k) * c2 + j], s);
        dw1[i] += s * sc;
    } else {                                                 // db1, dgamma, dbeta [K][c2]
        if (i >= (long long)K * c2) return;
        const int j = (int)(i % c2); const int k = (int)(i / c2);
        const float sc = p.gamma[k * c2 + j] * rsqrtf(p.var[k * c2 + j] + p.bn_eps);
        float sb = 0.f, sg = 0.f;
#pragma unroll 8
        for (int n = 0; n < N; ++n) { const float d = dbn[((size_t)n * K + k) * c2 + j]; sb += d; sg = fmaf(d, xhat[((size_t)n * K + k) * c2 + j], sg); }
        dbeta[i] += sb; dgamma[i] += sg; db1[i] += sb * sc;
    }
}

// ---------------------------------------------------------------------------------------------
// softmax + my_loss_cat + accuracy (+ gradient w.r.t. logits)
// ---------------------------------------------------------------------------------------------
// The loss couples the batch through the per-pixel class counts, so a pixel is owned by one CTA -- but by SL_J threads of it:
// thread (pixel lane, j) takes the images n = j, j + SL_J, ...; the counts and the pixel's loss are reduced over j through shared
// memory in a fixed order (deterministic).  A warp is 32 consecutive pixels of one image: 384-byte contiguous accesses.
// (One thread per pixel looping over the whole batch left only 65 536 threads for 285 MB of traffic: 118 us.)
constexpr int SL_J = 8;
// FROM_Y: the logits are not read but FORMED here from the head's per-input-pixel tap products Y[n, H/2, W/2, 16 taps x NC]
// (fp32; tbi_convt_scatter_y's sum, same order: bias, then the 2 x 2 contributing (ky, kx) -- bit-identical logits), so the
// [N,H,W,NC] logits tensor is never written or re-read and the scatter launch disappears.  W = width of the OUTPUT grid.
template <int NC, typename TD, bool FROM_Y>
__global__ void __launch_bounds__(32 * SL_J) softmax_loss_kernel(int N, int hw, const float* __restrict__ logits, const float* __restrict__ y,
                                                                float* __restrict__ probs, float* __restrict__ loss_map, int32_t* correct,
                                                                TD* __restrict__ dlogits, int dl_cs,
                                                                const float* __restrict__ ytap, int y_cs, const float* __restrict__ bias, int W) {
    __shared__ float red[SL_J][32][3];
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const bool live = p < hw;
    float cnt[3] = {0.f, 0.f, 0.f};
    if (live) {
        for (int n = j; n < N; n += SL_J) {
            const float* yy = y + ((size_t)n * hw + p) * NC;
#pragma unroll
            for (int c = 0; c < 3; ++c) cnt[c] += yy[c];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) red[j][lane][c] = cnt[c];
    __syncthreads();
    float sf[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < SL_J; ++q) t += red[q][lane][c];
        sf[c] = (1.f / (t + 1.f)) / (float)hw;
    }
    __syncthreads();
    int ok = 0;
    float ce = 0.f;
    if (live) {
        for (int n = j; n < N; n += SL_J) {
            const size_t o = ((size_t)n * hw + p) * NC;
            float z[NC], yy[NC];
            float m = -INFINITY;
#pragma unroll
            if (FROM_Y) {
                const int H = hw / W, oy = p / W, ox = p - oy * W, h2 = H >> 1, w2 = W >> 1;
#pragma unroll
                for (int c = 0; c < NC; ++c) z[c] = bias ? bias[c] : 0.f;
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int ky = ((oy + 1) & 1) + 2 * a, iy = (oy + 1 - ky) >> 1;
                    if (iy < 0 || iy >= h2) continue;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kx = ((ox + 1) & 1) + 2 * e, ix = (ox + 1 - kx) >> 1;
                        if (ix < 0 || ix >= w2) continue;
                        const float* src = ytap + (((size_t)n * h2 + iy) * w2 + ix) * y_cs + (ky * 4 + kx) * NC;
#pragma unroll
                        for (int c = 0; c < NC; ++c) z[c] += src[c];
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) { if (!FROM_Y) z[c] = logits[o + c]; yy[c] = y[o + c]; m = fmaxf(m, z[c]); }
            float sm = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) { z[c] = expf(z[c] - m); sm += z[c]; }
            const float inv = 1.f / sm;
            int am = 0, ay = 0;
#pragma unroll
            for (int c = 0; c < NC; ++c) { z[c] *= inv; if (z[c] > z[am]) am = c; if (yy[c] > yy[ay]) ay = c; }
            ok += (am == ay);
            float dldp[NC];
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                dldp[c] = 0.f;
                if (c < 3) {
                    ce += yy[c] * logf(z[c] + 1e-7f) * sf[c];
                    dldp[c] = -yy[c] * sf[c] / (z[c] + 1e-7f);
                }
                dot += z[c] * dldp[c];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                probs[o + c] = z[c];
                if (dlogits) stf(dlogits + ((size_t)n * hw + p) * dl_cs + c, z[c] * (dldp[c] - dot));
            }
        }
    }
    red[j][lane][0] = ce;
    __syncthreads();
    if (j == 0 && live) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < SL_J; ++q) t += red[q][lane][0];
        loss_map[p] = -t;
    }
    ok = (int)warp_sum((float)ok);
    if (lane == 0 && ok) atomicAdd(correct, ok);
}

// ---------------------------------------------------------------------------------------------
// packing, BN fold / param grads, Adam, misc
// ---------------------------------------------------------------------------------------------
// block-diagonal dense expansion of a grouped kernel (tbi_conv_dense_expand): same two layouts as below for groups = 1 over
// cin_total = groups*cin_g input channels, zero where input and output channel belong to different groups
template <typename T>
__global__ void pack_conv_expand_kernel(int mode, int ntaps, int groups, int cin_g, int cin, int cout_total, const float* __restrict__ w,
                                        const float* __restrict__ scale, T* __restrict__ out) {
    const int cout_g = cout_total / groups;                  // cin = groups*cin_g rounded up to 16: rows past the last group are zero
    const long long total = (long long)ntaps * cin * cout_total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int tap, ci, co;
        if (mode == 0) {            // out[co][tap][ci]
            ci = (int)(i % cin); long long t = i / cin; tap = (int)(t % ntaps); co = (int)(t / ntaps);
        } else {                    // out[ci][tap'][co]
            co = (int)(i % cout_total); long long t = i / cout_total; const int tp = (int)(t % ntaps); ci = (int)(t / ntaps);
            tap = ntaps - 1 - tp;
        }
        float v = 0.f;
        if (ci < groups * cin_g && ci / cin_g == co / cout_g) {
            v = w[((size_t)tap * cin_g + ci % cin_g) * cout_total + co];
            if (scale) v *= scale[co];
        }
        stf(out + i, v);
    }
}

// dw_hwio[tap][ci_g][co] += dense[tap][g(co)*cin_g + ci_g][co]   (the block-diagonal part of a dense weight gradient)
__global__ void wgrad_gather_blocks_kernel(int ntaps, int groups, int cin_g, int cin, int cout_total, const float* __restrict__ dense, float* dw) {
    const int cout_g = cout_total / groups;
    const long long total = (long long)ntaps * cin_g * cout_total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % cout_total); long long t = i / cout_total; const int cig = (int)(t % cin_g); const int tap = (int)(t / cin_g);
        dw[i] += dense[((size_t)tap * cin + (size_t)(co / cout_g) * cin_g + cig) * cout_total + co];
    }
}

template <typename T>
__global__ void pack_conv_kernel(int mode, int ntaps, int groups, int cin_g, int cout_total, const float* __restrict__ w,
                                 const float* __restrict__ scale, T* __restrict__ out) {
    const int cout_g = cout_total / groups;
    const long long total = (long long)ntaps * cin_g * cout_total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int tap, ci, co;
        if (mode == 0) {            // out[co][tap][ci]
            ci = (int)(i % cin_g); long long t = i / cin_g; tap = (int)(t % ntaps); co = (int)(t / ntaps);
        } else {                    // out[g*cin_g+ci][tap'][col] ; K = ntaps*cout_g
            const int col = (int)(i % cout_g); long long t = i / cout_g; const int tp = (int)(t % ntaps); const int row = (int)(t / ntaps);
            const int g = row / cin_g; ci = row % cin_g; tap = ntaps - 1 - tp; co = g * cout_g + col;
        }
        float v = w[((size_t)tap * cin_g + ci) * cout_total + co];
        if (scale) v *= scale[co];
        stf(out + i, v);
    }
}

struct ConvtTapTable { int n[4]; int off[4]; int ky[4][4]; int kx[4][4]; };

template <typename T>
__global__ void pack_convt_kernel(int mode, int k, int cin, int cout, int cpad, ConvtTapTable tt, const float* __restrict__ w,
                                  const float* __restrict__ scale, T* __restrict__ out) {
    const long long total = mode >= 2 ? (long long)cin * cpad : (long long)k * k * cin * (mode == 1 ? cpad : cout);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ky, kx, ci, co;
        if (mode == 3) {            // out[q][ci]: the transposed-conv forward as a plain GEMM per INPUT pixel (K = cin, N = q)
            ci = (int)(i % cin); const int q = (int)(i / cin);
            if (q >= k * k * cout) { stf(out + i, 0.f); continue; }
            co = q % cout; const int tap = q / cout; ky = tap / k; kx = tap % k;
        } else
        if (mode == 2) {            // out[ci][q], q = (ky*k+kx)*cout + co, zero-padded to cpad
            const int q = (int)(i % cpad); ci = (int)(i / cpad);
            if (q >= k * k * cout) { stf(out + i, 0.f); continue; }
            co = q % cout; const int tap = q / cout; ky = tap / k; kx = tap % k;
        } else
        if (mode == 0) {            // out[phase][co][t][ci], phase blocks at tt.off[ph]*cout*cin
            long long r = i; int ph = 0;
            while (ph < 3 && r >= (long long)tt.n[ph] * cout * cin) { r -= (long long)tt.n[ph] * cout * cin; ++ph; }
            ci = (int)(r % cin); long long t = r / cin; const int tp = (int)(t % tt.n[ph]); co = (int)(t / tt.n[ph]);
            ky = tt.ky[ph][tp]; kx = tt.kx[ph][tp];
        } else {                    // out[ci][ky*k+kx][co]
            co = (int)(i % cpad); long long t = i / cpad; const int tap = (int)(t % (k * k)); ci = (int)(t / (k * k));
            ky = tap / k; kx = tap % k;
            if (co >= cout) { stf(out + i, 0.f); continue; }
        }
        float v = w[(((size_t)ky * k + kx) * cout + co) * cin + ci];
        if (scale) v *= scale[co];
        stf(out + i, v);
    }
}

__global__ void bn_fold_kernel(int c, const float* gamma, const float* beta, const float* mean, const float* var,
                               const float* bias, float eps, float* scale, float* fbias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float b = bias ? bias[i] : 0.f;
    if (gamma) {
        const float s = gamma[i] * rsqrtf(var[i] + eps);
        scale[i] = s; fbias[i] = (b - mean[i]) * s + beta[i];
    } else { scale[i] = 1.f; fbias[i] = b; }
}

// one block per channel
__global__ void bn_param_grad_kernel(int c, long long k_outer, long long inner, long long outer_stride, long long co_stride,
                                     const float* w, float* dw, const float* bias, float* dbias, const float* gamma,
                                     const float* mean, const float* var, float eps, float* dgamma, float* dbeta) {
    __shared__ float red[32];
    const int ch = blockIdx.x;
    const long long K = k_outer * inner;
    const float istd = rsqrtf(var[ch] + eps);
    const float sc = gamma[ch] * istd;
    float l = 0.f;
    if (inner % 4 == 0 && outer_stride % 4 == 0 && co_stride % 4 == 0 && (((uintptr_t)w | (uintptr_t)dw) & 15) == 0) {
        // HWOI kernels: runs of `inner` contiguous floats per (tap, channel) -> 16-byte accesses, 4 independent chains
        const long long K4 = K / 4, inner4 = inner / 4;
#pragma unroll 4
        for (long long k = threadIdx.x; k < K4; k += blockDim.x) {
            const size_t a = (size_t)(k / inner4) * outer_stride + (size_t)ch * co_stride + (size_t)(k % inner4) * 4;
            float4 g = *reinterpret_cast<const float4*>(dw + a);
            const float4 ww = *reinterpret_cast<const float4*>(w + a);
            l = fmaf(ww.x, g.x, l); l = fmaf(ww.y, g.y, l); l = fmaf(ww.z, g.z, l); l = fmaf(ww.w, g.w, l);
            g.x *= sc; g.y *= sc; g.z *= sc; g.w *= sc;
            *reinterpret_cast<float4*>(dw + a) = g;
        }
    } else {
        for (long long k = threadIdx.x; k < K; k += blockDim.x) {
            const size_t a = (size_t)(k / inner) * outer_stride + (size_t)ch * co_stride + (size_t)(k % inner);
            const float g = dw[a];
            l = fmaf(w[a], g, l);
            dw[a] = g * sc;
        }
    }
    const float dot = block_reduce(l, red, false);
    if (threadIdx.x == 0) {
        const float draw = dbias[ch];
        const float b = bias ? bias[ch] : 0.f;
        dgamma[ch] += (dot + (b - mean[ch]) * draw) * istd;
        dbeta[ch] += draw;
        dbias[ch] = draw * sc;
    }
}

// HWIO kernels (output channel innermost): 32 consecutive channels per warp row so every access is a full 128-byte line,
// K split over blockIdx.y; dgamma is linear in the dot product, so the slices combine with one atomic each.
__global__ void __launch_bounds__(256) bn_param_grad_cols_kernel(int c, long long K, long long row_stride, const float* __restrict__ w, float* dw,
                                                                 const float* bias, float* dbias, const float* gamma, const float* mean,
                                                                 const float* var, float eps, float* dgamma, float* dbeta) {
    __shared__ float red[8][33];
    const int ch = blockIdx.x * 32 + threadIdx.x, kl = threadIdx.y;
    const long long per = (K + gridDim.y - 1) / gridDim.y;
    const long long k0 = (long long)blockIdx.y * per, k1 = min(K, k0 + per);
    float l = 0.f, istd = 0.f, sc = 0.f;
    if (ch < c) {
        istd = rsqrtf(var[ch] + eps);
        sc = gamma[ch] * istd;
#pragma unroll 4
        for (long long k = k0 + kl; k < k1; k += 8) {
            const size_t a = (size_t)k * row_stride + ch;
            const float g = dw[a];
            l = fmaf(w[a], g, l);
            dw[a] = g * sc;
        }
    }
    red[kl][threadIdx.x] = l;
    __syncthreads();
    if (kl == 0 && ch < c) {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) dot += red[q][threadIdx.x];
        if (blockIdx.y == 0) {
            const float draw = dbias[ch];
            const float b = bias ? bias[ch] : 0.f;
            dot += (b - mean[ch]) * draw;
            dbeta[ch] += draw;
            dbias[ch] = draw * sc;
        }
        atomicAdd(dgamma + ch, dot * istd);
    }
}

__global__ void adam_kernel(long long count, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const int32_t* step_count, float lr, float b1, float b2, float eps, float gs) {
    const float t = (float)(*step_count + 1);
    const float lr_t = lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
    const long long n4 = count / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            ma[k] = b1 * ma[k] + (1.f - b1) * gk;
            va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
            pa[k] -= lr_t * ma[k] / (sqrtf(va[k]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const float gk = g[i] * gs;
        const float mk = b1 * m[i] + (1.f - b1) * gk, vk = b2 * v[i] + (1.f - b2) * gk * gk;
        m[i] = mk; v[i] = vk; p[i] -= lr_t * mk / (sqrtf(vk) + eps);
    }
}
__global__ void adam_advance_kernel(int32_t* s) { if (threadIdx.x == 0 && blockIdx.x == 0) *s += 1; }

// Adam with its hyper-parameters on the device (hyper = {lr, grad_scale, clip_norm}): a captured CUDA graph keeps following
// a learning-rate schedule, and the global-norm clip of VisionTransformer.py:244 (tf.clip_by_global_norm) is one more factor
// on the gradient: g * clip / max(||g||, clip) with ||g|| = sqrt(*gnorm_sq) * grad_scale.
__global__ void adam_dev_kernel(long long count, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, const int32_t* step_count, const float* __restrict__ hyper,
                                const float* __restrict__ gnorm_sq, float b1, float b2, float eps) {
    const float t = (float)(*step_count + 1);
    const float lr_t = hyper[0] * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
    float gs = hyper[1];
    const float clip = hyper[2];
    if (clip > 0.f && gnorm_sq) { const float nrm = sqrtf(*gnorm_sq) * fabsf(gs); gs *= clip / fmaxf(nrm, clip); }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count / 4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            ma[k] = b1 * ma[k] + (1.f - b1) * gk;
            va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
            pa[k] -= lr_t * ma[k] / (sqrtf(va[k]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (count / 4) * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const float gk = g[i] * gs;
        const float mk = b1 * m[i] + (1.f - b1) * gk, vk = b2 * v[i] + (1.f - b2) * gk * gk;
        m[i] = mk; v[i] = vk; p[i] -= lr_t * mk / (sqrtf(vk) + eps);
    }
}

// out += sum x^2 (global gradient norm): 16-byte loads, warp shuffles, one atomic per block
__global__ void __launch_bounds__(256) sumsq_kernel(long long count, const float* __restrict__ x, float* out) {
    __shared__ float red[32];
    float l = 0.f;
    const long long n4 = count / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 q = reinterpret_cast<const float4*>(x)[i];
        l = fmaf(q.x, q.x, l); l = fmaf(q.y, q.y, l); l = fmaf(q.z, q.z, l); l = fmaf(q.w, q.w, l);
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) l = fmaf(x[i], x[i], l);
    const float tot = block_reduce(l, red, false);
    if (threadIdx.x == 0) atomicAdd(out, tot);
}

__global__ void dropout_mask_kernel(uint8_t* keep, long long count, unsigned long long seed, const int32_t* step_ptr) {
    const unsigned long long offset = step_ptr ? (unsigned long long)(*step_ptr) * (unsigned long long)count : 0ull;
    auto draw = [&](long long i) -> unsigned long long {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + offset + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        return ((z >> 33) & 1ull) << 1;                      // multiplier: 0 dropped, 2 kept
    };
    const bool vec = (reinterpret_cast<uintptr_t>(keep) & 7) == 0;
    const long long groups = vec ? count / 8 : 0;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < groups; gi += (long long)gridDim.x * blockDim.x) {
        unsigned long long pack = 0;                          // same per-element stream as the scalar form, one 8-byte store
#pragma unroll
        for (int j = 0; j < 8; ++j) pack |= draw(gi * 8 + j) << (8 * j);
        reinterpret_cast<unsigned long long*>(keep)[gi] = pack;
    }
    for (long long i = groups * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        keep[i] = (uint8_t)draw(i);
}

template <typename S, typename D>
__global__ void cast_kernel(long long count, const S* __restrict__ src, D* __restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        stf(dst + i, (float)src[i]);
}

}  // namespace

#define DISPATCH_TV(dtype, V, CALL_F32_V, CALL_F32_1, CALL_BF_V, CALL_BF_1)                     \
    do {                                                                                        \
        if ((dtype) == TBI_F32) { if ((V) > 1) { CALL_F32_V; } else { CALL_F32_1; } }           \
        else if ((dtype) == TBI_BF16) { if ((V) > 1) { CALL_BF_V; } else { CALL_BF_1; } }       \
        else return tbi_set_error(TBI_ERR_UNSUPPORTED, "dtype %d", (dtype));                    \
    } while (0)

extern "C" int tbi_avgpool2x2_fwd(int dtype, int n, int h, int w, const tbi_view* x, const tbi_view* y, void* stream) {
    TBI_CHECK(h % 2 == 0 && w % 2 == 0 && x->c == y->c && x->h == h && x->w == w && y->h == h / 2 && y->w == w / 2,
              TBI_ERR_BAD_SHAPE, "avgpool fwd: bad shapes");
    cudaStream_t s = (cudaStream_t)stream;
    const int V = pick_vec(dtype, {x, y});
    const long long work = (long long)n * (h / 2) * (w / 2) * (y->c / V);
    const unsigned g = grid_for(work, 256);
    DISPATCH_TV(dtype, V,
        (avgpool_fwd_kernel<float, 4><<<g, 256, 0, s>>>(n, h / 2, w / 2, *x, *y)),
        (avgpool_fwd_kernel<float, 1><<<g, 256, 0, s>>>(n, h / 2, w / 2, *x, *y)),
        (avgpool_fwd_kernel<__nv_bfloat16, 8><<<g, 256, 0, s>>>(n, h / 2, w / 2, *x, *y)),
        (avgpool_fwd_kernel<__nv_bfloat16, 1><<<g, 256, 0, s>>>(n, h / 2, w / 2, *x, *y)));
    TBI_CUDA_LAUNCH_CHECK("avgpool_fwd");
    return TBI_OK;
}

extern "C" int tbi_avgpool2x2_bwd(int dtype, int n, int h, int w, const tbi_view* dy, const tbi_view* dx, int accumulate,
                                  int dact, const tbi_view* dact_ref, void* stream) {
    TBI_CHECK(h % 2 == 0 && w % 2 == 0 && dx->c == dy->c && dx->h == h && dx->w == w && dy->h == h / 2 && dy->w == w / 2,
              TBI_ERR_BAD_SHAPE, "avgpool bwd: bad shapes");
    TBI_CHECK(dact == TBI_ACT_NONE || (dact_ref && dact_ref->ptr), TBI_ERR_BAD_SHAPE, "avgpool bwd: dact without ref");
    cudaStream_t s = (cudaStream_t)stream;
    tbi_view ref{}; if (dact != TBI_ACT_NONE) ref = *dact_ref;
    const int V = pick_vec(dtype, {dy, dx, dact != TBI_ACT_NONE ? dact_ref : nullptr});
    const long long work = (long long)n * (h / 2) * (w / 2) * (dy->c / V);
    const unsigned g = grid_for(work, 256);
    DISPATCH_TV(dtype, V,
        (avgpool_bwd_kernel<float, 4><<<g, 256, 0, s>>>(n, h / 2, w / 2, *dy, *dx, accumulate, dact, ref)),
        (avgpool_bwd_kernel<float, 1><<<g, 256, 0, s>>>(n, h / 2, w / 2, *dy, *dx, accumulate, dact, ref)),
        (avgpool_bwd_kernel<__nv_bfloat16, 8><<<g, 256, 0, s>>>(n, h / 2, w / 2, *dy, *dx, accumulate, dact, ref)),
        (avgpool_bwd_kernel<__nv_bfloat16, 1><<<g, 256, 0, s>>>(n, h / 2, w / 2, *dy, *dx, accumulate, dact, ref)));
    TBI_CUDA_LAUNCH_CHECK("avgpool_bwd");
    return TBI_OK;
}

extern "C" int tbi_act_bwd(int dtype, int64_t npix, int act, const tbi_view* dy, const tbi_view* y_ref, const uint8_t* keep,
                           const tbi_view* dz, void* stream) {
    TBI_CHECK(dy->c == dz->c && y_ref->c == dz->c, TBI_ERR_BAD_SHAPE, "act_bwd: channel mismatch");
    cudaStream_t s = (cudaStream_t)stream;
    const int V = pick_vec(dtype, {dy, y_ref, dz});
    const unsigned g = grid_for(npix * (dz->c / V), 256);
    DISPATCH_TV(dtype, V,
        (act_bwd_kernel<float, 4><<<g, 256, 0, s>>>(npix, act, *dy, *y_ref, keep, *dz)),
        (act_bwd_kernel<float, 1><<<g, 256, 0, s>>>(npix, act, *dy, *y_ref, keep, *dz)),
        (act_bwd_kernel<__nv_bfloat16, 8><<<g, 256, 0, s>>>(npix, act, *dy, *y_ref, keep, *dz)),
        (act_bwd_kernel<__nv_bfloat16, 1><<<g, 256, 0, s>>>(npix, act, *dy, *y_ref, keep, *dz)));
    TBI_CUDA_LAUNCH_CHECK("act_bwd");
    return TBI_OK;
}

extern "C" int tbi_convt_gather_dz(int dtype, int n, int h, int w, int ksize, int cout, const tbi_view* dz, const tbi_view* g, void* stream) {
    TBI_CHECK(ksize == 3 || ksize == 4, TBI_ERR_UNSUPPORTED, "convt_gather: ksize %d", ksize);
    TBI_CHECK(dz->h == 2 * h && dz->w == 2 * w && dz->c >= cout && g->h == h && g->w == w && g->c >= ksize * ksize * cout, TBI_ERR_BAD_SHAPE,
              "convt_gather: shapes");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned gr = grid_for((long long)n * h * w * ksize * ksize, 256);
    const int pad = ksize == 4 ? 1 : 0;
    if (dtype == TBI_BF16 && ksize == 4 && cout <= 4 && dz->c >= 4 && dz->cstride % 4 == 0 && dz->coff % 4 == 0 && ((uintptr_t)dz->ptr & 7) == 0 &&
        g->c == 64 && g->cstride == 64 && g->coff == 0 && ((uintptr_t)g->ptr & 31) == 0) {
        const unsigned gg = grid_for((long long)n * h * w, 256);
        switch (cout) {
            case 1: convt_gather4_bf16_kernel<1><<<gg, 256, 0, s>>>(n, h, w, *dz, *g); break;
            case 2: convt_gather4_bf16_kernel<2><<<gg, 256, 0, s>>>(n, h, w, *dz, *g); break;
            case 3: convt_gather4_bf16_kernel<3><<<gg, 256, 0, s>>>(n, h, w, *dz, *g); break;
            default: convt_gather4_bf16_kernel<4><<<gg, 256, 0, s>>>(n, h, w, *dz, *g); break;
        }
        TBI_CUDA_LAUNCH_CHECK("convt_gather4");
        return TBI_OK;
    }
    if (dtype == TBI_F32) convt_gather_kernel<float><<<gr, 256, 0, s>>>(n, h, w, ksize, pad, cout, *dz, *g);
    else if (dtype == TBI_BF16) convt_gather_kernel<__nv_bfloat16><<<gr, 256, 0, s>>>(n, h, w, ksize, pad, cout, *dz, *g);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "convt_gather dtype");
    TBI_CUDA_LAUNCH_CHECK("convt_gather");
    return TBI_OK;
}

extern "C" int tbi_convt_scatter_y(int y_dtype, int n, int h, int w, int ksize, int cout, const tbi_view* y, const float* bias, const tbi_view* out,
                                   void* stream) {
    TBI_CHECK(ksize == 4 && cout >= 1 && cout <= 4, TBI_ERR_UNSUPPORTED, "convt_scatter: ksize %d, cout %d (k = 4, cout <= 4 only)", ksize, cout);
    TBI_CHECK(y && out && y->h == h && y->w == w && y->c >= 16 * cout && out->h == 2 * h && out->w == 2 * w && out->c >= cout, TBI_ERR_BAD_SHAPE,
              "convt_scatter: shapes");
    TBI_CHECK(y_dtype == TBI_F32 || y_dtype == TBI_BF16, TBI_ERR_UNSUPPORTED, "convt_scatter: y dtype");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for((long long)n * 4 * h * w, 256);
#define TBI_SCATTER(CO) do { if (y_dtype == TBI_F32) convt_scatter4_kernel<float, CO><<<g, 256, 0, s>>>(n, h, w, *y, bias, *out); \
                             else convt_scatter4_kernel<__nv_bfloat16, CO><<<g, 256, 0, s>>>(n, h, w, *y, bias, *out); } while (0)
    switch (cout) {
        case 1: TBI_SCATTER(1); break;
        case 2: TBI_SCATTER(2); break;
        case 3: TBI_SCATTER(3); break;
        default: TBI_SCATTER(4); break;
    }
#undef TBI_SCATTER
    TBI_CUDA_LAUNCH_CHECK("convt_scatter");
    return TBI_OK;
}

extern "C" int tbi_accumulate(int dtype, int64_t npix, const tbi_view* src, const tbi_view* dst, void* stream) {
    TBI_CHECK(src->c == dst->c, TBI_ERR_BAD_SHAPE, "accumulate: channel mismatch");
    cudaStream_t s = (cudaStream_t)stream;
    const int V = pick_vec(dtype, {src, dst});
    const unsigned g = grid_for(npix * (dst->c / V), 256);
    DISPATCH_TV(dtype, V,
        (accumulate_kernel<float, 4><<<g, 256, 0, s>>>(npix, *src, *dst)),
        (accumulate_kernel<float, 1><<<g, 256, 0, s>>>(npix, *src, *dst)),
        (accumulate_kernel<__nv_bfloat16, 8><<<g, 256, 0, s>>>(npix, *src, *dst)),
        (accumulate_kernel<__nv_bfloat16, 1><<<g, 256, 0, s>>>(npix, *src, *dst)));
    TBI_CUDA_LAUNCH_CHECK("accumulate");
    return TBI_OK;
}

extern "C" int tbi_colsum(int dtype, int64_t npix, const tbi_view* x, float* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    // enough blocks to fill the machine even for the small deep-stage tensors (>= 16 pixels per block)
    long long blocks = (npix + 15) / 16;
    const long long cap = (long long)tbi_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const int ppb = (int)((npix + blocks - 1) / blocks);
    blocks = (npix + ppb - 1) / ppb;
    {   // dense fast path
        const int V = dtype == TBI_F32 ? 4 : 8;
        static const bool no_flat = getenv("TBI_COLSUM_NO_FLAT") != nullptr;
        if (!no_flat && (dtype == TBI_F32 || dtype == TBI_BF16) && x->cstride == x->c && x->coff == 0 && x->c % V == 0 && 256 % (x->c / V) == 0 &&
            (((uintptr_t)x->ptr) & 15) == 0) {
            const int cv = x->c / V;
            const long long nchunks = npix * cv;
            long long nb = (nchunks + 256 * 8 - 1) / (256 * 8);
            const long long capf = (long long)tbi_sm_count() * 4;
            if (nb > capf) nb = capf;
            if (nb < 1) nb = 1;
            if (dtype == TBI_F32) colsum_flat_kernel<float, 4><<<(unsigned)nb, 256, 0, s>>>(nchunks, cv, (const float*)x->ptr, out);
            else colsum_flat_kernel<__nv_bfloat16, 8><<<(unsigned)nb, 256, 0, s>>>(nchunks, cv, (const __nv_bfloat16*)x->ptr, out);
            TBI_CUDA_LAUNCH_CHECK("colsum_flat");
            return TBI_OK;
        }
    }
    if (pick_vec(dtype, {x}) > 1) {
        if (dtype == TBI_F32) colsum_vec_kernel<float, 4><<<(unsigned)blocks, 256, 256 * 4 * sizeof(float), s>>>(npix, *x, out, ppb);
        else colsum_vec_kernel<__nv_bfloat16, 8><<<(unsigned)blocks, 256, 256 * 8 * sizeof(float), s>>>(npix, *x, out, ppb);
        TBI_CUDA_LAUNCH_CHECK("colsum_vec");
        return TBI_OK;
    }
    if (dtype == TBI_F32) colsum_kernel<float><<<(unsigned)blocks, 256, 256 * sizeof(float), s>>>(npix, *x, out, ppb);
    else if (dtype == TBI_BF16) colsum_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 256 * sizeof(float), s>>>(npix, *x, out, ppb);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "colsum dtype");
    TBI_CUDA_LAUNCH_CHECK("colsum");
    return TBI_OK;
}

static int splitatt_check(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v) {
    TBI_CHECK(p->radix >= 1 && p->radix <= 4 && p->kpaths >= 1 && p->c >= 2, TBI_ERR_BAD_SHAPE, "splitatt: radix/kpaths/c");
    TBI_CHECK(u->c == p->kpaths * p->radix * p->c, TBI_ERR_BAD_SHAPE, "splitatt: u channels %d != K*R*c", u->c);
    if (v) TBI_CHECK(v->c == p->kpaths * p->c, TBI_ERR_BAD_SHAPE, "splitatt: v channels %d != K*c", v->c);
    TBI_CHECK(p->act == TBI_ACT_ELU || p->act == TBI_ACT_LRELU, TBI_ERR_UNSUPPORTED, "splitatt: act");
    return TBI_OK;
}

template <bool MUL>
static int splitatt_reduce_launch(const tbi_splitatt* p, const tbi_view* u, const tbi_view* dv, float* raw, cudaStream_t s) {
    const int hw = p->h * p->w;
    const int C = u->c;
    // the vector must not straddle a (k,r) block: c % V == 0
    int V = pick_vec(p->dtype, {u, dv});
    if (p->c % V != 0) V = 1;
    cudaError_t e = cudaMemsetAsync(raw, 0, sizeof(float) * (size_t)p->n * C, s);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt memset: %s", cudaGetErrorString(e));
    // blocks per image so that the grid is a few waves of the SMs
    int bpi = (tbi_sm_count() * 4 + p->n - 1) / p->n;
    const int max_bpi = (hw + 63) / 64;
    if (bpi > max_bpi) bpi = max_bpi;
    if (bpi < 1) bpi = 1;
    const int ppb = (hw + bpi - 1) / bpi;
    bpi = (hw + ppb - 1) / ppb;
    dim3 grid(bpi, p->n);
    const size_t smem = 256 * sizeof(float) * V;
    tbi_view dvv{}; if (dv) dvv = *dv;
    if (p->dtype == TBI_F32) {
        if (V > 1) splitatt_reduce_kernel<float, 4, MUL><<<grid, 256, smem, s>>>(hw, p->radix, p->c, *u, dvv, raw, ppb);
        else       splitatt_reduce_kernel<float, 1, MUL><<<grid, 256, smem, s>>>(hw, p->radix, p->c, *u, dvv, raw, ppb);
    } else if (p->dtype == TBI_BF16) {
        if (V > 1) splitatt_reduce_kernel<__nv_bfloat16, 8, MUL><<<grid, 256, smem, s>>>(hw, p->radix, p->c, *u, dvv, raw, ppb);
        else       splitatt_reduce_kernel<__nv_bfloat16, 1, MUL><<<grid, 256, smem, s>>>(hw, p->radix, p->c, *u, dvv, raw, ppb);
    } else return tbi_set_error(TBI_ERR_UNSUPPORTED, "splitatt dtype");
    TBI_CUDA_LAUNCH_CHECK("splitatt_reduce");
    return TBI_OK;
}

extern "C" int tbi_splitatt_gap(const tbi_splitatt* p, const tbi_view* u, void* stream) {
    int rc = splitatt_check(p, u, nullptr); if (rc) return rc;
    return splitatt_reduce_launch<false>(p, u, nullptr, p->att, (cudaStream_t)stream);
}

extern "C" int tbi_splitatt_combine(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, void* stream) {
    int rc = splitatt_check(p, u, v); if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    int V = pick_vec(p->dtype, {u, v});
    if (p->c % V != 0) V = 1;
    const int hw = p->h * p->w;
    const unsigned g = grid_for((long long)p->n * hw * (v->c / V), 256);
    DISPATCH_TV(p->dtype, V,
        (splitatt_combine_kernel<float, 4><<<g, 256, 0, s>>>(p->n, hw, p->kpaths, p->radix, p->c, *u, *v, p->att)),
        (splitatt_combine_kernel<float, 1><<<g, 256, 0, s>>>(p->n, hw, p->kpaths, p->radix, p->c, *u, *v, p->att)),
        (splitatt_combine_kernel<__nv_bfloat16, 8><<<g, 256, 0, s>>>(p->n, hw, p->kpaths, p->radix, p->c, *u, *v, p->att)),
        (splitatt_combine_kernel<__nv_bfloat16, 1><<<g, 256, 0, s>>>(p->n, hw, p->kpaths, p->radix, p->c, *u, *v, p->att)));
    TBI_CUDA_LAUNCH_CHECK("splitatt_combine");
    return TBI_OK;
}

extern "C" int tbi_split_attention_fwd(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, void* stream) {
    int rc = splitatt_check(p, u, v); if (rc) return rc;
    rc = tbi_splitatt_fwd_fused(p, u, v, (cudaStream_t)stream);          // one cooperative launch when it applies (bf16, aligned)
    if (rc != 0) return rc < 0 ? rc : TBI_OK;
    rc = tbi_splitatt_gap(p, u, stream); if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t base_smem = sizeof(float) * (p->c + p->c / 2 + 32);
    const size_t w_smem = sizeof(float) * ((size_t)p->c * (p->c / 2) * (1 + p->radix));
    if (base_smem + w_smem <= 100 * 1024) {
        static bool attr_done = false;
        if (!attr_done) { cudaFuncSetAttribute(splitatt_fc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr_done = true; }
        splitatt_fc_kernel<true><<<dim3(p->n, p->kpaths), 256, base_smem + w_smem, s>>>(*p);
    } else {
        splitatt_fc_kernel<false><<<dim3(p->n, p->kpaths), 256, base_smem, s>>>(*p);
    }
    TBI_CUDA_LAUNCH_CHECK("splitatt_fc");
    return tbi_splitatt_combine(p, u, v, stream);
}

extern "C" int tbi_split_attention_bwd(const tbi_splitatt* p, const tbi_view* u, const tbi_view* dv, const tbi_view* du,
                                       float* dw1, float* db1, float* dgamma, float* dbeta, float* dw2, float* db2,
                                       float* scratch, void* stream) {
    int rc = splitatt_check(p, u, dv); if (rc) return rc;
    TBI_CHECK(du->c == u->c, TBI_ERR_BAD_SHAPE, "splitatt bwd: du channels");
    cudaStream_t s = (cudaStream_t)stream;
    const int c = p->c, c2 = p->c / 2, K = p->kpaths, R = p->radix, N = p->n;
    const long long maxel = (long long)K * R * c2 * c;
    rc = tbi_splitatt_bwd_fused(p, u, dv, du, scratch, s);               // da + FC backward + dU in one cooperative launch
    if (rc < 0) return rc;
    if (rc == 1) {
        splitatt_param_grad_kernel<<<dim3((unsigned)((maxel + 127) / 128), 4), 128, 0, s>>>(*p, scratch, dw1, db1, dgamma, dbeta, dw2, db2);
        TBI_CUDA_LAUNCH_CHECK("splitatt_param_grad");
        return TBI_OK;
    }
    rc = splitatt_reduce_launch<true>(p, u, dv, scratch, s); if (rc) return rc;
    splitatt_fc_bwd_kernel<<<dim3(N, K), 256, sizeof(float) * (2 * c2 + 32), s>>>(*p, scratch);
    TBI_CUDA_LAUNCH_CHECK("splitatt_fc_bwd");
    splitatt_param_grad_kernel<<<dim3((unsigned)((maxel + 127) / 128), 4), 128, 0, s>>>(*p, scratch, dw1, db1, dgamma, dbeta, dw2, db2);
    TBI_CUDA_LAUNCH_CHECK("splitatt_param_grad");
    int V = pick_vec(p->dtype, {u, dv, du});
    if (c % V != 0) V = 1;
    const int hw = p->h * p->w;
    const float* dgap = scratch + (size_t)N * K * R * c;
    const unsigned g = grid_for((long long)N * hw * (dv->c / V), 256);
    DISPATCH_TV(p->dtype, V,
        (splitatt_du_kernel<float, 4><<<g, 256, 0, s>>>(N, hw, K, R, c, p->act, *u, *dv, *du, p->att, dgap)),
        (splitatt_du_kernel<float, 1><<<g, 256, 0, s>>>(N, hw, K, R, c, p->act, *u, *dv, *du, p->att, dgap)),
        (splitatt_du_kernel<__nv_bfloat16, 8><<<g, 256, 0, s>>>(N, hw, K, R, c, p->act, *u, *dv, *du, p->att, dgap)),
        (splitatt_du_kernel<__nv_bfloat16, 1><<<g, 256, 0, s>>>(N, hw, K, R, c, p->act, *u, *dv, *du, p->att, dgap)));
    TBI_CUDA_LAUNCH_CHECK("splitatt_du");
    return TBI_OK;
}

extern "C" int tbi_softmax_loss_fwd_bwd(int dlogits_dtype, int n, int h, int w, int nc, const float* logits, const float* y,
                                        float* probs, float* loss_map, int32_t* correct, void* dlogits, int dlogits_cstride,
                                        void* stream) {
    const int dl_cs = dlogits_cstride > 0 ? dlogits_cstride : nc;
    TBI_CHECK(dl_cs >= nc, TBI_ERR_BAD_SHAPE, "softmax_loss: dlogits_cstride %d < num_class %d", dl_cs, nc);
    TBI_CHECK(nc >= 3 && nc <= 4, TBI_ERR_UNSUPPORTED, "softmax_loss: num_class %d (my_loss_cat hard-codes 3 classes; 3..4 supported)", nc);
    cudaStream_t s = (cudaStream_t)stream;
    const int hw = h * w;
    const unsigned g = (hw + 31) / 32;
    if (dlogits_dtype == TBI_F32) {
        if (nc == 3) softmax_loss_kernel<3, float, false><<<g, 32 * SL_J, 0, s>>>(n, hw, logits, y, probs, loss_map, correct, (float*)dlogits, dl_cs, nullptr, 0, nullptr, w);
        else         softmax_loss_kernel<4, float, false><<<g, 32 * SL_J, 0, s>>>(n, hw, logits, y, probs, loss_map, correct, (float*)dlogits, dl_cs, nullptr, 0, nullptr, w);
    } else if (dlogits_dtype == TBI_BF16) {
        if (nc == 3) softmax_loss_kernel<3, __nv_bfloat16, false><<<g, 32 * SL_J, 0, s>>>(n, hw, logits, y, probs, loss_map, correct, (__nv_bfloat16*)dlogits, dl_cs, nullptr, 0, nullptr, w);
        else         softmax_loss_kernel<4, __nv_bfloat16, false><<<g, 32 * SL_J, 0, s>>>(n, hw, logits, y, probs, loss_map, correct, (__nv_bfloat16*)dlogits, dl_cs, nullptr, 0, nullptr, w);
    } else return tbi_set_error(TBI_ERR_UNSUPPORTED, "softmax_loss: dlogits dtype");
    TBI_CUDA_LAUNCH_CHECK("softmax_loss");
    return TBI_OK;
}

extern "C" int tbi_softmax_loss_fwd_bwd_taps(int dlogits_dtype, int n, int h, int w, int nc, const tbi_view* ytaps, const float* bias,
                                             const float* y, float* probs, float* loss_map, int32_t* correct, void* dlogits,
                                             int dlogits_cstride, void* stream) {
    const int dl_cs = dlogits_cstride > 0 ? dlogits_cstride : nc;
    TBI_CHECK(dl_cs >= nc, TBI_ERR_BAD_SHAPE, "softmax_loss_taps: dlogits_cstride %d < num_class %d", dl_cs, nc);
    TBI_CHECK(nc >= 3 && nc <= 4, TBI_ERR_UNSUPPORTED, "softmax_loss_taps: num_class %d", nc);
    TBI_CHECK((h % 2) == 0 && (w % 2) == 0 && ytaps && ytaps->ptr && ytaps->h == h / 2 && ytaps->w == w / 2 && ytaps->c >= 16 * nc && ytaps->coff == 0,
              TBI_ERR_BAD_SHAPE, "softmax_loss_taps: Y must be fp32 [n, h/2, w/2, >= 16*nc] for an [n, h, w] output grid");
    cudaStream_t s = (cudaStream_t)stream;
    const int hw = h * w;
    const unsigned g = (hw + 31) / 32;
    const float* yt = (const float*)ytaps->ptr;
    if (dlogits_dtype == TBI_F32) {
        if (nc == 3) softmax_loss_kernel<3, float, true><<<g, 32 * SL_J, 0, s>>>(n, hw, nullptr, y, probs, loss_map, correct, (float*)dlogits, dl_cs, yt, ytaps->cstride, bias, w);
        else         softmax_loss_kernel<4, float, true><<<g, 32 * SL_J, 0, s>>>(n, hw, nullptr, y, probs, loss_map, correct, (float*)dlogits, dl_cs, yt, ytaps->cstride, bias, w);
    } else if (dlogits_dtype == TBI_BF16) {
        if (nc == 3) softmax_loss_kernel<3, __nv_bfloat16, true><<<g, 32 * SL_J, 0, s>>>(n, hw, nullptr, y, probs, loss_map, correct, (__nv_bfloat16*)dlogits, dl_cs, yt, ytaps->cstride, bias, w);
        else         softmax_loss_kernel<4, __nv_bfloat16, true><<<g, 32 * SL_J, 0, s>>>(n, hw, nullptr, y, probs, loss_map, correct, (__nv_bfloat16*)dlogits, dl_cs, yt, ytaps->cstride, bias, w);
    } else return tbi_set_error(TBI_ERR_UNSUPPORTED, "softmax_loss_taps: dlogits dtype");
    TBI_CUDA_LAUNCH_CHECK("softmax_loss_taps");
    return TBI_OK;
}

extern "C" int tbi_convt_phase_taps(int ksize, int a, int b, int* ky, int* kx, int* dy, int* dx) {
    if (ksize != 3 && ksize != 4) return tbi_set_error(TBI_ERR_UNSUPPORTED, "convT ksize %d", ksize);
    const int pad = ksize == 4 ? 1 : 0;                 // TF 'same': k=4 -> torch padding 1; k=3 -> padding 0 + crop
    int n = 0;
    for (int y = 0; y < ksize; ++y) {
        if (((a + pad - y) & 1) != 0) continue;
        for (int x = 0; x < ksize; ++x) {
            if (((b + pad - x) & 1) != 0) continue;
            ky[n] = y; kx[n] = x; dy[n] = (a + pad - y) / 2; dx[n] = (b + pad - x) / 2; ++n;
        }
    }
    return n;
}

extern "C" int tbi_pack_conv_weights(int dtype, int mode, int ksize, int groups, int cin_g, int cout_total, const float* w_hwio,
                                     const float* scale, void* out, void* stream) {
    TBI_CHECK(groups >= 1 && cout_total % groups == 0, TBI_ERR_BAD_SHAPE, "pack_conv: groups");
    cudaStream_t s = (cudaStream_t)stream;
    const int ntaps = ksize * ksize;
    if (tbi_conv_dense_expand(dtype, groups, cin_g, cout_total / groups)) {
        const int cinp = (groups * cin_g + 15) & ~15;
        const unsigned ge = grid_for((long long)ntaps * cinp * cout_total, 256);
        pack_conv_expand_kernel<__nv_bfloat16><<<ge, 256, 0, s>>>(mode, ntaps, groups, cin_g, cinp, cout_total, w_hwio, scale, (__nv_bfloat16*)out);
        TBI_CUDA_LAUNCH_CHECK("pack_conv_expand");
        return TBI_OK;
    }
    const unsigned g = grid_for((long long)ntaps * cin_g * cout_total, 256);
    if (dtype == TBI_F32) pack_conv_kernel<float><<<g, 256, 0, s>>>(mode, ntaps, groups, cin_g, cout_total, w_hwio, scale, (float*)out);
    else if (dtype == TBI_BF16) pack_conv_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(mode, ntaps, groups, cin_g, cout_total, w_hwio, scale, (__nv_bfloat16*)out);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "pack dtype");
    TBI_CUDA_LAUNCH_CHECK("pack_conv");
    return TBI_OK;
}

int tbi_wgrad_gather_blocks(int ntaps, int groups, int cin_g, int cin_pad, int cout_total, const float* dense, float* dw, cudaStream_t s) {
    wgrad_gather_blocks_kernel<<<grid_for((long long)ntaps * cin_g * cout_total, 256), 256, 0, s>>>(ntaps, groups, cin_g, cin_pad, cout_total, dense, dw);
    TBI_CUDA_LAUNCH_CHECK("wgrad_gather_blocks");
    return TBI_OK;
}

extern "C" int tbi_pack_convt_weights(int dtype, int mode, int ksize, int cin, int cout, int cout_pad, const float* w_hwoi,
                                      const float* scale, void* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const int cpad = ((mode == 1 && cout_pad > cout) || mode >= 2) ? cout_pad : cout;
    TBI_CHECK(mode < 2 || cout_pad >= ksize * ksize * cout, TBI_ERR_BAD_SHAPE, "pack_convt mode %d: cout_pad %d < k*k*cout", mode, cout_pad);
    ConvtTapTable tt{};
    int off = 0;
    for (int ph = 0; ph < 4; ++ph) {
        int ky[TBI_MAX_TAPS], kx[TBI_MAX_TAPS], dy[TBI_MAX_TAPS], dx[TBI_MAX_TAPS];
        const int n = tbi_convt_phase_taps(ksize, ph >> 1, ph & 1, ky, kx, dy, dx);
        if (n < 0) return n;
        tt.n[ph] = n; tt.off[ph] = off; off += n;
        for (int t = 0; t < n; ++t) { tt.ky[ph][t] = ky[t]; tt.kx[ph][t] = kx[t]; }
    }
    const unsigned g = grid_for(mode >= 2 ? (long long)cin * cpad : (long long)ksize * ksize * cin * cpad, 256);
    if (dtype == TBI_F32) pack_convt_kernel<float><<<g, 256, 0, s>>>(mode, ksize, cin, cout, cpad, tt, w_hwoi, scale, (float*)out);
    else if (dtype == TBI_BF16) pack_convt_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(mode, ksize, cin, cout, cpad, tt, w_hwoi, scale, (__nv_bfloat16*)out);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "pack dtype");
    TBI_CUDA_LAUNCH_CHECK("pack_convt");
    return TBI_OK;
}

extern "C" int tbi_bn_fold(int c, const float* gamma, const float* beta, const float* mean, const float* var, const float* bias,
                           float eps, float* scale, float* fbias, void* stream) {
    bn_fold_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, gamma, beta, mean, var, bias, eps, scale, fbias);
    TBI_CUDA_LAUNCH_CHECK("bn_fold");
    return TBI_OK;
}

extern "C" int tbi_bn_param_grad(int c, int64_t k_outer, int64_t inner, int64_t outer_stride, int64_t co_stride, const float* w,
                                 float* dw, const float* bias, float* dbias, const float* gamma, const float* mean,
                                 const float* var, float eps, float* dgamma, float* dbeta, void* stream) {
    static const bool rows_only = getenv("TBI_BN_ROWS") != nullptr;
    if (inner == 1 && co_stride == 1 && !rows_only) {
        long long ks = k_outer / 64;
        if (ks < 1) ks = 1;
        if (ks > 32) ks = 32;
        bn_param_grad_cols_kernel<<<dim3((unsigned)((c + 31) / 32), (unsigned)ks), dim3(32, 8), 0, (cudaStream_t)stream>>>(
            c, k_outer, outer_stride, w, dw, bias, dbias, gamma, mean, var, eps, dgamma, dbeta);
    } else {
        bn_param_grad_kernel<<<c, 256, 0, (cudaStream_t)stream>>>(c, k_outer, inner, outer_stride, co_stride, w, dw, bias, dbias,
                                                                  gamma, mean, var, eps, dgamma, dbeta);
    }
    TBI_CUDA_LAUNCH_CHECK("bn_param_grad");
    return TBI_OK;
}

extern "C" int tbi_adam_multi(int64_t count, float* p, const float* g, float* m, float* v, const int32_t* step_count, float lr,
                              float b1, float b2, float eps, float grad_scale, void* stream) {
    TBI_CHECK((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, TBI_ERR_BAD_ALIGN, "adam: 16B alignment");
    const unsigned gsz = grid_for(count / 4 + 1, 256, 8);
    adam_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>(count, p, g, m, v, step_count, lr, b1, b2, eps, grad_scale);
    TBI_CUDA_LAUNCH_CHECK("adam");
    return TBI_OK;
}
extern "C" int tbi_adam_multi_dev(int64_t count, float* p, const float* g, float* m, float* v, const int32_t* step_count,
                                  const float* hyper, const float* gnorm_sq, float b1, float b2, float eps, void* stream) {
    TBI_CHECK((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, TBI_ERR_BAD_ALIGN, "adam: 16B alignment");
    TBI_CHECK(hyper != nullptr, TBI_ERR_BAD_SHAPE, "adam_dev: hyper is NULL");
    const unsigned gsz = grid_for(count / 4 + 1, 256, 8);
    adam_dev_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>(count, p, g, m, v, step_count, hyper, gnorm_sq, b1, b2, eps);
    TBI_CUDA_LAUNCH_CHECK("adam_dev");
    return TBI_OK;
}
extern "C" int tbi_sumsq(int64_t count, const float* x, float* out, void* stream) {
    TBI_CHECK(((uintptr_t)x & 15) == 0, TBI_ERR_BAD_ALIGN, "sumsq: 16B alignment");
    unsigned gsz = grid_for(count / 4 + 1, 256, 8);
    if (gsz > 1184) gsz = 1184;
    sumsq_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>(count, x, out);
    TBI_CUDA_LAUNCH_CHECK("sumsq");
    return TBI_OK;
}
extern "C" int tbi_adam_advance(int32_t* step_count, void* stream) {
    adam_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_count);
    TBI_CUDA_LAUNCH_CHECK("adam_advance");
    return TBI_OK;
}

extern "C" int tbi_dropout_mask(uint8_t* keep, int64_t count, uint64_t seed, const int32_t* step_ptr, void* stream) {
    dropout_mask_kernel<<<grid_for(count / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(keep, count, seed, step_ptr);
    TBI_CUDA_LAUNCH_CHECK("dropout_mask");
    return TBI_OK;
}

extern "C" int tbi_cast(int src_is_f32, int dst_dtype, int64_t count, const void* src, void* dst, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(count, 256);
    if (src_is_f32) {
        if (dst_dtype == TBI_F32) cast_kernel<float, float><<<g, 256, 0, s>>>(count, (const float*)src, (float*)dst);
        else cast_kernel<float, __nv_bfloat16><<<g, 256, 0, s>>>(count, (const float*)src, (__nv_bfloat16*)dst);
    } else {
        if (dst_dtype == TBI_F32) cast_kernel<double, float><<<g, 256, 0, s>>>(count, (const double*)src, (float*)dst);
        else cast_kernel<double, __nv_bfloat16><<<g, 256, 0, s>>>(count, (const double*)src, (__nv_bfloat16*)dst);
    }
    TBI_CUDA_LAUNCH_CHECK("cast");
    return TBI_OK;
}

// ---------------------------------------------------------------------------------------------
// One-launch weight preparation (tbi_prepare_run, tbi_bn_fold_multi)
// ---------------------------------------------------------------------------------------------
// The per-layer entry points above (tbi_pack_conv_weights / tbi_pack_convt_weights / tbi_bn_fold) cost ~85 launches per step;
// although they ran on a side stream under the stem's forward they took 0.35 ms of the 8.9 ms step (measured by skipping
// them).  Every packing mode is the same operation on a different index map:
//     out[out_tap[t] + a*out_a + b*out_b] = S[src_tap_index[t]*src_tap + a*src_a + b] * scale(co),   co = co_base + (co_is_a ? a : b)
// with S the fp32 master kernel of one tap as a matrix [A][B] (b contiguous) and scale = gamma/sqrt(var+eps) (BN folded) or 1.
// A table of such items (built once by the host, tbi_prep_item) is executed by ONE launch: a block moves a 32x32 tile through
// shared memory so that both the fp32 reads and the packed writes are coalesced whichever of a/b is contiguous in the output.
namespace {

template <typename T>
__global__ void __launch_bounds__(256) prepare_kernel(const tbi_prep_item* __restrict__ items, int nitems, int total_tiles, float eps) {
    // A block owns a CONTIGUOUS range of tiles: it finds its first item once and then walks the table (the first version
    // searched the table per tile from one thread -- seven dependent global loads in front of every 4 KB tile -- and ran at
    // 0.9 TB/s: 218 us for 200 MB).  Two tile buffers: one barrier per tile.
    __shared__ float tile[2][32][33];
    const int per = (total_tiles + gridDim.x - 1) / gridDim.x;
    int tix = blockIdx.x * per;
    const int end = min(total_tiles, tix + per);
    if (tix >= end) return;
    int cur;
    {
        int lo = 0, hi = nitems - 1;                        // last item with tile_begin <= tix (every thread: uniform, cached loads)
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (items[mid].tile_begin <= tix) lo = mid; else hi = mid - 1; }
        cur = lo;
    }
    int next_begin = cur + 1 < nitems ? items[cur + 1].tile_begin : 0x7fffffff;
    // per-item constants live in registers (read through a const reference to global memory the compiler re-loaded every field
    // after every store: ncu showed 8x the load requests the data needs and 134 M warp instructions, the kernel was issue-bound)
    int A = 0, B = 0, tiles_a = 1, tiles_b = 1, tile_begin = 0, co_base = 0, co_is_a = 0, src_a = 0, out_a = 0, out_b = 0;
    long long src_tap = 0;
    const float* src0 = nullptr; const float* gamma = nullptr; const float* var = nullptr; T* out0 = nullptr;
    auto load_item = [&](int i) {
        const tbi_prep_item& it = items[i];
        A = it.A; B = it.B; tiles_a = it.tiles_a; tiles_b = it.tiles_b; tile_begin = it.tile_begin; co_base = it.co_base; co_is_a = it.co_is_a;
        src_a = (int)it.src_a; out_a = (int)it.out_a; out_b = (int)it.out_b; src_tap = it.src_tap;
        src0 = it.src; gamma = it.gamma; var = it.var; out0 = (T*)it.out;
    };
    load_item(cur);
    for (int k = 0; tix < end; ++tix, k ^= 1) {
        if (tix >= next_begin) {
            while (tix >= next_begin) { ++cur; next_begin = cur + 1 < nitems ? items[cur + 1].tile_begin : 0x7fffffff; }
            load_item(cur);
        }
        int r = tix - tile_begin;
        const int tb = r % tiles_b; r /= tiles_b;
        const int ta = r % tiles_a; const int t = r / tiles_a;
        const int a0 = ta * 32, b0 = tb * 32;
        const float* src = src0 + (long long)items[cur].src_tap_index[t] * src_tap;
        // read: threads along b (contiguous in the source); the BN scale of a column (or of each of the thread's 4 rows) once
        const int b = b0 + threadIdx.x;
        const bool bok = b < B;
        float sb = 1.f;
        if (gamma && !co_is_a && bok) sb = gamma[co_base + b] * rsqrtf(var[co_base + b] + eps);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = a0 + threadIdx.y + 8 * j;
            float v = 0.f;
            if (a < A && bok) {
                v = src[a * src_a + b];
                if (gamma) v *= co_is_a ? gamma[co_base + a] * rsqrtf(var[co_base + a] + eps) : sb;
            }
            tile[k][threadIdx.y + 8 * j][threadIdx.x] = v;
        }
        __syncthreads();
        T* out = out0 + items[cur].out_tap[t];
        if (out_b == 1) {                                       // b contiguous in the output as well: straight copy
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int a = a0 + threadIdx.y + 8 * j;
                if (a < A && bok) stf(out + a * out_a + b, tile[k][threadIdx.y + 8 * j][threadIdx.x]);
            }
        } else {                                                // a contiguous (or strided) in the output: threads along a
            const int a = a0 + threadIdx.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int bb = b0 + threadIdx.y + 8 * j;
                if (a < A && bb < B) stf(out + a * out_a + bb * out_b, tile[k][threadIdx.x][threadIdx.y + 8 * j]);
            }
        }
        // buffer k is rewritten two tiles from now, behind the next tile's barrier: no second barrier needed
    }
}

__global__ void bn_fold_multi_kernel(const tbi_fold_item* __restrict__ items, float eps) {
    const tbi_fold_item& it = items[blockIdx.x];
    for (int i = threadIdx.x; i < it.c; i += blockDim.x) {
        const float b = it.bias ? it.bias[i] : 0.f;
        if (it.gamma) {
            const float s = it.gamma[i] * rsqrtf(it.var[i] + eps);
            it.scale[i] = s; it.fbias[i] = (b - it.mean[i]) * s + it.beta[i];
        } else { it.scale[i] = 1.f; it.fbias[i] = b; }
    }
}

}  // namespace

extern "C" int tbi_prepare_run(int dtype, const tbi_prep_item* items_dev, int nitems, int total_tiles, float bn_eps, void* stream) {
    TBI_CHECK(items_dev && nitems > 0 && total_tiles > 0, TBI_ERR_BAD_SHAPE, "prepare_run: empty table");
    long long grid = total_tiles;
    const long long cap = (long long)tbi_sm_count() * 32;
    if (grid > cap) grid = cap;
    if (dtype == TBI_F32) prepare_kernel<float><<<(unsigned)grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(items_dev, nitems, total_tiles, bn_eps);
    else if (dtype == TBI_BF16) prepare_kernel<__nv_bfloat16><<<(unsigned)grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(items_dev, nitems, total_tiles, bn_eps);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "prepare_run dtype");
    TBI_CUDA_LAUNCH_CHECK("prepare_run");
    return TBI_OK;
}

extern "C" int tbi_bn_fold_multi(const tbi_fold_item* items_dev, int nitems, float bn_eps, void* stream) {
    TBI_CHECK(items_dev && nitems > 0, TBI_ERR_BAD_SHAPE, "bn_fold_multi: empty table");
    bn_fold_multi_kernel<<<nitems, 256, 0, (cudaStream_t)stream>>>(items_dev, bn_eps);
    TBI_CUDA_LAUNCH_CHECK("bn_fold_multi");
    return TBI_OK;
}
