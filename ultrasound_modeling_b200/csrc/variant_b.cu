// Kernels that only Variant B of the reference needs (ResNest.py / Decoder.py): LayerNormalization over the channel
// axis (Keras: axis -1, eps 1e-3, biased variance) fused with its activation, and the split-attention of
// ResNest.py:171-199 whose R inputs are the SAME tensor and whose dense2 is shared between the R splits.
// Channel counts here are 3..256 and mostly not multiples of 8 (10, 21, 30, 42, 63, 85, 126, 255), so everything is
// written for arbitrary C: one warp per pixel, lanes stride over channels.
#include "tbi_common.cuh"

namespace {

// y = act(gamma * (x - mean_c) * rsqrt(var_c + eps) + beta), one warp per pixel.  x and y may alias.
template <typename T>
__global__ void __launch_bounds__(256) layernorm_c_fwd_kernel(long long npix, int c, tbi_view x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int act, tbi_view y) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp0; p < npix; p += nwarps) {
        const T* xp = (const T*)x.ptr + (size_t)p * x.cstride + x.coff;
        T* yp = (T*)y.ptr + (size_t)p * y.cstride + y.coff;
        float s = 0.f, ss = 0.f;
        for (int ch = lane; ch < c; ch += 32) { const float v = ldf(xp + ch); s += v; ss = fmaf(v, v, ss); }
        s = warp_sum(s); ss = warp_sum(ss);
        const float mean = s / (float)c;
        const float var = fmaxf(ss / (float)c - mean * mean, 0.f);
        const float istd = rsqrtf(var + eps);
        for (int ch = lane; ch < c; ch += 32) {
            const float v = (ldf(xp + ch) - mean) * istd * gamma[ch] + beta[ch];
            stf(yp + ch, act_apply(act, v));
        }
    }
}

// Backward of the above given x (the LayerNorm INPUT), y (its activated output, for act') and dy:
//   g = dy * act'(y) ; dgamma += g * xhat ; dbeta += g ; h = g * gamma ;
//   dx = istd * (h - mean_c(h) - xhat * mean_c(h * xhat))
// Parameter gradients: per-thread partial sums over the warp's pixels would need C registers; instead each block
// accumulates in shared memory (C <= 1024) and issues one atomic per channel.
template <typename T>
__global__ void __launch_bounds__(256) layernorm_c_bwd_kernel(long long npix, int c, tbi_view x, tbi_view y, tbi_view dy,
                                                              const float* __restrict__ gamma, float eps, int act, tbi_view dx,
                                                              float* dgamma, float* dbeta) {
    extern __shared__ float sm[];                 // [2][c]
    float* sg = sm; float* sb = sm + c;
    for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp0; p < npix; p += nwarps) {
        const T* xp = (const T*)x.ptr + (size_t)p * x.cstride + x.coff;
        const T* yp = (const T*)y.ptr + (size_t)p * y.cstride + y.coff;
        const T* gp = (const T*)dy.ptr + (size_t)p * dy.cstride + dy.coff;
        T* dp = (T*)dx.ptr + (size_t)p * dx.cstride + dx.coff;
        float s = 0.f, ss = 0.f;
        for (int ch = lane; ch < c; ch += 32) { const float v = ldf(xp + ch); s += v; ss = fmaf(v, v, ss); }
        s = warp_sum(s); ss = warp_sum(ss);
        const float mean = s / (float)c;
        const float istd = rsqrtf(fmaxf(ss / (float)c - mean * mean, 0.f) + eps);
        float sh = 0.f, shx = 0.f;
        for (int ch = lane; ch < c; ch += 32) {
            const float xhat = (ldf(xp + ch) - mean) * istd;
            const float g = ldf(gp + ch) * act_grad_from_out(act, ldf(yp + ch));
            atomicAdd(sg + ch, g * xhat); atomicAdd(sb + ch, g);
            const float h = g * gamma[ch];
            sh += h; shx = fmaf(h, xhat, shx);
        }
        sh = warp_sum(sh) / (float)c; shx = warp_sum(shx) / (float)c;
        for (int ch = lane; ch < c; ch += 32) {
            const float xhat = (ldf(xp + ch) - mean) * istd;
            const float h = ldf(gp + ch) * act_grad_from_out(act, ldf(yp + ch)) * gamma[ch];
            stf(dp + ch, istd * (h - sh - xhat * shx));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) { atomicAdd(dgamma + i, sg[i]); atomicAdd(dbeta + i, sb[i]); }
}

// raw[n][ch] += sum over a chunk of pixels of u[n,p,ch]
template <typename T>
__global__ void __launch_bounds__(256) shared_gap_kernel(int hw, int C, int c, int cpad, tbi_view u, float* raw, int pix_per_block) {
    const int n = blockIdx.y;
    const int pbeg = blockIdx.x * pix_per_block, pend = min(hw, pbeg + pix_per_block);
    const T* ub = (const T*)u.ptr + (size_t)n * hw * u.cstride + u.coff;
    // threads: channel lanes x pixel lanes
    const int cl = min(C, (int)blockDim.x), pl = blockDim.x / cl;
    const int lc = threadIdx.x % cl, lp = threadIdx.x / cl;
    if (lp >= pl) return;
    for (int ch = lc; ch < C; ch += cl) {
        float s = 0.f;
        const int pch = ch + (ch / c) * cpad;               // cardinal k's channels start at k*(c + cpad) in a padded record
        for (int p = pbeg + lp; p < pend; p += pl) s += ldf(ub + (size_t)p * u.cstride + pch);
        atomicAdd(raw + (size_t)n * C + ch, s);
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += red[i];
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
    return r;
}

// one block per (n, k): g = R * mean_hw(u) -> dense1 -> LayerNorm -> act -> dense2 (shared by the R splits) ->
// softmax over channels (R > 1) or sigmoid (R == 1); att[n][k][ch] = R * a  (V = sum_r U * a = R * U * a)
__global__ void __launch_bounds__(128) shared_fc_kernel(int hw, int K, int R, int c, const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ lng, const float* __restrict__ lnb, float eps, int act,
                                                        const float* __restrict__ w2, const float* __restrict__ b2, float* att) {
    extern __shared__ float sm[];
    const int c2 = c / 2;
    float* g = sm; float* h1 = sm + c; float* z = h1 + c2; float* red = z + c;
    const int n = blockIdx.x, k = blockIdx.y;
    float* a = att + ((size_t)n * K + k) * c;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) g[ch] = a[ch] * ((float)R / (float)hw);
    __syncthreads();
    float ls = 0.f, lss = 0.f;
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        float q = b1[k * c2 + j];
        const float* w = w1 + (size_t)k * c * c2 + j;
        for (int ch = 0; ch < c; ++ch) q = fmaf(g[ch], w[(size_t)ch * c2], q);
        h1[j] = q; ls += q; lss = fmaf(q, q, lss);
    }
    const float mean = block_sum(ls, red) / (float)c2;
    const float var = fmaxf(block_sum(lss, red) / (float)c2 - mean * mean, 0.f);
    const float istd = rsqrtf(var + eps);
    __syncthreads();
    for (int j = threadIdx.x; j < c2; j += blockDim.x)
        h1[j] = act_apply(act, (h1[j] - mean) * istd * lng[k * c2 + j] + lnb[k * c2 + j]);
    __syncthreads();
    float lmax = -INFINITY;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float q = b2[k * c + ch];
        const float* w = w2 + (size_t)k * c2 * c + ch;
        for (int j = 0; j < c2; ++j) q = fmaf(h1[j], w[(size_t)j * c], q);
        z[ch] = q; lmax = fmaxf(lmax, q);
    }
    if (R == 1) {
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) a[ch] = 1.f / (1.f + expf(-z[ch]));
        return;
    }
    const float m = block_max(lmax, red);
    float lsum = 0.f;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) { const float e = expf(z[ch] - m); z[ch] = e; lsum += e; }
    const float tot = block_sum(lsum, red);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) a[ch] = (float)R * z[ch] / tot;
}

// v[n,p,ch] = u[n,p,ch] * att[n][ch]
template <typename T>
__global__ void __launch_bounds__(256) shared_scale_kernel(long long total, int hw, int C, int c, int cpad, tbi_view u, tbi_view v, const float* __restrict__ att) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C); const long long pg = i / C; const int n = (int)(pg / hw);
        const int pch = ch + (ch / c) * cpad;
        const float x = ldf((const T*)u.ptr + (size_t)pg * u.cstride + u.coff + pch);
        stf((T*)v.ptr + (size_t)pg * v.cstride + v.coff + pch, x * att[(size_t)n * C + ch]);
    }
}

// ---- backward of the shared-input split attention --------------------------------------------------------------------
// su[n][ch] += sum_p u ; sdu[n][ch] += sum_p dv*u   (one pass over u and dv)
template <typename T>
__global__ void __launch_bounds__(256) shared_bwd_reduce_kernel(int hw, int C, int c, int cpad, tbi_view u, tbi_view dv, float* su, float* sdu, int pix_per_block) {
    const int n = blockIdx.y;
    const int pbeg = blockIdx.x * pix_per_block, pend = min(hw, pbeg + pix_per_block);
    const T* ub = (const T*)u.ptr + (size_t)n * hw * u.cstride + u.coff;
    const T* db = (const T*)dv.ptr + (size_t)n * hw * dv.cstride + dv.coff;
    const int cl = min(C, (int)blockDim.x), pl = blockDim.x / cl;
    const int lc = threadIdx.x % cl, lp = threadIdx.x / cl;
    if (lp >= pl) return;
    for (int ch = lc; ch < C; ch += cl) {
        float s = 0.f, sd = 0.f;
        const int pch = ch + (ch / c) * cpad;
        for (int p = pbeg + lp; p < pend; p += pl) {
            const float x = ldf(ub + (size_t)p * u.cstride + pch);
            s += x; sd = fmaf(ldf(db + (size_t)p * dv.cstride + pch), x, sd);
        }
        atomicAdd(su + (size_t)n * C + ch, s);
        atomicAdd(sdu + (size_t)n * C + ch, sd);
    }
}

// one block per (n, k): recompute the FC chain from su, then its backward.  V = U * att with att = R * a, so
// dL/da = R * sum_p dV*U.  Parameter gradients are ADDED with atomics (reduction over n and k blocks).
// On return su[n][k*c + ch] holds dgap * R / hw, the per-image constant of dU.
__global__ void __launch_bounds__(128) shared_fc_bwd_kernel(int hw, int K, int R, int c, const float* __restrict__ w1, const float* __restrict__ b1,
                                                            const float* __restrict__ lng, const float* __restrict__ lnb, float eps, int act,
                                                            const float* __restrict__ w2, const float* __restrict__ att, float* su,
                                                            const float* __restrict__ sdu, float* dw1, float* db1, float* dlng, float* dlnb,
                                                            float* dw2, float* db2) {
    extern __shared__ float sm[];
    const int c2 = c / 2;
    float* g = sm; float* xh = g + c; float* h1 = xh + c2; float* dz = h1 + c2; float* dq = dz + c; float* red = dq + c2;
    const int n = blockIdx.x, k = blockIdx.y;
    float* sun = su + ((size_t)n * K + k) * c;
    const float* sdn = sdu + ((size_t)n * K + k) * c;
    const float* an = att + ((size_t)n * K + k) * c;
    const float rs = (float)R / (float)hw;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) g[ch] = sun[ch] * rs;
    __syncthreads();
    float ls = 0.f, lss = 0.f;
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        float q = b1[k * c2 + j];
        const float* w = w1 + (size_t)k * c * c2 + j;
        for (int ch = 0; ch < c; ++ch) q = fmaf(g[ch], w[(size_t)ch * c2], q);
        xh[j] = q; ls += q; lss = fmaf(q, q, lss);
    }
    const float mean = block_sum(ls, red) / (float)c2;
    const float var = fmaxf(block_sum(lss, red) / (float)c2 - mean * mean, 0.f);
    const float istd = rsqrtf(var + eps);
    __syncthreads();
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        const float x = (xh[j] - mean) * istd;
        xh[j] = x;
        h1[j] = act_apply(act, x * lng[k * c2 + j] + lnb[k * c2 + j]);
    }
    // softmax / sigmoid backward
    float ldot = 0.f;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) ldot = fmaf(an[ch] / (float)R, (float)R * sdn[ch], ldot);
    const float dot = block_sum(ldot, red);                  // (also orders the h1 / xh writes above)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const float a = an[ch] / (float)R, da = (float)R * sdn[ch];
        const float z = R == 1 ? a * (1.f - a) * da : a * (da - dot);
        dz[ch] = z;
        atomicAdd(db2 + k * c + ch, z);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c2 * c; i += blockDim.x) {
        const int j = i / c, ch = i - j * c;
        atomicAdd(dw2 + (size_t)k * c2 * c + i, h1[j] * dz[ch]);
    }
    float lh = 0.f, lhx = 0.f;
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        const float* w = w2 + ((size_t)k * c2 + j) * c;
        float d = 0.f;
        for (int ch = 0; ch < c; ++ch) d = fmaf(dz[ch], w[ch], d);
        const float dy = d * act_grad_from_out(act, h1[j]);
        atomicAdd(dlng + k * c2 + j, dy * xh[j]);
        atomicAdd(dlnb + k * c2 + j, dy);
        const float hh = dy * lng[k * c2 + j];
        dq[j] = hh; lh += hh; lhx = fmaf(hh, xh[j], lhx);
    }
    const float m1 = block_sum(lh, red) / (float)c2;
    const float m2 = block_sum(lhx, red) / (float)c2;
    __syncthreads();
    for (int j = threadIdx.x; j < c2; j += blockDim.x) {
        const float d = istd * (dq[j] - m1 - xh[j] * m2);
        dq[j] = d;
        atomicAdd(db1 + k * c2 + j, d);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c * c2; i += blockDim.x) {
        const int ch = i / c2, j = i - ch * c2;
        atomicAdd(dw1 + (size_t)k * c * c2 + i, g[ch] * dq[j]);
    }
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const float* w = w1 + ((size_t)k * c + ch) * c2;
        float d = 0.f;
        for (int j = 0; j < c2; ++j) d = fmaf(dq[j], w[j], d);
        sun[ch] = d * rs;
    }
}

// du[n,p,ch] = dv[n,p,ch] * att[n][ch] + dgs[n][ch]
template <typename T>
__global__ void __launch_bounds__(256) shared_du_kernel(long long total, int hw, int C, int c, int cpad, tbi_view dv, tbi_view du, const float* __restrict__ att,
                                                        const float* __restrict__ dgs) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C); const long long pg = i / C; const int n = (int)(pg / hw);
        const int pch = ch + (ch / c) * cpad;
        const float g = ldf((const T*)dv.ptr + (size_t)pg * dv.cstride + dv.coff + pch);
        stf((T*)du.ptr + (size_t)pg * du.cstride + du.coff + pch, fmaf(g, att[(size_t)n * C + ch], dgs[(size_t)n * C + ch]));
    }
}

unsigned grid_cap(long long work, int per_block, int waves) {
    long long b = (work + per_block - 1) / per_block;
    const long long cap = (long long)tbi_sm_count() * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace

extern "C" int tbi_layernorm_c_fwd(int dtype, int64_t npix, int c, const tbi_view* x, const float* gamma, const float* beta, float eps,
                                   int act, const tbi_view* y, void* stream) {
    TBI_CHECK(x && y && gamma && beta && c > 0, TBI_ERR_BAD_SHAPE, "layernorm_c_fwd: null argument or c <= 0");
    TBI_CHECK(x->c == c && y->c == c, TBI_ERR_BAD_SHAPE, "layernorm_c_fwd: views have %d / %d channels, expected %d", x->c, y->c, c);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_cap(npix, 8, 16);
    if (dtype == TBI_F32) layernorm_c_fwd_kernel<float><<<g, 256, 0, s>>>(npix, c, *x, gamma, beta, eps, act, *y);
    else if (dtype == TBI_BF16) layernorm_c_fwd_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(npix, c, *x, gamma, beta, eps, act, *y);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "layernorm_c_fwd dtype");
    TBI_CUDA_LAUNCH_CHECK("layernorm_c_fwd");
    return TBI_OK;
}

extern "C" int tbi_layernorm_c_bwd(int dtype, int64_t npix, int c, const tbi_view* x, const tbi_view* y, const tbi_view* dy,
                                   const float* gamma, float eps, int act, const tbi_view* dx, float* dgamma, float* dbeta, void* stream) {
    TBI_CHECK(x && y && dy && dx && gamma && dgamma && dbeta && c > 0, TBI_ERR_BAD_SHAPE, "layernorm_c_bwd: null argument or c <= 0");
    TBI_CHECK(c <= 4096, TBI_ERR_UNSUPPORTED, "layernorm_c_bwd: c = %d > 4096", c);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_cap(npix, 8 * 16, 4);
    const size_t smem = 2 * (size_t)c * sizeof(float);
    if (dtype == TBI_F32) layernorm_c_bwd_kernel<float><<<g, 256, smem, s>>>(npix, c, *x, *y, *dy, gamma, eps, act, *dx, dgamma, dbeta);
    else if (dtype == TBI_BF16) layernorm_c_bwd_kernel<__nv_bfloat16><<<g, 256, smem, s>>>(npix, c, *x, *y, *dy, gamma, eps, act, *dx, dgamma, dbeta);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "layernorm_c_bwd dtype");
    TBI_CUDA_LAUNCH_CHECK("layernorm_c_bwd");
    return TBI_OK;
}

extern "C" int tbi_splitatt_shared_fwd(int dtype, int n, int h, int w, int kpaths, int radix, int c, const tbi_view* u, const tbi_view* v,
                                       const float* w1, const float* b1, const float* ln_gamma, const float* ln_beta, float ln_eps, int act,
                                       const float* w2, const float* b2, float* att, int cgroup, void* stream) {
    TBI_CHECK(u && v && w1 && b1 && ln_gamma && ln_beta && w2 && b2 && att, TBI_ERR_BAD_SHAPE, "splitatt_shared_fwd: null argument");
    if (cgroup <= 0) cgroup = c;
    const int cpad = cgroup - c;
    TBI_CHECK(c >= 2 && kpaths >= 1 && radix >= 1 && cpad >= 0 && u->c == kpaths * cgroup && v->c == kpaths * cgroup, TBI_ERR_BAD_SHAPE,
              "splitatt_shared_fwd: u/v must have kpaths*cgroup = %d channels (got %d, %d)", kpaths * cgroup, u->c, v->c);
    TBI_CHECK(c <= 2048, TBI_ERR_UNSUPPORTED, "splitatt_shared_fwd: c = %d > 2048", c);
    cudaStream_t s = (cudaStream_t)stream;
    const int hw = h * w, C = kpaths * c;
    if (cudaMemsetAsync(att, 0, (size_t)n * C * sizeof(float), s) != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt_shared_fwd: memset");
    int chunks = (tbi_sm_count() * 4 + n - 1) / n;
    if (chunks > (hw + 63) / 64) chunks = (hw + 63) / 64;
    if (chunks < 1) chunks = 1;
    const int ppb = (hw + chunks - 1) / chunks;
    const dim3 gg((unsigned)((hw + ppb - 1) / ppb), (unsigned)n);
    const unsigned ge = grid_cap((long long)n * hw * C, 256 * 4, 16);
    const size_t fsm = (size_t)(c + c / 2 + c + 32) * sizeof(float);
    if (dtype == TBI_F32) {
        shared_gap_kernel<float><<<gg, 256, 0, s>>>(hw, C, c, cpad, *u, att, ppb);
        shared_fc_kernel<<<dim3((unsigned)n, (unsigned)kpaths), 128, fsm, s>>>(hw, kpaths, radix, c, w1, b1, ln_gamma, ln_beta, ln_eps, act, w2, b2, att);
        shared_scale_kernel<float><<<ge, 256, 0, s>>>((long long)n * hw * C, hw, C, c, cpad, *u, *v, att);
    } else if (dtype == TBI_BF16) {
        shared_gap_kernel<__nv_bfloat16><<<gg, 256, 0, s>>>(hw, C, c, cpad, *u, att, ppb);
        shared_fc_kernel<<<dim3((unsigned)n, (unsigned)kpaths), 128, fsm, s>>>(hw, kpaths, radix, c, w1, b1, ln_gamma, ln_beta, ln_eps, act, w2, b2, att);
        shared_scale_kernel<__nv_bfloat16><<<ge, 256, 0, s>>>((long long)n * hw * C, hw, C, c, cpad, *u, *v, att);
    } else return tbi_set_error(TBI_ERR_UNSUPPORTED, "splitatt_shared_fwd dtype");
    TBI_CUDA_LAUNCH_CHECK("splitatt_shared_fwd");
    return TBI_OK;
}

extern "C" int tbi_splitatt_shared_bwd(int dtype, int n, int h, int w, int kpaths, int radix, int c, const tbi_view* u, const tbi_view* dv,
                                       const tbi_view* du, const float* w1, const float* b1, const float* ln_gamma, const float* ln_beta,
                                       float ln_eps, int act, const float* w2, const float* att, float* dw1, float* db1, float* dln_gamma,
                                       float* dln_beta, float* dw2, float* db2, float* scratch, int cgroup, void* stream) {
    TBI_CHECK(u && dv && du && w1 && b1 && ln_gamma && ln_beta && w2 && att && dw1 && db1 && dln_gamma && dln_beta && dw2 && db2 && scratch,
              TBI_ERR_BAD_SHAPE, "splitatt_shared_bwd: null argument");
    if (cgroup <= 0) cgroup = c;
    const int cpad = cgroup - c;
    TBI_CHECK(c >= 2 && kpaths >= 1 && radix >= 1 && cpad >= 0 && u->c == kpaths * cgroup && dv->c == kpaths * cgroup && du->c == kpaths * cgroup, TBI_ERR_BAD_SHAPE,
              "splitatt_shared_bwd: u/dv/du must have kpaths*cgroup = %d channels", kpaths * cgroup);
    TBI_CHECK(c <= 2048, TBI_ERR_UNSUPPORTED, "splitatt_shared_bwd: c = %d > 2048", c);
    cudaStream_t s = (cudaStream_t)stream;
    const int hw = h * w, C = kpaths * c;
    float* su = scratch; float* sdu = scratch + (size_t)n * C;
    if (cudaMemsetAsync(scratch, 0, 2 * (size_t)n * C * sizeof(float), s) != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt_shared_bwd: memset");
    int chunks = (tbi_sm_count() * 4 + n - 1) / n;
    if (chunks > (hw + 63) / 64) chunks = (hw + 63) / 64;
    if (chunks < 1) chunks = 1;
    const int ppb = (hw + chunks - 1) / chunks;
    const dim3 gg((unsigned)((hw + ppb - 1) / ppb), (unsigned)n);
    const unsigned ge = grid_cap((long long)n * hw * C, 256 * 4, 16);
    const size_t fsm = (size_t)(2 * c + 3 * (c / 2) + 32) * sizeof(float);
    if (dtype == TBI_F32) shared_bwd_reduce_kernel<float><<<gg, 256, 0, s>>>(hw, C, c, cpad, *u, *dv, su, sdu, ppb);
    else if (dtype == TBI_BF16) shared_bwd_reduce_kernel<__nv_bfloat16><<<gg, 256, 0, s>>>(hw, C, c, cpad, *u, *dv, su, sdu, ppb);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "splitatt_shared_bwd dtype");
    shared_fc_bwd_kernel<<<dim3((unsigned)n, (unsigned)kpaths), 128, fsm, s>>>(hw, kpaths, radix, c, w1, b1, ln_gamma, ln_beta, ln_eps, act, w2, att, su, sdu,
                                                                               dw1, db1, dln_gamma, dln_beta, dw2, db2);
    if (dtype == TBI_F32) shared_du_kernel<float><<<ge, 256, 0, s>>>((long long)n * hw * C, hw, C, c, cpad, *dv, *du, att, su);
    else shared_du_kernel<__nv_bfloat16><<<ge, 256, 0, s>>>((long long)n * hw * C, hw, C, c, cpad, *dv, *du, att, su);
    TBI_CUDA_LAUNCH_CHECK("splitatt_shared_bwd");
    return TBI_OK;
}
