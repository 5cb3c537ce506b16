// Fused split-attention tail (TBI_ResNest.py:175-207) as ONE cooperative launch per direction.
//
// forward : phase 1  every CTA sums its pixel chunk of U (all K*R*c channels) -> atomics into raw[n][K*R*c]
//           grid.sync
//           phase 2  every CTA recomputes the (tiny) FC chain of its image in shared memory
//                    (gap -> dense1 -> BN -> act -> dense2 x R -> softmax over channels / sigmoid) and recombines
//                    V = sum_r U_r * a_r over the SAME pixel chunk, which is still L2-resident: U crosses HBM once.
// backward: phase 1  da[n][k][r][c] = sum_pixels dV * U_r   (same chunking) ; grid.sync
//           phase 2  per-image FC backward in shared memory (softmax/sigmoid bwd, dense2^T, act', BN, dense1^T),
//                    then dU = (dV * a_r + dgap / HW) * act'(U) over the chunk.  The chunk-0 CTA of each image also
//                    leaves dz / dbn / xhat in the scratch buffer for the parameter-gradient kernel (reduction over n).
// A CTA keeps the same chunk in both phases, so the second read of U (and dV) hits L2 (126 MB) for every stage of the
// network.  Grid size is bounded by co-residency (cooperative launch).
#include "tbi_common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace {

template <typename T, int V> struct alignas(sizeof(T) * V) PackF { T v[V]; };
template <typename T, int V> __device__ __forceinline__ void ldp(const T* p, float (&f)[V]) {
    PackF<T, V> q = *reinterpret_cast<const PackF<T, V>*>(p);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = ldf(&q.v[i]);
}
template <typename T, int V> __device__ __forceinline__ void stp(T* p, const float (&f)[V]) {
    PackF<T, V> q;
#pragma unroll
    for (int i = 0; i < V; ++i) stf(&q.v[i], f[i]);
    *reinterpret_cast<PackF<T, V>*>(p) = q;
}

__device__ __forceinline__ float blk_reduce(float v, float* red, bool is_max) {
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

// phase 1 of both directions: per-channel sums over this CTA's pixel chunk (optionally times dv)
template <typename T, int V, bool MUL>
__device__ __forceinline__ void chunk_reduce(const tbi_splitatt& p, const tbi_view& u, const tbi_view& dv, float* raw, float* sm,
                                             int n, int pbeg, int pend) {
    const int hw = p.h * p.w, C = u.c, cv = C / V, R = p.radix, c = p.c;
    const int cl = min(cv, (int)blockDim.x), pl = blockDim.x / cl;
    const int lane_c = threadIdx.x % cl, lane_p = threadIdx.x / cl;
    const T* ub = (const T*)u.ptr + (size_t)n * hw * u.cstride + u.coff;
    const T* db = MUL ? (const T*)dv.ptr + (size_t)n * hw * dv.cstride + dv.coff : nullptr;
    for (int cv0 = 0; cv0 < cv; cv0 += cl) {
        const int ch = (cv0 + lane_c) * V;
        float s[V];
#pragma unroll
        for (int k = 0; k < V; ++k) s[k] = 0.f;
        if (ch < C && lane_p < pl) {
            const int kk = ch / (R * c), cc = ch % c;
#pragma unroll 4
            for (int px = pbeg + lane_p; px < pend; px += pl) {
                float a[V];
                ldp<T, V>(ub + (size_t)px * u.cstride + ch, a);
                if (MUL) {
                    float g[V];
                    ldp<T, V>(db + (size_t)px * dv.cstride + kk * c + cc, g);
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

// shared-memory layout of the per-image FC state (floats): g[K*c] | h1[K*c2] | att[K*R*c] | red[32] | tmp[...]
template <typename T, int V>
__global__ void __launch_bounds__(256) splitatt_fwd_fused_kernel(tbi_splitatt p, tbi_view u, tbi_view v, float* raw, int bpi, int ppb, int stage_w, int wofs) {
    extern __shared__ float sm[];
    cg::grid_group grid = cg::this_grid();
    const int hw = p.h * p.w, c = p.c, c2 = p.c / 2, R = p.radix, K = p.kpaths, C = u.c;
    const int n = blockIdx.x / bpi, chunk = blockIdx.x % bpi;
    const int pbeg = chunk * ppb, pend = min(hw, pbeg + ppb);
    // FC weights -> shared memory now (the loads overlap phase 1), so the FC chain between the phases never waits on L2
    float* sw1 = sm + wofs; float* sw2 = sw1 + K * c * c2;
    if (stage_w) {
        for (int i = threadIdx.x; i < K * c * c2; i += blockDim.x) sw1[i] = __ldg(p.w1 + i);
        for (int i = threadIdx.x; i < K * R * c2 * c; i += blockDim.x) sw2[i] = __ldg(p.w2 + i);
    }
    chunk_reduce<T, V, false>(p, u, v, raw, sm, n, pbeg, pend);
    __threadfence();
    grid.sync();
    // ---- FC chain for image n (all K cardinals) in shared memory
    float* g = sm; float* h1 = g + K * c; float* att = h1 + K * c2; float* red = att + K * R * c;
    const float inv_hw = 1.f / (float)hw;
    const float* rw = raw + (size_t)n * C;
    for (int i = threadIdx.x; i < K * c; i += blockDim.x) {
        const int k = i / c, ch = i % c;
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += __ldcg(rw + (k * R + r) * c + ch);
        g[i] = s * inv_hw;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * c2; i += blockDim.x) {
        const int k = i / c2, j = i % c2;
        float q = p.b1[i];
        const float* w1 = (stage_w ? sw1 : p.w1) + (size_t)k * c * c2 + j;
#pragma unroll 32
        for (int ch = 0; ch < c; ++ch) q = fmaf(g[k * c + ch], w1[(size_t)ch * c2], q);
        const float sc = p.gamma[i] * rsqrtf(p.var[i] + p.bn_eps);
        h1[i] = act_apply(p.act, (q - p.mean[i]) * sc + p.beta[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * R * c; i += blockDim.x) {
        const int ch = i % c, kr = i / c, k = kr / R;
        float z = p.b2[i];
        const float* w2 = (stage_w ? sw2 : p.w2) + (size_t)kr * c2 * c + ch;
#pragma unroll 32
        for (int j = 0; j < c2; ++j) z = fmaf(h1[k * c2 + j], w2[(size_t)j * c], z);
        att[i] = z;
    }
    __syncthreads();
    for (int kr = 0; kr < K * R; ++kr) {                    // softmax over the channel axis (reference quirk) / sigmoid
        float* a = att + kr * c;
        if (R == 1) {
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) a[ch] = 1.f / (1.f + expf(-a[ch]));
        } else {
            float lm = -INFINITY;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) lm = fmaxf(lm, a[ch]);
            const float mx = blk_reduce(lm, red, true);
            float ls = 0.f;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) { const float e = expf(a[ch] - mx); a[ch] = e; ls += e; }
            const float inv = 1.f / blk_reduce(ls, red, false);
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) a[ch] *= inv;
        }
        __syncthreads();
    }
    if (chunk == 0) {                                       // keep the per-image state for the backward pass
        for (int i = threadIdx.x; i < K * c; i += blockDim.x) p.gap[(size_t)n * K * c + i] = g[i];
        for (int i = threadIdx.x; i < K * c2; i += blockDim.x) p.h1[(size_t)n * K * c2 + i] = h1[i];
        for (int i = threadIdx.x; i < K * R * c; i += blockDim.x) p.att[(size_t)n * K * R * c + i] = att[i];
    }
    // ---- recombine over the same chunk (U is L2-resident)
    const int cvv = (K * c) / V;
    const T* ub = (const T*)u.ptr + (size_t)n * hw * u.cstride + u.coff;
    T* vb = (T*)v.ptr + (size_t)n * hw * v.cstride + v.coff;
    const int npx = pend - pbeg;
    for (int i = threadIdx.x; i < npx * cvv; i += blockDim.x) {
        const int co = (i % cvv) * V, px = pbeg + i / cvv;
        const int k = co / c, cc = co % c;
        float o[V];
#pragma unroll
        for (int q = 0; q < V; ++q) o[q] = 0.f;
        const float* a = att + (size_t)k * R * c + cc;
        const T* up = ub + (size_t)px * u.cstride + (size_t)k * R * c + cc;
        for (int r = 0; r < R; ++r) {
            float x[V];
            ldp<T, V>(up + (size_t)r * c, x);
#pragma unroll
            for (int q = 0; q < V; ++q) o[q] = fmaf(x[q], a[r * c + q], o[q]);
        }
        stp<T, V>(vb + (size_t)px * v.cstride + co, o);
    }
}

// scratch layout (as tbi_split_attention_bwd): dz [n][K][R][c] | dgap [n][K][c] | dbn [n][K][c2] | xhat [n][K][c2]
template <typename T, int V>
__global__ void __launch_bounds__(256) splitatt_bwd_fused_kernel(tbi_splitatt p, tbi_view u, tbi_view dv, tbi_view du, float* scratch, int bpi, int ppb,
                                                                 int stage_w, int wofs) {
    extern __shared__ float sm[];
    cg::grid_group grid = cg::this_grid();
    const int hw = p.h * p.w, c = p.c, c2 = p.c / 2, R = p.radix, K = p.kpaths, C = u.c, N = p.n;
    const int n = blockIdx.x / bpi, chunk = blockIdx.x % bpi;
    const int pbeg = chunk * ppb, pend = min(hw, pbeg + ppb);
    float* sw1 = sm + wofs; float* sw2 = sw1 + K * c * c2;
    if (stage_w) {
        for (int i = threadIdx.x; i < K * c * c2; i += blockDim.x) sw1[i] = __ldg(p.w1 + i);
        for (int i = threadIdx.x; i < K * R * c2 * c; i += blockDim.x) sw2[i] = __ldg(p.w2 + i);
    }
    const float* W1 = stage_w ? sw1 : p.w1; const float* W2 = stage_w ? sw2 : p.w2;
    chunk_reduce<T, V, true>(p, u, dv, scratch, sm, n, pbeg, pend);          // da accumulates in the dz region
    __threadfence();
    grid.sync();
    float* att = sm; float* dz = att + K * R * c; float* dq = dz + K * R * c; float* dg = dq + K * c2; float* red = dg + K * c;
    const float* gatt = p.att + (size_t)n * K * R * c;
    const float* gg = p.gap + (size_t)n * K * c;
    const float* gh1 = p.h1 + (size_t)n * K * c2;
    const float* da = scratch + (size_t)n * C;
    for (int i = threadIdx.x; i < K * R * c; i += blockDim.x) { att[i] = gatt[i]; dz[i] = __ldcg(da + i); }
    __syncthreads();
    for (int kr = 0; kr < K * R; ++kr) {
        float* a = att + kr * c; float* z = dz + kr * c;
        if (R == 1) {
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) z[ch] = a[ch] * (1.f - a[ch]) * z[ch];
        } else {
            float l = 0.f;
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) l += a[ch] * z[ch];
            const float dot = blk_reduce(l, red, false);
            for (int ch = threadIdx.x; ch < c; ch += blockDim.x) z[ch] = a[ch] * (z[ch] - dot);
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll 4
    for (int i = wid; i < K * c2; i += nwarp) {              // dh1 = sum_r dz_r W2_r^T : a warp per output, lanes over channels
        const int k = i / c2, j = i % c2;
        float s = 0.f;
        for (int r = 0; r < R; ++r) {
            const float* w2 = W2 + (((size_t)k * R + r) * c2 + j) * c;
            for (int ch = lane; ch < c; ch += 32) s = fmaf(dz[(k * R + r) * c + ch], w2[ch], s);
        }
        s = warp_sum(s);
        if (lane == 0) dq[i] = s;
    }
    __syncthreads();
    float* sdbn = scratch + (size_t)N * K * R * c + (size_t)N * K * c + (size_t)n * K * c2;
    float* sxh = sdbn + (size_t)N * K * c2;
    for (int i = threadIdx.x; i < K * c2; i += blockDim.x) {
        const int k = i / c2, j = i % c2;
        float q = p.b1[i];
        const float* w1 = W1 + (size_t)k * c * c2 + j;
#pragma unroll 32
        for (int ch = 0; ch < c; ++ch) q = fmaf(gg[k * c + ch], w1[(size_t)ch * c2], q);
        const float istd = rsqrtf(p.var[i] + p.bn_eps);
        const float d = dq[i] * act_grad_from_out(p.act, gh1[i]);
        if (chunk == 0) { sdbn[i] = d; sxh[i] = (q - p.mean[i]) * istd; }
        dq[i] = d * p.gamma[i] * istd;
    }
    __syncthreads();
#pragma unroll 4
    for (int i = wid; i < K * c; i += nwarp) {
        const int k = i / c, ch = i % c;
        float s = 0.f;
        const float* w1 = W1 + ((size_t)k * c + ch) * c2;
        for (int j = lane; j < c2; j += 32) s = fmaf(dq[k * c2 + j], w1[j], s);
        s = warp_sum(s);
        if (lane == 0) dg[i] = s;
    }
    __syncthreads();
    if (chunk == 0) {
        float* sdz = scratch + (size_t)n * C;                // overwrite da with dz for the parameter-gradient kernel
        float* sdg = scratch + (size_t)N * K * R * c + (size_t)n * K * c;
        // every CTA of the image has already copied da into shared memory: the grid.sync below orders this write after them
        for (int i = threadIdx.x; i < K * c; i += blockDim.x) sdg[i] = dg[i];
        (void)sdz;
    }
    grid.sync();
    if (chunk == 0) {
        float* sdz = scratch + (size_t)n * C;
        for (int i = threadIdx.x; i < K * R * c; i += blockDim.x) sdz[i] = dz[i];
    }
    // ---- dU over the same chunk
    const float inv_hw = 1.f / (float)hw;
    const int cvv = (K * c) / V;
    const T* ub = (const T*)u.ptr + (size_t)n * hw * u.cstride + u.coff;
    const T* db = (const T*)dv.ptr + (size_t)n * hw * dv.cstride + dv.coff;
    T* dub = (T*)du.ptr + (size_t)n * hw * du.cstride + du.coff;
    const int npx = pend - pbeg;
    for (int i = threadIdx.x; i < npx * cvv; i += blockDim.x) {
        const int co = (i % cvv) * V, px = pbeg + i / cvv;
        const int k = co / c, cc = co % c;
        float gv[V];
        ldp<T, V>(db + (size_t)px * dv.cstride + co, gv);
        const float* a = att + (size_t)k * R * c + cc;
        const float* dgp = dg + k * c + cc;
        const size_t uo = (size_t)k * R * c + cc;
        for (int r = 0; r < R; ++r) {
            float x[V], o[V];
            ldp<T, V>(ub + (size_t)px * u.cstride + uo + (size_t)r * c, x);
#pragma unroll
            for (int q = 0; q < V; ++q) o[q] = (gv[q] * a[r * c + q] + dgp[q] * inv_hw) * act_grad_from_out(p.act, x[q]);
            stp<T, V>(dub + (size_t)px * du.cstride + uo + (size_t)r * c, o);
        }
    }
}

struct FusedPlan { int bpi, ppb, grid, stage_w, wofs; size_t smem; };

template <typename KernelT>
bool plan_fused(const tbi_splitatt* p, KernelT kernel, size_t fc_floats, int V, FusedPlan* fp) {
    const int hw = p->h * p->w;
    const size_t base_floats = fc_floats > (size_t)256 * V ? fc_floats : (size_t)256 * V;
    const size_t w_floats = (size_t)p->kpaths * p->c * (p->c / 2) * (1 + p->radix);
    fp->stage_w = w_floats * sizeof(float) <= 48 * 1024 ? 1 : 0;
    fp->wofs = (int)((base_floats + 3) & ~(size_t)3);
    const size_t smem = sizeof(float) * (fp->wofs + (fp->stage_w ? w_floats : 0)) + 256;
    if (smem > 160 * 1024) return false;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return false; }
    const int cap = per_sm * tbi_sm_count();
    if (cap < p->n) return false;
    int bpi = cap / p->n;
    const int max_bpi = (hw + 63) / 64;
    if (bpi > max_bpi) bpi = max_bpi;
    static const int bpi_cap = getenv("TBI_SA_BPI") ? atoi(getenv("TBI_SA_BPI")) : 32;
    if (bpi > bpi_cap) bpi = bpi_cap;
    if (bpi < 1) bpi = 1;
    fp->ppb = (hw + bpi - 1) / bpi;
    fp->bpi = (hw + fp->ppb - 1) / fp->ppb;
    fp->grid = fp->bpi * p->n;
    fp->smem = smem;
    return true;
}

}  // namespace

// returns 1 if launched, 0 if the fused path does not apply (caller falls back to the multi-kernel path), <0 on error
int tbi_splitatt_fwd_fused(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, cudaStream_t s) {
    if (p->dtype != TBI_BF16) return 0;
    static const bool disabled = getenv("TBI_NO_FUSED_SPLITATT") != nullptr;
    if (disabled) return 0;
    const int V = 8;
    if (p->c % V != 0 || u->cstride % V || u->coff % V || v->cstride % V || v->coff % V || ((uintptr_t)u->ptr & 15) || ((uintptr_t)v->ptr & 15)) return 0;
    const int K = p->kpaths, R = p->radix, c = p->c;
    FusedPlan fp;
    auto kernel = splitatt_fwd_fused_kernel<__nv_bfloat16, 8>;
    if (!plan_fused(p, kernel, (size_t)K * c + K * (c / 2) + (size_t)K * R * c + 32, V, &fp)) return 0;
    cudaError_t e = cudaMemsetAsync(p->att, 0, sizeof(float) * (size_t)p->n * u->c, s);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt fused memset: %s", cudaGetErrorString(e));
    tbi_splitatt pp = *p; tbi_view uu = *u, vv = *v; float* raw = p->att; int bpi = fp.bpi, ppb = fp.ppb, stw = fp.stage_w, wofs = fp.wofs;
    void* args[] = {&pp, &uu, &vv, &raw, &bpi, &ppb, &stw, &wofs};
    e = cudaLaunchCooperativeKernel((void*)kernel, dim3(fp.grid), dim3(256), args, fp.smem, s);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt fused fwd launch: %s", cudaGetErrorString(e));
    return 1;
}

int tbi_splitatt_bwd_fused(const tbi_splitatt* p, const tbi_view* u, const tbi_view* dv, const tbi_view* du, float* scratch, cudaStream_t s) {
    if (p->dtype != TBI_BF16) return 0;
    static const bool disabled = getenv("TBI_NO_FUSED_SPLITATT") != nullptr;
    if (disabled) return 0;
    const int V = 8;
    if (p->c % V != 0 || u->cstride % V || u->coff % V || dv->cstride % V || dv->coff % V || du->cstride % V || du->coff % V ||
        ((uintptr_t)u->ptr & 15) || ((uintptr_t)dv->ptr & 15) || ((uintptr_t)du->ptr & 15)) return 0;
    const int K = p->kpaths, R = p->radix, c = p->c;
    FusedPlan fp;
    auto kernel = splitatt_bwd_fused_kernel<__nv_bfloat16, 8>;
    if (!plan_fused(p, kernel, (size_t)2 * K * R * c + K * (c / 2) + (size_t)K * c + 32, V, &fp)) return 0;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)p->n * u->c, s);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt fused memset: %s", cudaGetErrorString(e));
    tbi_splitatt pp = *p; tbi_view uu = *u, dd = *dv, du2 = *du; int bpi = fp.bpi, ppb = fp.ppb, stw = fp.stage_w, wofs = fp.wofs;
    void* args[] = {&pp, &uu, &dd, &du2, &scratch, &bpi, &ppb, &stw, &wofs};
    e = cudaLaunchCooperativeKernel((void*)kernel, dim3(fp.grid), dim3(256), args, fp.smem, s);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "splitatt fused bwd launch: %s", cudaGetErrorString(e));
    return 1;
}
