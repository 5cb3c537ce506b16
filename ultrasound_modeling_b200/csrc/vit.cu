// The ViT bridge of Variant B (reference: VisionTransformer.py:9-190): the pieces that are not convolutions.
//   * multi-head self-attention core for short sequences (80 tokens, 4 heads x 128): ONE CTA per (image, head) keeps Q, K, V
//     and the T x T score matrix in shared memory; forward = scores -> softmax -> context, backward = dV, dP, softmax', dQ, dK.
//     The reference divides the scores by sqrt(num_heads), not sqrt(head_dim) (VisionTransformer.py:42): the scale is an argument.
//   * exact (erf) GELU forward / backward (tf.keras.activations.gelu default, VisionTransformer.py:70)
//   * softmax + CategoricalCrossentropy(label_smoothing, reduction NONE) + compute_average_loss (VisionTransformer.py:205-206,
//     225-227): scalar loss = sum over all pixels / global batch, and its gradient w.r.t. the logits.
// The dense layers around them are 1x1 tap-GEMMs (tbi_conv2d_*), the LayerNorms tbi_layernorm_c_*: all of it is
// bandwidth / latency bound at 80 tokens (4.2 GFLOP per image against 5.1 for the convolutional encoder+decoder).
#include "tbi_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ void load_rows(float* dst, int pitch, const T* src, int rows, int d, int src_stride) {
    for (int i = threadIdx.x; i < rows * d; i += blockDim.x) { const int r = i / d, c = i - r * d; dst[r * pitch + c] = ldf(src + (size_t)r * src_stride + c); }
}

// q, k, v, ctx: [n, T, heads*d] (head h = channels [h*d, (h+1)*d)); probs: fp32 [n, heads, T, T]
template <typename T>
__global__ void __launch_bounds__(256) attention_fwd_kernel(int T_, int heads, int d, float scale, const T* __restrict__ q, const T* __restrict__ k,
                                                            const T* __restrict__ v, T* __restrict__ ctx, float* __restrict__ probs) {
    extern __shared__ float sm[];
    const int pitch = d + 1;                                   // +1: column walks of K hit distinct banks
    float* Q = sm; float* K = Q + T_ * pitch; float* V = K + T_ * pitch; float* S = V + T_ * pitch;     // S: [T][T+1]
    const int n = blockIdx.x / heads, h = blockIdx.x % heads, C = heads * d, sp = T_ + 1;
    const size_t base = (size_t)n * T_ * C + (size_t)h * d;
    load_rows(Q, pitch, q + base, T_, d, C); load_rows(K, pitch, k + base, T_, d, C); load_rows(V, pitch, v + base, T_, d, C);
    __syncthreads();
    for (int e = threadIdx.x; e < T_ * T_; e += blockDim.x) {
        const int i = e / T_, j = e - i * T_;
        float s = 0.f;
        for (int c = 0; c < d; ++c) s = fmaf(Q[i * pitch + c], K[j * pitch + c], s);
        S[i * sp + j] = s * scale;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < T_; i += nw) {                        // softmax over the keys (axis 3 of [n, heads, T, T])
        float m = -INFINITY;
        for (int j = lane; j < T_; j += 32) m = fmaxf(m, S[i * sp + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float z = 0.f;
        for (int j = lane; j < T_; j += 32) { const float p = __expf(S[i * sp + j] - m); S[i * sp + j] = p; z += p; }
        z = warp_sum(z);
        const float inv = 1.f / z;
        float* pr = probs + (((size_t)n * heads + h) * T_ + i) * T_;
        for (int j = lane; j < T_; j += 32) { const float p = S[i * sp + j] * inv; S[i * sp + j] = p; pr[j] = p; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < T_ * d; e += blockDim.x) {
        const int i = e / d, c = e - i * d;
        float o = 0.f;
        for (int j = 0; j < T_; ++j) o = fmaf(S[i * sp + j], V[j * pitch + c], o);
        stf(ctx + base + (size_t)i * C + c, o);
    }
}

// dctx -> dq, dk, dv given q, k, v and the forward's probabilities
template <typename T>
__global__ void __launch_bounds__(256) attention_bwd_kernel(int T_, int heads, int d, float scale, const T* __restrict__ q, const T* __restrict__ k,
                                                            const T* __restrict__ v, const float* __restrict__ probs, const T* __restrict__ dctx,
                                                            T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv) {
    extern __shared__ float sm[];
    const int pitch = d + 1, sp = T_ + 1;
    float* Q = sm; float* K = Q + T_ * pitch; float* V = K + T_ * pitch; float* dO = V + T_ * pitch; float* P = dO + T_ * pitch; float* dS = P + T_ * sp;
    const int n = blockIdx.x / heads, h = blockIdx.x % heads, C = heads * d;
    const size_t base = (size_t)n * T_ * C + (size_t)h * d;
    load_rows(Q, pitch, q + base, T_, d, C); load_rows(K, pitch, k + base, T_, d, C); load_rows(V, pitch, v + base, T_, d, C);
    load_rows(dO, pitch, dctx + base, T_, d, C);
    const float* pr = probs + ((size_t)n * heads + h) * T_ * T_;
    for (int e = threadIdx.x; e < T_ * T_; e += blockDim.x) P[(e / T_) * sp + (e % T_)] = pr[e];
    __syncthreads();
    // dV[j][c] = sum_i P[i][j] dO[i][c]
    for (int e = threadIdx.x; e < T_ * d; e += blockDim.x) {
        const int j = e / d, c = e - j * d;
        float a = 0.f;
        for (int i = 0; i < T_; ++i) a = fmaf(P[i * sp + j], dO[i * pitch + c], a);
        stf(dv + base + (size_t)j * C + c, a);
    }
    // dP[i][j] = sum_c dO[i][c] V[j][c]
    for (int e = threadIdx.x; e < T_ * T_; e += blockDim.x) {
        const int i = e / T_, j = e - i * T_;
        float a = 0.f;
        for (int c = 0; c < d; ++c) a = fmaf(dO[i * pitch + c], V[j * pitch + c], a);
        dS[i * sp + j] = a;
    }
    __syncthreads();
    // softmax backward per row, times the score scale: dS = P * (dP - sum_j dP*P) * scale
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < T_; i += nw) {
        float dot = 0.f;
        for (int j = lane; j < T_; j += 32) dot = fmaf(dS[i * sp + j], P[i * sp + j], dot);
        dot = warp_sum(dot);
        for (int j = lane; j < T_; j += 32) dS[i * sp + j] = P[i * sp + j] * (dS[i * sp + j] - dot) * scale;
    }
    __syncthreads();
    // dQ[i][c] = sum_j dS[i][j] K[j][c] ;  dK[j][c] = sum_i dS[i][j] Q[i][c]
    for (int e = threadIdx.x; e < T_ * d; e += blockDim.x) {
        const int i = e / d, c = e - i * d;
        float a = 0.f, b = 0.f;
        for (int j = 0; j < T_; ++j) { a = fmaf(dS[i * sp + j], K[j * pitch + c], a); b = fmaf(dS[j * sp + i], Q[j * pitch + c], b); }
        stf(dq + base + (size_t)i * C + c, a);
        stf(dk + base + (size_t)i * C + c, b);
    }
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
template <typename T>
__global__ void gelu_fwd_kernel(long long count, const T* __restrict__ x, T* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) stf(y + i, gelu_f(ldf(x + i)));
}
template <typename T>
__global__ void gelu_bwd_kernel(long long count, const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        stf(dx + i, ldf(dy + i) * gelu_grad_f(ldf(x + i)));
}

// one thread per pixel: p = softmax(z); Keras CategoricalCrossentropy on probabilities: y_s = y*(1-ls) + ls/nc,
// p_n = p / sum(p), clipped to [1e-7, 1 - 1e-7]; loss_pixel = -sum_c y_s log(p_n); total = sum_pixels / global_batch.
// dL/dz_j = p_j (g_j - sum_c p_c g_c) with g_c = -y_s,c / p_c inside the clip range, 0 outside (tf.clip_by_value's gradient).
template <int NC>
__global__ void __launch_bounds__(256) softmax_cce_kernel(long long npix, float smoothing, float inv_batch, const float* __restrict__ logits,
                                                          const float* __restrict__ y, float* __restrict__ probs, float* loss_sum,
                                                          float* __restrict__ dlogits) {
    __shared__ float red[32];
    float local = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        float z[NC], p[NC], ys[NC], g[NC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) { z[c] = logits[i * NC + c]; m = fmaxf(m, z[c]); }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) { p[c] = expf(z[c] - m); s += p[c]; }
        const float inv = 1.f / s;
        float dot = 0.f, lp = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            p[c] *= inv;
            probs[i * NC + c] = p[c];
            ys[c] = y[i * NC + c] * (1.f - smoothing) + smoothing / (float)NC;
            const bool inside = p[c] >= 1e-7f && p[c] <= 1.f - 1e-7f;
            const float pc = fminf(fmaxf(p[c], 1e-7f), 1.f - 1e-7f);
            lp -= ys[c] * logf(pc);
            g[c] = inside ? -ys[c] / p[c] : 0.f;
            dot = fmaf(p[c], g[c], dot);
        }
        local += lp;
        if (dlogits) {
#pragma unroll
            for (int c = 0; c < NC; ++c) dlogits[i * NC + c] = p[c] * (g[c] - dot) * inv_batch;
        }
    }
    float v = warp_sum(local);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) atomicAdd(loss_sum, v * inv_batch);
    }
}

unsigned ew_grid(long long work) {
    long long b = (work + 255) / 256;
    const long long cap = (long long)tbi_sm_count() * 16;
    return (unsigned)(b < 1 ? 1 : b > cap ? cap : b);
}

}  // namespace

static size_t attn_smem(int T_, int d, bool bwd) {
    return ((size_t)(bwd ? 4 : 3) * T_ * (d + 1) + (size_t)(bwd ? 2 : 1) * T_ * (T_ + 1)) * sizeof(float);
}

extern "C" int tbi_attention_fwd(int dtype, int n, int tokens, int heads, int head_dim, float scale, const void* q, const void* k, const void* v,
                                 void* ctx, float* probs, void* stream) {
    TBI_CHECK(q && k && v && ctx && probs && n > 0 && tokens > 0 && heads > 0 && head_dim > 0, TBI_ERR_BAD_SHAPE, "attention_fwd: null / empty argument");
    const size_t smem = attn_smem(tokens, head_dim, false);
    TBI_CHECK(smem <= 220 * 1024, TBI_ERR_UNSUPPORTED, "attention_fwd: %d tokens x %d channels per head needs %zu bytes of shared memory (the fused kernel "
              "holds Q, K, V and the score matrix of one head on chip; the reference's sequences are 80 tokens)", tokens, head_dim, smem);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (dtype == TBI_F32) {
        e = cudaFuncSetAttribute(attention_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess) attention_fwd_kernel<float><<<n * heads, 256, smem, s>>>(tokens, heads, head_dim, scale, (const float*)q, (const float*)k, (const float*)v, (float*)ctx, probs);
    } else if (dtype == TBI_BF16) {
        e = cudaFuncSetAttribute(attention_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess) attention_fwd_kernel<__nv_bfloat16><<<n * heads, 256, smem, s>>>(tokens, heads, head_dim, scale, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                                                               (const __nv_bfloat16*)v, (__nv_bfloat16*)ctx, probs);
    } else return tbi_set_error(TBI_ERR_UNSUPPORTED, "attention_fwd dtype");
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "attention_fwd: %s", cudaGetErrorString(e));
    TBI_CUDA_LAUNCH_CHECK("attention_fwd");
    return TBI_OK;
}

extern "C" int tbi_attention_bwd(int dtype, int n, int tokens, int heads, int head_dim, float scale, const void* q, const void* k, const void* v,
                                 const float* probs, const void* dctx, void* dq, void* dk, void* dv, void* stream) {
    TBI_CHECK(q && k && v && probs && dctx && dq && dk && dv && n > 0, TBI_ERR_BAD_SHAPE, "attention_bwd: null / empty argument");
    const size_t smem = attn_smem(tokens, head_dim, true);
    TBI_CHECK(smem <= 220 * 1024, TBI_ERR_UNSUPPORTED, "attention_bwd: %d tokens x %d channels per head needs %zu bytes of shared memory", tokens, head_dim, smem);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (dtype == TBI_F32) {
        e = cudaFuncSetAttribute(attention_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess) attention_bwd_kernel<float><<<n * heads, 256, smem, s>>>(tokens, heads, head_dim, scale, (const float*)q, (const float*)k, (const float*)v, probs,
                                                                                       (const float*)dctx, (float*)dq, (float*)dk, (float*)dv);
    } else if (dtype == TBI_BF16) {
        typedef __nv_bfloat16 B;
        e = cudaFuncSetAttribute(attention_bwd_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e == cudaSuccess) attention_bwd_kernel<B><<<n * heads, 256, smem, s>>>(tokens, heads, head_dim, scale, (const B*)q, (const B*)k, (const B*)v, probs, (const B*)dctx,
                                                                                   (B*)dq, (B*)dk, (B*)dv);
    } else return tbi_set_error(TBI_ERR_UNSUPPORTED, "attention_bwd dtype");
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "attention_bwd: %s", cudaGetErrorString(e));
    TBI_CUDA_LAUNCH_CHECK("attention_bwd");
    return TBI_OK;
}

extern "C" int tbi_gelu_fwd(int dtype, int64_t count, const void* x, void* y, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TBI_F32) gelu_fwd_kernel<float><<<ew_grid(count), 256, 0, s>>>(count, (const float*)x, (float*)y);
    else if (dtype == TBI_BF16) gelu_fwd_kernel<__nv_bfloat16><<<ew_grid(count), 256, 0, s>>>(count, (const __nv_bfloat16*)x, (__nv_bfloat16*)y);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "gelu_fwd dtype");
    TBI_CUDA_LAUNCH_CHECK("gelu_fwd");
    return TBI_OK;
}

extern "C" int tbi_gelu_bwd(int dtype, int64_t count, const void* x, const void* dy, void* dx, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TBI_F32) gelu_bwd_kernel<float><<<ew_grid(count), 256, 0, s>>>(count, (const float*)x, (const float*)dy, (float*)dx);
    else if (dtype == TBI_BF16) gelu_bwd_kernel<__nv_bfloat16><<<ew_grid(count), 256, 0, s>>>(count, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx);
    else return tbi_set_error(TBI_ERR_UNSUPPORTED, "gelu_bwd dtype");
    TBI_CUDA_LAUNCH_CHECK("gelu_bwd");
    return TBI_OK;
}

extern "C" int tbi_softmax_cce_fwd_bwd(int64_t npix, int nc, float label_smoothing, float global_batch, const float* logits, const float* y,
                                       float* probs, float* loss_sum, float* dlogits, void* stream) {
    TBI_CHECK(logits && y && probs && loss_sum && npix > 0 && global_batch > 0.f, TBI_ERR_BAD_SHAPE, "softmax_cce: null / empty argument");
    cudaStream_t s = (cudaStream_t)stream;
    const float inv = 1.f / global_batch;
    const unsigned g = ew_grid(npix);
    switch (nc) {
        case 2: softmax_cce_kernel<2><<<g, 256, 0, s>>>(npix, label_smoothing, inv, logits, y, probs, loss_sum, dlogits); break;
        case 3: softmax_cce_kernel<3><<<g, 256, 0, s>>>(npix, label_smoothing, inv, logits, y, probs, loss_sum, dlogits); break;
        case 4: softmax_cce_kernel<4><<<g, 256, 0, s>>>(npix, label_smoothing, inv, logits, y, probs, loss_sum, dlogits); break;
        default: return tbi_set_error(TBI_ERR_UNSUPPORTED, "softmax_cce: %d classes (2..4)", nc);
    }
    TBI_CUDA_LAUNCH_CHECK("softmax_cce");
    return TBI_OK;
}
