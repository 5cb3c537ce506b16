// Shared device/host helpers for libtbi_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/tbi_sm100.h"

// ---- error plumbing -------------------------------------------------------------------------
int  tbi_set_error(int code, const char* fmt, ...);
#define TBI_CHECK(cond, code, ...) do { if (!(cond)) return tbi_set_error((code), __VA_ARGS__); } while (0)
#define TBI_CUDA_LAUNCH_CHECK(what) do { cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "%s: %s", (what), cudaGetErrorString(e__)); } while (0)

static inline int tbi_dtype_size(int dtype) { return dtype == TBI_F32 ? 4 : 2; }
int tbi_sm_count();
int tbi_wgrad_gather_blocks(int ntaps, int groups, int cin_g, int cin_pad, int cout_total, const float* dense, float* dw, cudaStream_t s);

// ---- storage type <-> float -----------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- activations (Keras: ELU alpha=1, LeakyReLU slope 0.3) -----------------------------------
__device__ __forceinline__ float act_apply(int act, float v) {
    switch (act) {
        case TBI_ACT_ELU:   return v > 0.f ? v : expm1f(v);
        case TBI_ACT_LRELU: return v > 0.f ? v : 0.3f * v;
        case TBI_ACT_RELU:  return v > 0.f ? v : 0.f;
        default:            return v;
    }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
    switch (act) {
        case TBI_ACT_ELU:   return y > 0.f ? 1.f : y + 1.f;
        case TBI_ACT_LRELU: return y > 0.f ? 1.f : 0.3f;
        case TBI_ACT_RELU:  return y > 0.f ? 1.f : 0.f;
        default:            return 1.f;
    }
}

// ---- view addressing -------------------------------------------------------------------------
// element offset of (n,y,x,ch) in a view
__device__ __forceinline__ size_t view_off(const tbi_view& v, int n, int y, int x, int ch) {
    return (((size_t)n * v.h + y) * v.w + x) * (size_t)v.cstride + v.coff + ch;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Epilogue math shared by the SIMT and tcgen05 tap-GEMMs.  (n, oy, ox) is the OUTPUT pixel, co the
// output channel (global), v the fp32 accumulator.  T = storage type.
template <typename T>
__device__ __forceinline__ void epilogue_store(const tbi_epilogue& e, int n, int oy, int ox, int co, float v) {
    if (e.split_c > 0 && co >= e.split_c) {
        const int c2 = co - e.split_c;
        if (e.residual2.ptr) v += ldf((const T*)e.residual2.ptr + view_off(e.residual2, n, oy, ox, c2));
        stf((T*)e.out2.ptr + view_off(e.out2, n, oy, ox, c2), v);
        return;
    }
    if (e.bias) v += e.bias[co];
    if (e.drop_keep) v *= (float)e.drop_keep[(((size_t)n * e.out.h + oy) * e.out.w + ox) * e.out.c + co];
    v = act_apply(e.act, v);
    if (e.residual.ptr) v += ldf((const T*)e.residual.ptr + view_off(e.residual, n, oy, ox, co));
    if (e.dact != TBI_ACT_NONE) {
        const float yref = ldf((const T*)e.dact_ref.ptr + view_off(e.dact_ref, n, oy, ox, co));
        v *= act_grad_from_out(e.dact, yref);
        if (e.dact_keep) v *= (float)e.dact_keep[(((size_t)n * e.dact_ref.h + oy) * e.dact_ref.w + ox) * e.dact_ref.c + co];
    }
    if (e.out_f32) ((float*)e.out.ptr)[view_off(e.out, n, oy, ox, co)] = v;
    else stf((T*)e.out.ptr + view_off(e.out, n, oy, ox, co), v);
}

// "narrow" tensor-core epilogue: 16-column accumulator tiles written element by element through epilogue_store() (any output
// dtype).  Used when there are fewer than 8 output channels (the 3-class head) and for fp32 outputs of up to 64 columns (the
// head's forward as a per-input-pixel GEMM, engine.py) -- the vectorised epilogue writes bf16 only.
bool tbi_tapgemm_halo_supported(const tbi_tapgemm* d);          // persistent halo-addressed variant (spatial >= 16x8)
// "wide fp32" epilogue of the halo kernel: fp32 outputs written as 16-byte vectors straight from the accumulator registers
// (bias only: no activation, residual, act', dropout or split).  Used by the head's forward as a per-input-pixel GEMM
// (48 fp32 columns), where the element-wise narrow epilogue cost more than the 16x fewer MMAs saved.
static inline bool tbi_tc_f32wide(const tbi_tapgemm* d) {
    const tbi_epilogue& e = d->epi;
    return d->groups == 1 && e.out_f32 && d->cout_g >= 16 && d->cout_g % 4 == 0 && e.act == TBI_ACT_NONE && e.dact == TBI_ACT_NONE &&
           !e.residual.ptr && !e.drop_keep && e.split_c == 0 && e.out.cstride % 4 == 0 && e.out.coff % 4 == 0 &&
           ((uintptr_t)e.out.ptr & 15) == 0 && tbi_tapgemm_halo_supported(d);
}
static inline bool tbi_tc_narrow(const tbi_tapgemm* d) {
    return d->groups == 1 && (d->cout_g < 8 || (d->epi.out_f32 && d->cout_g <= 64 && !tbi_tc_f32wide(d)));
}

// entry points implemented per translation unit
int tbi_tapgemm_simt(const tbi_tapgemm* d, cudaStream_t s);
int tbi_tapwgrad_simt(const tbi_tapwgrad* d, cudaStream_t s);
int tbi_tapgemm_tc(const tbi_tapgemm* d, cudaStream_t s);       // tcgen05 implicit GEMM
int tbi_tapwgrad_tc(const tbi_tapwgrad* d, cudaStream_t s);
bool tbi_tapgemm_tc_supported(const tbi_tapgemm* d, const char** why);
bool tbi_tapwgrad_tc_supported(const tbi_tapwgrad* d, const char** why);
int64_t tbi_tapwgrad_tc_workspace(const tbi_tapwgrad* d);
struct tbi_wgrad_pair_plan { int ok; int m_is_a; int T; int m_pairs; int m_tiles; };
tbi_wgrad_pair_plan tbi_tapwgrad_pair_plan(const tbi_tapwgrad* d);                       // CTA-pair (cta_group::2) weight gradient, tapwgrad_tc2.cu
int tbi_tapwgrad_pair(const tbi_tapwgrad* d, const tbi_wgrad_pair_plan& pl, cudaStream_t s);
bool tbi_tapwgrad_small_supported(const tbi_tapwgrad* d);      // few-input-channel stride-1 conv wgrad (tap-packed M, TMEM-resident)
int tbi_tapwgrad_small(const tbi_tapwgrad* d, cudaStream_t s);
int tbi_splitatt_fwd_fused(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, cudaStream_t s);     // 1 launched, 0 not applicable
int tbi_splitatt_bwd_fused(const tbi_splitatt* p, const tbi_view* u, const tbi_view* dv, const tbi_view* du, float* scratch, cudaStream_t s);
int tbi_tapgemm_halo(const tbi_tapgemm* d, cudaStream_t s);
bool tbi_tapgemm_xpack_supported(const tbi_tapgemm* d);         // 3x3 stride-1, few channels: horizontal taps packed into N (tapgemm_xpack.cu)
int tbi_tapgemm_xpack(const tbi_tapgemm* d, cudaStream_t s);
int tbi_tapgemm_direct(const tbi_tapgemm* d, cudaStream_t s);   // few-input-channel direct conv (stem)
int tbi_tapwgrad_direct(const tbi_tapwgrad* d, cudaStream_t s);
bool tbi_tapgemm_direct_supported(const tbi_tapgemm* d);
bool tbi_tapwgrad_direct_supported(const tbi_tapwgrad* d);
