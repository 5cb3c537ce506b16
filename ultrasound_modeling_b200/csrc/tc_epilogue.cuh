// Epilogue shared by the tcgen05 tap-GEMM kernels: 8 consecutive output channels per call, 16-byte accesses.
#pragma once
#include "tbi_common.cuh"

namespace {

// ELU for bf16 outputs: exp via the SFU (abs error ~1e-7, far below bf16 resolution)
__device__ __forceinline__ float act_apply_fast(int act, float v) {
    switch (act) {
        case TBI_ACT_ELU:   return v > 0.f ? v : __expf(v) - 1.f;
        case TBI_ACT_LRELU: return v > 0.f ? v : 0.3f * v;
        case TBI_ACT_RELU:  return v > 0.f ? v : 0.f;
        default:            return v;
    }
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return q;
}

// same math as epilogue_store<bf16> (tbi_common.cuh) on 8 consecutive channels with 16-byte accesses
__device__ __forceinline__ void epilogue_store8(const tbi_epilogue& e, int n, int oy, int ox, int co, float (&v)[8]) {
    typedef __nv_bfloat16 T;
    if (e.split_c > 0 && co >= e.split_c) {
        const int c2 = co - e.split_c;
        if (e.residual2.ptr) {
            float r[8]; unpack8(*reinterpret_cast<const uint4*>((const T*)e.residual2.ptr + view_off(e.residual2, n, oy, ox, c2)), r);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += r[i];
        }
        *reinterpret_cast<uint4*>((T*)e.out2.ptr + view_off(e.out2, n, oy, ox, c2)) = pack8(v);
        return;
    }
    if (e.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + co)), b1 = __ldg(reinterpret_cast<const float4*>(e.bias + co + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (e.drop_keep) {
        const uint2 k = *reinterpret_cast<const uint2*>(e.drop_keep + (((size_t)n * e.out.h + oy) * e.out.w + ox) * e.out.c + co);
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_apply_fast(e.act, v[i]);
    if (e.residual.ptr) {
        float r[8]; unpack8(*reinterpret_cast<const uint4*>((const T*)e.residual.ptr + view_off(e.residual, n, oy, ox, co)), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    if (e.dact != TBI_ACT_NONE) {
        float r[8]; unpack8(*reinterpret_cast<const uint4*>((const T*)e.dact_ref.ptr + view_off(e.dact_ref, n, oy, ox, co)), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= act_grad_from_out(e.dact, r[i]);
        if (e.dact_keep) {
            const uint2 k = *reinterpret_cast<const uint2*>(e.dact_keep + (((size_t)n * e.dact_ref.h + oy) * e.dact_ref.w + ox) * e.dact_ref.c + co);
            const unsigned char* kb = reinterpret_cast<const unsigned char*>(&k);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
        }
    }
    *reinterpret_cast<uint4*>((T*)e.out.ptr + view_off(e.out, n, oy, ox, co)) = pack8(v);
}


}  // namespace
