// Epilogue shared by the tcgen05 tap-GEMM kernels.  One thread owns one output pixel (one TMEM lane) and
// walks its channels 8 at a time with 16-byte accesses.  Everything that does not depend on the channel
// (row pointers, which optional terms exist) is resolved once per tile in RowCtx; the activation and the
// fused activation-derivative are template parameters so the per-element code is branch-free.
#pragma once
#include "tbi_common.cuh"

namespace {

template <int ACT> __device__ __forceinline__ float act_fast(float v) {
    if (ACT == TBI_ACT_ELU)   return v > 0.f ? v : __expf(v) - 1.f;     // SFU exp: abs error ~1e-7 << bf16 resolution
    if (ACT == TBI_ACT_LRELU) return v > 0.f ? v : 0.3f * v;
    if (ACT == TBI_ACT_RELU)  return fmaxf(v, 0.f);
    return v;
}
template <int ACT> __device__ __forceinline__ float dact_fast(float y) {
    if (ACT == TBI_ACT_ELU)   return y > 0.f ? 1.f : y + 1.f;
    if (ACT == TBI_ACT_LRELU) return y > 0.f ? 1.f : 0.3f;
    if (ACT == TBI_ACT_RELU)  return y > 0.f ? 1.f : 0.f;
    return 1.f;
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

// per-thread, per-tile pointers to channel 0 of this thread's output pixel in every tensor the epilogue touches
struct RowCtx {
    __nv_bfloat16* out; __nv_bfloat16* out2;
    const __nv_bfloat16* res; const __nv_bfloat16* res2; const __nv_bfloat16* ref;
    const uint8_t* keep; const uint8_t* dkeep;
    const float* bias;
    int split_c;
};

__device__ __forceinline__ size_t pix_off(const tbi_view& v, int n, int y, int x) {
    return (((size_t)n * v.h + y) * v.w + x) * (size_t)v.cstride + v.coff;
}

__device__ __forceinline__ RowCtx make_row_ctx(const tbi_epilogue& e, int n, int oy, int ox) {
    typedef __nv_bfloat16 T;
    RowCtx r;
    r.out = (T*)e.out.ptr + pix_off(e.out, n, oy, ox);
    r.out2 = e.split_c > 0 ? (T*)e.out2.ptr + pix_off(e.out2, n, oy, ox) : nullptr;
    r.res = e.residual.ptr ? (const T*)e.residual.ptr + pix_off(e.residual, n, oy, ox) : nullptr;
    r.res2 = e.residual2.ptr ? (const T*)e.residual2.ptr + pix_off(e.residual2, n, oy, ox) : nullptr;
    r.ref = e.dact != TBI_ACT_NONE ? (const T*)e.dact_ref.ptr + pix_off(e.dact_ref, n, oy, ox) : nullptr;
    r.keep = e.drop_keep ? e.drop_keep + (((size_t)n * e.out.h + oy) * e.out.w + ox) * e.out.c : nullptr;
    r.dkeep = e.dact_keep ? e.dact_keep + (((size_t)n * e.dact_ref.h + oy) * e.dact_ref.w + ox) * e.dact_ref.c : nullptr;
    r.bias = e.bias;
    r.split_c = e.split_c;
    return r;
}

// same math as epilogue_store<bf16> (tbi_common.cuh) on 8 consecutive channels starting at co
template <int ACT, int DACT>
__device__ __forceinline__ void epilogue_store8(const RowCtx& r, int co, float (&v)[8]) {
    if (r.split_c > 0 && co >= r.split_c) {
        const int c2 = co - r.split_c;
        if (r.res2) {
            float t[8]; unpack8(*reinterpret_cast<const uint4*>(r.res2 + c2), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
        }
        *reinterpret_cast<uint4*>(r.out2 + c2) = pack8(v);
        return;
    }
    if (r.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(r.bias + co)), b1 = __ldg(reinterpret_cast<const float4*>(r.bias + co + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (r.keep) {
        const uint2 k = *reinterpret_cast<const uint2*>(r.keep + co);
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fast<ACT>(v[i]);
    if (r.res) {
        float t[8]; unpack8(*reinterpret_cast<const uint4*>(r.res + co), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (DACT != TBI_ACT_NONE) {
        float t[8]; unpack8(*reinterpret_cast<const uint4*>(r.ref + co), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= dact_fast<DACT>(t[i]);
        if (r.dkeep) {
            const uint2 k = *reinterpret_cast<const uint2*>(r.dkeep + co);
            const unsigned char* kb = reinterpret_cast<const unsigned char*>(&k);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
        }
    }
    *reinterpret_cast<uint4*>(r.out + co) = pack8(v);
}

// NCOLS accumulator columns (already in registers) -> fused epilogue for channels [col0, col0+NCOLS) of the group
template <int ACT, int DACT, int NCOLS>
__device__ __forceinline__ void epilogue_cols(const RowCtx& rc, const uint32_t (&r)[NCOLS], int col0, int cout_g, int cbase) {
#pragma unroll
    for (int j = 0; j < NCOLS; j += 8) {
        const int col = col0 + j;
        if (col < cout_g) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j + i]);
            epilogue_store8<ACT, DACT>(rc, cbase + col, v);
        }
    }
}

// runtime (act, dact) -> compile-time variant.  A forward epilogue has an activation and no derivative, a
// backward one has a derivative and no activation; anything else takes the generic (0,0)+scalar route upstream.
#define TBI_EPI_DISPATCH(ACTV, DACTV, CALL)                                                          \
    do {                                                                                             \
        if ((DACTV) == TBI_ACT_NONE) {                                                               \
            switch (ACTV) {                                                                          \
                case TBI_ACT_ELU:   { constexpr int A_ = TBI_ACT_ELU,   D_ = TBI_ACT_NONE; CALL; } break;   \
                case TBI_ACT_RELU:  { constexpr int A_ = TBI_ACT_RELU,  D_ = TBI_ACT_NONE; CALL; } break;   \
                case TBI_ACT_LRELU: { constexpr int A_ = TBI_ACT_LRELU, D_ = TBI_ACT_NONE; CALL; } break;   \
                default:            { constexpr int A_ = TBI_ACT_NONE,  D_ = TBI_ACT_NONE; CALL; } break;   \
            }                                                                                        \
        } else {                                                                                     \
            switch (DACTV) {                                                                         \
                case TBI_ACT_ELU:   { constexpr int A_ = TBI_ACT_NONE, D_ = TBI_ACT_ELU;   CALL; } break;   \
                case TBI_ACT_RELU:  { constexpr int A_ = TBI_ACT_NONE, D_ = TBI_ACT_RELU;  CALL; } break;   \
                default:            { constexpr int A_ = TBI_ACT_NONE, D_ = TBI_ACT_LRELU; CALL; } break;   \
            }                                                                                        \
        }                                                                                            \
    } while (0)

}  // namespace
