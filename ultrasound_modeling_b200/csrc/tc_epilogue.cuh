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

// Side inputs of one 8-channel group (slow path: ragged channel counts, chunks straddling the split point).
struct Side8 { uint4 res, ref; uint2 k; };      // k = forward dropout multiplier OR the act' dropout multiplier (never both)

template <int DACT>
__device__ __forceinline__ void load_side8(const RowCtx& r, int co, Side8& s) {
    if (r.split_c > 0 && co >= r.split_c) {
        if (r.res2) s.res = *reinterpret_cast<const uint4*>(r.res2 + (co - r.split_c));
        return;
    }
    if (r.res) s.res = *reinterpret_cast<const uint4*>(r.res + co);
    if (DACT != TBI_ACT_NONE) {
        s.ref = *reinterpret_cast<const uint4*>(r.ref + co);
        if (r.dkeep) s.k = *reinterpret_cast<const uint2*>(r.dkeep + co);
    } else if (r.keep) {
        s.k = *reinterpret_cast<const uint2*>(r.keep + co);
    }
}

// same math as epilogue_store<bf16> (tbi_common.cuh) on 8 consecutive channels starting at co
template <int ACT, int DACT>
__device__ __forceinline__ void finish_store8(const RowCtx& r, int co, float (&v)[8], const Side8& s) {
    if (r.split_c > 0 && co >= r.split_c) {
        if (r.res2) {
            float t[8]; unpack8(s.res, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
        }
        *reinterpret_cast<uint4*>(r.out2 + (co - r.split_c)) = pack8(v);
        return;
    }
    if (r.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(r.bias + co)), b1 = __ldg(reinterpret_cast<const float4*>(r.bias + co + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (DACT == TBI_ACT_NONE && r.keep) {
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&s.k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fast<ACT>(v[i]);
    if (r.res) {
        float t[8]; unpack8(s.res, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (DACT != TBI_ACT_NONE) {
        float t[8]; unpack8(s.ref, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= dact_fast<DACT>(t[i]);
        if (r.dkeep) {
            const unsigned char* kb = reinterpret_cast<const unsigned char*>(&s.k);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
        }
    }
    *reinterpret_cast<uint4*>(r.out + co) = pack8(v);
}

// Fast path: NCOLS in-range channels on one side of the split, as ONE basic block.  Loads and stores go through
// pointers the compiler must assume alias, and a group's loads issued only after the previous group's store cost one
// DRAM round trip per group (measured ~3000 cycles each).  So: no data-dependent control flow (an absent residual or
// bias reads a 32-byte zero block with stride 0; the dropout multiplier is a template flag) and every side input comes
// through the read-only path, which lets the scheduler lift the loads of all groups above the first store as far as
// the register budget allows.  (A residual that aliases the output -- in-place accumulation -- is read exactly once,
// by the thread that then overwrites it, so the read-only path is safe there too.)
static __device__ __align__(32) unsigned char g_epi_zero[32];

#ifndef TBI_EPI_STAGE
#define TBI_EPI_STAGE 1          // 1: explicit load-all-then-store-all (measured faster: 14.36 vs 14.86 ms/step); 0: leave the hoisting to the scheduler
#endif
template <int ACT, int DACT, bool HAS_K, int NCOLS>
__device__ __forceinline__ void epilogue_fast(__nv_bfloat16* out, const __nv_bfloat16* __restrict__ res, int res_step,
                                              const __nv_bfloat16* __restrict__ ref, const uint8_t* __restrict__ k,
                                              const float* bias, int bias_step, const uint32_t (&r)[NCOLS]) {
    constexpr int G = NCOLS / 8;
#if TBI_EPI_STAGE
    uint4 qres[G], qref[DACT != TBI_ACT_NONE ? G : 1];
    uint2 qk[HAS_K ? G : 1];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        qres[j] = *reinterpret_cast<const uint4*>(res + j * res_step);
        if (DACT != TBI_ACT_NONE) qref[j] = *reinterpret_cast<const uint4*>(ref + 8 * j);
        if (HAS_K) qk[j] = *reinterpret_cast<const uint2*>(k + 8 * j);
    }
#endif
#pragma unroll
    for (int j = 0; j < G; ++j) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
#if TBI_EPI_STAGE
        const uint4 q_res = qres[j];
        const uint4 q_ref = qref[DACT != TBI_ACT_NONE ? j : 0];
        const uint2 kk = qk[HAS_K ? j : 0];
#else
        const uint4 q_res = __ldg(reinterpret_cast<const uint4*>(res + j * res_step));
        uint4 q_ref; uint2 kk;
        if (DACT != TBI_ACT_NONE) q_ref = __ldg(reinterpret_cast<const uint4*>(ref + 8 * j));
        if (HAS_K) kk = __ldg(reinterpret_cast<const uint2*>(k + 8 * j));
#endif
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kk);
        if (DACT == TBI_ACT_NONE) {
            // plain loads on purpose: they stay next to their use (an L1 hit) instead of holding 8 registers per group
            const float4 b0 = *reinterpret_cast<const float4*>(bias + j * bias_step), b1 = *reinterpret_cast<const float4*>(bias + j * bias_step + 4);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            if (HAS_K) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_fast<ACT>(v[i]);
        {
            float t[8]; unpack8(q_res, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
        }
        if (DACT != TBI_ACT_NONE) {
            float t[8]; unpack8(q_ref, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= dact_fast<DACT>(t[i]);
            if (HAS_K) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
            }
        }
        *reinterpret_cast<uint4*>(out + 8 * j) = pack8(v);
    }
}

// NCOLS accumulator columns (already in registers) -> fused epilogue for channels [col0, col0+NCOLS) of the group
template <int ACT, int DACT, int NCOLS>
__device__ __forceinline__ void epilogue_cols(const RowCtx& rc, const uint32_t (&r)[NCOLS], int col0, int cout_g, int cbase) {
    const int co0 = cbase + col0;
    const __nv_bfloat16* zero16 = reinterpret_cast<const __nv_bfloat16*>(g_epi_zero);
    const float* zero32 = reinterpret_cast<const float*>(g_epi_zero);
    const bool whole = col0 + NCOLS <= cout_g;
    if (whole && rc.split_c > 0 && co0 >= rc.split_c) {                       // the pass-through half of a split output
        const int c2 = co0 - rc.split_c;
        epilogue_fast<TBI_ACT_NONE, TBI_ACT_NONE, false, NCOLS>(rc.out2 + c2, rc.res2 ? rc.res2 + c2 : zero16, rc.res2 ? 8 : 0,
                                                                 nullptr, nullptr, zero32, 0, r);
        return;
    }
    if (whole && (rc.split_c <= 0 || co0 + NCOLS <= rc.split_c) && !(DACT != TBI_ACT_NONE && rc.bias)) {
        const __nv_bfloat16* res = rc.res ? rc.res + co0 : zero16;
        const int res_step = rc.res ? 8 : 0;
        const float* bias = rc.bias ? rc.bias + co0 : zero32;
        const int bias_step = rc.bias ? 8 : 0;
        const uint8_t* k = DACT != TBI_ACT_NONE ? rc.dkeep : rc.keep;
        if (k) epilogue_fast<ACT, DACT, true, NCOLS>(rc.out + co0, res, res_step, rc.ref + co0, k + co0, bias, bias_step, r);
        else   epilogue_fast<ACT, DACT, false, NCOLS>(rc.out + co0, res, res_step, rc.ref + co0, nullptr, bias, bias_step, r);
        return;
    }
#pragma unroll
    for (int j = 0; j < NCOLS / 8; ++j) {
        const int col = col0 + 8 * j;
        if (col < cout_g) {
            Side8 side;
            load_side8<DACT>(rc, cbase + col, side);
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
            finish_store8<ACT, DACT>(rc, cbase + col, v, side);
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
