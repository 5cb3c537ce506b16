// Epilogue shared by the tcgen05 tap-GEMM kernels.  One thread owns one output pixel (one TMEM lane) and
// walks its channels 8 at a time with 16-byte accesses.  Everything that does not depend on the channel
// (row pointers, which optional terms exist) is resolved once per tile in RowCtx; the activation and the
// fused activation-derivative are template parameters so the per-element code is branch-free.
#pragma once
#include "tbi_common.cuh"

namespace {

// exp(v) for v <= 0 as one FMUL + one MUFU.  (__expf adds a denormal-range rescue around ex2 -- ~10 instructions and a
// handful of predicates per value -- which made the ELU epilogue issue-bound; below -126*ln2 the answer flushes to 0,
// i.e. ELU = -1, exactly what bf16 would round to anyway.)
__device__ __forceinline__ float exp_neg_fast(float v) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v * 1.4426950408889634f));
    return y;
}
template <int ACT> __device__ __forceinline__ float act_fast(float v) {
    if (ACT == TBI_ACT_ELU)   return v > 0.f ? v : exp_neg_fast(v) - 1.f;     // SFU exp: abs error ~1e-7 << bf16 resolution
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

// 32 bytes per thread per access when the row is 32-byte aligned (sm_100 256-bit global accesses): every access is then a
// whole 32-byte sector.  With 16-byte accesses each sector was written in two halves by two instructions (ncu: 2x the
// sector writes, L1TEX the busiest unit of the stem convolutions).
__device__ __forceinline__ void st_global_32B(void* p, const uint4& a, const uint4& b) {
    if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
    } else {
        reinterpret_cast<uint4*>(p)[0] = a; reinterpret_cast<uint4*>(p)[1] = b;
    }
}
__device__ __forceinline__ void ld_global_32B(const void* p, uint4& a, uint4& b) {
    if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
        asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p) : "memory");
    } else {
        a = reinterpret_cast<const uint4*>(p)[0]; b = reinterpret_cast<const uint4*>(p)[1];
    }
}

// per-thread, per-tile pointers to channel 0 of this thread's output pixel in every tensor the epilogue touches
struct RowCtx {
    __nv_bfloat16* out; __nv_bfloat16* out2;
    const __nv_bfloat16* res; const __nv_bfloat16* res2; const __nv_bfloat16* ref;
    const uint8_t* keep; const uint8_t* dkeep;
    const float* bias;
    int split_c;
    __device__ __forceinline__ __nv_bfloat16* p_out() const { return out; }
    __device__ __forceinline__ __nv_bfloat16* p_out2() const { return out2; }
    __device__ __forceinline__ const __nv_bfloat16* p_res() const { return res; }
    __device__ __forceinline__ const __nv_bfloat16* p_res2() const { return res2; }
    __device__ __forceinline__ const __nv_bfloat16* p_ref() const { return ref; }
    __device__ __forceinline__ const uint8_t* p_keep() const { return keep; }
    __device__ __forceinline__ const uint8_t* p_dkeep() const { return dkeep; }
    __device__ __forceinline__ int split() const { return split_c; }
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

// The same context in TWO registers: every tensor of the epilogue shares the output's pixel grid (checked on the host,
// tbi_tapgemm_tc_supported), so one linear pixel index plus the epilogue descriptor -- a kernel parameter, i.e. constant-bank
// operands -- gives each row pointer with one IMAD.WIDE where it is used.  The eight live 64-bit pointers of RowCtx were what
// pushed the persistent kernels' epilogue loops over their 96-register budget.
struct LeanRowCtx {
    const tbi_epilogue* e;
    int lin;
    const float* bias;
    typedef __nv_bfloat16 T;
    __device__ __forceinline__ T* p_out() const { return (T*)e->out.ptr + ((size_t)lin * e->out.cstride + e->out.coff); }
    __device__ __forceinline__ T* p_out2() const { return e->split_c > 0 ? (T*)e->out2.ptr + ((size_t)lin * e->out2.cstride + e->out2.coff) : nullptr; }
    __device__ __forceinline__ const T* p_res() const { return e->residual.ptr ? (const T*)e->residual.ptr + ((size_t)lin * e->residual.cstride + e->residual.coff) : nullptr; }
    __device__ __forceinline__ const T* p_res2() const { return e->residual2.ptr ? (const T*)e->residual2.ptr + ((size_t)lin * e->residual2.cstride + e->residual2.coff) : nullptr; }
    __device__ __forceinline__ const T* p_ref() const { return e->dact != TBI_ACT_NONE ? (const T*)e->dact_ref.ptr + ((size_t)lin * e->dact_ref.cstride + e->dact_ref.coff) : nullptr; }
    __device__ __forceinline__ const uint8_t* p_keep() const { return e->drop_keep ? e->drop_keep + (size_t)lin * e->out.c : nullptr; }
    __device__ __forceinline__ const uint8_t* p_dkeep() const { return e->dact_keep ? e->dact_keep + (size_t)lin * e->dact_ref.c : nullptr; }
    __device__ __forceinline__ int split() const { return e->split_c; }
};
__device__ __forceinline__ LeanRowCtx make_lean_row_ctx(const tbi_epilogue& e, int n, int oy, int ox, const float* bias) {
    LeanRowCtx r;
    r.e = &e; r.lin = (n * e.out.h + oy) * e.out.w + ox; r.bias = bias;
    return r;
}

// Side inputs of one 8-channel group (slow path: ragged channel counts, chunks straddling the split point).
struct Side8 { uint4 res, ref; uint2 k; };      // k = forward dropout multiplier OR the act' dropout multiplier (never both)

template <int DACT, typename Ctx>
__device__ __forceinline__ void load_side8(const Ctx& r, int co, Side8& s) {
    if (r.split() > 0 && co >= r.split()) {
        if (r.p_res2()) s.res = *reinterpret_cast<const uint4*>(r.p_res2() + (co - r.split()));
        return;
    }
    if (r.p_res()) s.res = *reinterpret_cast<const uint4*>(r.p_res() + co);
    if (DACT != TBI_ACT_NONE) {
        s.ref = *reinterpret_cast<const uint4*>(r.p_ref() + co);
        if (r.p_dkeep()) s.k = *reinterpret_cast<const uint2*>(r.p_dkeep() + co);
    } else if (r.p_keep()) {
        s.k = *reinterpret_cast<const uint2*>(r.p_keep() + co);
    }
}

// same math as epilogue_store<bf16> (tbi_common.cuh) on 8 consecutive channels starting at co
template <int ACT, int DACT, typename Ctx>
__device__ __forceinline__ void finish_store8(const Ctx& r, int co, float (&v)[8], const Side8& s) {
    if (r.split() > 0 && co >= r.split()) {
        if (r.p_res2()) {
            float t[8]; unpack8(s.res, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
        }
        *reinterpret_cast<uint4*>(r.p_out2() + (co - r.split())) = pack8(v);
        return;
    }
    if (r.bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(r.bias + co), b1 = *reinterpret_cast<const float4*>(r.bias + co + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (DACT == TBI_ACT_NONE && r.p_keep()) {
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&s.k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fast<ACT>(v[i]);
    if (r.p_res()) {
        float t[8]; unpack8(s.res, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (DACT != TBI_ACT_NONE) {
        float t[8]; unpack8(s.ref, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= dact_fast<DACT>(t[i]);
        if (r.p_dkeep()) {
            const unsigned char* kb = reinterpret_cast<const unsigned char*>(&s.k);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
        }
    }
    *reinterpret_cast<uint4*>(r.p_out() + co) = pack8(v);
}

// Fast path: NCOLS in-range channels on one side of the split.  Loads and stores go through pointers the compiler must
// assume alias, so a load placed after a store waits for nothing in hardware but cannot be hoisted by the compiler, and
// the epilogue degenerates into one DRAM/L2 round trip per 8-channel group (measured ~800-3000 cycles each).  Hence two
// phases: (A) every global side input of the chunk into registers, (B) math + stores with no global load in between.
// The bias comes from `bias` with plain loads inside phase B: the halo kernel points it at a shared-memory copy.
// Which optional inputs exist is a template parameter (HAS_RES, HAS_K): a variant holds registers only for what it reads.
template <int ACT, int DACT, bool HAS_RES, bool HAS_K, int NCOLS>
__device__ __forceinline__ void epilogue_fast(__nv_bfloat16* out, const __nv_bfloat16* res, const __nv_bfloat16* ref, const uint8_t* k,
                                              const float* bias, const uint32_t (&r)[NCOLS]) {
    constexpr int G = NCOLS / 8;
    uint4 qres[HAS_RES ? G : 1], qref[DACT != TBI_ACT_NONE ? G : 1];
    uint2 qk[HAS_K ? G : 1];
    static_assert(G % 2 == 0, "channel groups are processed in pairs (32-byte accesses)");
    if (HAS_RES) {
#pragma unroll
        for (int j = 0; j < G; j += 2) ld_global_32B(res + 8 * j, qres[j], qres[j + 1]);
    }
    if (DACT != TBI_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < G; j += 2) ld_global_32B(ref + 8 * j, qref[j], qref[j + 1]);
    }
    if (HAS_K) {
#pragma unroll
        for (int j = 0; j < G; ++j) qk[j] = *reinterpret_cast<const uint2*>(k + 8 * j);
    }
    uint4 held;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
        const unsigned char* kb = reinterpret_cast<const unsigned char*>(&qk[HAS_K ? j : 0]);
        if (DACT == TBI_ACT_NONE) {
            if (bias) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * j), b1 = *reinterpret_cast<const float4*>(bias + 8 * j + 4);
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (HAS_K) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_fast<ACT>(v[i]);
        if (HAS_RES) {
            float t[8]; unpack8(qres[HAS_RES ? j : 0], t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += t[i];
        }
        if (DACT != TBI_ACT_NONE) {
            float t[8]; unpack8(qref[DACT != TBI_ACT_NONE ? j : 0], t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= dact_fast<DACT>(t[i]);
            if (HAS_K) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= (float)kb[i];
            }
        }
        if (j & 1) st_global_32B(out + 8 * (j - 1), held, pack8(v));
        else held = pack8(v);
    }
}

// NCOLS accumulator columns (already in registers) -> fused epilogue for channels [col0, col0+NCOLS) of the group
template <int ACT, int DACT, int NCOLS, typename Ctx>
__device__ __forceinline__ void epilogue_cols(const Ctx& rc, const uint32_t (&r)[NCOLS], int col0, int cout_g, int cbase) {
    const int co0 = cbase + col0;
    const bool whole = col0 + NCOLS <= cout_g;
    const int split_c = rc.split();
    if (whole && split_c > 0 && co0 >= split_c) {                             // the pass-through half of a split output
        const int c2 = co0 - split_c;
        const __nv_bfloat16* res2 = rc.p_res2();
        if (res2) epilogue_fast<TBI_ACT_NONE, TBI_ACT_NONE, true, false, NCOLS>(rc.p_out2() + c2, res2 + c2, nullptr, nullptr, nullptr, r);
        else      epilogue_fast<TBI_ACT_NONE, TBI_ACT_NONE, false, false, NCOLS>(rc.p_out2() + c2, nullptr, nullptr, nullptr, nullptr, r);
        return;
    }
    if (whole && (split_c <= 0 || co0 + NCOLS <= split_c) && !(DACT != TBI_ACT_NONE && rc.bias)) {
        const uint8_t* k = DACT != TBI_ACT_NONE ? rc.p_dkeep() : rc.p_keep();
        const float* bias = rc.bias ? rc.bias + co0 : nullptr;
        const __nv_bfloat16* res = rc.p_res();
        const __nv_bfloat16* ref = DACT != TBI_ACT_NONE ? rc.p_ref() + co0 : nullptr;
        if (res) {
            if (k) epilogue_fast<ACT, DACT, true, true, NCOLS>(rc.p_out() + co0, res + co0, ref, k + co0, bias, r);
            else   epilogue_fast<ACT, DACT, true, false, NCOLS>(rc.p_out() + co0, res + co0, ref, nullptr, bias, r);
        } else {
            if (k) epilogue_fast<ACT, DACT, false, true, NCOLS>(rc.p_out() + co0, nullptr, ref, k + co0, bias, r);
            else   epilogue_fast<ACT, DACT, false, false, NCOLS>(rc.p_out() + co0, nullptr, ref, nullptr, bias, r);
        }
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
