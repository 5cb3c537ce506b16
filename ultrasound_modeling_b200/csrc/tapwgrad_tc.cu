// tcgen05 weight-gradient tap-GEMM (bf16 in, fp32 accumulate in TMEM) for sm_100a.
//
//   dw[tap][ci][co] += sum_pixels a[pixel + a_off(tap)][ci] * b[pixel + b_off(tap)][co]
//
// Per tap this is a GEMM with M = ci, N = co and K = PIXELS.  Activations are NHWC, so both operands
// arrive "MN-major" (the M/N index -- the channel -- is the contiguous one): every K-step loads, per
// 64-channel block, one 5-D TMA box [64 pixels x 64 channels] (128-byte rows, 128B swizzle) which is
// exactly the canonical MN-major SW128 UMMA layout (atom = 64 channels x 8 pixels; SBO = 1024 B between
// 8-pixel groups; LBO = one box = 8192 B between 64-channel blocks).  Shifted / stride-2 taps are box
// coordinates, out-of-image pixels are zero-filled by TMA (== they contribute nothing).
// One CTA owns T taps x one (128-ci, BN-co) tile of dw -- T accumulators side by side in TMEM -- and a contiguous range of
// pixel tiles (split-K); the partial results are reduced into the fp32 gradient with red.global.add (caller zeroes dw).
// T > 1 when the taps of the group read one operand at the same pixels (a transposed conv's taps all read the same input
// pixels; a stride-1 conv's taps all read the same output-gradient pixels): that operand is loaded once per K-step and the
// L2->shared-memory stream per MMA drops from 32 KB to ~20 KB (the T = 1 kernel sat on that stream: 64 flop/B).
// Warp roles as in tapgemm_tc.cu.
#include "tbi_common.cuh"
#include "tc_common.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int WG_THREADS = 192;
constexpr int KP = 64;                       // pixels per K-step
constexpr uint32_t BOX_BYTES = KP * 128;     // one [64 px x 64 ch] bf16 box

struct alignas(64) TcWgradParams {
    CUtensorMap a[2];
    CUtensorMap b;
    int n, gh, gw;
    int ltw, lth;                 // pixel tile: tw x th x tn = 64
    int tiles_x, tiles_y, tiles_b;
    int cin_g, cout_g, groups, c0;
    int ci_tiles, co_tiles, ntaps;
    int nb;                       // 64-channel blocks of the co tile (BN = 64*nb)
    int stages, tiles_per_cta;
    int ci_tile_base, co_tile_base;   // this launch covers ci tiles [base, base + ci_tiles) and co tiles [base, base + co_tiles)
    int b_cbase, b_cpix;
    signed char aqy[16], aqx[16], bqy[16], bqx[16], bay[16], bax[16];
    float* dw;
    long long tap_stride, ci_stride, co_stride;
};

__device__ __forceinline__ void wg_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!tc::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("tbi tcgen05 wgrad: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
    }
}

__host__ __device__ constexpr int wg_pow2(int v) { return v <= 32 ? 32 : v <= 64 ? 64 : v <= 128 ? 128 : v <= 256 ? 256 : 512; }

template <int NB, int T, bool SHARE_A>
__global__ void __launch_bounds__(WG_THREADS) tapwgrad_tc_kernel(const __grid_constant__ TcWgradParams p) {
    constexpr int BN = 64 * NB;
    constexpr int TCOLS = wg_pow2(T * BN);
    constexpr int A_BOXES = SHARE_A ? 2 : 2 * T, B_BOXES = SHARE_A ? NB * T : NB;      // boxes per stage (T == 1: 2 and NB)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int stages = p.stages;
    constexpr uint32_t STAGE_BYTES = (A_BOXES + B_BOXES) * BOX_BYTES;
    uint8_t* bar_base = smem + (size_t)stages * STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty = full + stages;
    uint64_t* tfull = empty + stages;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.a[0]); tc::prefetch_tmap(&p.b);
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<TCOLS>(tslot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tslot;

    int t = blockIdx.x;
    const int tap = (t % (p.ntaps / T)) * T; t /= (p.ntaps / T);          // first tap of this CTA's group
    const int co_t = t % p.co_tiles; const int ci_t = t / p.co_tiles;
    const int g = blockIdx.z;
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    const int tile_beg = blockIdx.y * p.tiles_per_cta;
    const int tile_end = min(total_tiles, tile_beg + p.tiles_per_cta);
    const int iters = tile_end - tile_beg;
    const int tw = 1 << p.ltw, th = 1 << p.lth, tn = KP >> (p.ltw + p.lth);
    const int ci0 = (ci_t + p.ci_tile_base) * 128, co0 = (co_t + p.co_tile_base) * BN;

    if (warp == 0) {
        if (lane == 0 && iters > 0) {
            // ===== TMA producer =====
            // channel coordinates of the two 64-wide A blocks (virtual concat: a block lives in one source)
            int asrc[2], ach[2];
            for (int j = 0; j < 2; ++j) {
                int ch = ci0 + 64 * j + (p.groups > 1 ? g * p.cin_g : 0);
                int src = 0;
                if (p.groups == 1 && ch >= p.c0 && p.c0 < p.cin_g) { src = 1; ch -= p.c0; }
                asrc[j] = src; ach[j] = ch;
            }
            for (int it = 0; it < iters; ++it) {
                const int s = it % stages;
                wg_wait(&empty[s], (((uint32_t)(it / stages)) & 1u) ^ 1u);
                int tt = tile_beg + it;
                const int tix = tt % p.tiles_x; tt /= p.tiles_x;
                const int tiy = tt % p.tiles_y; const int tib = tt / p.tiles_y;
                const int x0 = tix * tw, y0 = tiy * th, n0 = tib * tn;
                uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                tc::mbar_expect_tx(&full[s], STAGE_BYTES);
                // stage layout: [A boxes: (tap,) 64-ci block][B boxes: (tap,) 64-co block]; the shared operand has one copy
#pragma unroll
                for (int tt2 = 0; tt2 < (SHARE_A ? 1 : T); ++tt2)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        tc::tma_load_5d(st + (tt2 * 2 + j) * BOX_BYTES, &p.a[asrc[j]], &full[s], ach[j], x0 + (int)p.aqx[tap + tt2], 0, y0 + (int)p.aqy[tap + tt2], n0);
#pragma unroll
                for (int tt2 = 0; tt2 < (SHARE_A ? T : 1); ++tt2)
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        tc::tma_load_5d(st + (A_BOXES + tt2 * NB + j) * BOX_BYTES, &p.b, &full[s],
                                        p.b_cbase + g * p.cout_g + co0 + 64 * j + (int)p.bax[tap + tt2] * p.b_cpix,
                                        x0 + (int)p.bqx[tap + tt2], (int)p.bay[tap + tt2], y0 + (int)p.bqy[tap + tt2], n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && iters > 0) {
            // ===== MMA issuer: D[ci][co] += A^T[ci][px] * B[px][co], both operands MN-major =====
            const uint32_t idesc = tc::make_idesc_bf16(128, BN, 1, 1);
            for (int it = 0; it < iters; ++it) {
                const int s = it % stages;
                wg_wait(&full[s], ((uint32_t)(it / stages)) & 1u);
                tc::tc_fence_after();
                const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BOXES * BOX_BYTES;
#pragma unroll
                for (int tt2 = 0; tt2 < T; ++tt2) {
                    const uint32_t a_t = a_addr + (SHARE_A ? 0 : tt2 * 2) * BOX_BYTES, b_t = b_addr + (SHARE_A ? tt2 * NB : 0) * BOX_BYTES;
#pragma unroll
                    for (int k = 0; k < KP / 16; ++k) {
                        // 16 pixels = 2 swizzle atoms along K: +2048 B per step; LBO = box (next 64 channels), SBO = 1024 B
                        const uint64_t da = tc::make_smem_desc(a_t + k * 2048, BOX_BYTES, 1024, 2u);
                        const uint64_t db = tc::make_smem_desc(b_t + k * 2048, BOX_BYTES, 1024, 2u);
                        tc::umma_bf16(tmem_base + tt2 * BN, da, db, idesc, (it | k) != 0 ? 1u : 0u);
                    }
                }
                tc::umma_commit(&empty[s]);
            }
            tc::umma_commit(tfull);
        }
        __syncwarp();
    } else if (iters > 0) {
        // ===== epilogue: lane = ci row, columns = co; reduce into the fp32 gradient =====
        const int q = warp & 3;
        const int ci = ci0 + q * 32 + lane;
        const bool row_ok = ci < p.cin_g;
        wg_wait(tfull, 0);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        // split-K partials of many CTAs land on the same fp32 words: start each CTA at its own 32-column chunk so they do not walk
        // the addresses in lock step, and with unit cout stride (Conv2D, HWIO) send 16-byte vector reductions
        const bool vec = p.co_stride == 1 && (p.cout_g & 3) == 0 && (p.ci_stride & 3) == 0 && (p.tap_stride & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(p.dw) & 15) == 0);
        constexpr int CPT = BN / 32, CHUNKS = T * CPT;
        int idx = (int)((blockIdx.x * 5u + blockIdx.y * 3u + blockIdx.z) % (unsigned)CHUNKS);
#pragma unroll 1
        for (int k = 0; k < CHUNKS; ++k, idx = (idx + 1 == CHUNKS ? 0 : idx + 1)) {
            const int tt2 = idx / CPT, c = (idx - tt2 * CPT) * 32;
            float* base = p.dw + (size_t)(tap + tt2) * p.tap_stride + (size_t)ci * p.ci_stride;
            uint32_t r[32];
            tc::tmem_ld32(taddr + tt2 * BN + c, r);
            tc::tmem_ld_wait();
            if (!row_ok) continue;
            if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int co = co0 + c + j;
                    if (co < p.cout_g)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + (size_t)(g * p.cout_g + co)), "f"(__uint_as_float(r[j])),
                                     "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = co0 + c + j;
                    if (co < p.cout_g) atomicAdd(base + (size_t)(g * p.cout_g + co) * p.co_stride, __uint_as_float(r[j]));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TCOLS>(tmem_base);
}

bool wg_aligned_view(const tbi_view& v) {
    return v.ptr == nullptr || (v.cstride % 8 == 0 && v.coff % 8 == 0 && v.c % 8 == 0 && ((uintptr_t)v.ptr & 15) == 0);
}
int wg_ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// 5-D activation map with a [64 ch x tw x 1 x th x tn] box (see make_act_tmap in tapgemm_tc.cu)
int wg_act_tmap(CUtensorMap* out, const tbi_view& v, int n, int stride, int tw, int th, int tn, int* cbase, int* cpix) {
    uint64_t dims[5], strides[4];
    uint32_t box[5] = {64u, (uint32_t)tw, 1u, (uint32_t)th, (uint32_t)tn};
    const uint64_t px = (uint64_t)v.cstride * 2;
    void* base;
    if (stride == 1) {
        dims[0] = (uint64_t)v.c; dims[1] = (uint64_t)v.w; dims[2] = 1; dims[3] = (uint64_t)v.h; dims[4] = (uint64_t)n;
        strides[0] = px; strides[1] = px * v.w; strides[2] = px * v.w; strides[3] = px * v.w * v.h;
        base = (char*)v.ptr + (size_t)v.coff * 2;
        *cbase = 0; *cpix = 0;
    } else {
        dims[0] = (uint64_t)v.cstride * 2; dims[1] = (uint64_t)v.w / 2; dims[2] = 2; dims[3] = (uint64_t)v.h / 2; dims[4] = (uint64_t)n;
        strides[0] = px * 2; strides[1] = px * v.w; strides[2] = px * v.w * 2; strides[3] = px * v.w * v.h;
        base = v.ptr;
        *cbase = v.coff; *cpix = v.cstride;
    }
    return tbi_make_tmap_bf16(out, base, 5, dims, strides, box, 128);
}

template <int NB, int T, bool SHARE_A>
int launch_wg(const TcWgradParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tapwgrad_tc_kernel<NB, T, SHARE_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attr_err != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    tapwgrad_tc_kernel<NB, T, SHARE_A><<<grid, WG_THREADS, smem, s>>>(p);
    TBI_CUDA_LAUNCH_CHECK("tapwgrad_tc");
    return TBI_OK;
}

}  // namespace

bool tbi_tapwgrad_tc_supported(const tbi_tapwgrad* d, const char** why) {
#define NO(msg) do { *why = msg; return false; } while (0)
    if (d->dtype != TBI_BF16) NO("storage dtype is not bf16");
    if (!tbi_get_encode_tiled()) NO("no cuTensorMapEncodeTiled");
    if (d->ntaps < 1 || d->ntaps > TBI_MAX_TAPS) NO("ntaps");
    {
        const bool views_ok = d->a_src[0].cstride % 8 == 0 && d->a_src[0].coff % 8 == 0 && ((uintptr_t)d->a_src[0].ptr & 15) == 0 &&
                              d->b_src.cstride % 8 == 0 && d->b_src.coff % 8 == 0 && ((uintptr_t)d->b_src.ptr & 15) == 0;
        if (views_ok && tbi_tapwgrad_small_supported(d)) return true;
    }
    if (d->a_stride != 1) NO("a_stride");
    if (d->b_stride != 1 && d->b_stride != 2) NO("b_stride");
    if (d->groups > 1 && d->a_src[1].ptr) NO("groups with two sources");
    if (d->a_src[1].ptr && d->a_src[0].c % 64 != 0) NO("first source of a concat must be a multiple of 64 channels");
    if (d->cin_g % 8 != 0) NO("input channels per group must be a multiple of 8");
    if (d->groups > 1 && d->cout_g % 8 != 0) NO("grouped: output channels per group must be a multiple of 8");
    if (d->cin_g < 16) NO("too few input channels for the tensor-core path");
    if (d->b_src.c < d->cout_g * d->groups) NO("gradient view has fewer channels than cout");
    if (!wg_aligned_view(d->a_src[0]) || !wg_aligned_view(d->a_src[1]) || !wg_aligned_view(d->b_src)) NO("view not 16-byte aligned");
    if (d->b_stride == 2 && ((d->b_src.h | d->b_src.w) & 1)) NO("stride-2 gather needs even dims");
    for (int t = 0; t < d->ntaps; ++t)
        if (d->a_dy[t] < -64 || d->a_dy[t] > 64 || d->a_dx[t] < -64 || d->a_dx[t] > 64 || d->b_dy[t] < -64 || d->b_dy[t] > 64 || d->b_dx[t] < -64 || d->b_dx[t] > 64) NO("tap offset range");
    return true;
#undef NO
}

int64_t tbi_tapwgrad_tc_workspace(const tbi_tapwgrad*) { return 0; }   // partials are reduced with red.global.add: no workspace

int tbi_tapwgrad_tc(const tbi_tapwgrad* d, cudaStream_t s) {
    const char* why = "";
    if (!tbi_tapwgrad_tc_supported(d, &why)) return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapwgrad_tc: %s", why);
    if (tbi_tapwgrad_small_supported(d)) return tbi_tapwgrad_small(d, s);
    // CTA pairs (cta_group::2, tapwgrad_tc2.cu) take every pair of 128-channel tiles of the M operand; an odd last tile (e.g. the
    // 64 skip channels of upsample_4's 320-channel input) is left to the single-CTA kernel below
    const tbi_wgrad_pair_plan pl = tbi_tapwgrad_pair_plan(d);
    int rest_ci_base = 0, rest_ci_tiles = -1, rest_co_base = 0, rest_co_tiles = -1;
    if (pl.ok) {
        int rcp = tbi_tapwgrad_pair(d, pl, s); if (rcp) return rcp;
        if (pl.m_tiles % 2 == 0) {
            if (d->dbias) return tbi_colsum(d->dtype, (int64_t)d->n * d->b_src.h * d->b_src.w, &d->b_src, d->dbias, (void*)s);
            return TBI_OK;
        }
        if (pl.m_is_a) { rest_ci_base = pl.m_tiles - 1; rest_ci_tiles = 1; }
        else           { rest_co_base = pl.m_tiles - 1; rest_co_tiles = 1; }       // M = co in 128-channel tiles == this kernel's BN = 128 tiles
    }
    TcWgradParams p; memset(&p, 0, sizeof(p));
    int ltw = wg_ilog2_ceil(d->gw); if (ltw > 3) ltw = 3;
    int lth = wg_ilog2_ceil(d->gh); if (lth > 6 - ltw) lth = 6 - ltw;
    const int tw = 1 << ltw, th = 1 << lth, tn = KP >> (ltw + lth);
    p.n = d->n; p.gh = d->gh; p.gw = d->gw; p.ltw = ltw; p.lth = lth;
    p.tiles_x = (d->gw + tw - 1) / tw; p.tiles_y = (d->gh + th - 1) / th; p.tiles_b = (d->n + tn - 1) / tn;
    p.cin_g = d->cin_g; p.cout_g = d->cout_g; p.groups = d->groups;
    p.c0 = d->groups > 1 ? d->cin_g : d->a_src[0].c;
    p.ntaps = d->ntaps;
    p.ci_tiles = (d->cin_g + 127) / 128;
    p.nb = d->cout_g > 64 ? 2 : 1;
    const int bn = 64 * p.nb;
    p.co_tiles = (d->cout_g + bn - 1) / bn;
    if (rest_ci_tiles > 0) { p.ci_tile_base = rest_ci_base; p.ci_tiles = rest_ci_tiles; }
    if (rest_co_tiles > 0) { p.co_tile_base = rest_co_base; p.co_tiles = rest_co_tiles; }
    for (int t = 0; t < d->ntaps; ++t) {
        p.aqy[t] = (signed char)d->a_dy[t]; p.aqx[t] = (signed char)d->a_dx[t];
        if (d->b_stride == 1) { p.bqy[t] = (signed char)d->b_dy[t]; p.bqx[t] = (signed char)d->b_dx[t]; p.bay[t] = 0; p.bax[t] = 0; }
        else {
            const int ay = d->b_dy[t] & 1, ax = d->b_dx[t] & 1;
            p.bay[t] = (signed char)ay; p.bax[t] = (signed char)ax;
            p.bqy[t] = (signed char)((d->b_dy[t] - ay) / 2); p.bqx[t] = (signed char)((d->b_dx[t] - ax) / 2);
        }
    }
    int cb, cp;
    int rc = wg_act_tmap(&p.a[0], d->a_src[0], d->n, 1, tw, th, tn, &cb, &cp); if (rc) return rc;
    if (d->groups == 1 && d->a_src[1].ptr) { rc = wg_act_tmap(&p.a[1], d->a_src[1], d->n, 1, tw, th, tn, &cb, &cp); if (rc) return rc; }
    else p.a[1] = p.a[0];
    rc = wg_act_tmap(&p.b, d->b_src, d->n, d->b_stride, tw, th, tn, &p.b_cbase, &p.b_cpix); if (rc) return rc;
    p.dw = d->dw; p.tap_stride = d->tap_stride; p.ci_stride = d->ci_stride; p.co_stride = d->co_stride;

    // tap groups: T taps per CTA when they share one operand's pixels (see the header comment) and the layer is big enough to
    // be bound by the operand stream (both channel counts >= 64)
    static const bool no_multi = getenv("TBI_WGRAD_NO_MULTITAP") != nullptr;
    int T = 1; bool share_a = true;
    if (!no_multi && d->cin_g >= 64 && d->cout_g >= 64 && d->ntaps > 1) {
        bool a_same = true, b_same = true;
        for (int t = 1; t < d->ntaps; ++t) {
            a_same = a_same && p.aqy[t] == p.aqy[0] && p.aqx[t] == p.aqx[0];
            b_same = b_same && p.bqy[t] == p.bqy[0] && p.bqx[t] == p.bqx[0] && p.bay[t] == p.bay[0] && p.bax[t] == p.bax[0];
        }
        if ((a_same || b_same) && d->ntaps % 4 == 0) T = 4;
        else if ((a_same || b_same) && d->ntaps % 3 == 0) T = 3;
        share_a = a_same;
    }
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    const long long out_tiles = (long long)(d->ntaps / T) * p.ci_tiles * p.co_tiles * d->groups;
    long long ksplit = (4LL * tbi_sm_count() + out_tiles - 1) / out_tiles;      // T = 1: ~4 CTAs per SM in flight overall
    if (T > 1) {
        // one CTA per SM (it owns all of TMEM): pick the split-K factor with the shortest modelled time
        //   waves * (K-steps per CTA + E),  E = the epilogue's T*128*BN reduction atomics expressed in K-steps (~32, measured:
        //   upsample_1 went 104 -> 210 us when an 8-way split left 8 K-steps per CTA in front of that epilogue)
        const long long sms = tbi_sm_count();
        // (the cap was 16 while the reduction was scalar and contended; with vector reductions in rotated order a 49-way split of
        // stage 2's concats_2 -- 3 output tiles -- fills the GPU: 187 -> ~75 us)
        static const long long cap_env = getenv("TBI_WGRAD_KSPLIT_CAP") ? atoll(getenv("TBI_WGRAD_KSPLIT_CAP")) : 64;
        long long cap = total_tiles < cap_env ? total_tiles : cap_env;
        long long best = -1; ksplit = 1;
        for (long long ks = 1; ks <= cap; ++ks) {
            const long long ctas = out_tiles * ks, waves = (ctas + sms - 1) / sms;
            const long long cost = waves * ((total_tiles + ks - 1) / ks + 32);
            if (best < 0 || cost < best) { best = cost; ksplit = ks; }
        }
    }
    if (ksplit > total_tiles) ksplit = total_tiles;
    if (ksplit < 1) ksplit = 1;
    p.tiles_per_cta = (int)((total_tiles + ksplit - 1) / ksplit);
    ksplit = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const uint32_t stage_bytes = (T == 1 ? (2 + p.nb) : share_a ? (2 + T * p.nb) : (2 * T + p.nb)) * BOX_BYTES;
    int stages = (int)((T == 1 ? 96 : 192) * 1024 / stage_bytes);
    if (stages > p.tiles_per_cta) stages = p.tiles_per_cta;
    if (stages < 1) stages = 1;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    dim3 grid((unsigned)((d->ntaps / T) * p.ci_tiles * p.co_tiles), (unsigned)ksplit, (unsigned)d->groups);
    if (T == 1)      rc = p.nb == 2 ? launch_wg<2, 1, true>(p, grid, smem, s) : launch_wg<1, 1, true>(p, grid, smem, s);
    else if (T == 4) rc = p.nb == 2 ? (share_a ? launch_wg<2, 4, true>(p, grid, smem, s) : launch_wg<2, 4, false>(p, grid, smem, s))
                                    : (share_a ? launch_wg<1, 4, true>(p, grid, smem, s) : launch_wg<1, 4, false>(p, grid, smem, s));
    else             rc = p.nb == 2 ? (share_a ? launch_wg<2, 3, true>(p, grid, smem, s) : launch_wg<2, 3, false>(p, grid, smem, s))
                                    : (share_a ? launch_wg<1, 3, true>(p, grid, smem, s) : launch_wg<1, 3, false>(p, grid, smem, s));
    if (rc) return rc;
    if (d->dbias)                      // bias gradient = column sum of the (unshifted) output gradient
        return tbi_colsum(d->dtype, (int64_t)d->n * d->b_src.h * d->b_src.w, &d->b_src, d->dbias, (void*)s);
    return TBI_OK;
}
