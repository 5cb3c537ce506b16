// tcgen05 weight gradient for stride-1 convolutions with FEW input channels per group (cin_g <= 32: the
// full-resolution stem and stage-1 layers, where wgrad is HBM/dispatch-bound, not math-bound).
//
//   dw[tap][ci][co] += sum_pixels x[pixel + off(tap)][ci] * dz[pixel][co]          (3x3 dilation 1, or 1x1)
//
// Per 16x8 pixel tile the producer issues ONE box of x, 16 x (8+2) pixels (halo in x only), and ONE box of dz,
// (16+2) x 8 pixels (halo in y only).  Both operands are MN-major (channels contiguous, pixels = K).
// ALL NINE TAPS ARE ONE MMA: the three horizontal taps are packed into the M dimension -- an M block is kcA
// channels, consecutive M blocks start one PIXEL later (descriptor leading-byte-offset = row bytes), so M = 128
// covers (dx = -1, 0, +1, [unused]) x kcA input channels -- and the three vertical taps into the N dimension:
// an N block is kcB channels of dz, consecutive N blocks start one TILE ROW later (leading-byte-offset = 8 pixel
// rows), i.e. the K index is the pixel q that x is read around and N block n pairs it with dz[q + (n-1) W]:
//   D[(j, ci), (n, co)] = sum_q x[q + (j-1)][ci] * dz[q + (n-1) W][co] = dw[tap (dy = 1-n, dx = j-1)][ci][co].
// (A product x[p+off] dz[p] with p outside the image reads TMA zero fill; every in-image pair is counted once, in the
// tile that holds q.)  The accumulator (3 kcB TMEM columns) stays resident for the whole CTA: a CTA walks a contiguous
// range of pixel tiles (split-K across CTAs) and reduces its partial dw into the fp32 gradient with red.global.add
// once, at the end.
// MMAs per 128 pixels: 8 (K steps of 16 pixels), N = 3 kcB -- instead of 24 with one accumulator per dy (round 1)
// and 72 for tap-by-tap accumulation.  A tcgen05.mma costs ~49 cycles to dispatch whatever its N
// (profiles/r1_summary.md), so for these HBM-bound layers the dispatch count, not the math, was the limit.
#include "tbi_common.cuh"
#include "tc_common.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int WS_THREADS = 192;            // warps 0..3 epilogue, warp 4 TMA producer, warp 5 MMA issuer + TMEM owner
constexpr int WTW = 8, WTH = 16;

struct alignas(64) WsParams {
    CUtensorMap a, b;
    int n, gh, gw, tiles_x, tiles_y, total_tiles, tiles_per_cta;
    int cin_g, cout_g, groups;
    int kca, kcb;                 // channels per TMA row of x / dz (16 | 32 | 64), row bytes = 2*kc
    int nrow;                     // vertical taps (3, or 1 for a 1x1 conv); horizontal taps = nrow as well
    int halo;                     // 1 for 3x3, 0 for 1x1
    int stages, a_stage_bytes, b_stage_bytes, a_tx, b_tx;
    float* dw;
    long long tap_stride, ci_stride, co_stride;
    float* dbias;                 // bias gradient (column sums of dz) accumulated by the otherwise idle epilogue warps, or nullptr
};

template <int NB>   // N (= kcb) / 16
__global__ void __launch_bounds__(WS_THREADS) tapwgrad_small_kernel(const __grid_constant__ WsParams p) {
    constexpr int N = 16 * NB;
    constexpr int TMEM_COLS = 3 * N <= 64 ? 64 : 3 * N <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* done = empty + p.stages;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4 && lane == 0) {
        tc::prefetch_tmap(&p.a); tc::prefetch_tmap(&p.b);
        // a stage is free once its MMAs have completed AND (bias gradient fused) the four epilogue warps have summed its dz tile
        for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], p.dbias ? 5 : 1); }
        tc::mbar_init(done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc<TMEM_COLS>(tslot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tslot;

    const int g = blockIdx.y;
    const int tile_beg = blockIdx.x * p.tiles_per_cta;
    const int tile_end = min(p.total_tiles, tile_beg + p.tiles_per_cta);
    const int iters = tile_end - tile_beg;
    const int pitch = WTW + 2 * p.halo;                            // x: halo columns only; dz: halo rows only
    const uint32_t rowa = 2u * p.kca, rowb = 2u * p.kcb;

    if (warp == 4) {
        // ===================== TMA producer =====================
        const bool leader = tc::elect_one();
        uint32_t s = 0, par = 1;
        for (int it = 0; it < iters; ++it) {
            int tt = tile_beg + it;
            const int tix = tt % p.tiles_x; tt /= p.tiles_x;
            const int tiy = tt % p.tiles_y; const int n0 = tt / p.tiles_y;
            const int x0 = tix * WTW, y0 = tiy * WTH;
            tc::mbar_wait_bounded(&empty[s], par);
            if (leader) {
                uint8_t* st = smem + (size_t)s * stage_bytes;
                tc::mbar_expect_tx(&full[s], p.a_tx + p.b_tx);
                tc::tma_load_5d(st, &p.a, &full[s], g * p.cin_g, x0 - p.halo, 0, y0, n0);
                tc::tma_load_5d(st + p.a_stage_bytes, &p.b, &full[s], g * p.cout_g, x0, 0, y0 - p.halo, n0);
            }
            if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::make_idesc_bf16(128, p.nrow * N, 1, 1);              // both operands MN-major; N = (dy block, co)
        const uint32_t lay_a = p.kca == 64 ? 2u : p.kca == 32 ? 4u : 6u, lay_b = p.kcb == 64 ? 2u : p.kcb == 32 ? 4u : 6u;
        // A: M blocks of kca channels one pixel row apart (LBO = row bytes); K groups of 8 pixels = one tile row (SBO = pitch rows)
        const uint64_t da_base = tc::smem_desc_base(rowa, (uint32_t)pitch * rowa, lay_a);
        // B: N blocks of kcb channels one TILE ROW (8 pixels) apart; K groups of 8 pixels = one tile row as well
        const uint64_t db_base = tc::smem_desc_base(p.nrow > 1 ? 8u * rowb : rowb, 8u * rowb, lay_b);
        const uint32_t a_lo0 = (uint32_t)da_base, a_hi = (uint32_t)(da_base >> 32), b_lo0 = (uint32_t)db_base, b_hi = (uint32_t)(db_base >> 32);
        const uint32_t ring_lo = (tc::smem_u32(smem) & 0x3FFFFu) >> 4, stage_lo = (uint32_t)stage_bytes >> 4, boff_lo = (uint32_t)p.a_stage_bytes >> 4;
        const uint32_t a_kstep = (2u * pitch * rowa) >> 4, b_kstep = (16u * rowb) >> 4;
        uint32_t s = 0, par = 0, accum = 0;
        for (int it = 0; it < iters; ++it) {
            tc::mbar_wait_bounded(&full[s], par);
            tc::tc_fence_after();
            const uint32_t a0 = a_lo0 + ring_lo + s * stage_lo, b0 = b_lo0 + ring_lo + s * stage_lo + boff_lo;
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tc::umma_bf16_lh(tmem_base, a0 + ks * a_kstep, a_hi, b0 + ks * b_kstep, b_hi, idesc, (accum | ks) ? 1u : 0u);
                tc::umma_commit(&empty[s]);
            }
            accum = 1;
            if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; }
        }
        if (leader) tc::umma_commit(done);
        __syncwarp();
    } else if (iters > 0) {
        if (p.dbias) {
            // ---- bias gradient while the tiles stream by: dbias[co] += sum over the tile's 16 x 8 pixels of dz[pixel][co] ----
            // (the separate column-sum kernel re-read every dz tensor from HBM: 26 launches, 420 us per step, 0.83 of the HBM
            // peak, i.e. only fusion removes it).  The dz box is [18 or 16 rows][8 pixels][kcb channels], TMA-swizzled with
            // Swizzle<B,4,3> on the byte offset within the 1 KB-aligned stage (B = 3 / 2 / 1 for 128 / 64 / 32-byte pixel rows).
            // Thread t sums the 16-byte chunk (8 channels) t % nch of the pixels t / nch + k * (128 / nch).
            const int t = warp * 32 + lane;
            const int nch = p.kcb >> 3, c = t % nch, p0 = t / nch, pstep = 128 / nch;
            const uint32_t bmask = (uint32_t)(p.kcb == 64 ? 7 : p.kcb == 32 ? 3 : 1);
            float bs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[j] = 0.f;
            uint32_t s = 0, par = 0;
            for (int it = 0; it < iters; ++it) {
                tc::mbar_wait_bounded<true>(&full[s], par);
                const uint8_t* bt = smem + (size_t)s * stage_bytes + p.a_stage_bytes;
                for (int k = 0; k < nch; ++k) {
                    const int px = p0 + k * pstep;                                    // pixel of the 16 x 8 centre, row-major
                    const uint32_t o = (uint32_t)(px + 8 * p.halo) * rowb + (uint32_t)c * 16u;
                    const uint4 v = *reinterpret_cast<const uint4*>(bt + (o ^ (((o >> 7) & bmask) << 4)));
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); bs[2 * j] += f.x; bs[2 * j + 1] += f.y; }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[s]);
                if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; }
            }
            // lanes with equal t % nch hold the same channels: butterfly over the higher lane bits, then one atomic per warp
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = bs[j];
                for (int o = 16; o >= nch; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int co = c * 8 + j;
                if (lane < nch && co < p.cout_g) atomicAdd(p.dbias + g * p.cout_g + co, v);
            }
        }
        // ===================== epilogue: TMEM lane = (dx block j, ci) ; columns = co ; one accumulator per dy =====================
        const int m = warp * 32 + lane;
        const int j = m / p.kca, ci = m % p.kca;
        const bool row_ok = j < p.nrow && ci < p.cin_g;
        tc::mbar_wait_bounded<true>(done, 0);
        tc::tc_fence_after();
        // Every CTA adds a full (taps x cin x cout) partial into the same few thousand fp32 words.  Two things keep the L2
        // atomic units from serialising: each CTA starts at its own 16-column chunk (the CTAs finish together; without the
        // rotation all of them hit the same addresses in the same order), and with unit cout stride a chunk goes out as four
        // 16-byte vector reductions instead of sixteen scalar ones.
        const int cpn = N / 16, chunks = p.nrow * cpn;
        const bool vec = p.co_stride == 1 && (p.cout_g & 3) == 0 && (p.ci_stride & 3) == 0 && (p.tap_stride & 3) == 0 &&
                         ((reinterpret_cast<uintptr_t>(p.dw) & 15) == 0);
        int idx = (int)((blockIdx.x * 7u + blockIdx.y * 3u) % (unsigned)chunks);
#pragma unroll 1
        for (int k = 0; k < chunks; ++k, idx = (idx + 1 == chunks ? 0 : idx + 1)) {
            const int dy = idx / cpn, c = (idx - dy * cpn) * 16;       // N block dy holds the tap row (nrow - 1 - dy)
            float* base = p.dw + (size_t)((p.nrow - 1 - dy) * p.nrow + j) * p.tap_stride + (size_t)ci * p.ci_stride;
            uint32_t r[16];
            tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + dy * N + c, r);
            tc::tmem_ld_wait();
            if (!row_ok) continue;
            if (vec) {
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    if (c + q < p.cout_g) {
                        float* dst = base + (size_t)(g * p.cout_g + c + q);
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(r[q])), "f"(__uint_as_float(r[q + 1])),
                                     "f"(__uint_as_float(r[q + 2])), "f"(__uint_as_float(r[q + 3])) : "memory");
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int co = c + q;
                    if (co < p.cout_g) atomicAdd(base + (size_t)(g * p.cout_g + co) * p.co_stride, __uint_as_float(r[q]));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

int ws_act_tmap(CUtensorMap* out, const tbi_view& v, int n, int kc, int bw, int bh) {
    const uint64_t px = (uint64_t)v.cstride * 2;
    uint64_t dims[5] = {(uint64_t)v.c, (uint64_t)v.w, 1, (uint64_t)v.h, (uint64_t)n};
    uint64_t strides[4] = {px, px * v.w, px * v.w, px * v.w * v.h};
    uint32_t box[5] = {(uint32_t)kc, (uint32_t)bw, 1u, (uint32_t)bh, 1u};
    return tbi_make_tmap_bf16(out, (char*)v.ptr + (size_t)v.coff * 2, 5, dims, strides, box, kc * 2);
}

int pow2_ge16(int c) { int k = 16; while (k < c) k <<= 1; return k; }

template <int NB>
int launch_ws(const WsParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tapwgrad_small_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attr_err != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    tapwgrad_small_kernel<NB><<<grid, WS_THREADS, smem, s>>>(p);
    TBI_CUDA_LAUNCH_CHECK("tapwgrad_small");
    return TBI_OK;
}

}  // namespace

bool tbi_tapwgrad_small_supported(const tbi_tapwgrad* d) {
    static const bool disabled = getenv("TBI_TC_NO_WSMALL") != nullptr;
    if (disabled) return false;
    if (d->dtype != TBI_BF16 || d->a_stride != 1 || d->b_stride != 1 || d->a_src[1].ptr) return false;
    if (d->cin_g > 32 || d->cin_g % 8 != 0 || d->cout_g > 64 || d->cout_g % 8 != 0) return false;
    if (d->gh < 16 || d->gw < 8) return false;
    if (d->ntaps != 1 && d->ntaps != 9) return false;
    const int k = d->ntaps == 9 ? 3 : 1;
    for (int t = 0; t < d->ntaps; ++t) {
        if (d->b_dy[t] != 0 || d->b_dx[t] != 0) return false;
        if (d->a_dy[t] != t / k - k / 2 || d->a_dx[t] != t % k - k / 2) return false;      // row-major 3x3, dilation 1
    }
    return true;
}

int tbi_tapwgrad_small(const tbi_tapwgrad* d, cudaStream_t s) {
    TBI_CHECK(tbi_tapwgrad_small_supported(d), TBI_ERR_UNSUPPORTED, "tapwgrad_small: unsupported shape");
    WsParams p; memset(&p, 0, sizeof(p));
    p.n = d->n; p.gh = d->gh; p.gw = d->gw;
    p.tiles_x = (d->gw + WTW - 1) / WTW; p.tiles_y = (d->gh + WTH - 1) / WTH; p.total_tiles = d->n * p.tiles_x * p.tiles_y;
    p.cin_g = d->cin_g; p.cout_g = d->cout_g; p.groups = d->groups;
    p.kca = pow2_ge16(d->cin_g); p.kcb = pow2_ge16(d->cout_g);
    p.halo = d->ntaps == 9 ? 1 : 0; p.nrow = d->ntaps == 9 ? 3 : 1;
    const int bw = WTW + 2 * p.halo, bh = WTH + 2 * p.halo;
    int rc = ws_act_tmap(&p.a, d->a_src[0], d->n, p.kca, bw, WTH); if (rc) return rc;
    rc = ws_act_tmap(&p.b, d->b_src, d->n, p.kcb, WTW, bh); if (rc) return rc;
    p.a_tx = bw * WTH * p.kca * 2; p.b_tx = WTW * bh * p.kcb * 2;
    // the packed M blocks read up to 128/kca - 1 pixel rows past the halo box: keep that slack inside the stage
    p.a_stage_bytes = (p.a_tx + (128 / p.kca) * p.kca * 2 + 1023) & ~1023;
    p.b_stage_bytes = (p.b_tx + 1023) & ~1023;
    const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    int stages = (100 * 1024) / stage_bytes; if (stages > 8) stages = 8; if (stages < 2) stages = 2;
    p.stages = stages;
    p.dw = d->dw; p.tap_stride = d->tap_stride; p.ci_stride = d->ci_stride; p.co_stride = d->co_stride;
    p.dbias = d->dbias;                                    // column sums of dz ride along (the epilogue warps are idle in the main loop)
    int ctas = 2 * tbi_sm_count() / (d->groups > 0 ? d->groups : 1); if (ctas < 1) ctas = 1;
    if (ctas > p.total_tiles) ctas = p.total_tiles;
    p.tiles_per_cta = (p.total_tiles + ctas - 1) / ctas;
    ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    dim3 grid((unsigned)ctas, (unsigned)d->groups);
    switch (p.kcb / 16) {
        case 1:  rc = launch_ws<1>(p, grid, smem, s); break;
        case 2:  rc = launch_ws<2>(p, grid, smem, s); break;
        default: rc = launch_ws<4>(p, grid, smem, s); break;
    }
    return rc;
}
