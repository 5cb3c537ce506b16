// 3x3 stride-1 convolutions with FEW channels (cin per group <= 64, cout per group <= 64: the full-resolution stem, the
// stage-1 layers and their data gradients) as a persistent tcgen05 kernel that packs the three HORIZONTAL taps into the N
// dimension of the MMA.
//
// Why: these layers are HBM-bound on paper (32 -> 32 channels at 64 x 256 x 256: 537 MB, 82 us) but ran at 188 us in the halo
// kernel (tapgemm_halo.cu).  That kernel issues one MMA per (tap, 16-channel K step): 18 MMAs of M128 x N32 x K16 per 128
// pixels.  Each of them reads its 4 KB pixel operand from shared memory for 16 cycles of tensor-pipe work, and the measured
// rate was ~84 cycles per MMA with two CTAs per SM: ~1 500 cycles per 128 pixels per SM against ~700 for the HBM traffic.
// Here the pixel operand of a K step is read ONCE for all three dx taps:
//
//   D[q, (j, co)] = sum_{dy, ci} x[q + dy*W][ci] * w[dy][dx = j-1][ci][co]          N = 3*cout, K = 3*cin  (3 * cin/16 MMAs)
//   out[p, co]    = D[p-1, (0, co)] + D[p, (1, co)] + D[p+1, (2, co)]                done in the epilogue
//
// A tile is 8 rows x 16 columns of pixels q (M = 128; TMEM lane = 16*row + column), of which the inner 14 columns produce
// outputs: the column shift of the epilogue is then a warp shuffle by one lane (a warp holds two 16-pixel rows, the shifted-in
// values of columns 1..14 always come from the same row), and tiles advance by 14 columns.  The vertical taps are three UMMA
// descriptors into one (8+2) x 16 halo box, each starting a whole 16-pixel row later (8-row swizzle atoms stay aligned).
// MMAs per 112 output pixels: 6 (cin 32) or 3 (cin 16) instead of 18 / 9 per 128; shared-memory operand bytes per output pixel
// 0.37 KB instead of 0.70 KB.
//
// Weights (9 * cin * cout bf16, <= 72 KB) stay resident in shared memory for the CTA's lifetime, one CTA per SM owns all 512
// TMEM columns (four accumulator buffers), warps: 0..15 epilogue (four groups of four, group g takes every fourth tile), 16 TMA
// producer, 17 MMA issuer.  The epilogue is the shared fused one (tc_epilogue.cuh): bias / BN fold / activation / residual / activation
// derivative / split outputs all work as in the halo kernel.
#include "tbi_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int XP_EPI_WARPS = 16;
constexpr int XP_THREADS = 64 + 32 * XP_EPI_WARPS;
constexpr int XP_STRIDE = 128;                          // TMEM columns per accumulator buffer (N = 3*cout <= 96)
constexpr uint32_t XP_NBUF = 4, XP_LGB = 2;
constexpr int XTW = 16, XTH = 8, XOUT = 14;             // tile of q pixels; output columns per tile

struct alignas(64) XpParams {
    CUtensorMap a, b;
    int n, gh, gw, tiles_x, tiles_y, m_tiles, groups;
    int cin_g, cout_g, cout_total;
    int ksteps, row_bytes;
    int a_stages, a_stage_bytes, a_tx, b_blk_bytes, b_tx;
    int kcol[3][3];                 // [dy+1][dx+1] -> column block (tap index) of the weight pack
    int simple_ctx;                 // epilogue tensors share the conv's geometry: per-tile pointers from one pixel index
    tbi_epilogue epi;
};

__device__ __forceinline__ void xp_decode(const XpParams& p, int mt, int& x0, int& y0, int& n0) {
    const int tix = mt % p.tiles_x; mt /= p.tiles_x;
    const int tiy = mt % p.tiles_y; n0 = mt / p.tiles_y;
    x0 = tix * XOUT - 1; y0 = tiy * XTH;
}

template <int NC, int ACT, int DACT>          // NC = cout_g / 16
__device__ __forceinline__ void xp_epilogue(const XpParams& p, const float* sbias, uint32_t tmem_base, uint64_t* t_full, uint64_t* t_empty,
                                            int cg, int it_first, int it_stride, int warp, int lane) {
    // Four groups of four warps (one warp per TMEM lane quadrant); group g owns accumulator buffer g, i.e. every fourth tile.
    // The per-tile chain (barrier wait -> tcgen05.ld -> shuffles -> side-input loads -> math -> stores) is latency-bound; with two
    // groups the kernel ran at ~1 700 cycles per tile per SM, the barrier wake-up alone costing about a microsecond per tile.
    constexpr int COUT = 16 * NC, CH = 16;
    const int q = warp & 3, grp = warp >> 2;
    const int m = q * 32 + lane;
    const int xx = m & (XTW - 1), yy = m >> 4;
    // this group's tiles are it_first + grp*it_stride + k*(4*it_stride): decode the first one, then step (tix, tiy, n) by a fixed
    // increment with carries -- no division in the loop
    const int step = XP_NBUF * it_stride;
    const int s_x = step % p.tiles_x, s_y = (step / p.tiles_x) % p.tiles_y, s_n = step / (p.tiles_x * p.tiles_y);
    int i = it_first + grp * it_stride;
    int tix = i % p.tiles_x, tiy = (i / p.tiles_x) % p.tiles_y, n0 = i / (p.tiles_x * p.tiles_y);
    const uint32_t buf = (uint32_t)grp;
    const tbi_epilogue& e = p.epi;
    for (uint32_t k = 0; i < p.m_tiles; i += step, ++k) {
        const int gx = tix * XOUT - 1 + xx, gy = tiy * XTH + yy;
        const bool valid = xx >= 1 && xx <= XOUT && gx < p.gw && gy < p.gh;
        // two-register row context (tc_epilogue.cuh): every epilogue tensor shares the output's pixel grid
        const LeanRowCtx rc = make_lean_row_ctx(e, n0, gy + e.out_off_y, gx + e.out_off_x, e.bias ? sbias : nullptr);
        tix += s_x; if (tix >= p.tiles_x) { tix -= p.tiles_x; ++tiy; }
        tiy += s_y; if (tiy >= p.tiles_y) { tiy -= p.tiles_y; ++n0; }
        n0 += s_n;
        const uint32_t acc_it = (uint32_t)grp + XP_NBUF * k;
        tc::mbar_wait_bounded<true>(&t_full[buf], (acc_it >> XP_LGB) & 1u);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + buf * XP_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < COUT; c += CH) {
            uint32_t d0[CH], d1[CH], d2[CH];
            tc::tmem_ld16(taddr + c, d0); tc::tmem_ld16(taddr + COUT + c, d1); tc::tmem_ld16(taddr + 2 * COUT + c, d2);
            tc::tmem_ld_wait();
            if (c + CH >= COUT) {                                         // last read of this buffer: hand it back to the MMA warp
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&t_empty[buf]);
            }
            // out[p] = D[p-1][dx block 0] + D[p][block 1] + D[p+1][block 2]: neighbours of columns 1..14 are lanes m-1 / m+1
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const float l = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[k]), 1);
                const float r = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[k]), 1);
                d1[k] = __float_as_uint(__uint_as_float(d1[k]) + l + r);
            }
            if (valid) epilogue_cols<ACT, DACT, CH>(rc, d1, c, p.cout_g, cg * p.cout_g);
        }
    }
}

template <int NC>
__global__ void __launch_bounds__(XP_THREADS, 1) tapgemm_xpack_kernel(const __grid_constant__ XpParams p) {
    constexpr int STRIDE = XP_STRIDE, NBUF = (int)XP_NBUF;
    constexpr uint32_t LGB = XP_LGB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_res = smem + (size_t)p.a_stages * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_res + 9 * (size_t)p.b_blk_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + p.a_stages;
    uint64_t* t_full = a_empty + p.a_stages;               // [4]
    uint64_t* t_empty = t_full + 4;                        // [4]
    uint64_t* b_bar = t_empty + 4;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(b_bar + 1);
    float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tslot + 4) + 15) & ~static_cast<uintptr_t>(15));
    if (p.epi.bias)
        for (int i = threadIdx.x; i < p.cout_total; i += XP_THREADS) sbias[i] = p.epi.bias[i];

    const int cg = (int)(blockIdx.x % p.groups);
    const int it_first = (int)(blockIdx.x / p.groups), it_stride = (int)(gridDim.x / p.groups);
    constexpr int W_TMA = XP_EPI_WARPS, W_MMA = XP_EPI_WARPS + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == W_TMA && lane == 0) {
        tc::prefetch_tmap(&p.a); tc::prefetch_tmap(&p.b);
        for (int s = 0; s < p.a_stages; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NBUF; ++s) { tc::mbar_init(&t_full[s], 1); tc::mbar_init(&t_empty[s], 4); }
        tc::mbar_init(b_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == W_MMA) tc::tmem_alloc<512>(tslot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tslot;

    if (warp == W_TMA) {
        // ===================== TMA producer =====================
        const bool leader = tc::elect_one();
        if (it_first < p.m_tiles && leader) {                 // the group's nine weight blocks, once: [dy][dx][cout rows][cin]
            tc::mbar_expect_tx(b_bar, 9u * (uint32_t)p.b_tx);
            for (int by = 0; by < 3; ++by)
                for (int j = 0; j < 3; ++j)
                    tc::tma_load_2d(b_res + (size_t)(by * 3 + j) * p.b_blk_bytes, &p.b, b_bar, p.kcol[by][j] * p.cin_g, cg * p.cout_g);
        }
        uint32_t sa = 0, a_par = 1;
        for (int i = it_first; i < p.m_tiles; i += it_stride) {
            int x0, y0, n0;
            xp_decode(p, i, x0, y0, n0);
            tc::mbar_wait_bounded(&a_empty[sa], a_par);
            if (leader) {
                tc::mbar_expect_tx(&a_full[sa], (uint32_t)p.a_tx);
                tc::tma_load_4d(a_ring + (size_t)sa * p.a_stage_bytes, &p.a, &a_full[sa], cg * p.cin_g, x0, y0 - 1, n0);
            }
            if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::make_idesc_bf16(128, 3 * 16 * NC, 0, 0);
        const uint32_t layout = p.row_bytes == 128 ? 2u : p.row_bytes == 64 ? 4u : 6u;
        const uint64_t dbase = tc::smem_desc_base(16, 8u * (uint32_t)p.row_bytes, layout);      // K-major, 8-row groups back to back
        const uint32_t d_lo0 = (uint32_t)dbase, d_hi = (uint32_t)(dbase >> 32);
        const uint32_t a_ring_lo = (tc::smem_u32(a_ring) & 0x3FFFFu) >> 4, a_stage_lo = (uint32_t)p.a_stage_bytes >> 4;
        const uint32_t b_lo0 = d_lo0 + ((tc::smem_u32(b_res) & 0x3FFFFu) >> 4);
        const uint32_t a_row_lo = (uint32_t)(XTW * p.row_bytes) >> 4;            // one 16-pixel box row = one vertical tap step
        const uint32_t b_dy_lo = (uint32_t)(3 * p.b_blk_bytes) >> 4;
        if (it_first < p.m_tiles) tc::mbar_wait_bounded(b_bar, 0);
        uint32_t sa = 0, a_par = 0, acc_it = 0;
        for (int i = it_first; i < p.m_tiles; i += it_stride, ++acc_it) {
            const uint32_t buf = acc_it & (NBUF - 1u);
            tc::mbar_wait_bounded(&t_empty[buf], ((acc_it >> LGB) & 1u) ^ 1u);
            tc::mbar_wait_bounded(&a_full[sa], a_par);
            tc::tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * STRIDE;
            const uint32_t a_lo = d_lo0 + a_ring_lo + sa * a_stage_lo;
            if (leader) {
#pragma unroll
                for (int by = 0; by < 3; ++by) {
#pragma unroll 4
                    for (int k = 0; k < p.ksteps; ++k)
                        tc::umma_bf16_lh(tmem_d, a_lo + by * a_row_lo + 2 * k, d_hi, b_lo0 + by * b_dy_lo + 2 * k, d_hi, idesc, (by | k) ? 1u : 0u);
                }
                tc::umma_commit(&a_empty[sa]);
                tc::umma_commit(&t_full[buf]);
            }
            if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
        }
        __syncwarp();
    } else {
        TBI_EPI_DISPATCH(p.epi.act, p.epi.dact, (xp_epilogue<NC, A_, D_>(p, sbias, tmem_base, t_full, t_empty, cg, it_first, it_stride, warp, lane)));
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tc::tmem_dealloc<512>(tmem_base);
}

template <int NC>
int launch_xp(const XpParams& p, int grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tapgemm_xpack_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    if (attr_err != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    tapgemm_xpack_kernel<NC><<<grid, XP_THREADS, smem, s>>>(p);
    TBI_CUDA_LAUNCH_CHECK("tapgemm_xpack");
    return TBI_OK;
}

inline uint32_t xr1024(uint32_t x) { return (x + 1023u) & ~1023u; }

bool xp_tap_table(const tbi_tapgemm* d, int (&kcol)[3][3]) {
    if (d->ntaps != 9) return false;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) kcol[a][b] = -1;
    for (int t = 0; t < 9; ++t) {
        const int dy = d->dy[t], dx = d->dx[t];
        if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || kcol[dy + 1][dx + 1] >= 0) return false;
        kcol[dy + 1][dx + 1] = t;
    }
    return true;
}

}  // namespace

// precondition: tbi_tapgemm_tc_supported(d)
bool tbi_tapgemm_xpack_supported(const tbi_tapgemm* d) {
    // Measured (64 x 256 x 256, scratch/mb_conv.py): 32 -> 32 forward 172 us here vs 196 us in the halo kernel, its data gradient
    // 209 vs 213 us, 16 -> 32 forward 157 vs 141 us.  The shared-memory operand traffic is halved as designed, but the epilogue's
    // column shift (two shuffles + two adds per output value) makes the kernel issue-bound (ncu: 107 M warp instructions, issue
    // slots 56-60 % busy, tensor pipe 25 %).  Default: only the shape it wins on; TBI_TC_XPACK=all / =off override.
    static const char* mode = getenv("TBI_TC_XPACK");
    const bool all = mode && !strcmp(mode, "all"), off = mode && !strcmp(mode, "off");
    if (off || d->dtype != TBI_BF16) return false;
    if (!all && !(d->cin_g == 32 && d->cout_g == 32 && d->groups == 1 && d->epi.dact == TBI_ACT_NONE)) return false;
    if (d->in_stride != 1 || d->nphase > 1 || d->src[1].ptr) return false;
    if (d->cin_g != 16 && d->cin_g != 32 && d->cin_g != 64) return false;
    if (d->cout_g != 16 && d->cout_g != 32) return false;              // N = 3*cout <= 96: four accumulator buffers of 128 columns
    if (d->gw < 16 || d->gh < 8) return false;
    if (d->epi.out_stride > 1 || d->epi.out_f32) return false;
    if (tbi_tc_narrow(d) || tbi_tc_f32wide(d)) return false;
    int kcol[3][3];
    if (!xp_tap_table(d, kcol)) return false;
    const long long w_bytes = 9LL * d->cout_g * d->cin_g * 2;
    const long long stage = xr1024((uint32_t)((XTH + 2) * XTW * d->cin_g * 2));
    return w_bytes + 3 * stage + 4096 + (long long)d->cout_g * d->groups * 4 <= 200 * 1024;
}

int tbi_tapgemm_xpack(const tbi_tapgemm* d, cudaStream_t s) {
    TBI_CHECK(tbi_tapgemm_xpack_supported(d), TBI_ERR_UNSUPPORTED, "tapgemm_xpack: unsupported shape");
    XpParams p; memset(&p, 0, sizeof(p));
    xp_tap_table(d, p.kcol);
    p.n = d->n; p.gh = d->gh; p.gw = d->gw; p.groups = d->groups;
    p.tiles_x = (d->gw + XOUT - 1) / XOUT; p.tiles_y = (d->gh + XTH - 1) / XTH; p.m_tiles = d->n * p.tiles_x * p.tiles_y;
    p.cin_g = d->cin_g; p.cout_g = d->cout_g; p.cout_total = d->cout_g * d->groups;
    p.ksteps = d->cin_g / 16; p.row_bytes = d->cin_g * 2;
    p.epi = d->epi;
    {
        const tbi_epilogue& e = d->epi;
        auto same = [&](const tbi_view& v) { return v.ptr == nullptr || (v.h == d->gh && v.w == d->gw); };
        p.simple_ctx = (e.split_c == 0 && !e.drop_keep && !e.dact_keep && !e.out2.ptr && !e.residual2.ptr && e.out_off_y == 0 && e.out_off_x == 0 &&
                        same(e.out) && same(e.residual) && (e.dact == TBI_ACT_NONE || same(e.dact_ref))) ? 1 : 0;
    }
    {
        const tbi_view& v = d->src[0];
        const uint64_t px = (uint64_t)v.cstride * 2;
        uint64_t d4[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)d->n};
        uint64_t s4[3] = {px, px * v.w, px * v.w * v.h};
        uint32_t b4[4] = {(uint32_t)d->cin_g, (uint32_t)XTW, (uint32_t)(XTH + 2), 1u};
        int rc = tbi_make_tmap_bf16(&p.a, (char*)v.ptr + (size_t)v.coff * 2, 4, d4, s4, b4, p.row_bytes);
        if (rc) return rc;
        const uint64_t K = 9ull * d->cin_g;
        uint64_t dims[2] = {K, (uint64_t)p.cout_total};
        uint64_t strides[1] = {K * 2};
        uint32_t box[2] = {(uint32_t)d->cin_g, (uint32_t)d->cout_g};
        rc = tbi_make_tmap_bf16(&p.b, const_cast<void*>(d->w), 2, dims, strides, box, p.row_bytes);
        if (rc) return rc;
    }
    p.a_tx = (XTH + 2) * XTW * p.row_bytes; p.a_stage_bytes = (int)xr1024((uint32_t)p.a_tx);
    p.b_tx = d->cout_g * p.row_bytes; p.b_blk_bytes = p.b_tx;                  // a multiple of the 8-row swizzle atom (cout_g % 8 == 0)
    const long long fixed = 9LL * p.b_blk_bytes + 4096 + (long long)p.cout_total * 4;
    int as = (int)((200 * 1024 - fixed) / p.a_stage_bytes);
    if (as > 10) as = 10;
    p.a_stages = as;
    int per_group = tbi_sm_count() / d->groups; if (per_group < 1) per_group = 1;
    if (per_group > p.m_tiles) per_group = p.m_tiles;
    const int grid = per_group * d->groups;
    const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + 9 * (size_t)p.b_blk_bytes + 1024 + 512 + ((size_t)p.cout_total * 4 + 64) +
                        (size_t)(2 * p.a_stages + 16) * 8;
    return d->cout_g == 16 ? launch_xp<1>(p, grid, smem, s) : launch_xp<2>(p, grid, smem, s);
}
