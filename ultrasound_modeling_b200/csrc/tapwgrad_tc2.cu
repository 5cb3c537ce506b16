// tcgen05 weight-gradient tap-GEMM on CTA PAIRS (cta_group::2, M = 256) for sm_100a.
//
//   dw[tap][ci][co] += sum_pixels a[pixel + a_off(tap)][ci] * b[pixel + b_off(tap)][co]
//
// Why pairs.  Both operands of this GEMM come from shared memory (K = pixels, so both are MN-major TMA boxes).  A
// 128 x 128 x 16 MMA of one CTA reads 4 KB of A and 4 KB of B per 64 tensor-pipe cycles = 128 B/clk, the whole shared-memory
// bandwidth of an SM, and the TMA fills of the next stage (80 KB per 1024 cycles in tapwgrad_tc.cu) have to fit in beside
// them: ncu shows the tensor pipe 54-58 % active with L2 and DRAM far from their limits (profiles/r2_wgrad.md).  With
// cta_group::2 one instruction computes a 256 x 128 tile across the two SMs of a pair: each CTA supplies ITS 128 rows of the
// M operand and HALF of the N operand (64 of the 128 columns), so per CTA the MMAs read 6 KB per 64 cycles (96 B/clk) and a
// stage is 48 KB instead of 80 KB.
//
// Orientation.  T taps of a group share one operand's pixels (a transposed conv's taps all read the same input pixels; a
// stride-1 conv's taps all read the same output-gradient pixels).  The shared operand is made the M side (each CTA loads its
// 128 channels once per K step), the per-tap operand the N side (each CTA loads 64 channels per tap), whatever that means
// for ci/co: the host swaps the roles and the output strides accordingly.  T accumulators of 128 columns sit side by side
// in TMEM (T = 4: all 512 columns, one CTA per SM).
//
// Pipeline (per CTA): warp 0 = TMA producer (both CTAs load; every load signals the LEADER's full barrier through the
// .cta_group::2 form of cp.async.bulk.tensor), warp 1 = TMEM allocation in both CTAs + single-thread MMA issue in the leader
// (tcgen05.commit ... multicast::cluster releases the stage in both CTAs), warps 2..5 = epilogue (each CTA reduces its own 128
// rows into the fp32 gradient with red.global.add; consecutive lanes = consecutive M channels, which the host makes the
// unit-stride axis of dw whenever it can).
#include "tbi_common.cuh"
#include "tc_common.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int W2_THREADS = 192;
constexpr int KP2 = 64;                       // pixels per K step
constexpr uint32_t BOX2 = KP2 * 128;          // one [64 px x 64 ch] bf16 box

struct Operand {
    CUtensorMap map[2];           // virtual concat of two sources (second = first when single)
    int c0;                       // channels of the first source (>= channels: single source)
    int cbase, cpix;              // stride-2 parity view: channel coordinate = cbase + ch + ax * cpix
    signed char qy[16], qx[16], ay[16], ax[16];
};

struct alignas(64) Wgrad2Params {
    Operand m, n;                 // m: shared by the taps of a group, 256 channels per CTA pair; n: per tap, 128 per pair
    int ltw, lth;                 // pixel tile: tw x th x tn = 64
    int tiles_x, tiles_y, tiles_b;
    int cm, cn;                   // channel counts of the two operands
    int m_pairs, n_tiles, ntaps;
    int stages, tiles_per_cta;
    float* dw;
    long long tap_stride, m_stride, n_stride;
};

using tc::cluster_ctarank; using tc::cluster_sync_all; using tc::map_to_cta; using tc::tma_load_5d_pair;
using tc::tmem_alloc_pair; using tc::tmem_dealloc_pair; using tc::umma_bf16_pair; using tc::umma_commit_pair;

__device__ __forceinline__ void w2_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!tc::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("tbi tcgen05 wgrad (pair): mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
    }
}

__host__ __device__ constexpr int w2_pow2(int v) { return v <= 32 ? 32 : v <= 64 ? 64 : v <= 128 ? 128 : v <= 256 ? 256 : 512; }

// T <= 2: 256 TMEM columns and <= ~100 KB of shared memory per CTA, so TWO pairs share an SM pair and one pair's reduction
// epilogue (T*128*128 red.global.add per CTA, worth ~16-32 K steps) runs under the other's main loop; with T = 4 the
// accumulators fill TMEM and every CTA's epilogue is exposed (ncu: tensor pipe 48 % active on upsample_2's weight gradient,
// exactly main-loop cycles / elapsed cycles).
template <int T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W2_THREADS, T <= 2 ? 2 : 1) tapwgrad_pair_kernel(const __grid_constant__ Wgrad2Params p) {
    constexpr int TCOLS = w2_pow2(T * 128);
    constexpr int M_BOXES = 2, N_BOXES = T;                    // per CTA per stage: own 128 M channels, own 64 N channels of each tap
    constexpr uint32_t STAGE_BYTES = (M_BOXES + N_BOXES) * BOX2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int stages = p.stages;
    uint8_t* bar_base = smem + (size_t)stages * STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);    // used in the leader CTA only
    uint64_t* empty = full + stages;
    uint64_t* tfull = empty + stages;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader_cta = rank == 0;
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.m.map[0]); tc::prefetch_tmap(&p.n.map[0]);
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<TCOLS>(tslot);
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                         // both CTAs' barriers exist before anything is signalled across the pair
    tc::tc_fence_after();
    const uint32_t tmem_base = *tslot;

    int t = blockIdx.x >> 1;
    const int tap = (t % (p.ntaps / T)) * T; t /= (p.ntaps / T);           // first tap of this pair's group
    const int n_t = t % p.n_tiles; const int m_pair = t / p.n_tiles;
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    const int tile_beg = blockIdx.y * p.tiles_per_cta;
    const int tile_end = min(total_tiles, tile_beg + p.tiles_per_cta);
    const int iters = tile_end - tile_beg;
    const int tw = 1 << p.ltw, th = 1 << p.lth, tn = KP2 >> (p.ltw + p.lth);
    const int m0 = (m_pair * 2 + (int)rank) * 128;             // this CTA's M channels
    const int n0c = n_t * 128 + (int)rank * 64;                // this CTA's half of the N tile

    if (warp == 0) {
        if (lane == 0 && iters > 0) {
            // ===== TMA producer (both CTAs) =====
            int msrc[2], mch[2];
            for (int j = 0; j < 2; ++j) {
                int ch = m0 + 64 * j, src = 0;
                if (ch >= p.m.c0 && p.m.c0 < p.cm) { src = 1; ch -= p.m.c0; }
                msrc[j] = src; mch[j] = ch;
            }
            int nsrc = 0, nch = n0c;
            if (nch >= p.n.c0 && p.n.c0 < p.cn) { nsrc = 1; nch -= p.n.c0; }
            for (int it = 0; it < iters; ++it) {
                const int s = it % stages;
                w2_wait(&empty[s], (((uint32_t)(it / stages)) & 1u) ^ 1u);
                int tt = tile_beg + it;
                const int tix = tt % p.tiles_x; tt /= p.tiles_x;
                const int tiy = tt % p.tiles_y; const int tib = tt / p.tiles_y;
                const int x0 = tix * tw, y0 = tiy * th, nb0 = tib * tn;
                uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                if (leader_cta) tc::mbar_expect_tx(&full[s], 2u * STAGE_BYTES);          // the pair's bytes land on the leader's barrier
                const uint32_t bar = map_to_cta(tc::smem_u32(&full[s]), 0);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    tma_load_5d_pair(st + j * BOX2, &p.m.map[msrc[j]], bar, p.m.cbase + mch[j] + (int)p.m.ax[tap] * p.m.cpix,
                                     x0 + (int)p.m.qx[tap], (int)p.m.ay[tap], y0 + (int)p.m.qy[tap], nb0);
#pragma unroll
                for (int t2 = 0; t2 < T; ++t2)
                    tma_load_5d_pair(st + (M_BOXES + t2) * BOX2, &p.n.map[nsrc], bar, p.n.cbase + nch + (int)p.n.ax[tap + t2] * p.n.cpix,
                                     x0 + (int)p.n.qx[tap + t2], (int)p.n.ay[tap + t2], y0 + (int)p.n.qy[tap + t2], nb0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (leader_cta && lane == 0 && iters > 0) {
            // ===== MMA issuer (leader CTA, one thread): D[256 x 128] += M^T[256 x px] * N[px x 128], both MN-major =====
            const uint32_t idesc = tc::make_idesc_bf16(256, 128, 1, 1);
            for (int it = 0; it < iters; ++it) {
                const int s = it % stages;
                w2_wait(&full[s], ((uint32_t)(it / stages)) & 1u);
                tc::tc_fence_after();
                const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint32_t b_addr = a_addr + M_BOXES * BOX2;
#pragma unroll
                for (int t2 = 0; t2 < T; ++t2) {
#pragma unroll
                    for (int k = 0; k < KP2 / 16; ++k) {
                        // 16 pixels = 2 swizzle atoms along K: +2048 B per step; LBO = box (next 64 channels), SBO = 1024 B
                        const uint64_t da = tc::make_smem_desc(a_addr + k * 2048, BOX2, 1024, 2u);
                        const uint64_t db = tc::make_smem_desc(b_addr + t2 * BOX2 + k * 2048, BOX2, 1024, 2u);
                        umma_bf16_pair(tmem_base + t2 * 128, da, db, idesc, (it | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit_pair(&empty[s]);
            }
            umma_commit_pair(tfull);
        }
        __syncwarp();
    } else if (iters > 0) {
        // ===== epilogue (both CTAs): lane = M channel, columns = N channels of the whole 128-wide tile =====
        const int q = warp & 3;
        const int cm = m0 + q * 32 + lane;
        const bool row_ok = cm < p.cm;
        w2_wait(tfull, 0);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        // the split-K CTAs of one output tile finish together: each starts at its own 32-column chunk so that they do not send
        // their reductions to the same words in the same order
        constexpr int CHUNKS = T * 4;
        int idx = (int)((blockIdx.x * 5u + blockIdx.y * 3u + blockIdx.z) % (unsigned)CHUNKS);
#pragma unroll 1
        for (int k = 0; k < CHUNKS; ++k, idx = (idx + 1 == CHUNKS ? 0 : idx + 1)) {
            const int t2 = idx >> 2, c = (idx & 3) * 32;
            float* base = p.dw + (size_t)(tap + t2) * p.tap_stride + (size_t)cm * p.m_stride;
            uint32_t r[32];
            tc::tmem_ld32(taddr + t2 * 128 + c, r);
            tc::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int cn = n_t * 128 + c + j;
                    if (cn < p.cn) atomicAdd(base + (size_t)cn * p.n_stride, __uint_as_float(r[j]));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                         // the peer may still be signalling this CTA's barriers / reading its operands
    if (warp == 1) tmem_dealloc_pair<TCOLS>(tmem_base);
}

int w2_ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// 5-D activation map with a [64 ch x tw x 1 x th x tn] box (as wg_act_tmap in tapwgrad_tc.cu)
int w2_act_tmap(CUtensorMap* out, const tbi_view& v, int n, int stride, int tw, int th, int tn, int* cbase, int* cpix) {
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

template <int T>
int launch_w2(const Wgrad2Params& p, dim3 grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tapwgrad_pair_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024); });
    if (attr_err != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    tapwgrad_pair_kernel<T><<<grid, W2_THREADS, smem, s>>>(p);
    TBI_CUDA_LAUNCH_CHECK("tapwgrad_pair");
    return TBI_OK;
}

}  // namespace

// tbi_wgrad_pair_plan (tbi_common.cuh): which operand is the M side, how many taps per group, and which 128-channel tiles of
// the M operand the pair launch covers ([0, 2*m_pairs)); the caller runs the single-CTA kernel on what is left.
tbi_wgrad_pair_plan tbi_tapwgrad_pair_plan(const tbi_tapwgrad* d) {
    tbi_wgrad_pair_plan pl; memset(&pl, 0, sizeof(pl));
    static const bool off = getenv("TBI_WGRAD_NO_PAIR") != nullptr;
    if (off || d->groups != 1 || d->a_stride != 1 || d->ntaps < 1) return pl;
    if (d->cin_g < 64 || d->cout_g < 64) return pl;
    bool a_same = true, b_same = true;
    for (int t = 1; t < d->ntaps; ++t) {
        a_same = a_same && d->a_dy[t] == d->a_dy[0] && d->a_dx[t] == d->a_dx[0];
        b_same = b_same && d->b_dy[t] == d->b_dy[0] && d->b_dx[t] == d->b_dx[0];
    }
    int T = 1;
    if (d->ntaps > 1) {
        if (!(a_same || b_same)) return pl;
        static const int t_pref = getenv("TBI_WGRAD_PAIR_T") ? atoi(getenv("TBI_WGRAD_PAIR_T")) : 2;
        T = (d->ntaps % 4 == 0 && t_pref == 4) ? 4 : d->ntaps % 3 == 0 ? 3 : d->ntaps % 2 == 0 ? 2 : 1;
        if (T == 1) return pl;
    }
    // the M side is the operand the taps share; with one tap either will do: prefer the one with an even number of 128-tiles
    bool m_is_a;
    if (d->ntaps > 1) m_is_a = a_same;
    else {
        const int at = (d->cin_g + 127) / 128, bt = (d->cout_g + 127) / 128;
        const bool can_a = d->cout_g % 128 == 0 && at >= 2, can_b = d->cin_g % 128 == 0 && bt >= 2;
        if (can_a && can_b) m_is_a = (at % 2 == 0) || (bt % 2 != 0);
        else if (can_a) m_is_a = true;
        else if (can_b) m_is_a = false;
        else return pl;
    }
    const int cm = m_is_a ? d->cin_g : d->cout_g, cn = m_is_a ? d->cout_g : d->cin_g;
    if (cn % 128 != 0) return pl;                               // each CTA loads whole 64-channel boxes of the N operand
    // virtual concat: a 64-channel box must lie inside one source
    if (d->a_src[1].ptr && d->a_src[0].c % 64 != 0) return pl;
    if (d->b_stride == 2 && !m_is_a && d->ntaps > 1) return pl;  // a stride-2 gathered operand is per-tap: it must be the N side
    pl.m_tiles = (cm + 127) / 128;
    if (pl.m_tiles < 2) return pl;
    // an odd last tile is paired with an all-zero one (its CTA's loads are out of bounds = zero-filled, its epilogue writes
    // nothing): for upsample_4's 320 input channels that wastes 3/8 of the pairs' MMAs and is still faster than a second,
    // poorly filled launch of the single-CTA kernel (measured 546 us for pair + rest against 410 us single)
    pl.m_pairs = (pl.m_tiles + 1) / 2;
    pl.m_tiles = 2 * pl.m_pairs;
    pl.ok = 1; pl.m_is_a = m_is_a ? 1 : 0; pl.T = T;
    return pl;
}

// precondition: tbi_tapwgrad_tc_supported(d) and tbi_tapwgrad_pair_plan(d).ok
int tbi_tapwgrad_pair(const tbi_tapwgrad* d, const tbi_wgrad_pair_plan& pl, cudaStream_t s) {
    Wgrad2Params p; memset(&p, 0, sizeof(p));
    int ltw = w2_ilog2_ceil(d->gw); if (ltw > 3) ltw = 3;
    int lth = w2_ilog2_ceil(d->gh); if (lth > 6 - ltw) lth = 6 - ltw;
    const int tw = 1 << ltw, th = 1 << lth, tn = KP2 >> (ltw + lth);
    p.ltw = ltw; p.lth = lth;
    p.tiles_x = (d->gw + tw - 1) / tw; p.tiles_y = (d->gh + th - 1) / th; p.tiles_b = (d->n + tn - 1) / tn;
    p.ntaps = d->ntaps;
    Operand oa, ob; memset(&oa, 0, sizeof(oa)); memset(&ob, 0, sizeof(ob));
    int rc = w2_act_tmap(&oa.map[0], d->a_src[0], d->n, 1, tw, th, tn, &oa.cbase, &oa.cpix); if (rc) return rc;
    if (d->a_src[1].ptr) { int cb, cp; rc = w2_act_tmap(&oa.map[1], d->a_src[1], d->n, 1, tw, th, tn, &cb, &cp); if (rc) return rc; oa.c0 = d->a_src[0].c; }
    else { oa.map[1] = oa.map[0]; oa.c0 = d->cin_g; }
    rc = w2_act_tmap(&ob.map[0], d->b_src, d->n, d->b_stride, tw, th, tn, &ob.cbase, &ob.cpix); if (rc) return rc;
    ob.map[1] = ob.map[0]; ob.c0 = d->cout_g;
    for (int t = 0; t < d->ntaps; ++t) {
        oa.qy[t] = (signed char)d->a_dy[t]; oa.qx[t] = (signed char)d->a_dx[t]; oa.ay[t] = 0; oa.ax[t] = 0;
        if (d->b_stride == 1) { ob.qy[t] = (signed char)d->b_dy[t]; ob.qx[t] = (signed char)d->b_dx[t]; ob.ay[t] = 0; ob.ax[t] = 0; }
        else {
            const int ay = d->b_dy[t] & 1, ax = d->b_dx[t] & 1;
            ob.ay[t] = (signed char)ay; ob.ax[t] = (signed char)ax;
            ob.qy[t] = (signed char)((d->b_dy[t] - ay) / 2); ob.qx[t] = (signed char)((d->b_dx[t] - ax) / 2);
        }
    }
    if (pl.m_is_a) { p.m = oa; p.n = ob; p.cm = d->cin_g; p.cn = d->cout_g; p.m_stride = d->ci_stride; p.n_stride = d->co_stride; }
    else           { p.m = ob; p.n = oa; p.cm = d->cout_g; p.cn = d->cin_g; p.m_stride = d->co_stride; p.n_stride = d->ci_stride; }
    p.dw = d->dw; p.tap_stride = d->tap_stride;
    p.m_pairs = pl.m_pairs; p.n_tiles = p.cn / 128;
    const int T = pl.T;
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    const long long out_pairs = (long long)(d->ntaps / T) * p.m_pairs * p.n_tiles;
    // split-K: whole pairs per wave (one CTA per SM); cost model as in tapwgrad_tc.cu
    const long long pair_slots = (long long)(tbi_sm_count() / 2) * (T <= 2 ? 2 : 1);
    long long cap = total_tiles < 32 ? total_tiles : 32, best = -1, ksplit = 1;
    for (long long ks = 1; ks <= cap; ++ks) {
        const long long pairs = out_pairs * ks, waves = (pairs + pair_slots - 1) / pair_slots;
        const long long cost = waves * ((total_tiles + ks - 1) / ks + 32);
        if (best < 0 || cost < best) { best = cost; ksplit = ks; }
    }
    p.tiles_per_cta = (int)((total_tiles + ksplit - 1) / ksplit);
    ksplit = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const uint32_t stage_bytes = (2 + T) * BOX2;
    int stages = (int)((T <= 2 ? 100 : 200) * 1024 / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages > p.tiles_per_cta) stages = p.tiles_per_cta;
    if (stages < 1) stages = 1;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    dim3 grid((unsigned)(2 * out_pairs), (unsigned)ksplit, 1);
    switch (T) {
        case 4:  return launch_w2<4>(p, grid, smem, s);
        case 3:  return launch_w2<3>(p, grid, smem, s);
        case 2:  return launch_w2<2>(p, grid, smem, s);
        default: return launch_w2<1>(p, grid, smem, s);
    }
}
