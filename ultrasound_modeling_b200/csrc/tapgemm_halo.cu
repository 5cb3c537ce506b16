// Persistent, halo-addressed tcgen05 tap-GEMM (bf16 in, fp32 accumulate in TMEM) for sm_100a.
//
// Same contract as tapgemm_tc.cu, different schedule, for spatial sizes >= 16 x 8:
//   * the 128 output pixels of a tile are a 16 x 8 patch of ONE image;
//   * per 64-channel K chunk the producer issues ONE TMA box covering the patch plus its halo
//     ((16+ey) x (8+ex) pixels); every tap's A operand is a UMMA descriptor into that box, starting at
//     row (dy-oy)*(8+ex) + (dx-ox) with stride-byte-offset (8+ex)*row_bytes.  This works because the
//     hardware swizzle follows absolute shared-memory address bits (profiles/r1_umma_descriptor_probe.md).
//     A 3x3 conv thus reads its input once from L2 instead of 9 times; the 4 parity phases of a k4 s2
//     transposed conv share one 3x3 halo; the 16 taps of its dgrad use 4 halos (one per input parity).
//   * CTAs are persistent (grid = resident CTAs, static round-robin over tiles, N-tile fastest so
//     neighbouring CTAs hit the same halo in L2); two TMEM accumulator buffers let the epilogue of tile
//     i run under the TMA/MMA main loop of tile i+1; separate A (halo) and B (weights) smem rings.
//   * schedules on top of that (chosen per launch on the host, tbi_tapgemm_halo): STREAMED weights (one B stage per chunk and tap),
//     PAIR (cta_group::2: a cluster of two CTAs computes a 256 x 128 tile, each CTA streams half of every weight stage),
//     RESIDENT (a CTA keeps one N tile's whole weight slab in shared memory and walks pixel tiles), ALL-SLABS (a resident CTA
//     keeps EVERY N tile's slab and runs one accumulator per slab against each halo: one-chunk, all-epilogue GEMMs);
//   * the epilogue's side inputs (act' reference, residual, skip-gradient accumulator) are requested by the PRODUCER thread as L2
//     prefetch boxes (cp.async.bulk.prefetch.tensor) one tile ahead of the epilogue that reads them row by row;
//   * nothing in the role loops may touch local memory: with ~208 KB of shared memory per SM the L1 is a few KB and a spill
//     reload is an L2 round trip on the MMA -> epilogue -> MMA chain (profiles/r2_ncu_local_memory.md) -- iteration constants
//     come from the kernel parameters, per-CTA constants are re-read from shared memory, the row context is two registers.
// Warp roles (320 threads): warps 0..7 epilogue (two per TMEM lane quadrant), warp 8 TMA producer, warp 9 MMA issuer + TMEM owner
// (highest ids = highest scheduler priority for the two single-issuer warps).
#include "tbi_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int HT_EPI_WARPS = 8;                // two per TMEM lane quadrant, each takes half of the columns
constexpr int HT_THREADS = 64 + 32 * HT_EPI_WARPS;
constexpr int TW = 8, TH = 16;                 // tile = 16 rows x 8 columns of pixels = 128 GEMM rows

struct alignas(64) HaloParams {
    CUtensorMap a[2];
    CUtensorMap b;
    int n, gh, gw;
    int tiles_x, tiles_y, m_tiles, n_tiles, cgroups, nphase, total_tiles;
    int cin_g, cout_g, c0, cout_total;
    int kc, nchunks;
    int a_stages, b_stages, a_stage_bytes, b_stage_bytes, a_tx, b_tx;
    int pitch, row_bytes;
    int ngroups, ntaps;
    int g_ox[4], g_oy[4], g_ax[4], g_ay[4];
    unsigned short t_row[4][16];
    unsigned char t_grp[4][16], t_kidx[4][16];
    int a_cbase[2], a_cpix[2];
    int out_stride, ph_off_y[4], ph_off_x[4];
    int narrow;
    int f32wide;                   // fp32 output, 16-byte vector stores from the accumulator registers (bias only)
    int resident, nslabs;          // weights-resident mode: a CTA keeps one (N tile, group, phase) weight slab in smem
    int sh_x, sh_y;                // log2(tiles_x), log2(tiles_y) when both are powers of two, else -1 (divide)
    int flat;                      // resident + one halo per chunk + (ntaps, ksteps) has an unrolled issue loop
    int rank4;                     // stride-1 sources: 4-D tensor map (C, W, H, N) instead of the 5-D parity view
    int pair, m_pairs;             // PAIR mode (cta_group::2): iterations are (slab, pair of M tiles); m_pairs = ceil(m_tiles / 2)
    // iteration space of a CTA: it_first (from blockIdx), then += it_stride while < it_count.  Host-computed and read from the
    // constant bank: as kernel-computed values they were live across every role and SPILLED -- the reload sat in the loop
    // control of the epilogue, producer and MMA loops, and with ~210 KB of the SM's 228 KB in shared memory a local-memory
    // load is an L2 round trip (ncu: 14 % of the head data gradient's stall samples on the instruction behind that LDL).
    int it_stride, it_count;
    // ALL-SLABS mode (nsl > 1): a weights-resident CTA holds EVERY N tile's weights and walks pixel tiles only; each halo is
    // loaded once and multiplied against the nsl slabs in turn (accumulator buffers round-robin, one epilogue pass per slab).
    // For a one-chunk GEMM that is all epilogue (the head's data gradient: K = 64, N = 160) the operand crosses HBM once --
    // with one slab per CTA the five CTAs sharing a pixel tile drift apart and each re-reads it (ncu: 736 MB read for 402 MB).
    int nsl;
    // Side inputs of the epilogue (act' reference, residual, skip-gradient accumulator): the producer thread asks L2 for the
    // tile's box of each with ONE cp.async.bulk.prefetch.tensor when it loads the tile's first halo, i.e. a ring depth ahead of
    // the epilogue, whose one-row-per-thread loads then hit L2 instead of waiting a DRAM round trip per 32-column chunk.
    CUtensorMap side[3];
    int nside, side_split[3];      // side_split[k]: map k covers the channels >= split_c (out2 / residual2) instead of the main ones
    // The producer runs a whole halo ring (up to 8 tiles with resident weights) ahead of the epilogue; boxes requested that early
    // were evicted again before the epilogue read them (ncu, head data gradient: the act' reference crossed HBM twice).  The
    // request for tile i is therefore issued when the halo of tile i + pf_lag is loaded, pf_lag = ring depth in tiles - 1.
    int pf_lag;
    unsigned long long* trace;     // debug timeline buffer or nullptr (a kernel parameter: testing it costs no memory access)
    tbi_epilogue epi;
};

// optional timeline trace (debug): block 0 writes clock64 stamps, [role][tile][event]; enabled by tbi_debug_set_halo_trace()
// Compiled in only with -DTBI_HALO_TRACE (scratch/ timeline scripts): the stamps cost registers in every role.
#ifdef TBI_HALO_TRACE
__device__ __forceinline__ void trace(unsigned long long* tr, int role, int tile, int ev) {
    if (tr && blockIdx.x == 0 && tile < 64) tr[(role * 64 + tile) * 8 + ev] = clock64();
}
#define TBI_TRACE_PTR(p) ((p).trace)
#else
__device__ __forceinline__ void trace(unsigned long long*, int, int, int) {}
#define TBI_TRACE_PTR(p) ((unsigned long long*)nullptr)
#endif

static unsigned long long* g_halo_trace_host = nullptr;

// accumulator buffers in TMEM per CTA: narrow tiles (<= 32 columns) get four so the MMA warp can run three tiles ahead of the
// epilogue groups (their per-tile times vary; with two buffers the issuing thread waited ~400 cycles per tile for a free one)
__host__ __device__ constexpr int halo_nbuf(int bn) { return bn <= 32 ? 4 : 2; }

struct TileCoord { int x0, y0, n0, nc0, cg, ph; };

// Per-CTA constants of the roles (TMEM base address, the resident slab's (N tile, group, phase)) live in four shared-memory
// words and are re-read with ld.volatile.shared where they are used: as ordinary values they were live across the whole
// epilogue loop, ptxas spilled them, and a local-memory reload in these kernels is an L2 round trip (see it_stride above) --
// a shared-memory load is ~30 cycles and needs no register between uses.
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
#define SLOT_TMEM(ts) lds_u32(ts)
#define SLOT_NC0(ts)  ((int)lds_u32((ts) + 4))
#define SLOT_CG(ts)   ((int)lds_u32((ts) + 8))
#define SLOT_PH(ts)   ((int)lds_u32((ts) + 12))
// PAIR mode: (N tile, group, phase) of every slab, packed nc0 | cg << 16 | ph << 24, in the 32 words behind the four above --
// the five divisions of decode_tile() per iteration in every role (and the registers they held) become one shared-memory load
constexpr int PAIR_MAX_SLABS = 32;
__device__ __forceinline__ void slab_from_table(uint32_t ts, int slab, TileCoord& t) {
    const uint32_t v = lds_u32(ts + 16u + 4u * (uint32_t)slab);
    t.nc0 = (int)(v & 0xFFFFu); t.cg = (int)((v >> 16) & 0xFFu); t.ph = (int)(v >> 24);
}

__device__ __forceinline__ TileCoord decode_tile(const HaloParams& p, int tile, int BN) {
    TileCoord t;
    const int nt = tile % p.n_tiles; tile /= p.n_tiles;
    t.ph = tile % p.nphase; tile /= p.nphase;
    t.cg = tile % p.cgroups; tile /= p.cgroups;
    const int tix = tile % p.tiles_x; tile /= p.tiles_x;
    const int tiy = tile % p.tiles_y; t.n0 = tile / p.tiles_y;
    t.x0 = tix * TW; t.y0 = tiy * TH; t.nc0 = nt * BN;
    return t;
}
// resident mode: the slab part (nc0, cg, ph) is fixed per CTA, only the M-tile index moves
__device__ __forceinline__ void decode_mtile(const HaloParams& p, int mt, int& x0, int& y0, int& n0) {
    if (p.sh_x >= 0) {
        x0 = (mt & (p.tiles_x - 1)) * TW; mt >>= p.sh_x;
        y0 = (mt & (p.tiles_y - 1)) * TH; n0 = mt >> p.sh_y;
    } else {
        x0 = (mt % p.tiles_x) * TW; mt /= p.tiles_x;
        y0 = (mt % p.tiles_y) * TH; n0 = mt / p.tiles_y;
    }
}
// lane 0 does the divisions, the warp gets the result by shuffle
__device__ __forceinline__ TileCoord decode_tile_warp(const HaloParams& p, int tile, int BN, int lane, uint32_t ts, int mt) {
    if (p.resident) { TileCoord t; t.nc0 = SLOT_NC0(ts); t.cg = SLOT_CG(ts); t.ph = SLOT_PH(ts); decode_mtile(p, mt, t.x0, t.y0, t.n0); return t; }
    TileCoord t{};
    if (lane == 0) t = decode_tile(p, tile, BN);
    t.x0 = __shfl_sync(0xffffffffu, t.x0, 0); t.y0 = __shfl_sync(0xffffffffu, t.y0, 0); t.n0 = __shfl_sync(0xffffffffu, t.n0, 0);
    t.nc0 = __shfl_sync(0xffffffffu, t.nc0, 0); t.cg = __shfl_sync(0xffffffffu, t.cg, 0); t.ph = __shfl_sync(0xffffffffu, t.ph, 0);
    return t;
}

struct Rings {
    uint8_t* a_ring; uint8_t* b_ring;
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *t_full, *t_empty, *b_res;
    int slab, it_first;
};


// Weights-resident tiles whose taps all live in ONE halo load (stride-1 gathers): the issue stream of a tile is a fixed
// list of NTAPS*KSTEPS MMAs per K chunk.  Fully unrolled with the tap row offsets in registers it is ~5 instructions per
// MMA and the descriptor moves of successive MMAs are independent.  (The generic loop below spends ~30 dependent
// instructions per MMA; for N=32 tiles -- 16 cycles of math per MMA -- the single issuing thread was the bottleneck:
// ~210 cycles per MMA measured against ~49 for a free-running issue loop, profiles/r1_summary.md.)
template <int NTAPS, int KSTEPS>
__device__ __forceinline__ void resident_flat_mma_loop(const HaloParams& p, const Rings& R, uint32_t ts, uint32_t acc_cols, bool leader,
                                                       uint32_t idesc, uint32_t a_base, uint32_t a_stage_lo, uint32_t a_hi,
                                                       uint32_t b_base, uint32_t b_stage_lo, uint32_t b_hi, uint32_t row_lo, int ph, uint32_t nbuf) {
    uint32_t a_off[NTAPS];
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) a_off[t] = ((uint32_t)p.t_row[ph][t] & 0x3FFu) * row_lo;
    uint32_t sa = 0, a_par = 0, acc_it = 0;
    const uint32_t lgb = nbuf == 4 ? 2u : 1u;
    for (int i = R.it_first; i < p.it_count; i += p.it_stride, ++acc_it) {
        const uint32_t buf = acc_it & (nbuf - 1u);
        trace(TBI_TRACE_PTR(p), 1, acc_it, 0);
        tc::mbar_wait_bounded(&R.t_empty[buf], ((acc_it >> lgb) & 1u) ^ 1u);
        tc::tc_fence_after();
        trace(TBI_TRACE_PTR(p), 1, acc_it, 1);
        const uint32_t tmem_d = SLOT_TMEM(ts) + buf * acc_cols;
        uint32_t b_lo = b_base;
#pragma unroll 1
        for (int c = 0; c < p.nchunks; ++c) {
            tc::mbar_wait_bounded(&R.a_full[sa], a_par);
            tc::tc_fence_after();
            trace(TBI_TRACE_PTR(p), 1, acc_it, 2);
            const uint32_t a_lo = a_base + sa * a_stage_lo;
            if (leader) {
#pragma unroll
                for (int t = 0; t < NTAPS; ++t)
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                        tc::umma_bf16_lh(tmem_d, a_lo + a_off[t] + 2 * k, a_hi, b_lo + t * b_stage_lo + 2 * k, b_hi, idesc, (t | k) ? 1u : (c > 0 ? 1u : 0u));
                tc::umma_commit(&R.a_empty[sa]);
            }
            b_lo += NTAPS * b_stage_lo;
            if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
        }
        if (leader) tc::umma_commit(&R.t_full[buf]);
        trace(TBI_TRACE_PTR(p), 1, acc_it, 3);
    }
}

// ALL-SLABS variant of the loop above (one K chunk): per pixel tile one halo wait, then nsl accumulators in a row.
template <int NTAPS, int KSTEPS>
__device__ __forceinline__ void allslab_flat_mma_loop(const HaloParams& p, const Rings& R, uint32_t ts, uint32_t acc_cols, bool leader,
                                                      uint32_t idesc, uint32_t a_base, uint32_t a_stage_lo, uint32_t a_hi,
                                                      uint32_t b_base, uint32_t b_stage_lo, uint32_t b_hi, uint32_t row_lo, uint32_t nbuf) {
    uint32_t a_off[NTAPS];
#pragma unroll
    for (int t = 0; t < NTAPS; ++t) a_off[t] = ((uint32_t)p.t_row[0][t] & 0x3FFu) * row_lo;
    uint32_t sa = 0, a_par = 0, acc_it = 0;
    const uint32_t lgb = nbuf == 4 ? 2u : 1u;
    for (int i = R.it_first; i < p.it_count; i += p.it_stride) {
        tc::mbar_wait_bounded(&R.a_full[sa], a_par);
        const uint32_t a_lo = a_base + sa * a_stage_lo;
        uint32_t b_lo = b_base;
#pragma unroll 1
        for (int sl = 0; sl < p.nsl; ++sl, ++acc_it) {
            const uint32_t buf = acc_it & (nbuf - 1u);
            tc::mbar_wait_bounded(&R.t_empty[buf], ((acc_it >> lgb) & 1u) ^ 1u);
            tc::tc_fence_after();
            const uint32_t tmem_d = SLOT_TMEM(ts) + buf * acc_cols;
            if (leader) {
#pragma unroll
                for (int t = 0; t < NTAPS; ++t)
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                        tc::umma_bf16_lh(tmem_d, a_lo + a_off[t] + 2 * k, a_hi, b_lo + t * b_stage_lo + 2 * k, b_hi, idesc, (t | k) ? 1u : 0u);
                if (sl + 1 == p.nsl) tc::umma_commit(&R.a_empty[sa]);
                tc::umma_commit(&R.t_full[buf]);
            }
            b_lo += NTAPS * b_stage_lo;
        }
        if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
    }
}

// Streamed weights (one B stage per (chunk, tap)) with the tap count and the K steps per stage known at compile time: tap row
// offsets and group boundaries sit in registers, the K steps of a stage are back-to-back MMAs with immediate address
// increments (the generic loop below spends ~30 dependent instructions per MMA: table decode + loop-carried address
// arithmetic; this one took upsample_4 forward from 322 to 304 us, upsample_2 from 215 to 194 us).
// PAIR: the MMAs are cta_group::2 instructions (M = 256: this CTA's pixel tile and the peer's; each CTA supplies its halo and
// half of the weight stage), issued by the leader CTA only; commits are multicast so both CTAs' producers and epilogues see
// them, and the accumulator-free barrier collects the epilogue warps of both CTAs.
template <int BN, bool PAIR, int NTAPS, int KSTEPS>
__device__ __forceinline__ void streamed_mma_loop(const HaloParams& p, const Rings& R, uint32_t ts, bool leader, uint32_t idesc,
                                                  uint32_t a_base, uint32_t a_stage_lo, uint32_t a_hi, uint32_t b_base, uint32_t b_stage_lo,
                                                  uint32_t b_hi, uint32_t row_lo) {
    uint32_t toff[NTAPS];
    uint32_t first_mask = 0, last_mask = 0;
    int cur_ph = -1;
    uint32_t sa = 0, a_par = 0, sb = 0, b_par = 0, acc_it = 0;
    for (int i = R.it_first; i < p.it_count; i += p.it_stride, ++acc_it) {
        const int ph = p.nphase > 1 ? (PAIR ? (int)(lds_u32(ts + 16u + 4u * (uint32_t)(i % p.nslabs)) >> 24) : decode_tile(p, i, BN).ph) : 0;
        if (ph != cur_ph) {
            first_mask = last_mask = 0;
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
                toff[t] = ((uint32_t)p.t_row[ph][t] & 0x3FFu) * row_lo;
                const int g = p.t_grp[ph][t];
                if (t == 0 || p.t_grp[ph][t - 1] != g) first_mask |= 1u << t;
                if (t + 1 == NTAPS || p.t_grp[ph][t + 1] != g) last_mask |= 1u << t;
            }
            cur_ph = ph;
        }
        const uint32_t buf = acc_it & 1u;
        trace(TBI_TRACE_PTR(p), 1, acc_it, 0);
        if (PAIR) { if (!tc::mbar_try_wait_cluster(&R.t_empty[buf], ((acc_it >> 1) & 1u) ^ 1u)) { const long long t0 = clock64(); while (!tc::mbar_try_wait_cluster(&R.t_empty[buf], ((acc_it >> 1) & 1u) ^ 1u)) { if (clock64() - t0 > 6000000000LL) { printf("tbi tcgen05 (pair): accumulator barrier timeout (block %d)\n", blockIdx.x); __trap(); } } } }
        else tc::mbar_wait_bounded(&R.t_empty[buf], ((acc_it >> 1) & 1u) ^ 1u);
        tc::tc_fence_after();
        trace(TBI_TRACE_PTR(p), 1, acc_it, 1);
        const uint32_t tmem_d = SLOT_TMEM(ts) + buf * BN;
        uint32_t a_lo = 0;
        long long wait_a = 0, wait_b = 0, tw = 0;          // debug timeline only (p.trace): cycles this tile spent waiting for operands
#pragma unroll 1
        for (int c = 0; c < p.nchunks; ++c) {
#pragma unroll
            for (int t = 0; t < NTAPS; ++t) {
                if ((first_mask >> t) & 1u) {
                    if (TBI_TRACE_PTR(p)) tw = clock64();
                    tc::mbar_wait_bounded(&R.a_full[sa], a_par);
                    if (TBI_TRACE_PTR(p)) wait_a += clock64() - tw;
                    trace(TBI_TRACE_PTR(p), 1, acc_it, 2);
                    a_lo = a_base + sa * a_stage_lo;
                }
                if (TBI_TRACE_PTR(p)) tw = clock64();
                tc::mbar_wait_bounded(&R.b_full[sb], b_par);
                if (TBI_TRACE_PTR(p)) wait_b += clock64() - tw;
                tc::tc_fence_after();
                const uint32_t al = a_lo + toff[t], bl = b_base + sb * b_stage_lo;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                        if (PAIR) tc::umma_bf16_lh_pair(tmem_d, al + 2 * k, a_hi, bl + 2 * k, b_hi, idesc, (t | k) ? 1u : (c > 0 ? 1u : 0u));
                        else      tc::umma_bf16_lh(tmem_d, al + 2 * k, a_hi, bl + 2 * k, b_hi, idesc, (t | k) ? 1u : (c > 0 ? 1u : 0u));
                    }
                    if (PAIR) tc::umma_commit_pair(&R.b_empty[sb]); else tc::umma_commit(&R.b_empty[sb]);
                }
                if (++sb == (uint32_t)p.b_stages) { sb = 0; b_par ^= 1u; }
                if ((last_mask >> t) & 1u) {
                    if (leader) { if (PAIR) tc::umma_commit_pair(&R.a_empty[sa]); else tc::umma_commit(&R.a_empty[sa]); }
                    if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
                }
            }
        }
        if (leader) { if (PAIR) tc::umma_commit_pair(&R.t_full[buf]); else tc::umma_commit(&R.t_full[buf]); }
        trace(TBI_TRACE_PTR(p), 1, acc_it, 3);
        if (TBI_TRACE_PTR(p) && blockIdx.x == 0 && acc_it < 64) { p.trace[(64 + acc_it) * 8 + 4] = (unsigned long long)wait_a; p.trace[(64 + acc_it) * 8 + 5] = (unsigned long long)wait_b; }
    }
}

template <int BN, bool PAIR, int ACT, int DACT>
__device__ __forceinline__ void epilogue_role(const HaloParams& p, const Rings& R, const float* sbias, uint32_t ts, int warp, int lane) {
    // Two groups of four warps (one warp per TMEM lane quadrant).  Group g owns accumulator buffer g, i.e. every
    // second tile, so the epilogues of consecutive tiles overlap (the per-tile chain wait -> tcgen05.ld -> loads ->
    // math -> stores is latency-bound for small K).
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    constexpr uint32_t NBUF = halo_nbuf(BN), LGB = NBUF == 4 ? 2 : 1;
    const int q = warp & 3;
    const int grp = warp >> 2;                              // 0 or 1 == accumulator buffer
    const int m = q * 32 + lane;
    const int xx = m & (TW - 1), yy = m >> 3;
    uint32_t acc_it = 0;
    unsigned long long* tr = (warp == 0 && lane == 0) ? TBI_TRACE_PTR(p) : nullptr;
    const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
    // PAIR: the accumulator-free barrier the MMA thread waits on lives in the leader CTA
    // (one base address + 8 * buffer: a two-element array indexed by the buffer number lived in LOCAL memory, i.e. an L2 round trip
    // in front of every accumulator release)
    const uint32_t te_addr0 = PAIR ? tc::map_to_cta(tc::smem_u32(&R.t_empty[0]), 0) : 0u;
    for (int i = R.it_first; i < p.it_count; i += p.it_stride)
    for (int sl = 0; sl < p.nsl; ++sl, ++acc_it) {
        if ((int)(acc_it & 1u) != grp) continue;
        const uint32_t buf = acc_it & (NBUF - 1u);
        TileCoord t;
        bool tile_ok = true;
        if (PAIR) {
            slab_from_table(ts, i % p.nslabs, t);            // slab part: (N tile, group, phase)
            const int mt = 2 * (i / p.nslabs) + (int)rank;   // this CTA's pixel tile of the pair
            tile_ok = mt < p.m_tiles;                        // odd tile count: the last pair's second CTA repeats a tile and drops it
            decode_mtile(p, tile_ok ? mt : 0, t.x0, t.y0, t.n0);
        } else t = decode_tile_warp(p, i, BN, lane, ts, i);
        if (p.nsl > 1) t.nc0 = sl * BN;
        trace(tr, 2, acc_it, 0);
        const int gx = t.x0 + xx, gy = t.y0 + yy, n = t.n0;
        const bool valid = tile_ok && gx < p.gw && gy < p.gh;
        const int oy = gy * p.out_stride + (p.nphase > 1 ? p.ph_off_y[t.ph] : p.epi.out_off_y);
        const int ox = gx * p.out_stride + (p.nphase > 1 ? p.ph_off_x[t.ph] : p.epi.out_off_x);
        const LeanRowCtx rc = make_lean_row_ctx(p.epi, n, oy, ox, p.epi.bias ? sbias : nullptr);
        tc::mbar_wait_bounded<true>(&R.t_full[buf], (acc_it >> LGB) & 1u);
        tc::tc_fence_after();
        trace(tr, 2, acc_it, 1);
        const uint32_t taddr = SLOT_TMEM(ts) + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
        if constexpr (BN >= 32) {
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c, r);
                tc::tmem_ld_wait();
                if (c + 32 >= BN) {                                       // last read of this buffer: hand it back to the MMA warp
                    trace(tr, 2, acc_it, 2);
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) tc::mbar_arrive_cluster(te_addr0 + 8u * (buf & 1u)); else tc::mbar_arrive(&R.t_empty[buf]); }
                    trace(tr, 2, acc_it, 3);
                }
                if (valid && p.f32wide) {
                    // fp32 row of this pixel, columns [nc0 + c, +32) clipped to cout_g (a multiple of 8 here: 32-byte stores, whole sectors)
                    float* o = (float*)p.epi.out.ptr + pix_off(p.epi.out, n, oy, ox) + t.nc0 + c;
                    const float* bsm = p.epi.bias ? sbias + t.nc0 + c : nullptr;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        if (t.nc0 + c + j < p.cout_g) {
                            float v[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[j + k]) + (bsm ? bsm[j + k] : 0.f);
                            if (t.nc0 + c + j + 8 <= p.cout_g)
                                st_global_32B(o + j, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                                              make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
                            else
                                *reinterpret_cast<float4*>(o + j) = make_float4(v[0], v[1], v[2], v[3]);
                        }
                    }
                } else if (valid) epilogue_cols<ACT, DACT, 32>(rc, r, t.nc0 + c, p.cout_g, t.cg * p.cout_g);
            }
            trace(tr, 2, acc_it, 4);
        } else {
            uint32_t r[16];
            tc::tmem_ld16(taddr, r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&R.t_empty[buf]);
            if (valid && p.narrow) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (t.nc0 + j < p.cout_g) epilogue_store<__nv_bfloat16>(p.epi, n, oy, ox, t.cg * p.cout_g + t.nc0 + j, __uint_as_float(r[j]));
            } else if (valid) {
                epilogue_cols<ACT, DACT, 16>(rc, r, t.nc0, p.cout_g, t.cg * p.cout_g);
            }
        }
    }
}

template <int BN, bool PAIR>
__global__ void __launch_bounds__(HT_THREADS, 2) tapgemm_halo_kernel(const __grid_constant__ HaloParams p) {
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;            // columns per accumulator buffer
    constexpr uint32_t NBUF = halo_nbuf(BN), LGB = NBUF == 4 ? 2 : 1;
    constexpr int TMEM_COLS = NBUF * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    Rings R;
    R.a_ring = smem;
    R.b_ring = smem + (size_t)p.a_stages * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(R.b_ring + (size_t)p.b_stages * p.b_stage_bytes);
    R.a_full = bars;
    R.a_empty = R.a_full + p.a_stages;
    R.b_full = R.a_empty + p.a_stages;
    R.b_empty = R.b_full + p.b_stages;
    R.t_full = R.b_empty + p.b_stages;                     // [4] (NBUF used)
    R.t_empty = R.t_full + 4;                              // [4]
    R.b_res = R.t_empty + 4;                               // weights-resident slab loaded
    uint32_t* tslot = reinterpret_cast<uint32_t*>(R.b_res + 1);
    // the folded bias of every output channel, staged once: the epilogue reads it with shared-memory latency and the loads do not
    // sit behind the (possibly aliasing) global stores of the previous channel group
    float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tslot + 4 + (PAIR ? PAIR_MAX_SLABS : 0)) + 15) & ~static_cast<uintptr_t>(15));
    if (p.epi.bias && !p.narrow)
        for (int i = threadIdx.x; i < p.cout_total; i += HT_THREADS) sbias[i] = p.epi.bias[i];
    // tile iteration: streaming = round-robin over all tiles; resident = this CTA's slab x a strided set of M tiles
    R.slab = p.resident ? (int)(blockIdx.x % p.nslabs) : 0;
    R.it_first = p.resident ? (int)(blockIdx.x / p.nslabs) : (int)blockIdx.x;
    const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
    if (PAIR) {                    // a cluster of two CTAs walks (slab, tile pair) iterations; CTA `rank` owns tile 2*pair + rank
        R.it_first = (int)(blockIdx.x >> 1);
    }

    // the scheduler prefers higher warp ids: the two single-issuer warps get the highest ids so the epilogue math cannot starve them
    constexpr int W_TMA = HT_EPI_WARPS, W_MMA = HT_EPI_WARPS + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == W_TMA && lane == 0) {
        tc::prefetch_tmap(&p.a[0]); tc::prefetch_tmap(&p.b);
        if (p.c0 < p.cin_g * p.cgroups) tc::prefetch_tmap(&p.a[1]);
        for (int s = 0; s < p.a_stages; ++s) { tc::mbar_init(&R.a_full[s], 1); tc::mbar_init(&R.a_empty[s], 1); }
        for (int s = 0; s < p.b_stages; ++s) { tc::mbar_init(&R.b_full[s], 1); tc::mbar_init(&R.b_empty[s], 1); }
        for (int s = 0; s < (int)NBUF; ++s) { tc::mbar_init(&R.t_full[s], 1); tc::mbar_init(&R.t_empty[s], PAIR ? 8 : 4); }
        tc::mbar_init(R.b_res, 1);
        tc::fence_barrier_init();
    }
    if (warp == W_MMA) { if (PAIR) tc::tmem_alloc_pair<TMEM_COLS>(tslot); else tc::tmem_alloc<TMEM_COLS>(tslot); }
    if (threadIdx.x == 0) {
        TileCoord st{};
        if (p.resident) st = decode_tile(p, R.slab, BN);
        tslot[1] = (uint32_t)st.nc0; tslot[2] = (uint32_t)st.cg; tslot[3] = (uint32_t)st.ph;
    }
    if (PAIR && (int)threadIdx.x < p.nslabs) {
        const TileCoord st = decode_tile(p, (int)threadIdx.x, BN);
        tslot[4 + threadIdx.x] = (uint32_t)st.nc0 | ((uint32_t)st.cg << 16) | ((uint32_t)st.ph << 24);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (PAIR) tc::cluster_sync_all();                       // the peer's barriers exist before anything is signalled across the pair
    tc::tc_fence_after();
    const uint32_t ts = tc::smem_u32(tslot);                // [0] TMEM base (written by tcgen05.alloc), [1..3] the resident slab's nc0 / cg / ph

    if (warp == W_TMA) {
        // ===================== TMA producer: the whole warp walks the loop, one elected lane issues =====================
        const bool leader = tc::elect_one();
        uint32_t a_it = 0, sa = 0, a_par = 1, sb = 0, b_par = 1;       // ring stage + parity (empty barriers start "free")
        TileCoord slab_t{};
        if (p.resident) { slab_t.nc0 = SLOT_NC0(ts); slab_t.cg = SLOT_CG(ts); slab_t.ph = SLOT_PH(ts); }
        if (p.resident && R.it_first < p.it_count && leader) {            // the whole weight slab (all-slabs mode: every slab), once
            tc::mbar_expect_tx(R.b_res, (uint32_t)(p.nsl * p.nchunks * p.ntaps) * p.b_tx);
            for (int sl = 0; sl < p.nsl; ++sl)
                for (int c = 0; c < p.nchunks; ++c)
                    for (int tap = 0; tap < p.ntaps; ++tap)
                        tc::tma_load_2d(R.b_ring + (size_t)((sl * p.nchunks + c) * p.ntaps + tap) * p.b_stage_bytes, &p.b, R.b_res,
                                        (int)p.t_kidx[slab_t.ph][tap] * p.cin_g + c * p.kc,
                                        slab_t.ph * p.cout_total + slab_t.cg * p.cout_g + slab_t.nc0 + sl * BN);
        }
        auto side_prefetch = [&](int it) {
            TileCoord t;
            if (PAIR) {
                slab_from_table(ts, it % p.nslabs, t);
                int mt = 2 * (it / p.nslabs) + (int)rank;
                if (mt >= p.m_tiles) return;
                decode_mtile(p, mt, t.x0, t.y0, t.n0);
            } else if (p.resident) { t = slab_t; decode_mtile(p, it, t.x0, t.y0, t.n0); } else t = decode_tile(p, it, BN);
            if (p.nsl > 1) {
                // all-slabs: one box per side tensor spans the columns of every slab
                for (int k = 0; k < p.nside; ++k) tc::tma_prefetch_4d(&p.side[k], 0, t.x0, t.y0, t.n0);
            } else {
                const int co0 = t.cg * p.cout_g + t.nc0;
                const int in_split = (p.epi.split_c > 0 && co0 >= p.epi.split_c) ? 1 : 0;
                for (int k = 0; k < p.nside; ++k)
                    if (p.side_split[k] == in_split) tc::tma_prefetch_4d(&p.side[k], in_split ? co0 - p.epi.split_c : co0, t.x0, t.y0, t.n0);
            }
        };
        for (int i = R.it_first; i < p.it_count; i += p.it_stride) {
            TileCoord t;
            if (PAIR) {
                slab_from_table(ts, i % p.nslabs, t);
                int mt = 2 * (i / p.nslabs) + (int)rank;
                if (mt >= p.m_tiles) mt = p.m_tiles - 1;                 // odd tile count: load something valid, the epilogue drops it
                decode_mtile(p, mt, t.x0, t.y0, t.n0);
            } else if (p.resident) { t = slab_t; decode_mtile(p, i, t.x0, t.y0, t.n0); } else t = decode_tile(p, i, BN);
            if (p.nside && leader) {
                // the tile whose side boxes are requested now: pf_lag iterations behind the halo loads (see HaloParams::pf_lag)
                const int ip = i - p.pf_lag * p.it_stride;
                if (ip >= R.it_first) side_prefetch(ip);
            }
            for (int c = 0; c < p.nchunks; ++c) {
                const int ch = c * p.kc;
                int src = 0, cch = ch + (p.cgroups > 1 ? t.cg * p.cin_g : 0);
                if (p.cgroups == 1 && cch >= p.c0) { src = 1; cch -= p.c0; }
                int tap = 0;
                for (int grp = 0; grp < p.ngroups; ++grp) {
                    trace(TBI_TRACE_PTR(p), 0, a_it, 0);
                    tc::mbar_wait_bounded(&R.a_empty[sa], a_par);
                    trace(TBI_TRACE_PTR(p), 0, a_it, 1);
                    if (leader && PAIR) {
                        // both CTAs' halos complete on the LEADER's barrier (it expects the pair's bytes)
                        if (rank == 0) tc::mbar_expect_tx(&R.a_full[sa], 2u * (uint32_t)p.a_tx);
                        const uint32_t bar = tc::map_to_cta(tc::smem_u32(&R.a_full[sa]), 0);
                        if (p.rank4)
                            tc::tma_load_4d_pair(R.a_ring + (size_t)sa * p.a_stage_bytes, &p.a[src], bar, cch, t.x0 + p.g_ox[grp], t.y0 + p.g_oy[grp], t.n0);
                        else
                            tc::tma_load_5d_pair(R.a_ring + (size_t)sa * p.a_stage_bytes, &p.a[src], bar,
                                                 p.a_cbase[src] + cch + p.g_ax[grp] * p.a_cpix[src], t.x0 + p.g_ox[grp], p.g_ay[grp], t.y0 + p.g_oy[grp], t.n0);
                    } else if (leader) {
                        tc::mbar_expect_tx(&R.a_full[sa], p.a_tx);
                        if (p.rank4)
                            tc::tma_load_4d(R.a_ring + (size_t)sa * p.a_stage_bytes, &p.a[src], &R.a_full[sa], cch, t.x0 + p.g_ox[grp], t.y0 + p.g_oy[grp], t.n0);
                        else
                            tc::tma_load_5d(R.a_ring + (size_t)sa * p.a_stage_bytes, &p.a[src], &R.a_full[sa],
                                            p.a_cbase[src] + cch + p.g_ax[grp] * p.a_cpix[src], t.x0 + p.g_ox[grp], p.g_ay[grp], t.y0 + p.g_oy[grp], t.n0);
                    }
                    ++a_it;
                    if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
                    if (p.resident) continue;
                    while (tap < p.ntaps && p.t_grp[t.ph][tap] == grp) {
                        tc::mbar_wait_bounded(&R.b_empty[sb], b_par);
                        if (leader && PAIR) {                    // this CTA's half of the weight stage: rows [rank*BN/2, +BN/2) of the N tile
                            if (rank == 0) tc::mbar_expect_tx(&R.b_full[sb], 2u * (uint32_t)p.b_tx);
                            tc::tma_load_2d_pair(R.b_ring + (size_t)sb * p.b_stage_bytes, &p.b, tc::map_to_cta(tc::smem_u32(&R.b_full[sb]), 0),
                                                 (int)p.t_kidx[t.ph][tap] * p.cin_g + ch, t.ph * p.cout_total + t.cg * p.cout_g + t.nc0 + (int)rank * (BN / 2));
                        } else if (leader) {
                            tc::mbar_expect_tx(&R.b_full[sb], p.b_tx);
                            tc::tma_load_2d(R.b_ring + (size_t)sb * p.b_stage_bytes, &p.b, &R.b_full[sb], (int)p.t_kidx[t.ph][tap] * p.cin_g + ch,
                                            t.ph * p.cout_total + t.cg * p.cout_g + t.nc0);
                        }
                        if (++sb == (uint32_t)p.b_stages) { sb = 0; b_par ^= 1u; }
                        ++tap;
                    }
                }
            }
        }
        if (p.nside && leader && p.pf_lag > 0 && R.it_first < p.it_count) {       // the last pf_lag tiles' boxes
            const int cnt = (p.it_count - 1 - R.it_first) / p.it_stride + 1;    // iterations of this CTA
            for (int k = cnt > p.pf_lag ? cnt - p.pf_lag : 0; k < cnt; ++k) side_prefetch(R.it_first + k * p.it_stride);
        }
        __syncwarp();
    } else if (warp == W_MMA && (!PAIR || rank == 0)) {
        // ===================== MMA issuer: warp-uniform loop (descriptors live in uniform registers), elected lane issues =====================
        // (PAIR: the leader CTA issues for both; the peer's MMA warp only owns its half of the TMEM allocation)
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::make_idesc_bf16(PAIR ? 256 : 128, BN, 0, 0);
        const uint32_t layout = p.kc == 64 ? 2u : p.kc == 32 ? 4u : 6u;
        const uint64_t da_base = tc::smem_desc_base(16, (uint32_t)p.pitch * p.row_bytes, layout);       // 8-row group == one halo tile row
        const uint64_t db_base = tc::smem_desc_base(16, 8u * p.row_bytes, layout);
        const uint32_t a_lo0 = (uint32_t)da_base, a_hi = (uint32_t)(da_base >> 32), b_lo0 = (uint32_t)db_base, b_hi = (uint32_t)(db_base >> 32);
        const uint32_t a_ring_lo = (tc::smem_u32(R.a_ring) & 0x3FFFFu) >> 4, a_stage_lo = (uint32_t)p.a_stage_bytes >> 4;
        const uint32_t b_ring_lo = (tc::smem_u32(R.b_ring) & 0x3FFFFu) >> 4, b_stage_lo = (uint32_t)p.b_stage_bytes >> 4;
        const uint32_t row_lo = (uint32_t)p.row_bytes >> 4;
        const int ksteps = p.kc / 16;
        uint32_t a_it = 0, b_it = 0, acc_it = 0;
        const int slab_ph = p.resident ? SLOT_PH(ts) : 0;
        // per-phase tap table packed into registers: 12 bits per tap = row offset (10) | first-of-group (1) | last-of-group (1),
        // five taps per 64-bit word, consumed by a running shift -> no memory load sits in front of an MMA
        unsigned long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0; int cur_ph = -1;
        auto load_taps = [&](int ph) {
            tw0 = tw1 = tw2 = tw3 = 0;
            for (int t = 0; t < p.ntaps; ++t) {
                const int g = p.t_grp[ph][t];
                unsigned long long e = (unsigned long long)p.t_row[ph][t] & 0x3FFull;
                if (t == 0 || p.t_grp[ph][t - 1] != g) e |= 0x400ull;
                if (t + 1 == p.ntaps || p.t_grp[ph][t + 1] != g) e |= 0x800ull;
                e <<= 12 * (t % 5);
                if (t < 5) tw0 |= e; else if (t < 10) tw1 |= e; else if (t < 15) tw2 |= e; else tw3 |= e;
            }
            cur_ph = ph;
        };
        bool flat_done = false;
        if (p.resident && p.ngroups == 1 && p.flat && p.nsl > 1) {
            // ---- every slab resident, one K chunk, one tap: halo once, nsl accumulators ----
            if (R.it_first < p.it_count) tc::mbar_wait_bounded(R.b_res, 0);
            const uint32_t a_base = a_lo0 + a_ring_lo, b_base = b_lo0 + b_ring_lo;
#define TBI_ALLSLAB(KS) allslab_flat_mma_loop<1, KS>(p, R, ts, ACC_COLS, leader, idesc, a_base, a_stage_lo, a_hi, b_base, b_stage_lo, b_hi, row_lo, NBUF)
            switch (ksteps) {
                case 1:  TBI_ALLSLAB(1); break;
                case 2:  TBI_ALLSLAB(2); break;
                default: TBI_ALLSLAB(4); break;
            }
#undef TBI_ALLSLAB
            flat_done = true;
        } else if (p.resident && p.ngroups == 1 && p.flat) {
            // ---- weights resident, one halo per chunk: unrolled issue stream ----
            if (R.it_first < p.it_count) tc::mbar_wait_bounded(R.b_res, 0);
            const uint32_t a_base = a_lo0 + a_ring_lo, b_base = b_lo0 + b_ring_lo;
#define TBI_FLAT(NT, KS) resident_flat_mma_loop<NT, KS>(p, R, ts, ACC_COLS, leader, idesc, a_base, a_stage_lo, a_hi, b_base, b_stage_lo, b_hi, row_lo, slab_ph, NBUF)
            switch (p.ntaps * 8 + ksteps) {
                case 1 * 8 + 1: TBI_FLAT(1, 1); break;
                case 1 * 8 + 2: TBI_FLAT(1, 2); break;
                case 1 * 8 + 4: TBI_FLAT(1, 4); break;
                case 4 * 8 + 2: TBI_FLAT(4, 2); break;
                case 4 * 8 + 4: TBI_FLAT(4, 4); break;
                case 9 * 8 + 1: TBI_FLAT(9, 1); break;
                case 9 * 8 + 2: TBI_FLAT(9, 2); break;
                default:        TBI_FLAT(9, 4); break;
            }
#undef TBI_FLAT
            flat_done = true;
        }
        if (!flat_done && !p.resident && BN >= 64 && ksteps == 4) {
            // ---- weights streamed, 64-channel stages: unrolled issue stream per tap count (NBUF == 2 for BN >= 64) ----
            const uint32_t a_base = a_lo0 + a_ring_lo, b_base = b_lo0 + b_ring_lo;
#define TBI_STREAM(NT) streamed_mma_loop<BN, PAIR, NT, 4>(p, R, ts, leader, idesc, a_base, a_stage_lo, a_hi, b_base, b_stage_lo, b_hi, row_lo)
            if constexpr (BN >= 64) {
                switch (p.ntaps) {
                    case 1:  TBI_STREAM(1); flat_done = true; break;
                    case 4:  TBI_STREAM(4); flat_done = true; break;
                    case 9:  TBI_STREAM(9); flat_done = true; break;
                    case 16: TBI_STREAM(16); flat_done = true; break;
                    default: break;
                }
            }
#undef TBI_STREAM
        }
        if (flat_done) {
        } else if (p.resident) {
            // ---- weights resident: nothing but halo waits, descriptor adds and MMAs in the steady state ----
            if (R.it_first < p.it_count) { load_taps(slab_ph); tc::mbar_wait_bounded(R.b_res, 0); }
            uint32_t sa = 0, a_par = 0;
            const uint32_t a_base = a_lo0 + a_ring_lo, b_base = b_lo0 + b_ring_lo;
            for (int i = R.it_first; i < p.it_count; i += p.it_stride, ++acc_it) {
                const uint32_t buf = acc_it & (NBUF - 1u);
                trace(TBI_TRACE_PTR(p), 1, acc_it, 0);
                tc::mbar_wait_bounded(&R.t_empty[buf], ((acc_it >> LGB) & 1u) ^ 1u);
                tc::tc_fence_after();
                trace(TBI_TRACE_PTR(p), 1, acc_it, 1);
                const uint32_t tmem_d = SLOT_TMEM(ts) + buf * ACC_COLS;
                uint32_t accum = 0, b_lo = b_base, a_lo = 0;
#pragma unroll 1
                for (int c = 0; c < p.nchunks; ++c) {
                    unsigned long long cur = tw0;
#pragma unroll 1
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        if (tap == 5) cur = tw1; else if (tap == 10) cur = tw2; else if (tap == 15) cur = tw3;
                        const uint32_t e = (uint32_t)cur & 0xFFFu;
                        cur >>= 12;
                        if (e & 0x400u) {
                            tc::mbar_wait_bounded(&R.a_full[sa], a_par);
                            tc::tc_fence_after();
                            trace(TBI_TRACE_PTR(p), 1, acc_it, 2);
                            a_lo = a_base + sa * a_stage_lo;
                        }
                        const uint32_t al = a_lo + (e & 0x3FFu) * row_lo;
                        if (leader) {
#pragma unroll 1
                            for (int k = 0; k < ksteps; ++k) { tc::umma_bf16_lh(tmem_d, al + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, accum); accum = 1; }
                        }
                        b_lo += b_stage_lo;
                        if (e & 0x800u) {
                            if (leader) tc::umma_commit(&R.a_empty[sa]);
                            if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
                        }
                    }
                }
                if (leader) tc::umma_commit(&R.t_full[buf]);
                trace(TBI_TRACE_PTR(p), 1, acc_it, 3);
            }
        } else {
            // ---- weights streamed: one B stage per (chunk, tap); wrap-around stage counters, nothing but waits + MMAs ----
            uint32_t sa = 0, a_par = 0, sb = 0, b_par = 0;
            const uint32_t a_base = a_lo0 + a_ring_lo, b_base = b_lo0 + b_ring_lo;
            for (int i = R.it_first; i < p.it_count; i += p.it_stride, ++acc_it) {
                const int ph = (p.nphase > 1 ? decode_tile(p, i, BN).ph : 0);
                if (ph != cur_ph) load_taps(ph);
                const uint32_t buf = acc_it & (NBUF - 1u);
                trace(TBI_TRACE_PTR(p), 1, acc_it, 0);
                tc::mbar_wait_bounded(&R.t_empty[buf], ((acc_it >> LGB) & 1u) ^ 1u);
                tc::tc_fence_after();
                trace(TBI_TRACE_PTR(p), 1, acc_it, 1);
                const uint32_t tmem_d = SLOT_TMEM(ts) + buf * ACC_COLS;
                uint32_t accum = 0, a_lo = 0;
#pragma unroll 1
                for (int c = 0; c < p.nchunks; ++c) {
                    unsigned long long cur = tw0;
#pragma unroll 1
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        if (tap == 5) cur = tw1; else if (tap == 10) cur = tw2; else if (tap == 15) cur = tw3;
                        const uint32_t e = (uint32_t)cur & 0xFFFu;
                        cur >>= 12;
                        if (e & 0x400u) {
                            tc::mbar_wait_bounded(&R.a_full[sa], a_par);
                            trace(TBI_TRACE_PTR(p), 1, acc_it, 2);
                            a_lo = a_base + sa * a_stage_lo;
                        }
                        tc::mbar_wait_bounded(&R.b_full[sb], b_par);
                        tc::tc_fence_after();
                        const uint32_t al = a_lo + (e & 0x3FFu) * row_lo, bl = b_base + sb * b_stage_lo;
                        if (leader) {
#pragma unroll 1
                            for (int k = 0; k < ksteps; ++k) { tc::umma_bf16_lh(tmem_d, al + 2 * k, a_hi, bl + 2 * k, b_hi, idesc, accum); accum = 1; }
                            tc::umma_commit(&R.b_empty[sb]);
                        }
                        accum = 1;
                        if (++sb == (uint32_t)p.b_stages) { sb = 0; b_par ^= 1u; }
                        if (e & 0x800u) {
                            if (leader) tc::umma_commit(&R.a_empty[sa]);
                            if (++sa == (uint32_t)p.a_stages) { sa = 0; a_par ^= 1u; }
                        }
                    }
                }
                if (leader) tc::umma_commit(&R.t_full[buf]);
                trace(TBI_TRACE_PTR(p), 1, acc_it, 3);
            }
        }
        __syncwarp();
    } else if (warp < HT_EPI_WARPS) {
        // ===================== epilogue (warp w owns TMEM lanes [32*(w%4), +32)) =====================
        TBI_EPI_DISPATCH(p.epi.act, p.epi.dact, (epilogue_role<BN, PAIR, A_, D_>(p, R, sbias, ts, warp, lane)));
    }
    tc::tc_fence_before();
    __syncthreads();
    if (PAIR) tc::cluster_sync_all();                       // the peer may still signal this CTA's barriers / read its operands
    if (warp == W_MMA) { if (PAIR) tc::tmem_dealloc_pair<TMEM_COLS>(SLOT_TMEM(ts)); else tc::tmem_dealloc<TMEM_COLS>(SLOT_TMEM(ts)); }
}

inline uint32_t r1024(uint32_t x) { return (x + 1023u) & ~1023u; }

// 5-D activation map with a halo box [kc x (TW+ex) x 1 x (TH+ey) x 1]
int halo_act_tmap(CUtensorMap* out, const tbi_view& v, int n, int stride, int kc, int bw, int bh, int* cbase, int* cpix, bool rank4) {
    if (rank4 && stride == 1) {
        const uint64_t px4 = (uint64_t)v.cstride * 2;
        uint64_t d4[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)n};
        uint64_t s4[3] = {px4, px4 * v.w, px4 * v.w * v.h};
        uint32_t b4[4] = {(uint32_t)kc, (uint32_t)bw, (uint32_t)bh, 1u};
        *cbase = 0; *cpix = 0;
        return tbi_make_tmap_bf16(out, (char*)v.ptr + (size_t)v.coff * 2, 4, d4, s4, b4, kc * 2);
    }
    uint64_t dims[5], strides[4];
    uint32_t box[5] = {(uint32_t)kc, (uint32_t)bw, 1u, (uint32_t)bh, 1u};
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
    return tbi_make_tmap_bf16(out, base, 5, dims, strides, box, kc * 2);
}

template <int BN, bool PAIR>
int launch_halo(const HaloParams& p, int grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tapgemm_halo_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attr_err != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    if (PAIR) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(HT_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, tapgemm_halo_kernel<BN, PAIR>, p);
        if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "tapgemm_halo (pair): %s", cudaGetErrorString(e));
        return TBI_OK;
    }
    tapgemm_halo_kernel<BN, PAIR><<<grid, HT_THREADS, smem, s>>>(p);
    TBI_CUDA_LAUNCH_CHECK("tapgemm_halo");
    return TBI_OK;
}

int halo_pick_kc(const tbi_tapgemm* d) {
    const int c0 = d->groups > 1 ? d->cin_g : d->src[0].c;
    const int c1 = (d->groups == 1 && d->src[1].ptr) ? d->src[1].c : 0;
    for (int kc = 64; kc >= 16; kc >>= 1)
        if (c0 % kc == 0 && c1 % kc == 0) return kc;
    return 0;
}

struct TapPlan { int ngroups, ex, ey; int ox[4], oy[4], ax[4], ay[4]; int grp[4][16], row_dy[4][16], row_dx[4][16], kidx[4][16]; };

// group taps by input parity (stride 2) or into one group (stride 1); returns false if the halo would be too large
bool plan_taps(const tbi_tapgemm* d, TapPlan* tp) {
    const int nph = d->nphase > 1 ? d->nphase : 1;
    memset(tp, 0, sizeof(*tp));
    int lo_x[4], hi_x[4], lo_y[4], hi_y[4]; bool used[4] = {false, false, false, false};
    int qy[4][16], qx[4][16], g[4][16];
    for (int ph = 0; ph < nph; ++ph)
        for (int t = 0; t < d->ntaps; ++t) {
            const int dy = d->nphase > 1 ? d->ph_dy[ph][t] : d->dy[t], dx = d->nphase > 1 ? d->ph_dx[ph][t] : d->dx[t];
            int grp = 0, y = dy, x = dx, ay = 0, ax = 0;
            if (d->in_stride == 2) { ay = dy & 1; ax = dx & 1; y = (dy - ay) / 2; x = (dx - ax) / 2; grp = ay * 2 + ax; }
            qy[ph][t] = y; qx[ph][t] = x; g[ph][t] = grp;
            if (!used[grp]) { used[grp] = true; lo_x[grp] = hi_x[grp] = x; lo_y[grp] = hi_y[grp] = y; tp->ax[grp] = ax; tp->ay[grp] = ay; }
            else { if (x < lo_x[grp]) lo_x[grp] = x; if (x > hi_x[grp]) hi_x[grp] = x; if (y < lo_y[grp]) lo_y[grp] = y; if (y > hi_y[grp]) hi_y[grp] = y; }
        }
    // compact group ids in increasing order
    int remap[4], ng = 0;
    for (int i = 0; i < 4; ++i) { remap[i] = -1; if (used[i]) { remap[i] = ng; tp->ox[ng] = lo_x[i]; tp->oy[ng] = lo_y[i]; tp->ax[ng] = tp->ax[i]; tp->ay[ng] = tp->ay[i];
                                                               if (hi_x[i] - lo_x[i] > tp->ex) tp->ex = hi_x[i] - lo_x[i]; if (hi_y[i] - lo_y[i] > tp->ey) tp->ey = hi_y[i] - lo_y[i]; ++ng; } }
    tp->ngroups = ng;
    if (tp->ex > 2 || tp->ey > 2) return false;
    if (nph > 1 && ng != 1) return false;
    // per phase: taps sorted by group (stable), remember original tap index for the weight column block
    for (int ph = 0; ph < nph; ++ph) {
        int k = 0;
        for (int gi = 0; gi < ng; ++gi)
            for (int t = 0; t < d->ntaps; ++t)
                if (remap[g[ph][t]] == gi) { tp->grp[ph][k] = gi; tp->row_dy[ph][k] = qy[ph][t] - tp->oy[gi]; tp->row_dx[ph][k] = qx[ph][t] - tp->ox[gi]; tp->kidx[ph][k] = t; ++k; }
    }
    return true;
}

}  // namespace

// debug/test knob (not part of the public header): minimum number of (slab, tile-pair) iterations for PAIR mode; tests lower
// it to drive small shapes (odd tile counts, partial tiles) through the cta_group::2 path.  <= 0 restores the default.
static int g_pair_min = 0;
extern "C" int tbi_debug_set_pair_min(int v) { g_pair_min = v; return 0; }

// debug: device buffer of 3*64*8 uint64 (or nullptr to disable); not part of the public header
extern "C" int tbi_debug_set_halo_trace(void* buf) {
    g_halo_trace_host = (unsigned long long*)buf;
    return 0;
}

bool tbi_tapgemm_halo_supported(const tbi_tapgemm* d) {
    static const bool disabled = getenv("TBI_TC_NO_HALO") != nullptr;
    if (disabled) return false;
    if (d->gh < 16 || d->gw < 8) return false;
    TapPlan tp;
    return plan_taps(d, &tp);
}

// precondition: tbi_tapgemm_tc_supported(d) (alignment, dtype, channel multiples) and tbi_tapgemm_halo_supported(d)
int tbi_tapgemm_halo(const tbi_tapgemm* d, cudaStream_t s) {
    HaloParams p; memset(&p, 0, sizeof(p));
    TapPlan tp;
    if (!plan_taps(d, &tp)) return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapgemm_halo: tap pattern");
    const int kc = halo_pick_kc(d);
    p.n = d->n; p.gh = d->gh; p.gw = d->gw; p.trace = g_halo_trace_host;
    p.tiles_x = (d->gw + TW - 1) / TW; p.tiles_y = (d->gh + TH - 1) / TH; p.m_tiles = d->n * p.tiles_x * p.tiles_y;
    p.cgroups = d->groups; p.nphase = d->nphase > 1 ? d->nphase : 1;
    p.cin_g = d->cin_g; p.cout_g = d->cout_g; p.cout_total = d->cout_g * d->groups;
    p.c0 = d->groups > 1 ? d->cin_g * d->groups : d->src[0].c;
    p.kc = kc; p.nchunks = d->cin_g / kc; p.row_bytes = kc * 2;
    p.ngroups = tp.ngroups; p.ntaps = d->ntaps; p.pitch = TW + tp.ex;
    p.narrow = tbi_tc_narrow(d) ? 1 : 0;
    p.f32wide = tbi_tc_f32wide(d) ? 1 : 0;
    p.epi = d->epi;
    p.out_stride = d->nphase > 1 ? 2 : (d->epi.out_stride ? d->epi.out_stride : 1);
    for (int gi = 0; gi < tp.ngroups; ++gi) { p.g_ox[gi] = tp.ox[gi]; p.g_oy[gi] = tp.oy[gi]; p.g_ax[gi] = tp.ax[gi]; p.g_ay[gi] = tp.ay[gi]; }
    for (int ph = 0; ph < p.nphase; ++ph) {
        p.ph_off_y[ph] = d->ph_off_y[ph]; p.ph_off_x[ph] = d->ph_off_x[ph];
        for (int t = 0; t < d->ntaps; ++t) {
            p.t_row[ph][t] = (unsigned short)(tp.row_dy[ph][t] * p.pitch + tp.row_dx[ph][t]);
            p.t_grp[ph][t] = (unsigned char)tp.grp[ph][t]; p.t_kidx[ph][t] = (unsigned char)tp.kidx[ph][t];
        }
    }
    const int bw = TW + tp.ex, bh = TH + tp.ey;
    static const bool no_rank4 = getenv("TBI_TC_NO_RANK4") != nullptr;
    p.rank4 = (d->in_stride == 1 && !no_rank4) ? 1 : 0;
    int rc = halo_act_tmap(&p.a[0], d->src[0], d->n, d->in_stride, kc, bw, bh, &p.a_cbase[0], &p.a_cpix[0], p.rank4); if (rc) return rc;
    if (d->groups == 1 && d->src[1].ptr) { rc = halo_act_tmap(&p.a[1], d->src[1], d->n, d->in_stride, kc, bw, bh, &p.a_cbase[1], &p.a_cpix[1], p.rank4); if (rc) return rc; }
    else p.a[1] = p.a[0];
    int bn = 128;
    while (bn > 16 && bn / 2 >= d->cout_g) bn >>= 1;
    if (p.narrow) bn = 16;                                   // the element-wise epilogue exists for 16-column tiles only
    {
        // A weights-resident CTA is tied to ONE N tile.  With a ragged last tile (the head's data gradient: 160 = 128 + 32
        // columns over one 64-channel K chunk, all epilogue) the CTAs of the full tile carry 4x the work of the others, and block
        // ids are dealt to SMs round-robin, so with two slabs the heavy CTAs all land on the even SMs.  Equal 32-column slabs
        // instead (5 x 59 CTAs; the 16 KB halo is then read five times, from L2): 387 -> 310 us, step 7.68 -> 7.62 ms.
        static const bool ragged32 = getenv("TBI_HALO_NO_RAGGED_BN32") == nullptr;
        if (ragged32 && !p.narrow && bn >= 64 && d->cout_g % bn != 0 && d->cout_g % 32 == 0 && d->groups == 1 && p.nphase == 1 && kc == 64 && d->cin_g == 64) bn = 32;
    }
    p.n_tiles = (d->cout_g + bn - 1) / bn;
    p.total_tiles = p.m_tiles * p.cgroups * p.nphase * p.n_tiles;
    p.nslabs = p.n_tiles * p.cgroups * p.nphase;
    p.m_pairs = (p.m_tiles + 1) / 2;
    // PAIR (cta_group::2): streamed 128-column tiles of 64-channel stages with enough tile pairs for every SM pair; the
    // resident / small-K layers keep the single-CTA schedules (their weights are not re-streamed)
    static const bool no_pair = getenv("TBI_TC_NO_PAIR") != nullptr;
    const long long slab_bytes_full = (long long)(d->cin_g / kc) * d->ntaps * (long long)r1024((uint32_t)(bn * kc * 2));
    static const bool no_resident0 = getenv("TBI_TC_NO_RESIDENT") != nullptr;
    const bool would_be_resident = !no_resident0 && slab_bytes_full <= 72 * 1024 && p.nslabs <= 2 * tbi_sm_count() &&
                                   slab_bytes_full + 2 * (long long)r1024((uint32_t)(bw * bh * kc * 2)) <= 104 * 1024 && (d->cin_g / kc) * d->ntaps <= 192;
    p.pair = (!no_pair && !would_be_resident && bn == 128 && !p.narrow && kc == 64 && d->cout_g % 128 == 0 &&
              (d->ntaps == 1 || d->ntaps == 4 || d->ntaps == 9 || d->ntaps == 16) && p.nslabs <= PAIR_MAX_SLABS && d->cout_g < 65536 && d->groups < 256 &&
              (long long)p.m_pairs * p.nslabs >= (g_pair_min > 0 ? (long long)g_pair_min : (long long)tbi_sm_count())) ? 1 : 0;
    const int b_rows = p.pair ? bn / 2 : bn;                 // PAIR: each CTA streams half of the N tile's weight rows
    {
        const uint64_t K = (uint64_t)d->ntaps * d->cin_g;
        uint64_t dims[2] = {K, (uint64_t)p.cout_total * p.nphase};
        uint64_t strides[1] = {K * 2};
        uint32_t box[2] = {(uint32_t)kc, (uint32_t)b_rows};
        rc = tbi_make_tmap_bf16(&p.b, const_cast<void*>(d->w), 2, dims, strides, box, kc * 2);
        if (rc) return rc;
    }
    p.a_tx = bw * bh * kc * 2; p.b_tx = b_rows * kc * 2;
    p.a_stage_bytes = (int)r1024((uint32_t)p.a_tx); p.b_stage_bytes = (int)r1024((uint32_t)p.b_tx);
    // ~104 KB per CTA so that two CTAs share an SM (TBI_HALO_BUDGET_KB > 113 -> one CTA per SM with deeper rings)
    static const int budget_kb = getenv("TBI_HALO_BUDGET_KB") ? atoi(getenv("TBI_HALO_BUDGET_KB")) : 104;
    const int budget = budget_kb * 1024;
    const int max_ctas = (budget_kb > 113 ? 1 : 2) * tbi_sm_count();
    const long long slab_bytes = (long long)p.nchunks * d->ntaps * p.b_stage_bytes;
    static const bool no_resident = getenv("TBI_TC_NO_RESIDENT") != nullptr;
    p.resident = (!p.pair && !no_resident && slab_bytes <= 72 * 1024 && p.nslabs <= max_ctas && slab_bytes + 2 * p.a_stage_bytes <= budget &&
                  p.nchunks * d->ntaps <= 192) ? 1 : 0;
    {
        static const bool no_flat = getenv("TBI_TC_NO_FLAT") != nullptr;
        const int ks = kc / 16, nt = d->ntaps;
        const bool have = (nt == 1 && (ks == 1 || ks == 2 || ks == 4)) || (nt == 4 && (ks == 2 || ks == 4)) || (nt == 9 && (ks == 1 || ks == 2 || ks == 4));
        p.flat = (p.resident && p.ngroups == 1 && have && !no_flat) ? 1 : 0;
    }
    p.nsl = 1;
    {
        static const bool no_allslabs = getenv("TBI_HALO_NO_ALLSLABS") != nullptr;
        if (!no_allslabs && p.resident && p.flat && p.n_tiles > 1 && p.n_tiles <= 8 && p.cgroups == 1 && p.nphase == 1 && p.nchunks == 1 &&
            d->ntaps == 1 && bn <= 32 && (long long)p.n_tiles * slab_bytes + 2 * p.a_stage_bytes <= budget) {
            p.nsl = p.n_tiles;
            p.nslabs = 1;                                    // the CTA's slab arithmetic sees one slab starting at column 0
        }
    }
    {
        static const bool no_pf = getenv("TBI_HALO_NO_SIDE_PF") != nullptr;
        const tbi_epilogue& e = d->epi;
        p.nside = 0;
        if (!no_pf && !p.narrow && !p.f32wide && p.nphase == 1 && p.out_stride == 1 && e.out_off_x == 0 && e.out_off_y == 0) {
            auto add = [&](const tbi_view& v, int split) -> int {
                if (!v.ptr || ((v.cstride * 2) & 15) || ((v.coff * 2) & 15) || (((uintptr_t)v.ptr) & 15)) return TBI_OK;
                int cols = v.c < bn * p.nsl ? v.c : bn * p.nsl;
                cols &= ~7;
                if (cols <= 0) return TBI_OK;
                const uint64_t px = (uint64_t)v.cstride * 2;
                uint64_t d4[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)d->n};
                uint64_t s4[3] = {px, px * v.w, px * v.w * v.h};
                uint32_t b4[4] = {(uint32_t)cols, (uint32_t)TW, (uint32_t)TH, 1u};
                p.side_split[p.nside] = split;
                const int r = tbi_make_tmap_bf16(&p.side[p.nside], (char*)v.ptr + (size_t)v.coff * 2, 4, d4, s4, b4, 0);
                if (r == TBI_OK) ++p.nside;
                return r;
            };
            if (e.dact != TBI_ACT_NONE) { rc = add(e.dact_ref, 0); if (rc) return rc; }
            rc = add(e.residual, 0); if (rc) return rc;
            if (e.split_c > 0) { rc = add(e.residual2, 1); if (rc) return rc; }
        }
    }
    auto lg2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    p.sh_x = lg2(p.tiles_x); p.sh_y = lg2(p.tiles_y);
    if (p.sh_x < 0 || p.sh_y < 0) p.sh_x = p.sh_y = -1;
    int grid;
    if (p.resident) {
        // weights stay in smem for the CTA's lifetime; everything else is a deep ring of halo tiles
        p.b_stages = p.nsl * p.nchunks * d->ntaps;
        int as = (int)((budget - p.nsl * slab_bytes) / p.a_stage_bytes);
        if (as > 8) as = 8;
        p.a_stages = as;
        int per_slab = max_ctas / p.nslabs;
        if (per_slab > p.m_tiles) per_slab = p.m_tiles;
        if (per_slab < 1) per_slab = 1;
        grid = per_slab * p.nslabs;
    } else {
        // a halo stage feeds ntaps/ngroups weight tiles: two halo stages are enough, the weight ring gets the rest
        p.a_stages = budget_kb > 113 ? 4 : 2;
        int bs = (budget - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
        if (bs > (budget_kb > 113 ? 16 : 8)) bs = budget_kb > 113 ? 16 : 8;
        if (bs < 2) bs = 2;
        p.b_stages = bs;
        grid = max_ctas;
        if (grid > p.total_tiles) grid = p.total_tiles;
        if (p.pair) {                                        // clusters of two: one pair per (slab, tile pair) iteration slot
            const long long its = (long long)p.m_pairs * p.nslabs;
            long long pairs = max_ctas / 2;
            if (pairs > its) pairs = its;
            grid = (int)(2 * pairs);
        }
    }
    {
        static const int lag_env = getenv("TBI_HALO_PF_LAG") ? atoi(getenv("TBI_HALO_PF_LAG")) : -1;
        const int tiles_ahead = p.a_stages / (p.nchunks * p.ngroups);       // halo stages per tile = chunks x tap groups
        p.pf_lag = lag_env >= 0 ? lag_env : (tiles_ahead > 1 ? tiles_ahead - 1 : 0);      // one tile of lead (swept 0..4: 6.99 / 6.98 / 6.95 / 6.95 / 6.91 ms per step)
        if (p.pf_lag > tiles_ahead - 1) p.pf_lag = tiles_ahead > 1 ? tiles_ahead - 1 : 0;
    }
    p.it_stride = p.pair ? grid / 2 : p.resident ? grid / p.nslabs : grid;
    p.it_count = p.pair ? p.m_pairs * p.nslabs : p.resident ? p.m_tiles : p.total_tiles;
    const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_stage_bytes + 1024 + 512 + (p.pair ? 4 * PAIR_MAX_SLABS : 0) +
                        ((size_t)p.cout_total * 4 + 64) + (size_t)(2 * (p.a_stages + p.b_stages) + 8) * 8;
    if (p.pair) return launch_halo<128, true>(p, grid, smem, s);
    switch (bn) {
        case 128: return launch_halo<128, false>(p, grid, smem, s);
        case 64:  return launch_halo<64, false>(p, grid, smem, s);
        case 32:  return launch_halo<32, false>(p, grid, smem, s);
        default:  return launch_halo<16, false>(p, grid, smem, s);
    }
}
