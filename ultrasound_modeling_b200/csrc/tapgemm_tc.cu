// tcgen05 implicit-GEMM tap-GEMM (bf16 in, fp32 accumulate in TMEM) for sm_100a.
//
//   acc[pixel][co] = sum_{tap, ci} in[pixel + off(tap)][ci] * w[co][tap*cin_g + ci]
//
// One CTA computes a 128-pixel x BN-channel output tile.  The 128 pixels are a (tn x th x tw) box
// of the (N, H, W) pixel grid, so the A operand of every (tap, 64-channel chunk) K-step is ONE
// 5-D TMA box of the NHWC activation tensor shifted by the tap offset; out-of-image rows/columns
// are zero-filled by TMA, which is exactly SAME padding.  The virtual channel concat of two
// sources is two tensor maps walked one after the other along K.  A stride-2 gather (dgrad of a
// transposed conv) views the tensor as [N, H/2, 2, W/2, 2*C] so the parity becomes a box coordinate.
// Both operands are K-major with the 128/64/32-byte TMA swizzle matching the UMMA descriptor.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> fused bias/dropout/activation/residual/act' -> 16-byte stores).
// smem ring of `stages` {A,B} buffers with full/empty mbarriers; tcgen05.commit releases a stage.
// <= ~100 KB smem and <= 128 TMEM columns per CTA so two CTAs share an SM: one CTA's epilogue
// overlaps the other's main loop.
#include "tbi_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"
#include <mutex>

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 192;

struct alignas(64) TcGemmParams {
    CUtensorMap a[2];
    CUtensorMap b;
    int n, gh, gw;
    int ltw, lth;                 // log2(tile w), log2(tile h); tile n = 128 >> (ltw + lth)
    int tiles_x, tiles_y;
    int cin_g, cout_g, groups, c0, cout_total;
    int kc, stages;
    int nphase, ntaps;
    int out_stride;
    int narrow;                   // cout_g < 8 (e.g. the 3-class head): scalar epilogue, fp32 output allowed
    int a_cbase[2], a_cpix[2];
    int ph_off_y[4], ph_off_x[4];
    signed char qy[4][16], qx[4][16], ay[4][16], ax[4][16];
    tbi_epilogue epi;
};

__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
    // a wrong descriptor must become an error, not a hung GPU: trap after ~2 s of waiting
    tc::mbar_wait_bounded(bar, parity);
}

__host__ __device__ constexpr uint32_t round1024(uint32_t x) { return (x + 1023u) & ~1023u; }

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS) tapgemm_tc_kernel(const __grid_constant__ TcGemmParams p) {
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int kc = p.kc, stages = p.stages;
    const uint32_t a_tx = TILE_M * kc * 2, b_tx = BN * kc * 2;
    const uint32_t a_bytes = round1024(a_tx), b_bytes = round1024(b_tx);
    uint8_t* bar_base = smem + (size_t)stages * (a_bytes + b_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty = full + stages;
    uint64_t* tfull = empty + stages;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&p.a[0]); tc::prefetch_tmap(&p.b);
        if (p.c0 < p.cin_g * p.groups) tc::prefetch_tmap(&p.a[1]);
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tfull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tslot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tslot;

    // ---- which tile
    int t = blockIdx.x;
    const int tix = t % p.tiles_x; t /= p.tiles_x;
    const int tiy = t % p.tiles_y; const int tib = t / p.tiles_y;
    const int tw = 1 << p.ltw, th = 1 << p.lth;
    const int x0 = tix * tw, y0 = tiy * th, n0 = tib * (TILE_M >> (p.ltw + p.lth));
    const int nc0 = blockIdx.y * BN;                       // first output channel of the tile within the group
    const int g = blockIdx.z % p.groups, ph = blockIdx.z / p.groups;
    const int cpt = p.cin_g / kc;
    const int iters = p.ntaps * cpt;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int it = 0; it < iters; ++it) {
                const int s = it % stages;
                mbar_wait_guard(&empty[s], (((uint32_t)(it / stages)) & 1u) ^ 1u);
                const int tap = it / cpt, ch = (it - tap * cpt) * kc;
                int src = 0, cch = ch + (p.groups > 1 ? g * p.cin_g : 0);
                if (p.groups == 1 && cch >= p.c0) { src = 1; cch -= p.c0; }
                uint8_t* a_s = smem + (size_t)s * (a_bytes + b_bytes);
                uint8_t* b_s = a_s + a_bytes;
                tc::mbar_expect_tx(&full[s], a_tx + b_tx);
                tc::tma_load_5d(a_s, &p.a[src], &full[s], p.a_cbase[src] + cch + (int)p.ax[ph][tap] * p.a_cpix[src],
                                x0 + (int)p.qx[ph][tap], (int)p.ay[ph][tap], y0 + (int)p.qy[ph][tap], n0);
                tc::tma_load_2d(b_s, &p.b, &full[s], tap * p.cin_g + ch, ph * p.cout_total + g * p.cout_g + nc0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer (one thread) =====
            const uint32_t idesc = tc::make_idesc_bf16(TILE_M, BN, 0, 0);
            const uint32_t layout = kc == 64 ? 2u : kc == 32 ? 4u : 6u;      // 128B / 64B / 32B swizzle
            const uint32_t sbo = 8u * kc * 2u;                               // 8 rows of kc bf16
            for (int it = 0; it < iters; ++it) {
                const int s = it % stages;
                mbar_wait_guard(&full[s], ((uint32_t)(it / stages)) & 1u);
                tc::tc_fence_after();
                const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * (a_bytes + b_bytes));
                const uint32_t b_addr = a_addr + a_bytes;
                for (int k = 0; k < kc / 16; ++k) {
                    const uint64_t da = tc::make_smem_desc(a_addr + k * 32, 16, sbo, layout);
                    const uint64_t db = tc::make_smem_desc(b_addr + k * 32, 16, sbo, layout);
                    tc::umma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
                }
                tc::umma_commit(&empty[s]);                 // frees the smem stage when these MMAs retire
            }
            tc::umma_commit(tfull);                         // accumulator complete
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), +32) =====
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int xx = m & (tw - 1), yy = (m >> p.ltw) & (th - 1), nn = m >> (p.ltw + p.lth);
        const int gx = x0 + xx, gy = y0 + yy, n = n0 + nn;
        const bool valid = gx < p.gw && gy < p.gh && n < p.n;
        const int os = p.out_stride;
        const int oy = gy * os + (p.nphase > 1 ? p.ph_off_y[ph] : p.epi.out_off_y);
        const int ox = gx * os + (p.nphase > 1 ? p.ph_off_x[ph] : p.epi.out_off_x);
        RowCtx rc{};
        if (valid && !p.narrow) rc = make_row_ctx(p.epi, n, oy, ox);
        mbar_wait_guard(tfull, 0);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        if constexpr (BN >= 32) {
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + c, r);
                tc::tmem_ld_wait();
                if (valid) TBI_EPI_DISPATCH(p.epi.act, p.epi.dact, (epilogue_cols<A_, D_, 32>(rc, r, nc0 + c, p.cout_g, g * p.cout_g)));
            }
        } else {
            uint32_t r[16];
            tc::tmem_ld16(taddr, r);
            tc::tmem_ld_wait();
            if (valid && p.narrow) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (nc0 + j < p.cout_g) epilogue_store<__nv_bfloat16>(p.epi, n, oy, ox, g * p.cout_g + nc0 + j, __uint_as_float(r[j]));
            } else if (valid) {
                TBI_EPI_DISPATCH(p.epi.act, p.epi.dact, (epilogue_cols<A_, D_, 16>(rc, r, nc0, p.cout_g, g * p.cout_g)));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

bool aligned_view(const tbi_view& v) {
    return v.ptr == nullptr || (v.cstride % 8 == 0 && v.coff % 8 == 0 && v.c % 8 == 0 && ((uintptr_t)v.ptr & 15) == 0);
}
int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

}  // namespace

tbi_encode_tiled_fn tbi_get_encode_tiled() {
    static tbi_encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tbi_encode_tiled_fn>(p);
        else cudaGetLastError();
    });
    return fn;
}

int tbi_make_tmap_bf16(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int swizzle_bytes) {
    tbi_encode_tiled_fn enc = tbi_get_encode_tiled();
    TBI_CHECK(enc != nullptr, TBI_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TBI_CHECK(r == CUDA_SUCCESS, TBI_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return TBI_OK;
}

// 5-D activation map: (channel-ish, x, parity, y, n).  stride 1: parity dim has size 1.
// stride 2: the tensor [N, 2H, 2W, C] is viewed as [N, H, 2, W, 2*cstride]: the x parity selects a
// +cstride offset inside the innermost dimension, the y parity is coordinate 2.
static int make_act_tmap(CUtensorMap* out, const tbi_view& v, int n, int stride, int kc, int tw, int th, int tn, int* cbase, int* cpix) {
    uint64_t dims[5], strides[4];
    uint32_t box[5] = {(uint32_t)kc, (uint32_t)tw, 1u, (uint32_t)th, (uint32_t)tn};
    const uint64_t px = (uint64_t)v.cstride * 2;            // bytes per pixel record
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

static int pick_kc(const tbi_tapgemm* d) {
    const int c0 = d->groups > 1 ? d->cin_g : d->src[0].c;
    const int c1 = (d->groups == 1 && d->src[1].ptr) ? d->src[1].c : 0;
    for (int kc = 64; kc >= 16; kc >>= 1)
        if (c0 % kc == 0 && c1 % kc == 0) return kc;
    return 0;
}

bool tbi_tapgemm_tc_supported(const tbi_tapgemm* d, const char** why) {
#define NO(msg) do { *why = msg; return false; } while (0)
    if (d->dtype != TBI_BF16) NO("storage dtype is not bf16");
    if (!tbi_get_encode_tiled()) NO("no cuTensorMapEncodeTiled");
    if (d->ntaps < 1 || d->ntaps > TBI_MAX_TAPS) NO("ntaps");
    if (!(d->nphase <= 1 || (d->nphase == 4 && d->ntaps <= 4))) NO("nphase");
    if (d->in_stride != 1 && d->in_stride != 2) NO("in_stride");
    if (d->groups > 1 && d->src[1].ptr) NO("groups with two sources");
    if (pick_kc(d) == 0) NO("input channels per source are not a multiple of 16");
    const bool narrow = tbi_tc_narrow(d);
    if (d->cout_g % 8 != 0 && !narrow) NO("output channels per group are not a multiple of 8");
    if (!aligned_view(d->src[0]) || !aligned_view(d->src[1])) NO("source view not 16-byte aligned");
    if (d->in_stride == 2 && ((d->src[0].h | d->src[0].w) & 1)) NO("stride-2 gather needs even source dims");
    const tbi_epilogue& e = d->epi;
    const bool f32wide = tbi_tc_f32wide(d);
    if (e.out_f32 && !narrow && !f32wide) NO("fp32 output");
    if (e.act != TBI_ACT_NONE && e.dact != TBI_ACT_NONE) NO("activation and activation-derivative in one epilogue");
    if (!narrow && !f32wide) {
        if (!aligned_view(e.out) || !aligned_view(e.residual) || !aligned_view(e.dact_ref) || !aligned_view(e.out2) || !aligned_view(e.residual2))
            NO("epilogue view not 16-byte aligned");
        if (e.split_c % 8 != 0) NO("split_c");
        auto same_grid = [&](const tbi_view& v) { return !v.ptr || (v.h == e.out.h && v.w == e.out.w); };
        if (!same_grid(e.residual) || !same_grid(e.out2) || !same_grid(e.residual2) || (e.dact != TBI_ACT_NONE && !same_grid(e.dact_ref)))
            NO("epilogue tensors on a different pixel grid than the output");
    }
    if (e.bias && ((uintptr_t)e.bias & 15) && !narrow) NO("bias alignment");
    if (((uintptr_t)d->w & 15)) NO("weight alignment");
    if ((e.drop_keep && (((uintptr_t)e.drop_keep & 7) || e.out.c != e.out.cstride)) ||
        (e.dact_keep && (((uintptr_t)e.dact_keep & 7) || e.dact_ref.c != e.dact_ref.cstride))) NO("dropout mask layout");
    return true;
#undef NO
}

template <int BN>
static int launch_tc(const TcGemmParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tapgemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attr_err != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    tapgemm_tc_kernel<BN><<<grid, NUM_THREADS, smem, s>>>(p);
    TBI_CUDA_LAUNCH_CHECK("tapgemm_tc");
    return TBI_OK;
}

int tbi_tapgemm_tc(const tbi_tapgemm* d, cudaStream_t s) {
    const char* why = "";
    if (!tbi_tapgemm_tc_supported(d, &why)) return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapgemm_tc: %s", why);
    if (tbi_tapgemm_xpack_supported(d)) return tbi_tapgemm_xpack(d, s);     // 3x3, few channels: dx taps packed into N
    if (tbi_tapgemm_halo_supported(d)) return tbi_tapgemm_halo(d, s);       // large spatial extents: persistent halo schedule
    TcGemmParams p; memset(&p, 0, sizeof(p));
    const int kc = pick_kc(d);
    int ltw = ilog2_ceil(d->gw); if (ltw > 4) ltw = 4;
    int lth = ilog2_ceil(d->gh); if (lth > 7 - ltw) lth = 7 - ltw;
    const int tw = 1 << ltw, th = 1 << lth, tn = TILE_M >> (ltw + lth);
    p.n = d->n; p.gh = d->gh; p.gw = d->gw; p.ltw = ltw; p.lth = lth;
    p.tiles_x = (d->gw + tw - 1) / tw; p.tiles_y = (d->gh + th - 1) / th;
    const int tiles_b = (d->n + tn - 1) / tn;
    p.cin_g = d->cin_g; p.cout_g = d->cout_g; p.groups = d->groups; p.cout_total = d->cout_g * d->groups;
    p.c0 = d->groups > 1 ? d->cin_g * d->groups : d->src[0].c;
    p.kc = kc; p.nphase = d->nphase > 1 ? d->nphase : 1; p.ntaps = d->ntaps;
    p.narrow = tbi_tc_narrow(d) ? 1 : 0;
    p.epi = d->epi;
    p.out_stride = d->nphase > 1 ? 2 : (d->epi.out_stride ? d->epi.out_stride : 1);
    for (int ph = 0; ph < p.nphase; ++ph) {
        p.ph_off_y[ph] = d->ph_off_y[ph]; p.ph_off_x[ph] = d->ph_off_x[ph];
        for (int t = 0; t < d->ntaps; ++t) {
            const int dy = d->nphase > 1 ? d->ph_dy[ph][t] : d->dy[t], dx = d->nphase > 1 ? d->ph_dx[ph][t] : d->dx[t];
            if (d->in_stride == 1) { p.qy[ph][t] = (signed char)dy; p.qx[ph][t] = (signed char)dx; p.ay[ph][t] = 0; p.ax[ph][t] = 0; }
            else {                 // dy = 2*q + a with a in {0,1} (floor division)
                const int ay = dy & 1, ax = dx & 1;
                p.ay[ph][t] = (signed char)ay; p.ax[ph][t] = (signed char)ax;
                p.qy[ph][t] = (signed char)((dy - ay) / 2); p.qx[ph][t] = (signed char)((dx - ax) / 2);
            }
        }
    }
    int rc = make_act_tmap(&p.a[0], d->src[0], d->n, d->in_stride, kc, tw, th, tn, &p.a_cbase[0], &p.a_cpix[0]);
    if (rc) return rc;
    if (d->groups == 1 && d->src[1].ptr) {
        rc = make_act_tmap(&p.a[1], d->src[1], d->n, d->in_stride, kc, tw, th, tn, &p.a_cbase[1], &p.a_cpix[1]);
        if (rc) return rc;
    } else p.a[1] = p.a[0];
    int bn = 128;
    while (bn > 16 && bn / 2 >= d->cout_g) bn >>= 1;
    if (p.narrow) bn = 16;                                   // the element-wise epilogue exists for 16-column tiles only
    {
        const uint64_t K = (uint64_t)d->ntaps * d->cin_g;
        uint64_t dims[2] = {K, (uint64_t)p.cout_total * p.nphase};
        uint64_t strides[1] = {K * 2};
        uint32_t box[2] = {(uint32_t)kc, (uint32_t)bn};
        rc = tbi_make_tmap_bf16(&p.b, const_cast<void*>(d->w), 2, dims, strides, box, kc * 2);
        if (rc) return rc;
    }
    const int iters = d->ntaps * (d->cin_g / kc);
    const uint32_t stage_bytes = round1024(TILE_M * kc * 2) + round1024(bn * kc * 2);
    int stages = (int)(96 * 1024 / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages > iters) stages = iters;
    if (stages < 1) stages = 1;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    dim3 grid((unsigned)(p.tiles_x * p.tiles_y * tiles_b), (unsigned)((d->cout_g + bn - 1) / bn), (unsigned)(d->groups * p.nphase));
    switch (bn) {
        case 128: return launch_tc<128>(p, grid, smem, s);
        case 64:  return launch_tc<64>(p, grid, smem, s);
        case 32:  return launch_tc<32>(p, grid, smem, s);
        default:  return launch_tc<16>(p, grid, smem, s);
    }
}

