// extern "C" surface of libtbi_sm100.so (declared in include/tbi_sm100.h): error plumbing, the
// convolution entry points (which expand into tap-GEMM descriptors) and implementation dispatch.
#include "tbi_common.cuh"
#include <string.h>
#include <stdlib.h>
#include <mutex>

static thread_local char g_err[512] = "";

int tbi_set_error(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}

int tbi_sm_count() {
    static int n = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0; cudaDeviceProp p{};
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) n = p.multiProcessorCount;
        if (n <= 0) n = 148;
    });
    return n;
}

extern "C" int tbi_version(void) { return TBI_VERSION; }
extern "C" const char* tbi_last_error(void) { return g_err; }
extern "C" int tbi_device_ok(void) {
    int dev = 0; cudaDeviceProp p{};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return p.major == 10 ? 1 : 0;
}

// Which bf16 launches under TBI_IMPL_AUTO left the tensor cores for the CUDA-core tap-GEMM, and why (the last reason): a user
// can see that a layer shape fell off the tcgen05 path instead of silently running ~50x slower (tbi_fallback_stats).
#include <atomic>
static std::atomic<long long> g_fallback_gemm{0}, g_fallback_wgrad{0};
static char g_fallback_why[256] = "";
static std::mutex g_fallback_mu;
static void note_fallback(std::atomic<long long>& ctr, const char* kind, const char* why, int cin_g, int cout_g, int groups) {
    ctr.fetch_add(1, std::memory_order_relaxed);
    std::lock_guard<std::mutex> lk(g_fallback_mu);
    snprintf(g_fallback_why, sizeof(g_fallback_why), "%s (groups %d, cin/group %d, cout/group %d): %s", kind, groups, cin_g, cout_g, why);
}
extern "C" int tbi_fallback_stats(int64_t* tapgemm_simt, int64_t* tapwgrad_simt, int reset) {
    if (tapgemm_simt) *tapgemm_simt = g_fallback_gemm.load();
    if (tapwgrad_simt) *tapwgrad_simt = g_fallback_wgrad.load();
    if (reset) { g_fallback_gemm = 0; g_fallback_wgrad = 0; std::lock_guard<std::mutex> lk(g_fallback_mu); g_fallback_why[0] = 0; }
    return TBI_OK;
}
extern "C" const char* tbi_last_fallback(void) { return g_fallback_why; }

extern "C" int tbi_tapgemm_run(const tbi_tapgemm* d, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (d->impl == TBI_IMPL_SIMT) return tbi_tapgemm_simt(d, s);
    const char* why = "";
    const bool ok = tbi_tapgemm_tc_supported(d, &why);
    if (d->impl == TBI_IMPL_TCGEN05) {
        if (!ok) return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapgemm: tcgen05 path does not take this shape: %s", why);
        return tbi_tapgemm_tc(d, s);
    }
    if (ok) return tbi_tapgemm_tc(d, s);
    if (tbi_tapgemm_direct_supported(d)) return tbi_tapgemm_direct(d, s);
    if (d->dtype == TBI_BF16) note_fallback(g_fallback_gemm, "tap-GEMM", why, d->cin_g, d->cout_g, d->groups);
    return tbi_tapgemm_simt(d, s);
}

extern "C" int tbi_tapwgrad_run(const tbi_tapwgrad* d, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (d->impl == TBI_IMPL_SIMT) return tbi_tapwgrad_simt(d, s);
    const char* why = "";
    const bool ok = tbi_tapwgrad_tc_supported(d, &why);
    if (d->impl == TBI_IMPL_TCGEN05) {
        if (!ok) return tbi_set_error(TBI_ERR_UNSUPPORTED, "tapwgrad: tcgen05 path does not take this shape: %s", why);
        return tbi_tapwgrad_tc(d, s);
    }
    if (ok) return tbi_tapwgrad_tc(d, s);
    if (tbi_tapwgrad_direct_supported(d)) return tbi_tapwgrad_direct(d, s);
    if (d->dtype == TBI_BF16) note_fallback(g_fallback_wgrad, "weight-gradient tap-GEMM", why, d->cin_g, d->cout_g, d->groups);
    return tbi_tapwgrad_simt(d, s);
}

extern "C" int64_t tbi_workspace_bytes(const tbi_tapwgrad* d) {
    const char* why = "";
    if (d->impl == TBI_IMPL_SIMT || !tbi_tapwgrad_tc_supported(d, &why)) return 0;
    return tbi_tapwgrad_tc_workspace(d);
}

// ---------------------------------------------------------------------------------------------
// Conv2D (stride 1, SAME)
// ---------------------------------------------------------------------------------------------
static int fill_conv_taps(int ksize, int dilation, int* dy, int* dx) {
    int n = 0;
    for (int r = 0; r < ksize; ++r)
        for (int c = 0; c < ksize; ++c) { dy[n] = (r - ksize / 2) * dilation; dx[n] = (c - ksize / 2) * dilation; ++n; }
    return n;
}

// see include/tbi_sm100.h.  Expanded when the grouped form cannot use the tensor cores in some direction (K per group not a
// multiple of 16) while the dense form can, and the dense layer stays small.  The dense form reads pad16(cin) input channels:
// a total input width that is not a multiple of 16 (radix 3: 12 groups x 5 channels = 60) needs pixel records padded with
// zeros up to the next multiple of 16 (the engine stores such tensors that way); the pad rows of the pack are zero.
static inline int pad16(int c) { return (c + 15) & ~15; }
extern "C" int tbi_conv_dense_expand(int dtype, int groups, int cin_g, int cout_g) {
    static const bool off = getenv("TBI_NO_DENSE_EXPAND") != nullptr;
    if (off || dtype != TBI_BF16 || groups <= 1) return 0;
    if (cin_g % 16 == 0 && cout_g % 16 == 0) return 0;                       // grouped form is tensor-core eligible both ways
    const int cin = groups * cin_g, cout = groups * cout_g;
    return (cout % 16 == 0 && pad16(cin) <= 256 && cout <= 1024) ? 1 : 0;
}
extern "C" int64_t tbi_conv_packed_elems(int dtype, int ksize, int groups, int cin_g, int cout_total) {
    const int64_t grouped = (int64_t)ksize * ksize * cin_g * cout_total;
    return tbi_conv_dense_expand(dtype, groups, cin_g, cout_total / (groups > 0 ? groups : 1)) ? (int64_t)ksize * ksize * pad16(groups * cin_g) * cout_total : grouped;
}
extern "C" int64_t tbi_conv2d_wgrad_workspace(int dtype, int ksize, int groups, int cin_total, int cout_total) {
    if (groups <= 1 || !tbi_conv_dense_expand(dtype, groups, cin_total / groups, cout_total / groups)) return 0;
    return (int64_t)ksize * ksize * pad16(cin_total) * cout_total * (int64_t)sizeof(float);
}
// the expanded form addresses pad16(c) channels of a view that names c of them: the record must have the room
static int widen_view(tbi_view* v, int cpad, const char* what) {
    TBI_CHECK(v->cstride - v->coff >= cpad, TBI_ERR_BAD_SHAPE, "%s: block-diagonal expansion addresses %d channels, the pixel record has room for %d "
              "(store the tensor with zero-padded records)", what, cpad, v->cstride - v->coff);
    v->c = cpad;
    return TBI_OK;
}

static int check_conv_args(int ksize, int dilation) {
    TBI_CHECK(ksize == 1 || ksize == 3, TBI_ERR_UNSUPPORTED, "conv2d: ksize %d (1 or 3)", ksize);
    TBI_CHECK(dilation == 1 || dilation == 2 || dilation == 4 || dilation == 8, TBI_ERR_UNSUPPORTED, "conv2d: dilation %d", dilation);
    return TBI_OK;
}

extern "C" int tbi_conv2d_fwd(int dtype, int impl, int n, int h, int w, int ksize, int dilation, int groups,
                              const tbi_view* src0, const tbi_view* src1, int cout_total, const void* w_packed,
                              const tbi_epilogue* epi, void* stream) {
    int rc = check_conv_args(ksize, dilation); if (rc) return rc;
    TBI_CHECK(groups >= 1 && cout_total % groups == 0, TBI_ERR_BAD_SHAPE, "conv2d_fwd: cout %d %% groups %d", cout_total, groups);
    tbi_tapgemm d; memset(&d, 0, sizeof(d));
    d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = groups;
    d.src[0] = *src0; if (src1 && src1->ptr) d.src[1] = *src1;
    const int cin = src0->c + ((src1 && src1->ptr) ? src1->c : 0);
    TBI_CHECK(cin % groups == 0, TBI_ERR_BAD_SHAPE, "conv2d_fwd: cin %d %% groups %d", cin, groups);
    d.cin_g = cin / groups; d.cout_g = cout_total / groups;
    if (tbi_conv_dense_expand(dtype, groups, d.cin_g, d.cout_g)) {           // block-diagonal pack over pad16(cin) input channels
        d.groups = 1; d.cin_g = pad16(cin); d.cout_g = cout_total;
        rc = widen_view(&d.src[0], d.cin_g, "conv2d_fwd input"); if (rc) return rc;
    }
    d.in_stride = 1; d.ntaps = fill_conv_taps(ksize, dilation, d.dy, d.dx);
    d.w = w_packed; d.epi = *epi;
    if (d.epi.out_stride == 0) d.epi.out_stride = 1;
    return tbi_tapgemm_run(&d, stream);
}

extern "C" int tbi_conv2d_dgrad(int dtype, int impl, int n, int h, int w, int ksize, int dilation, int groups,
                                const tbi_view* dz, int cin_total, const void* w_packed_dgrad, const tbi_epilogue* epi,
                                void* stream) {
    int rc = check_conv_args(ksize, dilation); if (rc) return rc;
    TBI_CHECK(groups >= 1 && dz->c % groups == 0 && cin_total % groups == 0, TBI_ERR_BAD_SHAPE, "conv2d_dgrad: groups");
    tbi_tapgemm d; memset(&d, 0, sizeof(d));
    d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = groups;
    d.src[0] = *dz;
    d.cin_g = dz->c / groups;            // K side = forward output channels
    d.cout_g = cin_total / groups;       // produced = forward input channels
    d.in_stride = 1; d.ntaps = fill_conv_taps(ksize, dilation, d.dy, d.dx);
    d.w = w_packed_dgrad; d.epi = *epi;
    if (tbi_conv_dense_expand(dtype, groups, d.cout_g, d.cin_g)) {           // writes pad16(cin_total) channels; the pad lanes come out 0
        d.groups = 1; d.cin_g = dz->c; d.cout_g = pad16(cin_total);
        rc = widen_view(&d.epi.out, d.cout_g, "conv2d_dgrad output"); if (rc) return rc;
        if (d.epi.dact != TBI_ACT_NONE) { rc = widen_view(&d.epi.dact_ref, d.cout_g, "conv2d_dgrad act' reference"); if (rc) return rc; }
        if (d.epi.residual.ptr) { rc = widen_view(&d.epi.residual, d.cout_g, "conv2d_dgrad residual"); if (rc) return rc; }
    }
    if (d.epi.out_stride == 0) d.epi.out_stride = 1;
    return tbi_tapgemm_run(&d, stream);
}

extern "C" int tbi_conv2d_wgrad(int dtype, int impl, int n, int h, int w, int ksize, int dilation, int groups,
                                const tbi_view* x0, const tbi_view* x1, const tbi_view* dz, float* dw_hwio, float* dbias,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_conv_args(ksize, dilation); if (rc) return rc;
    tbi_tapwgrad d; memset(&d, 0, sizeof(d));
    d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = groups;
    d.a_src[0] = *x0; if (x1 && x1->ptr) d.a_src[1] = *x1;
    d.b_src = *dz;
    const int cin = x0->c + ((x1 && x1->ptr) ? x1->c : 0);
    TBI_CHECK(groups >= 1 && cin % groups == 0 && dz->c % groups == 0, TBI_ERR_BAD_SHAPE, "conv2d_wgrad: groups");
    d.cin_g = cin / groups; d.cout_g = dz->c / groups;
    d.a_stride = 1; d.b_stride = 1;
    d.ntaps = fill_conv_taps(ksize, dilation, d.a_dy, d.a_dx);
    d.dw = dw_hwio;
    d.tap_stride = (int64_t)d.cin_g * dz->c; d.ci_stride = dz->c; d.co_stride = 1;
    d.dbias = dbias; d.workspace = workspace; d.workspace_bytes = workspace_bytes;
    const int64_t need = tbi_conv2d_wgrad_workspace(dtype, ksize, groups, cin, dz->c);
    if (need > 0 && workspace && workspace_bytes >= need && impl != TBI_IMPL_SIMT) {
        // dense weight gradient into the scratch, then its block-diagonal part is added to the grouped HWIO gradient
        cudaStream_t s = (cudaStream_t)stream;
        if (cudaMemsetAsync(workspace, 0, (size_t)need, s) != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "conv2d_wgrad: memset");
        tbi_tapwgrad e = d;
        const int cinp = pad16(cin);
        e.groups = 1; e.cin_g = cinp; e.cout_g = dz->c;
        rc = widen_view(&e.a_src[0], cinp, "conv2d_wgrad input"); if (rc) return rc;
        e.dw = (float*)workspace; e.tap_stride = (int64_t)cinp * dz->c; e.ci_stride = dz->c; e.co_stride = 1;
        e.workspace = nullptr; e.workspace_bytes = 0;
        rc = tbi_tapwgrad_run(&e, stream); if (rc) return rc;
        return tbi_wgrad_gather_blocks(d.ntaps, groups, d.cin_g, cinp, dz->c, (const float*)workspace, dw_hwio, s);
    }
    return tbi_tapwgrad_run(&d, stream);
}

// ---------------------------------------------------------------------------------------------
// Conv2DTranspose (stride 2, TF 'same'):  oy = 2*iy - pad + ky,  pad = 1 (k=4) or 0 (k=3, cropped)
// ---------------------------------------------------------------------------------------------
extern "C" int tbi_conv2d_transpose_s2_fwd(int dtype, int impl, int n, int h, int w, int ksize, const tbi_view* src0,
                                           const tbi_view* src1, int cout, const void* w_packed, const tbi_epilogue* epi,
                                           void* stream) {
    TBI_CHECK(ksize == 3 || ksize == 4, TBI_ERR_UNSUPPORTED, "convT: ksize %d", ksize);
    const int cin = src0->c + ((src1 && src1->ptr) ? src1->c : 0);
    if (ksize == 4) {                    // every phase has 2x2 taps: one descriptor, one launch on the tcgen05 path
        tbi_tapgemm d; memset(&d, 0, sizeof(d));
        d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = 1;
        d.src[0] = *src0; if (src1 && src1->ptr) d.src[1] = *src1;
        d.cin_g = cin; d.cout_g = cout; d.in_stride = 1; d.w = w_packed; d.epi = *epi; d.epi.out_stride = 2;
        d.nphase = 4;
        for (int ph = 0; ph < 4; ++ph) {
            int ky[TBI_MAX_TAPS], kx[TBI_MAX_TAPS], dy[TBI_MAX_TAPS], dx[TBI_MAX_TAPS];
            d.ntaps = tbi_convt_phase_taps(ksize, ph >> 1, ph & 1, ky, kx, dy, dx);
            for (int t = 0; t < d.ntaps; ++t) { d.ph_dy[ph][t] = dy[t]; d.ph_dx[ph][t] = dx[t]; }
            d.ph_off_y[ph] = ph >> 1; d.ph_off_x[ph] = ph & 1;
        }
        return tbi_tapgemm_run(&d, stream);
    }
    size_t woff = 0;
    for (int ph = 0; ph < 4; ++ph) {
        tbi_tapgemm d; memset(&d, 0, sizeof(d));
        int ky[TBI_MAX_TAPS], kx[TBI_MAX_TAPS];
        d.ntaps = tbi_convt_phase_taps(ksize, ph >> 1, ph & 1, ky, kx, d.dy, d.dx);
        d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = 1;
        d.src[0] = *src0; if (src1 && src1->ptr) d.src[1] = *src1;
        d.cin_g = cin; d.cout_g = cout; d.in_stride = 1;
        d.w = (const char*)w_packed + woff * tbi_dtype_size(dtype);
        woff += (size_t)d.ntaps * cin * cout;
        d.epi = *epi;
        d.epi.out_stride = 2; d.epi.out_off_y = ph >> 1; d.epi.out_off_x = ph & 1;
        int rc = tbi_tapgemm_run(&d, stream); if (rc) return rc;
    }
    return TBI_OK;
}

extern "C" int tbi_conv2d_transpose_s2_dgrad(int dtype, int impl, int n, int h, int w, int ksize, const tbi_view* dz,
                                             int cin_total, const void* w_packed_dgrad, const tbi_epilogue* epi, void* stream) {
    TBI_CHECK(ksize == 3 || ksize == 4, TBI_ERR_UNSUPPORTED, "convT: ksize %d", ksize);
    TBI_CHECK(dz->h == 2 * h && dz->w == 2 * w, TBI_ERR_BAD_SHAPE, "convT dgrad: dz dims %dx%d != 2x(%dx%d)", dz->h, dz->w, h, w);
    const int pad = ksize == 4 ? 1 : 0;
    tbi_tapgemm d; memset(&d, 0, sizeof(d));
    d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = 1;
    d.src[0] = *dz; d.cin_g = dz->c; d.cout_g = cin_total; d.in_stride = 2;
    d.ntaps = 0;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) { d.dy[d.ntaps] = ky - pad; d.dx[d.ntaps] = kx - pad; ++d.ntaps; }
    d.w = w_packed_dgrad; d.epi = *epi;
    if (d.epi.out_stride == 0) d.epi.out_stride = 1;
    return tbi_tapgemm_run(&d, stream);
}

extern "C" int tbi_conv2d_transpose_s2_wgrad(int dtype, int impl, int n, int h, int w, int ksize, const tbi_view* x0,
                                             const tbi_view* x1, const tbi_view* dz, int cout, float* dw_hwoi, float* dbias,
                                             void* workspace, int64_t workspace_bytes, void* stream) {
    TBI_CHECK(ksize == 3 || ksize == 4, TBI_ERR_UNSUPPORTED, "convT: ksize %d", ksize);
    TBI_CHECK(dz->h == 2 * h && dz->w == 2 * w, TBI_ERR_BAD_SHAPE, "convT wgrad: dz dims");
    TBI_CHECK(cout >= 1 && cout <= dz->c, TBI_ERR_BAD_SHAPE, "convT wgrad: cout %d vs dz channels %d", cout, dz->c);
    const int pad = ksize == 4 ? 1 : 0;
    tbi_tapwgrad d; memset(&d, 0, sizeof(d));
    d.dtype = dtype; d.impl = impl; d.n = n; d.gh = h; d.gw = w; d.groups = 1;
    d.a_src[0] = *x0; if (x1 && x1->ptr) d.a_src[1] = *x1;
    d.b_src = *dz;
    const int cin = x0->c + ((x1 && x1->ptr) ? x1->c : 0);
    d.cin_g = cin; d.cout_g = cout; d.a_stride = 1; d.b_stride = 2;
    d.ntaps = 0;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) { d.b_dy[d.ntaps] = ky - pad; d.b_dx[d.ntaps] = kx - pad; ++d.ntaps; }
    d.dw = dw_hwoi;
    d.tap_stride = (int64_t)cin * cout; d.ci_stride = 1; d.co_stride = cin;
    d.dbias = nullptr;                   // a tap of a stride-2 gather does not visit every dz pixel: use colsum
    d.workspace = workspace; d.workspace_bytes = workspace_bytes;
    int rc = tbi_tapwgrad_run(&d, stream); if (rc) return rc;
    if (dbias) { tbi_view real = *dz; real.c = cout; return tbi_colsum(dtype, (int64_t)n * dz->h * dz->w, &real, dbias, stream); }
    return TBI_OK;
}
