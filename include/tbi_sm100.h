/* tbi_sm100.h -- C ABI of libtbi_sm100.so: the B200 (sm_100a) kernels under the TBI_ResNest hot path.
 *
 * The reference (silverlight6/Ultrasound_Modeling) has no FFI layer: its model code calls
 * tf.keras.layers.* directly and TensorFlow picks cuDNN/XLA kernels.  This ABI sits where those
 * library kernels sat.  Each entry point names the reference call site(s) it replaces.
 *
 * Conventions:
 *   - activations NHWC, contiguous; a tensor may be addressed as a channel slice of a wider
 *     pixel record (cstride = elements per pixel record, coff = first channel), which is how
 *     tf.concat (TBI_ResNest.py:110-122,139) is never materialised;
 *   - the caller owns every buffer, including workspaces; the library never allocates, frees or
 *     synchronises; every call enqueues on the cudaStream_t it is handed (passed as void*);
 *   - every call returns 0 or a negative TBI_ERR_*; tbi_last_error() gives the thread-local text;
 *     an unsupported shape is an error, never a fallback;
 *   - dtype is the STORAGE type of activations and packed weights (TBI_F32 | TBI_BF16);
 *     accumulation is always fp32; parameter gradients and optimizer state are always fp32.
 */
#ifndef TBI_SM100_H
#define TBI_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBI_VERSION 100

enum { TBI_OK = 0, TBI_ERR_BAD_SHAPE = -1, TBI_ERR_BAD_ALIGN = -2, TBI_ERR_UNSUPPORTED = -3, TBI_ERR_CUDA = -4 };
enum { TBI_F32 = 0, TBI_BF16 = 1 };
enum { TBI_ACT_NONE = 0, TBI_ACT_ELU = 1, TBI_ACT_LRELU = 2 /* slope 0.3 */, TBI_ACT_RELU = 3 };
/* which implementation executes a tap-GEMM: AUTO picks tcgen05 when the shape qualifies */
enum { TBI_IMPL_AUTO = 0, TBI_IMPL_SIMT = 1, TBI_IMPL_TCGEN05 = 2 };

#define TBI_MAX_TAPS 16

/* A (possibly strided) channel-slice view of an NHWC tensor. */
typedef struct {
    void*   ptr;       /* base of the tensor (element 0 of pixel 0)                    */
    int32_t h, w;      /* spatial dims                                                 */
    int32_t c;         /* channels in this view                                        */
    int32_t cstride;   /* elements per pixel record (>= coff + c)                      */
    int32_t coff;      /* first channel of the view inside the record                  */
} tbi_view;

/* Fused epilogue of a tap-GEMM, applied to the fp32 accumulator v of output element (pixel, co):
 *   v += bias[co]                        (bias != NULL; BN-affine is folded into weights+bias)
 *   v *= keep[pixel,co]                  (drop_keep != NULL; uint8 MULTIPLIER: 0 = dropped, 2 = kept (1/(1-rate)),
 *                                         1 = dropout disabled; tf.nn.dropout(x,0.5), TBI_ResNest.py:216)
 *   v  = act(v)
 *   v += residual[pixel,co]              (residual.ptr != NULL; may alias out -> accumulate)
 *   v *= act'(dact_ref[pixel,co]) [* dact_keep]     (dact != NONE; derivative from the stored
 *                                         forward OUTPUT; fuses the previous layer's activation
 *                                         backward into this dgrad)
 * Output channels >= split_c go to out2 instead (dgrad of a virtual concat); out2 takes only the
 * residual2 add (accumulation of a skip-connection gradient), no dact.
 */
typedef struct {
    const float*   bias;
    const uint8_t* drop_keep;     /* indexed like out (dense, c == cstride)                   */
    int32_t        act;
    tbi_view       residual;      /* ptr NULL = none                                          */
    int32_t        dact;          /* TBI_ACT_* whose derivative to apply, from dact_ref       */
    tbi_view       dact_ref;
    const uint8_t* dact_keep;     /* dropout keep-mask that followed dact_ref's layer, or NULL */
    tbi_view       out;           /* h,w = dims of the output tensor                          */
    int32_t        out_f32;       /* 1: out is float32 whatever dtype is (head logits)        */
    int32_t        out_stride;    /* 1, or 2 for a transposed-conv phase                      */
    int32_t        out_off_y, out_off_x;
    int32_t        split_c;       /* 0 = no split                                             */
    tbi_view       out2;
    tbi_view       residual2;
} tbi_epilogue;

/* One tap-GEMM:  acc[pixel p=(n,gy,gx)][co] = sum_tap sum_ci  in[n, gy*in_stride+dy[tap], gx*in_stride+dx[tap], ci]
 *                                                          * w[co][tap*cin_g + ci]
 * (out-of-range input pixels read as 0).  The input is the virtual concat of src[0] and src[1]
 * along channels.  groups > 1: input channel block g feeds output channel block g (single source).
 * Packed weights are K-major: [cout_total][ntaps*cin_g] in `dtype`.
 * Every convolution of the path is a list of these (tbi_conv2d_* / tbi_conv2d_transpose_s2_* build them).
 */
typedef struct {
    int32_t  dtype;
    int32_t  impl;             /* TBI_IMPL_*                                                  */
    int32_t  n, gh, gw;        /* pixel grid (GEMM M = n*gh*gw)                               */
    int32_t  groups;
    int32_t  cin_g, cout_g;    /* per-group K channels (sum over sources) and output channels */
    tbi_view src[2];           /* src[1].ptr NULL = single source                             */
    int32_t  in_stride;        /* 1, or 2 (dgrad of a stride-2 transposed conv)               */
    int32_t  ntaps;
    int32_t  dy[TBI_MAX_TAPS], dx[TBI_MAX_TAPS];
    const void* w;
    /* nphase == 4: the four output-parity phases of a k=4 stride-2 transposed conv in ONE launch.
     * Phase p uses taps ph_dy/ph_dx[p][0..ntaps), writes to (2*gy+ph_off_y[p], 2*gx+ph_off_x[p]) and its
     * weights are rows [p*cout_total, (p+1)*cout_total) of w.  nphase <= 1: dy/dx/epi offsets as given. */
    int32_t  nphase;
    int32_t  ph_dy[4][4], ph_dx[4][4];
    int32_t  ph_off_y[4], ph_off_x[4];
    tbi_epilogue epi;
} tbi_tapgemm;

/* Weight-gradient tap-GEMM (fp32 accumulate, atomically ADDED into dw, which the caller zeroes):
 *   dw[tap*tap_stride + ci*ci_stride + co*co_stride] +=
 *        sum_p a[n, gy*a_stride+a_dy[tap], gx*a_stride+a_dx[tap], ci] * b[n, gy*b_stride+b_dy[tap], gx*b_stride+b_dx[tap], co]
 * a = forward input (virtual concat of a_src[0..1]), b = gradient w.r.t. the pre-activation output.
 * groups as above (ci in [0,cin_g) of group g pairs with co in group g; ci index in dw is group-local,
 * co index is global).  dw strides let the result land directly in Keras HWIO / HWOI layout.
 */
typedef struct {
    int32_t  dtype;
    int32_t  impl;
    int32_t  n, gh, gw;
    int32_t  groups, cin_g, cout_g;
    tbi_view a_src[2];
    tbi_view b_src;
    int32_t  a_stride, b_stride;
    int32_t  ntaps;
    int32_t  a_dy[TBI_MAX_TAPS], a_dx[TBI_MAX_TAPS], b_dy[TBI_MAX_TAPS], b_dx[TBI_MAX_TAPS];
    float*   dw;
    int64_t  tap_stride, ci_stride, co_stride;
    float*   dbias;            /* optional: dbias[co] += sum_p b[p,co]  (computed once, tap 0)    */
    void*    workspace;        /* tcgen05 split-K partials; tbi_workspace_bytes() tells how much  */
    int64_t  workspace_bytes;
} tbi_tapwgrad;

int         tbi_version(void);
const char* tbi_last_error(void);
/* 1 if the running device is sm_100 and the tcgen05 kernels can launch; 0 otherwise */
int         tbi_device_ok(void);

/* How many bf16 tap-GEMM / weight-gradient launches issued with TBI_IMPL_AUTO ran on the CUDA-core kernel because the tcgen05
 * path refused their shape (cumulative per process; reset != 0 zeroes the counters), and the reason of the last one.     */
int         tbi_fallback_stats(int64_t* tapgemm_simt, int64_t* tapwgrad_simt, int reset);
const char* tbi_last_fallback(void);

/* ---- convolutions ------------------------------------------------------------------------
 * tbi_tapgemm_run / tbi_tapwgrad_run are the single execution entry points; the named conv
 * entry points below fill the descriptors for the shapes the reference uses and call them.
 * replaces: tf.keras.layers.Conv2D fwd + its GradientTape input/filter gradients
 *           (TBI_ResNest.py:83-91,140,143,162-168; ResNest.py:14-24,77-85,122-131; Decoder.py:11-25,103)
 */
int tbi_tapgemm_run(const tbi_tapgemm* d, void* stream);
int tbi_tapwgrad_run(const tbi_tapwgrad* d, void* stream);
int64_t tbi_workspace_bytes(const tbi_tapwgrad* d);

/* Conv2D k in {1,3}, stride 1, SAME, dilation in {1,2,4,8}, groups >= 1, <= 2 sources.
 * w_packed: see tbi_pack_conv_weights (mode FWD for fwd, DGRAD for dgrad).                      */
int tbi_conv2d_fwd(int dtype, int impl, int n, int h, int w, int ksize, int dilation, int groups,
                   const tbi_view* src0, const tbi_view* src1, int cout_total,
                   const void* w_packed, const tbi_epilogue* epi, void* stream);
/* dgrad: dz (cout_total channels) -> dx (cin_total channels) through epi->out / out2.            */
int tbi_conv2d_dgrad(int dtype, int impl, int n, int h, int w, int ksize, int dilation, int groups,
                     const tbi_view* dz, int cin_total, const void* w_packed_dgrad,
                     const tbi_epilogue* epi, void* stream);
/* wgrad into Keras (grouped) HWIO [k,k,cin_g,cout_total] fp32, ADDED; dbias optional.           */
int tbi_conv2d_wgrad(int dtype, int impl, int n, int h, int w, int ksize, int dilation, int groups,
                     const tbi_view* x0, const tbi_view* x1, const tbi_view* dz,
                     float* dw_hwio, float* dbias, void* workspace, int64_t workspace_bytes, void* stream);

/* Conv2DTranspose k in {3,4}, stride 2, TF 'same' (TBI_ResNest.py:124,210; Decoder.py:57-59,120).
 * fwd runs one tap-GEMM per output parity phase (4 phases); input dims h,w -> output 2h,2w.
 * w_packed: tbi_pack_convt_weights (FWD: [phase][cout][taps*cin]; DGRAD: [cin][k*k*cout]).       */
int tbi_conv2d_transpose_s2_fwd(int dtype, int impl, int n, int h, int w, int ksize,
                                const tbi_view* src0, const tbi_view* src1, int cout,
                                const void* w_packed, const tbi_epilogue* epi, void* stream);
int tbi_conv2d_transpose_s2_dgrad(int dtype, int impl, int n, int h, int w, int ksize,
                                  const tbi_view* dz /* 2h x 2w, cout ch */, int cin_total,
                                  const void* w_packed_dgrad, const tbi_epilogue* epi, void* stream);
/* wgrad into Keras HWOI [k,k,cout,cin] fp32, ADDED.                                             */
int tbi_conv2d_transpose_s2_wgrad(int dtype, int impl, int n, int h, int w, int ksize,
                                  const tbi_view* x0, const tbi_view* x1, const tbi_view* dz, int cout,
                                  float* dw_hwoi, float* dbias, void* workspace, int64_t workspace_bytes,
                                  void* stream);
/* (cout <= dz->c: dz may be zero-padded along channels to keep its pixel records 16-byte aligned,
 *  e.g. the 3-class head gradient stored with 16 channels; only the first cout are real.)             */

/* ---- weight packing (master fp32 Keras layout -> K-major compute copies, BN scale folded) ----
 * scale: per-output-channel multiplier (gamma/sqrt(var+eps)) or NULL.
 * mode 0 = FWD  : out[co][tap][ci_g]            = W[tap][ci_g][co] * scale[co]
 * mode 1 = DGRAD: out[g*cin_g+ci][tap'][co_g]   = W[flip(tap')][ci][g*cout_g+co_g] * scale[co]
 * W is (grouped) HWIO [k,k,cin_g,cout_total].                                                   */
int tbi_pack_conv_weights(int dtype, int mode, int ksize, int groups, int cin_g, int cout_total,
                          const float* w_hwio, const float* scale, void* out, void* stream);
/* W is HWOI [k,k,cout,cin].
 * mode 0 = FWD  : out[phase=(a,b)][co][t][ci] with t over the taps of that phase (order = tbi_convt_phase_taps)
 * mode 1 = DGRAD: out[ci][ky*k+kx][co]         = W[ky][kx][co][ci] * scale[co]                   */
int tbi_pack_convt_weights(int dtype, int mode, int ksize, int cin, int cout, int cout_pad,
                           const float* w_hwoi, const float* scale, void* out, void* stream);
/* cout_pad (mode 1 only; 0 or >= cout): the co axis of the DGRAD pack is zero-padded to cout_pad so that it
 * matches a channel-padded gradient tensor.                                                            */
/* Gathered-gradient form of a stride-2 transposed conv's backward (used for the narrow head, f_tran):
 *   G[n,i,j, (ky*k+kx)*cout + co] = dz[n, 2i-pad+ky, 2j-pad+kx, co]   (0 outside dz; channels >= k*k*cout untouched)
 * after which dW_hwoi (flattened [k*k*cout][cin]) = G^T X  and  dX = G W' are plain (1-tap) tap-GEMMs.
 * tbi_pack_convt_weights mode 2 writes W'[ci][q], q = (ky*k+kx)*cout+co, rows zero-padded to cout_pad.       */
int tbi_convt_gather_dz(int dtype, int n, int h, int w, int ksize, int cout, const tbi_view* dz, const tbi_view* g, void* stream);
/* forward counterpart: Y (fp32 or bf16: y_dtype) [n,h,w,>=16*cout] = the k=4 stride-2 transposed conv evaluated as ONE plain GEMM per input pixel
 * (weights packed with tbi_pack_convt_weights mode 3: [q = tap*cout + co][cin]); out fp32 [n,2h,2w,cout] = bias + the four
 * (tap, input pixel) contributions of every output pixel.  Used for the head f_tran (TBI_ResNest.py:124): with cout = 3 the
 * per-output-phase form issues 160 N=16 MMAs per 128 outputs and is MMA-dispatch-bound.                            */
int tbi_convt_scatter_y(int y_dtype, int n, int h, int w, int ksize, int cout, const tbi_view* y, const float* bias,
                        const tbi_view* out, void* stream);
/* taps of output-parity phase (a,b): returns count; ky/kx = kernel index, dy/dx = input offset.  */
int tbi_convt_phase_taps(int ksize, int a, int b, int* ky, int* kx, int* dy, int* dx);

/* One-launch weight preparation: every packing mode above is an index map
 *     out[out_tap[t] + a*out_a + b*out_b] = S[src_tap_index[t]*src_tap + a*src_a + b] * scale(co),  co = co_base + (co_is_a ? a : b)
 * of one tap's fp32 master matrix S[A][B] (b contiguous), scale = gamma/sqrt(var+eps) (BN folded) or 1 (gamma NULL).  The
 * caller builds a table of items once (device memory; tile_begin = running sum of ntaps*tiles_a*tiles_b with
 * tiles_x = ceil(X/32)), zero-fills padded destinations once, and tbi_prepare_run executes the whole table per step;
 * tbi_bn_fold_multi does the same for the per-channel scale / folded-bias vectors of all layers.                         */
typedef struct {
    const float* src; void* out; const float* gamma; const float* var;
    int32_t ntaps, A, B, co_is_a, co_base, tile_begin, tiles_a, tiles_b;
    int64_t src_tap, src_a, out_a, out_b;
    int32_t src_tap_index[TBI_MAX_TAPS];
    int64_t out_tap[TBI_MAX_TAPS];
} tbi_prep_item;
typedef struct { int32_t c, pad_; const float *gamma, *beta, *mean, *var, *bias; float *scale, *fbias; } tbi_fold_item;
int tbi_prepare_run(int dtype, const tbi_prep_item* items_dev, int nitems, int total_tiles, float bn_eps, void* stream);
int tbi_bn_fold_multi(const tbi_fold_item* items_dev, int nitems, float bn_eps, void* stream);

/* Folded BN-inference affine of a conv layer, and the matching parameter gradients.
 * fold:  scale[c] = gamma/sqrt(var+eps);  fbias[c] = (bias-mean)*scale+beta   (gamma NULL: scale=1, fbias=bias)
 * grads (after wgrad produced dw_raw = A^T dz and dbias_raw = colsum(dz), both w.r.t. pre-activation z):
 *   dgamma = (sum_k W[k,c]*dw_raw[k,c] + (bias-mean)*dbias_raw) / sqrt(var+eps);  dbeta = dbias_raw
 *   dw = dw_raw*scale (in place);  dbias = dbias_raw*scale (in place)
 * layout: co_stride/k_count describe dw/W as k_count rows per channel: element (k,c) at
 *   k_outer*outer_stride + c*co_stride + k_inner  with k = k_outer*inner + k_inner  (HWIO: inner=1, co_stride=1,
 *   outer_stride=cout;  HWOI: inner=cin, co_stride=cin, outer_stride=cout*cin).
 * replaces: BatchNormalization fwd/bwd in inference mode (TBI_ResNest.py:90,144,164,169,190,213). */
int tbi_bn_fold(int c, const float* gamma, const float* beta, const float* mean, const float* var,
                const float* bias, float eps, float* scale, float* fbias, void* stream);
int tbi_bn_param_grad(int c, int64_t k_outer, int64_t inner, int64_t outer_stride, int64_t co_stride,
                      const float* w, float* dw, const float* bias, float* dbias,
                      const float* gamma, const float* mean, const float* var, float eps,
                      float* dgamma, float* dbeta, void* stream);

/* ---- bandwidth-bound fused kernels ------------------------------------------------------------ */
/* AveragePooling2D(2,2) (TBI_ResNest.py:92-107; ResNest.py:25-28).  bwd: dx = dy/4 broadcast,
 * optionally accumulated into dx (accumulate=1) and/or multiplied by act'(dact_ref) (stem ELU).  */
int tbi_avgpool2x2_fwd(int dtype, int n, int h, int w, const tbi_view* x, const tbi_view* y, void* stream);
int tbi_avgpool2x2_bwd(int dtype, int n, int h, int w, const tbi_view* dy, const tbi_view* dx,
                       int accumulate, int dact, const tbi_view* dact_ref, void* stream);

/* Radix split-attention tail (TBI_ResNest.py:175-207; ResNest.py:153-199).
 * u: [n,h,w, K*R*c] channel order (k,r,c); v: [n,h,w,K*c].  Parameters per cardinal k:
 *   w1 [K][c][c/2], b1 [K][c/2], BN (gamma,beta,mean,var) [K][c/2], w2 [K][R][c/2][c], b2 [K][R][c].
 * softmax over the CHANNEL axis when R>1, sigmoid when R==1 (reference quirk, kept).
 * gap [n][K][c], h1 [n][K][c/2], att [n][K][R][c] are fp32 buffers the caller keeps for bwd.
 * fwd = gap pass + fc + recombine pass;  bwd = da pass + fc-bwd + dU pass (dU already multiplied
 * by act'(u): it is the gradient w.r.t. the pre-activation of the conv that produced u).         */
typedef struct {
    int32_t dtype, n, h, w, kpaths, radix, c;
    int32_t act;                       /* activation of dense1 and of the producer of u (ELU | LRELU) */
    float   bn_eps;
    const float *w1, *b1, *gamma, *beta, *mean, *var, *w2, *b2;
    float *gap, *h1, *att;
} tbi_splitatt;
int tbi_split_attention_fwd(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, void* stream);
/* scratch: fp32 [n][K][R][c] (da) + [n][K][c] (dgap).  Parameter grads are ADDED.                */
int tbi_split_attention_bwd(const tbi_splitatt* p, const tbi_view* u, const tbi_view* dv, const tbi_view* du,
                            float* dw1, float* db1, float* dgamma, float* dbeta, float* dw2, float* db2,
                            float* scratch, void* stream);
/* the two bandwidth passes alone (microbench config 2 times exactly these) */
int tbi_splitatt_gap(const tbi_splitatt* p, const tbi_view* u, void* stream);
int tbi_splitatt_combine(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, void* stream);

/* Softmax over classes + my_loss_cat (TBI_ResNest.py:125,234-248) + argmax accuracy (:48-51).
 * logits fp32 [n,h,w,nc]; y fp32 one-hot/soft [n,h,w,nc]; probs fp32 out; loss_map fp32 [h,w];
 * correct: int32 counter (ADDED; caller zeroes); dlogits (may be NULL; stored as dlogits_dtype)
 * = d sum(loss_map) / d logits.                                                                  */
int tbi_softmax_loss_fwd_bwd(int dlogits_dtype, int n, int h, int w, int nc, const float* logits, const float* y,
                             float* probs, float* loss_map, int32_t* correct, void* dlogits, int dlogits_cstride,
                             void* stream);
/* dlogits_cstride >= nc: elements per pixel record of dlogits (channels >= nc are left untouched).     */
/* The same with the logits FORMED inside the kernel from the head's per-input-pixel tap products (f_tran as a plain GEMM,
 * TBI_ResNest.py:124): ytaps fp32 [n, h/2, w/2, >= 16*nc] with column (ky*4+kx)*nc + c, bias fp32 [nc] or NULL;
 * logits[n,oy,ox,c] = bias[c] + sum of the 2 x 2 taps with 2i-1+ky = oy, 2j-1+kx = ox -- what tbi_convt_scatter_y would write,
 * bit for bit -- so the [n,h,w,nc] logits tensor is never materialised.  (h, w) = OUTPUT grid.                          */
int tbi_softmax_loss_fwd_bwd_taps(int dlogits_dtype, int n, int h, int w, int nc, const tbi_view* ytaps, const float* bias,
                                  const float* y, float* probs, float* loss_map, int32_t* correct, void* dlogits,
                                  int dlogits_cstride, void* stream);

/* dz = dy * act'(y_ref) [* keep]     (standalone activation backward where it cannot be fused)   */
int tbi_act_bwd(int dtype, int64_t npix, int act, const tbi_view* dy, const tbi_view* y_ref,
                const uint8_t* keep, const tbi_view* dz, void* stream);
/* dst += src  (identity shortcut gradient, TBI_ResNest.py:148 when Cin == Cout)                   */
int tbi_accumulate(int dtype, int64_t npix, const tbi_view* src, const tbi_view* dst, void* stream);
/* out[c] += sum_p x[p,c]  (bias gradient)                                                        */
int tbi_colsum(int dtype, int64_t npix, const tbi_view* x, float* out, void* stream);
/* keep-multiplier (0|2) generation for the always-on dropout (counter-based hash RNG); the stream position is
 * (*step_ptr)*count + i so that a CUDA-graph replay draws a fresh mask each step (step_ptr may be NULL) */
int tbi_dropout_mask(uint8_t* keep, int64_t count, uint64_t seed, const int32_t* step_ptr, void* stream);
/* Grouped convolutions whose per-group channel counts are too small for one tensor-core K step (the cardinal 3x3 convs of
 * the reference defaults radix=4,kpaths=4: 2, 4, 8 input channels per group; TBI_ResNest.py:157-168) are run as DENSE
 * convolutions over block-diagonal weights: 16x the flops of the grouped form, still ~50x faster than the CUDA-core path.
 * tbi_conv_dense_expand says whether a (dtype, groups, cin/group, cout/group) layer is treated that way by
 * tbi_pack_conv_weights / tbi_conv2d_{fwd,dgrad,wgrad}; tbi_conv_packed_elems is the element count of ONE packed copy
 * (either mode) the caller must allocate; tbi_conv2d_wgrad_workspace the scratch bytes tbi_conv2d_wgrad needs for it
 * (0 when not expanded; without the scratch the wgrad falls back to the grouped CUDA-core kernel).               */
int     tbi_conv_dense_expand(int dtype, int groups, int cin_g, int cout_g);
int64_t tbi_conv_packed_elems(int dtype, int ksize, int groups, int cin_g, int cout_total);
int64_t tbi_conv2d_wgrad_workspace(int dtype, int ksize, int groups, int cin_total, int cout_total);

/* ---- Variant B (ResNest.py / Decoder.py) ---------------------------------------------------------
 * LayerNormalization over the channel axis (Keras axis=-1, biased variance) fused with the activation
 * that follows it: y = act(gamma*(x-mean_c)/sqrt(var_c+eps)+beta).  x, y: npix pixel records of c channels
 * (views; may alias).  Replaces tf.keras.layers.LayerNormalization + LeakyReLU at ResNest.py:86-87,99-101,
 * 126-127,132-133,166-167 and Decoder.py:112-113,130-131.                                          */
int tbi_layernorm_c_fwd(int dtype, int64_t npix, int c, const tbi_view* x, const float* gamma, const float* beta,
                        float eps, int act, const tbi_view* y, void* stream);
/* backward of the above: x = the LayerNorm INPUT, y = its activated output (for act'), dy -> dx;
 * dgamma/dbeta are ADDED.                                                                          */
int tbi_layernorm_c_bwd(int dtype, int64_t npix, int c, const tbi_view* x, const tbi_view* y, const tbi_view* dy,
                        const float* gamma, float eps, int act, const tbi_view* dx, float* dgamma, float* dbeta,
                        void* stream);
/* split attention of ResNest.py:171-199: the R inputs are one and the same tensor U (cardinal.forward :136-147
 * applies the same conv1/conv2 R times) and dense2 is shared, so V = R * U * a with
 * a = softmax_c | sigmoid (R==1) of dense2(act(LN(dense1(R * mean_hw U)))).
 * u, v: [n,h,w,K*c]; w1 [K][c][c/2], b1 [K][c/2], LN gamma/beta [K][c/2], w2 [K][c/2][c], b2 [K][c];
 * att: fp32 scratch [n][K*c] (holds R*a on return).                                                 */
int tbi_splitatt_shared_fwd(int dtype, int n, int h, int w, int kpaths, int radix, int c, const tbi_view* u,
                            const tbi_view* v, const float* w1, const float* b1, const float* ln_gamma,
                            const float* ln_beta, float ln_eps, int act, const float* w2, const float* b2,
                            float* att, int cgroup, void* stream);
/* cgroup (0 = c): channels between the first channels of consecutive cardinals in the u / v / du / dv records.  With
 * cgroup > c every cardinal's slice is zero-padded (c = 10 in a 16-channel slot), which keeps the slices 16-byte aligned
 * for the tensor-core convolutions on either side; the pad lanes are neither read nor written here.              */

/* backward of the above.  att = the forward's att buffer (R*a); scratch: fp32 [2][n][K*c].  dV -> dU; the parameter
 * gradients (same shapes as the parameters) are ADDED.  One pass over (U, dV) for sum_p U and sum_p dV*U, the FC chain
 * and its backward per (image, cardinal), one pass dU = dV*att + dgap*R/HW.                            */
int tbi_splitatt_shared_bwd(int dtype, int n, int h, int w, int kpaths, int radix, int c, const tbi_view* u,
                            const tbi_view* dv, const tbi_view* du, const float* w1, const float* b1,
                            const float* ln_gamma, const float* ln_beta, float ln_eps, int act, const float* w2,
                            const float* att, float* dw1, float* db1, float* dln_gamma, float* dln_beta,
                            float* dw2, float* db2, float* scratch, int cgroup, void* stream);

/* ---- ViT bridge of Variant B (VisionTransformer.py:9-190) -------------------------------------------------------
 * Multi-head self-attention core for short sequences (the reference: 80 tokens, 4 heads of 128 channels): q, k, v, ctx are
 * [n, tokens, heads*head_dim] in `dtype` (head h = channels [h*head_dim, (h+1)*head_dim), the reference's split_heads :26-31),
 * probs fp32 [n, heads, tokens, tokens] = softmax_keys(q k^T * scale) (returned to the caller as the attention weights and
 * kept for the backward pass); ctx = probs v.  The reference scales by 1/sqrt(num_heads) (:42), so scale is an argument.
 * One CTA per (image, head) holds Q, K, V and the score matrix in shared memory: tokens*(head_dim+1)*12 + tokens^2*4 bytes
 * must fit in 220 KB (16 for the backward).  replaces: Attention.forward :33-54 between the four Dense layers.            */
int tbi_attention_fwd(int dtype, int n, int tokens, int heads, int head_dim, float scale, const void* q, const void* k,
                      const void* v, void* ctx, float* probs, void* stream);
int tbi_attention_bwd(int dtype, int n, int tokens, int heads, int head_dim, float scale, const void* q, const void* k,
                      const void* v, const float* probs, const void* dctx, void* dq, void* dk, void* dv, void* stream);
/* exact GELU, 0.5 x (1 + erf(x / sqrt 2)) (tf.keras.activations.gelu, Mlp.forward :67-73); bwd: dx = dy * gelu'(x)        */
int tbi_gelu_fwd(int dtype, int64_t count, const void* x, void* y, void* stream);
int tbi_gelu_bwd(int dtype, int64_t count, const void* x, const void* dy, void* dx, void* stream);
/* softmax over nc classes + tf.keras.losses.CategoricalCrossentropy(label_smoothing, reduction=NONE) on the probabilities +
 * tf.nn.compute_average_loss(global_batch_size) (VisionTransformer.py:205-206,225-227): *loss_sum += sum over the npix pixels
 * of the per-pixel loss / global_batch (caller zeroes it); probs fp32 [npix, nc]; dlogits (may be NULL) = d loss / d logits.
 * This is the label-smoothed variant of tbi_softmax_loss_fwd_bwd named in SURVEY 8(b).                                    */
int tbi_softmax_cce_fwd_bwd(int64_t npix, int nc, float label_smoothing, float global_batch, const float* logits,
                            const float* y, float* probs, float* loss_sum, float* dlogits, void* stream);

/* ---- device-side data path and evaluator epilogue --------------------------------------------
 * label2vec (Dataset.py:41-52, Dataset_2.py:6-20): scalar label map [npix] -> soft one-hot fp32 [npix, num_classes]
 *   3 classes: c2 = min(label-1, 1) where label >= 1.05; c1 = 1 - c2 where label > 0.95; c0 = 1 where label <= 0.95
 *   2 classes: (1 - label, label)                                                                                     */
int tbi_label2vec(int64_t npix, int num_classes, const float* label, float* y, void* stream);
/* DataAugs.dataAug (DataAugs.py:82-102) for a whole device-resident batch, out of place: x [n,h,w,c], label [n,h,w] fp32.
 * params: int32 [n][16], one sample's decisions in the order the reference draws them:
 *   [0] imageReduc on (r%3 != 0)  [1] clips (r%3)  [2..5],[6..9] clip k = (row, column, half height, half width)
 *   [10] shift on (t%2)  [11] rows  [12] columns  [13] direction  [14] noise on (t%3 != 0)
 * The stages are evaluated per output pixel at the source pixel it reads: imageReduc zeroes the image where the label is 0
 * (its erosion loop never fires in the reference), clip zeroes image and label in a rectangle, shift translates both with
 * zero fill, noisy adds N(0,1)/5000 (counter-based generator keyed by seed).  The reference's loop bounds (rows/columns
 * H-1 / W-1 are skipped by clip and shift) are reproduced.                                                                */
int tbi_data_aug(int n, int h, int w, int c, const float* x, const float* label, const int32_t* params, uint64_t seed,
                 float* x_out, float* label_out, void* stream);
/* evaluator epilogue (TBIEvaluator.py:238-252): probs = softmax(logits) (may be NULL), prob_out = probs[..., -1],
 * prob_o = 1 - p0 - 0.5 p1 + p2 (may be NULL)                                                                             */
int tbi_softmax_prob_maps(int64_t npix, int nc, const float* logits, float* probs, float* prob_out, float* prob_o, void* stream);
/* brain-mask pre-pass (TBIEvaluator.py:225-231): x[pixel, :] = 0 where round(mask_probs[pixel, 0]) == 1                  */
int tbi_apply_brain_mask(int64_t npix, int mask_classes, int c, const float* mask_probs, float* x, void* stream);

/* x fp32/fp64 host-layout NHWC -> storage dtype (device to device)                               */
int tbi_cast(int src_is_f32, int dst_dtype, int64_t count, const void* src, void* dst, void* stream);

/* Keras Adam (TBI_ResNest.py:28,46): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).
 * step_count: device int32 holding t-1; the kernel reads it; tbi_adam_advance increments it.
 * grad_scale multiplies gradients first (1/world for data-parallel averaging).                   */
int tbi_adam_multi(int64_t count, float* p, const float* g, float* m, float* v, const int32_t* step_count,
                   float lr, float b1, float b2, float eps, float grad_scale, void* stream);
int tbi_adam_advance(int32_t* step_count, void* stream);
/* Same update with the hyper-parameters read from DEVICE memory: hyper = {lr, grad_scale, clip_norm} (fp32[3]).  A captured
 * CUDA graph then follows learning-rate changes (the reference assigns schedules to .learning_rate, MainParallel.py:74-79).
 * clip_norm > 0 with gnorm_sq != NULL applies tf.clip_by_global_norm (VisionTransformer.py:244):
 *   g *= clip_norm / max(sqrt(*gnorm_sq) * |grad_scale|, clip_norm),  *gnorm_sq = sum of squares of the UNSCALED gradient
 * which tbi_sumsq accumulates (out += sum x^2; the caller zeroes out).                                       */
int tbi_adam_multi_dev(int64_t count, float* p, const float* g, float* m, float* v, const int32_t* step_count,
                       const float* hyper, const float* gnorm_sq, float b1, float b2, float eps, void* stream);
int tbi_sumsq(int64_t count, const float* x, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TBI_SM100_H */
