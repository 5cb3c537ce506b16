// Device-side data path and evaluator epilogue: what the reference does per pixel in Python loops on the host
// (DataAugs.py:6-102, Dataset.py:41-52 / Dataset_2.py:6-20, TBIEvaluator.py:225-252), as bandwidth-bound kernels over device-
// resident batches.  At the step rates of this path (>7 k img/s per GPU) the reference's O(H*W) Python loops per sample would
// be the bottleneck by 4-5 orders of magnitude.
#include "tbi_common.cuh"

namespace {

// label2vec: soft one-hot of the scalar label map (Dataset.py:41-52, Dataset_2.py:6-20)
//   3 classes: c2 = min(label - 1, 1) where label >= 1.05 else 0;  c1 = 1 - c2 where label > 0.95 else 0;  c0 = 1 where label <= 0.95 else 0
//   2 classes: (1 - label, label)
__global__ void label2vec_kernel(long long npix, int nc, const float* __restrict__ label, float* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const float l = label[i];
        if (nc == 3) {
            float c2 = l >= 1.05f ? l - 1.f : 0.f;
            c2 = c2 > 1.f ? 1.f : c2;
            y[i * 3 + 0] = l <= 0.95f ? 1.f : 0.f;
            y[i * 3 + 1] = l > 0.95f ? 1.f - c2 : 0.f;
            y[i * 3 + 2] = c2;
        } else {
            y[i * 2 + 0] = 1.f - l; y[i * 2 + 1] = l;
        }
    }
}

// One sample's augmentation decisions, drawn on the host in the reference's order (DataAugs.dataAug :82-102):
//   [0] reduce (r % 3 != 0)   [1] number of clips (r % 3)   [2..5], [6..9] clip k: centre row, centre column, half height, half width
//   [10] shift (t % 2)        [11] shift rows  [12] shift columns  [13] direction (1: read from (i+r, j+c), 0: from (i-r, j-c))
//   [14] noise (t % 3 != 0)   [15] unused
constexpr int AUG_WORDS = 16;

__device__ __forceinline__ float gauss_hash(unsigned long long seed, unsigned long long idx) {
    auto mix = [](unsigned long long z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
    const unsigned long long a = mix(seed + 0x9E3779B97F4A7C15ull * (2 * idx + 1)), b = mix(seed + 0x9E3779B97F4A7C15ull * (2 * idx + 2));
    const float u1 = ((float)(a >> 40) + 0.5f) * (1.f / 16777216.f), u2 = ((float)(b >> 40) + 0.5f) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);             // Box-Muller
}

// out[n,i,j,:] after imageReduc -> clip x nclip -> shift -> noisy, composed per OUTPUT pixel (each stage of the reference is a
// pure function of the previous image, so the chain is evaluated at the one source pixel the output pixel reads):
//   imageReduc (:52-79): its erosion loop never fires (`mask[i,j] > 1` on a 0/1 mask), what remains is "zero every input
//     channel where the LABEL is 0"; the label itself is unchanged.
//   clip (:26-37): zero image and label inside |i - r| < ra and |j - c| < ca, for i < H-1 and j < W-1 only (loop bounds).
//   shift (:6-23): out[i,j] = in[i +- r, j +- c] for i < H-1, j < W-1 when the source is inside the image, else 0.
//   noisy (:40-49): + N(0,1) / 5000 on every input channel.
__global__ void data_aug_kernel(int n, int h, int w, int c, const float* __restrict__ x, const float* __restrict__ label, const int* __restrict__ params,
                                unsigned long long seed, float* __restrict__ xo, float* __restrict__ lo) {
    const long long npix = (long long)n * h * w;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(p % w); const int i = (int)((p / w) % h); const int b = (int)(p / ((long long)w * h));
        const int* q = params + (size_t)b * AUG_WORDS;
        int si = i, sj = j; bool live = true;
        if (q[10]) {
            const int sg = q[13] ? 1 : -1;
            si = i + sg * q[11]; sj = j + sg * q[12];
            live = i < h - 1 && j < w - 1 && si >= 0 && si < h && sj >= 0 && sj < w;
        }
        float lab = 0.f; bool img_zero = true;
        if (live) {
            const long long sp = ((long long)b * h + si) * w + sj;
            lab = label[sp];
            img_zero = q[0] && lab == 0.f;
            for (int k = 0; k < q[1]; ++k) {
                const int r = q[2 + 4 * k], cc = q[3 + 4 * k], ra = q[4 + 4 * k], ca = q[5 + 4 * k];
                if (si < h - 1 && sj < w - 1 && r + ra > si && si > r - ra && cc + ca > sj && sj > cc - ca) { img_zero = true; lab = 0.f; }
            }
            for (int ch = 0; ch < c; ++ch) {
                float v = img_zero ? 0.f : x[sp * c + ch];
                if (q[14]) v += gauss_hash(seed, (unsigned long long)p * c + ch) * (1.f / 5000.f);
                xo[p * c + ch] = v;
            }
        } else {
            for (int ch = 0; ch < c; ++ch) xo[p * c + ch] = q[14] ? gauss_hash(seed, (unsigned long long)p * c + ch) * (1.f / 5000.f) : 0.f;
        }
        lo[p] = lab;
    }
}

// evaluator epilogue (TBIEvaluator.py:238-252): probs = softmax(logits); probOut = probs[..., -1]; probO = 1 - p0 - 0.5 p1 + p2
__global__ void softmax_prob_maps_kernel(long long npix, int nc, const float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ prob_out,
                                         float* __restrict__ prob_o) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        float z[4], m = -INFINITY, s = 0.f;
        for (int k = 0; k < nc; ++k) { z[k] = logits[i * nc + k]; m = fmaxf(m, z[k]); }
        for (int k = 0; k < nc; ++k) { z[k] = expf(z[k] - m); s += z[k]; }
        const float inv = 1.f / s;
        for (int k = 0; k < nc; ++k) { z[k] *= inv; if (probs) probs[i * nc + k] = z[k]; }
        prob_out[i] = z[nc - 1];
        if (prob_o) prob_o[i] = 1.f - z[0] - 0.5f * z[1] + (nc > 2 ? z[2] : 0.f);
    }
}

// brain-mask pre-pass (TBIEvaluator.py:225-231): x[..., :] = 0 where round(mask_probs[..., 0]) == 1
__global__ void apply_mask_kernel(long long npix, int nc, int c, const float* __restrict__ mask_probs, float* __restrict__ x) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        if (rintf(mask_probs[i * nc]) == 1.f)
            for (int ch = 0; ch < c; ++ch) x[i * c + ch] = 0.f;
    }
}

unsigned dgrid(long long work) {
    long long b = (work + 255) / 256;
    const long long cap = (long long)tbi_sm_count() * 16;
    return (unsigned)(b < 1 ? 1 : b > cap ? cap : b);
}

}  // namespace

extern "C" int tbi_label2vec(int64_t npix, int num_classes, const float* label, float* y, void* stream) {
    TBI_CHECK(label && y && npix > 0 && (num_classes == 2 || num_classes == 3), TBI_ERR_BAD_SHAPE, "label2vec: null argument or num_classes %d (2 or 3)", num_classes);
    label2vec_kernel<<<dgrid(npix), 256, 0, (cudaStream_t)stream>>>(npix, num_classes, label, y);
    TBI_CUDA_LAUNCH_CHECK("label2vec");
    return TBI_OK;
}

extern "C" int tbi_data_aug(int n, int h, int w, int c, const float* x, const float* label, const int32_t* params, uint64_t seed, float* x_out,
                            float* label_out, void* stream) {
    TBI_CHECK(x && label && params && x_out && label_out && n > 0 && h > 1 && w > 1 && c > 0, TBI_ERR_BAD_SHAPE, "data_aug: null / empty argument");
    TBI_CHECK(x != x_out && label != label_out, TBI_ERR_BAD_SHAPE, "data_aug: out of place only (a shifted pixel reads another pixel's input)");
    data_aug_kernel<<<dgrid((long long)n * h * w), 256, 0, (cudaStream_t)stream>>>(n, h, w, c, x, label, params, seed, x_out, label_out);
    TBI_CUDA_LAUNCH_CHECK("data_aug");
    return TBI_OK;
}

extern "C" int tbi_softmax_prob_maps(int64_t npix, int nc, const float* logits, float* probs, float* prob_out, float* prob_o, void* stream) {
    TBI_CHECK(logits && prob_out && npix > 0 && nc >= 2 && nc <= 4, TBI_ERR_BAD_SHAPE, "softmax_prob_maps: null argument or %d classes (2..4)", nc);
    softmax_prob_maps_kernel<<<dgrid(npix), 256, 0, (cudaStream_t)stream>>>(npix, nc, logits, probs, prob_out, prob_o);
    TBI_CUDA_LAUNCH_CHECK("softmax_prob_maps");
    return TBI_OK;
}

extern "C" int tbi_apply_brain_mask(int64_t npix, int mask_classes, int c, const float* mask_probs, float* x, void* stream) {
    TBI_CHECK(mask_probs && x && npix > 0 && mask_classes >= 1 && c > 0, TBI_ERR_BAD_SHAPE, "apply_brain_mask: null / empty argument");
    apply_mask_kernel<<<dgrid(npix), 256, 0, (cudaStream_t)stream>>>(npix, mask_classes, c, mask_probs, x);
    TBI_CUDA_LAUNCH_CHECK("apply_brain_mask");
    return TBI_OK;
}
