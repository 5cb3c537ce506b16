// Fused split-attention tail (TBI_ResNest.py:175-207): ONE launch per direction, one thread-block CLUSTER per image.
//
// The only coupling inside the op is per image (global average pool -> FC chain -> per-channel attention), so an image is
// given to a cluster of CS CTAs (CS <= 8, portable); each CTA owns a contiguous pixel chunk of that image in both bandwidth
// passes.  Nothing is exchanged through global memory and no grid-wide barrier exists:
//
// forward : pass 1   per-channel sums of U over the chunk -> part[] in this CTA's shared memory
//           cluster barrier; every CTA pulls the CS partial vectors through distributed shared memory -> gap
//           FC chain, SPLIT across the cluster: each CTA computes a 1/CS slice of dense1 (+BN+act) and pushes it into every
//           peer's shared memory, barrier, a 1/CS slice of dense2, push, barrier; softmax over channels (sigmoid for R=1)
//           is redundant per CTA (a warp per row).
//           pass 2   V = sum_r U_r * a_r over the SAME chunk with the SAME thread->data mapping as pass 1: the chunk is read
//           back from shared memory when it was small enough to be kept there, otherwise from L2 (it was read microseconds
//           earlier by this very CTA, whatever the total size of U: images are independent, so no tensor-wide L2 residency
//           is needed as it was with a grid-wide barrier).
// backward: pass 1   da[k][r][c] = sum_pixels dV * U_r ; barrier + pull
//           softmax/sigmoid backward (redundant), dh1 slice (+ recomputed pre-BN value for xhat) -> dq push, barrier,
//           dgap slice push, barrier; slice owners write dz / dbn / xhat / dgap of the image to the scratch buffer for the
//           parameter-gradient kernel (reduction over n).
//           pass 2   dU_r = (dV * a_r + dgap / HW) * act'(U_r).
// Images run out of step with each other (no grid barrier), so one cluster's FC chain hides behind other clusters' passes once
// there is more than one wave of clusters.  (Tried and dropped: delaying the odd clusters of the first wave so that one half
// streams while the other half is in its FC chain -- slower at every delay, scratch/trace_splitatt.py shows why: half the CTAs do
// not saturate HBM.)
#include "tbi_common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace {

constexpr int V = 8;                                        // bf16 elements per 16-byte access

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[V]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[V]) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return q;
}
__device__ __forceinline__ uint4 ld16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(__nv_bfloat16* p, const uint4& q) { *reinterpret_cast<uint4*>(p) = q; }

// Thread -> data mapping shared by both passes: lane_c = tid % cvv owns the 8-channel vector co = lane_c*8 of the K*c
// attention channels (all R radix copies of it), lane_p = tid / cvv walks the chunk's pixels with stride pl = NT / cvv.
struct Map {
    int cvv, pl, lane_c, lane_p, co, kk, cc, nit;
};
__device__ __forceinline__ Map make_map(int NT, int KC, int c, int npx) {
    Map m;
    m.cvv = KC / V; m.pl = NT / m.cvv;
    m.lane_c = threadIdx.x % m.cvv; m.lane_p = threadIdx.x / m.cvv;
    m.co = m.lane_c * V; m.kk = m.co / c; m.cc = m.co % c;
    m.nit = npx > 0 ? (npx + m.pl - 1) / m.pl : 0;
    return m;
}

// pass 1 of both directions.  s[r][k] += U_r (x dV); optionally keeps the raw 16-byte vectors in shared memory
// (cache_u[(it*R + r)*NT + tid], cache_d[it*NT + tid]: private to the thread, conflict-free, no barrier needed).
template <int R, bool MUL, int NT>
__device__ __forceinline__ void pass1(const Map& m, const __nv_bfloat16* up, int ucs, const __nv_bfloat16* dp, int dcs, int c,
                                      int npx, float (&s)[R][V], uint4* cache_u, uint4* cache_d) {
    constexpr int UN = MUL ? (NT == 512 ? (R == 1 ? 4 : R == 2 ? 2 : 1) : (R == 1 ? 8 : R == 2 ? 4 : 2)) : (R == 1 ? 8 : R == 2 ? 4 : R == 3 ? 2 : 1);   // 5..8 independent 16-byte loads in flight per thread
    const int tid = threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < V; ++k) s[r][k] = 0.f;
    for (int it = 0; it < m.nit; it += UN) {
        uint4 q[UN][R], g[UN];
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            const int px = m.lane_p + (it + i) * m.pl;
            if (px < npx) {
#pragma unroll
                for (int r = 0; r < R; ++r) q[i][r] = ld16(up + (size_t)px * ucs + r * c);
                if (MUL) g[i] = ld16(dp + (size_t)px * dcs);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) q[i][r] = make_uint4(0u, 0u, 0u, 0u);
                if (MUL) g[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
#pragma unroll
        for (int i = 0; i < UN; ++i) {
            float gv[V];
            if (MUL) unpack8(g[i], gv);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float x[V];
                unpack8(q[i][r], x);
#pragma unroll
                for (int k = 0; k < V; ++k) s[r][k] = MUL ? fmaf(x[k], gv[k], s[r][k]) : s[r][k] + x[k];
            }
            if (cache_u != nullptr && it + i < m.nit) {
#pragma unroll
                for (int r = 0; r < R; ++r) cache_u[((size_t)(it + i) * R + r) * NT + tid] = q[i][r];
                if (MUL) cache_d[(size_t)(it + i) * NT + tid] = g[i];
            }
        }
    }
}

// block-level reduction of the per-thread sums over the pixel lanes, PUSHED into every peer's parts[rank][(kk*R + r)*c + cc + k]
// (distributed shared memory): after cluster barrier A every CTA holds all CS partial vectors locally -- a pull would cost a
// remote round trip (measured ~1.3 us under load) on the critical path between the two passes
template <int R, int NT>
__device__ __forceinline__ void reduce_to_part(const Map& m, int c, float (&s)[R][V], float* tmp, float* parts, cg::cluster_group& cl,
                                               int CS, int rank, int C) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool shfl = m.cvv < 32 && (m.cvv & (m.cvv - 1)) == 0;       // a warp holds 32/cvv pixel lanes of every channel vector
    const int rows = shfl ? NT / 32 : m.pl;
    const int row = shfl ? wid : m.lane_p;
    const int width = m.cvv * V;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (shfl) {
            for (int o = 16; o >= m.cvv; o >>= 1)
#pragma unroll
                for (int k = 0; k < V; ++k) s[r][k] += __shfl_xor_sync(0xffffffffu, s[r][k], o);
        }
        if ((!shfl && m.lane_p < m.pl) || (shfl && lane < m.cvv)) {
            float4* d = reinterpret_cast<float4*>(tmp + (size_t)row * width + m.co);
            d[0] = make_float4(s[r][0], s[r][1], s[r][2], s[r][3]);
            d[1] = make_float4(s[r][4], s[r][5], s[r][6], s[r][7]);
        }
        __syncthreads();
        for (int idx = tid; idx < width; idx += NT) {
            float tot = 0.f;
            for (int q = 0; q < rows; ++q) tot += tmp[(size_t)q * width + idx];
            const int e = rank * C + ((idx / c) * R + r) * c + idx % c;
            for (int rk = 0; rk < CS; ++rk) cl.map_shared_rank(parts, rk)[e] = tot;     // remote stores do not stall the thread
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void slice_of(int total, int rank, int S, int& lo, int& hi) {
    lo = (int)((long long)total * rank / S);
    hi = (int)((long long)total * (rank + 1) / S);
}

// Every thread of every CTA of the cluster arrives (release: its shared-memory writes, local and remote, are ordered before
// the barrier) and waits (acquire).  Subsumes __syncthreads().  The cooperative-groups cluster.sync() wraps the same two
// instructions in two extra CTA barriers.
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// split form
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// out[i] for i in [lo,hi): consecutive threads take consecutive outputs, the input axis is split over `parts` thread groups
// and summed through tmp.  partial(i, part, parts) returns this thread's partial sum; finish(i, sum) consumes the total.
// Must be called by the whole CTA; ends with a CTA barrier.
template <typename F, typename G>
__device__ __forceinline__ void sliced_matvec(int lo, int hi, int n_in, int NT, float* tmp, F partial, G finish) {
    const int tid = threadIdx.x;
    for (int o0 = lo; o0 < hi; o0 += NT) {                  // uniform over the CTA
        const int no = min(NT, hi - o0);
        const int parts = min(NT / no, n_in);
        const int ol = tid % no, part = tid / no;
        if (part < parts) tmp[part * no + ol] = partial(o0 + ol, part, parts);
        __syncthreads();
        if (tid < no) {
            float s = 0.f;
            for (int q = 0; q < parts; ++q) s += tmp[q * no + tid];
            finish(o0 + tid, s);
        }
        __syncthreads();
    }
}

// same with two sums per output
template <typename F, typename G>
__device__ __forceinline__ void sliced_matvec2(int lo, int hi, int n_in, int NT, float* tmp, F partial, G finish) {
    const int tid = threadIdx.x;
    float2* tmp2 = reinterpret_cast<float2*>(tmp);
    for (int o0 = lo; o0 < hi; o0 += NT) {                  // uniform over the CTA
        const int no = min(NT, hi - o0);
        const int parts = min(NT / no, n_in);
        const int ol = tid % no, part = tid / no;
        if (part < parts) tmp2[part * no + ol] = partial(o0 + ol, part, parts);
        __syncthreads();
        if (tid < no) {
            float2 s = make_float2(0.f, 0.f);
            for (int q = 0; q < parts; ++q) { const float2 t = tmp2[q * no + tid]; s.x += t.x; s.y += t.y; }
            finish(o0 + tid, s);
        }
        __syncthreads();
    }
}

struct FcPlan { int s1, s2, staged; unsigned long long* trace; };

// FC layers: a layer whose weights are small is computed REDUNDANTLY by every CTA of the cluster (S = 1: no exchange, no
// cluster barrier); a large one is SLICED over the cluster (S = CS: each CTA computes 1/CS of the outputs, pushes them into
// every peer's shared memory, cluster barrier).  When `staged`, the weights (slices) and per-output parameters were copied
// to shared memory by cp.async at kernel start, behind pass 1, so the chain between the two passes never waits on L2/HBM.

// debug timeline (scratch/trace_splitatt.py): thread 0 of every CTA stamps clock64 at the phase boundaries
unsigned long long* g_sa_trace_host = nullptr;
__device__ __forceinline__ void stamp(unsigned long long* tr, int slot) {
    if (tr != nullptr && threadIdx.x == 0) {
        unsigned long long v;
        if (slot == 0 || slot == 15) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); else v = (unsigned long long)clock64();
        tr[(size_t)blockIdx.x * 16 + slot] = v;
    }
}

template <int R, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) splitatt_fwd_cluster_kernel(tbi_splitatt p, tbi_view u, tbi_view v, int cache, FcPlan fc) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = NT >> 5;
    const int hw = p.h * p.w, c = p.c, c2 = p.c / 2, K = p.kpaths, C = K * R * c, KC = K * c, KC2 = K * c2;
    const int n = blockIdx.x / CS;
    const int ppb = (hw + CS - 1) / CS, pbeg = min(hw, rank * ppb), pend = min(hw, pbeg + ppb), npx = pend - pbeg;
    const int n1max = (KC2 + fc.s1 - 1) / fc.s1, n2max = (C + fc.s2 - 1) / fc.s2;
    int lo1, hi1, lo2, hi2;
    slice_of(KC2, fc.s1 > 1 ? rank : 0, fc.s1, lo1, hi1);   // dense1 outputs of this CTA
    slice_of(C, fc.s2 > 1 ? rank : 0, fc.s2, lo2, hi2);     // dense2 outputs of this CTA
    const int no1 = hi1 - lo1, no2 = hi2 - lo2;
    float* parts = sm;                                      // [CS][C] every CTA's channel sums (pushed by the peers)
    float* g = parts + CS * C;                              // [KC]
    float* h1 = g + KC;                                     // [KC2] (slices pushed by the peers when s1 > 1)
    float* att = h1 + KC2;                                  // [C]   (slices pushed by the peers when s2 > 1)
    float* tmp = att + C;                                   // [NT*V]
    float* sw1 = tmp + NT * V;                              // [c][no1]      staged only
    float* sp1 = sw1 + (size_t)c * n1max;                   // [5][n1max]    b1, gamma, beta, mean, var
    float* sw2 = sp1 + 5 * n1max;                           // [c2][no2]
    float* sb2 = sw2 + (size_t)c2 * n2max;                  // [no2]
    const int staged_floats = fc.staged ? ((n1max * (c + 5) + n2max * (c2 + 1) + 3) & ~3) : 0;
    uint4* cache_u = cache ? reinterpret_cast<uint4*>(sw1 + staged_floats) : nullptr;
    if (fc.staged) {
        for (int idx = tid; idx < c * no1; idx += NT) {
            const int ch = idx / no1, i = lo1 + idx - ch * no1, k = i / c2, j = i - k * c2;
            cp_async4(sw1 + idx, p.w1 + ((size_t)k * c + ch) * c2 + j);
        }
        for (int idx = tid; idx < no1; idx += NT) {
            cp_async4(sp1 + idx, p.b1 + lo1 + idx);
            cp_async4(sp1 + n1max + idx, p.gamma + lo1 + idx);
            cp_async4(sp1 + 2 * n1max + idx, p.beta + lo1 + idx);
            cp_async4(sp1 + 3 * n1max + idx, p.mean + lo1 + idx);
            cp_async4(sp1 + 4 * n1max + idx, p.var + lo1 + idx);
        }
        for (int idx = tid; idx < c2 * no2; idx += NT) {
            const int j = idx / no2, i = lo2 + idx - j * no2, kr = i / c, ch = i - kr * c;
            cp_async4(sw2 + idx, p.w2 + ((size_t)kr * c2 + j) * c + ch);
        }
        for (int idx = tid; idx < no2; idx += NT) cp_async4(sb2 + idx, p.b2 + lo2 + idx);
    }
    const Map m = make_map(NT, KC, c, npx);
    stamp(fc.trace, 0); stamp(fc.trace, 1);
    cluster_arrive();                                       // "this CTA is running": waited for in front of the first remote store
    const __nv_bfloat16* up = (const __nv_bfloat16*)u.ptr + ((size_t)n * hw + pbeg) * u.cstride + u.coff + (size_t)m.kk * R * c + m.cc;
    {
        float s[R][V];
        pass1<R, false, NT>(m, up, u.cstride, nullptr, 0, c, npx, s, cache_u, nullptr);
        stamp(fc.trace, 2);
        cluster_wait();                                     // every peer has started (they arrived at their first instruction)
        reduce_to_part<R, NT>(m, c, s, tmp, parts, cl, CS, rank, C);
    }
    cp_async_wait_all();
    stamp(fc.trace, 3);
    cluster_barrier();                                      // A: every CTA's part[] is complete (and this CTA's staged copies)
    stamp(fc.trace, 4);
    const float inv_hw = 1.f / (float)hw;
    for (int i = tid; i < KC; i += NT) {
        const int k = i / c, ch = i - k * c;
        float s = 0.f;
        for (int rk = 0; rk < CS; ++rk)
#pragma unroll
            for (int r = 0; r < R; ++r) s += parts[rk * C + (k * R + r) * c + ch];
        g[i] = s * inv_hw;
    }
    __syncthreads();
    stamp(fc.trace, 5);
    sliced_matvec(lo1, hi1, c, NT, tmp,                     // dense1 + BN + act
        [&](int i, int part_, int parts) {
            const int k = i / c2, j = i - k * c2;
            const float* gk = g + k * c;
            float s = 0.f;
            if (fc.staged) {
                const float* w = sw1 + (i - lo1);
#pragma unroll 8
                for (int ch = part_; ch < c; ch += parts) s = fmaf(gk[ch], w[ch * no1], s);
            } else {
                const float* w = p.w1 + (size_t)k * c * c2 + j;
#pragma unroll 8
                for (int ch = part_; ch < c; ch += parts) s = fmaf(gk[ch], __ldg(w + (size_t)ch * c2), s);
            }
            return s;
        },
        [&](int i, float s) {
            const int ol = i - lo1;
            const float b1 = fc.staged ? sp1[ol] : p.b1[i], ga = fc.staged ? sp1[n1max + ol] : p.gamma[i];
            const float be = fc.staged ? sp1[2 * n1max + ol] : p.beta[i], mu = fc.staged ? sp1[3 * n1max + ol] : p.mean[i];
            const float va = fc.staged ? sp1[4 * n1max + ol] : p.var[i];
            const float hv = act_apply(p.act, (s + b1 - mu) * (ga * rsqrtf(va + p.bn_eps)) + be);
            if (fc.s1 > 1) { for (int rk = 0; rk < CS; ++rk) cl.map_shared_rank(h1, rk)[i] = hv; }
            else h1[i] = hv;
        });
    if (fc.s1 > 1) cluster_barrier();                       // B: h1 complete everywhere
    stamp(fc.trace, 6);
    sliced_matvec(lo2, hi2, c2, NT, tmp,                    // dense2
        [&](int i, int part_, int parts) {
            const int kr = i / c, ch = i - kr * c, k = kr / R;
            const float* hk = h1 + k * c2;
            float s = 0.f;
            if (fc.staged) {
                const float* w = sw2 + (i - lo2);
#pragma unroll 8
                for (int j = part_; j < c2; j += parts) s = fmaf(hk[j], w[j * no2], s);
            } else {
                const float* w = p.w2 + (size_t)kr * c2 * c + ch;
#pragma unroll 8
                for (int j = part_; j < c2; j += parts) s = fmaf(hk[j], __ldg(w + (size_t)j * c), s);
            }
            return s;
        },
        [&](int i, float s) {
            const float z = s + (fc.staged ? sb2[i - lo2] : p.b2[i]);
            if (fc.s2 > 1) { for (int rk = 0; rk < CS; ++rk) cl.map_shared_rank(att, rk)[i] = z; }
            else att[i] = z;
        });
    if (fc.s2 > 1) cluster_barrier();                       // C: logits complete everywhere; no remote access after this point
    stamp(fc.trace, 7);
    for (int kr = wid; kr < K * R; kr += nwarp) {           // softmax over the channel axis (reference quirk) / sigmoid
        float* a = att + kr * c;
        if (R == 1) {
            for (int ch = lane; ch < c; ch += 32) a[ch] = 1.f / (1.f + expf(-a[ch]));
        } else {
            float mx = -INFINITY;
            for (int ch = lane; ch < c; ch += 32) mx = fmaxf(mx, a[ch]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float ls = 0.f;
            for (int ch = lane; ch < c; ch += 32) { const float e = expf(a[ch] - mx); a[ch] = e; ls += e; }
            const float inv = 1.f / warp_sum(ls);
            for (int ch = lane; ch < c; ch += 32) a[ch] *= inv;
        }
    }
    __syncthreads();
    stamp(fc.trace, 8);
    if (rank == 0) {                                        // keep the per-image state for the backward pass
        for (int i = tid; i < KC; i += NT) p.gap[(size_t)n * KC + i] = g[i];
        for (int i = tid; i < KC2; i += NT) p.h1[(size_t)n * KC2 + i] = h1[i];
        for (int i = tid; i < C; i += NT) p.att[(size_t)n * C + i] = att[i];
    }
    // ---- pass 2: recombine over the same chunk, same mapping, LAST pixels first (the most recently read lines are the ones
    // most likely still in L2 when the tensors in flight exceed it); the chunk's lines are dead after this read (.cs)
    {
        float a[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < V; ++k) a[r][k] = att[(m.kk * R + r) * c + m.cc + k];
        __nv_bfloat16* vp = (__nv_bfloat16*)v.ptr + ((size_t)n * hw + pbeg) * v.cstride + v.coff + m.co;
        constexpr int UN = R == 1 ? 8 : R == 2 ? 4 : R == 3 ? 2 : 1;
        for (int it = ((m.nit - 1) / UN) * UN; it >= 0; it -= UN) {
            uint4 q[UN][R];
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int px = m.lane_p + (it + i) * m.pl;
                if (px < npx) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        q[i][r] = cache_u ? cache_u[((size_t)(it + i) * R + r) * NT + tid]
                                          : __ldcs(reinterpret_cast<const uint4*>(up + (size_t)px * u.cstride + r * c));
                }
            }
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int px = m.lane_p + (it + i) * m.pl;
                if (px < npx) {
                    float o[V];
#pragma unroll
                    for (int k = 0; k < V; ++k) o[k] = 0.f;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float x[V];
                        unpack8(q[i][r], x);
#pragma unroll
                        for (int k = 0; k < V; ++k) o[k] = fmaf(x[k], a[r][k], o[k]);
                    }
                    st16(vp + (size_t)px * v.cstride, pack8(o));
                }
            }
        }
    }
    stamp(fc.trace, 9); stamp(fc.trace, 15);
}

// scratch layout (as tbi_split_attention_bwd): dz [n][K][R][c] | dgap [n][K][c] | dbn [n][K][c2] | xhat [n][K][c2]
#ifndef SA_BWD_MINB
#define SA_BWD_MINB 2
#endif
template <int R, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? SA_BWD_MINB : 2) splitatt_bwd_cluster_kernel(tbi_splitatt p, tbi_view u, tbi_view dv, tbi_view du, float* scratch,
                                                                                      int cache, FcPlan fc) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = NT >> 5;
    const int hw = p.h * p.w, c = p.c, c2 = p.c / 2, K = p.kpaths, C = K * R * c, KC = K * c, KC2 = K * c2, N = p.n;
    const int n = blockIdx.x / CS;
    const int ppb = (hw + CS - 1) / CS, pbeg = min(hw, rank * ppb), pend = min(hw, pbeg + ppb), npx = pend - pbeg;
    const int n1max = (KC2 + fc.s1 - 1) / fc.s1, n2max = (KC + fc.s2 - 1) / fc.s2;
    int lo1, hi1, lo2, hi2;
    slice_of(KC2, fc.s1 > 1 ? rank : 0, fc.s1, lo1, hi1);   // dh1 / dq outputs of this CTA
    slice_of(KC, fc.s2 > 1 ? rank : 0, fc.s2, lo2, hi2);    // dgap outputs of this CTA
    const int no1 = hi1 - lo1, no2 = hi2 - lo2;
    float* parts = sm;                                      // [CS][C] every CTA's da sums (pushed by the peers)
    float* att = parts + CS * C;                            // [C]
    float* dz = att + C;                                    // [C]
    float* gg = dz + C;                                     // [KC]
    float* hh = gg + KC;                                    // [KC2]
    float* dq = hh + KC2;                                   // [KC2] (slices pushed by the peers when s1 > 1)
    float* dg = dq + KC2;                                   // [KC]  (slices pushed by the peers when s2 > 1)
    float* tmp = dg + KC;                                   // [NT*V]
    // staged FC operands, output index fastest with an odd pitch (conflict-free for the copy and for the matvec)
    const int ld1 = n1max | 1, ld2 = n2max | 1;
    float* swz = tmp + NT * V;                              // [R*c][ld1]    W2: (r, ch) x output j
    float* swq = swz + (size_t)R * c * ld1;                 // [c][ld1]      W1: ch x output j
    float* sp1 = swq + (size_t)c * ld1;                     // [4][n1max]    b1, gamma, mean, var
    float* swg = sp1 + 4 * n1max;                           // [c2][ld2]     W1: j x output (k, ch)
    const int staged_floats = fc.staged ? (((R + 1) * c * ld1 + 4 * n1max + c2 * ld2 + 3) & ~3) : 0;
    const Map m = make_map(NT, KC, c, npx);
    stamp(fc.trace, 0); stamp(fc.trace, 1);
    cluster_arrive();                                       // "this CTA is running": waited for in front of the first remote store
    uint4* cache_u = cache ? reinterpret_cast<uint4*>(swz + staged_floats) : nullptr;
    uint4* cache_d = cache ? cache_u + (size_t)m.nit * R * NT : nullptr;
    if (fc.staged) {
        for (int idx = tid; idx < no1 * R * c; idx += NT) { // global address contiguous in ch
            const int ol = idx / (R * c), e = idx - ol * R * c, r = e / c, ch = e - r * c, i = lo1 + ol, k = i / c2, j = i - k * c2;
            cp_async4(swz + (size_t)e * ld1 + ol, p.w2 + (((size_t)k * R + r) * c2 + j) * c + ch);
        }
        for (int idx = tid; idx < no1 * c; idx += NT) {     // global address contiguous in j
            const int ch = idx / no1, ol = idx - ch * no1, i = lo1 + ol, k = i / c2, j = i - k * c2;
            cp_async4(swq + (size_t)ch * ld1 + ol, p.w1 + ((size_t)k * c + ch) * c2 + j);
        }
        for (int idx = tid; idx < no1; idx += NT) {
            cp_async4(sp1 + idx, p.b1 + lo1 + idx);
            cp_async4(sp1 + n1max + idx, p.gamma + lo1 + idx);
            cp_async4(sp1 + 2 * n1max + idx, p.mean + lo1 + idx);
            cp_async4(sp1 + 3 * n1max + idx, p.var + lo1 + idx);
        }
        for (int idx = tid; idx < no2 * c2; idx += NT) {    // rows lo2.. of W1 are contiguous
            const int ol = idx / c2, j = idx - ol * c2;
            cp_async4(swg + (size_t)j * ld2 + ol, p.w1 + (size_t)lo2 * c2 + idx);
        }
    }
    for (int i = tid; i < C; i += NT) att[i] = p.att[(size_t)n * C + i];
    for (int i = tid; i < KC; i += NT) gg[i] = p.gap[(size_t)n * KC + i];
    for (int i = tid; i < KC2; i += NT) hh[i] = p.h1[(size_t)n * KC2 + i];
    const __nv_bfloat16* up = (const __nv_bfloat16*)u.ptr + ((size_t)n * hw + pbeg) * u.cstride + u.coff + (size_t)m.kk * R * c + m.cc;
    const __nv_bfloat16* dp = (const __nv_bfloat16*)dv.ptr + ((size_t)n * hw + pbeg) * dv.cstride + dv.coff + m.co;
    {
        float s[R][V];
        pass1<R, true, NT>(m, up, u.cstride, dp, dv.cstride, c, npx, s, cache_u, cache_d);
        stamp(fc.trace, 2);
        cluster_wait();                                     // every peer has started (they arrived at their first instruction)
        reduce_to_part<R, NT>(m, c, s, tmp, parts, cl, CS, rank, C);
    }
    cp_async_wait_all();
    stamp(fc.trace, 3);
    cluster_barrier();                                      // A
    stamp(fc.trace, 4);
    for (int i = tid; i < C; i += NT) {
        float s = 0.f;
        for (int rk = 0; rk < CS; ++rk) s += parts[rk * C + i];
        dz[i] = s;                                          // da
    }
    __syncthreads();
    stamp(fc.trace, 5);
    for (int kr = wid; kr < K * R; kr += nwarp) {           // softmax / sigmoid backward: da -> dz
        const float* a = att + kr * c; float* z = dz + kr * c;
        if (R == 1) {
            for (int ch = lane; ch < c; ch += 32) z[ch] = a[ch] * (1.f - a[ch]) * z[ch];
        } else {
            float l = 0.f;
            for (int ch = lane; ch < c; ch += 32) l = fmaf(a[ch], z[ch], l);
            const float dot = warp_sum(l);
            for (int ch = lane; ch < c; ch += 32) z[ch] = a[ch] * (z[ch] - dot);
        }
    }
    __syncthreads();
    float* sdz = scratch + (size_t)n * C;
    float* sdg = scratch + (size_t)N * C + (size_t)n * KC;
    float* sdbn = scratch + (size_t)N * C + (size_t)N * KC + (size_t)n * KC2;
    float* sxh = sdbn + (size_t)N * KC2;
    if (rank == 0) for (int i = tid; i < C; i += NT) sdz[i] = dz[i];
    const bool own1 = fc.s1 > 1 || rank == 0, own2 = fc.s2 > 1 || rank == 0;      // who writes the image's FC state to scratch
    sliced_matvec2(lo1, hi1, c, NT, tmp,                    // dh1 = sum_r dz_r W2_r^T and the pre-BN value of dense1
        [&](int i, int part_, int parts) {
            const int k = i / c2, j = i - k * c2, ol = i - lo1;
            const float* z = dz + k * R * c;
            const float* gk = gg + k * c;
            float s = 0.f, q = 0.f;
            if (fc.staged) {
#pragma unroll 4
                for (int e = part_; e < R * c; e += parts) s = fmaf(z[e], swz[(size_t)e * ld1 + ol], s);
#pragma unroll 4
                for (int ch = part_; ch < c; ch += parts) q = fmaf(gk[ch], swq[(size_t)ch * ld1 + ol], q);
            } else {
                for (int e = part_; e < R * c; e += parts) {
                    const int r = e / c, ch = e - r * c;
                    s = fmaf(z[e], __ldg(p.w2 + (((size_t)k * R + r) * c2 + j) * c + ch), s);
                }
                for (int ch = part_; ch < c; ch += parts) q = fmaf(gk[ch], __ldg(p.w1 + ((size_t)k * c + ch) * c2 + j), q);
            }
            return make_float2(s, q);
        },
        [&](int i, float2 sq) {
            const int ol = i - lo1;
            const float b1 = fc.staged ? sp1[ol] : p.b1[i], ga = fc.staged ? sp1[n1max + ol] : p.gamma[i];
            const float mu = fc.staged ? sp1[2 * n1max + ol] : p.mean[i], va = fc.staged ? sp1[3 * n1max + ol] : p.var[i];
            const float istd = rsqrtf(va + p.bn_eps);
            const float d = sq.x * act_grad_from_out(p.act, hh[i]);
            if (own1) { sdbn[i] = d; sxh[i] = (sq.y + b1 - mu) * istd; }
            const float dqv = d * ga * istd;
            if (fc.s1 > 1) { for (int rk = 0; rk < CS; ++rk) cl.map_shared_rank(dq, rk)[i] = dqv; }
            else dq[i] = dqv;
        });
    if (fc.s1 > 1) cluster_barrier();                       // B: dq complete everywhere
    stamp(fc.trace, 6);
    sliced_matvec(lo2, hi2, c2, NT, tmp,                    // dgap = dq W1^T
        [&](int i, int part_, int parts) {
            const int k = i / c, ol = i - lo2;
            const float* dk = dq + k * c2;
            float s = 0.f;
            if (fc.staged) {
#pragma unroll 4
                for (int j = part_; j < c2; j += parts) s = fmaf(dk[j], swg[(size_t)j * ld2 + ol], s);
            } else {
                const float* w1 = p.w1 + (size_t)i * c2;    // row (k*c + ch) of W1
                for (int j = part_; j < c2; j += parts) s = fmaf(dk[j], __ldg(w1 + j), s);
            }
            return s;
        },
        [&](int i, float s) {
            if (own2) sdg[i] = s;
            if (fc.s2 > 1) { for (int rk = 0; rk < CS; ++rk) cl.map_shared_rank(dg, rk)[i] = s; }
            else dg[i] = s;
        });
    if (fc.s2 > 1) cluster_barrier();                       // C: dgap complete everywhere; no remote access after this point
    stamp(fc.trace, 7); stamp(fc.trace, 8);
    // ---- pass 2: dU over the same chunk, same mapping, last pixels first; dead lines read and written with .cs
    {
        const float inv_hw = 1.f / (float)hw;
        // 512-thread CTAs run at 64 registers (two CTAs per SM): the attention values stay in shared memory there
        constexpr bool A_REGS = NT == 256;
        float a[A_REGS ? R : 1][V], dgs[V];
#pragma unroll
        for (int k = 0; k < V; ++k) dgs[k] = dg[m.co + k] * inv_hw;
        if (A_REGS) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < V; ++k) a[r][k] = att[(m.kk * R + r) * c + m.cc + k];
        }
        const float* arow = att + m.kk * R * c + m.cc;
        __nv_bfloat16* dup = (__nv_bfloat16*)du.ptr + ((size_t)n * hw + pbeg) * du.cstride + du.coff + (size_t)m.kk * R * c + m.cc;
        constexpr int UN = NT == 512 ? (R == 1 ? 4 : R == 2 ? 2 : 1) : (R == 1 ? 8 : R == 2 ? 4 : 2);
        for (int it = ((m.nit - 1) / UN) * UN; it >= 0; it -= UN) {
            uint4 q[UN][R], gq[UN];
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int px = m.lane_p + (it + i) * m.pl;
                if (px < npx) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        q[i][r] = cache_u ? cache_u[((size_t)(it + i) * R + r) * NT + tid]
                                          : __ldcs(reinterpret_cast<const uint4*>(up + (size_t)px * u.cstride + r * c));
                    gq[i] = cache_u ? cache_d[(size_t)(it + i) * NT + tid] : __ldcs(reinterpret_cast<const uint4*>(dp + (size_t)px * dv.cstride));
                }
            }
#pragma unroll
            for (int i = 0; i < UN; ++i) {
                const int px = m.lane_p + (it + i) * m.pl;
                if (px < npx) {
                    float gv[V];
                    unpack8(gq[i], gv);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float x[V], o[V];
                        unpack8(q[i][r], x);
#pragma unroll
                        for (int k = 0; k < V; ++k) o[k] = fmaf(gv[k], A_REGS ? a[r][k] : arow[r * c + k], dgs[k]) * act_grad_from_out(p.act, x[k]);
                        __stcs(reinterpret_cast<uint4*>(dup + (size_t)px * du.cstride + r * c), pack8(o));
                    }
                }
            }
        }
    }
    stamp(fc.trace, 9); stamp(fc.trace, 15);
}

struct ClusterPlan { int cs, nt, cache; size_t smem; FcPlan fc; };

int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

// bwd = 0: forward kernel, 1: backward kernel.  The shared-memory carve-up here must match the kernels'.
bool make_plan(const tbi_splitatt* p, int bwd, ClusterPlan* pl) {
    const int K = p->kpaths, R = p->radix, c = p->c, c2 = c / 2;
    const int hw = p->h * p->w, KC = K * c, KC2 = K * c2, C = K * R * c, cvv = KC / V;
    int cs = env_int("TBI_SA_CS", 0);
    if (cs != 1 && cs != 2 && cs != 4 && cs != 8 && cs != 16) {
        cs = 8;
        while (cs > 1 && (hw + cs - 1) / cs < 32) cs >>= 1;
    }
    const int npx = (hw + cs - 1) / cs;
    int nt = env_int("TBI_SA_NT", 0);
    // backward: 512-thread CTAs (64 registers, attention values in shared memory) measured slower than 256-thread ones
    // (66.1 / 42.5 / 36.1 us against 61.4 / 40.9 / 30.4 us on the three config-2 shapes)
    if (nt != 256 && nt != 512) nt = (!bwd && (long long)npx * cvv >= 512 * 4) ? 512 : 256;
    if (cvv < 1 || cvv > nt || nt % cvv != 0) return false;
    const int nit = (npx + nt / cvv - 1) / (nt / cvv);
    // FC plan: small layers redundantly per CTA, large ones sliced over the cluster.  Weight floats each step reads over all K:
    const size_t w_a = bwd ? (size_t)K * (R + 1) * c2 * c : (size_t)K * c * c2;
    const size_t w_b = bwd ? (size_t)K * c * c2 : (size_t)K * R * c2 * c;
    const int out_a = KC2, out_b = bwd ? KC : C;
    const int slice_mode = env_int("TBI_SA_SLICE", 0);      // 0 auto, 1 never slice, 2 always slice (tests)
    // slicing costs remote stores + a cluster barrier on the critical path, redundant computation costs L2 reads of the whole
    // layer by every CTA (measured at c = 128, dense2 = 64 KB: sliced 20.1 us, redundant 21.5 us per forward launch)
    const size_t small = (size_t)env_int("TBI_SA_SMALL_KB", 32) * 1024 / sizeof(float);
    pl->fc.s1 = (cs > 1 && (slice_mode == 2 || (slice_mode == 0 && w_a > small))) ? cs : 1;
    pl->fc.s2 = (cs > 1 && (slice_mode == 2 || (slice_mode == 0 && w_b > small))) ? cs : 1;
    const int n1 = (out_a + pl->fc.s1 - 1) / pl->fc.s1, n2 = (out_b + pl->fc.s2 - 1) / pl->fc.s2;
    size_t staged_floats = bwd ? (size_t)(R + 1) * c * (n1 | 1) + 4 * (size_t)n1 + (size_t)c2 * (n2 | 1)
                               : (size_t)n1 * (c + 5) + (size_t)n2 * (c2 + 1);
    staged_floats = (staged_floats + 3) & ~(size_t)3;
    // shared memory is not free: the CTAs of a launch start over a window that grows with their shared-memory size (measured
    // with scratch/trace_splitatt.py: 0.6 us at 12 KB, 2.3 us at 24 KB, 5.3 us at 61 KB per CTA), so the forward kernel stages
    // only small parameter sets; the backward matvecs need the transposed staged copies to read shared memory conflict-free
    const size_t stage_cap = (size_t)env_int("TBI_SA_STAGE_KB", bwd ? 64 : 16) * 1024;
    pl->fc.staged = (env_int("TBI_SA_STAGE", 1) != 0 && staged_floats * sizeof(float) <= stage_cap) ? 1 : 0;
    const size_t fc_floats = bwd ? (size_t)(cs + 2) * C + 2 * KC + 2 * KC2 : (size_t)(cs + 1) * C + KC + KC2;
    const size_t base = sizeof(float) * (fc_floats + (size_t)nt * V + (pl->fc.staged ? staged_floats : 0));
    const size_t cache_bytes = (size_t)nit * (bwd ? R + 1 : R) * nt * 16;
    const size_t cache_cap = (size_t)env_int("TBI_SA_CACHE_KB", 100) * 1024;
    pl->cs = cs; pl->nt = nt; pl->fc.trace = g_sa_trace_host;
    pl->cache = (cache_cap > 0 && base + cache_bytes <= cache_cap) ? 1 : 0;
    pl->smem = base + (pl->cache ? cache_bytes : 0);
    return pl->smem <= 200 * 1024;
}

template <typename... Args>
int launch_cluster(void (*kernel)(Args..., FcPlan), ClusterPlan& pl, int n_images, cudaStream_t s, const char* what, Args... args, FcPlan* fc_arg) {
    cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && pl.cs > 8) e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "%s attributes: %s", what, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_images * pl.cs)); cfg.blockDim = dim3((unsigned)pl.nt);
    cfg.dynamicSmemBytes = pl.smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)pl.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, args..., *fc_arg);
    if (e != cudaSuccess) return tbi_set_error(TBI_ERR_CUDA, "%s launch (cluster %d, %d threads, %zu B smem): %s", what, pl.cs, pl.nt, pl.smem, cudaGetErrorString(e));
    return 1;
}

bool aligned16(const tbi_view* v) { return v->cstride % V == 0 && v->coff % V == 0 && ((uintptr_t)v->ptr & 15) == 0; }

}  // namespace

extern "C" int tbi_debug_set_splitatt_trace(void* dev_buf) { g_sa_trace_host = (unsigned long long*)dev_buf; return 0; }

// returns 1 if launched, 0 if the fused path does not apply (caller falls back to the multi-kernel path), <0 on error
int tbi_splitatt_fwd_fused(const tbi_splitatt* p, const tbi_view* u, const tbi_view* v, cudaStream_t s) {
    if (p->dtype != TBI_BF16) return 0;
    if (getenv("TBI_NO_FUSED_SPLITATT") != nullptr) return 0;
    if (p->c % V != 0 || !aligned16(u) || !aligned16(v)) return 0;
    const int R = p->radix;
    ClusterPlan pl;
    if (!make_plan(p, 0, &pl)) return 0;
#define TBI_SA_FWD(RR, NN) launch_cluster<tbi_splitatt, tbi_view, tbi_view, int>(splitatt_fwd_cluster_kernel<RR, NN>, pl, p->n, s, "splitatt fused fwd", *p, *u, *v, pl.cache, &pl.fc)
    if (pl.nt == 512) {
        switch (R) { case 1: return TBI_SA_FWD(1, 512); case 2: return TBI_SA_FWD(2, 512); case 3: return TBI_SA_FWD(3, 512); case 4: return TBI_SA_FWD(4, 512); }
    } else {
        switch (R) { case 1: return TBI_SA_FWD(1, 256); case 2: return TBI_SA_FWD(2, 256); case 3: return TBI_SA_FWD(3, 256); case 4: return TBI_SA_FWD(4, 256); }
    }
#undef TBI_SA_FWD
    return 0;
}

int tbi_splitatt_bwd_fused(const tbi_splitatt* p, const tbi_view* u, const tbi_view* dv, const tbi_view* du, float* scratch, cudaStream_t s) {
    if (p->dtype != TBI_BF16) return 0;
    if (getenv("TBI_NO_FUSED_SPLITATT") != nullptr) return 0;
    if (p->c % V != 0 || !aligned16(u) || !aligned16(dv) || !aligned16(du)) return 0;
    const int R = p->radix;
    ClusterPlan pl;
    if (!make_plan(p, 1, &pl)) return 0;
#define TBI_SA_BWD(RR, NN) launch_cluster<tbi_splitatt, tbi_view, tbi_view, tbi_view, float*, int>(splitatt_bwd_cluster_kernel<RR, NN>, pl, p->n, s, "splitatt fused bwd", *p, *u, *dv, *du, scratch, pl.cache, &pl.fc)
    if (pl.nt == 512) {
        switch (R) { case 1: return TBI_SA_BWD(1, 512); case 2: return TBI_SA_BWD(2, 512); case 3: return TBI_SA_BWD(3, 512); case 4: return TBI_SA_BWD(4, 512); }
    } else {
        switch (R) { case 1: return TBI_SA_BWD(1, 256); case 2: return TBI_SA_BWD(2, 256); case 3: return TBI_SA_BWD(3, 256); case 4: return TBI_SA_BWD(4, 256); }
    }
#undef TBI_SA_BWD
    return 0;
}
