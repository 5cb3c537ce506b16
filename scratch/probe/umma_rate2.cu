// Probe 2: why do the 18 small MMAs of a stem-conv tile (M=128, N=32, K=16, SW64, halo pitch) take ~4000 cycles in the
// halo kernel when a free-running issue loop dispatches one every ~50 cycles?  Emulates the tile protocol step by step:
//   flags bit0: epilogue warps drain the accumulator (tcgen05.ld) and the MMA thread waits for the buffer to be free
//   flags bit1: epilogue writes 64 B per thread to global
//   flags bit2: a producer warp streams 11.5 KB per tile from global into smem with bulk copies (mbarrier paced)
//   grid: 1 CTA, or 2 CTAs per SM on every SM
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../ultrasound_modeling_b200/csrc/tc_common.cuh"

constexpr int MMAS = 18, STAGE = 11520, NST = 4;

template <int N>
__global__ void __launch_bounds__(192) tile_kernel(long long* out, const uint8_t* src, uint8_t* dst, int tiles, int flags) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t t_full[2], t_empty[2], a_full[NST], a_empty[NST];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&t_full[b], 1); tc::mbar_init(&t_empty[b], 4); }
        for (int s = 0; s < NST; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<64>(&slot);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = slot;
    const bool epi = flags & 1, st = flags & 2, prod = flags & 4;
    if (warp == 0 && prod) {
        if (lane == 0) {
            for (int i = 0; i < tiles; ++i) {
                const int s = i % NST; const uint32_t par = (i / NST) & 1;
                tc::mbar_wait(&a_empty[s], par ^ 1u);
                tc::mbar_expect_tx(&a_full[s], STAGE);
                const uint8_t* g = src + ((size_t)(blockIdx.x * tiles + i) % 4096) * STAGE;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(tc::smem_u32(smem + 16384 + s * 12288)), "l"(g), "r"(STAGE), "r"(tc::smem_u32(&a_full[s])) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            tc::fence_proxy_async();
            const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
            const uint64_t db = tc::make_smem_desc(tc::smem_u32(smem), 16, 512, 4);
            const long long t0 = clock64();
            long long issue = 0;
            for (int i = 0; i < tiles; ++i) {
                const uint32_t buf = i & 1;
                if (epi) { tc::mbar_wait(&t_empty[buf], ((i >> 1) & 1u) ^ 1u); tc::tc_fence_after(); }
                const int s = i % NST;
                if (prod) { tc::mbar_wait(&a_full[s], (i / NST) & 1); tc::tc_fence_after(); }
                const uint64_t da = tc::make_smem_desc(tc::smem_u32(smem + 16384 + (prod ? s * 12288 : 0)), 16, 640, 4);
                const long long a = clock64();
                for (int m = 0; m < MMAS; ++m)
                    tc::umma_bf16(tmem + buf * 32, da + 4 * (m >> 1) + 2 * (m & 1), db + 64 * (m >> 1) + 2 * (m & 1), idesc, m ? 1u : 0u);
                issue += clock64() - a;
                if (prod) tc::umma_commit(&a_empty[s]);
                tc::umma_commit(&t_full[buf]);
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = issue; }
        }
    } else if (warp >= 2 && epi) {
        const int q = warp & 3;
        for (int i = 0; i < tiles; ++i) {
            const uint32_t buf = i & 1;
            tc::mbar_wait(&t_full[buf], (i >> 1) & 1u);
            tc::tc_fence_after();
            uint32_t r[32];
            tc::tmem_ld32(tmem + buf * 32 + ((uint32_t)(q * 32) << 16), r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&t_empty[buf]);
            if (st) {
                uint4* o = reinterpret_cast<uint4*>(dst + ((size_t)((blockIdx.x * tiles + i) % 8192) * 128 + (q * 32 + lane)) * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = make_uint4(r[8 * j], r[8 * j + 1] ^ r[8 * j + 2], r[8 * j + 3] ^ r[8 * j + 4], r[8 * j + 5] ^ r[8 * j + 6] ^ r[8 * j + 7]);
            }
        }
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 1) tc::tmem_dealloc<64>(tmem);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    uint8_t *src, *dst; cudaMalloc(&src, (size_t)4096 * STAGE); cudaMalloc(&dst, (size_t)8192 * 128 * 64);
    cudaMemset(src, 0, (size_t)4096 * STAGE);
    cudaFuncSetAttribute(tile_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    const int tiles = 200;
    for (int grid : {1, 148, 296}) for (int flags : {0, 1, 3, 4, 5, 7}) {
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) { tile_kernel<32><<<grid, 192, 80 * 1024>>>(d, src, dst, tiles, flags); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("grid %d flags %d: %s\n", grid, flags, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("grid=%3d epi=%d store=%d producer=%d : %.0f cycles/tile (CTA 0), issue loop %.0f cycles per 18 MMAs\n", grid, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1,
               (double)h[0] / tiles, (double)h[1] / tiles);
        fflush(stdout);
    }
    return 0;
}
