// Probe: cost per tcgen05.mma (M=128, K=16, bf16) as a function of N and of the number of independent TMEM
// accumulators the issue stream round-robins over.  Operands: garbage in smem (values irrelevant).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../ultrasound_modeling_b200/csrc/tc_common.cuh"

template <int N, int NACC>
__global__ void rate_kernel(long long* out, int iters, int layout, int sbo_a, int start_a, int kstep) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { tc::mbar_init(&done, 1); tc::fence_barrier_init(); }
    if (warp == 1) tc::tmem_alloc<512>(&slot);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
        const uint32_t a_addr = tc::smem_u32(smem), b_addr = tc::smem_u32(smem + 16384);
        const uint64_t da = tc::make_smem_desc(a_addr + start_a, 16, sbo_a, layout), db = tc::make_smem_desc(b_addr, 16, layout == 2 ? 1024 : layout == 4 ? 512 : 256, layout);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t acc = tmem + (uint32_t)(i & (NACC - 1)) * N;
            tc::umma_bf16(acc, da + 2 * (i & (kstep - 1)), db + 2 * (i & (kstep - 1)), idesc, i >= NACC ? 1u : 0u);
        }
        const long long t1 = clock64();
        tc::umma_commit(&done);
        while (!tc::mbar_try_wait(&done, 0)) {}
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

template <int N, int nacc> void run(long long* d, int iters, int layout = 2, int sbo_a = 1024, int start_a = 0, int kstep = 4) {
    cudaFuncSetAttribute(rate_kernel<N, nacc>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    long long h[2];
    rate_kernel<N, nacc><<<1, 64, 64 * 1024>>>(d, iters, layout, sbo_a, start_a, kstep); cudaDeviceSynchronize();           // warm
    rate_kernel<N, nacc><<<1, 64, 64 * 1024>>>(d, iters, layout, sbo_a, start_a, kstep);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d nacc=%d: %s\n", N, nacc, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("layout=%d sbo_a=%4d start_a=%4d | N=%3d  accumulators=%d  MMAs=%d : issue %.1f cyc/MMA, issue+complete %.1f cyc/MMA  (math floor %d cyc)\n", layout, sbo_a, start_a, N, nacc, iters,
           (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256);
    fflush(stdout);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    run<32, 1>(d, 256);                                   // SW128 reference
    run<32, 1>(d, 256, 2, 1280, 0, 4);                    // SW128, halo pitch 10 rows
    run<32, 1>(d, 256, 2, 1280, 11 * 128, 4);             // + unaligned start row
    run<32, 1>(d, 256, 4, 512, 0, 2);                     // SW64 canonical
    run<32, 1>(d, 256, 4, 640, 0, 2);                     // SW64, halo pitch 10 rows (stem conv 32 ch)
    run<32, 1>(d, 256, 4, 640, 11 * 64, 2);               // + unaligned start row
    run<32, 1>(d, 256, 6, 256, 0, 1);                     // SW32 canonical
    run<32, 1>(d, 256, 6, 320, 11 * 32, 1);               // SW32 halo
    run<128, 1>(d, 256, 2, 1280, 11 * 128, 4);            // big N with halo addressing
    return 0;
}
