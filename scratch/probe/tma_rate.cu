// Probe 3: TMA tensor-load throughput per SM as a function of the box shape (inner row bytes x rows), 2 CTAs/SM, NST loads in flight each.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include "../../ultrasound_modeling_b200/csrc/tc_common.cuh"
int tbi_make_tmap_bf16(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

constexpr int NST = 6;
__global__ void fill(uint32_t* p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)(i * 2654435761u) ^ (uint32_t)(i >> 7); }
__global__ void __launch_bounds__(64) tma_kernel(const __grid_constant__ CUtensorMap map, long long* out, int loads, int stage_bytes, int tx, int tiles_x, int tiles_y, int n) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[NST], empty[NST];
    if (threadIdx.x == 0) { for (int s = 0; s < NST; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); } tc::fence_barrier_init(); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = tiles_x * tiles_y * n;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < loads; ++i) {
            const int s = i % NST;
            tc::mbar_wait(&empty[s], ((i / NST) & 1) ^ 1);
            int t = (blockIdx.x + i * gridDim.x) % total;
            const int x0 = (t % tiles_x) * 8; t /= tiles_x; const int y0 = (t % tiles_y) * 16; const int nn = t / tiles_y;
            tc::mbar_expect_tx(&full[s], tx);
            tc::tma_load_4d(smem + (size_t)s * stage_bytes, &map, &full[s], 0, x0 - 1, y0 - 1, nn);
        }
    } else if (warp == 1 && lane == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < loads; ++i) {
            const int s = i % NST;
            tc::mbar_wait(&full[s], (i / NST) & 1);
            tc::mbar_arrive(&empty[s]);
        }
        if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    const int N = 64, H = 256, W = 256;
    void* buf; cudaMalloc(&buf, (size_t)N * H * W * 128 * 2);
    fill<<<4096, 256>>>((uint32_t*)buf, (size_t)N * H * W * 128 / 2); cudaDeviceSynchronize();
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    struct Case { int c, bw, bh; } cases[] = {{16, 10, 18}, {32, 10, 18}, {64, 10, 18}, {32, 8, 16}, {64, 8, 16}, {64, 10, 9}, {32, 10, 36}};
    for (auto cs : cases) {
        const int hh = H;      // keep the tensor <= 1 GB-ish but far larger than L2
        CUtensorMap m;
        uint64_t dims[4] = {(uint64_t)cs.c, (uint64_t)W, (uint64_t)hh, (uint64_t)N};
        uint64_t strides[3] = {(uint64_t)cs.c * 2, (uint64_t)cs.c * 2 * W, (uint64_t)cs.c * 2 * W * hh};
        uint32_t box[4] = {(uint32_t)cs.c, (uint32_t)cs.bw, (uint32_t)cs.bh, 1};
        if (tbi_make_tmap_bf16(&m, buf, 4, dims, strides, box, cs.c * 2 > 128 ? 128 : cs.c * 2)) { printf("tmap failed\n"); return 1; }
        const int tx = cs.c * 2 * cs.bw * cs.bh, stage = (tx + 1023) & ~1023;
        const int loads = 400;
        for (int grid : {148, 296}) {
            for (int rep = 0; rep < 2; ++rep) tma_kernel<<<grid, 64, NST * stage + 1024>>>(m, d, loads, stage, tx, W / 8, hh / 16, N);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double cyc = (double)h / loads;
            printf("C=%3d (%3d B rows) box %2dx%2d = %3d rows, %5d B | %d CTA/SM: %.0f cycles/load/CTA -> %.1f B/cycle/SM, %.1f cycles/row/SM\n", cs.c, cs.c * 2, cs.bw, cs.bh, cs.bw * cs.bh, tx,
                   grid / 148, cyc, tx / cyc * (grid / 148), cyc / (cs.bw * cs.bh) / (grid / 148));
            fflush(stdout);
        }
    }
    return 0;
}
