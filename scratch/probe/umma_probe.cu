// Probe: does a K-major SW128 UMMA A-operand work when its start address is an arbitrary 128-byte row of a
// TMA-written tile (rows written with the address-based 128B swizzle), with SBO != 1024 ?
// A tile in smem: ROWS x 64 bf16 (128 B rows), written by TMA with SWIZZLE_128B from a [ROWS][64] global matrix G.
// Operand rows: m -> smem row r0 + (m/8)*pitch + (m%8)   (pitch in rows), i.e. SBO = pitch*128.
// B = 64x64 identity (K-major, N=64) so D[m][n] = G[row(m)][n] for k-chunk columns.
// Prints the max error for each (r0, pitch, base_offset mode).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../ultrasound_modeling_b200/csrc/tc_common.cuh"

constexpr int ROWS = 256;

__global__ void probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float* out, int r0, int pitch, int bo_mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_s = smem;                       // ROWS*128 = 32768
    uint8_t* b_s = smem + 32768;               // 64*128 = 8192
    uint64_t* bar = (uint64_t*)(smem + 40960);
    uint64_t* done = bar + 1;
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(done, 1); tc::fence_barrier_init(); }
    if (warp == 1) tc::tmem_alloc<64>(slot);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar, 32768 + 8192);
        tc::tma_load_2d(a_s, &ta, bar, 0, 0);
        tc::tma_load_2d(b_s, &tb, bar, 0, 0);
        { long long t0 = clock64(); while (!tc::mbar_try_wait(bar, 0)) { if (clock64() - t0 > 2000000000LL) { printf("tma wait timeout\n"); __trap(); } } }
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc_bf16(128, 64, 0, 0);
        const uint32_t a_addr = tc::smem_u32(a_s) + r0 * 128, b_addr = tc::smem_u32(b_s);
        for (int k = 0; k < 4; ++k) {
            uint64_t da = tc::make_smem_desc(a_addr + k * 32, 16, pitch * 128, 2u);
            if (bo_mode == 1) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
            const uint64_t db = tc::make_smem_desc(b_addr + k * 32, 16, 1024, 2u);
            tc::umma_bf16(tmem, da, db, idesc, k != 0);
        }
        tc::umma_commit(done);
    }
    __syncwarp();
    if (warp >= 2) {
        { long long t0 = clock64(); while (!tc::mbar_try_wait(done, 0)) { if (clock64() - t0 > 2000000000LL) { printf("mma wait timeout\n"); __trap(); } } }
        tc::tc_fence_after();
        const int q = warp & 3;
        uint32_t r[32];
        for (int c = 0; c < 64; c += 32) {
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, r);
            tc::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[(q * 32 + lane) * 64 + c + j] = __uint_as_float(r[j]);
        }
    }
    tc::tc_fence_before(); __syncthreads();
    if (warp == 1) tc::tmem_dealloc<64>(tmem);
}

tbi_encode_tiled_fn tbi_get_encode_tiled() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return (tbi_encode_tiled_fn)p;
}

int main() {
    std::vector<__nv_bfloat16> G(ROWS * 64), I(64 * 64);
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) G[r * 64 + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251) - 125.f);
    for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) I[r * 64 + c] = __float2bfloat16(r == c ? 1.f : 0.f);
    __nv_bfloat16 *dG, *dI; float* dO;
    cudaMalloc(&dG, G.size() * 2); cudaMalloc(&dI, I.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dG, G.data(), G.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dI, I.data(), I.size() * 2, cudaMemcpyHostToDevice);
    auto enc = tbi_get_encode_tiled();
    CUtensorMap ta, tb;
    cuuint64_t gd[2] = {64, ROWS}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
    enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dG, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gd2[2] = {64, 64}; cuuint32_t bx2[2] = {64, 64};
    enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dI, gd2, gs, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> O(128 * 64);
    const int cases[][2] = {{0, 8}, {8, 8}, {1, 8}, {3, 8}, {0, 10}, {11, 10}, {1, 16}, {2, 16}, {7, 16}, {5, 12}, {40, 12}};
    for (auto& cs : cases) for (int bo = 0; bo < 2; ++bo) {
        cudaMemset(dO, 0, 128 * 64 * 4);
        probe_kernel<<<1, 192, 64 * 1024>>>(ta, tb, dO, cs[0], cs[1], bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("r0=%d pitch=%d bo=%d: CUDA error %s\n", cs[0], cs[1], bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0; int bad = 0;
        for (int m = 0; m < 128; ++m) { const int row = cs[0] + (m / 8) * cs[1] + (m % 8);
            for (int c = 0; c < 64; ++c) { double d = fabs(O[m * 64 + c] - __bfloat162float(G[row * 64 + c])); if (d > err) err = d; if (d > 0.5) ++bad; } }
        printf("r0=%2d pitch=%2d base_offset_mode=%d : max err %.1f  bad %d/8192  %s\n", cs[0], cs[1], bo, err, bad, err < 0.5 ? "OK" : "WRONG"); fflush(stdout);
    }
    return 0;
}
