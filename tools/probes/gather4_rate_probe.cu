// How fast does TMA tile::gather4 move 128-byte row slices (the kept-token gather of the attention loaders)?
// Every SM runs one CTA; W warps each elect one thread that keeps BATCH gather4 loads (4 rows x 128 B) in flight on its own
// mbarrier and shared-memory region, over random rows of a [R, 2304] bf16 matrix (the qkv layout).  Prints bytes/clk/SM.
// Compared with: the same bytes as cp.async 16-byte row gathers by W full warps, and as dense TMA boxes of 64 rows.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rajni_vit_b200/csrc gather4_rate_probe.cu -o gather4_rate_probe -lcuda
#include <cstdio>
#include <cuda.h>
#include <vector>
#include "common.cuh"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return 0; } }
using namespace rajni;

constexpr int kMaxWarps = 8, kBatch = 8, kIters = 64;          // per warp: 8 gathers = 4 KB per batch

__global__ void __launch_bounds__(kMaxWarps * 32, 1)
gather_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* base, const int* rows, int n_rows, int warps, int mode, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kMaxWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < kMaxWarps; ++i) mbar_init(&bars[i], mode == 1 ? 32 : 1); mbar_fence_init(); }
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        uint8_t* region = smem + warp * (kBatch * 512);
        const int* my = rows + ((blockIdx.x * kMaxWarps + warp) * kIters * kBatch * 4) % (n_rows - kIters * kBatch * 4);
        for (int it = 0; it < kIters; ++it) {
            if (mode == 0) {                       // TMA gather4, one thread per warp
                if (lane == 0) {
                    mbar_expect_tx(&bars[warp], kBatch * 512);
                    for (int g = 0; g < kBatch; ++g) {
                        const int* r = my + (it * kBatch + g) * 4;
                        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                                     :: "r"(smem_u32(region + g * 512)), "l"(&tmap), "r"(smem_u32(&bars[warp])), "r"(64 * (warp % 12)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
                    }
                }
                mbar_wait(&bars[warp], it & 1);
            } else if (mode == 1) {                // cp.async 16 B per thread: 8 lanes per row, 4 rows per instruction
                for (int g = 0; g < kBatch; ++g) {
                    const int row = my[(it * kBatch + g) * 4 + (lane >> 3)];
                    const __nv_bfloat16* src = base + (long long)row * 2304 + 64 * (warp % 12) + (lane & 7) * 8;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(region + g * 512 + lane * 16)), "l"(src) : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(&bars[warp])) : "memory");
                mbar_wait(&bars[warp], it & 1);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    const int R = 50432, C3 = 2304;
    __nv_bfloat16* d; cudaMalloc(&d, (size_t)R * C3 * 2); cudaMemset(d, 0, (size_t)R * C3 * 2);
    std::vector<int> h(1 << 20);
    // kept-token-like pattern: ascending rows with ~12 % skipped, wrapping inside the matrix
    unsigned s = 12345; int r = 0;
    for (auto& v : h) { s = s * 1664525u + 1013904223u; r += 1 + ((s >> 24) < 31); if (r >= R) r = 0; v = r; }
    int* rows; cudaMalloc(&rows, h.size() * 4); cudaMemcpy(rows, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    long long* cyc; cudaMallocManaged(&cyc, 148 * 8);
    cuInit(0);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)C3, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C3 * 2};
    cuuint32_t box[2] = {64u, 1u}; cuuint32_t es[2] = {1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode rc=%d\n", (int)rc); return 1; }
    const int smem = kMaxWarps * kBatch * 512;
    cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 2, 4, 8}) {
            for (int rep = 0; rep < 2; ++rep) {
                gather_kernel<<<148, kMaxWarps * 32, smem>>>(m, d, rows, (int)h.size(), warps, mode, cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
            }
            long long mx = 0; for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
            const double bytes = (double)warps * kIters * kBatch * 512;
            printf("%-22s %d issuing warp(s), %d x 512 B in flight each: %7.2f bytes/clk/SM  (%lld cycles)\n", mode == 0 ? "TMA gather4" : "cp.async 16 B rows", warps, kBatch, bytes / mx, mx);
        }
    return 0;
}
