// Does cp.async.bulk.tensor.2d ... tile::gather4 work with a cuTensorMapEncodeTiled map, and with which box?
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rajni_vit_b200/csrc gather4_probe.cu -o gather4_probe -lcuda
#include <cstdio>
#include <cuda.h>
#include <vector>
#include "common.cuh"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return 0; } }
using namespace rajni;

__global__ void k(const __grid_constant__ CUtensorMap tmap, int col, int r0, int r1, int r2, int r3, uint16_t* out, int* status) {
    __shared__ __align__(1024) uint16_t buf[8 * 64];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) buf[i] = 0xffff;
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 4 * 128);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     :: "r"(smem_u32(buf)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        // bounded wait so a wrong byte count cannot hang the box
        int ok = 0;
        for (int it = 0; it < 2000000 && !ok; ++it) ok = mbar_test(&bar, 0);
        *status = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = buf[i];
}

int main() {
    const int R = 1024, C3 = 2304;
    std::vector<uint16_t> h((size_t)R * C3);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C3; ++c) h[(size_t)r * C3 + c] = (uint16_t)((r << 6) ^ (c & 63) ^ ((c >> 6) << 12));
    uint16_t *d, *out; int* status;
    cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMallocManaged(&out, 1024); cudaMallocManaged(&status, 4);
    for (int boxrows : {1, 4}) {
        CUtensorMap m;
        cuuint64_t dims[2] = {(cuuint64_t)C3, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C3 * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
        cuInit(0);
        CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %d: encode rc=%d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int col = 64 * 5, rows[4] = {5, 17, 100, 3};
        *status = -1;
        k<<<1, 128>>>(m, col, rows[0], rows[1], rows[2], rows[3], out, status);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  run: %s, barrier completed=%d\n", cudaGetErrorString(e), *status);
        if (e != cudaSuccess) return 1;
        // decode: smem row i (128 B), 16-byte chunk c holds (with 128B swizzle) logical chunk c ^ (i & 7)
        for (int i = 0; i < 4; ++i) {
            int okc = 0;
            for (int c = 0; c < 8; ++c)
                for (int e2 = 0; e2 < 8; ++e2) {
                    const int lc = c ^ (i & 7);
                    const uint16_t want = h[(size_t)rows[i] * C3 + col + lc * 8 + e2];
                    okc += out[i * 64 + c * 8 + e2] == want;
                }
            printf("  smem row %d: %d/64 elements match global row %d (swizzled)   first=%04x\n", i, okc, rows[i], out[i * 64]);
        }
    }
    return 0;
}
