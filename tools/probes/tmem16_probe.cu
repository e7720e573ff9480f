// Which (lane, column) does each register of tcgen05.ld.16x256b.x8 / tcgen05.st.16x128b.x8 touch, and may the
// 16-lane window start at lane 16 of a warp's quadrant?  Ground truth for the row-split softmax of attention_tc.cu.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rajni_vit_b200/csrc tmem16_probe.cu -o tmem16_probe
#include <cstdio>
#include "common.cuh"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return 0; } }
using namespace rajni;

__device__ __forceinline__ void ld16x256_x8(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16x128_x8(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr),
           "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

__global__ void __launch_bounds__(256) probe(uint32_t* out_ld, uint32_t* out_st) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, hsel = warp >> 2;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    // pattern: TMEM[lane L][col c] = L*1000 + c, columns 0..127, written with the well-understood 32x32b shape
    if (hsel == 0) {
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            for (int j = 0; j < 32; ++j) v[j] = (q * 32 + lane) * 1000 + c0 + j;
            tmem_st32(tm + ((uint32_t)(q * 32) << 16) + c0, v);
        }
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    {   // 16x256b.x8 load of columns 0..63 from the 16-lane window starting at lane q*32 + hsel*16
        uint32_t v[32];
        ld16x256_x8(tm + ((uint32_t)(q * 32 + hsel * 16) << 16), v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out_ld[(warp * 32 + lane) * 32 + j] = v[j];
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    {   // 16x128b.x8 store into columns 256.. : value encodes (warp, lane, reg)
        uint32_t v[16];
        for (int j = 0; j < 16; ++j) v[j] = warp * 100000 + lane * 100 + j;
        st16x128_x8(tm + ((uint32_t)(q * 32 + hsel * 16) << 16) + 256, v);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (hsel == 0) {   // read columns 256..287 back with 32x32b
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + 256, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out_st[(q * 32 + lane) * 32 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    uint32_t *ld, *st;
    cudaMallocManaged(&ld, 256 * 32 * 4); cudaMallocManaged(&st, 128 * 32 * 4);
    probe<<<1, 256>>>(ld, st);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    // expected load layout: reg 4g+{0,1} = (row t/4, cols 8g+2(t%4)+{0,1}); reg 4g+{2,3} = (row t/4+8, same cols)
    int bad = 0;
    for (int w = 0; w < 8; ++w) for (int t = 0; t < 32; ++t) for (int j = 0; j < 32; ++j) {
        int g = j / 4, r = j % 4;
        int lane = (w & 3) * 32 + (w >> 2) * 16 + t / 4 + (r >= 2 ? 8 : 0), col = 8 * g + 2 * (t % 4) + (r & 1);
        uint32_t got = ld[(w * 32 + t) * 32 + j];
        if (got != (uint32_t)(lane * 1000 + col)) { if (bad < 12) printf("LD mismatch warp %d thr %d reg %d: got lane %u col %u, expected lane %d col %d\n", w, t, j, got / 1000, got % 1000, lane, col); ++bad; }
    }
    printf("16x256b.x8 load: %d mismatches against the mma-fragment layout (lane window may start at +16: %s)\n", bad, bad ? "?" : "yes");
    // expected store layout: reg 2g+{0} -> (row t/4, col 4g + t%4), reg 2g+1 -> (row t/4+8, col 4g+t%4)
    bad = 0;
    for (int L = 0; L < 128; ++L) for (int c = 0; c < 32; ++c) {
        uint32_t got = st[L * 32 + c];
        int q = L / 32, in = L % 32, hsel = in / 16, r = in % 16;
        int w = hsel * 4 + q, t = (r % 8) * 4 + c % 4, j = 2 * (c / 4) + (r >= 8 ? 1 : 0);
        uint32_t exp = w * 100000 + t * 100 + j;
        if (got != exp) { if (bad < 12) printf("ST mismatch lane %d col %d: got warp %u thr %u reg %u, expected warp %d thr %d reg %d\n", L, c, got / 100000, (got / 100) % 1000, got % 100, w, t, j); ++bad; }
    }
    printf("16x128b.x8 store: %d mismatches\n", bad);
    return 0;
}
