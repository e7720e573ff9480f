// Micro-benchmarks of tcgen05.mma / tcgen05.ld issue rates on one SM (ground truth for the attention design).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rajni_vit_b200/csrc mma_probe.cu -o mma_probe
#include <cstdio>
#include "common.cuh"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return 0; } }
using namespace rajni;

// mode 0: SS, A K-major, B K-major     mode 1: SS, B MN-major     mode 2: TS (A from TMEM), B MN-major   mode 3: TS, B K-major
__global__ void __launch_bounds__(128) mma_kernel(int mode, int N, int reps, long long* out, int nacc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(&slot, 512); if (lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); } }
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, (mode == 1 || mode == 2) ? 1 : 0);
        const uint32_t sa = base, sb = base + 32768;
        // descriptors are built once; the timed loop is MMA issue only
        uint64_t ad[4], bd[4];
        uint32_t dd[4], at[4];
        for (int k = 0; k < 4; ++k) {
            ad[k] = umma_desc_sw128(sa + k * 32, 16, 1024);
            bd[k] = (mode == 1 || mode == 2) ? umma_desc_sw128(sb + k * 2048, 16, 1024) : umma_desc_sw128(sb + k * 32, 16, 1024);
            dd[k] = tm + 256 + (k % nacc) * 64;
            at[k] = tm + k * 8;
        }
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode >= 2) umma_bf16_ts(dd[k], at[k], bd[k], idesc, r >= 4);
                else umma_bf16(dd[k], ad[k], bd[k], idesc, r >= 4);
            }
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// warps x reps LDTM.x32 (4 KB per warp-instruction)
__global__ void __launch_bounds__(256) ldtm_kernel(int reps, int st, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t trow = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t v[32];
        if (st) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = acc + j;
            tmem_st32(trow + (r & 7) * 32, v);
            acc += r;
        } else {
            tmem_ld32(trow + (r & 7) * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
        }
    }
    if (st) tmem_st_wait();
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 256 + threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

// LDTM shapes: 0 = 32x32b.x32 (4 KB), 1 = 16x256b.x8 (4 KB), 2 = 16x256b.x4 (2 KB); `depth` loads in flight per wait
__global__ void __launch_bounds__(512) ldtm2_kernel(int shape, int depth, int reps, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t trow = slot + ((uint32_t)((warp & 3) * 32 + ((warp >> 2) & 1) * 16 * (shape != 0)) << 16) + (warp >> 3) * 256;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t a[32], b[32], c[16], d[16];
        if (shape == 0) { tmem_ld32(trow, a); if (depth > 1) tmem_ld32(trow + 32, b); }
        else if (shape == 1) { tmem_ld16x256_x8(trow, a); if (depth > 1) tmem_ld16x256_x8(trow + 64, b); }
        else { tmem_ld16x256_x4(trow, c); if (depth > 1) tmem_ld16x256_x4(trow + 32, d); }
        tmem_ld_wait();
        if (shape < 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= a[j] ^ (depth > 1 ? b[j] : 0u);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= c[j] ^ (depth > 1 ? d[j] : 0u);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

// MUFU.EX2 / FFMA / FFMA2 issue rates: `warps` warps, 8 independent chains per thread
__global__ void __launch_bounds__(1024) alu_kernel(int which, int reps, long long* out, float* sink) {
    float x[8];
    uint64_t y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = f2pack(x[i], x[i] + 0.5f); }
    const uint64_t c2 = f2pack(0.999f, 1.001f), d2 = f2pack(1e-3f, 2e-3f);
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (which == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            else if (which == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(0.999f), "f"(1e-3f));
            else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(c2), "l"(d2));
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; f2unpack(y[i], a, b); acc += x[i] + a + b; }
    sink[threadIdx.x] = acc;
}

int main() {
    {
        long long* o; float* sk;
        cudaMallocManaged(&o, 64); cudaMallocManaged(&sk, 4096 * 4);
        const char* nm[] = {"MUFU.EX2", "FFMA", "FFMA2"};
        for (int which = 0; which < 3; ++which)
            for (int warps : {4, 8, 16, 32}) {
                alu_kernel<<<1, warps * 32>>>(which, 256, o, sk);
                cudaDeviceSynchronize();
                alu_kernel<<<1, warps * 32>>>(which, 256, o, sk);
                cudaDeviceSynchronize();
                printf("%-8s %2d warps: %.2f lane-ops/clk/SM\n", nm[which], warps, 256.0 * 8 * warps * 32 / (double)o[0]);
            }
    }
    long long* out; uint32_t* sink;
    cudaMallocManaged(&out, 4096); cudaMallocManaged(&sink, 1 << 20);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* names[] = {"SS  K-major B ", "SS  MN-major B", "TS  MN-major B", "TS  K-major B "};
    const int reps = 256;
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {16, 32, 64, 128, 176, 208, 256}) {
            if ((mode == 1 || mode == 2) && N > 64) continue;   // MN-major probe uses a single 64-wide block
            mma_kernel<<<1, 128, 100 * 1024>>>(mode, N, reps, out, 1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
            mma_kernel<<<1, 128, 100 * 1024>>>(mode, N, reps, out, 1);
            cudaDeviceSynchronize();
            printf("%s M=128 N=%3d K=16: issue %.1f cyc/mma, complete %.1f cyc/mma (model floor %.0f)\n", names[mode], N,
                   (double)out[0] / reps, (double)out[1] / reps, 128.0 * N / 256);
        }
    for (int mode : {0, 2})
        for (int nacc : {1, 2, 3, 4}) {
            mma_kernel<<<1, 128, 100 * 1024>>>(mode, 64, reps, out, nacc);
            cudaDeviceSynchronize();
            mma_kernel<<<1, 128, 100 * 1024>>>(mode, 64, reps, out, nacc);
            cudaDeviceSynchronize();
            printf("%s M=128 N= 64 K=16, %d independent accumulators: issue %.1f cyc/mma, complete %.1f cyc/mma\n", names[mode], nacc,
                   (double)out[0] / reps, (double)out[1] / reps);
        }
    {
        const char* shp[] = {"32x32b.x32 (4KB)", "16x256b.x8 (4KB)", "16x256b.x4 (2KB)"};
        const int bytes[] = {4096, 4096, 2048};
        for (int shape = 0; shape < 3; ++shape)
            for (int depth : {1, 2})
                for (int warps : {4, 8, 16}) {
                    ldtm2_kernel<<<1, warps * 32>>>(shape, depth, 512, out, sink);
                    cudaDeviceSynchronize();
                    ldtm2_kernel<<<1, warps * 32>>>(shape, depth, 512, out, sink);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("ldtm2: %s\n", cudaGetErrorString(e)); return 1; }
                    printf("LDTM %s depth %d, %2d warps: %.1f cyc per wait, %.1f B/cyc/SM\n", shp[shape], depth, warps,
                           (double)out[0] / 512, (double)bytes[shape] * depth * warps * 512 / out[0]);
                }
    }
    for (int st = 0; st < 2; ++st)
        for (int warps : {1, 4, 8}) {
            ldtm_kernel<<<1, warps * 32>>>(512, st, out, sink);
            cudaDeviceSynchronize();
            ldtm_kernel<<<1, warps * 32>>>(512, st, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("ldtm: %s\n", cudaGetErrorString(e)); return 1; }
            printf("%s.x32 %d warps: %.1f cyc per warp-instruction, %.1f B/cyc/SM\n", st ? "STTM" : "LDTM", warps,
                   (double)out[0] / 512, 4096.0 * warps * 512 / out[0]);
        }
    return 0;
}
