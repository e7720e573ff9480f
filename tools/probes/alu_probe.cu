// Per-SM issue rates of MUFU.EX2, FFMA, FFMA2 and a software exp2 (Cody-Waite + cubic, FMA pipe only) as a function of
// the number of resident warps.  Clean loops (template-selected body, 8 independent chains per thread).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rajni_vit_b200/csrc alu_probe.cu -o alu_probe
#include <cstdio>
#include "common.cuh"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return 0; } }
using namespace rajni;

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^x for x <= 0 on the FMA/ALU pipes: x = n + f, f in [-1, 0]: 2^f by a cubic (rel err ~1e-4), exponent patched in by integer add
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -120.f);
    const float n = floorf(x);                 // FRND (or add-magic)
    const float f = x - n;                     // [0,1)
    float p = 0.0555054f;
    p = fmaf(p, f, 0.2402265f);
    p = fmaf(p, f, 0.6931472f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + ((int)n << 23));
}

template <int WHICH>
__global__ void __launch_bounds__(1024) alu_kernel(int reps, long long* out, float* sink) {
    float x[8];
    uint64_t y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = -(threadIdx.x * 1e-3f + i); y[i] = f2pack(x[i], x[i] + 0.5f); }
    const uint64_t c2 = f2pack(0.999f, 1.001f), d2 = f2pack(-1e-3f, -2e-3f);
    float acc0 = 0.f;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (WHICH == 0) x[i] = -ex2(x[i]);
            else if (WHICH == 1) x[i] = fmaf(x[i], 0.999f, -1e-3f);
            else if (WHICH == 2) y[i] = fma2(y[i], c2, d2);
            else if (WHICH == 3) x[i] = -ex2_poly(x[i]);
            else if (WHICH == 4) { x[i] = -ex2(x[i]); y[i] = fma2(y[i], c2, d2); y[i] = fma2(y[i], c2, d2); }     // 1 MUFU : 2 FFMA2
            else if (WHICH == 5) x[i] = __uint_as_float(float2_to_bf16x2(x[i], x[(i + 1) & 7]));        // F2FP.BF16 pack alone
            else if (WHICH == 6 || WHICH == 7 || WHICH == 8) {
                // the softmax exp pass per element pair: 2 FFMA, 2 MUFU.EX2, 2 FADD, 1 pack
                const float e0 = ex2(fmaf(x[i], 0.18f, -0.05f)), e1 = ex2(fmaf(x[(i + 1) & 7], 0.18f, -0.05f));
                acc0 += e0 + e1;
                uint32_t pk;
                if (WHICH == 6) pk = float2_to_bf16x2(e0, e1);
                else if (WHICH == 7) pk = __byte_perm(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632);   // round half up, integer pipe
                else pk = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);                                      // truncate
                x[i] = -__uint_as_float(pk & 0x3fff3fffu);
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    float acc = acc0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; f2unpack(y[i], a, b); acc += x[i] + a + b; }
    sink[threadIdx.x] = acc;
}

template <int WHICH> void run(const char* name, long long* o, float* sk) {
    for (int warps : {4, 8, 12, 16, 32}) {
        alu_kernel<WHICH><<<1, warps * 32>>>(512, o, sk);
        cudaDeviceSynchronize();
        alu_kernel<WHICH><<<1, warps * 32>>>(512, o, sk);
        cudaDeviceSynchronize();
        printf("%-22s %2d warps (%d/SMSP): %6.2f lane-results/clk/SM, %5.1f clk between a warp's consecutive ops\n", name, warps, warps / 4,
               512.0 * 8 * warps * 32 / (double)o[0], (double)o[0] / (512.0 * 8));
    }
}

int main() {
    long long* o; float* sk;
    cudaMallocManaged(&o, 64); cudaMallocManaged(&sk, 4096 * 4);
    run<0>("MUFU.EX2", o, sk);
    run<1>("FFMA", o, sk);
    run<2>("FFMA2 (2 fma each)", o, sk);
    run<3>("exp2 poly (FMA pipe)", o, sk);
    run<4>("MUFU + 2 FFMA2 mix", o, sk);
    run<5>("F2FP.BF16 pack", o, sk);
    run<6>("exp pair, cvt pack", o, sk);
    run<7>("exp pair, int round", o, sk);
    run<8>("exp pair, truncate", o, sk);
    return 0;
}
