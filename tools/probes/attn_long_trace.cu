// Event timeline of the long-sequence attention kernel on CTA 0 (clock64 stamps).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRAJNI_ATTN_TRACE -I../../rajni_vit_b200/csrc attn_long_trace.cu ../../rajni_vit_b200/csrc/gemm_tcgen05.o -o attn_long_trace
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rajni_vit_b200/csrc/attention_long.cu"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return cudaGetLastError() == cudaSuccess ? 0 : -2; } bool pdl_enabled() { return false; } }
int main(int argc, char** argv) {
    const int B = 128, H = 12, C = 768;
    const int N = argc > 1 ? atoi(argv[1]) : 577, Np = argc > 2 ? atoi(argv[2]) : 577;
    std::vector<__nv_bfloat16> h((size_t)B * N * 3 * C);
    srand(1);
    for (auto& x : h) x = __float2bfloat16((rand() % 2001 - 1000) / 500.f);
    std::vector<int> rm((size_t)B * Np);
    for (int b = 0; b < B; ++b) for (int j = 0; j < Np; ++j) rm[b * Np + j] = b * N + (int)((long long)j * N / Np);
    __nv_bfloat16 *qkv, *out; int* rmap;
    cudaMalloc(&qkv, h.size() * 2); cudaMalloc(&out, (size_t)B * Np * C * 2); cudaMalloc(&rmap, rm.size() * 4);
    cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(rmap, rm.data(), rm.size() * 4, cudaMemcpyHostToDevice);
    for (int i = 0; i < 3; ++i) rajni::launch_attention_long(qkv, Np < N ? rmap : nullptr, out, B, N, Np, C, H, 0.125f, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static long long tr[32 * 32];
    cudaMemcpyFromSymbol(tr, rajni::g_al_trace, sizeof(tr));
    const char* names[32] = {"ld:qfree", "ld:k1.0", "ld:k1.1", "ld:k1.2", "ld:kv.0", "ld:kv.1", "ld:kv.2", "", "mm:qfull", "mm:p1done", "mm:s0", "mm:pv0", "mm:pv1", "mm:pv2", "", "",
                             "sm:begin", "sm:max0", "sm:max1", "sm:max2", "sm:p0", "sm:p1", "sm:p2", "sm:O", "sm:st"};
    const long long t0 = tr[0];
    printf("item");
    for (int s = 0; s < 25; ++s) if (names[s][0]) printf(" %9s", names[s]);
    printf("\n");
    for (int n = 0; n < 12; ++n) {
        printf("%4d", n);
        for (int s = 0; s < 25; ++s) if (names[s][0]) printf(" %9lld", tr[n * 32 + s] ? tr[n * 32 + s] - t0 : -1);
        printf("\n");
    }
    return 0;
}
