// Event timeline of the attention kernel's pipeline on CTA 0 (clock64 stamps), to see which stage paces a unit.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRAJNI_ATTN_TRACE -I../../rajni_vit_b200/csrc attn_trace.cu -o attn_trace
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rajni_vit_b200/csrc/attention_tc.cu"
namespace rajni { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return cudaGetLastError() == cudaSuccess ? 0 : -2; } bool pdl_enabled() { return false; } int current_device() { return 0; } }
int main(int argc, char** argv) {
    const int B = 256, H = 12, C = 768;
    const int N = argc > 1 ? atoi(argv[1]) : 197, Np = argc > 2 ? atoi(argv[2]) : 173;
    std::vector<__nv_bfloat16> h((size_t)B * N * 3 * C);
    srand(1);
    for (auto& x : h) x = __float2bfloat16((rand() % 2001 - 1000) / 500.f);
    std::vector<int> rm((size_t)B * Np);
    for (int b = 0; b < B; ++b) for (int j = 0; j < Np; ++j) rm[b * Np + j] = b * N + (j * N) / Np;
    __nv_bfloat16 *qkv, *out; int* rmap;
    cudaMalloc(&qkv, h.size() * 2); cudaMalloc(&out, (size_t)B * Np * C * 2); cudaMalloc(&rmap, rm.size() * 4);
    cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(rmap, rm.data(), rm.size() * 4, cudaMemcpyHostToDevice);
    for (int i = 0; i < 3; ++i) rajni::launch_attention_tc(qkv, Np < N ? rmap : nullptr, out, B, N, Np, C, H, 0.125f, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    static long long tr[64 * 32];
    cudaMemcpyFromSymbol(tr, rajni::g_attn_trace, sizeof(tr));
    const char* names[32] = {"ld:empty", "ld:issued", "mma:full", "", "pv0:beg", "pv0:end", "s0:beg", "s0:end", "pv1:beg", "pv1:end", "s1:beg", "s1:end",
                             "sm0:S", "sm0:max", "sm0:P", "sm0:O", "sm0:Ord", "sm0:st", "", "", "sm1:S", "sm1:max", "sm1:P", "sm1:O", "sm1:Ord", "sm1:st"};
    const long long t0 = tr[0];
    printf("unit");
    for (int s = 0; s < 26; ++s) if (names[s][0]) printf(" %9s", names[s]);
    printf("\n");
    for (int n = 0; n < 20; ++n) {
        printf("%4d", n);
        for (int s = 0; s < 26; ++s) if (names[s][0]) printf(" %9lld", tr[n * 32 + s] ? tr[n * 32 + s] - t0 : -1);
        printf("\n");
    }
    return 0;
}
