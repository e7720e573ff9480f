// Multi-head attention over kept tokens with the token gather fused into the loads.
//
//   out[b, i, h*64:(h+1)*64] = softmax_j( q_i . k_j * scale ) v_j        attention.py:45-54
//   token j of image b lives at global qkv row  row_map[b*Np + j]        attention.py:42-43
//
// Round-1 kernel: flash-style single pass, warp-level mma.sync (m16n8k16, bf16 -> fp32),
// one CTA per (image, head, 64-query tile), 4 warps x 16 query rows, K/V streamed in
// 64-key blocks through a double-buffered cp.async ring with 128-byte XOR-swizzled rows
// (conflict-free ldmatrix).  The gather costs nothing extra: every 128-byte head slice of
// a token row is fetched by its own row index.  (The tcgen05/TMEM version of this kernel is
// the next step; attention is 3-9 % of the path's flops, the GEMMs went first.)
#include <cstdlib>

#include "common.cuh"

namespace rajni {

constexpr int kAttThreads = 128;
constexpr int kAttBQ = 64;      // queries per CTA
constexpr int kAttBK = 64;      // keys per block

struct AttnParams {
    const __nv_bfloat16* qkv;
    const int32_t* row_map;
    __nv_bfloat16* out;
    int N_src, Np, C, H;
    float scale_log2;           // scale * log2(e)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = smem_u32(smem_dst);
    const int bytes = valid ? 16 : 0;       // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile: [64 rows][64 bf16] with 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
__device__ __forceinline__ __nv_bfloat16* tile_ptr(__nv_bfloat16* tile, int r, int chunk) {
    return tile + r * 64 + ((chunk ^ (r & 7)) << 3);
}

// load 64 token rows (one head slice, 128 B each) of plane `plane` (0=q,1=k,2=v) into a tile
__device__ __forceinline__ void load_tile(__nv_bfloat16* tile, const AttnParams& p, int b, int h, int plane, int tok0) {
    const int chunk = threadIdx.x & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (threadIdx.x >> 3) + 16 * i;
        const int tok = tok0 + r;
        const bool ok = tok < p.Np;
        long long grow = 0;
        if (ok) grow = p.row_map ? (long long)__ldg(p.row_map + (long long)b * p.Np + tok) : (long long)b * p.N_src + tok;
        const __nv_bfloat16* src = p.qkv + grow * 3 * p.C + plane * p.C + h * 64 + chunk * 8;
        cp_async16(tile_ptr(tile, r, chunk), src, ok);
    }
}

__global__ void __launch_bounds__(kAttThreads) attention_kernel(const AttnParams p) {
    __shared__ __align__(128) __nv_bfloat16 s_q[kAttBQ * 64];
    __shared__ __align__(128) __nv_bfloat16 s_k[2][kAttBK * 64];
    __shared__ __align__(128) __nv_bfloat16 s_v[2][kAttBK * 64];

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int q0 = qt * kAttBQ;
    const int nkb = (p.Np + kAttBK - 1) / kAttBK;
    griddep_launch();
    griddep_wait();

    load_tile(s_q, p, b, h, 0, q0);
    load_tile(s_k[0], p, b, h, 1, 0);
    load_tile(s_v[0], p, b, h, 2, 0);
    cp_async_commit();

    uint32_t qf[4][4];                      // A fragments of this warp's 16 query rows, 4 k-steps
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int kb = 0; kb < nkb; ++kb) {
        const int cur = kb & 1;
        if (kb + 1 < nkb) {
            load_tile(s_k[cur ^ 1], p, b, h, 1, (kb + 1) * kAttBK);
            load_tile(s_v[cur ^ 1], p, b, h, 2, (kb + 1) * kAttBK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kb == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                ldmatrix_x4(qf[kk], tile_ptr(s_q, warp * 16 + (lane & 15), kk * 2 + (lane >> 4)));
        }
        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t bf[4];
                const int key = (jp * 2 + (lane >> 4)) * 8 + (lane & 7);
                ldmatrix_x4(bf, tile_ptr(s_k[cur], key, kk * 2 + ((lane >> 3) & 1)));
                mma_bf16_16816(s[jp * 2], qf[kk], bf[0], bf[1]);
                mma_bf16_16816(s[jp * 2 + 1], qf[kk], bf[2], bf[3]);
            }
        }
        // ---- mask keys beyond Np, online softmax (base-2, scale folded in)
        const int key0 = kb * kAttBK;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kcol = key0 + j * 8 + tig * 2;
            if (kcol >= p.Np) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (kcol + 1 >= p.Np) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);          // finite: key 0 of block 0 is always valid
            corr[r] = exp2f((m_run[r] - m_new) * p.scale_log2);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];                   // P as A fragments for 4 k-steps of 16 keys
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f((s[j][0] - m_run[0]) * p.scale_log2);
            const float p1 = exp2f((s[j][1] - m_run[0]) * p.scale_log2);
            const float p2 = exp2f((s[j][2] - m_run[1]) * p.scale_log2);
            const float p3 = exp2f((s[j][3] - m_run[1]) * p.scale_log2);
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            pf[j >> 1][(j & 1) * 2 + 0] = float2_to_bf16x2(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = float2_to_bf16x2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1]; }
        // ---- O += P V
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int jd = 0; jd < 4; ++jd) {
                uint32_t bf[4];
                const int key = t * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
                ldmatrix_x4_trans(bf, tile_ptr(s_v[cur], key, jd * 2 + (lane >> 4)));
                mma_bf16_16816(o[jd * 2], pf[t], bf[0], bf[1]);
                mma_bf16_16816(o[jd * 2 + 1], pf[t], bf[2], bf[3]);
            }
        }
        __syncthreads();                    // everyone is done with buffer `cur` before it is refilled
    }
    // ---- normalise and store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = h * 64 + j * 8 + tig * 2;
        if (row_a < p.Np)
            *reinterpret_cast<uint32_t*>(p.out + ((long long)b * p.Np + row_a) * p.C + col) = float2_to_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
        if (row_b < p.Np)
            *reinterpret_cast<uint32_t*>(p.out + ((long long)b * p.Np + row_b) * p.C + col) = float2_to_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
}

int launch_attention_tc(const void* qkv, const int32_t* row_map, void* out, int B, int N_src, int Np,
                        int C, int H, float scale, int reverse, cudaStream_t stream);   // attention_tc.cu
int launch_attention_long(const void* qkv, const int32_t* row_map, void* out, int B, int N_src, int Np,
                          int C, int H, float scale, int reverse, cudaStream_t stream);  // attention_long.cu

}  // namespace rajni

using namespace rajni;

extern "C" int rajni_attention_fwd(const void* qkv, const int32_t* row_map, void* out,
                                   int B, int N_src, int Np, int C, int H, float scale, int reverse, void* stream) {
    RAJNI_REQUIRE(qkv && out, RAJNI_EINVAL, "rajni_attention_fwd: null pointer");
    RAJNI_REQUIRE(B > 0 && Np > 0 && N_src >= Np && H > 0 && C == H * 64, RAJNI_EINVAL,
                  "rajni_attention_fwd: B=%d N_src=%d Np=%d C=%d H=%d (head dim must be 64)", B, N_src, Np, C, H);
    RAJNI_REQUIRE(row_map || N_src == Np, RAJNI_EINVAL, "rajni_attention_fwd: N_src != Np needs a row_map");
    RAJNI_REQUIRE(B <= 65535 && H <= 65535, RAJNI_EINVAL, "rajni_attention_fwd: B or H exceeds grid limits");
    // tcgen05 kernels: attention_tc for sequences that fit one TMEM score tile (every 224-px config), attention_long
    // (key blocks, two passes) for longer ones (the 577-token config).  RAJNI_ATTN_LEGACY=1 selects the round-1
    // mma.sync kernel below instead (kept for A/B timing and as an independent implementation for the tests).
    static const bool legacy = getenv("RAJNI_ATTN_LEGACY") != nullptr;
    if (!legacy) {
        int rc = Np <= 256 ? launch_attention_tc(qkv, row_map, out, B, N_src, Np, C, H, scale, reverse, static_cast<cudaStream_t>(stream))
                           : launch_attention_long(qkv, row_map, out, B, N_src, Np, C, H, scale, reverse, static_cast<cudaStream_t>(stream));
        if (rc != 0) return rc < 0 ? rc : RAJNI_OK;
    }
    AttnParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.row_map = row_map;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.N_src = N_src; p.Np = Np; p.C = C; p.H = H;
    p.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((Np + kAttBQ - 1) / kAttBQ, H, B);
    cudaError_t le = launch_kernel(attention_kernel, grid, dim3(kAttThreads), 0, static_cast<cudaStream_t>(stream), 1, p);
    RAJNI_REQUIRE(le == cudaSuccess, RAJNI_ECUDA, "attention_fwd: launch failed: %s", cudaGetErrorString(le));
    count_launch();
    return check_launch("attention_fwd");
}
