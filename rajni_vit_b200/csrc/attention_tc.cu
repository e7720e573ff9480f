// tcgen05 attention over kept tokens, one CTA per (image, head), Np <= 256 keys.
//
//   out[b, i, h*64:(h+1)*64] = softmax_j( q_i . k_j * scale ) v_j        attention.py:45-54
//   token j of image b is read from global qkv row row_map[b*Np + j]     attention.py:42-43 (gather fused)
//
// Per CTA: the head's Q (up to 2 tiles of 128 rows), K and V slices are gathered row by row into
// 128-byte-swizzled shared memory with cp.async (every row is one 128 B head slice; rows past Np are
// zero-filled).  Then for each 128-query tile
//   S = Q K^T       : tcgen05.mma, A = Q (K-major), B = K (K-major), N = Np rounded to 16, fp32 in TMEM
//   P = softmax(S)  : 128 threads, one row each, two passes over the TMEM row; P is written back to
//                     TMEM as packed bf16 (aliasing the S columns already consumed)
//   O = P V         : tcgen05.mma with A = P from TMEM and B = V from smem (MN-major), N = 64
//   epilogue        : O * (1/rowsum) -> bf16 -> global
// TMEM: 256 columns per CTA (S at [0,256), P at [0,128), O at [128,192)); ~85 KB smem => 2 CTAs per SM,
// which is what overlaps one CTA's gather/softmax with the other's MMAs.
#include "common.cuh"

namespace rajni {

constexpr int kTcThreads = 160;          // warps 0-3: loads + softmax (one TMEM sub-partition each); warp 4: MMA issuer
constexpr int kTcTmemCols = 256;
constexpr int kTcOCol = 128;

struct AttnTcParams {
    const __nv_bfloat16* qkv;
    const int32_t* row_map;
    __nv_bfloat16* out;
    int N_src, Np, Np_pad, C, H, nqt;
    float scale_log2;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const void* gsrc, bool valid) {
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_dst), "l"(gsrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kTcThreads) attention_tc_kernel(const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int q_rows = p.nqt * 128;
    const uint32_t s_q = smem_base;                                 // [q_rows][128 B]
    const uint32_t s_k = s_q + q_rows * 128;                        // [Np_pad][128 B]
    const uint32_t s_v = s_k + ((p.Np_pad * 128 + 1023) & ~1023);   // [Np_pad][128 B]
    __shared__ __align__(8) uint64_t bar_s, bar_o, bar_p, bar_done;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_rows[256];             // global qkv row of every kept token of this image

    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 4) {
        tmem_alloc(&tmem_slot, kTcTmemCols);
        if (lane == 0) {
            mbar_init(&bar_s, 1);
            mbar_init(&bar_o, 1);
            mbar_init(&bar_p, 128);
            mbar_init(&bar_done, 128);
            mbar_fence_init();
        }
    }
    for (int j = tid; j < p.Np; j += kTcThreads)
        s_rows[j] = p.row_map ? __ldg(p.row_map + (long long)b * p.Np + j) : b * p.N_src + j;
    __syncthreads();
    // ---- gather Q, K, V head slices: 8 lanes move one 128-byte row
    {
        const int chunk = tid & 7;
        const long long head_off = (long long)h * 64 + chunk * 8;
        for (int r = tid >> 3; r < q_rows + 2 * p.Np_pad; r += kTcThreads / 8) {
            int tok, plane;
            uint32_t dst;
            if (r < q_rows) { tok = r; plane = 0; dst = s_q + r * 128; }
            else if (r < q_rows + p.Np_pad) { tok = r - q_rows; plane = 1; dst = s_k + tok * 128; }
            else { tok = r - q_rows - p.Np_pad; plane = 2; dst = s_v + tok * 128; }
            const bool ok = tok < p.Np;
            const long long grow = ok ? (long long)s_rows[tok] : 0;
            cp_async16_zfill(dst + ((chunk ^ (tok & 7)) << 4), p.qkv + grow * 3 * p.C + plane * p.C + head_off, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_async_smem();                 // cp.async wrote through the generic proxy; UMMA reads through the async proxy
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 4) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, p.Np_pad, 0, 0);
            const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);          // B = V is MN-major
            for (int qt = 0; qt < p.nqt; ++qt) {
                const uint32_t ph = qt & 1;
                if (qt > 0) { mbar_wait(&bar_done, ph ^ 1); tc_fence_after(); }   // previous tile's O has been read
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, umma_desc_sw128(s_q + qt * 128 * 128 + k * 32, 16, 1024),
                              umma_desc_sw128(s_k + k * 32, 16, 1024), idesc_s, k != 0);
                umma_commit(&bar_s);
                mbar_wait(&bar_p, ph);                                         // P is in TMEM
                tc_fence_after();
                for (int k = 0; k < p.Np_pad / 16; ++k)
                    umma_bf16_ts(tmem_base + kTcOCol, tmem_base + k * 8,
                                 umma_desc_sw128(s_v + k * 2048, 16, 1024), idesc_o, k != 0);
                umma_commit(&bar_o);
            }
        }
    } else {
        // ================= softmax + epilogue: thread = query row = TMEM lane =================
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int qt = 0; qt < p.nqt; ++qt) {
            const uint32_t ph = qt & 1;
            mbar_wait(&bar_s, ph);
            tc_fence_after();
            float mx = -INFINITY;
            for (int c0 = 0; c0 < p.Np_pad; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(trow + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < p.Np) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
            const float mb = mx * p.scale_log2;
            float sum = 0.f;
            for (int c0 = 0; c0 < p.Np_pad; c0 += 32) {
                uint32_t v[32], pk[16];
                tmem_ld32(trow + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float e0 = (c0 + j < p.Np) ? exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2, -mb)) : 0.f;
                    float e1 = (c0 + j + 1 < p.Np) ? exp2f(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mb)) : 0.f;
                    sum += e0 + e1;
                    pk[j >> 1] = float2_to_bf16x2(e0, e1);
                }
                tmem_st16(trow + (c0 >> 1), pk);         // P (bf16 pairs) trails the S columns it overwrites
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bar_p);
            // ---- O = P V done -> normalise, store
            mbar_wait(&bar_o, ph);
            tc_fence_after();
            const float inv = 1.f / sum;
            const int q = qt * 128 + warp * 32 + lane;
            uint32_t o0[32], o1[32];
            tmem_ld32(trow + kTcOCol, o0);
            tmem_ld32(trow + kTcOCol + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bar_done);
            if (q < p.Np) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.Np + q) * p.C + h * 64);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    dst[c] = make_uint4(float2_to_bf16x2(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv),
                                        float2_to_bf16x2(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv),
                                        float2_to_bf16x2(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv),
                                        float2_to_bf16x2(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv));
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    dst[4 + c] = make_uint4(float2_to_bf16x2(__uint_as_float(o1[8 * c]) * inv, __uint_as_float(o1[8 * c + 1]) * inv),
                                            float2_to_bf16x2(__uint_as_float(o1[8 * c + 2]) * inv, __uint_as_float(o1[8 * c + 3]) * inv),
                                            float2_to_bf16x2(__uint_as_float(o1[8 * c + 4]) * inv, __uint_as_float(o1[8 * c + 5]) * inv),
                                            float2_to_bf16x2(__uint_as_float(o1[8 * c + 6]) * inv, __uint_as_float(o1[8 * c + 7]) * inv));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTcTmemCols);
    }
}

// returns 1 if the tcgen05 kernel handled the call, 0 if the shape is outside its range, <0 on error
int launch_attention_tc(const void* qkv, const int32_t* row_map, void* out, int B, int N_src, int Np,
                        int C, int H, float scale, cudaStream_t stream) {
    if (Np > 256) return 0;
    AttnTcParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.row_map = row_map;
    p.out = static_cast<__nv_bfloat16*>(out);
    p.N_src = N_src; p.Np = Np; p.C = C; p.H = H;
    p.Np_pad = (Np + 15) & ~15;
    p.nqt = (Np + 127) / 128;
    p.scale_log2 = scale * 1.4426950408889634f;
    const int kv_bytes = (p.Np_pad * 128 + 1023) & ~1023;
    const int smem = p.nqt * 128 * 128 + 2 * kv_bytes + 1024;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "attention_tc: smem attribute: %s", cudaGetErrorString(e));
        attr_smem = 100 * 1024;
    }
    attention_tc_kernel<<<dim3(H, B), kTcThreads, smem, stream>>>(p);
    count_launch();
    int rc = check_launch("attention_tc");
    return rc ? rc : 1;
}

}  // namespace rajni
