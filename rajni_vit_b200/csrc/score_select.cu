// Fused token scoring + selection for one image per CTA.
//
//   score  : rajni/wrapper/importance.py:5-34   (CLS attention x sigmoid(z-scored value norm))
//   select : rajni/wrapper/attention.py:31-39   (top-k of scores[:,1:], ascending, CLS prepended)
//   carry  : rajni/wrapper/attention.py:58      (next_scores = scores[keep_idx])
//
// HBM-bound: the only large traffic is ONE pass over the K and V planes of the
// image's qkv tile (2*N*C bf16, contiguous 4C bytes per token), read with 16-byte
// streaming loads, one warp per token row.  Everything else lives in shared memory:
// per-head CLS logits [H][N], the head-averaged value rows [N][64] (fp32, so the
// centred norm is computed exactly like the reference without a second HBM pass),
// and the selection state (radix-select histogram, flags, prefix sums).
// bf16 tiles in, fp32 arithmetic throughout.
#include "common.cuh"

namespace rajni {

constexpr int kSelThreads = 512;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kHeadDim = 64;

struct ScoreSelectParams {
    const __nv_bfloat16* qkv;   // [B,N,3C] or null (select-only)
    const float* scores_in;     // [B,N] when qkv == null
    float* scores_out;          // [B,N] or null
    int32_t* keep_idx;          // [B,keep+1] or null (score-only)
    float* next_scores;         // [B,keep+1]
    int32_t* row_map;           // [B*(keep+1)] or null
    const float* pre_logit;     // [B][H*N] CLS logits already computed by score_stream_kernel, or null
    const float* pre_vm;        // [B][N*64] head-averaged value rows already computed, or null
    int N, C, H, keep;
    float eps;
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                       // scratch may still be read from a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = (lane < kSelWarps) ? scratch[lane] : 0.f;
    return warp_sum(t);                    // every warp reduces the same 16 values
}

// order-preserving float -> uint key (larger float => larger key); any NaN, whatever its sign bit, gets the largest key:
// torch.topk treats NaN as greater than every number (attention.py:35)
__device__ __forceinline__ uint32_t float_key(float f) {
    uint32_t u = __float_as_uint(f + 0.0f);        // -0.0 -> +0.0: they compare equal in the reference
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Select the `keep` largest of score[1..N-1] (ties: lower index first), always keep token 0,
// and emit ascending indices.  All threads of the CTA call this; score[] is in smem.
__device__ void select_and_emit(const float* score, uint32_t* s_hist, int* s_misc, int* s_warp_off,
                                const ScoreSelectParams& p, int b) {
    const int N = p.N, keep = p.keep, tid = threadIdx.x;
    // ---- radix select: key of the keep-th largest patch score, 8 bits per pass
    uint32_t prefix = 0, mask = 0;
    int remaining = keep;
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int n = 1 + tid; n < N; n += kSelThreads) {
            uint32_t k = float_key(score[n]);
            if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 0xff], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // warp 0: each lane owns 8 consecutive bins, highest bins in lane 0
            uint32_t c[8];
            uint32_t lane_total = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = s_hist[255 - (tid * 8 + i)]; lane_total += c[i]; }
            uint32_t incl = lane_total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            uint32_t before = incl - lane_total;          // elements in strictly higher bins of other lanes
            if (before < (uint32_t)remaining && incl >= (uint32_t)remaining) {
                uint32_t run = before;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (run < (uint32_t)remaining && run + c[i] >= (uint32_t)remaining) {
                        s_misc[0] = 255 - (tid * 8 + i);   // the digit
                        s_misc[1] = remaining - (int)run;  // still to take inside that bin
                    }
                    run += c[i];
                }
            }
        }
        __syncthreads();
        prefix |= (uint32_t)s_misc[0] << shift;
        mask |= 0xffu << shift;
        remaining = s_misc[1];
        __syncthreads();
    }
    const uint32_t thresh = prefix;        // exact key of the keep-th largest
    // `remaining` of the elements equal to thresh are kept, lowest index first.

    // ---- flags + ascending compaction. Thread t owns tokens [t*ipt, (t+1)*ipt).
    const int ipt = (N + kSelThreads - 1) / kSelThreads;
    const int n0 = tid * ipt;
    int n_gt = 0, n_eq = 0;
    for (int i = 0; i < ipt; ++i) {
        int n = n0 + i;
        if (n >= 1 && n < N) {
            uint32_t k = float_key(score[n]);
            n_gt += (k > thresh);
            n_eq += (k == thresh);
        }
    }
    if (n0 == 0 && N > 0) n_gt += 1;        // CLS is always kept (counted as "greater")
    // block exclusive scan of (n_gt, n_eq) packed: both < 2^15
    int packed = (n_gt << 16) | n_eq;
    int incl = packed;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_off[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < kSelWarps) ? s_warp_off[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kSelWarps) s_warp_off[lane] = wi - w;   // exclusive warp offsets
    }
    __syncthreads();
    int excl = incl - packed + s_warp_off[warp];
    int gt_before = excl >> 16, eq_before = excl & 0xffff;
    const int stride = keep + 1;
    for (int i = 0; i < ipt; ++i) {
        int n = n0 + i;
        if (n < N) {
            bool take;
            int pos;
            if (n == 0) {
                take = true;
                pos = 0;
            } else {
                uint32_t k = float_key(score[n]);
                bool gt = k > thresh, eq = k == thresh;
                take = gt || (eq && eq_before < remaining);
                pos = gt_before + min(eq_before, remaining);
                eq_before += eq;
            }
            if (take) {
                p.keep_idx[(size_t)b * stride + pos] = n;
                p.next_scores[(size_t)b * stride + pos] = score[n];
                if (p.row_map) p.row_map[(size_t)b * stride + pos] = b * N + n;
            }
            gt_before += take && (n == 0 || float_key(score[n]) > thresh);
        }
    }
}

// One token row: CLS logits of every head (q pre-scaled) and the head-averaged value row.  Shared by the fused kernel
// (destinations in shared memory) and the streaming kernel of the split path (destinations in global scratch), so both
// paths produce bit-identical numbers.
template <int CPL>
__device__ __forceinline__ void score_row(const __nv_bfloat16* row, const float (&q)[CPL][8], int lane, int chunks, int C, int H,
                                          int N, int n, float* logit_dst, float* vm_dst, float inv_h) {
    uint4 kk[CPL], vv[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        int j = lane + 32 * i;
        bool ok = j < chunks;
        kk[i] = ok ? ld_stream16(row + C + j * 8) : make_uint4(0, 0, 0, 0);
        vv[i] = ok ? ld_stream16(row + 2 * C + j * 8) : make_uint4(0, 0, 0, 0);
    }
    float va[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        float2 k0 = bf16x2_to_float2(kk[i].x), k1 = bf16x2_to_float2(kk[i].y);
        float2 k2 = bf16x2_to_float2(kk[i].z), k3 = bf16x2_to_float2(kk[i].w);
        float dot = q[i][0] * k0.x;
        dot = fmaf(q[i][1], k0.y, dot); dot = fmaf(q[i][2], k1.x, dot); dot = fmaf(q[i][3], k1.y, dot);
        dot = fmaf(q[i][4], k2.x, dot); dot = fmaf(q[i][5], k2.y, dot); dot = fmaf(q[i][6], k3.x, dot);
        dot = fmaf(q[i][7], k3.y, dot);
        // 8 lanes share a head (64 dims = 8 chunks)
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        int h = (lane >> 3) + 4 * i;
        if ((lane & 7) == 0 && h < H) logit_dst[(size_t)h * N + n] = dot;
        float2 v0 = bf16x2_to_float2(vv[i].x), v1 = bf16x2_to_float2(vv[i].y);
        float2 v2 = bf16x2_to_float2(vv[i].z), v3 = bf16x2_to_float2(vv[i].w);
        va[0] += v0.x; va[1] += v0.y; va[2] += v1.x; va[3] += v1.y;
        va[4] += v2.x; va[5] += v2.y; va[6] += v3.x; va[7] += v3.y;
    }
    // lanes l, l+8, l+16, l+24 hold the same 8 dims of different heads
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        va[e] += __shfl_xor_sync(0xffffffffu, va[e], 8);
        va[e] += __shfl_xor_sync(0xffffffffu, va[e], 16);
    }
    if (lane < 8) {
        float* dst = vm_dst + (size_t)n * kHeadDim + lane * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = va[e] * inv_h;       // importance.py:24
    }
}

// CLS query of the image, pre-scaled by 1/sqrt(64) (exact: power of two)
template <int CPL>
__device__ __forceinline__ void load_cls_query(const __nv_bfloat16* img, int lane, int chunks, float (&q)[CPL][8]) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        int j = lane + 32 * i;
        uint4 u = (j < chunks) ? ld_stream16(img + j * 8) : make_uint4(0, 0, 0, 0);
        float2 a = bf16x2_to_float2(u.x), bb = bf16x2_to_float2(u.y), c = bf16x2_to_float2(u.z), d = bf16x2_to_float2(u.w);
        q[i][0] = a.x * 0.125f; q[i][1] = a.y * 0.125f; q[i][2] = bb.x * 0.125f; q[i][3] = bb.y * 0.125f;
        q[i][4] = c.x * 0.125f; q[i][5] = c.y * 0.125f; q[i][6] = d.x * 0.125f; q[i][7] = d.y * 0.125f;
    }
}

// Split path, kernel 1: when a batch has far fewer images than the GPU has SMs (vit_large at 32 images per GPU), one CTA
// per image leaves most SMs idle and each CTA latency-bound.  This kernel spreads the K/V pass over (image, 16-row block)
// CTAs and leaves the per-token results in global scratch; score_select_kernel then starts from them (pre_logit / pre_vm).
constexpr int kStreamThreads = 256;
constexpr int kStreamRows = 16;
template <int CPL>
__global__ void __launch_bounds__(kStreamThreads) score_stream_kernel(const __nv_bfloat16* qkv, float* logit_g, float* vm_g,
                                                                      int N, int C, int H) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, chunks = C >> 3;
    const __nv_bfloat16* img = qkv + (size_t)b * N * 3 * C;
    float q[CPL][8];
    load_cls_query<CPL>(img, lane, chunks, q);
    const float inv_h = 1.0f / (float)H;
    float* logit_dst = logit_g + (size_t)b * ((H * N + 3) & ~3);     // per-image stride padded to 16 bytes
    float* vm_dst = vm_g + (size_t)b * N * kHeadDim;
    for (int n = blockIdx.x * kStreamRows + warp; n < min(N, (int)(blockIdx.x + 1) * kStreamRows); n += kStreamThreads / 32)
        score_row<CPL>(img + (size_t)n * 3 * C, q, lane, chunks, C, H, N, n, logit_dst, vm_dst, inv_h);
}

// CPL = 16-byte chunks per lane per plane = ceil(C / 256).
// 64 registers per thread so that two CTAs (two images) share an SM: with one CTA per SM the 256 images of a batch
// ran as two latency-bound waves on 148 SMs (57 us at N=197); resident together they take 45 us.
template <int CPL>
__global__ void __launch_bounds__(kSelThreads, 2) score_select_kernel(const ScoreSelectParams p) {
    extern __shared__ __align__(16) float smem[];
    const int N = p.N, C = p.C, H = p.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;

    griddep_launch();
    griddep_wait();
    float* s_score = smem;                         // [N]
    float* s_scratch = s_score + N;                // [64 * 8] reduction scratch (also 16-warp scratch)
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_scratch + 512);   // [256]
    int* s_misc = reinterpret_cast<int*>(s_hist + 256);                // [4]
    int* s_warp_off = s_misc + 4;                                      // [32]
    float* s_mu = reinterpret_cast<float*>(s_warp_off + 32);           // [64]
    float* s_hstat = s_mu + 64;                                        // [2*H] max, sum per head
    float* s_r = s_hstat + 64;                                         // [N]
    float* s_logit = s_r + N;                                          // [H][N]
    float* s_vm = s_logit + (size_t)H * N;                             // [N][64]
    // (s_vm start is 16-byte aligned when H*N + 2N is a multiple of 4; we only use scalar access)

    if (p.qkv != nullptr) {
        const int chunks = C >> 3;                 // 16-byte chunks per plane row
        const __nv_bfloat16* img = p.qkv + (size_t)b * N * 3 * C;
        const float inv_h = 1.0f / (float)H;
        if (p.pre_logit != nullptr) {
            // split path: the K/V pass already ran (score_stream_kernel); fetch its per-token results (L2-resident)
            const float* gl1 = p.pre_logit + (size_t)b * ((H * N + 3) & ~3);
            const float4* gl = reinterpret_cast<const float4*>(gl1);
            const float4* gv = reinterpret_cast<const float4*>(p.pre_vm + (size_t)b * N * kHeadDim);
            for (int i = tid; i < (H * N) / 4; i += kSelThreads) {
                const float4 v = __ldg(gl + i);
                s_logit[4 * i] = v.x; s_logit[4 * i + 1] = v.y; s_logit[4 * i + 2] = v.z; s_logit[4 * i + 3] = v.w;
            }
            for (int i = ((H * N) / 4) * 4 + tid; i < H * N; i += kSelThreads) s_logit[i] = __ldg(gl1 + i);
            for (int i = tid; i < N * kHeadDim / 4; i += kSelThreads) {
                const float4 v = __ldg(gv + i);
                s_vm[4 * i] = v.x; s_vm[4 * i + 1] = v.y; s_vm[4 * i + 2] = v.z; s_vm[4 * i + 3] = v.w;
            }
        } else {
            float q[CPL][8];
            load_cls_query<CPL>(img, lane, chunks, q);
            // ---- the single HBM pass: one warp per token row, K plane then V plane
#pragma unroll 2
            for (int n = warp; n < N; n += kSelWarps)
                score_row<CPL>(img + (size_t)n * 3 * C, q, lane, chunks, C, H, N, n, s_logit, s_vm, inv_h);
        }
        __syncthreads();

        // ---- mean over tokens of the head-averaged value (importance.py:25)
        {
            int d = tid & 63, g = tid >> 6;        // 8 groups of 64 threads
            float acc = 0.f;
            for (int n = g; n < N; n += kSelThreads / 64) acc += s_vm[(size_t)n * kHeadDim + d];
            s_scratch[g * 64 + d] = acc;
            __syncthreads();
            if (tid < 64) {
                float t = 0.f;
#pragma unroll
                for (int gg = 0; gg < kSelThreads / 64; ++gg) t += s_scratch[gg * 64 + tid];
                s_mu[tid] = t / (float)N;
            }
            __syncthreads();
        }
        // ---- r[n] = || vm[n] - mu ||  (importance.py:27)
        for (int n = warp; n < N; n += kSelWarps) {
            float a = s_vm[(size_t)n * kHeadDim + lane] - s_mu[lane];
            float c = s_vm[(size_t)n * kHeadDim + lane + 32] - s_mu[lane + 32];
            float ss = warp_sum(a * a + c * c);
            if (lane == 0) s_r[n] = sqrtf(ss);
        }
        // ---- per-head softmax statistics over all N tokens (importance.py:20)
        for (int h = warp; h < H; h += kSelWarps) {
            float* l = s_logit + (size_t)h * N;
            float m = -INFINITY;
            for (int n = lane; n < N; n += 32) m = fmaxf(m, l[n]);
            m = warp_max(m);
            float s = 0.f;
            for (int n = lane; n < N; n += 32) {
                float e = expf(l[n] - m);
                l[n] = e;
                s += e;
            }
            s = warp_sum(s);
            if (lane == 0) s_hstat[h] = s;
        }
        __syncthreads();
        // ---- z-score of r with the unbiased std (importance.py:28-32)
        float part = 0.f;
        for (int n = tid; n < N; n += kSelThreads) part += s_r[n];
        const float mu = block_sum(part, s_scratch) / (float)N;
        part = 0.f;
        for (int n = tid; n < N; n += kSelThreads) { float d = s_r[n] - mu; part += d * d; }
        const float var = block_sum(part, s_scratch) / (float)(N - 1);
        const float sd = sqrtf(var) + p.eps;
        for (int n = tid; n < N; n += kSelThreads) {
            float a = 0.f;
            for (int h = 0; h < H; ++h) a += s_logit[(size_t)h * N + n] / s_hstat[h];
            a *= inv_h;                                                    // importance.py:21
            float z = (s_r[n] - mu) / sd;
            float sc = a * (1.0f / (1.0f + expf(-z)));                     // importance.py:32-34
            s_score[n] = sc;
            if (p.scores_out) p.scores_out[(size_t)b * N + n] = sc;
        }
    } else {
        for (int n = tid; n < N; n += kSelThreads) s_score[n] = p.scores_in[(size_t)b * N + n];
    }
    __syncthreads();
    if (p.keep_idx != nullptr) select_and_emit(s_score, s_hist, s_misc, s_warp_off, p, b);
}

static size_t score_smem_bytes(int N, int H, bool with_score) {
    size_t floats = (size_t)N + 512 + 256 + 4 + 32 + 64 + 64 + N;
    if (with_score) floats += (size_t)H * N + (size_t)N * kHeadDim;
    return floats * sizeof(float);
}

static int launch_score_select(const ScoreSelectParams& p, int B, cudaStream_t stream) {
    const bool with_score = p.qkv != nullptr;
    size_t smem = score_smem_bytes(p.N, p.H, with_score);
    RAJNI_REQUIRE(smem <= 227 * 1024, RAJNI_EINVAL, "score_select: N=%d H=%d needs %zu B of shared memory", p.N, p.H, smem);
    int cpl = with_score ? (p.C + 255) / 256 : 1;
    void (*kern)(const ScoreSelectParams) = nullptr;
    switch (cpl) {
        case 1: kern = score_select_kernel<1>; break;
        case 2: kern = score_select_kernel<2>; break;
        case 3: kern = score_select_kernel<3>; break;
        case 4: kern = score_select_kernel<4>; break;
        default: RAJNI_REQUIRE(false, RAJNI_EINVAL, "score_select: C=%d > 1024 unsupported", p.C);
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "score_select: smem attribute: %s", cudaGetErrorString(e));
    e = launch_kernel(kern, dim3(B), dim3(kSelThreads), smem, stream, 1, p);
    count_launch();
    RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "score_select: launch failed: %s", cudaGetErrorString(e));
    return check_launch("score_select");
}

static size_t split_workspace_floats(int B, int N, int H) {
    return (size_t)B * (((size_t)H * N + 3) & ~(size_t)3) + (size_t)B * N * kHeadDim;
}

// split path, kernel 1 (see score_stream_kernel)
static int launch_score_stream(const __nv_bfloat16* qkv, int B, int N, int C, int H, float* ws, cudaStream_t stream) {
    float* logit_g = ws;
    float* vm_g = ws + (size_t)B * (((size_t)H * N + 3) & ~(size_t)3);
    void (*kern)(const __nv_bfloat16*, float*, float*, int, int, int) = nullptr;
    switch ((C + 255) / 256) {
        case 1: kern = score_stream_kernel<1>; break;
        case 2: kern = score_stream_kernel<2>; break;
        case 3: kern = score_stream_kernel<3>; break;
        case 4: kern = score_stream_kernel<4>; break;
        default: RAJNI_REQUIRE(false, RAJNI_EINVAL, "score_select: C=%d > 1024 unsupported", C);
    }
    cudaError_t e = launch_kernel(kern, dim3((N + kStreamRows - 1) / kStreamRows, B), dim3(kStreamThreads), 0, stream, 1,
                                  qkv, logit_g, vm_g, N, C, H);
    count_launch();
    RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "score_stream: launch failed: %s", cudaGetErrorString(e));
    return check_launch("score_stream");
}

static int check_score_shape(int B, int N, int C, int H) {
    RAJNI_REQUIRE(B > 0 && N >= 2 && N <= 4096, RAJNI_EINVAL, "score: bad B=%d N=%d", B, N);
    RAJNI_REQUIRE(H > 0 && C == H * kHeadDim, RAJNI_EINVAL, "score: head dim must be 64 (C=%d H=%d)", C, H);
    return 0;
}

}  // namespace rajni

using namespace rajni;

extern "C" int rajni_importance(const void* qkv, int B, int N, int C, int H, float eps,
                                float* scores, void* stream) {
    RAJNI_REQUIRE(qkv && scores, RAJNI_EINVAL, "rajni_importance: null pointer");
    if (int rc = check_score_shape(B, N, C, H)) return rc;
    ScoreSelectParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.scores_out = scores;
    p.N = N; p.C = C; p.H = H; p.keep = 0; p.eps = eps;
    return launch_score_select(p, B, static_cast<cudaStream_t>(stream));
}

extern "C" int rajni_select(const float* scores, int B, int N, int keep, int32_t* keep_idx,
                            float* next_scores, int32_t* row_map, void* stream) {
    RAJNI_REQUIRE(scores && keep_idx && next_scores, RAJNI_EINVAL, "rajni_select: null pointer");
    RAJNI_REQUIRE(B > 0 && N >= 2 && N <= 16384, RAJNI_EINVAL, "rajni_select: bad B=%d N=%d", B, N);
    RAJNI_REQUIRE(keep >= 1, RAJNI_EINVAL, "rajni_select: keep=%d < 1", keep);
    RAJNI_REQUIRE(keep <= N - 1, RAJNI_ERANGE, "selected index k out of range (keep=%d, patches=%d)", keep, N - 1);
    ScoreSelectParams p{};
    p.scores_in = scores;
    p.keep_idx = keep_idx; p.next_scores = next_scores; p.row_map = row_map;
    p.N = N; p.C = 0; p.H = 0; p.keep = keep; p.eps = 0.f;
    return launch_score_select(p, B, static_cast<cudaStream_t>(stream));
}

extern "C" int rajni_score_select(const void* qkv, int B, int N, int C, int H, int keep, float eps,
                                  float* scores, int32_t* keep_idx, float* next_scores,
                                  int32_t* row_map, void* stream) {
    RAJNI_REQUIRE(qkv && keep_idx && next_scores, RAJNI_EINVAL, "rajni_score_select: null pointer");
    if (int rc = check_score_shape(B, N, C, H)) return rc;
    RAJNI_REQUIRE(keep >= 1, RAJNI_EINVAL, "rajni_score_select: keep=%d < 1", keep);
    RAJNI_REQUIRE(keep <= N - 1, RAJNI_ERANGE, "selected index k out of range (keep=%d, patches=%d)", keep, N - 1);
    ScoreSelectParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.scores_out = scores;
    p.keep_idx = keep_idx; p.next_scores = next_scores; p.row_map = row_map;
    p.N = N; p.C = C; p.H = H; p.keep = keep; p.eps = eps;
    return launch_score_select(p, B, static_cast<cudaStream_t>(stream));
}

extern "C" size_t rajni_score_select_workspace_bytes(int B, int N, int C, int H) {
    (void)C;
    if (B <= 0 || N <= 0 || H <= 0) return 0;
    return split_workspace_floats(B, N, H) * sizeof(float);
}

extern "C" int rajni_score_select_split(const void* qkv, int B, int N, int C, int H, int keep, float eps,
                                        float* scores, int32_t* keep_idx, float* next_scores, int32_t* row_map,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    RAJNI_REQUIRE(qkv && keep_idx && next_scores && workspace, RAJNI_EINVAL, "rajni_score_select_split: null pointer");
    if (int rc = check_score_shape(B, N, C, H)) return rc;
    RAJNI_REQUIRE(keep >= 1, RAJNI_EINVAL, "rajni_score_select_split: keep=%d < 1", keep);
    RAJNI_REQUIRE(keep <= N - 1, RAJNI_ERANGE, "selected index k out of range (keep=%d, patches=%d)", keep, N - 1);
    RAJNI_REQUIRE(B <= 65535, RAJNI_EINVAL, "rajni_score_select_split: B=%d exceeds the grid limit", B);
    RAJNI_REQUIRE(workspace_bytes >= rajni_score_select_workspace_bytes(B, N, C, H) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  RAJNI_EINVAL, "rajni_score_select_split: workspace too small (%zu B) or not 16-byte aligned", workspace_bytes);
    float* ws = static_cast<float*>(workspace);
    auto s = static_cast<cudaStream_t>(stream);
    if (int rc = launch_score_stream(static_cast<const __nv_bfloat16*>(qkv), B, N, C, H, ws, s)) return rc;
    ScoreSelectParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.scores_out = scores;
    p.keep_idx = keep_idx; p.next_scores = next_scores; p.row_map = row_map;
    p.pre_logit = ws;
    p.pre_vm = ws + (size_t)B * (((size_t)H * N + 3) & ~(size_t)3);
    p.N = N; p.C = C; p.H = H; p.keep = keep; p.eps = eps;
    return launch_score_select(p, B, s);
}
