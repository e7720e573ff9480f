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
#include <algorithm>

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
    int N, C, H, keep;
    float eps;
};

// shared memory of one image's tail (statistics + selection)
struct SelSmem {
    float* score;        // [N]
    float* scratch;      // [512] reduction scratch
    uint32_t* hist;      // [256]
    int* misc;           // [4]
    int* warp_off;       // [32]
    float* mu;           // [64]
    float* hstat;        // [64] per-head softmax sums
    float* r;            // [N]
    float* logit;        // [H][N]
    float* tail;         // whatever follows (the fused kernel keeps the value rows here)
};
__device__ __forceinline__ SelSmem carve_sel_smem(float* smem, int N, int H) {
    SelSmem s;
    s.score = smem;
    s.scratch = s.score + N;
    s.hist = reinterpret_cast<uint32_t*>(s.scratch + 512);
    s.misc = reinterpret_cast<int*>(s.hist + 256);
    s.warp_off = s.misc + 4;
    s.mu = reinterpret_cast<float*>(s.warp_off + 32);
    s.hstat = s.mu + 64;
    s.r = s.hstat + 64;
    s.logit = s.r + N;
    s.tail = s.logit + (size_t)H * N;
    return s;
}

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
// The CLS query sits in registers (q) or, where the registers are wanted for a second row in flight, in shared memory
// (qs: floats 0-3 of chunk j at qs[4*j], floats 4-7 at qs[C/2 + 4*j]; conflict-free 16-byte reads).  Same values, same order.
template <int CPL, bool kQShared>
__device__ __forceinline__ void score_row(const __nv_bfloat16* row, const float (&q)[CPL][8], const float* qs, int lane, int chunks,
                                          int C, int H, int N, int n, float* logit_dst, float* vm_dst, float inv_h) {
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
        float4 qa, qb;
        if (kQShared) {
            const int j = min(lane + 32 * i, chunks - 1);
            qa = *reinterpret_cast<const float4*>(qs + 4 * j);
            qb = *reinterpret_cast<const float4*>(qs + (C >> 1) + 4 * j);
        } else {
            qa = make_float4(q[i][0], q[i][1], q[i][2], q[i][3]);
            qb = make_float4(q[i][4], q[i][5], q[i][6], q[i][7]);
        }
        float dot = qa.x * k0.x;
        dot = fmaf(qa.y, k0.y, dot); dot = fmaf(qa.z, k1.x, dot); dot = fmaf(qa.w, k1.y, dot);
        dot = fmaf(qb.x, k2.x, dot); dot = fmaf(qb.y, k2.y, dot); dot = fmaf(qb.z, k3.x, dot);
        dot = fmaf(qb.w, k3.y, dot);
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

// The per-image tail: statistics of the value rows and the CLS logits, scores, selection.  Every thread of a 512-thread CTA
// calls it after a __syncthreads() that made sm.logit (shared) and vm (shared in the fused kernel, global scratch written by
// other CTAs of the same launch in the overlapped one: read past L1) complete.  Same instruction order for both, so the two
// kernels agree bit for bit.
template <bool kVmGlobal>
__device__ __forceinline__ float vm_ld(const float* p) { return kVmGlobal ? __ldcg(p) : *p; }

template <bool kVmGlobal>
__device__ void score_tail(const SelSmem& sm, const float* vm, const ScoreSelectParams& p, int b) {
    const int N = p.N, H = p.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv_h = 1.0f / (float)H;
    // ---- mean over tokens of the head-averaged value (importance.py:25)
    {
        int d = tid & 63, g = tid >> 6;        // 8 groups of 64 threads
        float acc = 0.f;
        for (int n = g; n < N; n += kSelThreads / 64) acc += vm_ld<kVmGlobal>(vm + (size_t)n * kHeadDim + d);
        sm.scratch[g * 64 + d] = acc;
        __syncthreads();
        if (tid < 64) {
            float t = 0.f;
#pragma unroll
            for (int gg = 0; gg < kSelThreads / 64; ++gg) t += sm.scratch[gg * 64 + tid];
            sm.mu[tid] = t / (float)N;
        }
        __syncthreads();
    }
    // ---- r[n] = || vm[n] - mu ||  (importance.py:27)
    for (int n = warp; n < N; n += kSelWarps) {
        float a = vm_ld<kVmGlobal>(vm + (size_t)n * kHeadDim + lane) - sm.mu[lane];
        float c = vm_ld<kVmGlobal>(vm + (size_t)n * kHeadDim + lane + 32) - sm.mu[lane + 32];
        float ss = warp_sum(a * a + c * c);
        if (lane == 0) sm.r[n] = sqrtf(ss);
    }
    // ---- per-head softmax statistics over all N tokens (importance.py:20)
    for (int h = warp; h < H; h += kSelWarps) {
        float* l = sm.logit + (size_t)h * N;
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
        if (lane == 0) sm.hstat[h] = s;
    }
    __syncthreads();
    // ---- z-score of r with the unbiased std (importance.py:28-32)
    float part = 0.f;
    for (int n = tid; n < N; n += kSelThreads) part += sm.r[n];
    const float mu = block_sum(part, sm.scratch) / (float)N;
    part = 0.f;
    for (int n = tid; n < N; n += kSelThreads) { float d = sm.r[n] - mu; part += d * d; }
    const float var = block_sum(part, sm.scratch) / (float)(N - 1);
    const float sd = sqrtf(var) + p.eps;
    for (int n = tid; n < N; n += kSelThreads) {
        float a = 0.f;
        for (int h = 0; h < H; ++h) a += sm.logit[(size_t)h * N + n] / sm.hstat[h];
        a *= inv_h;                                                    // importance.py:21
        float z = (sm.r[n] - mu) / sd;
        float sc = a * (1.0f / (1.0f + expf(-z)));                     // importance.py:32-34
        sm.score[n] = sc;
        if (p.scores_out) p.scores_out[(size_t)b * N + n] = sc;
    }
    __syncthreads();
    if (p.keep_idx != nullptr) select_and_emit(sm.score, sm.hist, sm.misc, sm.warp_off, p, b);
}

// CPL = 16-byte chunks per lane per plane = ceil(C / 256).
// One CTA per image, everything in shared memory, no scratch: the stand-alone entry points (rajni_importance, rajni_select,
// rajni_score_select).  64 registers per thread so that two CTAs (two images) share an SM.  All images of a batch are
// resident at once and march in lock-step - pass, then tail - so the tail (~25 % of the time) is never hidden; the model path
// uses score_overlap_kernel below.
template <int CPL>
__global__ void __launch_bounds__(kSelThreads, 2) score_select_kernel(const ScoreSelectParams p) {
    extern __shared__ __align__(16) float smem[];
    const int N = p.N, C = p.C, H = p.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;

    griddep_launch();
    griddep_wait();
    const SelSmem sm = carve_sel_smem(smem, N, H);
    if (p.qkv != nullptr) {
        float* s_vm = sm.tail;                     // [N][64] (scalar access only)
        const int chunks = C >> 3;                 // 16-byte chunks per plane row
        const __nv_bfloat16* img = p.qkv + (size_t)b * N * 3 * C;
        float q[CPL][8];
        load_cls_query<CPL>(img, lane, chunks, q);
        // ---- the single HBM pass: one warp per token row, K plane then V plane
#pragma unroll 2
        for (int n = warp; n < N; n += kSelWarps)
            score_row<CPL, false>(img + (size_t)n * 3 * C, q, nullptr, lane, chunks, C, H, N, n, sm.logit, s_vm, 1.0f / (float)H);
        __syncthreads();
        score_tail<false>(sm, s_vm, p, b);
    } else {
        for (int n = tid; n < N; n += kSelThreads) sm.score[n] = p.scores_in[(size_t)b * N + n];
        __syncthreads();
        if (p.keep_idx != nullptr) select_and_emit(sm.score, sm.hist, sm.misc, sm.warp_off, p, b);
    }
}

// The model path.  CTA (blk, b) streams rows [blk*rpb, (blk+1)*rpb) of image b - two rows per warp, both in flight - and
// leaves the per-token results (CLS logits [H][N], head-averaged value rows [N][64], fp32) in scratch, which stays in L2.
// The CTA that finishes an image's LAST block (a counter per image; it puts the counter back to zero) runs the image's tail
// from that scratch while the other CTAs on the SM and on the chip keep streaming: pass and tail overlap across images, and
// the grid is ~7 CTAs per image instead of one, so small batches fill the GPU too.
struct ScoreScratch {
    int* count;       // [B] zero before the first launch; every launch leaves it zero
    float* logit;     // [B][pad4(H*N)]
    float* vm;        // [B][N*64]
    int nblk, rpb;    // row blocks per image, rows per block
};

template <int CPL>
__global__ void __launch_bounds__(kSelThreads, 2) score_overlap_kernel(const ScoreSelectParams p, const ScoreScratch w) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_last;
    const int N = p.N, C = p.C, H = p.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, blk = blockIdx.x;

    griddep_launch();
    griddep_wait();
    const int chunks = C >> 3;
    const __nv_bfloat16* img = p.qkv + (size_t)b * N * 3 * C;
    const size_t lstride = ((size_t)H * N + 3) & ~(size_t)3;
    float* gl = w.logit + (size_t)b * lstride;
    float* gv = w.vm + (size_t)b * N * kHeadDim;
    {
        // CLS query, pre-scaled by 1/8 (exact), into shared memory
        for (int j = tid; j < chunks; j += kSelThreads) {
            const uint4 u = ld_stream16(img + j * 8);
            const float2 a = bf16x2_to_float2(u.x), bb = bf16x2_to_float2(u.y), c = bf16x2_to_float2(u.z), d = bf16x2_to_float2(u.w);
            *reinterpret_cast<float4*>(smem + 4 * j) = make_float4(a.x * 0.125f, a.y * 0.125f, bb.x * 0.125f, bb.y * 0.125f);
            *reinterpret_cast<float4*>(smem + (C >> 1) + 4 * j) = make_float4(c.x * 0.125f, c.y * 0.125f, d.x * 0.125f, d.y * 0.125f);
        }
        __syncthreads();
        const float dummy[CPL][8] = {};
        const int n_end = min(N, (blk + 1) * w.rpb);
#pragma unroll 2
        for (int n = blk * w.rpb + warp; n < n_end; n += kSelWarps)
            score_row<CPL, true>(img + (size_t)n * 3 * C, dummy, smem, lane, chunks, C, H, N, n, gl, gv, 1.0f / (float)H);
    }
    __threadfence();                               // this thread's scratch writes are visible device-wide ...
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(&w.count[b], 1); // ... before the arrival is
        s_last = (old == w.nblk - 1);
        if (s_last) w.count[b] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const SelSmem sm = carve_sel_smem(smem, N, H);
    {
        const float4* g4 = reinterpret_cast<const float4*>(gl);
        for (int i = tid; i < (H * N) / 4; i += kSelThreads) {
            const float4 v = __ldcg(g4 + i);
            sm.logit[4 * i] = v.x; sm.logit[4 * i + 1] = v.y; sm.logit[4 * i + 2] = v.z; sm.logit[4 * i + 3] = v.w;
        }
        for (int i = ((H * N) / 4) * 4 + tid; i < H * N; i += kSelThreads) sm.logit[i] = __ldcg(gl + i);
    }
    __syncthreads();
    score_tail<true>(sm, gv, p, b);
}

static size_t score_smem_bytes(int N, int H, bool with_logit, bool with_vm) {
    size_t floats = (size_t)N + 512 + 256 + 4 + 32 + 64 + 64 + N;
    if (with_logit) floats += (size_t)H * N;
    if (with_vm) floats += (size_t)N * kHeadDim;
    return floats * sizeof(float);
}

static int launch_score_select(const ScoreSelectParams& p, int B, cudaStream_t stream) {
    const bool with_score = p.qkv != nullptr;
    size_t smem = score_smem_bytes(p.N, p.H, with_score, with_score);
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

constexpr int kOverlapRows = 2 * kSelWarps;        // rows of one CTA: two per warp

static size_t scratch_count_ints(int B) { return ((size_t)B + 3) & ~(size_t)3; }
static size_t scratch_floats(int B, int N, int H) {
    return scratch_count_ints(B) + (size_t)B * (((size_t)H * N + 3) & ~(size_t)3) + (size_t)B * N * kHeadDim;
}

static int launch_score_overlap(const ScoreSelectParams& p, int B, float* ws, cudaStream_t stream) {
    ScoreScratch w;
    w.count = reinterpret_cast<int*>(ws);
    w.logit = ws + scratch_count_ints(B);
    w.vm = w.logit + (size_t)B * (((size_t)p.H * p.N + 3) & ~(size_t)3);
    w.nblk = (p.N + kOverlapRows - 1) / kOverlapRows;
    w.rpb = (p.N + w.nblk - 1) / w.nblk;
    size_t smem = std::max(score_smem_bytes(p.N, p.H, true, false), (size_t)p.C * sizeof(float));   // tail state / the CLS query
    RAJNI_REQUIRE(smem <= 227 * 1024, RAJNI_EINVAL, "score_select: N=%d H=%d needs %zu B of shared memory", p.N, p.H, smem);
    void (*kern)(const ScoreSelectParams, const ScoreScratch) = nullptr;
    switch ((p.C + 255) / 256) {
        case 1: kern = score_overlap_kernel<1>; break;
        case 2: kern = score_overlap_kernel<2>; break;
        case 3: kern = score_overlap_kernel<3>; break;
        case 4: kern = score_overlap_kernel<4>; break;
        default: RAJNI_REQUIRE(false, RAJNI_EINVAL, "score_select: C=%d > 1024 unsupported", p.C);
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "score_select: smem attribute: %s", cudaGetErrorString(e));
    e = launch_kernel(kern, dim3(w.nblk, B), dim3(kSelThreads), smem, stream, 1, p, w);
    count_launch();
    RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "score_select: launch failed: %s", cudaGetErrorString(e));
    return check_launch("score_select");
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
    return scratch_floats(B, N, H) * sizeof(float);
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
    ScoreSelectParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.scores_out = scores;
    p.keep_idx = keep_idx; p.next_scores = next_scores; p.row_map = row_map;
    p.N = N; p.C = C; p.H = H; p.keep = keep; p.eps = eps;
    return launch_score_overlap(p, B, static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}
