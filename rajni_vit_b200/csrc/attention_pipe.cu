// Pipelined tcgen05 attention over kept tokens (Np <= 224 keys, head dim 64): every 224-px configuration.
//
//   out[b, i, h*64:(h+1)*64] = softmax_j( q_i . k_j * scale ) v_j        attention.py:45-54
//   token j of image b is read from global qkv row row_map[b*Np + j]     attention.py:42-43 (gather fused)
//
// One CTA per SM walks a sequence of 128-query TILES (an (image, head) item has one tile when Np <= 128, else two that
// share K/V).  Each stage of the softmax has its own warps and the score tile is multi-buffered in TMEM, so the
// MUFU-bound exponentials of tile g overlap the row maxima of tile g+1, the read-out of tile g-1 and the tensor pipe:
//
//   warps 0-7  EXP     thread = (query row, half of the key columns): p = exp2(s*scale*log2e - max), bf16 P written over
//                      S in TMEM (each half in place behind its own read pointer), partial row sums to shared memory.
//                      Two warps per scheduler keep MUFU.EX2 saturated (16/clk/SM); they do nothing else.
//   warps 8-11 HELPER  thread = query row: (a) exact row maximum of S(g) as soon as the tensor pipe delivers it - one
//                      tile AHEAD of the exp warps; (b) O(g) = P V out of TMEM, times 1/rowsum, bf16, swizzled shared
//                      memory, one TMA store per warp (3-d map clipped at the image's Np rows).
//   warp 12    MMA     one lane issues S(g) = Q K^T (SS, 4 k-steps) and O(g) = P V (TS: A = P from TMEM, V MN-major)
//   warps 13-15 LOAD   dense call: one thread, three TMA boxes per item (3-d map: rows past the image are ZERO-filled, so
//                      one image's Inf/NaN can never reach another's output); gathered call: cp.async row gather by row_map
//
// TMEM (512 columns): nbuf score buffers of s_stride columns (2 at Np_pad 208 ... 4 at Np_pad <= 96), then one or two
// 64-column O buffers.  Softmax is exactly two-pass (true row maximum), like the reference.
// Per 128x208 tile: MUFU floor 1664 cycles, tensor pipe 1800 (S 4x144 + PV 13x94, profiles/r1_mma_ldtm_mufu_probe.txt).
#include <cuda.h>

#include "common.cuh"

namespace rajni {

constexpr int kApHelpWarp0 = 8;
constexpr int kApMmaWarp = 12;
constexpr int kApLoaderWarp0 = 13;
constexpr int kApLoaderThreads = 96;
constexpr int kApLoaderGroups = kApLoaderThreads / 8;                          // 8 lanes move one token's 128-byte head slice
constexpr int kApSweeps = (224 + kApLoaderGroups - 1) / kApLoaderGroups;       // sweeps of the groups over <= 224 token rows
constexpr int kApThreads = 512;
constexpr int kApMaxStages = 4;
constexpr int kApMaxBufs = 4;
constexpr int kApSumSlots = 8;
constexpr int kApOutStage = 4 * 4096;                                          // 32 rows x 128 B per helper warp
constexpr int kApAuxBytes = kApMaxBufs * 128 * 4 + kApSumSlots * 2 * 128 * 4;  // row maxima, partial row sums
constexpr int kApBarBytes = 256;
constexpr int kApSmemBudget = 227 * 1024 - 1024 - kApOutStage - kApAuxBytes - kApBarBytes;

#ifdef RAJNI_ATTN_TRACE
__device__ long long g_ap_trace[64 * 16];
#define AP_TRACE(g, slot) do { if (blockIdx.x == 0 && (g) < 64) g_ap_trace[(g) * 16 + (slot)] = clock64(); } while (0)
#else
#define AP_TRACE(g, slot) do { } while (0)
#endif

struct AttnPipeParams {
    const __nv_bfloat16* qkv;
    const int32_t* row_map;
    int N_src, Np, Np_pad, C, H, BH, tpi;      // tpi = 128-row tiles per (image, head) item
    int nbuf, s_stride, n_obuf, o_col;         // score buffer b at column b*s_stride; O buffer ob at o_col + 64*ob
    int split;                                  // key columns [0, split) -> exp warps 0-3, [split, Np_pad) -> warps 4-7
    int plane_bytes, stages, reverse;
    float scale_log2;
};

__device__ __forceinline__ void ap_cp_async16(uint32_t smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ap_cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ap_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ap_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ap_fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// Register re-partitioning between warpgroups (4 consecutive warps): the load/MMA warpgroup gives registers back, the
// helper warpgroup (two 32-column TMEM loads in flight + its tile cursors) takes them.  8*128 + 4*184 + 4*72 = 2048 = 64 K / 32.
template <int N> __device__ __forceinline__ void ap_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N)); }
template <int N> __device__ __forceinline__ void ap_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N)); }

// warp-uniform non-blocking barrier probe (lane 0 tests, everybody gets its answer)
__device__ __forceinline__ bool ap_test_uniform(uint64_t* bar, uint32_t parity) {
    return __shfl_sync(0xffffffffu, (int)mbar_test(bar, parity), 0) != 0;
}

__device__ __forceinline__ int ap_item(const AttnPipeParams& p, int n) {
    const int it = (int)blockIdx.x + n * (int)gridDim.x;
    return p.reverse ? p.BH - 1 - it : it;
}

// position of a tile in the CTA's sequence, kept with counters instead of divisions
struct ApCursor {
    int g = 0, n = 0, j = 0;          // tile index, item index, tile inside the item
    int buf = 0, stage = 0;
    uint32_t buf_ph = 0, stage_ph = 0;
    __device__ __forceinline__ void advance(const AttnPipeParams& p) {
        ++g;
        if (++buf == p.nbuf) { buf = 0; buf_ph ^= 1; }
        if (++j == p.tpi) {
            j = 0;
            ++n;
            if (++stage == p.stages) { stage = 0; stage_ph ^= 1; }
        }
    }
};

__global__ void __launch_bounds__(kApThreads, 1)
attention_pipe_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out, const AttnPipeParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int stage_bytes = 3 * p.plane_bytes;
    // [stages][Q|K|V planes] [output staging: 4 warps x 4 KB] [row maxima][partial row sums][barriers]
    // (a tile's Q operand is read as 128 rows from row 0 / 128 of the Q plane: the over-read lands in the stage's K plane)
    const uint32_t out_stage0 = smem_base + p.stages * stage_bytes;
    float* mrow = reinterpret_cast<float*>(smem_gen + p.stages * stage_bytes + kApOutStage);     // [nbuf][128]  max * scale * log2e
    float* sums = mrow + kApMaxBufs * 128;                                                         // [8 slots][2 halves][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sums + kApSumSlots * 2 * 128);
    uint64_t* qk_full = bars;                        // [4] loader -> MMA
    uint64_t* v_full = bars + 4;                     // [4] loader -> MMA
    uint64_t* stage_empty = bars + 8;                // [4] MMA -> loader (tcgen05.commit after the item's last P V)
    uint64_t* s_full = bars + 12;                    // [4 bufs] MMA -> helpers (S ready)
    uint64_t* m_ready = bars + 16;                   // [4 bufs] helpers -> exp warps (row maxima in shared memory)
    uint64_t* p_full = bars + 20;                    // [4 bufs] exp warps -> MMA (P in TMEM)
    uint64_t* o_full = bars + 24;                    // [2] MMA -> helpers
    uint64_t* o_empty = bars + 26;                   // [2] helpers -> MMA (O read out)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == kApMmaWarp) {
        tmem_alloc(tmem_slot, 512);
        if (lane == 0) {
            for (int i = 0; i < kApMaxStages; ++i) {
                // gathered: one cp.async-completion arrival per loader thread; dense: one arrival + TMA transaction bytes
                mbar_init(&qk_full[i], p.row_map ? kApLoaderThreads : 1);
                mbar_init(&v_full[i], p.row_map ? kApLoaderThreads : 1);
                mbar_init(&stage_empty[i], 1);
            }
            for (int i = 0; i < kApMaxBufs; ++i) {
                mbar_init(&s_full[i], 1);
                mbar_init(&m_ready[i], 4);           // one arrival per helper warp
                mbar_init(&p_full[i], 8);            // one arrival per exp warp
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&o_full[i], 1);
                mbar_init(&o_empty[i], 4);
            }
            mbar_fence_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();
    griddep_wait();
    const int n_mine = (p.BH - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // items of this CTA
    const int G = n_mine * p.tpi;                                                          // tiles of this CTA
    const int Np = p.Np, Np_pad = p.Np_pad;

    if (warp >= kApMmaWarp) {
      // warpgroup 3 (MMA issuer + loaders) hands registers back: ONE setmaxnreg site for its four warps
      ap_reg_dec<72>();
      if (warp >= kApLoaderWarp0 && p.row_map == nullptr) {
        // ================= dense loader: one TMA box per plane; rows past the image's N_src are zero-filled =================
        if (tid == kApLoaderWarp0 * 32) {
            tma_prefetch_desc(&tmap_qkv);
            const uint32_t plane_tx = (uint32_t)Np_pad * 128u;
            int stage = 0;
            uint32_t phase = 0;
            for (int n = 0; n < n_mine; ++n) {
                const int item = ap_item(p, n);
                const int b = item / p.H, h = item - b * p.H;
                mbar_wait(&stage_empty[stage], phase ^ 1);
                mbar_expect_tx(&qk_full[stage], 2u * plane_tx);
                mbar_expect_tx(&v_full[stage], plane_tx);
                uint8_t* sq = smem_gen + stage * stage_bytes;
                tma_load_3d(sq, &tmap_qkv, &qk_full[stage], h * 64, 0, b);
                tma_load_3d(sq + p.plane_bytes, &tmap_qkv, &qk_full[stage], p.C + h * 64, 0, b);
                tma_load_3d(sq + 2 * p.plane_bytes, &tmap_qkv, &v_full[stage], 2 * p.C + h * 64, 0, b);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
      } else if (warp >= kApLoaderWarp0) {
        // ================= gather loaders: 8 lanes move one token's 128-byte head slice per plane (cp.async) =================
        const int lt = tid - kApLoaderWarp0 * 32;
        const int grp = lt >> 3, chunk = lt & 7;
        const long long C3 = 3LL * p.C;
        int stage = 0;
        uint32_t phase = 0;
        for (int n = 0; n < n_mine; ++n) {
            const int item = ap_item(p, n);
            const int b = item / p.H, h = item - b * p.H;
            // global row of every token this lane group moves: the index loads are issued together, before the wait
            int grow[kApSweeps];
#pragma unroll
            for (int i = 0; i < kApSweeps; ++i) {
                const int j = grp + i * kApLoaderGroups;
                grow[i] = j < Np ? __ldg(p.row_map + (long long)b * Np + j) : -1;
            }
            mbar_wait(&stage_empty[stage], phase ^ 1);
            const uint32_t sq = smem_base + stage * stage_bytes, sk = sq + p.plane_bytes, sv = sk + p.plane_bytes;
            const __nv_bfloat16* base = p.qkv + h * 64 + chunk * 8;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {                    // pass 0: Q and K (what S needs), pass 1: V
#pragma unroll
                for (int i = 0; i < kApSweeps; ++i) {
                    const int j = grp + i * kApLoaderGroups;
                    if (j >= Np_pad) continue;
                    const bool ok = grow[i] >= 0;
                    const __nv_bfloat16* src = base + (long long)(ok ? grow[i] : 0) * C3;
                    const uint32_t off = j * 128 + ((chunk ^ (j & 7)) << 4);
                    if (pass == 0) {
                        if (ok) {                                     // rows past Np: dead query rows / masked key columns
                            ap_cp_async16(sq + off, src, 16);
                            ap_cp_async16(sk + off, src + p.C, 16);
                        }
                    } else {
                        ap_cp_async16(sv + off, src + 2 * p.C, ok ? 16 : 0);      // zero-filled: 0 * V must stay 0
                    }
                }
                ap_cp_async_arrive(pass == 0 ? &qk_full[stage] : &v_full[stage]);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      } else if (warp == kApMmaWarp) {
        // ================= MMA issuer =================
        // The tensor pipe executes in issue order:  ... PV(g-1) S(g+1) PV(g) S(g+2) ...  S(g) may be issued once
        // PV(g-nbuf), which read P out of the same buffer, is queued; PV(g) once P(g) is complete and its O buffer is free.
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, Np_pad, 0, 0);
            const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);          // B = V is MN-major
            const int nk = Np_pad / 16;
            const uint64_t qd0 = umma_desc_sw128(smem_base, 16, 1024);
            const uint64_t kd0 = umma_desc_sw128(smem_base + p.plane_bytes, 16, 1024);
            const uint64_t vd0 = umma_desc_sw128(smem_base + 2 * p.plane_bytes, 16, 1024);
            ApCursor s, v;
            int ob = 0, o_use0 = 0, o_use1 = 0;                               // O buffer of the next PV; how often each has been used
            while (v.g < G) {
                bool did = false;
                if (s.g < G && s.g - v.g < p.nbuf && mbar_test(&qk_full[s.stage], s.stage_ph)) {
                    tc_fence_after();
                    AP_TRACE(s.g, 0);
                    const uint64_t qd = qd0 + (uint64_t)((s.stage * stage_bytes + s.j * (128 * 128)) >> 4);
                    const uint64_t kd = kd0 + (uint64_t)((s.stage * stage_bytes) >> 4);
                    const uint32_t d = tmem_base + s.buf * p.s_stride;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_s, k != 0);
                    umma_commit(&s_full[s.buf]);
                    AP_TRACE(s.g, 1);
                    s.advance(p);
                    did = true;
                }
                const int o_use = ob ? o_use1 : o_use0;
                if (v.g < s.g && mbar_test(&p_full[v.buf], v.buf_ph) && mbar_test(&v_full[v.stage], v.stage_ph) &&
                    (o_use == 0 || mbar_test(&o_empty[ob], (o_use - 1) & 1))) {
                    tc_fence_after();
                    AP_TRACE(v.g, 2);
                    const uint64_t vd = vd0 + (uint64_t)((v.stage * stage_bytes) >> 4);
                    const uint32_t pb = tmem_base + v.buf * p.s_stride;
                    const uint32_t d = tmem_base + p.o_col + ob * 64;
                    for (int k = 0; k < nk; ++k) {
                        const int key0 = k * 16;
                        const uint32_t a = pb + (key0 < p.split ? (key0 >> 1) : p.split + ((key0 - p.split) >> 1));
                        umma_bf16_ts(d, a, vd + (uint64_t)(k * (2048 >> 4)), idesc_o, k != 0);
                    }
                    umma_commit(&o_full[ob]);
                    if (v.j == p.tpi - 1) umma_commit(&stage_empty[v.stage]);     // the item's last product: the stage may be refilled
                    AP_TRACE(v.g, 3);
                    if (ob) ++o_use1; else ++o_use0;
                    if (++ob == p.n_obuf) ob = 0;
                    v.advance(p);
                    did = true;
                }
                if (!did) __nanosleep(20);
            }
        }
      }
    } else if (warp >= kApHelpWarp0) {
        // ================= helpers: row maxima one tile ahead of the exp warps, and the O epilogue =================
        ap_reg_inc<184>();
        const int hq = warp - kApHelpWarp0;                                  // TMEM lane quadrant
        const int row = hq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(hq * 32) << 16);
        const float sl2 = p.scale_log2;
        const uint32_t s_out = out_stage0 + hq * 4096;
        ApCursor m, e;                                                        // next tile whose maxima / whose O is due
        int ob = 0;
        uint32_t o_ph0 = 0, o_ph1 = 0;
        int slot = 0;
        while (e.g < G) {
            if (m.g < G && ap_test_uniform(&s_full[m.buf], m.buf_ph)) {
                // ---- (a) exact row maximum of S(m.g)
                tc_fence_after();
                if (hq == 0 && lane == 0) AP_TRACE(m.g, 4);
                if (m.j * 128 + hq * 32 < Np) {
                    const uint32_t sb = lane_base + m.buf * p.s_stride;
                    uint32_t va[32], vb[32];
                    float mx = -INFINITY;
                    auto max32 = [&](const uint32_t (&cur)[32], int c0) {
                        if (c0 + 32 <= Np) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) mx = ap_fmax3(mx, __uint_as_float(cur[j]), __uint_as_float(cur[j + 1]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (c0 + j < Np) mx = fmaxf(mx, __uint_as_float(cur[j]));
                        }
                    };
                    for (int c0 = 0; c0 < Np; c0 += 64) {
                        // two loads in flight, both unconditional (the second may read past Np: still inside the 512 columns,
                        // masked in max32)
                        tmem_ld32(sb + c0, va);
                        tmem_ld32(sb + c0 + 32, vb);
                        tmem_ld_wait();
                        max32(va, c0);
                        max32(vb, c0 + 32);
                    }
                    mrow[m.buf * 128 + row] = mx * sl2;
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&m_ready[m.buf]);
                if (hq == 0 && lane == 0) AP_TRACE(m.g, 5);
                m.advance(p);
                continue;
            }
            if (e.g < m.g && ap_test_uniform(&o_full[ob], ob ? o_ph1 : o_ph0)) {
                // ---- (b) O(e.g) = P V is complete: normalise, store
                tc_fence_after();
                if (hq == 0 && lane == 0) AP_TRACE(e.g, 6);
                const bool live = e.j * 128 + hq * 32 < Np;
                uint32_t o0[32], o1[32];
                if (live) {
                    const uint32_t ocol = lane_base + p.o_col + ob * 64;
                    tmem_ld32(ocol, o0);
                    tmem_ld32(ocol + 32, o1);
                    tmem_ld_wait();
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[ob]);
                if (live) {
                    const float inv = 1.f / (sums[(slot * 2 + 0) * 128 + row] + sums[(slot * 2 + 1) * 128 + row]);
                    if (lane == 0) bulk_wait_group_read<0>();                 // the previous store has read the staging buffer
                    __syncwarp();
                    const uint32_t srow = s_out + lane * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t (&src)[32] = c < 4 ? o0 : o1;
                        const int j = (c & 3) * 8;
                        ap_sts128(srow + ((c ^ (lane & 7)) << 4),
                                  float2_to_bf16x2(__uint_as_float(src[j]) * inv, __uint_as_float(src[j + 1]) * inv),
                                  float2_to_bf16x2(__uint_as_float(src[j + 2]) * inv, __uint_as_float(src[j + 3]) * inv),
                                  float2_to_bf16x2(__uint_as_float(src[j + 4]) * inv, __uint_as_float(src[j + 5]) * inv),
                                  float2_to_bf16x2(__uint_as_float(src[j + 6]) * inv, __uint_as_float(src[j + 7]) * inv));
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int item = ap_item(p, e.n);
                        const int b = item / p.H, h = item - b * p.H;
                        tma_store_3d(&tmap_out, s_out, h * 64, e.j * 128 + hq * 32, b);
                        bulk_commit_group();
                    }
                }
                if (hq == 0 && lane == 0) AP_TRACE(e.g, 7);
                if (ob) o_ph1 ^= 1; else o_ph0 ^= 1;
                if (++ob == p.n_obuf) ob = 0;
                if (++slot == kApSumSlots) slot = 0;
                e.advance(p);
                continue;
            }
            __nanosleep(20);
        }
        if (lane == 0) bulk_wait_group<0>();                                  // output stores complete before the CTA retires
    } else {
        // ================= exp warps: thread = (query row, half of the key columns) =================
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane;
        const int cb = half ? p.split : 0, ce = half ? Np_pad : p.split;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const float sl2 = p.scale_log2;
        ApCursor c;
        int slot = 0;
        for (; c.g < G; c.advance(p)) {
            mbar_wait(&m_ready[c.buf], c.buf_ph);
            if (warp == 0 && lane == 0) AP_TRACE(c.g, 8);
            if (c.j * 128 + q * 32 < Np) {
                const float mb = mrow[c.buf * 128 + row];
                tc_fence_after();
                const uint32_t sb = lane_base + c.buf * p.s_stride;
                // P chunk of S columns [c0, c0+16) lands on columns cb + (c0-cb)/2 ..+8: behind this thread's read pointer
                uint32_t va[16], vb[16];
                float sum0 = 0.f, sum1 = 0.f;
                auto exp16 = [&](const uint32_t (&cur)[16], uint32_t (&nxt)[16], int c0) {
                    tmem_ld_wait();
                    if (c0 + 16 < ce) tmem_ld16(sb + c0 + 16, nxt);
                    uint32_t pk[8];
                    if (c0 + 16 <= Np) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float e0 = ap_ex2(fmaf(__uint_as_float(cur[j]), sl2, -mb));
                            const float e1 = ap_ex2(fmaf(__uint_as_float(cur[j + 1]), sl2, -mb));
                            sum0 += e0;
                            sum1 += e1;
                            pk[j >> 1] = float2_to_bf16x2(e0, e1);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float e0 = (c0 + j < Np) ? ap_ex2(fmaf(__uint_as_float(cur[j]), sl2, -mb)) : 0.f;
                            const float e1 = (c0 + j + 1 < Np) ? ap_ex2(fmaf(__uint_as_float(cur[j + 1]), sl2, -mb)) : 0.f;
                            sum0 += e0;
                            sum1 += e1;
                            pk[j >> 1] = float2_to_bf16x2(e0, e1);
                        }
                    }
                    tmem_st8(sb + cb + ((c0 - cb) >> 1), pk);
                };
                if (cb < ce) {
                    tmem_ld16(sb + cb, va);
                    for (int c0 = cb; c0 < ce; c0 += 32) {
                        exp16(va, vb, c0);
                        if (c0 + 16 < ce) exp16(vb, va, c0 + 16);
                    }
                }
                sums[(slot * 2 + half) * 128 + row] = sum0 + sum1;
                tmem_st_wait();
                tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[c.buf]);
            if (warp == 0 && lane == 0) AP_TRACE(c.g, 9);
            if (++slot == kApSumSlots) slot = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kApMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int make_tmap_bf16_3d_box(CUtensorMap* map, const void* base, long long batch, long long rows, long long cols, int box_rows);   // gemm_tcgen05.cu
int make_tmap_bf16_3d_ld(CUtensorMap* map, const void* base, long long batch, long long rows, long long cols, int box_rows);

static int ap_num_sms() {
    static int n_dev[kMaxDevices] = {};
    int& n = n_dev[current_device()];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
        if (n <= 0) n = 148;
    }
    return n;
}

// returns 1 if this kernel handled the call, 0 if the shape is outside its range, <0 on error
int launch_attention_pipe(const void* qkv, const int32_t* row_map, void* out, int B, int N_src, int Np,
                          int C, int H, float scale, int reverse, cudaStream_t stream) {
    const int Np_pad = (Np + 15) & ~15;
    if (Np_pad > 224) return 0;
    AttnPipeParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv);
    p.row_map = row_map;
    p.N_src = N_src; p.Np = Np; p.Np_pad = Np_pad; p.C = C; p.H = H;
    p.reverse = reverse;
    p.BH = B * H;
    p.tpi = Np > 128 ? 2 : 1;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.s_stride = (Np_pad + 31) & ~31;
    p.n_obuf = 2 * p.s_stride + 128 <= 512 ? 2 : 1;
    p.nbuf = (512 - 64 * p.n_obuf) / p.s_stride;
    if (p.nbuf > kApMaxBufs) p.nbuf = kApMaxBufs;
    p.o_col = 512 - 64 * p.n_obuf;
    p.split = ((Np_pad / 2) + 15) & ~15;
    p.plane_bytes = (Np_pad * 128 + 1023) & ~1023;
    p.stages = kApSmemBudget / (3 * p.plane_bytes);
    if (p.stages > kApMaxStages) p.stages = kApMaxStages;
    RAJNI_REQUIRE(p.stages >= 2 && p.nbuf >= 2, RAJNI_EINVAL, "attention_pipe: Np=%d leaves %d stage(s), %d score buffer(s)", Np, p.stages, p.nbuf);
    // the last stage's Q operand of tile 1 is read 128 rows deep from row 128: keep that inside the allocation
    const int smem = p.stages * 3 * p.plane_bytes + kApOutStage + kApAuxBytes + kApBarBytes + 1024;
    CUtensorMap tmap;
    if (row_map) {
        if (int rc = make_tmap_bf16_3d_box(&tmap, out, B, Np, C, 32)) return rc;        // placeholder (unused by the gather loaders)
    } else {
        if (int rc = make_tmap_bf16_3d_ld(&tmap, qkv, B, N_src, 3LL * C, Np_pad)) return rc;
    }
    CUtensorMap tmap_out;
    RAJNI_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, RAJNI_EINVAL, "attention_pipe: out must be 16-byte aligned (TMA store)");
    if (int rc = make_tmap_bf16_3d_box(&tmap_out, out, B, Np, C, 32)) return rc;
    static int attr_smem_dev[kMaxDevices] = {};
    int& attr_smem = attr_smem_dev[current_device()];
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(attention_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "attention_pipe: smem attribute (%d B): %s", smem, cudaGetErrorString(e));
        attr_smem = smem;
    }
    int grid = p.BH < ap_num_sms() ? p.BH : ap_num_sms();
    static const int cta_cap = getenv("RAJNI_ATTN_MAX_CTAS") ? atoi(getenv("RAJNI_ATTN_MAX_CTAS")) : 0;      // experiments: share the GPU
    if (cta_cap > 0 && grid > cta_cap) grid = cta_cap;
    cudaError_t le = launch_kernel(attention_pipe_kernel, dim3(grid), dim3(kApThreads), (size_t)smem, stream, 1, tmap, tmap_out, p);
    count_launch();
    RAJNI_REQUIRE(le == cudaSuccess, RAJNI_ECUDA, "attention_pipe: launch failed: %s", cudaGetErrorString(le));
    int rc = check_launch("attention_pipe");
    return rc ? rc : 1;
}

}  // namespace rajni
