// HBM-bound row kernels: gather-compaction, LayerNorm, patch im2col.
// All use 16-byte vector accesses, contiguous along the fastest dimension per warp.
#include "common.cuh"

namespace rajni {

// ------------------------------------------------------------------ gather rows
// dst[r,:] = src[row_map[r],:]   attention.py:42-43 (qkv rows), model.py:55-56 (residual rows)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src,
                                                          const int32_t* __restrict__ row_map,
                                                          uint4* __restrict__ dst,
                                                          long long total_chunks, int chunks_per_row) {
    griddep_launch();
    griddep_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent 16-byte loads in flight per thread
    for (; i + 3 * stride < total_chunks; i += 4 * stride) {
        uint4 v[4];
        long long idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            idx[u] = i + u * stride;
            int r = (int)(idx[u] / chunks_per_row);
            int c = (int)(idx[u] - (long long)r * chunks_per_row);
            v[u] = ld_stream16(src + (long long)__ldg(row_map + r) * chunks_per_row + c);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) st_stream16(dst + idx[u], v[u]);
    }
    for (; i < total_chunks; i += stride) {
        int r = (int)(i / chunks_per_row);
        int c = (int)(i - (long long)r * chunks_per_row);
        st_stream16(dst + i, ld_stream16(src + (long long)__ldg(row_map + r) * chunks_per_row + c));
    }
}

// ------------------------------------------------------------------ LayerNorm
// One warp per row, the row held in registers (CPL 16-byte chunks per lane), fp32 math:
// mean, then centred variance (biased, like nn.LayerNorm), y = (x-mean)*rstd*gamma+beta.
template <int CPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long in_stride,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ y, int rows, int C) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int chunks = C >> 3;
    const __nv_bfloat16* xr = x + (long long)row * in_stride;
    float v[CPL][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        int j = lane + 32 * i;
        if (j < chunks) {
            uint4 u = ld_stream16(xr + j * 8);
            float2 a = bf16x2_to_float2(u.x), b = bf16x2_to_float2(u.y), c = bf16x2_to_float2(u.z), d = bf16x2_to_float2(u.w);
            v[i][0] = a.x; v[i][1] = a.y; v[i][2] = b.x; v[i][3] = b.y;
            v[i][4] = c.x; v[i][5] = c.y; v[i][6] = d.x; v[i][7] = d.y;
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[i][e];
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
        }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        if (lane + 32 * i < chunks) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { float d = v[i][e] - mean; sq += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
    __nv_bfloat16* yr = y + (long long)row * C;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        int j = lane + 32 * i;
        if (j < chunks) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + j * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + j * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + j * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + j * 8 + 4));
            uint4 o;
            o.x = float2_to_bf16x2((v[i][0] - mean) * rstd * g0.x + b0.x, (v[i][1] - mean) * rstd * g0.y + b0.y);
            o.y = float2_to_bf16x2((v[i][2] - mean) * rstd * g0.z + b0.z, (v[i][3] - mean) * rstd * g0.w + b0.w);
            o.z = float2_to_bf16x2((v[i][4] - mean) * rstd * g1.x + b1.x, (v[i][5] - mean) * rstd * g1.y + b1.y);
            o.w = float2_to_bf16x2((v[i][6] - mean) * rstd * g1.z + b1.z, (v[i][7] - mean) * rstd * g1.w + b1.w);
            st_stream16(yr + j * 8, o);
        }
    }
}

// ------------------------------------------------------------------ patch im2col (+ CLS rows)
// One thread moves one (patch, channel, ky) strip: 16 pixels in, 16 bf16 out (two 16-byte
// stores = one full 32-byte sector).  px is the fastest thread index, so a warp reads
// contiguous image rows.  Column order (c, ky, kx) matches Conv2d weight.view(C, -1).
// DT: 0 = bf16, 1 = fp32, 2 = uint8 pixels normalised here exactly as torchvision's ToTensor + Normalize do in fp32
// ((u / 255 - mean[c]) / std[c], IEEE divisions; run.py:62-70), so the bf16 result equals that of the fp32 pipeline.
struct ImageNorm { float mean[3], std[3]; };
template <int DT>
__global__ void __launch_bounds__(256) im2col16_kernel(const void* __restrict__ images, const ImageNorm norm, int B, int S,
                                                       __nv_bfloat16* __restrict__ cols,
                                                       const uint4* __restrict__ cls_pos0,
                                                       uint4* __restrict__ x, int C,
                                                       float2* __restrict__ row_stats, long long stats_ld, int stats_slots,
                                                       float cls_sum, float cls_sumsq) {
    griddep_launch();
    griddep_wait();
    const int G = S >> 4;                       // patches per side
    const long long strips = (long long)B * G * 3 * 16 * G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < strips; i += stride) {
        int px = (int)(i % G);
        long long t = i / G;
        int ky = (int)(t & 15); t >>= 4;
        int c = (int)(t % 3); t /= 3;
        int py = (int)(t % G);
        int b = (int)(t / G);
        const long long pix = (((long long)b * 3 + c) * S + (py * 16 + ky)) * S + px * 16;
        uint4 o0, o1;
        if (DT == 2) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(images) + pix));
            const float mean = norm.mean[c], sd = norm.std[c];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    f[j] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)((w[k] >> (8 * j)) & 0xffu), 255.0f), mean), sd);
                o[2 * k] = float2_to_bf16x2(f[0], f[1]);
                o[2 * k + 1] = float2_to_bf16x2(f[2], f[3]);
            }
            o0 = make_uint4(o[0], o[1], o[2], o[3]);
            o1 = make_uint4(o[4], o[5], o[6], o[7]);
        } else if (DT == 1) {
            const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(images) + pix);
            float4 a = __ldg(src), bb = __ldg(src + 1), cc = __ldg(src + 2), d = __ldg(src + 3);
            o0 = make_uint4(float2_to_bf16x2(a.x, a.y), float2_to_bf16x2(a.z, a.w),
                            float2_to_bf16x2(bb.x, bb.y), float2_to_bf16x2(bb.z, bb.w));
            o1 = make_uint4(float2_to_bf16x2(cc.x, cc.y), float2_to_bf16x2(cc.z, cc.w),
                            float2_to_bf16x2(d.x, d.y), float2_to_bf16x2(d.z, d.w));
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(images) + pix);
            o0 = __ldg(src);
            o1 = __ldg(src + 1);
        }
        const long long m = ((long long)b * G + py) * G + px;
        uint4* dst = reinterpret_cast<uint4*>(cols + m * 768 + c * 256 + ky * 16);
        dst[0] = o0;
        dst[1] = o1;
    }
    // CLS rows: x[b,0,:] = cls_token + pos_embed[0]   (model.py:35-37)
    const int cchunks = C >> 3;
    const long long P1 = (long long)G * G + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * cchunks; i += stride) {
        int b = (int)(i / cchunks), j = (int)(i % cchunks);
        x[(long long)b * P1 * cchunks + j] = __ldg(cls_pos0 + j);
    }
    // LayerNorm partials of the CLS rows (the patch rows' partials come from the embed GEMM's epilogue)
    if (row_stats != nullptr) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * stats_slots; i += stride) {
            int b = (int)(i / stats_slots), s = (int)(i % stats_slots);
            row_stats[(long long)s * stats_ld + (long long)b * P1] = s == 0 ? make_float2(cls_sum, cls_sumsq) : make_float2(0.f, 0.f);
        }
    }
}

}  // namespace rajni

using namespace rajni;

extern "C" int rajni_gather_rows(const void* src, const int32_t* row_map, void* dst,
                                 int rows_out, int row_elems, void* stream) {
    RAJNI_REQUIRE(src && row_map && dst, RAJNI_EINVAL, "rajni_gather_rows: null pointer");
    RAJNI_REQUIRE(rows_out > 0 && row_elems > 0 && row_elems % 8 == 0, RAJNI_EINVAL,
                  "rajni_gather_rows: rows=%d row_elems=%d (must be a positive multiple of 8)", rows_out, row_elems);
    const int cpr = row_elems / 8;
    const long long total = (long long)rows_out * cpr;
    long long blocks = (total + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaError_t e = launch_kernel(gather_rows_kernel, dim3((int)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), 1,
                                  static_cast<const uint4*>(src), row_map, static_cast<uint4*>(dst), total, cpr);
    count_launch();
    RAJNI_REQUIRE(e == cudaSuccess, RAJNI_ECUDA, "gather_rows: launch failed: %s", cudaGetErrorString(e));
    return check_launch("gather_rows");
}

extern "C" int rajni_layernorm(const void* x, long long in_row_stride, const float* gamma,
                               const float* beta, float eps, void* y, int rows, int C, void* stream) {
    RAJNI_REQUIRE(x && gamma && beta && y, RAJNI_EINVAL, "rajni_layernorm: null pointer");
    RAJNI_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 2048 && in_row_stride % 8 == 0, RAJNI_EINVAL,
                  "rajni_layernorm: rows=%d C=%d stride=%lld unsupported", rows, C, in_row_stride);
    const int cpl = (C / 8 + 31) / 32;
    const int wpb = 8;
    dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
    auto s = static_cast<cudaStream_t>(stream);
    auto xb = static_cast<const __nv_bfloat16*>(x);
    auto yb = static_cast<__nv_bfloat16*>(y);
    cudaError_t le = cudaSuccess;
    switch (cpl) {
        case 1: le = launch_kernel(layernorm_kernel<1>, grid, block, 0, s, 1, xb, in_row_stride, gamma, beta, eps, yb, rows, C); break;
        case 2: le = launch_kernel(layernorm_kernel<2>, grid, block, 0, s, 1, xb, in_row_stride, gamma, beta, eps, yb, rows, C); break;
        case 3: le = launch_kernel(layernorm_kernel<3>, grid, block, 0, s, 1, xb, in_row_stride, gamma, beta, eps, yb, rows, C); break;
        case 4: le = launch_kernel(layernorm_kernel<4>, grid, block, 0, s, 1, xb, in_row_stride, gamma, beta, eps, yb, rows, C); break;
        default: le = launch_kernel(layernorm_kernel<8>, grid, block, 0, s, 1, xb, in_row_stride, gamma, beta, eps, yb, rows, C); break;
    }
    count_launch();
    RAJNI_REQUIRE(le == cudaSuccess, RAJNI_ECUDA, "layernorm: launch failed: %s", cudaGetErrorString(le));
    return check_launch("layernorm");
}

extern "C" int rajni_patch_im2col(const void* images, int image_dtype, int B, int S, int patch,
                                  void* cols, const void* cls_pos0, void* x, int C,
                                  float* row_stats, long long row_stats_ld, int stats_slots,
                                  float cls_sum, float cls_sumsq, const float* norm, void* stream) {
    RAJNI_REQUIRE(images && cols && cls_pos0 && x, RAJNI_EINVAL, "rajni_patch_im2col: null pointer");
    RAJNI_REQUIRE(image_dtype >= RAJNI_IMG_BF16 && image_dtype <= RAJNI_IMG_U8, RAJNI_EINVAL, "rajni_patch_im2col: image_dtype %d", image_dtype);
    RAJNI_REQUIRE(image_dtype != RAJNI_IMG_U8 || norm != nullptr, RAJNI_EINVAL, "rajni_patch_im2col: uint8 images need norm (mean[3], std[3])");
    ImageNorm nm{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
    if (image_dtype == RAJNI_IMG_U8)
        for (int i = 0; i < 3; ++i) {
            nm.mean[i] = norm[i];
            nm.std[i] = norm[3 + i];
            RAJNI_REQUIRE(nm.std[i] > 0.f, RAJNI_EINVAL, "rajni_patch_im2col: std[%d] must be positive", i);
        }
    RAJNI_REQUIRE(patch == 16 && S > 0 && S % 16 == 0 && B > 0 && C % 8 == 0, RAJNI_EINVAL,
                  "rajni_patch_im2col: patch=%d S=%d B=%d C=%d unsupported (patch must be 16)", patch, S, B, C);
    RAJNI_REQUIRE(row_stats == nullptr || (stats_slots > 0 && row_stats_ld >= (long long)B * ((S / 16) * (S / 16) + 1)), RAJNI_EINVAL,
                  "rajni_patch_im2col: row_stats needs stats_slots > 0 and row_stats_ld >= B*(P+1)");
    const int G = S / 16;
    const long long strips = (long long)B * G * 3 * 16 * G;
    long long blocks = (strips + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    auto s = static_cast<cudaStream_t>(stream);
    auto kern = image_dtype == RAJNI_IMG_U8 ? im2col16_kernel<2> : image_dtype == RAJNI_IMG_F32 ? im2col16_kernel<1> : im2col16_kernel<0>;
    cudaError_t le = launch_kernel(kern, dim3((int)blocks), dim3(256), 0, s, 1,
                                   images, nm, B, S, static_cast<__nv_bfloat16*>(cols), static_cast<const uint4*>(cls_pos0),
                                   static_cast<uint4*>(x), C, reinterpret_cast<float2*>(row_stats), row_stats_ld, stats_slots,
                                   cls_sum, cls_sumsq);
    count_launch();
    RAJNI_REQUIRE(le == cudaSuccess, RAJNI_ECUDA, "patch_im2col: launch failed: %s", cudaGetErrorString(le));
    return check_launch("patch_im2col");
}
