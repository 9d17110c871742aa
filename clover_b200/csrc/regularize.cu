// Stochastic regularisers of the Clover training path (sm_100a).
//
//   clv_dropout        y = residual + x * keep / (1 - p)      nn.Dropout of HF BertEmbeddings / BertSelfOutput /
//                                                              BertOutput (transformers 4.6.1; call sites
//                                                              bert_from_hugface.py:30, cross_transformer.py:110) and
//                                                              of the heads (ssl_head.py:108-109,211-212,292-293;
//                                                              qa_head.py:12,60).  The backward pass is the same
//                                                              kernel applied to dy (the mask is regenerated).
//   clv_rows_scale     y[r,:] = x[r,:] * scale[r / rows_per_group]   timm DropPath of SwinTransformerBlock3D
//                                                              (swin_transformer_3d.py:499,503) on the gradient
//                                                              side; the forward side is the GEMM epilogue row scale.
//   clv_keep_mask      the keep decisions themselves (uint8), for tests and for inspection.
//
// The random stream is counter based: element i of a call draws keep = hash(seed, offset + i) >= p * 2^32 with the
// splitmix64 finaliser (common.cuh: rand_u32).  No state lives on the device; the host hands out disjoint offsets,
// so forward and backward (and the attention kernels, which use the same function on the (b, h, i, j) index)
// regenerate identical masks.
#include "common.cuh"
#include "clover_b200.h"

namespace clv {

CLV_DEVICE float4 ld4(const void* p, int is_bf16, long long i) {
  if (is_bf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
}
CLV_DEVICE void st4(void* p, int is_bf16, long long i, float4 v) {
  if (is_bf16)
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  else
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
}

__global__ void __launch_bounds__(256) dropout_kernel(const void* x, int x_bf16, const void* res, int res_bf16, void* y,
                                                      int y_bf16, long long n4, uint32_t thresh, float inv_keep,
                                                      unsigned long long seed, unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ld4(x, x_bf16, i * 4);
    const unsigned long long e = offset + (unsigned long long)i * 4;
    v.x *= keep_scale(seed, e, thresh, inv_keep);
    v.y *= keep_scale(seed, e + 1, thresh, inv_keep);
    v.z *= keep_scale(seed, e + 2, thresh, inv_keep);
    v.w *= keep_scale(seed, e + 3, thresh, inv_keep);
    if (res) {
      const float4 r = ld4(res, res_bf16, i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    st4(y, y_bf16, i * 4, v);
  }
}

__global__ void __launch_bounds__(256) keep_mask_kernel(unsigned char* out, long long n, uint32_t thresh,
                                                        unsigned long long seed, unsigned long long offset) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = rand_u32(seed, offset + (unsigned long long)i) >= thresh ? 1 : 0;
}

__global__ void __launch_bounds__(256) rows_scale_kernel(const void* x, int x_bf16, void* y, int y_bf16, long long rows,
                                                         int nvec, const float* scale, long long rows_per_group) {
  const long long total = rows * nvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / nvec;
    const float s = __ldg(scale + r / rows_per_group);
    float4 v = ld4(x, x_bf16, i * 4);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    st4(y, y_bf16, i * 4, v);
  }
}

static int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace clv

using namespace clv;

extern "C" unsigned int clv_dropout_threshold(float p) { return drop_threshold(p); }
extern "C" unsigned int clv_rand_u32(unsigned long long seed, unsigned long long idx) { return rand_u32(seed, idx); }

extern "C" int clv_dropout(const void* x, int x_is_bf16, const void* residual, int residual_is_bf16, void* y, int y_is_bf16,
                           long long n, float p, unsigned long long seed, unsigned long long offset, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && y && n >= 0 && n % 4 == 0, "clv_dropout: n must be a multiple of 4 (got %lld)", n);
  CLV_REQUIRE(p >= 0.f && p < 1.f, "clv_dropout: p must be in [0, 1) (got %f)", (double)p);
  if (n == 0) return 0;
  dropout_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(x, x_is_bf16, residual, residual_is_bf16, y, y_is_bf16, n / 4,
                                                          drop_threshold(p), 1.0f / (1.0f - p), seed, offset);
  return after_launch("dropout_kernel");
}

extern "C" int clv_keep_mask(unsigned char* out, long long n, float p, unsigned long long seed, unsigned long long offset,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(out && n >= 0 && p >= 0.f && p < 1.f, "clv_keep_mask: bad arguments");
  if (n == 0) return 0;
  keep_mask_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, n, drop_threshold(p), seed, offset);
  return after_launch("keep_mask_kernel");
}

extern "C" int clv_rows_scale(const void* x, int x_is_bf16, void* y, int y_is_bf16, long long rows, int C, const float* scale,
                              long long rows_per_group, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && y && scale && C > 0 && C % 4 == 0 && rows_per_group > 0, "clv_rows_scale: bad arguments");
  if (rows == 0) return 0;
  rows_scale_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, stream>>>(x, x_is_bf16, y, y_is_bf16, rows, C / 4, scale,
                                                                     rows_per_group);
  return after_launch("rows_scale_kernel");
}
