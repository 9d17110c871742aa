// Multi-tensor AdamW step for the fp32 master weights of the Clover training path (sm_100a), SURVEY.md 8(f1).
//
// Replaces, in ONE pass over parameters + gradients + moments, what the reference spreads over
// core/hooks/mmcv_Fp16OptimizerHook.py:96-149 (copy grads to fp32 masters, unscale by the loss scale, isfinite check with a
// host sync, clip_grad_norm_ :max_norm 15/5/50, optimizer.step, copy masters back to the half weights) and torch.optim.AdamW
// with the per-parameter lr / weight-decay of its paramwise_cfg (configs/exp_local/pretrain_webvid_cc3m.py:129-137):
//   g' = g * grad_scale * min(1, max_norm / (||g * grad_scale|| + 1e-6))
//   p *= 1 - lr * wd;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr / (1-b1^t) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// and, in the same pass, refreshes the bf16 operand copy of the weight that the tcgen05 GEMMs read (so no cast pass runs
// after the step).  A non-finite gradient norm skips the whole step on the device (the reference's overflow skip) without a
// host round trip: the flag is returned through `status` for the loss scaler to read whenever it wants.
//
// HBM-bound: 16 B read + 12 B (+2 B) written per parameter; one CTA per 16 Ki-element chunk of one tensor.
#include "common.cuh"
#include "clover_b200.h"

namespace clv {

struct AdamTensor {            // mirrors clv_adamw_tensor_t
  float* p; const float* g; float* m; float* v; __nv_bfloat16* w16; long long n; float lr, wd;
};

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const AdamTensor* __restrict__ ts, const int* __restrict__ chunk_tensor,
                                                          const long long* __restrict__ chunk_off, int chunk, float* out) {
  const AdamTensor t = ts[chunk_tensor[blockIdx.x]];
  const long long o = chunk_off[blockIdx.x];
  const int n = (int)min((long long)chunk, t.n - o);
  const float* g = t.g + o;
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(g) + i);
      s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) s += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < n; i += 256) s += g[i] * g[i];
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
    s += __shfl_xor_sync(0xffu, s, 4); s += __shfl_xor_sync(0xffu, s, 2); s += __shfl_xor_sync(0xffu, s, 1);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

struct AdamHyper { float b1, b2, eps, inv_bc1, inv_sqrt_bc2, grad_scale, max_norm; const float* step_dev; };

CLV_DEVICE void adam_elem(float& p, float g, float& m, float& v, float coef, float lr, float decay, const AdamHyper& h) {
  g *= coef;
  m = fmaf(h.b1, m, (1.f - h.b1) * g);
  v = fmaf(h.b2, v, (1.f - h.b2) * g * g);
  const float denom = fmaf(sqrtf(v), h.inv_sqrt_bc2, h.eps);
  p = fmaf(-lr * h.inv_bc1, m / denom, p * decay);
}

__global__ void __launch_bounds__(256) adamw_kernel(const AdamTensor* __restrict__ ts, const int* __restrict__ chunk_tensor,
                                                    const long long* __restrict__ chunk_off, int chunk, AdamHyper h,
                                                    const float* __restrict__ sqnorm, float* status) {
  float coef = h.grad_scale;
  if (h.step_dev) {                              // device-resident step counter: this is update number *step_dev + 1
    const float t1 = *h.step_dev + 1.f;
    h.inv_bc1 = 1.f / (1.f - powf(h.b1, t1));
    h.inv_sqrt_bc2 = rsqrtf(1.f - powf(h.b2, t1));
  }
  if (sqnorm) {
    const float nrm = sqrtf(*sqnorm) * fabsf(h.grad_scale);
    if (!isfinite(nrm)) {                        // overflow: skip the step (reference: Fp16OptimizerHook skips + shrinks the scale)
      if (blockIdx.x == 0 && threadIdx.x == 0) { status[0] = nrm; status[1] = 1.f; }
      return;
    }
    if (h.max_norm > 0.f) coef *= fminf(1.f, h.max_norm / (nrm + 1e-6f));
    if (blockIdx.x == 0 && threadIdx.x == 0) { status[0] = nrm; status[1] = 0.f; }
  }
  const AdamTensor t = ts[chunk_tensor[blockIdx.x]];
  const long long o = chunk_off[blockIdx.x];
  const int n = (int)min((long long)chunk, t.n - o);
  float* p = t.p + o; const float* g = t.g + o; float* m = t.m + o; float* v = t.v + o;
  __nv_bfloat16* w = t.w16 ? t.w16 + o : nullptr;
  const float decay = 1.f - t.lr * t.wd;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!w || (reinterpret_cast<uintptr_t>(w) & 7) == 0);
  int done = 0;
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
      adam_elem(P.x, G.x, M.x, V.x, coef, t.lr, decay, h); adam_elem(P.y, G.y, M.y, V.y, coef, t.lr, decay, h);
      adam_elem(P.z, G.z, M.z, V.z, coef, t.lr, decay, h); adam_elem(P.w, G.w, M.w, V.w, coef, t.lr, decay, h);
      reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
      if (w) reinterpret_cast<uint2*>(w)[i] = make_uint2(pack_bf16(P.x, P.y), pack_bf16(P.z, P.w));
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < n; i += 256) {
    float P = p[i], M = m[i], V = v[i];
    adam_elem(P, g[i], M, V, coef, t.lr, decay, h);
    p[i] = P; m[i] = M; v[i] = V;
    if (w) w[i] = __float2bfloat16_rn(P);
  }
}

// the step counter advances only when the update was applied (a skipped overflow step must not age the bias correction)
__global__ void adam_step_advance_kernel(float* step_dev, const float* status, int check) {
  if (!check || status[1] == 0.f) *step_dev += 1.f;
}

}  // namespace clv

using namespace clv;

static_assert(sizeof(clv_adamw_tensor_t) == sizeof(AdamTensor), "clv_adamw_tensor_t layout");

extern "C" int clv_adamw_step(const clv_adamw_tensor_t* tensors_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                              int n_chunks, int chunk_elems, float beta1, float beta2, float eps, int step, float grad_scale,
                              float max_grad_norm, int check_finite, float* status_dev, float* step_dev, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(tensors_dev && chunk_tensor_dev && chunk_offset_dev && n_chunks >= 0 && chunk_elems > 0 && chunk_elems % 4 == 0,
              "clv_adamw_step: bad arguments");
  CLV_REQUIRE((step >= 1 || step_dev) && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "clv_adamw_step: bad hyper-parameters");
  const bool need_norm = max_grad_norm > 0.f || check_finite;
  CLV_REQUIRE(!need_norm || status_dev, "clv_adamw_step: status_dev (fp32[3]) is required for clipping / the finite check");
  if (n_chunks == 0) return 0;
  const AdamTensor* ts = reinterpret_cast<const AdamTensor*>(tensors_dev);
  float* sq = nullptr;
  if (need_norm) {
    sq = status_dev + 2;
    CLV_CHECK_CUDA(cudaMemsetAsync(sq, 0, sizeof(float), stream));
    grad_sqnorm_kernel<<<n_chunks, 256, 0, stream>>>(ts, chunk_tensor_dev, chunk_offset_dev, chunk_elems, sq);
    if (int rc = after_launch("grad_sqnorm_kernel")) return rc;
  }
  AdamHyper h;
  h.b1 = beta1; h.b2 = beta2; h.eps = eps;
  const int host_step = step >= 1 ? step : 1;
  h.inv_bc1 = (float)(1.0 / (1.0 - pow((double)beta1, host_step)));
  h.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, host_step)));
  h.grad_scale = grad_scale; h.max_norm = max_grad_norm; h.step_dev = step_dev;
  adamw_kernel<<<n_chunks, 256, 0, stream>>>(ts, chunk_tensor_dev, chunk_offset_dev, chunk_elems, h, sq, status_dev);
  if (int rc = after_launch("adamw_kernel")) return rc;
  if (step_dev) {
    adam_step_advance_kernel<<<1, 1, 0, stream>>>(step_dev, status_dev, need_norm ? 1 : 0);
    return after_launch("adam_step_advance_kernel");
  }
  return 0;
}
