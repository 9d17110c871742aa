// fp32 loss kernels (the reference runs its losses under @force_fp32).
//
//  * tri-modal exclusive-NCE + pair-wise ranking  (losses/contrastive_loss.py:103-161) and the
//    symmetric InfoNCE of NormSoftmaxLoss (:40-68): cosine normalisation, similarity matrices,
//    masked row/column log-sum-exp with warp-shuffle reductions, diagonal picks, hinge; analytic
//    backward (dS, then dE through the normalisation).
//  * softmax focal / cross-entropy over a vocabulary row (losses/focal_loss.py:61-72,
//    cross_entropy_loss.py:74-81): one CTA per row, online max/sum, fused gradient.
#include <algorithm>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

// ------------------------------------------------------------------------------------------
// generic fp32 SIMT GEMM: C[m,n] = alpha * sum_k A(m,k) * B(k,n) (+ C if accumulate), arbitrary strides
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long long sam, long long sak,
                                                    const float* __restrict__ B, long long sbk, long long sbn,
                                                    float* __restrict__ C, long long ldc, int M, int N, int K, float alpha,
                                                    int accumulate) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i / 64, mm = i % 64;
      const int kA = (sak == 1) ? (i % 16) : kk, mA = (sak == 1) ? (i / 16) : mm;   // coalesce along the unit stride
      sA[kA][mA] = (m0 + mA < M && k0 + kA < K) ? A[(long long)(m0 + mA) * sam + (long long)(k0 + kA) * sak] : 0.f;
      const int kB = (sbk == 1) ? (i % 16) : kk, nB = (sbk == 1) ? (i / 16) : mm;
      sB[kB][nB] = (n0 + nB < N && k0 + kB < K) ? B[(long long)(k0 + kB) * sbk + (long long)(n0 + nB) * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float* c = C + (long long)m * ldc + n;
        *c = accumulate ? (*c + alpha * acc[i][j]) : alpha * acc[i][j];
      }
    }
}

static int sgemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                 long long ldc, int M, int N, int K, float alpha, int accumulate, cudaStream_t stream) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  sgemm_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, alpha, accumulate);
  return after_launch("sgemm_kernel");
}

// ------------------------------------------------------------------------------------------
// cosine normalisation  x / max(||x||, eps)   (contrastive_loss.py:20-25)
// ------------------------------------------------------------------------------------------
__global__ void cos_norm_fwd_kernel(const float* x, float* xn, float* inv_norm, long long rows, int D, float eps) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = x[r * D + c]; s += v * v; }
  s = warp_sum(s);
  const float n = sqrtf(s);
  const float inv = 1.0f / fmaxf(n, eps);
  for (int c = lane; c < D; c += 32) xn[r * D + c] = x[r * D + c] * inv;
  if (lane == 0) inv_norm[r] = (n > eps) ? inv : -inv;    // sign bit marks the clamped branch
}
// dx = (dxn - xn <xn, dxn>) * inv   (or dxn * inv on the clamped branch)
__global__ void cos_norm_bwd_kernel(const float* xn, const float* dxn, const float* inv_norm, float* dx, long long rows,
                                    int D) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float invs = inv_norm[r];
  const float inv = fabsf(invs);
  float dot = 0.f;
  if (invs > 0.f) {
    for (int c = lane; c < D; c += 32) dot += xn[r * D + c] * dxn[r * D + c];
    dot = warp_sum(dot);
  }
  for (int c = lane; c < D; c += 32) dx[r * D + c] = (dxn[r * D + c] - xn[r * D + c] * dot) * inv;
}

// ------------------------------------------------------------------------------------------
// exclusive NCE statistics.  S: [nblk, Bg, Bg] (already / t).
// rowstat[i] = {M_i, Z_0, Z_1, Z_2} with Z_k = sum_m offdiag_m + diag_k (all shifted by M_i);
// for nblk == 1 this is the plain row softmax.  colstat[k, j] = column log-sum-exp.
// ------------------------------------------------------------------------------------------
__global__ void nce_row_kernel(const float* S, int nblk, int Bg, float* rowstat) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= Bg) return;
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int k = 0; k < nblk; ++k)
    for (int j = lane; j < Bg; j += 32) mx = fmaxf(mx, S[((long long)k * Bg + i) * Bg + j]);
  mx = warp_max(mx);
  float off = 0.f;
  for (int k = 0; k < nblk; ++k)
    for (int j = lane; j < Bg; j += 32)
      if (j != i) off += expf(S[((long long)k * Bg + i) * Bg + j] - mx);
  off = warp_sum(off);
  if (lane == 0) {
    rowstat[i * 4 + 0] = mx;
    for (int k = 0; k < 3; ++k)
      rowstat[i * 4 + 1 + k] = k < nblk ? off + expf(S[((long long)k * Bg + i) * Bg + i] - mx) : 1.f;
  }
}
__global__ void nce_col_kernel(const float* S, int nblk, int Bg, float* colstat) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblk * Bg) return;
  const int k = idx / Bg, j = idx % Bg;
  const float* p = S + (long long)k * Bg * Bg + j;
  float mx = -INFINITY;
  for (int i = 0; i < Bg; ++i) mx = fmaxf(mx, p[(long long)i * Bg]);
  float s = 0.f;
  for (int i = 0; i < Bg; ++i) s += expf(p[(long long)i * Bg] - mx);
  colstat[idx] = mx + logf(s);
}
// out[0] = nce loss (loss_v + loss_t), out[1] = ranking loss, active[i] = hinge active flag
__global__ void nce_final_kernel(const float* S, int nblk, int Bg, const float* rowstat, const float* colstat,
                                 float margin, int use_rank, float* out, float* active) {
  __shared__ float red[3][32];
  float lv = 0.f, lt = 0.f, lr = 0.f;
  for (int i = threadIdx.x; i < Bg; i += blockDim.x) {
    const float M = rowstat[i * 4];
    for (int k = 0; k < nblk; ++k) {
      const float d = S[((long long)k * Bg + i) * Bg + i];
      lv += d - M - logf(rowstat[i * 4 + 1 + k]);
      lt += d - colstat[k * Bg + i];
    }
    if (use_rank && nblk >= 2) {
      const float h = margin - (S[(long long)i * Bg + i] - S[((long long)Bg + i) * Bg + i]);
      active[i] = h > 0.f ? 1.f : 0.f;
      lr += fmaxf(h, 0.f);
    }
  }
  lv = warp_sum(lv); lt = warp_sum(lt); lr = warp_sum(lr);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = lv; red[1][w] = lt; red[2][w] = lr; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    lv = lane < nw ? red[0][lane] : 0.f; lt = lane < nw ? red[1][lane] : 0.f; lr = lane < nw ? red[2][lane] : 0.f;
    lv = warp_sum(lv); lt = warp_sum(lt); lr = warp_sum(lr);
    if (lane == 0) {
      out[0] = -lv / Bg - lt / (nblk * (float)Bg);
      out[1] = lr / Bg;
    }
  }
}
// dS[k,i,j] from saved statistics and the two upstream scalars
__global__ void nce_ds_kernel(const float* S, int nblk, int Bg, const float* rowstat, const float* colstat,
                              const float* active, const float* g_nce, const float* g_rank, int use_rank, float* dS) {
  const long long total = (long long)nblk * Bg * Bg;
  const float gn = g_nce ? *g_nce : 0.f, gr = (g_rank && use_rank) ? *g_rank : 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % Bg); const int i = (int)((idx / Bg) % Bg); const int k = (int)(idx / ((long long)Bg * Bg));
    const float s = S[idx];
    const float e = expf(s - rowstat[i * 4]);
    float rowp;
    if (i != j) {
      rowp = 0.f;
      for (int kk = 0; kk < nblk; ++kk) rowp += e / rowstat[i * 4 + 1 + kk];
    } else {
      rowp = e / rowstat[i * 4 + 1 + k] - 1.f;
    }
    const float colp = expf(s - colstat[k * Bg + j]) - (i == j ? 1.f : 0.f);
    float g = gn * (rowp / Bg + colp / (nblk * (float)Bg));
    if (i == j && gr != 0.f && nblk >= 2) {
      if (k == 0) g -= gr * active[i] / Bg;
      if (k == 1) g += gr * active[i] / Bg;
    }
    dS[idx] = g;
  }
}

// ------------------------------------------------------------------------------------------
// softmax focal / CE over vocabulary rows.  One CTA per row.
// stats[r] = {lse, ce, valid};  sums[0] += focal, sums[1] += valid
// ------------------------------------------------------------------------------------------
__global__ void focal_fwd_kernel(const float* logits, long long ld, int V, const long long* target, long long ignore_index,
                                 float gamma, float* stats, float* sums) {
  const long long r = blockIdx.x;
  const long long t = target[r];
  __shared__ float red[32];
  __shared__ float bc;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (t == ignore_index) {
    if (threadIdx.x == 0) { stats[r * 3] = 0.f; stats[r * 3 + 1] = 0.f; stats[r * 3 + 2] = 0.f; }
    return;
  }
  const float* p = logits + r * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, p[c]);
  mx = warp_max(mx);
  if (lane == 0) red[w] = mx;
  __syncthreads();
  if (w == 0) { mx = lane < nw ? red[lane] : -INFINITY; mx = warp_max(mx); if (lane == 0) bc = mx; }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) s += expf(p[c] - mx);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < nw; ++i) tot += red[i];
    const float lse = mx + logf(tot);
    const float ce = lse - p[t];
    const float pt = expf(-ce);
    const float focal = gamma == 0.f ? ce : powf(fmaxf(1.f - pt, 0.f), gamma) * ce;
    stats[r * 3] = lse; stats[r * 3 + 1] = ce; stats[r * 3 + 2] = 1.f;
    atomicAdd(sums + 0, focal);
    atomicAdd(sums + 1, 1.f);
  }
}
__global__ void focal_finalize_kernel(const float* sums, float* loss) { loss[0] = sums[0] / fmaxf(sums[1], 1.f); }
// dlogits[r, c] = g / n_valid * dfocal/dce * (softmax - onehot); padded columns [V, Vpad) get 0
__global__ void focal_bwd_kernel(const float* logits, long long ld, int V, int Vpad, const long long* target, float gamma,
                                 const float* stats, const float* sums, const float* g_loss, void* dlogits, int d_bf16,
                                 long long ld_d) {
  const long long r = blockIdx.x;
  const float valid = stats[r * 3 + 2];
  float coef = 0.f, lse = 0.f;
  if (valid != 0.f) {
    const float ce = stats[r * 3 + 1];
    const float pt = expf(-ce);
    const float om = fmaxf(1.f - pt, 0.f);
    const float dfd = gamma == 0.f ? 1.f : powf(om, gamma) + gamma * powf(om, gamma - 1.f) * pt * ce;
    coef = (*g_loss) * dfd / fmaxf(sums[1], 1.f);
    lse = stats[r * 3];
  }
  const long long t = target[r];
  const float* p = logits + r * ld;
  for (int c = threadIdx.x; c < Vpad; c += blockDim.x) {
    float g = 0.f;
    if (valid != 0.f && c < V) g = coef * (expf(p[c] - lse) - (c == t ? 1.f : 0.f));
    if (d_bf16) reinterpret_cast<__nv_bfloat16*>(dlogits)[r * ld_d + c] = __float2bfloat16(g);
    else reinterpret_cast<float*>(dlogits)[r * ld_d + c] = g;
  }
}

// ------------------------------------------------------------------------------------------
// retrieval evaluation: rank of the ground-truth column in every row of a score matrix, i.e. the
// position np.argsort(-scores, axis=1) gives it (core/evaluation/accuracy.py:447-449); ties keep column order
// (stable sort).  One warp per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) retrieval_rank_kernel(const float* __restrict__ S, long long ld, int rows, int cols,
                                                             const int* __restrict__ gt, int* __restrict__ rank) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int g = gt ? gt[r] : r;
  const float* row = S + (long long)r * ld;
  const float sg = row[g];
  int cnt = 0;
  for (int j = lane; j < cols; j += 32) {
    const float v = row[j];
    cnt += (v > sg) || (v == sg && j < g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) rank[r] = cnt;
}

// x / ||x|| with zero rows left untouched (mmaction/utils/numpy_norm.py:5-8)
__global__ void l2_normalize_kernel(const float* x, float* xn, long long rows, int D) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = x[r * D + c]; s += v * v; }
  s = warp_sum(s);
  const float n = sqrtf(s);
  const float inv = n == 0.f ? 1.0f : 1.0f / n;
  for (int c = lane; c < D; c += 32) xn[r * D + c] = x[r * D + c] * inv;
}

}  // namespace clv

using namespace clv;

// Workspace layout (floats) for nblk similarity blocks of Bg x Bg over D-dim embeddings:
//   Xn[(nblk+1), Bg, D] | inv[(nblk+1), Bg] | S[nblk,Bg,Bg] | rowstat[Bg,4] | colstat[nblk,Bg] | active[Bg] | dS[nblk,Bg,Bg] | dXn[(nblk+1),Bg,D]
extern "C" long long clv_nce_workspace_floats(int nblk, int Bg, int D) {
  const long long e = (long long)(nblk + 1) * Bg * D;
  return e + (long long)(nblk + 1) * Bg + 2LL * nblk * Bg * Bg + 4LL * Bg + (long long)nblk * Bg + Bg + e;
}

struct NceWs { float *Xn, *inv, *S, *rowstat, *colstat, *active, *dS, *dXn; };
static NceWs carve(float* ws, int nblk, int Bg, int D) {
  NceWs w; const long long e = (long long)(nblk + 1) * Bg * D;
  w.Xn = ws; w.inv = w.Xn + e; w.S = w.inv + (long long)(nblk + 1) * Bg; w.rowstat = w.S + (long long)nblk * Bg * Bg;
  w.colstat = w.rowstat + 4LL * Bg; w.active = w.colstat + (long long)nblk * Bg; w.dS = w.active + Bg;
  w.dXn = w.dS + (long long)nblk * Bg * Bg;
  return w;
}

// emb: (nblk+1) pointers: query-side matrix first (video), then the key-side matrices (text, text_mask, text_recon).
extern "C" int clv_nce_rank_fwd(const float* const* emb, int nblk, int Bg, int D, float temperature, float margin,
                                int use_rank, float eps, float* workspace, float* out_losses, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(emb && workspace && out_losses && (nblk == 1 || nblk == 3) && Bg > 0 && D > 0, "clv_nce_rank_fwd: bad arguments");
  NceWs w = carve(workspace, nblk, Bg, D);
  const int wpb = 8;
  for (int m = 0; m <= nblk; ++m) {
    cos_norm_fwd_kernel<<<(Bg + wpb - 1) / wpb, wpb * 32, 0, stream>>>(emb[m], w.Xn + (long long)m * Bg * D,
                                                                     w.inv + (long long)m * Bg, Bg, D, eps);
    if (int rc = after_launch("cos_norm_fwd_kernel")) return rc;
  }
  for (int k = 0; k < nblk; ++k)
    if (int rc = sgemm(w.Xn, D, 1, w.Xn + (long long)(k + 1) * Bg * D, 1, D, w.S + (long long)k * Bg * Bg, Bg, Bg, Bg, D,
                       1.0f / temperature, 0, stream)) return rc;
  nce_row_kernel<<<(Bg + wpb - 1) / wpb, wpb * 32, 0, stream>>>(w.S, nblk, Bg, w.rowstat);
  if (int rc = after_launch("nce_row_kernel")) return rc;
  nce_col_kernel<<<(nblk * Bg + 127) / 128, 128, 0, stream>>>(w.S, nblk, Bg, w.colstat);
  if (int rc = after_launch("nce_col_kernel")) return rc;
  nce_final_kernel<<<1, 256, 0, stream>>>(w.S, nblk, Bg, w.rowstat, w.colstat, margin, use_rank, out_losses, w.active);
  return after_launch("nce_final_kernel");
}

// grads: (nblk+1) output pointers [Bg, D] fp32 (d loss / d emb[m]); g_nce / g_rank device scalars.
extern "C" int clv_nce_rank_bwd(int nblk, int Bg, int D, float temperature, int use_rank, float* workspace,
                                const float* g_nce, const float* g_rank, float* const* grads, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(workspace && grads && (nblk == 1 || nblk == 3), "clv_nce_rank_bwd: bad arguments");
  NceWs w = carve(workspace, nblk, Bg, D);
  const long long total = (long long)nblk * Bg * Bg;
  nce_ds_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, stream>>>(w.S, nblk, Bg, w.rowstat, w.colstat,
                                                                                      w.active, g_nce, g_rank, use_rank, w.dS);
  if (int rc = after_launch("nce_ds_kernel")) return rc;
  const float it = 1.0f / temperature;
  for (int k = 0; k < nblk; ++k) {
    // dVn (+)= dS_k . Xn_k / t
    if (int rc = sgemm(w.dS + (long long)k * Bg * Bg, Bg, 1, w.Xn + (long long)(k + 1) * Bg * D, D, 1, w.dXn, D, Bg, D, Bg, it,
                       k > 0, stream)) return rc;
    // dXn_k = dS_k^T . Vn / t
    if (int rc = sgemm(w.dS + (long long)k * Bg * Bg, 1, Bg, w.Xn, D, 1, w.dXn + (long long)(k + 1) * Bg * D, D, Bg, D, Bg, it, 0,
                       stream)) return rc;
  }
  const int wpb = 8;
  for (int m = 0; m <= nblk; ++m) {
    cos_norm_bwd_kernel<<<(Bg + wpb - 1) / wpb, wpb * 32, 0, stream>>>(w.Xn + (long long)m * Bg * D, w.dXn + (long long)m * Bg * D,
                                                                     w.inv + (long long)m * Bg, grads[m], Bg, D);
    if (int rc = after_launch("cos_norm_bwd_kernel")) return rc;
  }
  return 0;
}

extern "C" int clv_softmax_focal_fwd(const float* logits, long long ld, long long rows, int V, const long long* target,
                                     long long ignore_index, float gamma, float* stats /*[rows,3]*/, float* sums /*[2]*/,
                                     float* loss, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(logits && target && stats && sums && loss && V > 0, "clv_softmax_focal_fwd: bad arguments");
  CLV_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), stream));
  if (rows > 0) {
    focal_fwd_kernel<<<(unsigned)rows, 256, 0, stream>>>(logits, ld, V, target, ignore_index, gamma, stats, sums);
    if (int rc = after_launch("focal_fwd_kernel")) return rc;
  }
  focal_finalize_kernel<<<1, 1, 0, stream>>>(sums, loss);
  return after_launch("focal_finalize_kernel");
}

extern "C" int clv_softmax_focal_bwd(const float* logits, long long ld, long long rows, int V, int Vpad, const long long* target,
                                     float gamma, const float* stats, const float* sums, const float* g_loss, void* dlogits,
                                     int dlogits_is_bf16, long long ld_d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(logits && target && stats && sums && g_loss && dlogits && Vpad >= V, "clv_softmax_focal_bwd: bad arguments");
  if (rows == 0) return 0;
  focal_bwd_kernel<<<(unsigned)rows, 256, 0, stream>>>(logits, ld, V, Vpad, target, gamma, stats, sums, g_loss, dlogits,
                                                      dlogits_is_bf16, ld_d);
  return after_launch("focal_bwd_kernel");
}

// scores[i, j] = <a_i / |a_i|, b_j / |b_j|>  (fp32).  workspace: (n_a + n_b) * D floats.
extern "C" int clv_cosine_scores(const float* a, int n_a, const float* b, int n_b, int D, float* scores, long long ld_scores,
                                 float* workspace, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(a && b && scores && workspace && n_a > 0 && n_b > 0 && D > 0 && ld_scores >= n_b, "clv_cosine_scores: bad arguments");
  float* an = workspace;
  float* bn = workspace + (long long)n_a * D;
  l2_normalize_kernel<<<(n_a + 7) / 8, 256, 0, stream>>>(a, an, n_a, D);
  if (int rc = after_launch("l2_normalize_kernel")) return rc;
  l2_normalize_kernel<<<(n_b + 7) / 8, 256, 0, stream>>>(b, bn, n_b, D);
  if (int rc = after_launch("l2_normalize_kernel")) return rc;
  return sgemm(an, D, 1, bn, 1, D, scores, ld_scores, n_a, n_b, D, 1.0f, 0, stream);
}

extern "C" int clv_retrieval_ranks(const float* scores, long long ld, int rows, int cols, const int* gt_col, int* rank_out,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(scores && rank_out && rows > 0 && cols > 0 && ld >= cols, "clv_retrieval_ranks: bad arguments");
  CLV_REQUIRE(gt_col || rows <= cols, "clv_retrieval_ranks: diagonal ground truth needs rows <= cols");
  retrieval_rank_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(scores, ld, rows, cols, gt_col, rank_out);
  return after_launch("retrieval_rank_kernel");
}
