// Row-mapped LayerNorm: the norm1 / norm2 hot path of SwinTransformerBlock3D (swin_transformer_3d.py:450,483)
// and every plain nn.LayerNorm on dense [rows, C] activations (HF BERT add&norm, final Swin norm).
//
// HBM-bound.  The generic kernels in norm.cu carry every layout variant of the model in one body and pay for it
// in registers (126-222 per thread -> 12-25 % occupancy, ~2.2 TB/s).  These kernels keep only what the hot path
// needs so that a row costs ~100 issued instructions forward / ~260 backward:
//   * rows are dense and contiguous (pitch == C); a row is owned by L = 8, 16 or 32 lanes (32/L rows per warp),
//     each lane holding V float4 -- narrow rows (C = 128) still move 2 KB per warp iteration;
//   * the fused roll + window_partition (or window_reverse + roll back) permutation of one clip arrives as an
//     int32 row map built on the host from the closed forms (clover_b200/tables.py): mapped row of source row s
//     is (s / period) * period + map[s % period].  Forward writes y at the mapped row (LN1 -> window order);
//     backward reads dy at the mapped row (LN1) or writes the bf16 copy of dx there (LN2 -> proj operand);
//   * statistics are indexed by the source row; dgamma / dbeta live in registers, are folded through shared
//     memory once per CTA and leave with one atomic per column per CTA.
#include <algorithm>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

struct LnrArgs {
  const void* x;
  const float* gamma; const float* beta; float eps;
  float* mean; float* rstd;
  unsigned rows; int C;
  const int* row_map; unsigned period;
  // forward
  void* y; int y_mapped;
  // backward
  const void* dy; int dy_mapped;
  const float* dres; float* dx;
  __nv_bfloat16* dx16; int dx16_mapped;
  float* dgamma; float* dbeta;
  float* dxsum;                  // [C] column sums of the final dx (bias gradient of the nn.Linear that produced x's input)
  const float* copy_scale; unsigned copy_scale_rows;   // per-sample DropPath factor applied to dx16 and dxsum (not dx)
  // mask-token blend after the patch-embed LN: y = LN(x) (1 - w[s]) + token w[s]; the backward accumulates dtoken in the
  // dxsum accumulators (the two never occur together)
  const float* row_w; const float* token;
};

template <bool BF>
CLV_DEVICE float4 lnr_ld(const void* base, size_t elem) {
  if (BF) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem);
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
}
template <bool BF>
CLV_DEVICE void lnr_st(void* base, size_t elem, float4 v) {
  if (BF) {
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem) = v;
  }
}
template <int L>
CLV_DEVICE float group_sum(float v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// The kernels below are latency-bound on their row loads (ncu r02a: long-scoreboard stalls, 12-25 % occupancy because the
// column accumulators live in registers).  Rather than spending registers on a second row in flight, every row group asks L2
// for the lines of the row it will process NEXT iteration before it starts computing the current one: the bytes in flight
// double without a single extra live register, and the next iteration's loads hit L2.
CLV_DEVICE void lnr_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int L>
CLV_DEVICE void lnr_prefetch_row(const void* base, size_t row_byte_offset, int row_bytes, int sub) {
  const char* p = reinterpret_cast<const char*>(base) + row_byte_offset;
  for (int b = sub * 128; b < row_bytes; b += L * 128) lnr_prefetch_l2(p + b);
}
CLV_DEVICE unsigned lnr_mapped(const LnrArgs& a, unsigned s) {
  const unsigned q = s / a.period;
  return q * a.period + (unsigned)__ldg(a.row_map + (s - q * a.period));
}

template <int V, int L, bool XB, bool YB>
__global__ void __launch_bounds__(256) lnr_fwd_kernel(LnrArgs a) {
  constexpr int G = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, grp = lane / L;
  const unsigned warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.0f / (float)a.C;
  for (unsigned r0 = warp_global * G; r0 < a.rows; r0 += nwarps * G) {
    const unsigned s = r0 + grp;
    const bool live = s < a.rows;
    {
      const unsigned sn = s + nwarps * G;
      if (sn < a.rows) lnr_prefetch_row<L>(a.x, (size_t)sn * a.C * (XB ? 2 : 4), a.C * (XB ? 2 : 4), sub);
    }
    float4 v[V];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = live ? lnr_ld<XB>(a.x, (size_t)s * a.C + 4 * (sub + L * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mu = group_sum<L>(sum) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j].x -= mu; v[j].y -= mu; v[j].z -= mu; v[j].w -= mu;
      sq = fmaf(v[j].x, v[j].x, sq); sq = fmaf(v[j].y, v[j].y, sq); sq = fmaf(v[j].z, v[j].z, sq); sq = fmaf(v[j].w, v[j].w, sq);
    }
    const float rs = rsqrtf(group_sum<L>(sq) * inv_c + a.eps);
    if (!live) continue;
    const unsigned orow = a.y_mapped ? lnr_mapped(a, s) : s;
    const float bw = a.row_w ? __ldg(a.row_w + s) : 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = 4 * (sub + L * j);
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta + c));
      float4 o;
      o.x = fmaf(v[j].x * rs, g.x, b.x); o.y = fmaf(v[j].y * rs, g.y, b.y);
      o.z = fmaf(v[j].z * rs, g.z, b.z); o.w = fmaf(v[j].w * rs, g.w, b.w);
      if (bw != 0.f) {
        const float4 tk = __ldg(reinterpret_cast<const float4*>(a.token + c));
        o.x = o.x * (1.f - bw) + tk.x * bw; o.y = o.y * (1.f - bw) + tk.y * bw;
        o.z = o.z * (1.f - bw) + tk.z * bw; o.w = o.w * (1.f - bw) + tk.w * bw;
      }
      lnr_st<YB>(a.y, (size_t)orow * a.C + c, o);
    }
    if (sub == 0 && a.mean) { a.mean[s] = mu; a.rstd[s] = rs; }
  }
}

// DRES: a residual-stream gradient is added (Swin blocks).  Without it (BERT / fusion / final norms) the V float4 that would hold
// it are not allocated, which lets the V = 6 (C = 768) instantiation fit two CTAs per SM: these kernels are latency-bound on
// their row loads and 8 warps per SM left the 29184 x 768 fusion LayerNorm at 1.4 TB/s.
template <int V, int L, bool XB, bool DYB, bool DXS, bool DRES>
__global__ void __launch_bounds__(256, ((V <= 4 || (V <= 6 && !DRES && !DXS)) ? 2 : 1)) lnr_bwd_kernel(LnrArgs a) {
  constexpr int G = 32 / L;
  extern __shared__ float red[];          // [3][C]: dgamma | dbeta | dx column sums of this CTA
  const int lane = threadIdx.x & 31, sub = lane % L, grp = lane / L;
  const unsigned warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.0f / (float)a.C;
  for (int c = threadIdx.x; c < 3 * a.C; c += blockDim.x) red[c] = 0.f;
  float4 dg[V], db[V], ds[DXS ? V : 1];
#pragma unroll
  for (int j = 0; j < V; ++j) dg[j] = db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < (DXS ? V : 1); ++j) ds[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (unsigned r0 = warp_global * G; r0 < a.rows; r0 += nwarps * G) {
    const unsigned s = r0 + grp;
    const bool live = s < a.rows;
    const unsigned m = (live && (a.dy_mapped | a.dx16_mapped)) ? lnr_mapped(a, s) : s;
    const unsigned dyrow = a.dy_mapped ? m : s;
    {
      const unsigned sn = s + nwarps * G;
      if (sn < a.rows) {
        lnr_prefetch_row<L>(a.x, (size_t)sn * a.C * (XB ? 2 : 4), a.C * (XB ? 2 : 4), sub);
        const unsigned dn = a.dy_mapped ? lnr_mapped(a, sn) : sn;
        lnr_prefetch_row<L>(a.dy, (size_t)dn * a.C * (DYB ? 2 : 4), a.C * (DYB ? 2 : 4), sub);
        if (DRES) lnr_prefetch_row<L>(a.dres, (size_t)sn * a.C * 4, a.C * 4, sub);
      }
    }
    float4 xh[V], d[V], rr[DRES ? V : 1];
    float mu = 0.f, rs = 0.f;
    if (live) { mu = a.mean[s]; rs = a.rstd[s]; }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = 4 * (sub + L * j);
      xh[j] = d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (DRES) rr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        xh[j] = lnr_ld<XB>(a.x, (size_t)s * a.C + c);
        d[j] = lnr_ld<DYB>(a.dy, (size_t)dyrow * a.C + c);
        if (DRES) rr[j] = *reinterpret_cast<const float4*>(a.dres + (size_t)s * a.C + c);
      }
    }
    if (DXS && a.row_w) {            // d(LN out) = dy (1 - w); the token takes dy w
      const float bw = live ? __ldg(a.row_w + s) : 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        ds[j].x = fmaf(d[j].x, bw, ds[j].x); ds[j].y = fmaf(d[j].y, bw, ds[j].y);
        ds[j].z = fmaf(d[j].z, bw, ds[j].z); ds[j].w = fmaf(d[j].w, bw, ds[j].w);
        d[j].x *= (1.f - bw); d[j].y *= (1.f - bw); d[j].z *= (1.f - bw); d[j].w *= (1.f - bw);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = 4 * (sub + L * j);
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
      xh[j].x = (xh[j].x - mu) * rs; xh[j].y = (xh[j].y - mu) * rs; xh[j].z = (xh[j].z - mu) * rs; xh[j].w = (xh[j].w - mu) * rs;
      dg[j].x = fmaf(d[j].x, xh[j].x, dg[j].x); dg[j].y = fmaf(d[j].y, xh[j].y, dg[j].y);
      dg[j].z = fmaf(d[j].z, xh[j].z, dg[j].z); dg[j].w = fmaf(d[j].w, xh[j].w, dg[j].w);
      db[j].x += d[j].x; db[j].y += d[j].y; db[j].z += d[j].z; db[j].w += d[j].w;
      d[j].x *= g.x; d[j].y *= g.y; d[j].z *= g.z; d[j].w *= g.w;       // d <- dy * gamma
      s1 = fmaf(d[j].x, xh[j].x, s1); s1 = fmaf(d[j].y, xh[j].y, s1); s1 = fmaf(d[j].z, xh[j].z, s1); s1 = fmaf(d[j].w, xh[j].w, s1);
      s2 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
    }
    s1 = group_sum<L>(s1) * inv_c;
    s2 = group_sum<L>(s2) * inv_c;
    if (!live || (!a.dx && !a.dx16)) continue;
    const unsigned crow = a.dx16_mapped ? m : s;
    const float sc = a.copy_scale ? __ldg(a.copy_scale + s / a.copy_scale_rows) : 1.0f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = 4 * (sub + L * j);
      float4 o;
      const float4 r4 = DRES ? rr[DRES ? j : 0] : make_float4(0.f, 0.f, 0.f, 0.f);
      o.x = fmaf(rs, d[j].x - s2 - xh[j].x * s1, r4.x); o.y = fmaf(rs, d[j].y - s2 - xh[j].y * s1, r4.y);
      o.z = fmaf(rs, d[j].z - s2 - xh[j].z * s1, r4.z); o.w = fmaf(rs, d[j].w - s2 - xh[j].w * s1, r4.w);
      if (a.dx) *reinterpret_cast<float4*>(a.dx + (size_t)s * a.C + c) = o;
      o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc;          // the branch copy / bias gradient carry the DropPath factor
      if (a.dx16) lnr_st<true>(a.dx16, (size_t)crow * a.C + c, o);
      if (DXS && !a.row_w) { ds[j].x += o.x; ds[j].y += o.y; ds[j].z += o.z; ds[j].w += o.w; }
    }
  }

  if (!a.dgamma && !DXS) return;
  // fold the row groups of the warp, then the CTA's warps through shared memory, then one atomic per column
#pragma unroll
  for (int j = 0; j < V; ++j) {
#pragma unroll
    for (int o = L; o < 32; o <<= 1) {
      dg[j].x += __shfl_xor_sync(0xffffffffu, dg[j].x, o); dg[j].y += __shfl_xor_sync(0xffffffffu, dg[j].y, o);
      dg[j].z += __shfl_xor_sync(0xffffffffu, dg[j].z, o); dg[j].w += __shfl_xor_sync(0xffffffffu, dg[j].w, o);
      db[j].x += __shfl_xor_sync(0xffffffffu, db[j].x, o); db[j].y += __shfl_xor_sync(0xffffffffu, db[j].y, o);
      db[j].z += __shfl_xor_sync(0xffffffffu, db[j].z, o); db[j].w += __shfl_xor_sync(0xffffffffu, db[j].w, o);
      if (DXS) {
        ds[j].x += __shfl_xor_sync(0xffffffffu, ds[j].x, o); ds[j].y += __shfl_xor_sync(0xffffffffu, ds[j].y, o);
        ds[j].z += __shfl_xor_sync(0xffffffffu, ds[j].z, o); ds[j].w += __shfl_xor_sync(0xffffffffu, ds[j].w, o);
      }
    }
  }
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = 4 * (sub + L * j);
      atomicAdd(red + c, dg[j].x); atomicAdd(red + c + 1, dg[j].y); atomicAdd(red + c + 2, dg[j].z); atomicAdd(red + c + 3, dg[j].w);
      atomicAdd(red + a.C + c, db[j].x); atomicAdd(red + a.C + c + 1, db[j].y);
      atomicAdd(red + a.C + c + 2, db[j].z); atomicAdd(red + a.C + c + 3, db[j].w);
      if (DXS) {
        atomicAdd(red + 2 * a.C + c, ds[j].x); atomicAdd(red + 2 * a.C + c + 1, ds[j].y);
        atomicAdd(red + 2 * a.C + c + 2, ds[j].z); atomicAdd(red + 2 * a.C + c + 3, ds[j].w);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    if (a.dgamma) {
      atomicAdd(a.dgamma + c, red[c]);
      atomicAdd(a.dbeta + c, red[a.C + c]);
    }
    if (DXS) atomicAdd(a.dxsum + c, red[2 * a.C + c]);
  }
}

// (V, L) choice: the fewest lanes per row that keep V <= 4, else a full warp with V <= 8
static bool lnr_shape(int C, int& V, int& L) {
  if (C % 4) return false;
  const int nvec = C / 4;
  for (int l : {8, 16, 32}) {
    if (nvec % l == 0 && (nvec / l == 3 || nvec / l == 4)) { V = nvec / l; L = l; return true; }
  }
  if (nvec % 32 == 0 && (nvec / 32 == 6 || nvec / 32 == 8)) { V = nvec / 32; L = 32; return true; }
  return false;
}

#define LNR_SHAPES(KERNEL, XB, OB, ...)                                       \
  if (L == 8 && V == 3) KERNEL<3, 8, XB, OB> __VA_ARGS__;                     \
  else if (L == 8 && V == 4) KERNEL<4, 8, XB, OB> __VA_ARGS__;                \
  else if (L == 16 && V == 3) KERNEL<3, 16, XB, OB> __VA_ARGS__;              \
  else if (L == 16 && V == 4) KERNEL<4, 16, XB, OB> __VA_ARGS__;              \
  else if (L == 32 && V == 3) KERNEL<3, 32, XB, OB> __VA_ARGS__;              \
  else if (L == 32 && V == 4) KERNEL<4, 32, XB, OB> __VA_ARGS__;              \
  else if (L == 32 && V == 6) KERNEL<6, 32, XB, OB> __VA_ARGS__;              \
  else KERNEL<8, 32, XB, OB> __VA_ARGS__;

#define LNR_SHAPES_B2(XB, OB, S, R, ...)                                      \
  if (L == 8 && V == 3) lnr_bwd_kernel<3, 8, XB, OB, S, R> __VA_ARGS__;       \
  else if (L == 8 && V == 4) lnr_bwd_kernel<4, 8, XB, OB, S, R> __VA_ARGS__;  \
  else if (L == 16 && V == 3) lnr_bwd_kernel<3, 16, XB, OB, S, R> __VA_ARGS__; \
  else if (L == 16 && V == 4) lnr_bwd_kernel<4, 16, XB, OB, S, R> __VA_ARGS__; \
  else if (L == 32 && V == 3) lnr_bwd_kernel<3, 32, XB, OB, S, R> __VA_ARGS__; \
  else if (L == 32 && V == 4) lnr_bwd_kernel<4, 32, XB, OB, S, R> __VA_ARGS__; \
  else if (L == 32 && V == 6) lnr_bwd_kernel<6, 32, XB, OB, S, R> __VA_ARGS__; \
  else lnr_bwd_kernel<8, 32, XB, OB, S, R> __VA_ARGS__;
#define LNR_SHAPES_B(XB, OB, S, ...)                                          \
  if (has_dres) { LNR_SHAPES_B2(XB, OB, S, true, __VA_ARGS__) } else { LNR_SHAPES_B2(XB, OB, S, false, __VA_ARGS__) }

#define LNR_DISPATCH(KERNEL, xb, ob, ...)                                     \
  if (xb) { if (ob) { LNR_SHAPES(KERNEL, true, true, __VA_ARGS__) } else { LNR_SHAPES(KERNEL, true, false, __VA_ARGS__) } } \
  else    { if (ob) { LNR_SHAPES(KERNEL, false, true, __VA_ARGS__) } else { LNR_SHAPES(KERNEL, false, false, __VA_ARGS__) } }

static int lnr_fill(LnrArgs& a, const clv_lnr_desc_t* d, int& V, int& L) {
  CLV_REQUIRE(d && d->x && d->gamma && d->beta, "lnr: null pointer");
  CLV_REQUIRE(lnr_shape(d->C, V, L), "lnr: unsupported width %d (see clv_lnr_supported)", d->C);
  CLV_REQUIRE(d->rows >= 0 && d->rows < 2000000000LL, "lnr: row count out of range");
  CLV_REQUIRE(!d->row_map || (d->map_period > 0 && d->rows % d->map_period == 0), "lnr: rows must be a multiple of map_period");
  a = LnrArgs{};
  a.x = d->x; a.gamma = d->gamma; a.beta = d->beta; a.eps = d->eps; a.mean = d->mean; a.rstd = d->rstd;
  a.rows = (unsigned)d->rows; a.C = d->C; a.row_map = d->row_map; a.period = d->row_map ? (unsigned)d->map_period : 1u;
  CLV_REQUIRE(!d->row_blend || d->blend_token, "lnr: row_blend needs blend_token");
  a.row_w = d->row_blend; a.token = d->blend_token;
  return 0;
}

}  // namespace clv

using namespace clv;

extern "C" int clv_lnr_supported(int C) {
  int V, L;
  return lnr_shape(C, V, L) ? 1 : 0;
}

extern "C" int clv_lnr_fwd(const clv_lnr_desc_t* d, void* y, int y_is_bf16, int y_mapped, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LnrArgs a; int V, L;
  if (int rc = lnr_fill(a, d, V, L)) return rc;
  CLV_REQUIRE(y != nullptr, "lnr_fwd: null output");
  CLV_REQUIRE(!y_mapped || d->row_map, "lnr_fwd: y_mapped without a row map");
  a.y = y; a.y_mapped = y_mapped;
  if (a.rows == 0) return 0;
  const int rows_per_block = 8 * (32 / L);
  const long long blocks = std::min<long long>(((long long)a.rows + rows_per_block - 1) / rows_per_block, (long long)num_sms() * 8);
  const bool xb = d->x_is_bf16 != 0, yb = y_is_bf16 != 0;
  LNR_DISPATCH(lnr_fwd_kernel, xb, yb, <<<(int)blocks, 256, 0, stream>>>(a));
  return after_launch("lnr_fwd_kernel");
}

extern "C" int clv_lnr_bwd(const clv_lnr_desc_t* d, const clv_lnr_bwd_t* b, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LnrArgs a; int V, L;
  if (int rc = lnr_fill(a, d, V, L)) return rc;
  CLV_REQUIRE(b && b->dy && d->mean && d->rstd, "lnr_bwd: null pointer");
  CLV_REQUIRE(!(b->dy_mapped || b->dx_bf16_mapped) || d->row_map, "lnr_bwd: mapped rows without a row map");
  CLV_REQUIRE((b->dgamma == nullptr) == (b->dbeta == nullptr), "lnr_bwd: dgamma and dbeta go together");
  a.dy = b->dy; a.dy_mapped = b->dy_mapped; a.dres = b->dres; a.dx = b->dx;
  a.dx16 = reinterpret_cast<__nv_bfloat16*>(b->dx_bf16); a.dx16_mapped = b->dx_bf16_mapped;
  a.dgamma = b->dgamma; a.dbeta = b->dbeta;
  CLV_REQUIRE(!b->copy_scale || b->copy_scale_rows > 0, "lnr_bwd: copy_scale needs copy_scale_rows > 0");
  a.copy_scale = b->copy_scale; a.copy_scale_rows = b->copy_scale ? (unsigned)b->copy_scale_rows : 1u;
  if (a.rows == 0) return 0;
  const int rows_per_block = 8 * (32 / L);
  const long long blocks = std::min<long long>(((long long)a.rows + rows_per_block - 1) / rows_per_block, (long long)num_sms() * 2);
  const size_t smem = 3 * (size_t)a.C * sizeof(float);
  const bool xb = d->x_is_bf16 != 0, dyb = b->dy_is_bf16 != 0;
  const bool has_dres = b->dres != nullptr;
  a.dxsum = b->dxsum;
  if (a.row_w) {
    // patch-embed LN with the mask-token blend: fp32 x and dy, dtoken rides in the column-sum accumulators
    CLV_REQUIRE(b->dtoken && !b->dxsum && !xb && !dyb && !b->copy_scale, "lnr_bwd: the blend needs dtoken, fp32 x / dy, no dxsum / copy_scale");
    a.dxsum = b->dtoken;
    LNR_SHAPES_B(false, false, true, <<<(int)blocks, 256, smem, stream>>>(a))
  } else if (a.dxsum) {
    // fp32 residual stream, bf16 dy: the Swin block configuration (the only producer of a fused bias gradient)
    CLV_REQUIRE(!xb && dyb && (b->dx || b->dx_bf16), "lnr_bwd: dxsum needs fp32 x, bf16 dy and a dx output");
    LNR_SHAPES_B(false, true, true, <<<(int)blocks, 256, smem, stream>>>(a))
  } else if (xb) {
    if (dyb) { LNR_SHAPES_B(true, true, false, <<<(int)blocks, 256, smem, stream>>>(a)) } else { LNR_SHAPES_B(true, false, false, <<<(int)blocks, 256, smem, stream>>>(a)) }
  } else {
    if (dyb) { LNR_SHAPES_B(false, true, false, <<<(int)blocks, 256, smem, stream>>>(a)) } else { LNR_SHAPES_B(false, false, false, <<<(int)blocks, 256, smem, stream>>>(a)) }
  }
  return after_launch("lnr_bwd_kernel");
}
