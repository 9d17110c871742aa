// LayerNorm family (HBM-bound, fp32 statistics), one warp per row, vectorised 128-bit accesses.
//
// Forward variants fold the layout work that the reference does with separate full-tensor copies
// into the load/store of the normalisation itself:
//   * window gather  : LN(norm1) + F.pad + torch.roll + window_partition
//                      (swin_transformer_3d.py:450-466, 271-283)  -> bf16 windows (B_,N,C)
//   * merge gather   : PatchMerging 2x2 strided slices + cat + LN(4C)   (:521-541)
//   * additive terms : fusion encoder's  + vis_space_pos + vis_tempor_pos + token_type  then LN,
//                      written straight into the concatenated [video ; text] buffer
//                      (cross_transformer.py:84-108)
//   * mask-token blend after the patch-embed LN  (:222-230)
// Backward mirrors them (scatter instead of gather) and can add the residual-stream gradient and
// emit a bf16 copy (optionally in window order) for the next GEMM's operand.
#include <algorithm>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

struct LnArgs {
  const void* x; int x_bf16; long long ld_x;
  const float* gamma; const float* beta; float eps;
  void* y; int y_bf16; long long ld_y;
  float* mean; float* rstd;
  long long rows; int C;                  // rows = number of OUTPUT rows, C = normalised width
  int mode;                               // 0 plain, 1 window gather, 2 merge gather
  WindowGeom geom;                        // mode 1
  int mB, mD, mH, mW, mC;                 // mode 2: input (B,D,H,W,mC), C == 4*mC
  const float* add0;                      // [C] broadcast
  const float* add1; int div1, mod1;      // [mod1, C] indexed (row / div1) % mod1
  const float* add2; int div2, mod2;
  long long group_rows, group_stride, row_offset;   // output row = (r / group_rows)*group_stride + r % group_rows + row_offset
  const long long* blend_mask; const float* blend_token; int bH, bW, mh, mw, bD;   // mask (B,mh,mw) int64
  const long long* row_index;             // mode 0: source row = row_index[r] (embedding lookup)
};

CLV_DEVICE float4 ld4(const void* base, int is_bf16, long long elem_off) {
  if (is_bf16) {
    uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
}
CLV_DEVICE void st4(void* base, int is_bf16, long long elem_off, float4 v) {
  if (is_bf16) {
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) =
        make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off) = v;
  }
}

// Source (row, column) of vector `i` (4 elements) of output row r; false for zero padding.
CLV_DEVICE bool ln_src(const LnArgs& a, long long r, long long src_row, int i, long long& srow, int& col) {
  if (a.mode == 2) {
    const int e = i * 4;
    const int part = e / a.mC;
    col = e % a.mC;
    const unsigned W2 = (a.mW + 1) / 2, H2 = (a.mH + 1) / 2;
    const unsigned r32 = (unsigned)r;
    int w2 = (int)(r32 % W2); unsigned t = r32 / W2;
    int h2 = (int)(t % H2); t /= H2;
    int d = (int)(t % (unsigned)a.mD); int b = (int)(t / (unsigned)a.mD);
    // PatchMerging channel blocks: [(even h, even w), (odd h, even w), (even h, odd w), (odd h, odd w)]
    const int h = 2 * h2 + (part & 1), w = 2 * w2 + (part >> 1);
    if (h >= a.mH || w >= a.mW) return false;
    srow = (((long long)b * a.mD + d) * a.mH + h) * a.mW + w;
    return true;
  }
  col = i * 4;
  srow = src_row;
  return src_row >= 0;
}

CLV_DEVICE long long ln_out_row(const LnArgs& a, long long r) {
  if (a.group_rows > 0) {
    const unsigned r32 = (unsigned)r, gr = (unsigned)a.group_rows;
    return (long long)(r32 / gr) * a.group_stride + (r32 % gr) + a.row_offset;
  }
  return r;
}

CLV_DEVICE float ln_blend_weight(const LnArgs& a, long long r) {
  // row r = (b, d, h, w) of the (B, bD, bH, bW) token grid; mask (B, mh, mw)
  const unsigned r32 = (unsigned)r;
  int w_ = (int)(r32 % (unsigned)a.bW); unsigned t = r32 / (unsigned)a.bW;
  int h_ = (int)(t % (unsigned)a.bH); t /= (unsigned)a.bH;
  int b_ = (int)(t / (unsigned)a.bD);
  return (float)a.blend_mask[((long long)b_ * a.mh + h_ / (a.bH / a.mh)) * a.mw + w_ / (a.bW / a.mw)];
}

CLV_DEVICE float4 ln_load_x(const LnArgs& a, long long r, long long src_row, int i) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  long long srow; int col;
  if (ln_src(a, r, src_row, i, srow, col)) v = ld4(a.x, a.x_bf16, srow * a.ld_x + col);
  if (a.add0) { float4 t = *reinterpret_cast<const float4*>(a.add0 + i * 4); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
  if (a.add1) { float4 t = *reinterpret_cast<const float4*>(a.add1 + (((unsigned)r / (unsigned)a.div1) % (unsigned)a.mod1) * a.C + i * 4); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
  if (a.add2) { float4 t = *reinterpret_cast<const float4*>(a.add2 + (((unsigned)r / (unsigned)a.div2) % (unsigned)a.mod2) * a.C + i * 4); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
  return v;
}

// One warp normalises RPW rows per iteration (RPW > 1 for narrow rows keeps enough bytes in flight per warp).
template <int VPL, int RPW>
__global__ void __launch_bounds__(256) ln_fwd_kernel(LnArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = a.C >> 2;
  const float inv_c = 1.0f / (float)a.C;
  for (long long r0 = warp_global * RPW; r0 < a.rows; r0 += nwarps * RPW) {
    float4 v[RPW][VPL];
    long long src_row[RPW];
    bool live[RPW], pad[RPW];
    float sum[RPW];
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      const long long r = r0 + q;
      live[q] = r < a.rows;
      src_row[q] = live[q] ? (a.row_index ? a.row_index[r] : r) : 0;
      if (a.mode == 1 && live[q]) src_row[q] = window_row_to_src(a.geom, r);
      pad[q] = a.mode == 1 && src_row[q] < 0;      // zero padding is applied AFTER norm1 in the reference
      sum[q] = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = lane + 32 * j;
        v[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live[q] && !pad[q] && i < nvec) {
          v[q][j] = ln_load_x(a, r, src_row[q], i);
          sum[q] += v[q][j].x + v[q][j].y + v[q][j].z + v[q][j].w;
        }
      }
    }
    float mu[RPW], rs[RPW];
#pragma unroll
    for (int q = 0; q < RPW; ++q) mu[q] = warp_sum(sum[q]) * inv_c;
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec) {
          const float dx = v[q][j].x - mu[q], dy = v[q][j].y - mu[q], dz = v[q][j].z - mu[q], dw = v[q][j].w - mu[q];
          sq += dx * dx + dy * dy + dz * dz + dw * dw;
        }
      }
      rs[q] = rsqrtf(warp_sum(sq) * inv_c + a.eps);
    }
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      const long long r = r0 + q;
      if (!live[q]) continue;
      const long long orow = ln_out_row(a, r);
      if (pad[q]) {
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          const int i = lane + 32 * j;
          if (i < nvec) st4(a.y, a.y_bf16, orow * a.ld_y + (long long)i * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (lane == 0 && a.mean) { a.mean[r] = 0.f; a.rstd[r] = 0.f; }
        continue;
      }
      const float bw = a.blend_mask ? ln_blend_weight(a, r) : 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec) {
          const float4 g = *reinterpret_cast<const float4*>(a.gamma + i * 4);
          const float4 b = *reinterpret_cast<const float4*>(a.beta + i * 4);
          float4 o;
          o.x = (v[q][j].x - mu[q]) * rs[q] * g.x + b.x; o.y = (v[q][j].y - mu[q]) * rs[q] * g.y + b.y;
          o.z = (v[q][j].z - mu[q]) * rs[q] * g.z + b.z; o.w = (v[q][j].w - mu[q]) * rs[q] * g.w + b.w;
          if (a.blend_mask) {
            const float4 tk = *reinterpret_cast<const float4*>(a.blend_token + i * 4);
            o.x = o.x * (1.f - bw) + tk.x * bw; o.y = o.y * (1.f - bw) + tk.y * bw;
            o.z = o.z * (1.f - bw) + tk.z * bw; o.w = o.w * (1.f - bw) + tk.w * bw;
          }
          st4(a.y, a.y_bf16, orow * a.ld_y + (long long)i * 4, o);
        }
      }
      if (lane == 0 && a.mean) { a.mean[r] = mu[q]; a.rstd[r] = rs[q]; }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct LnBwdArgs {
  LnArgs f;                                // forward description (x, gamma, mean, rstd, geometry, adds, blend)
  const void* dy; int dy_bf16; long long ld_dy;     // indexed by ln_out_row(r)
  float* dx; long long ld_dx;              // fp32, at the SOURCE rows
  const float* dres; long long ld_dres;    // optional residual-stream gradient added to dx (source rows)
  void* dx_copy; int dx_copy_bf16; long long ld_copy;  // optional copy of the final dx
  int copy_window_map; WindowGeom copy_geom;           // copy row = src_row_to_window(copy_geom, s)
  float* dgamma; float* dbeta;             // [C] fp32, atomically accumulated (caller zero-fills)
  float* dtoken;                           // [C] d(mask_token), blend only
  int dx_dense;                            // write dx (and copy) at row r instead of the source row
};

template <int VPL, int RPW>
__global__ void __launch_bounds__(128) ln_bwd_kernel(LnBwdArgs a) {
  const LnArgs& f = a.f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = f.C >> 2;
  const float inv_c = 1.0f / (float)f.C;
  float4 dg[VPL], db[VPL], dt[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) dg[j] = db[j] = dt[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (long long r0 = warp_global * RPW; r0 < f.rows; r0 += nwarps * RPW) {
    float4 xh[RPW][VPL], gy[RPW][VPL];
    long long src_row[RPW];
    bool live[RPW];
    float s1[RPW], s2[RPW], rsv[RPW];
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      const long long r = r0 + q;
      live[q] = r < f.rows;
      src_row[q] = live[q] ? (f.row_index ? f.row_index[r] : r) : 0;
      if (f.mode == 1 && live[q]) {
        src_row[q] = window_row_to_src(f.geom, r);
        if (src_row[q] < 0) live[q] = false;       // padding rows carry no gradient
      }
      s1[q] = s2[q] = 0.f;
      rsv[q] = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) xh[q][j] = gy[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!live[q]) continue;
      const long long orow = ln_out_row(f, r);
      const float mu = f.mean[r], rs = f.rstd[r];
      rsv[q] = rs;
      const float bw = f.blend_mask ? ln_blend_weight(f, r) : 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec) {
          const float4 v = ln_load_x(f, r, src_row[q], i);
          xh[q][j] = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
          float4 d = ld4(a.dy, a.dy_bf16, orow * a.ld_dy + (long long)i * 4);
          if (f.blend_mask) {
            dt[j].x += d.x * bw; dt[j].y += d.y * bw; dt[j].z += d.z * bw; dt[j].w += d.w * bw;
            d.x *= (1.f - bw); d.y *= (1.f - bw); d.z *= (1.f - bw); d.w *= (1.f - bw);
          }
          dg[j].x += d.x * xh[q][j].x; dg[j].y += d.y * xh[q][j].y; dg[j].z += d.z * xh[q][j].z; dg[j].w += d.w * xh[q][j].w;
          db[j].x += d.x; db[j].y += d.y; db[j].z += d.z; db[j].w += d.w;
          const float4 g = *reinterpret_cast<const float4*>(f.gamma + i * 4);
          gy[q][j] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
          s1[q] += gy[q][j].x * xh[q][j].x + gy[q][j].y * xh[q][j].y + gy[q][j].z * xh[q][j].z + gy[q][j].w * xh[q][j].w;
          s2[q] += gy[q][j].x + gy[q][j].y + gy[q][j].z + gy[q][j].w;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RPW; ++q) { s1[q] = warp_sum(s1[q]) * inv_c; s2[q] = warp_sum(s2[q]) * inv_c; }
    if (!a.dx) continue;
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
      if (!live[q]) continue;
      const long long r = r0 + q;
      const float rs = rsv[q];
      long long crow_w = 0;
      if (a.dx_copy && a.copy_window_map && f.mode != 2) crow_w = src_row_to_window(a.copy_geom, a.dx_dense ? r : src_row[q]);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec) {
          float4 o;
          o.x = rs * (gy[q][j].x - s2[q] - xh[q][j].x * s1[q]); o.y = rs * (gy[q][j].y - s2[q] - xh[q][j].y * s1[q]);
          o.z = rs * (gy[q][j].z - s2[q] - xh[q][j].z * s1[q]); o.w = rs * (gy[q][j].w - s2[q] - xh[q][j].w * s1[q]);
          long long srow; int col;
          if (!ln_src(f, r, src_row[q], i, srow, col)) continue;
          if (a.dx_dense) srow = r;
          if (a.dres) {
            const float4 rr = *reinterpret_cast<const float4*>(a.dres + srow * a.ld_dres + col);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          *reinterpret_cast<float4*>(a.dx + srow * a.ld_dx + col) = o;
          if (a.dx_copy) {
            long long crow = srow;
            if (a.copy_window_map) crow = (f.mode != 2) ? crow_w : src_row_to_window(a.copy_geom, srow);
            st4(a.dx_copy, a.dx_copy_bf16, crow * a.ld_copy + col, o);
          }
        }
      }
    }
  }

  // reduce dgamma / dbeta (/ dtoken) across the CTA's warps, then one atomic per column per CTA
  extern __shared__ float red[];   // [nwarps_in_block][C]
  const int nw = blockDim.x >> 5;
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? a.dgamma : (pass == 1 ? a.dbeta : a.dtoken);
    if (!dst) continue;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) *reinterpret_cast<float4*>(red + wib * f.C + i * 4) = pass == 0 ? dg[j] : (pass == 1 ? db[j] : dt[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += red[w * f.C + c];
      atomicAdd(dst + c, s);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// PatchMerging LayerNorm (swin_transformer_3d.py:535-542) for the Swin widths 4C = 512 / 1024 / 2048.
// The generic kernels above give a whole 4C-wide row to one warp (16 float4 per lane at 4C = 2048, plus as many
// accumulators in the backward: > 255 registers, local-memory spills, 0.5-1.5 TB/s).  Here a merged row belongs to
// T = C/4 threads (1, 2 or 4 warps) and every thread holds ONE float4 of each of the four 2x2 neighbours:
// vector j of thread t is column j*C + 4t of the merged row = column 4t of source row (2*h2 + (j & 1), 2*w2 + (j >> 1)),
// so all loads / stores are contiguous C-wide rows of x / dx.  Row statistics cross the warps of a row through shared memory.
// ------------------------------------------------------------------------------------------
struct LnmArgs {
  const float* x; const float* gamma; const float* beta; float eps;
  float* mean; float* rstd;
  unsigned rows; int C;                 // merged rows, source width (merged width 4C)
  int D, H, W;                          // source token grid per clip
  __nv_bfloat16* y;                     // forward: [rows, 4C]
  const __nv_bfloat16* dy; float* dx;   // backward: dy [rows, 4C]; dx at the source rows [B*D*H*W, C]
  float* dgamma; float* dbeta;
  __nv_bfloat16* dx16; float* dxsum;    // optional bf16 copy of dx and its column sums [C], both times copy_scale[clip]
  const float* copy_scale; unsigned copy_scale_rows;
};

// source rows of the four channel blocks of merged row r: [(even h, even w), (odd h, even w), (even h, odd w), (odd h, odd w)];
// -1 = zero padding (odd H / W)
CLV_DEVICE void lnm_sources(const LnmArgs& a, unsigned r, long long (&src)[4]) {
  const unsigned W2 = (a.W + 1) / 2, H2 = (a.H + 1) / 2;
  const unsigned w2 = r % W2; unsigned t = r / W2;
  const unsigned h2 = t % H2; t /= H2;             // t = b * D + d
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int h = 2 * (int)h2 + (j & 1), w = 2 * (int)w2 + (j >> 1);
    src[j] = (h < a.H && w < a.W) ? ((long long)t * a.H + h) * a.W + w : -1;
  }
}

CLV_DEVICE void lnm_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// ask L2 for the lines of the merged row this thread's row group handles NEXT iteration (one thread per 128-byte line): the
// kernels are latency-bound on their loads, this doubles the bytes in flight without a live register
template <bool WITH_DY>
CLV_DEVICE void lnm_prefetch_next(const LnmArgs& a, unsigned rn, int t) {
  if (rn >= a.rows) return;
  if ((t & 7) == 0) {
    long long sn[4];
    lnm_sources(a, rn, sn);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (sn[j] >= 0) lnm_prefetch_l2(a.x + sn[j] * a.C + 4 * t);
  }
  if (WITH_DY && (t & 15) == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) lnm_prefetch_l2(a.dy + (size_t)rn * 4 * a.C + j * a.C + 4 * t);
  }
}

template <int T>
CLV_DEVICE float lnm_row_sum(float v, float* slot, int warp_in_row) {
  v = warp_sum(v);
  if constexpr (T == 32) return v;
  if ((threadIdx.x & 31) == 0) slot[warp_in_row] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) s += slot[w];
  return s;
}

template <int T>
__global__ void __launch_bounds__(256) lnm_fwd_kernel(LnmArgs a) {
  constexpr int R = 256 / T;                       // merged rows per CTA iteration
  __shared__ float part[2][R][4];
  const int g = threadIdx.x / T, t = threadIdx.x % T, wir = t >> 5;
  const float inv_c = 1.0f / (float)(4 * a.C);
  float4 gam[4], bet[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    gam[j] = __ldg(reinterpret_cast<const float4*>(a.gamma + j * a.C + 4 * t));
    bet[j] = __ldg(reinterpret_cast<const float4*>(a.beta + j * a.C + 4 * t));
  }
  for (unsigned base = blockIdx.x * R; base < a.rows; base += gridDim.x * R) {
    const unsigned r = base + g;
    const bool live = r < a.rows;
    long long src[4] = {-1, -1, -1, -1};
    if (live) lnm_sources(a, r, src);
    lnm_prefetch_next<false>(a, r + gridDim.x * R, t);
    float4 v[4];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = src[j] >= 0 ? *reinterpret_cast<const float4*>(a.x + src[j] * a.C + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mu = lnm_row_sum<T>(sum, part[0][g], wir) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j].x -= mu; v[j].y -= mu; v[j].z -= mu; v[j].w -= mu;
      sq = fmaf(v[j].x, v[j].x, sq); sq = fmaf(v[j].y, v[j].y, sq); sq = fmaf(v[j].z, v[j].z, sq); sq = fmaf(v[j].w, v[j].w, sq);
    }
    const float rs = rsqrtf(lnm_row_sum<T>(sq, part[1][g], wir) * inv_c + a.eps);
    if (!live) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 o;
      o.x = fmaf(v[j].x * rs, gam[j].x, bet[j].x); o.y = fmaf(v[j].y * rs, gam[j].y, bet[j].y);
      o.z = fmaf(v[j].z * rs, gam[j].z, bet[j].z); o.w = fmaf(v[j].w * rs, gam[j].w, bet[j].w);
      *reinterpret_cast<uint2*>(a.y + (size_t)r * 4 * a.C + j * a.C + 4 * t) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
    if (t == 0) { a.mean[r] = mu; a.rstd[r] = rs; }
  }
}

template <int T>
__global__ void __launch_bounds__(256, 2) lnm_bwd_kernel(LnmArgs a) {
  constexpr int R = 256 / T;
  extern __shared__ float red[];                   // [2][4C]: dgamma | dbeta of this CTA
  __shared__ float part[2][2][R][4];               // [iteration parity][s1 | s2][row][warp of the row]
  const int g = threadIdx.x / T, t = threadIdx.x % T, wir = t >> 5;
  const int C4 = 4 * a.C;
  const float inv_c = 1.0f / (float)C4;
  for (int c = threadIdx.x; c < 2 * C4; c += blockDim.x) red[c] = 0.f;
  float4 gam[4], dg[4], db[4];
  float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);       // column sums of the scaled dx: the four neighbours share their columns
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    gam[j] = __ldg(reinterpret_cast<const float4*>(a.gamma + j * a.C + 4 * t));
    dg[j] = db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int par = 0;
  for (unsigned base = blockIdx.x * R; base < a.rows; base += gridDim.x * R, par ^= 1) {
    const unsigned r = base + g;
    const bool live = r < a.rows;
    long long src[4] = {-1, -1, -1, -1};
    float mu = 0.f, rs = 0.f;
    if (live) { lnm_sources(a, r, src); mu = a.mean[r]; rs = a.rstd[r]; }
    lnm_prefetch_next<true>(a, r + gridDim.x * R, t);
    float4 xh[4], d[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      xh[j] = d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        if (src[j] >= 0) xh[j] = *reinterpret_cast<const float4*>(a.x + src[j] * a.C + 4 * t);
        const uint2 u = *reinterpret_cast<const uint2*>(a.dy + (size_t)r * C4 + j * a.C + 4 * t);
        const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
        d[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
        // a padded neighbour is a real (zero) element of the normalised row: it takes part in the statistics and in dgamma
        xh[j].x = (xh[j].x - mu) * rs; xh[j].y = (xh[j].y - mu) * rs; xh[j].z = (xh[j].z - mu) * rs; xh[j].w = (xh[j].w - mu) * rs;
      }
      dg[j].x = fmaf(d[j].x, xh[j].x, dg[j].x); dg[j].y = fmaf(d[j].y, xh[j].y, dg[j].y);
      dg[j].z = fmaf(d[j].z, xh[j].z, dg[j].z); dg[j].w = fmaf(d[j].w, xh[j].w, dg[j].w);
      db[j].x += d[j].x; db[j].y += d[j].y; db[j].z += d[j].z; db[j].w += d[j].w;
      d[j].x *= gam[j].x; d[j].y *= gam[j].y; d[j].z *= gam[j].z; d[j].w *= gam[j].w;       // d <- dy * gamma
      s1 = fmaf(d[j].x, xh[j].x, s1); s1 = fmaf(d[j].y, xh[j].y, s1); s1 = fmaf(d[j].z, xh[j].z, s1); s1 = fmaf(d[j].w, xh[j].w, s1);
      s2 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if constexpr (T > 32) {
      if ((threadIdx.x & 31) == 0) { part[par][0][g][wir] = s1; part[par][1][g][wir] = s2; }
      __syncthreads();
      s1 = s2 = 0.f;
#pragma unroll
      for (int w = 0; w < T / 32; ++w) { s1 += part[par][0][g][w]; s2 += part[par][1][g][w]; }
    }
    s1 *= inv_c; s2 *= inv_c;
    if (!live) continue;
    float sc = 1.0f;
    if (a.copy_scale) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (src[j] >= 0) { sc = __ldg(a.copy_scale + (unsigned)src[j] / a.copy_scale_rows); break; }   // one clip per merged row
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (src[j] < 0) continue;
      float4 o;
      o.x = rs * (d[j].x - s2 - xh[j].x * s1); o.y = rs * (d[j].y - s2 - xh[j].y * s1);
      o.z = rs * (d[j].z - s2 - xh[j].z * s1); o.w = rs * (d[j].w - s2 - xh[j].w * s1);
      *reinterpret_cast<float4*>(a.dx + src[j] * a.C + 4 * t) = o;
      if (a.dx16) {
        o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc;
        *reinterpret_cast<uint2*>(a.dx16 + src[j] * a.C + 4 * t) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        ds.x += o.x; ds.y += o.y; ds.z += o.z; ds.w += o.w;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = j * a.C + 4 * t;
    atomicAdd(red + c, dg[j].x); atomicAdd(red + c + 1, dg[j].y); atomicAdd(red + c + 2, dg[j].z); atomicAdd(red + c + 3, dg[j].w);
    atomicAdd(red + C4 + c, db[j].x); atomicAdd(red + C4 + c + 1, db[j].y);
    atomicAdd(red + C4 + c + 2, db[j].z); atomicAdd(red + C4 + c + 3, db[j].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    atomicAdd(a.dgamma + c, red[c]);
    atomicAdd(a.dbeta + c, red[C4 + c]);
  }
  if (a.dxsum) {                                   // fold the row groups' column sums through the (now free) staging area
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) red[c] = 0.f;
    __syncthreads();
    atomicAdd(red + 4 * t, ds.x); atomicAdd(red + 4 * t + 1, ds.y); atomicAdd(red + 4 * t + 2, ds.z); atomicAdd(red + 4 * t + 3, ds.w);
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) atomicAdd(a.dxsum + c, red[c]);
  }
}

static bool lnm_eligible(const LnArgs& f) {
  return f.mode == 2 && !f.x_bf16 && f.ld_x == f.mC && (f.mC == 128 || f.mC == 256 || f.mC == 512) && !f.add0 && !f.add1 &&
         !f.add2 && f.group_rows == 0 && !f.blend_mask && !f.row_index && f.mean && f.rstd;
}
static LnmArgs lnm_args(const LnArgs& f) {
  LnmArgs m{};
  m.x = reinterpret_cast<const float*>(f.x); m.gamma = f.gamma; m.beta = f.beta; m.eps = f.eps; m.mean = f.mean; m.rstd = f.rstd;
  m.rows = (unsigned)f.rows; m.C = f.mC; m.D = f.mD; m.H = f.mH; m.W = f.mW;
  return m;
}

static void fill_geom(WindowGeom& g, const clv_window_geom_t* w) {
  g.B = w->B; g.D = w->D; g.H = w->H; g.W = w->W; g.wd = w->wd; g.wh = w->wh; g.ww = w->ww;
  g.sd = w->sd; g.sh = w->sh; g.sw = w->sw;
  g.Dp = (w->D + w->wd - 1) / w->wd * w->wd; g.Hp = (w->H + w->wh - 1) / w->wh * w->wh;
  g.Wp = (w->W + w->ww - 1) / w->ww * w->ww;
  g.nD = g.Dp / g.wd; g.nH = g.Hp / g.wh; g.nW = g.Wp / g.ww; g.N = g.wd * g.wh * g.ww; g.nWin = g.nD * g.nH * g.nW;
}

static int build_ln_args(LnArgs& a, const clv_ln_desc_t* d) {
  CLV_REQUIRE(d && d->x && d->gamma && d->beta, "layernorm: null pointer");
  CLV_REQUIRE(d->C > 0 && d->C % 4 == 0 && d->C <= 4096, "layernorm: C must be a multiple of 4 and <= 4096 (got %d)", d->C);
  a = LnArgs{};
  a.x = d->x; a.x_bf16 = d->x_is_bf16; a.ld_x = d->ld_x;
  a.gamma = d->gamma; a.beta = d->beta; a.eps = d->eps;
  a.mean = d->mean; a.rstd = d->rstd; a.rows = d->rows; a.C = d->C;
  a.mode = 0;
  if (d->window) { a.mode = 1; fill_geom(a.geom, d->window);
    CLV_REQUIRE((long long)a.geom.B * a.geom.nWin * a.geom.N == d->rows, "layernorm: window geometry/rows mismatch"); }
  if (d->merge_C > 0) {
    CLV_REQUIRE(!d->window, "layernorm: window and merge gathers are exclusive");
    a.mode = 2; a.mB = d->merge_B; a.mD = d->merge_D; a.mH = d->merge_H; a.mW = d->merge_W; a.mC = d->merge_C;
    CLV_REQUIRE(d->C == 4 * d->merge_C && d->merge_C % 4 == 0, "layernorm: merge gather needs C == 4*merge_C");
    CLV_REQUIRE((long long)a.mB * a.mD * ((a.mH + 1) / 2) * ((a.mW + 1) / 2) == d->rows, "layernorm: merge rows mismatch");
  }
  a.add0 = d->add0; a.add1 = d->add1; a.div1 = d->div1 > 0 ? d->div1 : 1; a.mod1 = d->mod1 > 0 ? d->mod1 : 1;
  a.add2 = d->add2; a.div2 = d->div2 > 0 ? d->div2 : 1; a.mod2 = d->mod2 > 0 ? d->mod2 : 1;
  a.group_rows = d->group_rows; a.group_stride = d->group_stride; a.row_offset = d->row_offset;
  CLV_REQUIRE(d->rows < 2000000000LL, "layernorm: more than 2^31 rows");
  a.row_index = d->row_index;
  CLV_REQUIRE(!a.row_index || a.mode == 0, "layernorm: row_index only with the plain mode");
  a.blend_mask = d->blend_mask; a.blend_token = d->blend_token;
  a.bD = d->blend_D; a.bH = d->blend_H; a.bW = d->blend_W; a.mh = d->blend_mh; a.mw = d->blend_mw;
  if (a.blend_mask) CLV_REQUIRE(a.blend_token && a.bH > 0 && a.mh > 0 && a.bH % a.mh == 0 && a.bW % a.mw == 0,
                                "layernorm: bad mask-token blend geometry");
  return 0;
}

// rows per warp iteration: 4 for C <= 128, 2 for C <= 256 (narrow rows are latency-bound otherwise)
#define LN_DISPATCH(KERNEL, nvec_per_lane, ...)                    \
  switch (nvec_per_lane) {                                         \
    case 1: KERNEL<1, 4> __VA_ARGS__; break;                       \
    case 2: KERNEL<2, 2> __VA_ARGS__; break;                       \
    case 3: case 4: KERNEL<4, 1> __VA_ARGS__; break;               \
    case 5: case 6: case 7: case 8: KERNEL<8, 1> __VA_ARGS__; break;  \
    case 9: case 10: case 11: case 12: case 13: case 14: case 15: case 16: KERNEL<16, 1> __VA_ARGS__; break; \
    default: KERNEL<32, 1> __VA_ARGS__; break;                     \
  }

static int ln_rows_per_warp(int vpl) { return vpl == 1 ? 4 : (vpl == 2 ? 2 : 1); }

}  // namespace clv

using namespace clv;

extern "C" int clv_layernorm_fwd(const clv_ln_desc_t* d, void* y, int y_is_bf16, long long ld_y, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LnArgs a;
  if (int rc = build_ln_args(a, d)) return rc;
  CLV_REQUIRE(y != nullptr, "layernorm_fwd: null output");
  a.y = y; a.y_bf16 = y_is_bf16; a.ld_y = ld_y;
  if (a.rows == 0) return 0;
  if (lnm_eligible(a) && y_is_bf16 && ld_y == a.C) {
    LnmArgs m = lnm_args(a);
    m.y = reinterpret_cast<__nv_bfloat16*>(y);
    const int T = a.mC / 4, R = 256 / T;
    const int blocks = (int)std::min<long long>((a.rows + R - 1) / R, (long long)num_sms() * 8);
    if (T == 32) lnm_fwd_kernel<32><<<blocks, 256, 0, stream>>>(m);
    else if (T == 64) lnm_fwd_kernel<64><<<blocks, 256, 0, stream>>>(m);
    else lnm_fwd_kernel<128><<<blocks, 256, 0, stream>>>(m);
    return after_launch("lnm_fwd_kernel launch");
  }
  const int vpl = (a.C / 4 + 31) / 32;
  const int warps_per_block = 8;
  const long long rows_per_block = (long long)warps_per_block * ln_rows_per_warp(vpl);
  long long blocks = (a.rows + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  LN_DISPATCH(ln_fwd_kernel, vpl, <<<(int)blocks, warps_per_block * 32, 0, stream>>>(a));
  return after_launch("ln_fwd_kernel launch");
}

extern "C" int clv_layernorm_bwd(const clv_ln_desc_t* d, const clv_ln_bwd_t* b, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LnBwdArgs a{};
  if (int rc = build_ln_args(a.f, d)) return rc;
  CLV_REQUIRE(b && b->dy && d->mean && d->rstd, "layernorm_bwd: null pointer");
  a.dy = b->dy; a.dy_bf16 = b->dy_is_bf16; a.ld_dy = b->ld_dy;
  a.dx = b->dx; a.ld_dx = b->ld_dx; a.dres = b->dres; a.ld_dres = b->ld_dres;
  a.dx_copy = b->dx_copy; a.dx_copy_bf16 = b->dx_copy_is_bf16; a.ld_copy = b->ld_copy;
  a.copy_window_map = b->copy_window != nullptr;
  if (b->copy_window) fill_geom(a.copy_geom, b->copy_window);
  a.dgamma = b->dgamma; a.dbeta = b->dbeta; a.dtoken = b->dtoken; a.dx_dense = b->dx_dense;
  const bool lnm_copy_ok = !a.dx_copy || (a.dx_copy_bf16 && a.ld_copy == a.f.mC && !b->copy_window);
  CLV_REQUIRE(!b->dxsum || a.dx_copy, "layernorm_bwd: dxsum comes with the bf16 dx_copy");
  CLV_REQUIRE(!b->copy_scale || b->copy_scale_rows > 0, "layernorm_bwd: copy_scale needs copy_scale_rows > 0");
  CLV_REQUIRE(a.f.C <= 2048, "layernorm_bwd: C must be <= 2048 (got %d)", a.f.C);
  if (a.f.rows == 0) return 0;
  if (lnm_eligible(a.f) && a.dy_bf16 && a.ld_dy == a.f.C && a.dx && a.ld_dx == a.f.mC && !a.dres && lnm_copy_ok && !a.dtoken &&
      !a.dx_dense && a.dgamma && a.dbeta) {
    LnmArgs m = lnm_args(a.f);
    m.dy = reinterpret_cast<const __nv_bfloat16*>(a.dy); m.dx = a.dx; m.dgamma = a.dgamma; m.dbeta = a.dbeta;
    m.dx16 = reinterpret_cast<__nv_bfloat16*>(a.dx_copy); m.dxsum = b->dxsum;
    m.copy_scale = b->copy_scale; m.copy_scale_rows = b->copy_scale ? (unsigned)b->copy_scale_rows : 1u;
    const int T = a.f.mC / 4, R = 256 / T;
    const int blocks = (int)std::min<long long>((a.f.rows + R - 1) / R, (long long)num_sms() * 2);
    const size_t smem = 2 * (size_t)a.f.C * sizeof(float);
    if (T == 32) lnm_bwd_kernel<32><<<blocks, 256, smem, stream>>>(m);
    else if (T == 64) lnm_bwd_kernel<64><<<blocks, 256, smem, stream>>>(m);
    else lnm_bwd_kernel<128><<<blocks, 256, smem, stream>>>(m);
    return after_launch("lnm_bwd_kernel launch");
  }
  CLV_REQUIRE(!b->dxsum && !b->copy_scale, "layernorm_bwd: dxsum / copy_scale only with the merge gather at the Swin widths");
  const int vpl = (a.f.C / 4 + 31) / 32;
  const int warps_per_block = 4;
  const long long rows_per_block = (long long)warps_per_block * ln_rows_per_warp(vpl);
  long long blocks = (a.f.rows + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)warps_per_block * a.f.C * sizeof(float);
  LN_DISPATCH(ln_bwd_kernel, vpl, <<<(int)blocks, warps_per_block * 32, smem, stream>>>(a));
  return after_launch("ln_bwd_kernel launch");
}
