// HBM-bound helper kernels of the Clover hot path: casts, im2col-free patchify for PatchEmbed3D,
// grouped column sums (bias / positional-embedding gradients, average pooling), row copies with an
// affine term (fusion concat), broadcast adds and embedding-gradient scatter.
#include <algorithm>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

CLV_DEVICE float4 ldv4(const void* base, int is_bf16, long long off) {
  if (is_bf16) {
    uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
}
CLV_DEVICE void stv4(void* base, int is_bf16, long long off, float4 v) {
  if (is_bf16)
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  else
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = v;
}

// ---------------------------------------------------------------------------------------------
__global__ void cast_kernel(const void* src, int src_bf16, void* dst, int dst_bf16, long long n4, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = ldv4(src, src_bf16, i * 4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    stv4(dst, dst_bf16, i * 4, v);
  }
}

// y = GELU(x)  (mode 0)   or   y = dy * GELU'(x)  (mode 1); exact erf form (nn.GELU default)
// y = tanh(x)  (mode 2)   or   y = dy * (1 - x^2) with x = the saved tanh OUTPUT  (mode 3)   (nn.Tanh of ITMHead)
__global__ void gelu_kernel(const void* x, int x_bf16, const void* dy, int dy_bf16, void* y, int y_bf16, long long n4, int mode) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldv4(x, x_bf16, i * 4);
    float4 o;
    if (mode == 0) {
      o = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
    } else if (mode == 1) {
      const float4 d = ldv4(dy, dy_bf16, i * 4);
      o = make_float4(d.x * gelu_erf_grad(v.x), d.y * gelu_erf_grad(v.y), d.z * gelu_erf_grad(v.z), d.w * gelu_erf_grad(v.w));
    } else if (mode == 2) {
      o = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    } else {
      const float4 d = ldv4(dy, dy_bf16, i * 4);
      o = make_float4(d.x * (1.f - v.x * v.x), d.y * (1.f - v.y * v.y), d.z * (1.f - v.z * v.z), d.w * (1.f - v.w * v.w));
    }
    stv4(y, y_bf16, i * 4, o);
  }
}

// ---------------------------------------------------------------------------------------------
// PatchEmbed3D as a GEMM (swin_transformer_3d.py:665,671-681): Conv3d with kernel == stride is a
// [tokens, Cin*pd*ph*pw] x [Cin*pd*ph*pw, C] product.  One CTA stages the Cin*pd*ph input rows of
// one token row (b, d, hp) through shared memory (coalesced reads) and writes the patch matrix
// rows (column order (c, kd, kh, kw) = the conv weight's flattening) with coalesced bf16 stores.
// Out-of-range input (the F.pad of :675-680) reads as zero.
// TIN = float: clips normalised by the CPU pipeline (the shipped configs).  TIN = unsigned char: raw frames, normalised here
// as (v - mean[c]) * inv_std[c] -- the GPUNormalize module hook of utils/module_hooks.py:35-87 folded into the load, so the
// host->device copy carries 1 byte per sample instead of 4 and no separate normalisation pass runs.
template <typename TIN>
__global__ void patchify_kernel(const TIN* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int Cin, int F,
                                int H, int W, int pd, int ph, int pw, int D, int Hp, int Wp, const float* __restrict__ mean,
                                const float* __restrict__ inv_std) {
  extern __shared__ float tile[];            // [Cin*pd*ph][Wp*pw + 1]  (odd pitch: phase 2 walks the rows, not the columns)
  const int hp = blockIdx.x % Hp, d = (blockIdx.x / Hp) % D, b = blockIdx.x / (Hp * D);
  const int nrow = Cin * pd * ph, roww = Wp * pw, rs = roww + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // phase 1: one input row (c, kd, kh) per warp iteration, lanes along the contiguous W axis (index math once per row)
  for (int r = warp; r < nrow; r += nwarps) {
    const int kh = r % ph, kd = (r / ph) % pd, c = r / (ph * pd);
    const int f = d * pd + kd, h = hp * ph + kh;
    const bool row_ok = f < F && h < H;
    const TIN* src = x + (((long long)b * Cin + c) * F + f) * H * W + (long long)h * W;
    const float m = mean ? __ldg(mean + c) : 0.f, is = mean ? __ldg(inv_std + c) : 1.f;
    for (int col = lane; col < roww; col += 32) {
      float v = 0.f;
      if (row_ok && col < W) v = ((float)src[col] - m) * is;
      tile[r * rs + col] = v;
    }
  }
  __syncthreads();
  // phase 2: one token (wp) per warp iteration, lanes along the K = nrow * pw patch columns (pairs: pw is even)
  const int K = nrow * pw;
  const long long row0 = ((long long)(b * D + d) * Hp + hp) * Wp;
  for (int wp = warp; wp < Wp; wp += nwarps) {
    __nv_bfloat16* dst = out + (row0 + wp) * K;
    const float* srcw = tile + wp * pw;
    for (int k = lane * 2; k < K; k += 64) {
      const int r = k / pw, kw = k - r * pw;           // both elements of the pair share r (pw even, k even)
      const float* p = srcw + r * rs + kw;
      *reinterpret_cast<uint32_t*>(dst + k) = pack_bf16(p[0], p[1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[g, c] (+)= scale * sum over rows r with (r / div) % mod == g of x[r, c].
// grid = (col tiles of 128, row chunks, mod).  Used for nn.Linear bias gradients (mod = 1),
// AdaptiveAvgPool3d of ssl_head.py:105 (groups = clips), and the gradients of vis_space_pos /
// vis_tempor_pos / BERT position embeddings.
// bf16 input whose width / pitch is only a multiple of 4: 8-byte loads
__global__ void __launch_bounds__(256) grouped_colsum_bf16x4_kernel(const void* x, long long ld, long long rows, int C, int div,
                                                                     int mod, float scale, float* out) {
  const int g = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 32 + lane) * 4;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long full = rows / ((long long)div * mod);
  const long long rem = rows - full * div * mod;
  long long in_group = full * div;
  {
    const long long start = (long long)g * div;
    if (rem > start) in_group += min((long long)div, rem - start);
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c0 < C)
    for (long long k = (long long)blockIdx.y * nw + wib; k < in_group; k += (long long)gridDim.y * nw) {
      const long long r = ((k / div) * mod + g) * div + (k % div);
      const float4 v = ldv4(x, 1, r * ld + c0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  __shared__ float4 red[8][32];
  red[wib][lane] = acc;
  __syncthreads();
  if (wib == 0 && c0 < C) {
    for (int w = 1; w < nw; ++w) { const float4 v = red[w][lane]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    float* o = out + (long long)g * C + c0;
    atomicAdd(o + 0, acc.x * scale); atomicAdd(o + 1, acc.y * scale); atomicAdd(o + 2, acc.z * scale); atomicAdd(o + 3, acc.w * scale);
  }
}

template <int VEC>   // VEC columns per lane: 8 for bf16 input (16-byte loads), 4 for fp32
__global__ void __launch_bounds__(256) grouped_colsum_kernel(const void* x, long long ld, long long rows, int C, int div, int mod,
                                                              float scale, float* out) {
  const int g = blockIdx.z;
  const int lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 32 + lane) * VEC;
  const int wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // rows of group g: r = (q*mod + g)*div + t
  const long long full = rows / ((long long)div * mod);                 // complete periods
  const long long rem = rows - full * div * mod;
  long long in_group = full * div;
  {
    const long long start = (long long)g * div;
    if (rem > start) in_group += min((long long)div, rem - start);
  }
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  if (c0 < C) {
    const long long step = (long long)gridDim.y * nw;
#pragma unroll 4
    for (long long k = (long long)blockIdx.y * nw + wib; k < in_group; k += step) {
      const long long r = (mod == 1) ? k : ((k / div) * mod + g) * div + (k % div);
      if (VEC == 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + r * ld + c0);
        const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
        acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
        acc[4 % VEC] += f2.x; acc[5 % VEC] += f2.y; acc[6 % VEC] += f3.x; acc[7 % VEC] += f3.y;
      } else {
        const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + r * ld + c0);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      }
    }
  }
  __shared__ float red[8][32 * VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) red[wib][lane * VEC + e] = acc[e];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VEC; i += blockDim.x) {
    const int c = blockIdx.x * 32 * VEC + i;
    if (c < C) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += red[w][i];
      atomicAdd(out + (long long)g * C + c, s * scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// y[orow(r), :] = x[irow(r), :] + add0 + bvec[(r / bdiv), :] * bscale
// with grouped row maps  row(r) = (r / group_rows) * group_stride + r % group_rows + offset.
struct RowsAffineArgs {
  const void* x; int x_bf16; long long ld_x; long long in_group_rows, in_group_stride, in_offset;
  void* y; int y_bf16; long long ld_y; long long out_group_rows, out_group_stride, out_offset;
  const float* add0; const float* bvec; long long bdiv; float bscale;
  long long rows; int C;
};
__global__ void rows_affine_kernel(RowsAffineArgs a) {
  const int nvec = a.C / 4;
  const long long total = a.rows * nvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned r32 = (unsigned)(i / nvec); const long long r = r32; const int c = (int)(i % nvec) * 4;
    const long long ir = a.in_group_rows > 0 ? (long long)(r32 / (unsigned)a.in_group_rows) * a.in_group_stride + r32 % (unsigned)a.in_group_rows + a.in_offset : r;
    const long long orow = a.out_group_rows > 0 ? (long long)(r32 / (unsigned)a.out_group_rows) * a.out_group_stride + r32 % (unsigned)a.out_group_rows + a.out_offset : r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.x) v = ldv4(a.x, a.x_bf16, ir * a.ld_x + c);
    if (a.add0) { const float4 t = *reinterpret_cast<const float4*>(a.add0 + c); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    if (a.bvec) {
      const float4 t = *reinterpret_cast<const float4*>(a.bvec + (r / a.bdiv) * a.C + c);
      v.x += t.x * a.bscale; v.y += t.y * a.bscale; v.z += t.z * a.bscale; v.w += t.w * a.bscale;
    }
    stv4(a.y, a.y_bf16, orow * a.ld_y + c, v);
  }
}

// dst[index[r], :] += src[r, :]   (word-embedding gradient; duplicates allowed -> atomics)
__global__ void scatter_add_rows_kernel(const float* src, const long long* index, float* dst, long long rows, int C) {
  const long long total = rows * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / C; const int c = (int)(i % C);
    atomicAdd(dst + index[r] * C + c, src[i]);
  }
}

static int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace clv

using namespace clv;

extern "C" int clv_cast(const void* src, int src_is_bf16, void* dst, int dst_is_bf16, long long n, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(src && dst && n >= 0 && n % 4 == 0, "clv_cast: n must be a multiple of 4 (got %lld)", n);
  if (n == 0) return 0;
  cast_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(src, src_is_bf16, dst, dst_is_bf16, n / 4, scale);
  return after_launch("cast_kernel");
}

extern "C" int clv_gelu(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, void* y, int y_is_bf16, long long n,
                        void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && y && n >= 0 && n % 4 == 0, "clv_gelu: n must be a multiple of 4 (got %lld)", n);
  if (n == 0) return 0;
  gelu_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(x, x_is_bf16, dy, dy_is_bf16, y, y_is_bf16, n / 4, dy ? 1 : 0);
  return after_launch("gelu_kernel");
}

extern "C" int clv_tanh(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, void* y, int y_is_bf16, long long n,
                        void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && y && n >= 0 && n % 4 == 0, "clv_tanh: n must be a multiple of 4 (got %lld)", n);
  if (n == 0) return 0;
  gelu_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(x, x_is_bf16, dy, dy_is_bf16, y, y_is_bf16, n / 4, dy ? 3 : 2);
  return after_launch("gelu_kernel(tanh)");
}

extern "C" int clv_patchify(const float* x, void* out_bf16, int B, int Cin, int F, int H, int W, int pd, int ph, int pw,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && out_bf16 && B > 0 && pw % 2 == 0, "clv_patchify: bad arguments");
  const int D = (F + pd - 1) / pd, Hp = (H + ph - 1) / ph, Wp = (W + pw - 1) / pw;
  const size_t smem = (size_t)Cin * pd * ph * (Wp * pw + 1) * sizeof(float);
  CLV_REQUIRE(smem <= 200 * 1024, "clv_patchify: row tile too large (%zu bytes)", smem);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(patchify_kernel<float>), (int)smem)) return rc;
  patchify_kernel<float><<<B * D * Hp, 256, smem, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, Cin, F, H, W, pd,
                                                           ph, pw, D, Hp, Wp, nullptr, nullptr);
  return after_launch("patchify_kernel");
}

extern "C" int clv_patchify_u8(const unsigned char* x, const float* mean, const float* inv_std, void* out_bf16, int B, int Cin,
                               int F, int H, int W, int pd, int ph, int pw, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && mean && inv_std && out_bf16 && B > 0 && pw % 2 == 0, "clv_patchify_u8: bad arguments");
  const int D = (F + pd - 1) / pd, Hp = (H + ph - 1) / ph, Wp = (W + pw - 1) / pw;
  const size_t smem = (size_t)Cin * pd * ph * (Wp * pw + 1) * sizeof(float);
  CLV_REQUIRE(smem <= 200 * 1024, "clv_patchify_u8: row tile too large (%zu bytes)", smem);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(patchify_kernel<unsigned char>), (int)smem)) return rc;
  patchify_kernel<unsigned char><<<B * D * Hp, 256, smem, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, Cin, F, H, W,
                                                                   pd, ph, pw, D, Hp, Wp, mean, inv_std);
  return after_launch("patchify_kernel<u8>");
}

extern "C" int clv_grouped_colsum(const void* x, int x_is_bf16, long long ld, long long rows, int C, int div, int mod,
                                  float scale, float* out, int accumulate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(x && out && C > 0 && C % 4 == 0 && div > 0 && mod > 0, "clv_grouped_colsum: bad arguments");
  if (!accumulate) CLV_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)mod * C * sizeof(float), stream));
  if (rows == 0) return 0;
  const bool wide = x_is_bf16 && C % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int vec = wide ? 8 : 4;
  const int col_tiles = (C + 32 * vec - 1) / (32 * vec);
  const long long per_group = (rows + mod - 1) / mod;
  long long chunks = (per_group + 63) / 64;
  const long long cap = std::max<long long>(1, (long long)num_sms() * 8 / ((long long)col_tiles * mod));
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid(col_tiles, (unsigned)chunks, mod);
  if (wide) {
    grouped_colsum_kernel<8><<<grid, 256, 0, stream>>>(x, ld, rows, C, div, mod, scale, out);
  } else if (!x_is_bf16) {
    grouped_colsum_kernel<4><<<grid, 256, 0, stream>>>(x, ld, rows, C, div, mod, scale, out);
  } else {
    grouped_colsum_bf16x4_kernel<<<grid, 256, 0, stream>>>(x, ld, rows, C, div, mod, scale, out);
  }
  return after_launch("grouped_colsum_kernel");
}

extern "C" int clv_rows_affine(const clv_rows_affine_t* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(d && d->y && d->C > 0 && d->C % 4 == 0, "clv_rows_affine: bad arguments");
  if (d->rows == 0) return 0;
  RowsAffineArgs a{};
  a.x = d->x; a.x_bf16 = d->x_is_bf16; a.ld_x = d->ld_x;
  a.in_group_rows = d->in_group_rows; a.in_group_stride = d->in_group_stride; a.in_offset = d->in_offset;
  a.y = d->y; a.y_bf16 = d->y_is_bf16; a.ld_y = d->ld_y;
  a.out_group_rows = d->out_group_rows; a.out_group_stride = d->out_group_stride; a.out_offset = d->out_offset;
  a.add0 = d->add0; a.bvec = d->bvec; a.bdiv = d->bdiv > 0 ? d->bdiv : 1; a.bscale = d->bscale;
  a.rows = d->rows; a.C = d->C;
  rows_affine_kernel<<<grid_for(a.rows * (a.C / 4), 256), 256, 0, stream>>>(a);
  return after_launch("rows_affine_kernel");
}

extern "C" int clv_scatter_add_rows(const float* src, const long long* index, float* dst, long long rows, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(src && index && dst && C > 0, "clv_scatter_add_rows: bad arguments");
  if (rows == 0) return 0;
  scatter_add_rows_kernel<<<grid_for(rows * C, 256), 256, 0, stream>>>(src, index, dst, rows, C);
  return after_launch("scatter_add_rows_kernel");
}
