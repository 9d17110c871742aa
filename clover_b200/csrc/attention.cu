// Fused small-sequence attention core (forward + backward) on packed qkv rows.
//
// One CTA per (batch element = window or caption, head).  The whole key/value set of a window
// (N <= 392 tokens) or caption+video sequence (S <= 432) lives in shared memory, so the
// (B_, nH, N, N) score tensor that the reference materialises and streams through HBM ~8 times
// per block (swin_transformer_3d.py:380-397) never leaves the SM: scores, relative-position bias
// (table lookup through a per-token code, closed form of :345-359), shift mask (region ids, closed
// form of compute_mask :548-562), BERT key-padding mask, online softmax and P.V are fused.
//
// Backward = 3 launches: prep (D_i = <dO_i, O_i>), dQ (+ relative-position-bias table gradient,
// the index_put_(accumulate) of the reference's autograd) and dK/dV; P is recomputed from the
// saved log-sum-exp.  Tensor-core math: mma.sync.m16n8k16 bf16 with fp32 accumulation (legacy
// HMMA path; head_dim 32 keeps this kernel exp/HBM bound).
#include "common.cuh"
#include "clover_b200.h"

namespace clv {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float NEG_BIG = -1.0e30f;

struct AttnArgs {
  int batch, seq, heads;
  const __nv_bfloat16* qkv;      // [batch*seq, 3*heads*HD]
  __nv_bfloat16* out;            // [batch*seq, heads*HD]
  float* lse;                    // [batch, heads, seq]
  const float* bias_table; int table_len; const int* rel_code; int code_off;
  const int* region; int nwin;
  const float* key_mask;
  uint32_t drop_thresh; float drop_inv_keep; unsigned long long drop_seed, drop_offset;   // drop_thresh == 0: off
  float* probs_mean;             // [batch, seq, seq] (clv_attention_probs_mean)
  // backward
  const __nv_bfloat16* dout; const float* dsum; __nv_bfloat16* dqkv; float q_scale; float* dbias;
};

CLV_DEVICE void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
CLV_DEVICE void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
CLV_DEVICE void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
CLV_DEVICE void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
CLV_DEVICE void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Copy rows [row0, row0+nrows) of one (batch, head) slice (HD columns starting at gcol of a matrix
// with `ld` columns) into smem [nrows][HD+8]; rows >= seq are zero-filled.
template <int HD>
CLV_DEVICE void load_rows(__nv_bfloat16* s, const __nv_bfloat16* g, long long ld, int row0, int nrows, int seq,
                          int tid, int nthreads) {
  constexpr int CH = HD / 8;       // 16-byte chunks per row
  constexpr int STRIDE = HD + 8;
  for (int i = tid; i < nrows * CH; i += nthreads) {
    const int r = i / CH, c = i % CH;
    __nv_bfloat16* dst = s + r * STRIDE + c * 8;
    if (row0 + r < seq) cp_async16(dst, g + (long long)(row0 + r) * ld + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// A-operand fragments (16 rows x HD) from a row-major smem tile.
template <int HD>
CLV_DEVICE void load_a_frags(uint32_t (&a)[HD / 16][4], const __nv_bfloat16* tile, int lane) {
  constexpr int STRIDE = HD + 8;
#pragma unroll
  for (int k = 0; k < HD / 16; ++k) ldsm_x4(a[k], tile + (lane & 15) * STRIDE + k * 16 + (lane >> 4) * 8);
}

// acc[16 x 64] = A(16 x HD) * R[n0 .. n0+63][HD]^T   (R row-major [n][k], "K-like" operand)
template <int HD>
CLV_DEVICE void mma_a_rt(float (&acc)[8][4], const uint32_t (&a)[HD / 16][4], const __nv_bfloat16* R, int n0,
                         int ntiles_valid, int lane) {
  constexpr int STRIDE = HD + 8;
#pragma unroll
  for (int nt = 0; nt < 8; nt += 2) {
    if (nt < ntiles_valid) {
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        uint32_t b[4];
        ldsm_x4(b, R + (n0 + nt * 8 + (lane & 7) + ((lane >> 4) << 3)) * STRIDE + k * 16 + ((lane >> 3) & 1) * 8);
        mma16816(acc[nt], a[k], b[0], b[1]);
        mma16816(acc[nt + 1], a[k], b[2], b[3]);
      }
    }
  }
}

// o[16 x HD] += P(16 x 64, fp32 accum layout -> bf16) * R[k0 .. k0+63][HD]   (R row-major [k][n], "V-like")
template <int HD>
CLV_DEVICE void mma_p_r(float (&o)[HD / 8][4], const float (&p)[8][4], const __nv_bfloat16* R, int k0,
                        int ntiles_valid, int lane) {
  constexpr int STRIDE = HD + 8;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (kk * 2 < ntiles_valid) {
      uint32_t a[4];
      a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
      a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
      a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
      a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < HD / 8; nt += 2) {
        uint32_t b[4];
        ldsm_x4_t(b, R + (k0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * STRIDE + nt * 8 + (lane >> 4) * 8);
        mma16816(o[nt], a, b[0], b[1]);
        mma16816(o[nt + 1], a, b[2], b[3]);
      }
    }
  }
}

// 2^x on the MUFU pipe without exp2f's denormal fix-up (the arguments here are <= 0 or differences to a row maximum / lse)
CLV_DEVICE float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct BiasCtx {
  const float* table;   // smem [table_len] for this head, or nullptr
  const int* code;      // smem [npad]
  const int* region;    // smem [npad] or nullptr
  const float* kmask;   // smem [npad] or nullptr
  int code_off, seq;
};

// additive score term for (query i, key j); NEG_BIG for keys/queries outside the sequence.
// MODE 1 = the BERT / fusion case (key mask only): no table / region tests in the element loop.
template <int MODE>
CLV_DEVICE float score_bias(const BiasCtx& c, int i, int j) {
  if (j >= c.seq || i >= c.seq) return NEG_BIG;
  if (MODE == 1) return c.kmask[j];
  float b = 0.f;
  if (c.table) b += c.table[c.code[i] - c.code[j] + c.code_off];
  if (c.region) b += (c.region[i] != c.region[j]) ? -100.0f : 0.0f;
  if (c.kmask) b += c.kmask[j];
  return b;
}

// keep / (1 - p) of attention probability (b, h, i, j): element ((b*heads + h)*seq + i)*seq + j of the stream
CLV_DEVICE float attn_keep(const AttnArgs& a, int bh, int i, int j) {
  const unsigned long long e = a.drop_offset + ((unsigned long long)bh * a.seq + (unsigned)i) * a.seq + (unsigned)j;
  return keep_scale(a.drop_seed, e, a.drop_thresh, a.drop_inv_keep);
}

// shared-memory carve-up helpers ------------------------------------------------------------
CLV_DEVICE int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <int HD, int MODE>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnArgs a) {
  constexpr int STRIDE = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int seq = a.seq, npad = round_up(seq, 16);
  const int bh = blockIdx.x, b = bh / a.heads, h = bh % a.heads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sV = sK + npad * STRIDE;
  __nv_bfloat16* sQ = sV + npad * STRIDE;                      // per-warp 16-row staging
  float* sTable = reinterpret_cast<float*>(sQ + nwarps * 16 * STRIDE);
  int* sCode = reinterpret_cast<int*>(sTable + (a.bias_table ? a.table_len : 0));
  int* sRegion = sCode + npad;
  float* sMask = reinterpret_cast<float*>(sRegion + npad);

  const long long ld = 3LL * a.heads * HD;
  const __nv_bfloat16* gq = a.qkv + (long long)b * seq * ld + h * HD;
  const __nv_bfloat16* gk = gq + a.heads * HD;
  const __nv_bfloat16* gv = gk + a.heads * HD;
  load_rows<HD>(sK, gk, ld, 0, npad, seq, tid, blockDim.x);
  load_rows<HD>(sV, gv, ld, 0, npad, seq, tid, blockDim.x);
  if (a.bias_table)
    for (int i = tid; i < a.table_len; i += blockDim.x) sTable[i] = a.bias_table[(long long)i * a.heads + h];
  for (int i = tid; i < npad; i += blockDim.x) {
    sCode[i] = (a.rel_code && i < seq) ? a.rel_code[i] : 0;
    sRegion[i] = (a.region && i < seq) ? a.region[(long long)(b % a.nwin) * seq + i] : 0;
    sMask[i] = (a.key_mask && i < seq) ? a.key_mask[(long long)b * seq + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  BiasCtx bc{a.bias_table ? sTable : nullptr, sCode, a.region ? sRegion : nullptr, a.key_mask ? sMask : nullptr,
             a.code_off, seq};
  __nv_bfloat16* myQ = sQ + warp * 16 * STRIDE;
  const int g = lane >> 2, q4 = lane & 3;
  const int nqt = npad / 16;
  for (int qt = warp; qt < nqt; qt += nwarps) {
    __syncwarp();
    load_rows<HD>(myQ, gq, ld, qt * 16, 16, seq, lane, 32);
    cp_async_wait_all();
    __syncwarp();
    uint32_t qa[HD / 16][4];
    load_a_frags<HD>(qa, myQ, lane);
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = NEG_BIG, m1 = NEG_BIG, l0 = 0.f, l1 = 0.f;
    const int i0 = qt * 16 + g, i1 = i0 + 8;
    for (int kc = 0; kc < npad; kc += 64) {
      const int ntv = min(8, (npad - kc) / 8);
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      mma_a_rt<HD>(s, qa, sK, kc, ntv, lane);
      float mx0 = NEG_BIG, mx1 = NEG_BIG;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = kc + nt * 8 + q4 * 2;
        if (nt < ntv) {
          s[nt][0] += score_bias<MODE>(bc, i0, j); s[nt][1] += score_bias<MODE>(bc, i0, j + 1);
          s[nt][2] += score_bias<MODE>(bc, i1, j); s[nt][3] += score_bias<MODE>(bc, i1, j + 1);
        } else {
          s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = NEG_BIG;
        }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float c0 = fast_ex2((m0 - mn0) * LOG2E), c1 = fast_ex2((m1 - mn1) * LOG2E);
      m0 = mn0; m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = fast_ex2((s[nt][0] - mn0) * LOG2E); s[nt][1] = fast_ex2((s[nt][1] - mn0) * LOG2E);
        s[nt][2] = fast_ex2((s[nt][2] - mn1) * LOG2E); s[nt][3] = fast_ex2((s[nt][3] - mn1) * LOG2E);
        rs0 += s[nt][0] + s[nt][1]; rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      if (a.drop_thresh) {     // dropout acts on the normalised probabilities; the normaliser l keeps every term
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int j = kc + nt * 8 + q4 * 2;
          s[nt][0] *= attn_keep(a, bh, i0, j); s[nt][1] *= attn_keep(a, bh, i0, j + 1);
          s[nt][2] *= attn_keep(a, bh, i1, j); s[nt][3] *= attn_keep(a, bh, i1, j + 1);
        }
      }
      mma_p_r<HD>(o, s, sV, kc, ntv, lane);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      *reinterpret_cast<uint32_t*>(myQ + g * STRIDE + nt * 8 + q4 * 2) = pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
      *reinterpret_cast<uint32_t*>(myQ + (g + 8) * STRIDE + nt * 8 + q4 * 2) = pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
    }
    if (q4 == 0) {
      float* lp = a.lse + ((long long)b * a.heads + h) * seq;
      if (i0 < seq) lp[i0] = m0 + logf(l0);
      if (i1 < seq) lp[i1] = m1 + logf(l1);
    }
    __syncwarp();
    constexpr int CH = HD / 8;
    for (int i = lane; i < 16 * CH; i += 32) {
      const int r = i / CH, c = i % CH;
      const int row = qt * 16 + r;
      if (row < seq)
        *reinterpret_cast<uint4*>(a.out + ((long long)b * seq + row) * (a.heads * HD) + h * HD + c * 8) =
            *reinterpret_cast<const uint4*>(myQ + r * STRIDE + c * 8);
    }
  }
}

// D[b,h,i] = sum_c dO[i,c] * O[i,c]
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* out, const __nv_bfloat16* dout, float* dsum, long long rows,
                                     int heads, int hd, int seq) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, head)
  if (idx >= rows * heads) return;
  const long long row = idx / heads; const int h = (int)(idx % heads);
  const uint4* po = reinterpret_cast<const uint4*>(out + row * heads * hd + h * hd);
  const uint4* pd = reinterpret_cast<const uint4*>(dout + row * heads * hd + h * hd);
  float s = 0.f;
  for (int c = 0; c < hd / 8; ++c) {
    const uint4 u = po[c], v = pd[c];
    float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
    float2 b0 = unpack_bf16(v.x), b1 = unpack_bf16(v.y), b2 = unpack_bf16(v.z), b3 = unpack_bf16(v.w);
    s += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
  }
  const long long bidx = row / seq; const int i = (int)(row % seq);
  dsum[(bidx * heads + h) * seq + i] = s;
}

// dQ (+ d bias table).  Resident: K, V.  Per warp: Q and dO tiles of 16 rows.
template <int HD, int MODE>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(AttnArgs a) {
  constexpr int STRIDE = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int seq = a.seq, npad = round_up(seq, 16);
  const int bh = blockIdx.x, b = bh / a.heads, h = bh % a.heads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sV = sK + npad * STRIDE;
  __nv_bfloat16* sQ = sV + npad * STRIDE;                      // per-warp: Q tile then dO tile
  float* sTable = reinterpret_cast<float*>(sQ + nwarps * 32 * STRIDE);
  float* sHist = sTable + (a.bias_table ? a.table_len : 0);
  int* sCode = reinterpret_cast<int*>(sHist + (a.dbias ? a.table_len : 0));
  int* sRegion = sCode + npad;
  float* sMask = reinterpret_cast<float*>(sRegion + npad);

  const long long ld = 3LL * a.heads * HD, ldo = (long long)a.heads * HD;
  const __nv_bfloat16* gq = a.qkv + (long long)b * seq * ld + h * HD;
  const __nv_bfloat16* gk = gq + a.heads * HD;
  const __nv_bfloat16* gv = gk + a.heads * HD;
  const __nv_bfloat16* gdo = a.dout + (long long)b * seq * ldo + h * HD;
  load_rows<HD>(sK, gk, ld, 0, npad, seq, tid, blockDim.x);
  load_rows<HD>(sV, gv, ld, 0, npad, seq, tid, blockDim.x);
  if (a.bias_table)
    for (int i = tid; i < a.table_len; i += blockDim.x) sTable[i] = a.bias_table[(long long)i * a.heads + h];
  if (a.dbias)
    for (int i = tid; i < a.table_len; i += blockDim.x) sHist[i] = 0.f;
  for (int i = tid; i < npad; i += blockDim.x) {
    sCode[i] = (a.rel_code && i < seq) ? a.rel_code[i] : 0;
    sRegion[i] = (a.region && i < seq) ? a.region[(long long)(b % a.nwin) * seq + i] : 0;
    sMask[i] = (a.key_mask && i < seq) ? a.key_mask[(long long)b * seq + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  BiasCtx bc{a.bias_table ? sTable : nullptr, sCode, a.region ? sRegion : nullptr, a.key_mask ? sMask : nullptr,
             a.code_off, seq};
  __nv_bfloat16* myQ = sQ + warp * 32 * STRIDE;
  __nv_bfloat16* myDO = myQ + 16 * STRIDE;
  const int g = lane >> 2, q4 = lane & 3;
  const float* lse = a.lse + ((long long)b * a.heads + h) * seq;
  const float* dsum = a.dsum + ((long long)b * a.heads + h) * seq;
  const int nqt = npad / 16;
  for (int qt = warp; qt < nqt; qt += nwarps) {
    __syncwarp();
    load_rows<HD>(myQ, gq, ld, qt * 16, 16, seq, lane, 32);
    load_rows<HD>(myDO, gdo, ldo, qt * 16, 16, seq, lane, 32);
    cp_async_wait_all();
    __syncwarp();
    uint32_t qa[HD / 16][4], da[HD / 16][4];
    load_a_frags<HD>(qa, myQ, lane);
    load_a_frags<HD>(da, myDO, lane);
    const int i0 = qt * 16 + g, i1 = i0 + 8;
    const float lse0 = i0 < seq ? lse[i0] : 0.f, lse1 = i1 < seq ? lse[i1] : 0.f;
    const float d0 = i0 < seq ? dsum[i0] : 0.f, d1 = i1 < seq ? dsum[i1] : 0.f;
    float dq[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kc = 0; kc < npad; kc += 64) {
      const int ntv = min(8, (npad - kc) / 8);
      float s[8][4], dp[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
      mma_a_rt<HD>(s, qa, sK, kc, ntv, lane);
      mma_a_rt<HD>(dp, da, sV, kc, ntv, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = kc + nt * 8 + q4 * 2;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (nt < ntv) {
          const float b0 = score_bias<MODE>(bc, i0, j), b1 = score_bias<MODE>(bc, i0, j + 1);
          const float b2 = score_bias<MODE>(bc, i1, j), b3 = score_bias<MODE>(bc, i1, j + 1);
          p0 = b0 <= NEG_BIG ? 0.f : fast_ex2((s[nt][0] + b0 - lse0) * LOG2E);
          p1 = b1 <= NEG_BIG ? 0.f : fast_ex2((s[nt][1] + b1 - lse0) * LOG2E);
          p2 = b2 <= NEG_BIG ? 0.f : fast_ex2((s[nt][2] + b2 - lse1) * LOG2E);
          p3 = b3 <= NEG_BIG ? 0.f : fast_ex2((s[nt][3] + b3 - lse1) * LOG2E);
        }
        if (a.drop_thresh) {   // dP = dP_dropped * keep / (1 - p)
          dp[nt][0] *= attn_keep(a, bh, i0, j); dp[nt][1] *= attn_keep(a, bh, i0, j + 1);
          dp[nt][2] *= attn_keep(a, bh, i1, j); dp[nt][3] *= attn_keep(a, bh, i1, j + 1);
        }
        s[nt][0] = p0 * (dp[nt][0] - d0); s[nt][1] = p1 * (dp[nt][1] - d0);
        s[nt][2] = p2 * (dp[nt][2] - d1); s[nt][3] = p3 * (dp[nt][3] - d1);
        if (a.dbias && nt < ntv) {
          if (i0 < seq && j < seq) atomicAdd(&sHist[sCode[i0] - sCode[j] + a.code_off], s[nt][0]);
          if (i0 < seq && j + 1 < seq) atomicAdd(&sHist[sCode[i0] - sCode[j + 1] + a.code_off], s[nt][1]);
          if (i1 < seq && j < seq) atomicAdd(&sHist[sCode[i1] - sCode[j] + a.code_off], s[nt][2]);
          if (i1 < seq && j + 1 < seq) atomicAdd(&sHist[sCode[i1] - sCode[j + 1] + a.code_off], s[nt][3]);
        }
      }
      mma_p_r<HD>(dq, s, sK, kc, ntv, lane);
    }
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      *reinterpret_cast<uint32_t*>(myQ + g * STRIDE + nt * 8 + q4 * 2) = pack_bf16(dq[nt][0] * a.q_scale, dq[nt][1] * a.q_scale);
      *reinterpret_cast<uint32_t*>(myQ + (g + 8) * STRIDE + nt * 8 + q4 * 2) = pack_bf16(dq[nt][2] * a.q_scale, dq[nt][3] * a.q_scale);
    }
    __syncwarp();
    constexpr int CH = HD / 8;
    for (int i = lane; i < 16 * CH; i += 32) {
      const int r = i / CH, c = i % CH;
      const int row = qt * 16 + r;
      if (row < seq)
        *reinterpret_cast<uint4*>(a.dqkv + ((long long)b * seq + row) * ld + h * HD + c * 8) =
            *reinterpret_cast<const uint4*>(myQ + r * STRIDE + c * 8);
    }
  }
  if (a.dbias) {
    __syncthreads();
    for (int i = tid; i < a.table_len; i += blockDim.x) {
      const float v = sHist[i];
      if (v != 0.f) atomicAdd(a.dbias + (long long)i * a.heads + h, v);
    }
  }
}

// dK, dV.  Resident: Q, dO (+ lse, D).  Per warp: K and V tiles of 16 keys.
template <int HD, int MODE>
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(AttnArgs a) {
  constexpr int STRIDE = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int seq = a.seq, npad = round_up(seq, 16);
  const int bh = blockIdx.x, b = bh / a.heads, h = bh % a.heads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sDO = sQ + npad * STRIDE;
  __nv_bfloat16* sKV = sDO + npad * STRIDE;                    // per-warp: K tile then V tile
  float* sTable = reinterpret_cast<float*>(sKV + nwarps * 32 * STRIDE);
  float* sLse = sTable + (a.bias_table ? a.table_len : 0);
  float* sD = sLse + npad;
  int* sCode = reinterpret_cast<int*>(sD + npad);
  int* sRegion = sCode + npad;
  float* sMask = reinterpret_cast<float*>(sRegion + npad);

  const long long ld = 3LL * a.heads * HD, ldo = (long long)a.heads * HD;
  const __nv_bfloat16* gq = a.qkv + (long long)b * seq * ld + h * HD;
  const __nv_bfloat16* gk = gq + a.heads * HD;
  const __nv_bfloat16* gv = gk + a.heads * HD;
  const __nv_bfloat16* gdo = a.dout + (long long)b * seq * ldo + h * HD;
  load_rows<HD>(sQ, gq, ld, 0, npad, seq, tid, blockDim.x);
  load_rows<HD>(sDO, gdo, ldo, 0, npad, seq, tid, blockDim.x);
  if (a.bias_table)
    for (int i = tid; i < a.table_len; i += blockDim.x) sTable[i] = a.bias_table[(long long)i * a.heads + h];
  const float* lse = a.lse + ((long long)b * a.heads + h) * seq;
  const float* dsum = a.dsum + ((long long)b * a.heads + h) * seq;
  for (int i = tid; i < npad; i += blockDim.x) {
    sCode[i] = (a.rel_code && i < seq) ? a.rel_code[i] : 0;
    sRegion[i] = (a.region && i < seq) ? a.region[(long long)(b % a.nwin) * seq + i] : 0;
    sMask[i] = (a.key_mask && i < seq) ? a.key_mask[(long long)b * seq + i] : 0.f;
    sLse[i] = i < seq ? lse[i] : 0.f;
    sD[i] = i < seq ? dsum[i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  BiasCtx bc{a.bias_table ? sTable : nullptr, sCode, a.region ? sRegion : nullptr, a.key_mask ? sMask : nullptr,
             a.code_off, seq};
  __nv_bfloat16* myK = sKV + warp * 32 * STRIDE;
  __nv_bfloat16* myV = myK + 16 * STRIDE;
  const int g = lane >> 2, q4 = lane & 3;
  const int nkt = npad / 16;
  for (int kt = warp; kt < nkt; kt += nwarps) {
    __syncwarp();
    load_rows<HD>(myK, gk, ld, kt * 16, 16, seq, lane, 32);
    load_rows<HD>(myV, gv, ld, kt * 16, 16, seq, lane, 32);
    cp_async_wait_all();
    __syncwarp();
    uint32_t ka[HD / 16][4], va[HD / 16][4];
    load_a_frags<HD>(ka, myK, lane);
    load_a_frags<HD>(va, myV, lane);
    const int j0 = kt * 16 + g, j1 = j0 + 8;
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qc = 0; qc < npad; qc += 64) {
      const int ntv = min(8, (npad - qc) / 8);
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
        dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
      }
      mma_a_rt<HD>(st, ka, sQ, qc, ntv, lane);       // S^T = K_tile Q^T
      mma_a_rt<HD>(dpt, va, sDO, qc, ntv, lane);     // dP^T = V_tile dO^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int i = qc + nt * 8 + q4 * 2;          // query index (column)
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        if (nt < ntv) {
          const float b0 = score_bias<MODE>(bc, i, j0), b1 = score_bias<MODE>(bc, i + 1, j0);
          const float b2 = score_bias<MODE>(bc, i, j1), b3 = score_bias<MODE>(bc, i + 1, j1);
          p0 = b0 <= NEG_BIG ? 0.f : fast_ex2((st[nt][0] + b0 - sLse[i]) * LOG2E);
          p1 = b1 <= NEG_BIG ? 0.f : fast_ex2((st[nt][1] + b1 - sLse[i + 1]) * LOG2E);
          p2 = b2 <= NEG_BIG ? 0.f : fast_ex2((st[nt][2] + b2 - sLse[i]) * LOG2E);
          p3 = b3 <= NEG_BIG ? 0.f : fast_ex2((st[nt][3] + b3 - sLse[i + 1]) * LOG2E);
          float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
          if (a.drop_thresh) {
            k0 = attn_keep(a, bh, i, j0); k1 = attn_keep(a, bh, i + 1, j0);
            k2 = attn_keep(a, bh, i, j1); k3 = attn_keep(a, bh, i + 1, j1);
          }
          dpt[nt][0] = p0 * (dpt[nt][0] * k0 - sD[i]); dpt[nt][1] = p1 * (dpt[nt][1] * k1 - sD[i + 1]);
          dpt[nt][2] = p2 * (dpt[nt][2] * k2 - sD[i]); dpt[nt][3] = p3 * (dpt[nt][3] * k3 - sD[i + 1]);
          p0 *= k0; p1 *= k1; p2 *= k2; p3 *= k3;      // dV uses the dropped probabilities
        } else {
          dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
        }
        st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
      }
      mma_p_r<HD>(dv, st, sDO, qc, ntv, lane);       // dV += P^T dO
      mma_p_r<HD>(dk, dpt, sQ, qc, ntv, lane);       // dK += dS^T Q
    }
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
      *reinterpret_cast<uint32_t*>(myK + g * STRIDE + nt * 8 + q4 * 2) = pack_bf16(dk[nt][0], dk[nt][1]);
      *reinterpret_cast<uint32_t*>(myK + (g + 8) * STRIDE + nt * 8 + q4 * 2) = pack_bf16(dk[nt][2], dk[nt][3]);
      *reinterpret_cast<uint32_t*>(myV + g * STRIDE + nt * 8 + q4 * 2) = pack_bf16(dv[nt][0], dv[nt][1]);
      *reinterpret_cast<uint32_t*>(myV + (g + 8) * STRIDE + nt * 8 + q4 * 2) = pack_bf16(dv[nt][2], dv[nt][3]);
    }
    __syncwarp();
    constexpr int CH = HD / 8;
    for (int i = lane; i < 32 * CH; i += 32) {
      const int r = i / CH, c = i % CH;          // r < 16: dK rows, r >= 16: dV rows
      const int row = kt * 16 + (r & 15);
      if (row < seq)
        *reinterpret_cast<uint4*>(a.dqkv + ((long long)b * seq + row) * ld + (r < 16 ? 1 : 2) * a.heads * HD + h * HD + c * 8) =
            *reinterpret_cast<const uint4*>(myK + r * STRIDE + c * 8);
    }
  }
}

// probs_mean[b, i, :] = mean_h softmax_j(q_i . k_j + additive terms).  One CTA per (query row, batch element); the
// thread block walks the heads, each thread owning keys tid, tid + blockDim, ...  (evaluation only, tiny).
template <int HD>
__global__ void __launch_bounds__(128) attn_probs_mean_kernel(AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  float* acc = reinterpret_cast<float*>(smem_attn);            // [seq]
  float* sq = acc + a.seq;                                     // [HD]
  __shared__ float red[4];
  const int i = blockIdx.x, b = blockIdx.y, seq = a.seq, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ld = 3LL * a.heads * HD;
  const __nv_bfloat16* base = a.qkv + (long long)b * seq * ld;
  for (int j = tid; j < seq; j += blockDim.x) acc[j] = 0.f;
  const int ci = a.rel_code ? a.rel_code[i] : 0;
  const int ri = a.region ? a.region[(long long)(b % a.nwin) * seq + i] : 0;
  for (int h = 0; h < a.heads; ++h) {
    __syncthreads();
    if (tid < HD) sq[tid] = __bfloat162float(base[(long long)i * ld + h * HD + tid]);
    __syncthreads();
    float sc[4];                                               // seq <= 512 = 4 * 128
    float mx = NEG_BIG;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = tid + u * 128;
      sc[u] = NEG_BIG;
      if (j < seq) {
        const uint4* kr = reinterpret_cast<const uint4*>(base + (long long)j * ld + (a.heads + h) * HD);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          const uint4 kv = kr[c];
          const float2 k0 = unpack_bf16(kv.x), k1 = unpack_bf16(kv.y), k2 = unpack_bf16(kv.z), k3 = unpack_bf16(kv.w);
          const float* q = sq + c * 8;
          s += q[0] * k0.x + q[1] * k0.y + q[2] * k1.x + q[3] * k1.y + q[4] * k2.x + q[5] * k2.y + q[6] * k3.x + q[7] * k3.y;
        }
        if (a.bias_table) s += a.bias_table[(long long)(ci - a.rel_code[j] + a.code_off) * a.heads + h];
        if (a.region) s += (a.region[(long long)(b % a.nwin) * seq + j] != ri) ? -100.0f : 0.0f;
        if (a.key_mask) s += a.key_mask[(long long)b * seq + j];
        sc[u] = s;
        mx = fmaxf(mx, s);
      }
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sc[u] = (tid + u * 128 < seq) ? fast_ex2((sc[u] - mx) * LOG2E) : 0.f;
      sum += sc[u];
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float inv = 1.0f / ((red[0] + red[1] + red[2] + red[3]) * a.heads);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (tid + u * 128 < seq) acc[tid + u * 128] += sc[u] * inv;
  }
  __syncthreads();
  float* o = a.probs_mean + ((long long)b * seq + i) * seq;
  for (int j = tid; j < seq; j += blockDim.x) o[j] = acc[j];
}

int launch_attn_bwd_prep(const void* out, const void* dout, float* dsum, long long rows, int heads, int hd, int seq,
                         cudaStream_t stream) {
  const long long n = rows * heads;
  attn_bwd_prep_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(out),
                                                                 reinterpret_cast<const __nv_bfloat16*>(dout), dsum, rows, heads, hd, seq);
  return after_launch("attn_bwd_prep_kernel");
}

static int attn_common_checks(const clv_attn_desc_t* d) {
  CLV_REQUIRE(d != nullptr, "attention: null descriptor");
  CLV_REQUIRE(d->head_dim == 32 || d->head_dim == 64, "attention: head_dim must be 32 or 64 (got %d)", d->head_dim);
  CLV_REQUIRE(d->seq > 0 && d->seq <= 512 && d->batch > 0 && d->heads > 0, "attention: bad shape batch=%d seq=%d heads=%d",
              d->batch, d->seq, d->heads);
  CLV_REQUIRE(!d->bias_table || (d->rel_code && d->table_len > 0), "attention: bias_table needs rel_code/table_len");
  CLV_REQUIRE(!d->region || d->nwin > 0, "attention: region needs nwin");
  CLV_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f, "attention: drop_p must be in [0, 1) (got %f)", (double)d->drop_p);
  return 0;
}

static void fill_args(AttnArgs& a, const clv_attn_desc_t* d) {
  a = AttnArgs{};
  a.batch = d->batch; a.seq = d->seq; a.heads = d->heads;
  a.bias_table = d->bias_table; a.table_len = d->table_len; a.rel_code = d->rel_code; a.code_off = d->code_off;
  a.region = d->region; a.nwin = d->nwin > 0 ? d->nwin : 1; a.key_mask = d->key_mask;
  if (d->drop_p > 0.f) {
    a.drop_thresh = drop_threshold(d->drop_p); a.drop_inv_keep = 1.0f / (1.0f - d->drop_p);
    a.drop_seed = d->drop_seed; a.drop_offset = d->drop_offset;
  }
}

template <typename K>
static int launch_attn(K kern, const AttnArgs& a, int nwarps, size_t smem, cudaStream_t stream, const char* what) {
  CLV_REQUIRE(smem <= 227 * 1024, "%s: needs %zu bytes of shared memory (seq too long)", what, smem);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)smem)) return rc;
  kern<<<a.batch * a.heads, nwarps * 32, smem, stream>>>(a);
  return after_launch(what);
}

static int pick_warps(size_t resident_bytes) { return resident_bytes > 48 * 1024 ? 8 : 4; }

}  // namespace clv

using namespace clv;

extern "C" int clv_attention_fwd(const clv_attn_desc_t* d, const void* qkv, void* out, float* lse, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = attn_common_checks(d)) return rc;
  CLV_REQUIRE(qkv && out && lse, "attention_fwd: null pointer");
  AttnArgs a; fill_args(a, d);
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse;
  const int hd = d->head_dim, npad = (d->seq + 15) / 16 * 16, stride = hd + 8;
  const size_t resident = (size_t)2 * npad * stride * 2;
  const int nw = pick_warps(resident);
  const size_t smem = resident + (size_t)nw * 16 * stride * 2 + (d->bias_table ? d->table_len * 4 : 0) + (size_t)npad * 12;
  const bool mask_only = d->key_mask && !d->bias_table && !d->region;
  if (hd == 32) return launch_attn(attn_fwd_kernel<32, 0>, a, nw, smem, stream, "attn_fwd_kernel<32>");
  if (mask_only) return launch_attn(attn_fwd_kernel<64, 1>, a, nw, smem, stream, "attn_fwd_kernel<64,mask>");
  return launch_attn(attn_fwd_kernel<64, 0>, a, nw, smem, stream, "attn_fwd_kernel<64>");
}

extern "C" int clv_attention_bwd(const clv_attn_desc_t* d, const void* qkv, const void* out, const void* dout,
                                 const float* lse, void* dqkv, float q_scale, float* dbias_table, float* dsum_ws,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = attn_common_checks(d)) return rc;
  CLV_REQUIRE(qkv && out && dout && lse && dqkv && dsum_ws, "attention_bwd: null pointer");
  CLV_REQUIRE(!dbias_table || d->bias_table, "attention_bwd: dbias_table without bias_table");
  AttnArgs a; fill_args(a, d);
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); a.out = nullptr; a.lse = const_cast<float*>(lse);
  a.dout = reinterpret_cast<const __nv_bfloat16*>(dout); a.dsum = dsum_ws;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); a.q_scale = q_scale; a.dbias = dbias_table;
  const int hd = d->head_dim, npad = (d->seq + 15) / 16 * 16, stride = hd + 8;
  const long long rows = (long long)d->batch * d->seq;
  {
    const long long n = rows * d->heads;
    attn_bwd_prep_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(out), a.dout,
                                                                   dsum_ws, rows, d->heads, hd, d->seq);
    if (int rc = after_launch("attn_bwd_prep_kernel")) return rc;
  }
  const size_t resident = (size_t)2 * npad * stride * 2;
  const int nw = pick_warps(resident);
  const size_t tab = d->bias_table ? (size_t)d->table_len * 4 : 0;
  const size_t smem_dq = resident + (size_t)nw * 32 * stride * 2 + tab + (dbias_table ? tab : 0) + (size_t)npad * 12;
  const size_t smem_dkv = resident + (size_t)nw * 32 * stride * 2 + tab + (size_t)npad * 20;
  int rc;
  const bool mask_only = d->key_mask && !d->bias_table && !d->region;
  if (hd == 32) {
    rc = launch_attn(attn_bwd_dq_kernel<32, 0>, a, nw, smem_dq, stream, "attn_bwd_dq_kernel<32>");
    if (!rc) rc = launch_attn(attn_bwd_dkv_kernel<32, 0>, a, nw, smem_dkv, stream, "attn_bwd_dkv_kernel<32>");
  } else if (mask_only) {
    rc = launch_attn(attn_bwd_dq_kernel<64, 1>, a, nw, smem_dq, stream, "attn_bwd_dq_kernel<64,mask>");
    if (!rc) rc = launch_attn(attn_bwd_dkv_kernel<64, 1>, a, nw, smem_dkv, stream, "attn_bwd_dkv_kernel<64,mask>");
  } else {
    rc = launch_attn(attn_bwd_dq_kernel<64, 0>, a, nw, smem_dq, stream, "attn_bwd_dq_kernel<64>");
    if (!rc) rc = launch_attn(attn_bwd_dkv_kernel<64, 0>, a, nw, smem_dkv, stream, "attn_bwd_dkv_kernel<64>");
  }
  return rc;
}

extern "C" int clv_attention_probs_mean(const clv_attn_desc_t* d, const void* qkv, float* probs_mean, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = attn_common_checks(d)) return rc;
  CLV_REQUIRE(qkv && probs_mean, "attention_probs_mean: null pointer");
  CLV_REQUIRE(d->drop_p == 0.f, "attention_probs_mean: evaluation only (drop_p must be 0)");
  AttnArgs a; fill_args(a, d);
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv); a.probs_mean = probs_mean;
  const size_t smem = (size_t)(d->seq + d->head_dim) * sizeof(float);
  dim3 grid(d->seq, d->batch);
  if (d->head_dim == 32) attn_probs_mean_kernel<32><<<grid, 128, smem, stream>>>(a);
  else attn_probs_mean_kernel<64><<<grid, 128, smem, stream>>>(a);
  return after_launch("attn_probs_mean_kernel");
}
