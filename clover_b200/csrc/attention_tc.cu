// Window-attention core on tcgen05 / TMEM / TMA (sm_100a), head_dim 32 -- forward.
//
// Work unit = (window b, head h, query tile t of <= 128 rows).  Per unit:
//   TMA   : Q tile [128 x 32], K [NK x 32], V [NK x 32] of head h straight out of the packed qkv rows
//           (64-byte rows, SWIZZLE_64B) -- the reference's reshape/permute of
//           swin_transformer_3d.py:376-377 is just the TMA coordinate;
//   UMMA  : S = Q K^T  (M=128, N=NK<=2x208, K=32) into tensor memory;
//   warps : one thread per query row (= TMEM lane): + relative-position bias (table[code_i - code_j + off],
//           :382-385) + shift mask (region ids, :388-390), row max, exp2, row sum -- no shuffles, the
//           row lives in one thread; P is written back in place as packed bf16;
//   UMMA  : O = P V  with P as the TMEM A-operand and V as an MN-major smem operand (K = NK);
//   warps : O / rowsum -> bf16 -> global, log-sum-exp saved for the backward.
// Roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax / epilogue.
// Two CTAs per SM (256 TMEM columns each) hide each other's MMA / TMA latency; a CTA double-buffers
// the next unit's Q/K/V in shared memory.
#include <algorithm>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

constexpr int TC_HD = 32;
constexpr int TC_ROWB = TC_HD * 2;           // bytes per Q/K/V row (64)
constexpr int TC_THREADS = 192;
constexpr float TC_LOG2E = 1.4426950408889634f;

struct AttnTcArgs {
  int batch, seq, heads;
  int nk;                       // keys padded to a multiple of 32
  int n_qt, rows_per_tile;
  int kv_boxes, kv_box_rows;    // TMA boxes per K / V load
  int kb_bytes;                 // bytes reserved per K (or V) buffer, multiple of 1024
  int tmem_cols;
  long long units;
  __nv_bfloat16* out;
  float* lse;
  const float* bias_table; int table_len; const int* rel_code; int code_off;
  const int* region; int nwin;
};

CLV_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (an integer round-trip would demote every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = 8192 + 2 * a.kb_bytes;
  float* sTable = reinterpret_cast<float*>(smem + 2 * stage_bytes);
  int* sCode = reinterpret_cast<int*>(sTable + ((a.table_len + 3) & ~3));
  int* sReg = sCode + a.nk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sReg + a.nk);   // 16-byte aligned by construction
  uint64_t* full_bar = bars;          // [2]
  uint64_t* empty_bar = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * TC_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1); mbar_init(s_free, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
  if (warp >= 2) {
    for (int j = threadIdx.x - 64; j < a.nk; j += 128) sCode[j] = (a.rel_code && j < a.seq) ? a.rel_code[j] : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        const int t = (int)(u % a.n_qt);
        const long long bh = u / a.n_qt;
        const int b = (int)(bh % a.batch), h = (int)(bh / a.batch);
        const int stage = it & 1;
        const uint32_t phase = (it >> 1) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sQ = smem + stage * stage_bytes;
        uint8_t* sK = sQ + 8192;
        uint8_t* sV = sK + a.kb_bytes;
        mbar_expect_tx(&full_bar[stage], 8192 + 2 * a.nk * TC_ROWB);
        const int row0 = b * a.seq;
        tma_load_2d(sQ, &tm_q, &full_bar[stage], h * TC_HD, row0 + t * a.rows_per_tile);
        for (int i = 0; i < a.kv_boxes; ++i) {
          tma_load_2d(sK + i * a.kv_box_rows * TC_ROWB, &tm_kv, &full_bar[stage], C + h * TC_HD, row0 + i * a.kv_box_rows);
          tma_load_2d(sV + i * a.kv_box_rows * TC_ROWB, &tm_kv, &full_bar[stage], 2 * C + h * TC_HD, row0 + i * a.kv_box_rows);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_pv = make_idesc_bf16(128, TC_HD, 0, 1);
      // S-MMA column chunks (N <= 256, multiple of 16)
      const int n0 = a.nk <= 256 ? a.nk : ((a.nk / 2 + 15) & ~15);
      const int n1 = a.nk - n0;
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        const int stage = it & 1;
        const uint32_t phase = (it >> 1) & 1;
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(s_free, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t k_addr = q_addr + 8192;
        const uint32_t v_addr = k_addr + a.kb_bytes;
#pragma unroll
        for (int k = 0; k < TC_HD / 16; ++k) {
          const uint64_t da = make_smem_desc(q_addr + k * 32, 16, 512, 4);
          umma_bf16_ss(tmem_base, da, make_smem_desc(k_addr + k * 32, 16, 512, 4), make_idesc_bf16(128, n0, 0, 0), k > 0);
        }
        if (n1 > 0) {
#pragma unroll
          for (int k = 0; k < TC_HD / 16; ++k) {
            const uint64_t da = make_smem_desc(q_addr + k * 32, 16, 512, 4);
            umma_bf16_ss(tmem_base + n0, da, make_smem_desc(k_addr + n0 * TC_ROWB + k * 32, 16, 512, 4),
                         make_idesc_bf16(128, n1, 0, 0), k > 0);
          }
        }
        umma_commit(s_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        const uint32_t tmem_o = tmem_base + (a.nk - TC_HD);
        for (int kk = 0; kk < a.nk / 16; ++kk)
          umma_bf16_ts(tmem_o, tmem_base + kk * 8, make_smem_desc(v_addr + kk * 1024, 16, 512, 4), idesc_pv, kk > 0);
        umma_commit(o_full);
        umma_commit(&empty_bar[stage]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool use_region = a.region != nullptr;
    int cur_h = -1;
    uint32_t it = 0;
    for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
      const int t = (int)(u % a.n_qt);
      const long long bh = u / a.n_qt;
      const int b = (int)(bh % a.batch), h = (int)(bh / a.batch);
      const int i = t * a.rows_per_tile + r;
      const bool valid = r < a.rows_per_tile && i < a.seq;
      const bool warp_active = (quarter * 32) < a.rows_per_tile && (t * a.rows_per_tile + quarter * 32) < a.seq;
      // ---- per-unit tables
      named_bar_sync(1, 128);
      if (h != cur_h) {
        if (a.bias_table)
          for (int x = tid; x < a.table_len; x += 128) sTable[x] = a.bias_table[(long long)x * a.heads + h];
        cur_h = h;
      }
      if (use_region) {
        const int* rg = a.region + (long long)(b % a.nwin) * a.seq;
        for (int j = tid; j < a.nk; j += 128) sReg[j] = j < a.seq ? rg[j] : 0;
      }
      named_bar_sync(1, 128);
      const int ci = (valid ? sCode[i] : 0) + a.code_off;
      const int ri = (valid && use_region) ? sReg[i] : 0;

      mbar_wait(s_full, it & 1);
      tc_fence_after();
      float m = -1.0e30f, l = 0.f;
      if (warp_active) {
        // ---- pass 1: s += bias (+ mask); row max; biased scores written back to TMEM
#pragma unroll 1
        for (int c0 = 0; c0 < a.nk; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          if (a.bias_table) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int4 c4 = *reinterpret_cast<const int4*>(sCode + c0 + j);      // broadcast 128-bit load
              v[j] = __float_as_uint(__uint_as_float(v[j]) + sTable[ci - c4.x]);
              v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + sTable[ci - c4.y]);
              v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + sTable[ci - c4.z]);
              v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + sTable[ci - c4.w]);
            }
          }
          if (use_region) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int4 r4 = *reinterpret_cast<const int4*>(sReg + c0 + j);
              if (r4.x != ri) v[j] = __float_as_uint(__uint_as_float(v[j]) - 100.0f);
              if (r4.y != ri) v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) - 100.0f);
              if (r4.z != ri) v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) - 100.0f);
              if (r4.w != ri) v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) - 100.0f);
            }
          }
          if (c0 + 32 > a.seq) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j >= a.seq) v[j] = __float_as_uint(-1.0e30f);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { lo[j] = v[j]; hi[j] = v[16 + j]; }
          tmem_st_32x16(taddr + c0, lo);
          tmem_st_32x16(taddr + c0 + 16, hi);
        }
        tmem_st_wait();
        // ---- pass 2: p = exp(s - m), row sum, packed bf16 P written in place (columns [0, nk/2))
        const float mL = m * TC_LOG2E;
#pragma unroll 1
        for (int c0 = 0; c0 < a.nk; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = ex2(fmaf(__uint_as_float(v[2 * j]), TC_LOG2E, -mL));
            const float p1 = ex2(fmaf(__uint_as_float(v[2 * j + 1]), TC_LOG2E, -mL));
            l += p0 + p1;
            pk[j] = pack_bf16(p0, p1);
          }
          tmem_st_32x16(taddr + (c0 >> 1), pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      // ---- epilogue: O / l -> bf16 -> global; lse
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      if (warp_active) {
        uint32_t o[32];
        tmem_ld_32x32(taddr + (a.nk - TC_HD), o);
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.0f / l;
          uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)b * a.seq + i) * C + h * TC_HD);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_bf16(__uint_as_float(o[q * 8]) * inv, __uint_as_float(o[q * 8 + 1]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv));
          a.lse[((long long)b * a.heads + h) * a.seq + i] = m + logf(l);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, a.tmem_cols);
}

}  // namespace clv

using namespace clv;

// head_dim 32 forward on tensor memory; same contract as clv_attention_fwd (key_mask unsupported).
extern "C" int clv_attention_fwd_tc(const clv_attn_desc_t* d, const void* qkv, void* out, float* lse, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(d && qkv && out && lse, "attention_fwd_tc: null pointer");
  CLV_REQUIRE(d->head_dim == 32 && !d->key_mask, "attention_fwd_tc: head_dim 32 without key mask only");
  CLV_REQUIRE(d->drop_p == 0.f, "attention_fwd_tc: attention dropout is not supported (attn_drop is 0 in Video Swin)");
  CLV_REQUIRE(d->seq >= 33 && d->seq <= 416, "attention_fwd_tc: seq must be in [33, 416] (got %d)", d->seq);
  CLV_REQUIRE(!d->bias_table || (d->rel_code && d->table_len > 0), "attention_fwd_tc: bias_table needs rel_code");
  AttnTcArgs a{};
  a.batch = d->batch; a.seq = d->seq; a.heads = d->heads;
  a.nk = (d->seq + 31) / 32 * 32;
  a.n_qt = (d->seq + 127) / 128;
  a.rows_per_tile = (d->seq + a.n_qt - 1) / a.n_qt;
  a.kv_boxes = a.nk <= 256 ? 1 : 2;
  a.kv_box_rows = a.nk / a.kv_boxes;
  a.kb_bytes = (a.nk * TC_ROWB + 1023) / 1024 * 1024;
  a.tmem_cols = a.nk <= 256 ? 256 : 512;
  a.units = (long long)d->batch * d->heads * a.n_qt;
  a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.bias_table = d->bias_table; a.table_len = d->bias_table ? d->table_len : 0; a.rel_code = d->rel_code; a.code_off = d->code_off;
  a.region = d->region; a.nwin = d->nwin > 0 ? d->nwin : 1;
  const long long rows = (long long)d->batch * d->seq;
  const long long ld = 3LL * d->heads * TC_HD;
  CUtensorMap tq, tkv;
  if (int rc = make_tmap_bf16_2d(&tq, qkv, ld, rows, ld, TC_HD, 128, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tkv, qkv, ld, rows, ld, TC_HD, a.kv_box_rows, 64)) return rc;
  const size_t smem = 1024 + 2 * (size_t)(8192 + 2 * a.kb_bytes) + (size_t)((a.table_len + 3) & ~3) * 4 + (size_t)a.nk * 8 + 8 + 128;
  CLV_REQUIRE(smem <= 227 * 1024, "attention_fwd_tc: %zu bytes of shared memory needed", smem);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_tc_kernel), (int)smem)) return rc;
  const int per_sm = (a.tmem_cols == 256 && smem <= 110 * 1024) ? 2 : 1;
  const int grid = (int)std::min<long long>(a.units, (long long)num_sms() * per_sm);
  attn_fwd_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(tq, tkv, a);
  return after_launch("attn_fwd_tc_kernel");
}

// =================================================================================================
// Backward (head_dim 32, seq <= 224).  Unit = (window b, head h); inner loop over key tiles t.
//   UMMA 1 : S^T = K_t Q^T  and  dP^T = V_t dO^T   (M = 128 keys, N = NQ queries, K = 32)   -> TMEM
//   warps  : one thread per key row j: p = exp(s + bias - lse_i), ds = p (dp - D_i); P^T and dS^T are
//            written back in place as packed bf16 (TMEM A-operands); dS^T also goes to shared memory
//            (MN-major, SWIZZLE_128B) for the dQ product and to global (bf16) for the bias-table gradient
//   UMMA 2 : dV_t = P^T dO, dK_t = dS^T Q  (A from TMEM), dQ += dS K_t (A from smem, accumulated in TMEM
//            across the key tiles of the unit)
//   warps  : dK_t / dV_t (and at the end of the unit dQ * q_scale) -> bf16 -> packed dqkv rows.
// The reference's autograd reduces d(bias) over all windows with index_put_(accumulate) on an fp16 dS
// tensor; here dS^T is stored once in bf16 and reduced by dbias_reduce_kernel into the (2535, nH) table.
// =================================================================================================
struct AttnTcBwdArgs {
  int batch, seq, heads;
  int nq;                       // queries padded to a multiple of 32 (<= 224)
  int n_kt, rows_per_tile;      // key tiling
  int n_mq;                     // 128-row query tiles of the dQ accumulator
  int qb_bytes;                 // bytes per Q (or dO) buffer, multiple of 1024
  int tmem_cols;
  int col_split;                // queries [0, col_split) are handled by softmax warps 2-5, the rest by warps 6-9
  long long units;
  const float* lse; const float* dsum;
  __nv_bfloat16* dqkv; float q_scale;
  __nv_bfloat16* ds_out;        // [batch, heads, seq, nq] bf16 or nullptr
  const float* bias_table; int table_len; const int* rel_code; int code_off;
  const int* region; int nwin;
};

constexpr int TC_BWD_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two column halves x four lane quarters)

__global__ void __launch_bounds__(TC_BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_full, const __grid_constant__ CUtensorMap tm_qkv_tile,
                   const __grid_constant__ CUtensorMap tm_do_full, AttnTcBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (an integer round-trip would demote every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: [Q0 dO0 Q1 dO1] [K0 V0 K1 V1] [dS^T tile] tables barriers
  uint8_t* sQdO = smem;
  uint8_t* sKV = sQdO + 4 * a.qb_bytes;
  uint8_t* sDS = sKV + 4 * 8192;
  const int ds_bytes = a.n_mq * 2 * 16384;
  float* sTable = reinterpret_cast<float*>(sDS + ds_bytes);
  float* sLse = sTable + ((a.table_len + 3) & ~3);
  float* sD = sLse + a.nq;
  int* sCode = reinterpret_cast<int*>(sD + a.nq);
  int* sReg = sCode + a.nq;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sReg + a.nq);   // 16-byte aligned by construction
  uint64_t* qdo_full = bars;        // [2]
  uint64_t* qdo_empty = bars + 2;   // [2]
  uint64_t* kv_full = bars + 4;     // [2]
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* st_full = bars + 8;
  uint64_t* p_ready = bars + 9;
  uint64_t* mma2_done = bars + 10;
  uint64_t* acc_free = bars + 11;
  uint64_t* dq_free = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * TC_HD;
  const int NQ = a.nq;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv_full); tma_prefetch_desc(&tm_qkv_tile); tma_prefetch_desc(&tm_do_full);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
    }
    mbar_init(st_full, 1); mbar_init(p_ready, 8); mbar_init(mma2_done, 1); mbar_init(acc_free, 8); mbar_init(dq_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
  if (warp >= 2) {
    const int tid = threadIdx.x - 64;
    for (int j = tid; j < NQ; j += 256) sCode[j] = (a.rel_code && j < a.seq) ? a.rel_code[j] : 0;
    for (int x = tid; x < ds_bytes / 16; x += 256) reinterpret_cast<uint4*>(sDS)[x] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t col_dv = NQ - TC_HD, col_dp = NQ, col_dk = 2 * NQ - TC_HD, col_dq = 2 * NQ;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0, tt = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        const int b = (int)(u % a.batch), h = (int)(u / a.batch);
        const int us = it & 1;
        mbar_wait(&qdo_empty[us], ((it >> 1) & 1) ^ 1);
        uint8_t* sQ = sQdO + us * 2 * a.qb_bytes;
        uint8_t* sDO = sQ + a.qb_bytes;
        mbar_expect_tx(&qdo_full[us], 2 * NQ * TC_ROWB);
        const int row0 = b * a.seq;
        tma_load_2d(sQ, &tm_qkv_full, &qdo_full[us], h * TC_HD, row0);
        tma_load_2d(sDO, &tm_do_full, &qdo_full[us], h * TC_HD, row0);
        for (int t = 0; t < a.n_kt; ++t, ++tt) {
          const int ts = tt & 1;
          mbar_wait(&kv_empty[ts], ((tt >> 1) & 1) ^ 1);
          uint8_t* sK = sKV + ts * 2 * 8192;
          mbar_expect_tx(&kv_full[ts], 2 * 8192);
          tma_load_2d(sK, &tm_qkv_tile, &kv_full[ts], C + h * TC_HD, row0 + t * a.rows_per_tile);
          tma_load_2d(sK + 8192, &tm_qkv_tile, &kv_full[ts], 2 * C + h * TC_HD, row0 + t * a.rows_per_tile);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_st = make_idesc_bf16(128, NQ, 0, 0);
      const uint32_t idesc_ts = make_idesc_bf16(128, TC_HD, 0, 1);
      const uint32_t idesc_dq = make_idesc_bf16(128, TC_HD, 1, 1);
      uint32_t it = 0, tt = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        const int us = it & 1;
        mbar_wait(&qdo_full[us], (it >> 1) & 1);
        mbar_wait(dq_free, (it & 1) ^ 1);
        const uint32_t q_addr = smem_u32(sQdO + us * 2 * a.qb_bytes);
        const uint32_t do_addr = q_addr + a.qb_bytes;
        for (int t = 0; t < a.n_kt; ++t, ++tt) {
          const int ts = tt & 1;
          mbar_wait(&kv_full[ts], (tt >> 1) & 1);
          mbar_wait(acc_free, (tt & 1) ^ 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sKV + ts * 2 * 8192);
          const uint32_t v_addr = k_addr + 8192;
#pragma unroll
          for (int k = 0; k < TC_HD / 16; ++k)
            umma_bf16_ss(tmem_base, make_smem_desc(k_addr + k * 32, 16, 512, 4), make_smem_desc(q_addr + k * 32, 16, 512, 4),
                         idesc_st, k > 0);
#pragma unroll
          for (int k = 0; k < TC_HD / 16; ++k)
            umma_bf16_ss(tmem_base + col_dp, make_smem_desc(v_addr + k * 32, 16, 512, 4),
                         make_smem_desc(do_addr + k * 32, 16, 512, 4), idesc_st, k > 0);
          umma_commit(st_full);
          mbar_wait(p_ready, tt & 1);
          tc_fence_after();
          for (int kk = 0; kk < NQ / 16; ++kk) {   // dV_t = P^T dO ; dK_t = dS^T Q   (K = queries)
            // packed bf16 operands live in two column segments (one per softmax warp group)
            const uint32_t pc = kk * 16 < a.col_split ? kk * 8 : a.col_split + ((kk * 16 - a.col_split) >> 1);
            umma_bf16_ts(tmem_base + col_dv, tmem_base + pc, make_smem_desc(do_addr + kk * 1024, 16, 512, 4), idesc_ts, kk > 0);
            umma_bf16_ts(tmem_base + col_dk, tmem_base + col_dp + pc, make_smem_desc(q_addr + kk * 1024, 16, 512, 4), idesc_ts,
                         kk > 0);
          }
          const uint32_t ds_addr = smem_u32(sDS);
          for (int mq = 0; mq < a.n_mq; ++mq)      // dQ[mq] += dS K_t   (K = 128 keys of this tile)
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_ss(tmem_base + col_dq + mq * TC_HD, make_smem_desc(ds_addr + mq * 2 * 16384 + ks * 2048, 16384, 1024, 2),
                           make_smem_desc(k_addr + ks * 1024, 16, 512, 4), idesc_dq, (t > 0 || ks > 0) ? 1u : 0u);
          umma_commit(mma2_done);
          umma_commit(&kv_empty[ts]);
          if (t == a.n_kt - 1) umma_commit(&qdo_empty[us]);
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool use_region = a.region != nullptr;
    const int c_begin = chalf ? a.col_split : 0, c_end = chalf ? NQ : a.col_split;
    int cur_h = -1;
    uint32_t it = 0, tt = 0;
    for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
      const int b = (int)(u % a.batch), h = (int)(u / a.batch);
      named_bar_sync(1, 256);
      if (h != cur_h) {
        if (a.bias_table)
          for (int x = tid; x < a.table_len; x += 256) sTable[x] = a.bias_table[(long long)x * a.heads + h];
        cur_h = h;
      }
      {
        const float* lp = a.lse + ((long long)b * a.heads + h) * a.seq;
        const float* dp = a.dsum + ((long long)b * a.heads + h) * a.seq;
        const int* rg = use_region ? a.region + (long long)(b % a.nwin) * a.seq : nullptr;
        for (int j = tid; j < NQ; j += 256) {
          sLse[j] = j < a.seq ? lp[j] * TC_LOG2E : 1.0e30f;      // queries outside the window get p = 0
          sD[j] = j < a.seq ? dp[j] : 0.f;
          sReg[j] = (rg && j < a.seq) ? rg[j] : 0;
        }
      }
      named_bar_sync(1, 256);
      for (int t = 0; t < a.n_kt; ++t, ++tt) {
        const int j = t * a.rows_per_tile + r;
        const bool valid = r < a.rows_per_tile && j < a.seq;
        const bool warp_active = (quarter * 32) < a.rows_per_tile && (t * a.rows_per_tile + quarter * 32) < a.seq;
        const int cj = a.code_off - (valid ? sCode[j] : 0);
        const int rj = valid ? sReg[j] : 0;
        uint8_t* ds_row = sDS + (r >> 3) * 1024 + (r & 7) * 128;
        __nv_bfloat16* ds_g = a.ds_out ? a.ds_out + (((long long)b * a.heads + h) * a.seq + (valid ? j : 0)) * NQ : nullptr;
        mbar_wait(st_full, tt & 1);
        tc_fence_after();
        if (warp_active) {
#pragma unroll 1
          for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            const uint32_t pc = c_begin + ((c0 - c_begin) >> 1);      // packed columns of this warp group's segment
            uint32_t v[32], w[32];
            tmem_ld_32x32(taddr + c0, v);
            tmem_ld_32x32(taddr + col_dp + c0, w);
            tmem_ld_wait();
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              const int i = c0 + x;
              const int4 c4 = *reinterpret_cast<const int4*>(sCode + i);           // broadcast 128-bit loads
              const float4 l4 = *reinterpret_cast<const float4*>(sLse + i);
              const float4 d4 = *reinterpret_cast<const float4*>(sD + i);
              float s[4] = {__uint_as_float(v[x]), __uint_as_float(v[x + 1]), __uint_as_float(v[x + 2]), __uint_as_float(v[x + 3])};
              if (a.bias_table) {
                s[0] += sTable[c4.x + cj]; s[1] += sTable[c4.y + cj]; s[2] += sTable[c4.z + cj]; s[3] += sTable[c4.w + cj];
              }
              if (use_region) {
                const int4 r4 = *reinterpret_cast<const int4*>(sReg + i);
                if (r4.x != rj) s[0] -= 100.0f;
                if (r4.y != rj) s[1] -= 100.0f;
                if (r4.z != rj) s[2] -= 100.0f;
                if (r4.w != rj) s[3] -= 100.0f;
              }
              const float p0 = ex2(fmaf(s[0], TC_LOG2E, -l4.x)), p1 = ex2(fmaf(s[1], TC_LOG2E, -l4.y));
              const float p2 = ex2(fmaf(s[2], TC_LOG2E, -l4.z)), p3 = ex2(fmaf(s[3], TC_LOG2E, -l4.w));
              const float g0 = p0 * (__uint_as_float(w[x]) - d4.x), g1 = p1 * (__uint_as_float(w[x + 1]) - d4.y);
              const float g2 = p2 * (__uint_as_float(w[x + 2]) - d4.z), g3 = p3 * (__uint_as_float(w[x + 3]) - d4.w);
              pk[x >> 1] = pack_bf16(p0, p1);
              pk[(x >> 1) + 1] = pack_bf16(p2, p3);
              dk[x >> 1] = valid ? pack_bf16(g0, g1) : 0u;
              dk[(x >> 1) + 1] = valid ? pack_bf16(g2, g3) : 0u;
            }
            tmem_st_32x16(taddr + pc, pk);
            tmem_st_32x16(taddr + col_dp + pc, dk);
            // dS^T row -> shared memory (MN-major A operand of the dQ product, 64-query chunks, 128B swizzle)
            uint8_t* chunk = ds_row + (c0 >> 6) * 16384;
            const int half = (c0 >> 5) & 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int unit = (half * 4 + q) ^ (r & 7);
              *reinterpret_cast<uint4*>(chunk + unit * 16) = make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);
            }
            if (ds_g && valid) {
              uint4* g = reinterpret_cast<uint4*>(ds_g + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) g[q] = make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);
            }
          }
          tmem_st_wait();
        } else {
          for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            uint8_t* chunk = ds_row + (c0 >> 6) * 16384;
            const int half = (c0 >> 5) & 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(chunk + ((half * 4 + q) ^ (r & 7)) * 16) = make_uint4(0, 0, 0, 0);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);

        mbar_wait(mma2_done, tt & 1);
        tc_fence_after();
        if (warp_active) {
          // dV rows by warp group 0, dK rows by warp group 1
          uint32_t o[32];
          tmem_ld_32x32(taddr + (chalf ? col_dk : col_dv), o);
          tmem_ld_wait();
          if (valid) {
            uint4* g = reinterpret_cast<uint4*>(a.dqkv + ((long long)b * a.seq + j) * (3 * C) + (chalf ? 1 : 2) * C + h * TC_HD);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              g[q] = make_uint4(pack_bf16(__uint_as_float(o[q * 8]), __uint_as_float(o[q * 8 + 1])),
                                pack_bf16(__uint_as_float(o[q * 8 + 2]), __uint_as_float(o[q * 8 + 3])),
                                pack_bf16(__uint_as_float(o[q * 8 + 4]), __uint_as_float(o[q * 8 + 5])),
                                pack_bf16(__uint_as_float(o[q * 8 + 6]), __uint_as_float(o[q * 8 + 7])));
          }
        }
        if (t == a.n_kt - 1) {
          // dQ of the whole unit (all key tiles accumulated); query row i = mq*128 + r; tiles alternate between warp groups
          for (int mq = chalf; mq < a.n_mq; mq += 2) {
            const int i = mq * 128 + r;
            if (mq * 128 + quarter * 32 < a.seq) {
              uint32_t oq[32];
              tmem_ld_32x32(taddr + col_dq + mq * TC_HD, oq);
              tmem_ld_wait();
              if (i < a.seq) {
                uint4* gq = reinterpret_cast<uint4*>(a.dqkv + ((long long)b * a.seq + i) * (3 * C) + h * TC_HD);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  gq[q] = make_uint4(pack_bf16(__uint_as_float(oq[q * 8]) * a.q_scale, __uint_as_float(oq[q * 8 + 1]) * a.q_scale),
                                     pack_bf16(__uint_as_float(oq[q * 8 + 2]) * a.q_scale, __uint_as_float(oq[q * 8 + 3]) * a.q_scale),
                                     pack_bf16(__uint_as_float(oq[q * 8 + 4]) * a.q_scale, __uint_as_float(oq[q * 8 + 5]) * a.q_scale),
                                     pack_bf16(__uint_as_float(oq[q * 8 + 6]) * a.q_scale, __uint_as_float(oq[q * 8 + 7]) * a.q_scale));
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(acc_free);
          if (t == a.n_kt - 1) mbar_arrive(dq_free);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, a.tmem_cols);
}

// dTable[(code_i - code_j + off), h] += sum_b dS^T[b, h, j, i]   (bf16 dS^T, fp32 accumulation)
// grid = (seq rows j, heads, batch splits); block = nq/8 threads (8 queries = one 16-byte load each).
__global__ void dbias_reduce_kernel(const __nv_bfloat16* ds, int batch, int heads, int seq, int nq, const int* rel_code,
                                    int code_off, float* dtable) {
  const int j = blockIdx.x, h = blockIdx.y;
  const int i0 = threadIdx.x * 8;
  if (i0 >= nq) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)heads * seq * nq;
  const __nv_bfloat16* p = ds + ((long long)h * seq + j) * nq + i0;
  const int per = (batch + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * per, b1 = min(batch, b0 + per);
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + b * stride));
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
    acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
  }
  const int cj = rel_code[j];
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (i0 + e < seq) atomicAdd(dtable + (long long)(rel_code[i0 + e] - cj + code_off) * heads + h, acc[e]);
}

extern "C" long long clv_attention_bwd_tc_workspace_bytes(const clv_attn_desc_t* d, int with_dbias) {
  if (!d) return 0;
  const long long nq = (d->seq + 31) / 32 * 32;
  long long bytes = (long long)d->batch * d->heads * d->seq * 4;                       // D = rowsum(dO * O)
  if (with_dbias) bytes += (long long)d->batch * d->heads * d->seq * nq * 2 + 256;     // bf16 dS^T
  return bytes;
}

extern "C" int clv_attention_bwd_tc(const clv_attn_desc_t* d, const void* qkv, const void* out, const void* dout,
                                    const float* lse, void* dqkv, float q_scale, float* dbias_table, void* workspace,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(d && qkv && out && dout && lse && dqkv && workspace, "attention_bwd_tc: null pointer");
  CLV_REQUIRE(d->head_dim == 32 && !d->key_mask, "attention_bwd_tc: head_dim 32 without key mask only");
  CLV_REQUIRE(d->drop_p == 0.f, "attention_bwd_tc: attention dropout is not supported (attn_drop is 0 in Video Swin)");
  CLV_REQUIRE(d->seq >= 33 && d->seq <= 224, "attention_bwd_tc: seq must be in [33, 224] (got %d)", d->seq);
  CLV_REQUIRE(!dbias_table || d->bias_table, "attention_bwd_tc: dbias_table without bias_table");
  AttnTcBwdArgs a{};
  a.batch = d->batch; a.seq = d->seq; a.heads = d->heads;
  a.nq = (d->seq + 31) / 32 * 32;
  a.n_kt = (d->seq + 127) / 128;
  a.rows_per_tile = (d->seq + a.n_kt - 1) / a.n_kt;
  a.n_mq = (a.nq + 127) / 128;
  a.qb_bytes = (a.nq * TC_ROWB + 1023) / 1024 * 1024;
  const int need_cols = 2 * a.nq + a.n_mq * TC_HD;
  a.tmem_cols = need_cols <= 256 ? 256 : 512;
  a.units = (long long)d->batch * d->heads;
  // second warp group needs its own in-place packing segment that stays clear of the accumulator columns [nq-32, nq)
  a.col_split = a.nq >= 128 ? std::min(128, (a.nq - 64) / 32 * 32) : a.nq;
  if (a.col_split <= 0) a.col_split = a.nq;
  float* dsum = reinterpret_cast<float*>(workspace);
  a.lse = lse; a.dsum = dsum;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); a.q_scale = q_scale;
  const long long dsum_bytes = ((long long)d->batch * d->heads * d->seq * 4 + 255) / 256 * 256;
  a.ds_out = dbias_table ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + dsum_bytes) : nullptr;
  a.bias_table = d->bias_table; a.table_len = d->bias_table ? d->table_len : 0; a.rel_code = d->rel_code; a.code_off = d->code_off;
  a.region = d->region; a.nwin = d->nwin > 0 ? d->nwin : 1;
  const long long rows = (long long)d->batch * d->seq;
  if (int rc = launch_attn_bwd_prep(out, dout, dsum, rows, d->heads, TC_HD, d->seq, stream)) return rc;
  const long long ld = 3LL * d->heads * TC_HD, ldo = (long long)d->heads * TC_HD;
  CUtensorMap tfull, ttile, tdo;
  if (int rc = make_tmap_bf16_2d(&tfull, qkv, ld, rows, ld, TC_HD, a.nq, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&ttile, qkv, ld, rows, ld, TC_HD, 128, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tdo, dout, ldo, rows, ldo, TC_HD, a.nq, 64)) return rc;
  const size_t smem = 1024 + 4 * (size_t)a.qb_bytes + 4 * 8192 + (size_t)a.n_mq * 2 * 16384 + (size_t)((a.table_len + 3) & ~3) * 4 +
                      (size_t)a.nq * 16 + 8 + 14 * 8 + 16;
  CLV_REQUIRE(smem <= 227 * 1024, "attention_bwd_tc: %zu bytes of shared memory needed", smem);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_tc_kernel), (int)smem)) return rc;
  const int grid = (int)std::min<long long>(a.units, (long long)num_sms());
  attn_bwd_tc_kernel<<<grid, TC_BWD_THREADS, smem, stream>>>(tfull, ttile, tdo, a);
  if (int rc = after_launch("attn_bwd_tc_kernel")) return rc;
  if (dbias_table) {
    const int zsplit = std::max(1, std::min(32, d->batch / 32));
    dim3 g(d->seq, d->heads, zsplit);
    dbias_reduce_kernel<<<g, a.nq / 8, 0, stream>>>(a.ds_out, d->batch, d->heads, d->seq, a.nq, d->rel_code, d->code_off, dbias_table);
    if (int rc = after_launch("dbias_reduce_kernel")) return rc;
  }
  return 0;
}
