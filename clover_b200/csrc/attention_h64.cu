// BERT / fusion self-attention core (head_dim 64, additive per-key padding mask, optional attention-probability dropout)
// on tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces HF BertSelfAttention's softmax(Q K^T / 8 + M) V (transformers 4.6.1; reference call sites
// bert_from_hugface.py:30 and cross_transformer.py:109-110) for the text encoder (S = 32 / 40), the fusion encoder
// (S = T*49 + L = 228 / 432) and its backward.  Same packed-qkv contract as clv_attention_fwd / clv_attention_bwd.
//
//   forward   unit = (sample b, head h, 128-query tile):  S = Q K^T (M = 128, N = keys, K = 64) in tensor memory; one thread
//             per query row adds the key mask, takes max / exp2 / sum, applies the dropout keep factor and writes P back in
//             place as packed bf16; O = P V with P as the TMEM A operand and V as an MN-major shared-memory operand.
//   backward  two kernels that recompute the probabilities from the saved log-sum-exp (no atomics, no dS round trip):
//             dK|dV  unit = (b, h, 128-key tile), loop over 64-query chunks:  S^T = K Q_c^T, dP^T = V dO_c^T (N = 64) ->
//                    one thread per key row: p, dS -> packed in place -> dV += P_drop^T dO_c, dK += dS^T Q_c (A from TMEM);
//             dQ     unit = (b, h, 128-query tile), loop over 64-key chunks:  S = Q K_c^T, dP = dO V_c^T -> one thread per
//                    query row -> dS packed in place -> dQ += dS K_c (A from TMEM, K_c as MN-major operand).
//             Both hold 256 TMEM columns, so two CTAs share an SM and hide each other's MMA / softmax phases.
// Roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 one thread per TMEM lane.
#include <algorithm>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

constexpr int H64_HD = 64;
constexpr int H64_ROWB = 128;              // bytes per Q / K / V row of one head (SWIZZLE_128B)
constexpr int H64_THREADS = 192;
constexpr int H64_CH = 64;                 // chunk of the backward inner loops
constexpr float H64_LOG2E = 1.4426950408889634f;
constexpr float H64_LN2 = 0.6931471805599453f;

CLV_DEVICE float h64_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// unit u = (h * batch + b) * n_tiles + t with 32-bit divisions (the launchers check the unit count; the 64-bit forms are
// ~300-cycle subroutine calls between two tiles of every role)
CLV_DEVICE void h64_unit(uint32_t u, int n_tiles, int batch, int& t, int& b, int& h) {
  const uint32_t bh = u / (uint32_t)n_tiles;
  t = (int)(u - bh * (uint32_t)n_tiles);
  const uint32_t hh = bh / (uint32_t)batch;
  h = (int)hh;
  b = (int)(bh - hh * (uint32_t)batch);
}

struct H64Args {
  int batch, seq, heads;
  int nk;                        // forward: keys padded to a multiple of 32
  int n0, n1;                    // forward: N of the two S-MMA column chunks (n1 may be 0)
  int kv_boxes, kv_box_rows, kb_bytes, stages;
  int n_tiles;                   // 128-row tiles per (b, h)
  int n_chunks;                  // backward: 64-row chunks of the inner loop
  int tmem_cols, col_o;
  long long units;
  const float* key_mask;         // [batch, seq] additive or nullptr
  __nv_bfloat16* out; float* lse;             // forward outputs
  const float* lse_in; const float* dsum;     // backward inputs: lse [b,h,i], D [b,h,i]
  __nv_bfloat16* dqkv; float q_scale;
  uint32_t drop_thresh; float drop_inv_keep; unsigned long long drop_seed, drop_offset;   // drop_thresh == 0: off
};

// K-major operand (rows x 64 bf16, 128-byte rows, SWIZZLE_128B): k-th 16-element K step
CLV_DEVICE uint64_t h64_desc_k(uint32_t addr, int k) { return make_smem_desc(addr + k * 32, 16, 1024, 2); }
// MN-major operand ([K rows][64 bf16] tile: N = head_dim contiguous): kk-th group of 16 K rows
CLV_DEVICE uint64_t h64_desc_mn(uint32_t addr, int kk) { return make_smem_desc(addr + kk * 2048, 16, 1024, 2); }

CLV_DEVICE void h64_store_row64(__nv_bfloat16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32], float scale) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d[q] = make_uint4(pack_bf16(__uint_as_float(a[q * 8]) * scale, __uint_as_float(a[q * 8 + 1]) * scale),
                      pack_bf16(__uint_as_float(a[q * 8 + 2]) * scale, __uint_as_float(a[q * 8 + 3]) * scale),
                      pack_bf16(__uint_as_float(a[q * 8 + 4]) * scale, __uint_as_float(a[q * 8 + 5]) * scale),
                      pack_bf16(__uint_as_float(a[q * 8 + 6]) * scale, __uint_as_float(a[q * 8 + 7]) * scale));
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d[4 + q] = make_uint4(pack_bf16(__uint_as_float(b[q * 8]) * scale, __uint_as_float(b[q * 8 + 1]) * scale),
                          pack_bf16(__uint_as_float(b[q * 8 + 2]) * scale, __uint_as_float(b[q * 8 + 3]) * scale),
                          pack_bf16(__uint_as_float(b[q * 8 + 4]) * scale, __uint_as_float(b[q * 8 + 5]) * scale),
                          pack_bf16(__uint_as_float(b[q * 8 + 6]) * scale, __uint_as_float(b[q * 8 + 7]) * scale));
}

// =================================================================================================================
// Forward
// =================================================================================================================
__global__ void __launch_bounds__(H64_THREADS, 1)
attn64_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, H64Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = 16384 + 2 * a.kb_bytes;
  float* sMask = reinterpret_cast<float*>(smem + a.stages * stage_bytes);     // [nk] key mask * log2 e (-1e30 beyond seq)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + a.nk);
  uint64_t* full_bar = bars;          // [2]
  uint64_t* empty_bar = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * H64_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1); mbar_init(s_free, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t nst = (uint32_t)a.stages;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        int t, b, h;
        h64_unit((uint32_t)u, a.n_tiles, a.batch, t, b, h);
        const uint32_t stage = it % nst, phase = (it / nst) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sQ = smem + stage * stage_bytes;
        uint8_t* sK = sQ + 16384;
        uint8_t* sV = sK + a.kb_bytes;
        mbar_expect_tx(&full_bar[stage], 16384 + 2 * a.nk * H64_ROWB);
        const int row0 = b * a.seq;
        tma_load_2d(sQ, &tm_q, &full_bar[stage], h * H64_HD, row0 + t * 128);
        for (int i = 0; i < a.kv_boxes; ++i) {
          tma_load_2d(sK + i * a.kv_box_rows * H64_ROWB, &tm_kv, &full_bar[stage], C + h * H64_HD, row0 + i * a.kv_box_rows);
          tma_load_2d(sV + i * a.kv_box_rows * H64_ROWB, &tm_kv, &full_bar[stage], 2 * C + h * H64_HD, row0 + i * a.kv_box_rows);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_pv = make_idesc_bf16(128, H64_HD, 0, 1);
      const uint32_t idesc_s0 = make_idesc_bf16(128, a.n0, 0, 0);
      const uint32_t idesc_s1 = make_idesc_bf16(128, a.n1 > 0 ? a.n1 : 16, 0, 0);
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        const uint32_t stage = it % nst, phase = (it / nst) & 1;
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(s_free, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t k_addr = q_addr + 16384;
        const uint32_t v_addr = k_addr + a.kb_bytes;
#pragma unroll
        for (int k = 0; k < H64_HD / 16; ++k)
          umma_bf16_ss(tmem_base, h64_desc_k(q_addr, k), h64_desc_k(k_addr, k), idesc_s0, k > 0);
        if (a.n1 > 0) {
#pragma unroll
          for (int k = 0; k < H64_HD / 16; ++k)
            umma_bf16_ss(tmem_base + a.n0, h64_desc_k(q_addr, k), h64_desc_k(k_addr + a.n0 * H64_ROWB, k), idesc_s1, k > 0);
        }
        umma_commit(s_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        const uint32_t tmem_o = tmem_base + a.col_o;
        for (int kk = 0; kk < a.nk / 16; ++kk)
          umma_bf16_ts(tmem_o, tmem_base + kk * 8, h64_desc_mn(v_addr, kk), idesc_pv, kk > 0);
        umma_commit(o_full);
        umma_commit(&empty_bar[stage]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int cur_b = -1;
    uint32_t it = 0;
    for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
      int t, b, h;
      h64_unit((uint32_t)u, a.n_tiles, a.batch, t, b, h);
      const int i = t * 128 + r;
      const bool valid = i < a.seq;
      const bool warp_active = (t * 128 + quarter * 32) < a.seq;
      if (b != cur_b) {                       // key mask of this sample (the previous unit's readers are past pass 1)
        named_bar_sync(1, 128);
        const float* km = a.key_mask ? a.key_mask + (long long)b * a.seq : nullptr;
        for (int j = tid; j < a.nk; j += 128) sMask[j] = j < a.seq ? (km ? km[j] * H64_LOG2E : 0.f) : -1.0e30f;
        cur_b = b;
        named_bar_sync(1, 128);
      }
      const unsigned long long drop_row =
          a.drop_offset + (((unsigned long long)b * a.heads + h) * a.seq + (unsigned)(valid ? i : 0)) * a.seq;

      mbar_wait(s_full, it & 1);
      tc_fence_after();
      float m = -1.0e30f, l = 0.f;
      if (warp_active) {
        // ---- pass 1: x = s * log2 e + mask; row max; x written back
#pragma unroll 1
        for (int c0 = 0; c0 < a.nk; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 mk = *reinterpret_cast<const float4*>(sMask + c0 + j);          // broadcast 128-bit load
            v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), H64_LOG2E, mk.x));
            v[j + 1] = __float_as_uint(fmaf(__uint_as_float(v[j + 1]), H64_LOG2E, mk.y));
            v[j + 2] = __float_as_uint(fmaf(__uint_as_float(v[j + 2]), H64_LOG2E, mk.z));
            v[j + 3] = __float_as_uint(fmaf(__uint_as_float(v[j + 3]), H64_LOG2E, mk.w));
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
        // ---- pass 2: p = 2^(x - m); row sum over ALL terms; dropout keep factor on the stored P; packed bf16 in place
#pragma unroll 1
        for (int c0 = 0; c0 < a.nk; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p0 = h64_ex2(__uint_as_float(v[2 * j]) - m);
            float p1 = h64_ex2(__uint_as_float(v[2 * j + 1]) - m);
            l += p0 + p1;
            if (a.drop_thresh) {
              p0 *= keep_scale(a.drop_seed, drop_row + (unsigned)(c0 + 2 * j), a.drop_thresh, a.drop_inv_keep);
              p1 *= keep_scale(a.drop_seed, drop_row + (unsigned)(c0 + 2 * j + 1), a.drop_thresh, a.drop_inv_keep);
            }
            pk[j] = pack_bf16(p0, p1);
          }
          tmem_st_32x16(taddr + (c0 >> 1), pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      // ---- epilogue: O / l -> bf16 -> global; lse (natural log)
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      if (warp_active) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(taddr + a.col_o, o0);
        tmem_ld_32x32(taddr + a.col_o + 32, o1);
        tmem_ld_wait();
        if (valid) {
          h64_store_row64(a.out + ((long long)b * a.seq + i) * C + h * H64_HD, o0, o1, 1.0f / l);
          a.lse[((long long)b * a.heads + h) * a.seq + i] = (m + log2f(l)) * H64_LN2;
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

// =================================================================================================================
// Backward.  ROWS_ARE_KEYS = true : dK | dV kernel (TMEM lanes = keys of the tile, chunks = 64 queries)
//            ROWS_ARE_KEYS = false: dQ kernel      (TMEM lanes = queries of the tile, chunks = 64 keys)
// TMEM columns: [0,64) S / S^T chunk (packed P^T or dS written in place), [64,128) dP / dP^T chunk (packed dS^T in place),
//               [128,192) dV or dQ accumulator, [192,256) dK accumulator.
// Shared memory: resident tile pair (K_t,V_t | Q_t,dO_t) 2 x 16 KB x 2 stages, chunk pair 2 x 8 KB x 2 stages,
//                per-(b,h) vectors: column terms (-lse * log2 e, D) or (mask * log2 e).
// =================================================================================================================
template <bool ROWS_ARE_KEYS>
__global__ void __launch_bounds__(H64_THREADS, 2)
attn64_bwd_kernel(const __grid_constant__ CUtensorMap tm_tile, const __grid_constant__ CUtensorMap tm_tile_do,
                  const __grid_constant__ CUtensorMap tm_chunk, const __grid_constant__ CUtensorMap tm_chunk_do, H64Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sTile = smem;                          // [2][A 16 KB | B 16 KB]   (K_t | V_t) or (Q_t | dO_t)
  uint8_t* sChunk = smem + 2 * 32768;             // [2][A 8 KB | B 8 KB]     (Q_c | dO_c) or (K_c | V_c)
  const int ncol = a.n_chunks * H64_CH;
  float* sCol0 = reinterpret_cast<float*>(sChunk + 2 * 16384);   // [2][ncol]: keys-kernel: -lse_i * log2 e ; dQ-kernel: mask_j * log2 e
  float* sCol1 = sCol0 + 2 * ncol;                               // [2][ncol]: keys-kernel: D_i
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCol1 + 2 * ncol);
  uint64_t* tile_full = bars;         // [2]
  uint64_t* tile_empty = bars + 2;    // [2]
  uint64_t* ch_full = bars + 4;       // [2]
  uint64_t* ch_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* p_ready = bars + 9;
  uint64_t* acc_full = bars + 10;
  uint64_t* acc_free = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * H64_HD;
  // column offsets inside the packed qkv row of the tile operands and of the chunk operands
  const int tile_col = ROWS_ARE_KEYS ? C : 0;          // K (then V at +C)   |  Q (dO comes from its own tensor)
  const int chunk_col = ROWS_ARE_KEYS ? 0 : C;         // Q (dO own tensor)  |  K (then V at +C)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_tile); tma_prefetch_desc(&tm_tile_do); tma_prefetch_desc(&tm_chunk); tma_prefetch_desc(&tm_chunk_do);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tile_full[s], 1); mbar_init(&tile_empty[s], 1); mbar_init(&ch_full[s], 1); mbar_init(&ch_empty[s], 1);
    }
    mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(acc_full, 1); mbar_init(acc_free, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0, cc = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        int t, b, h;
        h64_unit((uint32_t)u, a.n_tiles, a.batch, t, b, h);
        const int row0 = b * a.seq;
        const uint32_t us = it & 1;
        mbar_wait(&tile_empty[us], ((it >> 1) & 1) ^ 1);
        uint8_t* sA = sTile + us * 32768;
        mbar_expect_tx(&tile_full[us], 32768);
        tma_load_2d(sA, &tm_tile, &tile_full[us], tile_col + h * H64_HD, row0 + t * 128);
        if (ROWS_ARE_KEYS) tma_load_2d(sA + 16384, &tm_tile, &tile_full[us], 2 * C + h * H64_HD, row0 + t * 128);
        else tma_load_2d(sA + 16384, &tm_tile_do, &tile_full[us], h * H64_HD, row0 + t * 128);
        for (int c = 0; c < a.n_chunks; ++c, ++cc) {
          const uint32_t cs = cc & 1;
          mbar_wait(&ch_empty[cs], ((cc >> 1) & 1) ^ 1);
          uint8_t* sC = sChunk + cs * 16384;
          mbar_expect_tx(&ch_full[cs], 16384);
          tma_load_2d(sC, &tm_chunk, &ch_full[cs], chunk_col + h * H64_HD, row0 + c * H64_CH);
          if (ROWS_ARE_KEYS) tma_load_2d(sC + 8192, &tm_chunk_do, &ch_full[cs], h * H64_HD, row0 + c * H64_CH);
          else tma_load_2d(sC + 8192, &tm_chunk, &ch_full[cs], 2 * C + h * H64_HD, row0 + c * H64_CH);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(128, H64_CH, 0, 0);        // scores: both operands K-major
      const uint32_t idesc_acc = make_idesc_bf16(128, H64_HD, 0, 1);      // accumulators: A from TMEM, B MN-major
      uint32_t it = 0, cc = 0;
      for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
        const uint32_t us = it & 1;
        mbar_wait(&tile_full[us], (it >> 1) & 1);
        const uint32_t ta_addr = smem_u32(sTile + us * 32768), tb_addr = ta_addr + 16384;
        for (int c = 0; c < a.n_chunks; ++c, ++cc) {
          const uint32_t cs = cc & 1;
          mbar_wait(&ch_full[cs], (cc >> 1) & 1);
          tc_fence_after();
          const uint32_t ca_addr = smem_u32(sChunk + cs * 16384), cb_addr = ca_addr + 8192;
          // S(^T) chunk = tileA . chunkA^T ; dP(^T) chunk = tileB . chunkB^T     (M = 128 tile rows, N = 64 chunk rows, K = 64)
#pragma unroll
          for (int k = 0; k < H64_HD / 16; ++k)
            umma_bf16_ss(tmem_base, h64_desc_k(ta_addr, k), h64_desc_k(ca_addr, k), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < H64_HD / 16; ++k)
            umma_bf16_ss(tmem_base + 64, h64_desc_k(tb_addr, k), h64_desc_k(cb_addr, k), idesc_s, k > 0);
          umma_commit(s_full);
          mbar_wait(p_ready, cc & 1);
          if (c == 0) mbar_wait(acc_free, (it & 1) ^ 1);        // the previous unit's accumulators have been read
          tc_fence_after();
          if (ROWS_ARE_KEYS) {
            // dV_t += P_drop^T_c dO_c (packed at [0,32)) ; dK_t += dS^T_c Q_c (packed at [64,96));  K = 64 queries
#pragma unroll
            for (int kk = 0; kk < H64_CH / 16; ++kk) {
              umma_bf16_ts(tmem_base + 128, tmem_base + kk * 8, h64_desc_mn(cb_addr, kk), idesc_acc, (c > 0 || kk > 0) ? 1u : 0u);
              umma_bf16_ts(tmem_base + 192, tmem_base + 64 + kk * 8, h64_desc_mn(ca_addr, kk), idesc_acc, (c > 0 || kk > 0) ? 1u : 0u);
            }
          } else {
            // dQ_t += dS_c K_c (packed at [0,32));  K = 64 keys
#pragma unroll
            for (int kk = 0; kk < H64_CH / 16; ++kk)
              umma_bf16_ts(tmem_base + 128, tmem_base + kk * 8, h64_desc_mn(ca_addr, kk), idesc_acc, (c > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&ch_empty[cs]);
        }
        umma_commit(acc_full);
        umma_commit(&tile_empty[us]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    long long cur_bh = -1;
    uint32_t it = 0, cc = 0, nvec = 0;
    for (long long u = blockIdx.x; u < a.units; u += gridDim.x, ++it) {
      int t, b, h;
      h64_unit((uint32_t)u, a.n_tiles, a.batch, t, b, h);
      const int row = t * 128 + r;                          // key index j (dK|dV kernel) or query index i (dQ kernel)
      const bool valid = row < a.seq;
      const long long stat = ((long long)b * a.heads + h) * a.seq;
      const long long bh = (long long)h * a.batch + b;
      // ---- per-(b,h) column vectors, double-buffered so that a unit never waits for the previous one's readers
      if (bh != cur_bh) {
        nvec ^= 1;
        float* c0 = sCol0 + nvec * ncol;
        float* c1 = sCol1 + nvec * ncol;
        for (int x = tid; x < ncol; x += 128) {
          if (ROWS_ARE_KEYS) {        // columns are queries: -lse_i * log2 e (+huge beyond seq -> p = 0) and D_i
            c0[x] = x < a.seq ? -a.lse_in[stat + x] * H64_LOG2E : -1.0e30f;
            c1[x] = x < a.seq ? a.dsum[stat + x] : 0.f;
          } else {                    // columns are keys: mask_j * log2 e (-huge beyond seq)
            c0[x] = x < a.seq ? (a.key_mask ? a.key_mask[(long long)b * a.seq + x] * H64_LOG2E : 0.f) : -1.0e30f;
          }
        }
        cur_bh = bh;
        named_bar_sync(1, 128);
      }
      const float* col0 = sCol0 + nvec * ncol;
      const float* col1 = sCol1 + nvec * ncol;
      // per-row constants
      float row_c, row_d = 0.f;
      if (ROWS_ARE_KEYS) {
        row_c = (valid && a.key_mask) ? a.key_mask[(long long)b * a.seq + row] * H64_LOG2E : 0.f;
      } else {
        row_c = valid ? -a.lse_in[stat + row] * H64_LOG2E : -1.0e30f;
        row_d = valid ? a.dsum[stat + row] : 0.f;
      }
      const unsigned long long drop_bh = a.drop_offset + (unsigned long long)((long long)b * a.heads + h) * a.seq * a.seq;
      const bool warp_active = (t * 128 + quarter * 32) < a.seq;

      for (int c = 0; c < a.n_chunks; ++c, ++cc) {
        mbar_wait(s_full, cc & 1);
        tc_fence_after();
        if (warp_active) {
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t s[32], dp[32], pk[16], dk[16];
            tmem_ld_32x32(taddr + hf * 32, s);
            tmem_ld_32x32(taddr + 64 + hf * 32, dp);
            tmem_ld_wait();
            const int x0 = c * H64_CH + hf * 32;             // first column (query i or key j) of this half
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t0 = *reinterpret_cast<const float4*>(col0 + x0 + j);
              float4 t1 = make_float4(row_d, row_d, row_d, row_d);
              if (ROWS_ARE_KEYS) t1 = *reinterpret_cast<const float4*>(col1 + x0 + j);
              const float tc[4] = {t0.x, t0.y, t0.z, t0.w}, td[4] = {t1.x, t1.y, t1.z, t1.w};
              float p[4], g[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                p[e] = h64_ex2(fmaf(__uint_as_float(s[j + e]), H64_LOG2E, tc[e] + row_c));
                float dpe = __uint_as_float(dp[j + e]);
                float pd = p[e];
                if (a.drop_thresh) {
                  const unsigned col = (unsigned)(x0 + j + e);
                  const unsigned long long idx = ROWS_ARE_KEYS ? drop_bh + (unsigned long long)col * a.seq + (unsigned)(valid ? row : 0)
                                                               : drop_bh + (unsigned long long)(valid ? row : 0) * a.seq + col;
                  const float ks = keep_scale(a.drop_seed, idx, a.drop_thresh, a.drop_inv_keep);
                  dpe *= ks;                               // dP = dP_dropped * keep / (1 - p)
                  pd *= ks;                                // dV uses the dropped probabilities
                }
                g[e] = p[e] * (dpe - td[e]);
                p[e] = pd;
              }
              pk[j >> 1] = pack_bf16(p[0], p[1]); pk[(j >> 1) + 1] = pack_bf16(p[2], p[3]);
              dk[j >> 1] = pack_bf16(g[0], g[1]); dk[(j >> 1) + 1] = pack_bf16(g[2], g[3]);
            }
            // packed operands: keys kernel P_drop^T at [0,32) and dS^T at [64,96); dQ kernel dS at [0,32).  Half hf fills packed
            // columns [16 hf, 16 hf + 16): half 0's stores land in [0,16) / [64,80), which half 1 (fp32 columns [32,64) /
            // [96,128)) never reads, and half 1's stores land in [16,32) / [80,96), already consumed by half 0.
            if (ROWS_ARE_KEYS) {
              tmem_st_32x16(taddr + hf * 16, pk);
              tmem_st_32x16(taddr + 64 + hf * 16, dk);
            } else {
              tmem_st_32x16(taddr + hf * 16, dk);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
      }

      // ---- unit epilogue: accumulators -> bf16 -> packed dqkv rows
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      if (warp_active) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(taddr + 128, o0);
        tmem_ld_32x32(taddr + 160, o1);
        tmem_ld_wait();
        __nv_bfloat16* grow = a.dqkv + ((long long)b * a.seq + row) * (3 * C) + h * H64_HD;
        if (ROWS_ARE_KEYS) {
          if (valid) h64_store_row64(grow + 2 * C, o0, o1, 1.0f);        // dV
          tmem_ld_32x32(taddr + 192, o0);
          tmem_ld_32x32(taddr + 224, o1);
          tmem_ld_wait();
          if (valid) h64_store_row64(grow + C, o0, o1, 1.0f);            // dK (q is pre-scaled, so dS^T q_scaled is d/dk)
        } else {
          if (valid) h64_store_row64(grow, o0, o1, a.q_scale);           // dQ w.r.t. the unscaled q
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

static int h64_checks(const clv_attn_desc_t* d, const char* who) {
  CLV_REQUIRE(d != nullptr, "%s: null descriptor", who);
  CLV_REQUIRE(d->head_dim == 64 && !d->bias_table && !d->region, "%s: head_dim 64 with an optional key mask only", who);
  CLV_REQUIRE(d->batch > 0 && d->heads > 0 && d->seq >= 1 && d->seq <= 448, "%s: seq must be in [1, 448] (got %d)", who, d->seq);
  CLV_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f, "%s: drop_p must be in [0, 1)", who);
  CLV_REQUIRE((long long)d->batch * d->seq < 2000000000LL, "%s: too many rows", who);
  return 0;
}

static void h64_fill(H64Args& a, const clv_attn_desc_t* d) {
  a = H64Args{};
  a.batch = d->batch; a.seq = d->seq; a.heads = d->heads;
  a.key_mask = d->key_mask;
  a.n_tiles = (d->seq + 127) / 128;
  a.n_chunks = (d->seq + H64_CH - 1) / H64_CH;
  a.units = (long long)d->batch * d->heads * a.n_tiles;
  if (d->drop_p > 0.f) {
    a.drop_thresh = drop_threshold(d->drop_p); a.drop_inv_keep = 1.0f / (1.0f - d->drop_p);
    a.drop_seed = d->drop_seed; a.drop_offset = d->drop_offset;
  }
}

}  // namespace clv

using namespace clv;

extern "C" int clv_attention_tc64_supported(int seq) { return seq >= 1 && seq <= 448; }

extern "C" int clv_attention_fwd_tc64(const clv_attn_desc_t* d, const void* qkv, void* out, float* lse, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = h64_checks(d, "attention_fwd_tc64")) return rc;
  CLV_REQUIRE(qkv && out && lse, "attention_fwd_tc64: null pointer");
  H64Args a; h64_fill(a, d);
  CLV_REQUIRE(a.units < (1LL << 31), "attention_tc64: %lld units exceed the 32-bit unit index", a.units);
  a.nk = (d->seq + 31) / 32 * 32;
  a.n0 = a.nk <= 256 ? a.nk : ((a.nk / 2 + 15) & ~15);
  a.n1 = a.nk - a.n0;
  a.kv_boxes = a.nk <= 256 ? 1 : 2;
  a.kv_box_rows = a.nk / a.kv_boxes;
  a.kb_bytes = (a.nk * H64_ROWB + 1023) / 1024 * 1024;
  // O accumulator: behind the scores when they are short; otherwise it aliases the LAST 64 score columns, which are dead once
  // pass 2 has packed P into columns [0, nk/2) (nk >= 128 keeps the two apart) -- 228 keys then fit 256 columns and two
  // CTAs share an SM, each hiding the other's serial S-MMA -> softmax -> PV-MMA -> epilogue chain
  a.col_o = a.nk >= 128 ? a.nk - H64_HD : a.nk;
  const int need = std::max(a.nk, a.col_o + H64_HD);
  a.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
  a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse;
  const size_t stage = 16384 + 2 * (size_t)a.kb_bytes;
  // two smem stages only when two CTAs per SM still fit next to each other (small sequences); else one stage, more CTAs
  a.stages = (2 * stage + (size_t)a.nk * 4 + 2048 <= 100 * 1024) ? 2 : 1;
  const long long rows = (long long)d->batch * d->seq;
  const long long ld = 3LL * d->heads * H64_HD;
  CUtensorMap tq, tkv;
  if (int rc = make_tmap_bf16_2d(&tq, qkv, ld, rows, ld, H64_HD, 128, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&tkv, qkv, ld, rows, ld, H64_HD, a.kv_box_rows, 128)) return rc;
  const size_t smem = 1024 + a.stages * stage + (size_t)a.nk * 4 + 128 + 64;
  CLV_REQUIRE(smem <= 227 * 1024, "attention_fwd_tc64: %zu bytes of shared memory needed", smem);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn64_fwd_kernel), (int)smem)) return rc;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(512 / a.tmem_cols, (220 * 1024) / smem), 4));
  const int grid = (int)std::min<long long>(a.units, (long long)num_sms() * per_sm);
  attn64_fwd_kernel<<<grid, H64_THREADS, smem, stream>>>(tq, tkv, a);
  return after_launch("attn64_fwd_kernel");
}

extern "C" int clv_attention_bwd_tc64(const clv_attn_desc_t* d, const void* qkv, const void* out, const void* dout,
                                      const float* lse, void* dqkv, float q_scale, float* dsum_ws, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = h64_checks(d, "attention_bwd_tc64")) return rc;
  CLV_REQUIRE(qkv && out && dout && lse && dqkv && dsum_ws, "attention_bwd_tc64: null pointer");
  H64Args a; h64_fill(a, d);
  CLV_REQUIRE(a.units < (1LL << 31), "attention_tc64: %lld units exceed the 32-bit unit index", a.units);
  a.lse_in = lse; a.dsum = dsum_ws; a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); a.q_scale = q_scale;
  const long long rows = (long long)d->batch * d->seq;
  if (int rc = launch_attn_bwd_prep(out, dout, dsum_ws, rows, d->heads, H64_HD, d->seq, stream)) return rc;
  const long long ld = 3LL * d->heads * H64_HD, ldo = (long long)d->heads * H64_HD;
  CUtensorMap t_tile, t_tile_do, t_chunk, t_chunk_do;
  if (int rc = make_tmap_bf16_2d(&t_tile, qkv, ld, rows, ld, H64_HD, 128, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&t_tile_do, dout, ldo, rows, ldo, H64_HD, 128, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&t_chunk, qkv, ld, rows, ld, H64_HD, H64_CH, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&t_chunk_do, dout, ldo, rows, ldo, H64_HD, H64_CH, 128)) return rc;
  const size_t smem = 1024 + 2 * 32768 + 2 * 16384 + 4 * (size_t)a.n_chunks * H64_CH * 4 + 128 + 64;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn64_bwd_kernel<true>), (int)smem)) return rc;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn64_bwd_kernel<false>), (int)smem)) return rc;
  const int grid = (int)std::min<long long>(a.units, (long long)num_sms() * 2);
  attn64_bwd_kernel<true><<<grid, H64_THREADS, smem, stream>>>(t_tile, t_tile_do, t_chunk, t_chunk_do, a);
  if (int rc = after_launch("attn64_bwd_kernel<dkv>")) return rc;
  attn64_bwd_kernel<false><<<grid, H64_THREADS, smem, stream>>>(t_tile, t_tile_do, t_chunk, t_chunk_do, a);
  return after_launch("attn64_bwd_kernel<dq>");
}
