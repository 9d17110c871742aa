// Window attention core for the Video Swin windows actually used by Clover: head_dim 32, spatial window 7x7,
// tokens in (d, h, w) order, N = 49 * wd (wd = 2, 4, 6, 8 frames-of-tokens).  tcgen05 / TMEM / TMA, sm_100a.
//
// The first tcgen05 version (attention_tc.cu, kept for other window shapes) spends 15 (forward) / 26 (backward)
// issued instructions per score element on index arithmetic, mask compares, broadcast loads and selects while the
// tensor pipe idles at < 1 %.  The softmax exponent (MUFU, 16/clk/SM) allows 8 issue slots per element per warp,
// so this version moves everything that is not the exponent off the CUDA cores:
//   * relative-position bias (swin_transformer_3d.py:345-359,382-385): idx(i,j) = code_i - code_j + off with
//     code = d*169 + h*13 + w.  The column token of every TMEM element is a compile-time constant here, so
//     the gather is ONE shared-memory load with an immediate offset from a per-thread base (table staged per head,
//     pre-multiplied by log2 e) and the add is folded into the log2-domain FMA;
//   * shift mask (compute_mask :548-562, added at :388-390): 0 / -100 by region equality.  Regions factor per axis
//     (3 x 3 x 3), so "+100 per matching axis, -300" is a rank-11 term; it rides in a 16-wide K-extension of the
//     Q K^T product (bf16 one-hots scaled by 10; -256 and -44 against two ones columns) -- an entry that differs in
//     1..3 axes gets -100..-300, all of which vanish from the softmax exactly like the reference's -100;
//   * backward: -lse_i and -D_i (hi/lo bf16 pairs) ride in the same K-extension of K Q^T and V dO^T, so a thread
//     computes p = ex2(fma(s, log2e, bias)) and ds = p * dp and nothing else; rows outside the window are never
//     touched (their dS^T rows in shared memory stay zero from the start).
// Layouts, barriers and roles follow attention_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, the rest one thread
// per TMEM lane (query row forward / key row backward).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

constexpr int W7_HD = 32;
constexpr int W7_ROWB = 64;            // bytes per Q/K/V row of one head
constexpr int W7_XROWB = 32;           // bytes per K-extension row (16 bf16)
constexpr int W7_TILE = 98;            // rows per tile = two 49-token slabs
constexpr int W7_SH = 169, W7_SW = 13; // code strides of the reference's bias table for a (., 7, 7) window: (2*7-1)^2 and 2*7-1
// Strides of the STAGED table the kernels look up (attn_w7_table_kernel re-lays the used depth range out with them): both are
// unique decodes (PW >= 13, PH >= 12 * PW + 13) and PW = 7, PH = 49 (mod 32), so the staged address of token t = 49 d + 7 h + w
// is t (mod 32) plus a warp-uniform constant -- the 32 lanes of a warp (32 consecutive query rows forward, key rows backward)
// hit 32 different banks.  With the reference strides 20-40 % of the lanes collided pairwise and every bias load took two
// shared-memory wavefronts (ncu: 22.5 M wavefronts for 12.8 M loads, the busiest unit of both kernels).
constexpr int W7_PH = 497, W7_PW = 39;
static_assert(W7_PW % 32 == 7 && W7_PH % 32 == 49 % 32 && W7_PW >= 13 && W7_PH >= 12 * W7_PW + 13, "staged-table strides");
constexpr float W7_LOG2E = 1.4426950408889634f;
constexpr float W7_LN2 = 0.6931471805599453f;

// code of body column c (two slabs of 49 tokens) relative to the body's first slab
__host__ __device__ constexpr int w7_code(int c) { return (c / 49) * W7_PH + ((c % 49) / 7) * W7_PW + (c % 7); }

CLV_DEVICE float w7_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
CLV_DEVICE float w7_max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// ---- TMEM access shapes not in common.cuh -------------------------------------------------------------------
CLV_DEVICE void tmem_ld_32x2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
CLV_DEVICE void tmem_ld_32x4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
CLV_DEVICE void tmem_st_32x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
CLV_DEVICE void tmem_st_32x2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
CLV_DEVICE void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
CLV_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// =================================================================================================================
// Forward.  Unit = (window b, head h, query tile t of 98 rows).
// =================================================================================================================
struct W7FwdArgs {
  int batch, heads, seq, n_qt;
  int tile_rows;                // fwd2: rows of every query tile but the last (a multiple of 32)
  int nmma;                     // S columns issued to the tensor core: seq rounded up to 16
  int n0;                       // first N chunk of the S product (<= 256), the rest is nmma - n0
  int kb_bytes, kx_bytes;       // bytes reserved per K (or V) buffer / per key-side extension buffer (multiples of 1024)
  int tmem_cols, col_o;
  long long units;
  __nv_bfloat16* out; float* lse;
  const float* table_t; int table_len, table_ld; int code_off;   // [heads, table_ld] fp32, pre-multiplied by log2 e
  int has_ext, nwin;
  int early;                    // score MMAs of tile i+1 issued before the O read-out of tile i has been acknowledged
  long long* dbg;               // phase-cycle sums (DBG instantiation only)
  int q_rows;                   // query rows per tile: 98 (two slabs) or 128 (every TMEM lane used; the last tile is ragged)
  int split_body;               // > 0: the P V products of key bodies [0, split_body) are issued while the softmax still packs the rest
};

constexpr int W7_FWD_THREADS = 192;

// bias table -> staged layout [heads][ld], times log2 e: entry (dz + wd - 1) * PH + (dy + 6) * PW + (dx + 6) of head h holds
// table[(dz + cfg_wd - 1) * 169 + (dy + 6) * 13 + (dx + 6), h] for the 2 wd - 1 depth offsets a wd-deep window can produce
// (one coalesced 6-30 KB copy per head change in the kernels); padding entries are zero and never read
__global__ void attn_w7_table_kernel(const float* table, float* table_t, int ld, int heads, int wd, int cfg_wd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ld * heads) return;
  const int h = idx / ld, x = idx - h * ld;
  const int zz = x / W7_PH, rem = x - zz * W7_PH, yy = rem / W7_PW, xx = rem - yy * W7_PW;
  const bool used = zz <= 2 * (wd - 1) && yy <= 12 && xx <= 12;
  const int src = (zz - (wd - 1) + cfg_wd - 1) * W7_SH + yy * W7_SW + xx;
  table_t[idx] = used ? table[(long long)src * heads + h] * W7_LOG2E : 0.f;
}

// contiguous unit range of CTA c: consecutive units share the head (slowest index), so the staged table is reloaded
// at most once or twice per CTA
CLV_DEVICE void w7_unit_range(long long units, long long& u0, long long& u1) {
  const long long base = units / gridDim.x, rem = units % gridDim.x;
  const long long c = blockIdx.x;
  u0 = c * base + (c < rem ? c : rem);
  u1 = u0 + base + (c < rem ? 1 : 0);
}

// (tile t, window b, head h) of unit u = (h * batch + b) * n_t + t, advanced incrementally: the 64-bit divisions by run-time
// values cost several hundred cycles per unit when they sit between two tiles of the softmax warps
struct W7Idx {
  int t, b, h;
  CLV_DEVICE void init(long long u, int n_t, int batch) {
    t = (int)(u % n_t);
    const long long bh = u / n_t;
    b = (int)(bh % batch); h = (int)(bh / batch);
  }
  CLV_DEVICE void next(int n_t, int batch) {
    if (++t == n_t) { t = 0; if (++b == batch) { b = 0; ++h; } }
  }
};

// pass 1 on CNT (<= 32) consecutive body columns starting at static column C0: x = s*log2e + bias*log2e; running max
// (four independent running maxima: a single one is a chain of 98 dependent FMNMX3 per row that the in-order issue cannot hide)
template <int C0, int CNT, int NREG>
CLV_DEVICE void w7_fwd_p1(uint32_t (&v)[NREG], const float* tb, float (&m)[4]) {
#pragma unroll
  for (int x = 0; x < CNT; ++x) v[x] = __float_as_uint(fmaf(__uint_as_float(v[x]), W7_LOG2E, tb[-w7_code(C0 + x)]));
#pragma unroll
  for (int x = 0; x + 1 < CNT; x += 2) m[(x >> 1) & 3] = w7_max3(m[(x >> 1) & 3], __uint_as_float(v[x]), __uint_as_float(v[x + 1]));
  if (CNT & 1) m[0] = fmaxf(m[0], __uint_as_float(v[CNT - 1]));
}
// pass 2: p = 2^(x - m), partial row sums, packed bf16 pairs
template <int CNT, int NREG>
CLV_DEVICE void w7_fwd_p2(const uint32_t (&v)[NREG], float m, float& l0, float& l1, uint32_t* pk) {
#pragma unroll
  for (int x = 0; x < CNT; x += 2) {
    const float p0 = w7_ex2(__uint_as_float(v[x]) - m), p1 = w7_ex2(__uint_as_float(v[x + 1]) - m);
    l0 += p0; l1 += p1;
    pk[x >> 1] = pack_bf16(p0, p1);
  }
}

// DBG: tools/attn_microbench.py --phases 1 only (tunable w7_fwd_dbg = device buffer) -- softmax warp 0 of every CTA sums the cycles it spends in each phase of a tile
// (wait for S, pass 1, pass 2, wait for O, epilogue) into a.dbg[blockIdx.x * 8 ..]
template <bool DBG>
__global__ void __launch_bounds__(W7_FWD_THREADS, 2)
attn_w7_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv0,
                   const __grid_constant__ CUtensorMap tm_kv1, const __grid_constant__ CUtensorMap tm_qx,
                   const __grid_constant__ CUtensorMap tm_kx0, const __grid_constant__ CUtensorMap tm_kx1, W7FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = 8192 + 2 * a.kb_bytes + (a.has_ext ? 4096 + a.kx_bytes : 0);
  float* sTable = reinterpret_cast<float*>(smem + 2 * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTable + a.table_ld);
  uint64_t* full_bar = bars;          // [2]
  uint64_t* empty_bar = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint64_t* p_half = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * W7_HD;
  const int n1 = a.nmma - a.n0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv0);
    if (n1 > 0) tma_prefetch_desc(&tm_kv1);
    if (a.has_ext) { tma_prefetch_desc(&tm_qx); tma_prefetch_desc(&tm_kx0); }
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1); mbar_init(s_free, 4); mbar_init(p_half, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long u_begin, u_end;
  w7_unit_range(a.units, u_begin, u_end);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      W7Idx ix; ix.init(u_begin, a.n_qt, a.batch);
      for (long long u = u_begin; u < u_end; ++u, ++it, ix.next(a.n_qt, a.batch)) {
        const int t = ix.t, b = ix.b, h = ix.h;
        const int stage = it & 1;
        mbar_wait(&empty_bar[stage], ((it >> 1) & 1) ^ 1);
        uint8_t* sQ = smem + stage * stage_bytes;
        uint8_t* sK = sQ + 8192;
        uint8_t* sV = sK + a.kb_bytes;
        uint8_t* sQx = sV + a.kb_bytes;
        uint8_t* sKx = sQx + 4096;
        mbar_expect_tx(&full_bar[stage], 8192 + 2 * a.nmma * W7_ROWB + (a.has_ext ? 4096 + a.nmma * W7_XROWB : 0));
        const int row0 = b * a.seq;
        tma_load_2d(sQ, &tm_q, &full_bar[stage], h * W7_HD, row0 + t * a.q_rows);
        tma_load_2d(sK, &tm_kv0, &full_bar[stage], C + h * W7_HD, row0);
        tma_load_2d(sV, &tm_kv0, &full_bar[stage], 2 * C + h * W7_HD, row0);
        if (n1 > 0) {
          tma_load_2d(sK + a.n0 * W7_ROWB, &tm_kv1, &full_bar[stage], C + h * W7_HD, row0 + a.n0);
          tma_load_2d(sV + a.n0 * W7_ROWB, &tm_kv1, &full_bar[stage], 2 * C + h * W7_HD, row0 + a.n0);
        }
        if (a.has_ext) {
          const int xrow0 = (b % a.nwin) * a.seq;
          tma_load_2d(sQx, &tm_qx, &full_bar[stage], 0, xrow0 + t * a.q_rows);
          tma_load_2d(sKx, &tm_kx0, &full_bar[stage], 0, xrow0);
          if (n1 > 0) tma_load_2d(sKx + a.n0 * W7_XROWB, &tm_kx1, &full_bar[stage], 0, xrow0 + a.n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_pv = make_idesc_bf16(128, W7_HD, 0, 1);
      const uint32_t idesc_s0 = make_idesc_bf16(128, a.n0, 0, 0);
      const uint32_t idesc_s1 = make_idesc_bf16(128, n1 > 0 ? n1 : 16, 0, 0);
      uint32_t it = 0;
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int stage = it & 1;
        mbar_wait(&full_bar[stage], (it >> 1) & 1);
        // early: the score MMAs of this tile are issued straight behind the P V products of the previous one (the tensor pipe runs
        // them in order, so P has been consumed before S overwrites it): the softmax warps find the scores ready when they return
        // from the previous tile's O read-out.  Only the P V products below, which overwrite O, wait for that read-out (s_free).
        if (!a.early) mbar_wait(s_free, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t k_addr = q_addr + 8192;
        const uint32_t v_addr = k_addr + a.kb_bytes;
        const uint32_t qx_addr = v_addr + a.kb_bytes;
        const uint32_t kx_addr = qx_addr + 4096;
#pragma unroll
        for (int k = 0; k < W7_HD / 16; ++k)
          umma_bf16_ss(tmem_base, make_smem_desc(q_addr + k * 32, 16, 512, 4), make_smem_desc(k_addr + k * 32, 16, 512, 4), idesc_s0, k > 0);
        if (a.has_ext)
          umma_bf16_ss(tmem_base, make_smem_desc(qx_addr, 16, 256, 6), make_smem_desc(kx_addr, 16, 256, 6), idesc_s0, 1);
        if (n1 > 0) {
#pragma unroll
          for (int k = 0; k < W7_HD / 16; ++k)
            umma_bf16_ss(tmem_base + a.n0, make_smem_desc(q_addr + k * 32, 16, 512, 4),
                         make_smem_desc(k_addr + a.n0 * W7_ROWB + k * 32, 16, 512, 4), idesc_s1, k > 0);
          if (a.has_ext)
            umma_bf16_ss(tmem_base + a.n0, make_smem_desc(qx_addr, 16, 256, 6),
                         make_smem_desc(kx_addr + a.n0 * W7_XROWB, 16, 256, 6), idesc_s1, 1);
        }
        umma_commit(s_full);
        const uint32_t tmem_o = tmem_base + a.col_o;
        int kk = 0;
        if (a.split_body > 0) {
          // the packed probabilities of the first key bodies are complete: their P V products run on the tensor pipe while the
          // softmax warps exponentiate the remaining bodies (whole 16-key steps only; the straddling step waits for p_ready)
          mbar_wait(p_half, it & 1);
          if (a.early) mbar_wait(s_free, (it & 1) ^ 1);
          tc_fence_after();
          const int k_half = a.split_body * W7_TILE / 16;
          for (; kk < k_half; ++kk)
            umma_bf16_ts(tmem_o, tmem_base + kk * 8, make_smem_desc(v_addr + kk * 1024, 16, 512, 4), idesc_pv, kk > 0);
          mbar_wait(p_ready, it & 1);
        } else {
          mbar_wait(p_ready, it & 1);
          if (a.early) mbar_wait(s_free, (it & 1) ^ 1);    // the previous tile's O has been read out of tensor memory
        }
        tc_fence_after();
        for (; kk < a.nmma / 16; ++kk)
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
    const int n_body = a.seq / W7_TILE;
    int cur_h = -1;
    uint32_t it = 0;
    long long ph[5] = {0, 0, 0, 0, 0};
    long long tk = 0;
    W7Idx ix; ix.init(u_begin, a.n_qt, a.batch);
    for (long long u = u_begin; u < u_end; ++u, ++it, ix.next(a.n_qt, a.batch)) {
      const int t = ix.t, b = ix.b, h = ix.h;
      // rows of this query tile (the last one may be ragged); a warp whose 32 lanes are all beyond it skips the tile.  Lanes
      // beyond the tile inside an active warp mirror the warp's first row: same address as lane 0 in every bias load (a
      // broadcast, no second wavefront -- mirroring row 0 of the tile put them on lane 0's bank with another address)
      const int rows_t = min(a.q_rows, a.seq - t * a.q_rows);
      const bool valid = r < rows_t;
      const bool warp_active = quarter * 32 < rows_t;
      const int i = t * a.q_rows + (valid ? r : (warp_active ? quarter * 32 : 0));
      if (h != cur_h) {                       // stage this head's bias column, pre-multiplied by log2 e
        named_bar_sync(1, 128);
        const float4* src = reinterpret_cast<const float4*>(a.table_t + (long long)h * a.table_ld);
        for (int x = tid; x < a.table_ld / 4; x += 128) reinterpret_cast<float4*>(sTable)[x] = __ldg(src + x);
        cur_h = h;
        named_bar_sync(1, 128);
      }
      // sTable[(code_i + off) - code_j]: per-thread base, static column offsets
      const float* tb0 = sTable + ((i / 49) * W7_PH + ((i % 49) / 7) * W7_PW + (i % 7) + a.code_off);

      if constexpr (DBG) tk = clock64();
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      if constexpr (DBG) { const long long c = clock64(); ph[0] += c - tk; tk = c; }
      float m = -1.0e30f, l = 0.f;
      if (warp_active) {
        // ---- pass 1: x = (s + bias) * log2 e written back; row max
        const float* tb = tb0;
        float m4[4] = {-1.0e30f, -1.0e30f, -1.0e30f, -1.0e30f};
#pragma unroll 1
        for (int body = 0; body < n_body; ++body, tb -= 2 * W7_PH) {
          const uint32_t tc = taddr + body * W7_TILE;
          uint32_t va[32], vb[32], vd[2];
          tmem_ld_32x32(tc, va); tmem_ld_wait();
          tmem_ld_32x32(tc + 32, vb);
          w7_fwd_p1<0, 32>(va, tb, m4); tmem_st_32x32(tc, va);
          tmem_ld_wait();
          tmem_ld_32x32(tc + 64, va);
          w7_fwd_p1<32, 32>(vb, tb, m4); tmem_st_32x32(tc + 32, vb);
          tmem_ld_wait();
          tmem_ld_32x2(tc + 96, vd);
          w7_fwd_p1<64, 32>(va, tb, m4); tmem_st_32x32(tc + 64, va);
          tmem_ld_wait();
          w7_fwd_p1<96, 2>(vd, tb, m4); tmem_st_32x2(tc + 96, vd[0], vd[1]);
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        tmem_st_wait();
        if constexpr (DBG) { const long long c = clock64(); ph[1] += c - tk; tk = c; }
        // ---- pass 2: p = 2^(x - m); packed bf16 P written in place (columns [0, seq/2))
        float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
        for (int body = 0; body < n_body; ++body) {
          const uint32_t tc = taddr + body * W7_TILE, tp = taddr + body * (W7_TILE / 2);
          uint32_t va[32], vb[32], vd[2], pk[16];
          tmem_ld_32x32(tc, va); tmem_ld_wait();
          tmem_ld_32x32(tc + 32, vb);
          w7_fwd_p2<32>(va, m, l0, l1, pk); tmem_st_32x16(tp, pk);
          tmem_ld_wait();
          tmem_ld_32x32(tc + 64, va);
          w7_fwd_p2<32>(vb, m, l0, l1, pk); tmem_st_32x16(tp + 16, pk);
          tmem_ld_wait();
          tmem_ld_32x2(tc + 96, vd);
          w7_fwd_p2<32>(va, m, l0, l1, pk); tmem_st_32x16(tp + 32, pk);
          tmem_ld_wait();
          w7_fwd_p2<2>(vd, m, l0, l1, pk); tmem_st_32x1(tp + 48, pk[0]);
          if (body + 1 == a.split_body) {       // first key bodies packed: release their P V products
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_half);
          }
        }
        l = l0 + l1;
        // keys [seq, nmma) of the P V product: zero probabilities
        const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tmem_st_32x8(taddr + a.seq / 2, z);
        tmem_st_wait();
      }
      if (!warp_active && a.split_body > 0 && lane == 0) mbar_arrive(p_half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      if constexpr (DBG) { const long long c = clock64(); ph[2] += c - tk; tk = c; }

      // ---- epilogue: O / l -> bf16 -> global; lse (natural log)
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      if constexpr (DBG) { const long long c = clock64(); ph[3] += c - tk; tk = c; }
      if (warp_active) {
        uint32_t o[32];
        tmem_ld_32x32(taddr + a.col_o, o);
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.0f / l;
          uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)b * a.seq + i) * C + h * W7_HD);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(pack_bf16(__uint_as_float(o[q * 8]) * inv, __uint_as_float(o[q * 8 + 1]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv));
          a.lse[((long long)b * a.heads + h) * a.seq + i] = (m + log2f(l)) * W7_LN2;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if constexpr (DBG) { const long long c = clock64(); ph[4] += c - tk; tk = c; }
    }
    if constexpr (DBG) {
      if (quarter == 0 && lane == 0 && a.dbg) {
        for (int k = 0; k < 5; ++k) a.dbg[(long long)blockIdx.x * 8 + k] = ph[k];
        a.dbg[(long long)blockIdx.x * 8 + 5] = (long long)it;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, a.tmem_cols);
}

// =================================================================================================================
// Forward, second generation.  Same math and operands as attn_w7_fwd_kernel; what changes is how the softmax is spread
// over the SM (ncu r01c: 12 warps / SM, 19 % warps active, issue slots 42 % busy, MUFU 35 % -- latency-bound):
//   * EIGHT softmax warps per CTA instead of four: warps q and q + 4 share TMEM lane quarter q and split the key columns
//     of a query row at SPLIT (a multiple of 16): half 0 owns [0, SPLIT), half 1 owns [SPLIT, SEQ).  The halves exchange
//     their partial row max (before the exponentials) and row sum (before the epilogue) through shared memory and a
//     64-thread named barrier; each half packs its P in place at the START of its own column range, so no half ever writes
//     a column the other one still reads, and the P V product takes its A operand from two TMEM segments;
//   * query tiles are cut at multiples of 32 rows (196 = 96 + 100, 392 = 96 + 96 + 96 + 104) instead of 98-row slabs
//     whose fourth warp carried two rows: 7 instead of 8 warp-passes per 196-token window (13 instead of 16 at 392);
//   * 320 threads at <= 102 registers keep two CTAs (20 warps) resident per SM.
// =================================================================================================================
constexpr int W7_FWD2_THREADS = 320;

template <int CNT>
CLV_DEVICE void w7_tmem_ld(uint32_t taddr, uint32_t (&r)[CNT]) {
  if constexpr (CNT == 32) tmem_ld_32x32(taddr, r);
  else if constexpr (CNT == 16) tmem_ld_32x16(taddr, r);
  else if constexpr (CNT == 8)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  else if constexpr (CNT == 4) tmem_ld_32x4(taddr, r);
  else tmem_ld_32x2(taddr, r);
}
template <int CNT>
CLV_DEVICE void w7_tmem_st(uint32_t taddr, const uint32_t (&r)[CNT]) {
  if constexpr (CNT == 32) tmem_st_32x32(taddr, r);
  else if constexpr (CNT == 16) tmem_st_32x16(taddr, r);
  else if constexpr (CNT == 8) tmem_st_32x8(taddr, r);
  else if constexpr (CNT == 4)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  else if constexpr (CNT == 2) tmem_st_32x2(taddr, r[0], r[1]);
  else tmem_st_32x1(taddr, r[0]);
}
__host__ __device__ constexpr int w7_pow2_chunk(int rem) { return rem >= 32 ? 32 : (rem >= 16 ? 16 : (rem >= 8 ? 8 : (rem >= 4 ? 4 : 2))); }
// table offset of key column c of the window (two 49-token slabs per 98 columns)
__host__ __device__ constexpr int w7_col_code(int c) { return w7_code(c % 98) + (c / 98) * 2 * W7_PH; }

// pass 1 on the static column range [C, END): x = s * log2 e + bias * log2 e written back; running row max
template <int C, int END>
CLV_DEVICE void w7f2_pass1(uint32_t taddr, const float* tb, float& m) {
  if constexpr (C < END) {
    constexpr int CNT = w7_pow2_chunk(END - C);
    uint32_t v[CNT];
    w7_tmem_ld<CNT>(taddr + C, v);
    tmem_ld_wait();
#pragma unroll
    for (int x = 0; x < CNT; ++x) v[x] = __float_as_uint(fmaf(__uint_as_float(v[x]), W7_LOG2E, tb[-w7_col_code(C + x)]));
#pragma unroll
    for (int x = 0; x + 1 < CNT; x += 2) m = w7_max3(m, __uint_as_float(v[x]), __uint_as_float(v[x + 1]));
    w7_tmem_st<CNT>(taddr + C, v);
    w7f2_pass1<C + CNT, END>(taddr, tb, m);
  }
}
// pass 2 on [C, END): p = 2^(x - m), partial row sums, packed bf16 pairs stored at column PB + (C - LO) / 2
template <int C, int END, int LO, int PB>
CLV_DEVICE void w7f2_pass2(uint32_t taddr, float m, float& l0, float& l1) {
  if constexpr (C < END) {
    constexpr int CNT = w7_pow2_chunk(END - C);
    uint32_t v[CNT], pk[CNT / 2];
    w7_tmem_ld<CNT>(taddr + C, v);
    tmem_ld_wait();
#pragma unroll
    for (int x = 0; x < CNT; x += 2) {
      const float p0 = w7_ex2(__uint_as_float(v[x]) - m), p1 = w7_ex2(__uint_as_float(v[x + 1]) - m);
      l0 += p0; l1 += p1;
      pk[x >> 1] = pack_bf16(p0, p1);
    }
    w7_tmem_st<CNT / 2>(taddr + PB + (C - LO) / 2, pk);
    w7f2_pass2<C + CNT, END, LO, PB>(taddr, m, l0, l1);
  }
}

template <int SEQ>
__global__ void __launch_bounds__(W7_FWD2_THREADS, 2)
attn_w7_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv0,
                    const __grid_constant__ CUtensorMap tm_kv1, const __grid_constant__ CUtensorMap tm_qx,
                    const __grid_constant__ CUtensorMap tm_kx0, const __grid_constant__ CUtensorMap tm_kx1, W7FwdArgs a) {
  constexpr int NMMA = (SEQ + 15) / 16 * 16;
  constexpr int SPLIT = (SEQ / 2) / 16 * 16;          // 48, 96, 144, 192
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = 8192 + 2 * a.kb_bytes + (a.has_ext ? 4096 + a.kx_bytes : 0);
  float* sTable = reinterpret_cast<float*>(smem + 2 * stage_bytes);
  float* sMax = sTable + a.table_ld;                  // [2][128] partial row max of the two column halves
  float* sSum = sMax + 256;                           // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 256);
  uint64_t* full_bar = bars;          // [2]
  uint64_t* empty_bar = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * W7_HD;
  const int n1 = a.nmma - a.n0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv0);
    if (n1 > 0) tma_prefetch_desc(&tm_kv1);
    if (a.has_ext) { tma_prefetch_desc(&tm_qx); tma_prefetch_desc(&tm_kx0); }
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 8); mbar_init(o_full, 1); mbar_init(s_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long u_begin, u_end;
  w7_unit_range(a.units, u_begin, u_end);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      W7Idx ix; ix.init(u_begin, a.n_qt, a.batch);
      for (long long u = u_begin; u < u_end; ++u, ++it, ix.next(a.n_qt, a.batch)) {
        const int t = ix.t, b = ix.b, h = ix.h;
        const int stage = it & 1;
        mbar_wait(&empty_bar[stage], ((it >> 1) & 1) ^ 1);
        uint8_t* sQ = smem + stage * stage_bytes;
        uint8_t* sK = sQ + 8192;
        uint8_t* sV = sK + a.kb_bytes;
        uint8_t* sQx = sV + a.kb_bytes;
        uint8_t* sKx = sQx + 4096;
        mbar_expect_tx(&full_bar[stage], 8192 + 2 * a.nmma * W7_ROWB + (a.has_ext ? 4096 + a.nmma * W7_XROWB : 0));
        const int row0 = b * SEQ;
        tma_load_2d(sQ, &tm_q, &full_bar[stage], h * W7_HD, row0 + t * a.tile_rows);
        tma_load_2d(sK, &tm_kv0, &full_bar[stage], C + h * W7_HD, row0);
        tma_load_2d(sV, &tm_kv0, &full_bar[stage], 2 * C + h * W7_HD, row0);
        if (n1 > 0) {
          tma_load_2d(sK + a.n0 * W7_ROWB, &tm_kv1, &full_bar[stage], C + h * W7_HD, row0 + a.n0);
          tma_load_2d(sV + a.n0 * W7_ROWB, &tm_kv1, &full_bar[stage], 2 * C + h * W7_HD, row0 + a.n0);
        }
        if (a.has_ext) {
          const int xrow0 = (b % a.nwin) * SEQ;
          tma_load_2d(sQx, &tm_qx, &full_bar[stage], 0, xrow0 + t * a.tile_rows);
          tma_load_2d(sKx, &tm_kx0, &full_bar[stage], 0, xrow0);
          if (n1 > 0) tma_load_2d(sKx + a.n0 * W7_XROWB, &tm_kx1, &full_bar[stage], 0, xrow0 + a.n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_pv = make_idesc_bf16(128, W7_HD, 0, 1);
      const uint32_t idesc_s0 = make_idesc_bf16(128, a.n0, 0, 0);
      const uint32_t idesc_s1 = make_idesc_bf16(128, n1 > 0 ? n1 : 16, 0, 0);
      uint32_t it = 0;
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int stage = it & 1;
        mbar_wait(&full_bar[stage], (it >> 1) & 1);
        mbar_wait(s_free, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t k_addr = q_addr + 8192;
        const uint32_t v_addr = k_addr + a.kb_bytes;
        const uint32_t qx_addr = v_addr + a.kb_bytes;
        const uint32_t kx_addr = qx_addr + 4096;
#pragma unroll
        for (int k = 0; k < W7_HD / 16; ++k)
          umma_bf16_ss(tmem_base, make_smem_desc(q_addr + k * 32, 16, 512, 4), make_smem_desc(k_addr + k * 32, 16, 512, 4), idesc_s0, k > 0);
        if (a.has_ext)
          umma_bf16_ss(tmem_base, make_smem_desc(qx_addr, 16, 256, 6), make_smem_desc(kx_addr, 16, 256, 6), idesc_s0, 1);
        if (n1 > 0) {
#pragma unroll
          for (int k = 0; k < W7_HD / 16; ++k)
            umma_bf16_ss(tmem_base + a.n0, make_smem_desc(q_addr + k * 32, 16, 512, 4),
                         make_smem_desc(k_addr + a.n0 * W7_ROWB + k * 32, 16, 512, 4), idesc_s1, k > 0);
          if (a.has_ext)
            umma_bf16_ss(tmem_base + a.n0, make_smem_desc(qx_addr, 16, 256, 6),
                         make_smem_desc(kx_addr + a.n0 * W7_XROWB, 16, 256, 6), idesc_s1, 1);
        }
        umma_commit(s_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        const uint32_t tmem_o = tmem_base + a.col_o;
#pragma unroll
        for (int kk = 0; kk < NMMA / 16; ++kk) {      // P of keys [16 kk, 16 kk + 16): first or second packed segment
          const uint32_t pc = kk * 16 < SPLIT ? kk * 8 : SPLIT + ((kk * 16 - SPLIT) >> 1);
          umma_bf16_ts(tmem_o, tmem_base + pc, make_smem_desc(v_addr + kk * 1024, 16, 512, 4), idesc_pv, kk > 0);
        }
        umma_commit(o_full);
        umma_commit(&empty_bar[stage]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int cur_h = -1;
    uint32_t it = 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int t = (int)(u % a.n_qt);
      const long long bh = u / a.n_qt;
      const int b = (int)(bh % a.batch), h = (int)(bh / a.batch);
      const int rows_here = (t == a.n_qt - 1) ? SEQ - t * a.tile_rows : a.tile_rows;
      const bool valid = r < rows_here;
      const bool warp_active = quarter * 32 < rows_here;
      const int i = t * a.tile_rows + (valid ? r : 0);
      if (h != cur_h) {                       // stage this head's bias column, pre-multiplied by log2 e
        named_bar_sync(1, 256);
        const float4* src = reinterpret_cast<const float4*>(a.table_t + (long long)h * a.table_ld);
        for (int x = tid; x < a.table_ld / 4; x += 256) reinterpret_cast<float4*>(sTable)[x] = __ldg(src + x);
        cur_h = h;
        named_bar_sync(1, 256);
      }
      const float* tb = sTable + ((i / 49) * W7_PH + ((i % 49) / 7) * W7_PW + (i % 7) + a.code_off);

      mbar_wait(s_full, it & 1);
      tc_fence_after();
      float m = -1.0e30f, l = 0.f;
      if (warp_active) {
        if (half == 0) w7f2_pass1<0, SPLIT>(taddr, tb, m);
        else w7f2_pass1<SPLIT, SEQ>(taddr, tb, m);
        tmem_st_wait();
        sMax[half * 128 + r] = m;
        named_bar_sync(2 + quarter, 64);
        m = fmaxf(m, sMax[(half ^ 1) * 128 + r]);
        float l0 = 0.f, l1 = 0.f;
        if (half == 0) {
          w7f2_pass2<0, SPLIT, 0, 0>(taddr, m, l0, l1);
        } else {
          w7f2_pass2<SPLIT, SEQ, SPLIT, SPLIT>(taddr, m, l0, l1);
          // keys [SEQ, NMMA) of the P V product: zero probabilities (8 packed columns cover the <= 7 pad columns)
          const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          tmem_st_32x8(taddr + SPLIT + (SEQ - SPLIT) / 2, z);
        }
        tmem_st_wait();
        sSum[half * 128 + r] = l0 + l1;
        named_bar_sync(2 + quarter, 64);
        l = (l0 + l1) + sSum[(half ^ 1) * 128 + r];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);

      // ---- epilogue: each half normalises and stores 16 of the 32 output columns; half 0 also writes lse (natural log)
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      if (warp_active) {
        uint32_t o[16];
        tmem_ld_32x16(taddr + a.col_o + half * 16, o);
        tmem_ld_wait();
        if (valid) {
          const float inv = 1.0f / l;
          uint4* dst = reinterpret_cast<uint4*>(a.out + ((long long)b * SEQ + i) * C + h * W7_HD + half * 16);
#pragma unroll
          for (int q = 0; q < 2; ++q)
            dst[q] = make_uint4(pack_bf16(__uint_as_float(o[q * 8]) * inv, __uint_as_float(o[q * 8 + 1]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv),
                                pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv));
          if (half == 0) a.lse[((long long)b * a.heads + h) * SEQ + i] = (m + log2f(l)) * W7_LN2;
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
// Backward (seq = 98 or 196).  Unit = (window b, head h); key tiles t of 98 rows.
//   UMMA 1 : S^T = [K_t | kx] [Q | e]^T,  dP^T = [V_t | vx] [dO | e]^T      (M = 128 keys, N = NQ queries, K = 48)
//            e_i  = (-lse_hi, -lse_lo, -D_hi, -D_lo, 10*onehot(region_i) x9, 1, 1, 0)      (prep kernel, per b,h,i)
//            kx_j = (1, 1, 0, 0, 10*onehot(region_j) x9, -256, -44, 0),  vx = (0, 0, 1, 1, 0, ...)
//            so S^T arrives as s + mask - lse and dP^T as dp - D.
//   warps  : key row per thread; p = 2^(s' log2e + bias log2e), ds = p * dp'; P^T / dS^T packed in place (two column
//            segments, one per warp group), dS^T also to shared memory (A of the dQ product) and to global (bf16)
//            for the bias-table gradient.
//   UMMA 2 : dV_t = P^T dO, dK_t = dS^T Q (A from TMEM), dQ += dS K_t (A from smem, accumulated over key tiles).
// =================================================================================================================
struct W7BwdArgs {
  int batch, heads, seq, n_kt;
  int nq;                        // MMA N / K extent over queries: seq rounded up to 16 (208 or 112)
  int qb_bytes, eb_bytes;        // bytes per Q (or dO) buffer / per e buffer
  int n_mq;                      // 128-row query tiles of the dQ accumulator
  int col_dp, col_dv, col_dk, col_dq, tmem_cols;
  long long units;
  __nv_bfloat16* dqkv; float q_scale;
  int dump_ds;                   // write dS^T [batch, heads, seq, nq] (bf16) with TMA stores for the bias-table gradient
  const float* table_t; int table_len, table_ld; int code_off;
  int has_kx, nwin;
  int pipe;                      // issue dV / dK steps chunk by chunk while the softmax warps are still working
  __nv_bfloat16* dkv_part;       // bwd2 with 392 keys: dK | dV partial sums of the second query half, bf16 [rows, 2 C]
  long long ds_half_rows;        // bwd2 with 392 keys: row offset of the second query half in the dS^T dump
  int early;                     // bwd2: score MMAs of tile g + 1 are issued before the dQ products of tile g
  int l2_hint;                   // bwd2, dump_ds == 2: streamed operands evict-first, the accumulation buffers evict-last
  int ds_spans;                  // dump_ds == 2 (bwd2): dS^T tiles are ADDED (TMA reduce, bf16) into per-CTA buffers
                                 // [gridDim.x][ds_spans heads][query halves][keys][nq] that stay in L2; ds_spans = max number of
                                 // heads one CTA's contiguous unit range touches
};

constexpr int W7_BWD_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2-9: two column groups x four lane quarters
constexpr int W7_SPLIT = 96;           // queries [0, 96) -> warp group 0, [96, seq) -> warp group 1

// D_i = <dO_i, O_i>;  e rows for the K-extension (see above).  One thread per (row, head).
__global__ void attn_w7_prep_kernel(const __nv_bfloat16* out, const __nv_bfloat16* dout, const float* lse,
                                    const __nv_bfloat16* q_ext, int nwin, __nv_bfloat16* e, long long rows, int heads, int seq) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * heads) return;
  const long long row = idx / heads; const int h = (int)(idx % heads);
  const uint4* po = reinterpret_cast<const uint4*>(out + row * heads * W7_HD + h * W7_HD);
  const uint4* pd = reinterpret_cast<const uint4*>(dout + row * heads * W7_HD + h * W7_HD);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < W7_HD / 8; ++c) {
    const uint4 u = po[c], v = pd[c];
    float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
    float2 b0 = unpack_bf16(v.x), b1 = unpack_bf16(v.y), b2 = unpack_bf16(v.z), b3 = unpack_bf16(v.w);
    s += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
  }
  const long long b = row / seq; const int i = (int)(row % seq);
  const long long er = (b * heads + h) * seq + i;
  const float nl = -lse[er], nd = -s;
  const __nv_bfloat16 lh = __float2bfloat16_rn(nl), dh = __float2bfloat16_rn(nd);
  const __nv_bfloat16 ll = __float2bfloat16_rn(nl - __bfloat162float(lh)), dl = __float2bfloat16_rn(nd - __bfloat162float(dh));
  uint4 lo, hi;
  if (q_ext) {
    const uint4* src = reinterpret_cast<const uint4*>(q_ext + ((long long)(b % nwin) * seq + i) * 16);
    lo = src[0]; hi = src[1];
  } else {
    lo = make_uint4(0, 0, 0, 0); hi = make_uint4(0, 0, 0, 0);
  }
  __nv_bfloat162 p0(lh, ll), p1(dh, dl);
  lo.x = *reinterpret_cast<uint32_t*>(&p0);
  lo.y = *reinterpret_cast<uint32_t*>(&p1);
  uint4* dst = reinterpret_cast<uint4*>(e + er * 16);
  dst[0] = lo; dst[1] = hi;
}

// One 32-column chunk of the backward element loop.  C0 = static query column of v[0] (also its dS^T smem position).
template <int C0, int CNT, int NREG>
CLV_DEVICE void w7_bwd_chunk(const uint32_t (&v)[NREG], const uint32_t (&w)[NREG], const float* tbj, uint32_t* pk, uint32_t* dk) {
#pragma unroll
  for (int x = 0; x < CNT; x += 2) {
    const float p0 = w7_ex2(fmaf(__uint_as_float(v[x]), W7_LOG2E, tbj[w7_code((C0 + x) % 98) + ((C0 + x) / 98) * 2 * W7_PH]));
    const float p1 = w7_ex2(fmaf(__uint_as_float(v[x + 1]), W7_LOG2E, tbj[w7_code((C0 + x + 1) % 98) + ((C0 + x + 1) / 98) * 2 * W7_PH]));
    pk[x >> 1] = pack_bf16(p0, p1);
    dk[x >> 1] = pack_bf16(p0 * __uint_as_float(w[x]), p1 * __uint_as_float(w[x + 1]));
  }
}

template <int SEQ>
__global__ void __launch_bounds__(W7_BWD_THREADS, 1)
attn_w7_bwd_kernel(const __grid_constant__ CUtensorMap tm_q_full, const __grid_constant__ CUtensorMap tm_kv_tile,
                   const __grid_constant__ CUtensorMap tm_do_full, const __grid_constant__ CUtensorMap tm_e,
                   const __grid_constant__ CUtensorMap tm_kx, const __grid_constant__ CUtensorMap tm_ds, W7BwdArgs a) {
  constexpr int NQ = (SEQ + 15) / 16 * 16;
  constexpr int SPLIT = SEQ > W7_SPLIT ? W7_SPLIT : SEQ;     // SEQ == 98: group 0 takes [0, 96), group 1 the last two columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: [Q0 dO0 e0 | Q1 dO1 e1] [K0 V0 kx0 | K1 V1 kx1] [vx] [dS^T tile] table barriers
  const int unit_bytes = 2 * a.qb_bytes + a.eb_bytes;
  constexpr int tile_bytes = 2 * 8192 + 4096;
  uint8_t* sQdO = smem;
  uint8_t* sKV = sQdO + 2 * unit_bytes;
  uint8_t* sVx = sKV + 2 * tile_bytes;
  uint8_t* sDS = sVx + 4096;
  const int ds_bytes = a.n_mq * 2 * 16384;
  float* sTable = reinterpret_cast<float*>(sDS + ds_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTable + a.table_ld);
  uint64_t* qdo_full = bars;        // [2]
  uint64_t* qdo_empty = bars + 2;   // [2]
  uint64_t* kv_full = bars + 4;     // [2]
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* st_full = bars + 8;
  uint64_t* dvk_done = bars + 9;
  uint64_t* mma2_done = bars + 10;
  uint64_t* acc_free = bars + 11;
  uint64_t* dq_free = bars + 12;
  uint64_t* chunk_ready = bars + 13;   // [NCH] one per 32-query chunk of the packed P^T / dS^T operands
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13 + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * W7_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q_full); tma_prefetch_desc(&tm_kv_tile); tma_prefetch_desc(&tm_do_full); tma_prefetch_desc(&tm_e);
    if (a.has_kx) tma_prefetch_desc(&tm_kx);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
    }
    mbar_init(st_full, 1); mbar_init(dvk_done, 1); mbar_init(mma2_done, 1); mbar_init(acc_free, 8); mbar_init(dq_free, 8);
    for (int c = 0; c < 8; ++c) mbar_init(&chunk_ready[c], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, a.tmem_cols);
  if (warp >= 2) {
    const int tid = threadIdx.x - 64;
    for (int x = tid; x < ds_bytes / 16; x += 256) reinterpret_cast<uint4*>(sDS)[x] = make_uint4(0, 0, 0, 0);
    // constant key-side extensions: vx = (0,0,1,1,0...) for V dO^T; kx = (1,1,0...) when no shift-mask table is given.
    // SWIZZLE_32B K-major rows: 16-byte chunk c of row r sits at r*32 + ((c ^ (r >> 2)) & 1) * 16.
    const uint32_t one2 = 0x3F803F80u;
    for (int x = tid; x < 128; x += 256) {
      uint4* rowp = reinterpret_cast<uint4*>(sVx + x * 32);
      const int sw = (x >> 2) & 1;
      rowp[sw] = make_uint4(0, one2, 0, 0);
      rowp[sw ^ 1] = make_uint4(0, 0, 0, 0);
      if (!a.has_kx) {
        for (int s = 0; s < 2; ++s) {
          uint4* kp = reinterpret_cast<uint4*>(sKV + s * tile_bytes + 2 * 8192 + x * 32);
          kp[sw] = make_uint4(one2, 0, 0, 0);
          kp[sw ^ 1] = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long u_begin, u_end;
  w7_unit_range(a.units, u_begin, u_end);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0, tt = 0;
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int b = (int)(u % a.batch), h = (int)(u / a.batch);
        const int us = it & 1;
        mbar_wait(&qdo_empty[us], ((it >> 1) & 1) ^ 1);
        uint8_t* sQ = sQdO + us * unit_bytes;
        uint8_t* sDO = sQ + a.qb_bytes;
        uint8_t* sE = sDO + a.qb_bytes;
        mbar_expect_tx(&qdo_full[us], 2 * NQ * W7_ROWB + NQ * W7_XROWB);
        const int row0 = b * SEQ;
        tma_load_2d(sQ, &tm_q_full, &qdo_full[us], h * W7_HD, row0);
        tma_load_2d(sDO, &tm_do_full, &qdo_full[us], h * W7_HD, row0);
        tma_load_2d(sE, &tm_e, &qdo_full[us], 0, (int)(((long long)b * a.heads + h) * SEQ));
        for (int t = 0; t < a.n_kt; ++t, ++tt) {
          const int ts = tt & 1;
          mbar_wait(&kv_empty[ts], ((tt >> 1) & 1) ^ 1);
          uint8_t* sK = sKV + ts * tile_bytes;
          mbar_expect_tx(&kv_full[ts], 2 * 8192 + (a.has_kx ? 4096 : 0));
          tma_load_2d(sK, &tm_kv_tile, &kv_full[ts], C + h * W7_HD, row0 + t * W7_TILE);
          tma_load_2d(sK + 8192, &tm_kv_tile, &kv_full[ts], 2 * C + h * W7_HD, row0 + t * W7_TILE);
          if (a.has_kx) tma_load_2d(sK + 2 * 8192, &tm_kx, &kv_full[ts], 0, (b % a.nwin) * SEQ + t * W7_TILE);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_st = make_idesc_bf16(128, NQ, 0, 0);
      const uint32_t idesc_ts = make_idesc_bf16(128, W7_HD, 0, 1);
      const uint32_t idesc_dq = make_idesc_bf16(128, W7_HD, 1, 1);
      uint32_t it = 0, tt = 0;
      for (long long u = u_begin; u < u_end; ++u, ++it) {
        const int us = it & 1;
        mbar_wait(&qdo_full[us], (it >> 1) & 1);
        mbar_wait(dq_free, (it & 1) ^ 1);
        const uint32_t q_addr = smem_u32(sQdO + us * unit_bytes);
        const uint32_t do_addr = q_addr + a.qb_bytes;
        const uint32_t e_addr = do_addr + a.qb_bytes;
        const uint64_t desc_e = make_smem_desc(e_addr, 16, 256, 6);
        for (int t = 0; t < a.n_kt; ++t, ++tt) {
          const int ts = tt & 1;
          mbar_wait(&kv_full[ts], (tt >> 1) & 1);
          mbar_wait(acc_free, (tt & 1) ^ 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(sKV + ts * tile_bytes);
          const uint32_t v_addr = k_addr + 8192;
          const uint32_t kx_addr = v_addr + 8192;
#pragma unroll
          for (int k = 0; k < W7_HD / 16; ++k)
            umma_bf16_ss(tmem_base, make_smem_desc(k_addr + k * 32, 16, 512, 4), make_smem_desc(q_addr + k * 32, 16, 512, 4),
                         idesc_st, k > 0);
          umma_bf16_ss(tmem_base, make_smem_desc(kx_addr, 16, 256, 6), desc_e, idesc_st, 1);
#pragma unroll
          for (int k = 0; k < W7_HD / 16; ++k)
            umma_bf16_ss(tmem_base + a.col_dp, make_smem_desc(v_addr + k * 32, 16, 512, 4),
                         make_smem_desc(do_addr + k * 32, 16, 512, 4), idesc_st, k > 0);
          umma_bf16_ss(tmem_base + a.col_dp, make_smem_desc(smem_u32(sVx), 16, 256, 6), desc_e, idesc_st, 1);
          if (a.dump_ds) tma_store_wait_read();      // the previous tile's dS^T stores have left shared memory
          umma_commit(st_full);
          // dV_t = P^T dO ; dK_t = dS^T Q (K = queries, 16 per step), issued chunk by chunk as the softmax warps
          // publish 32 queries of packed operands: the tensor pipe works while the rest of the tile is still being
          // exponentiated.  Chunks alternate between the two warp groups (they run concurrently).
          // Accumulator columns (SEQ 196): dV in the spare columns [480,512); dK aliases dP^T columns [64,96), which are
          // dead once warp group 0 has published its last chunk (its packed operands end at column 48) -- the dK steps
          // of earlier chunks are deferred until then.
          // SEQ 98 (256 columns): both alias columns [56,88) of their region, dead after chunk 2.
          constexpr int NCH = (NQ / 16 + 1) / 2;
          constexpr bool BIG = SEQ > W7_SPLIT + 2;
          constexpr int order196[8] = {0, 3, 1, 4, 2, 5, 6, 6};
          constexpr int order98[4] = {2, 0, 1, 3};
          constexpr int DK_OPEN = BIG ? 4 : 0;          // position in the order after which dK may accumulate
          uint32_t acc_v = 0, acc_k = 0;
          if (!a.pipe) {
            for (int c = 0; c < NCH; ++c) mbar_wait(&chunk_ready[c], tt & 1);
            tc_fence_after();
          }
#pragma unroll 1
          for (int o = 0; o < NCH; ++o) {
            const int c = BIG ? order196[o & 7] : order98[o & 3];
            if (a.pipe) {
              mbar_wait(&chunk_ready[c], tt & 1);
              tc_fence_after();
            }
            // packed operands: queries [0, SPLIT) at columns [0, SPLIT/2); queries [SPLIT, NQ) at SPLIT + (q - SPLIT)/2
            for (int kk = 2 * c; kk < 2 * c + 2 && kk < NQ / 16; ++kk) {
              const uint32_t pc = kk * 16 < SPLIT ? kk * 8 : SPLIT + ((kk * 16 - SPLIT) >> 1);
              umma_bf16_ts(tmem_base + a.col_dv, tmem_base + pc, make_smem_desc(do_addr + kk * 1024, 16, 512, 4), idesc_ts, acc_v);
              acc_v = 1;
            }
            if (o >= DK_OPEN) {
              for (int oo = (o == DK_OPEN ? 0 : o); oo <= o; ++oo) {
                const int cc = BIG ? order196[oo & 7] : order98[oo & 3];
                for (int kk = 2 * cc; kk < 2 * cc + 2 && kk < NQ / 16; ++kk) {
                  const uint32_t pc = kk * 16 < SPLIT ? kk * 8 : SPLIT + ((kk * 16 - SPLIT) >> 1);
                  umma_bf16_ts(tmem_base + a.col_dk, tmem_base + a.col_dp + pc, make_smem_desc(q_addr + kk * 1024, 16, 512, 4), idesc_ts, acc_k);
                  acc_k = 1;
                }
              }
            }
          }
          umma_commit(dvk_done);
          if (a.dump_ds) {
            // dS^T tile (all chunks published and fenced) -> global [b, h, key row, query]: 64-query boxes of 98 rows,
            // same 128-byte swizzle as the shared tile; columns beyond nq are clipped by the tensor map
            const int grow = (int)((((long long)u % a.batch) * a.heads + (u / a.batch)) * SEQ) + t * W7_TILE;
            for (int q = 0; q < 2 * a.n_mq; ++q) tma_store_2d(&tm_ds, sDS + q * 16384, q * 64, grow);
            tma_store_commit();
          }
          const uint32_t ds_addr = smem_u32(sDS);
          for (int mq = 0; mq < a.n_mq; ++mq)      // dQ[mq] += dS K_t   (K = 128 keys of this tile; rows >= 98 of dS^T are zero)
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_ss(tmem_base + a.col_dq + mq * W7_HD, make_smem_desc(ds_addr + mq * 2 * 16384 + ks * 2048, 16384, 1024, 2),
                           make_smem_desc(k_addr + ks * 1024, 16, 512, 4), idesc_dq, (t > 0 || ks > 0) ? 1u : 0u);
          umma_commit(mma2_done);
          umma_commit(&kv_empty[ts]);
          if (t == a.n_kt - 1) umma_commit(&qdo_empty[us]);
        }
      }
      if (a.dump_ds) tma_store_wait_all();
    }
  } else {
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool valid = r < W7_TILE;
    const bool warp_active = quarter * 32 < W7_TILE;
    uint8_t* ds_row = sDS + (r >> 3) * 1024 + (r & 7) * 128;
    const int rsw = r & 7;
    int cur_h = -1;
    uint32_t it = 0, tt = 0;
    for (long long u = u_begin; u < u_end; ++u, ++it) {
      const int b = (int)(u % a.batch), h = (int)(u / a.batch);
      if (h != cur_h) {
        named_bar_sync(1, 256);
        const float4* src = reinterpret_cast<const float4*>(a.table_t + (long long)h * a.table_ld);
        for (int x = tid; x < a.table_ld / 4; x += 256) reinterpret_cast<float4*>(sTable)[x] = __ldg(src + x);
        cur_h = h;
        named_bar_sync(1, 256);
      }
      for (int t = 0; t < a.n_kt; ++t, ++tt) {
        const int j = t * W7_TILE + (valid ? r : quarter * 32);     // lanes beyond the tile mirror their warp's first row (broadcast bias loads)
        // sTable[code_i + (off - code_j)]: per-thread base, static query offsets
        const float* tbj = sTable + (a.code_off - ((j / 49) * W7_PH + ((j % 49) / 7) * W7_PW + (j % 7)));
        mbar_wait(st_full, tt & 1);
        tc_fence_after();
        if (warp_active) {
          const uint32_t ts0 = taddr, td0 = taddr + a.col_dp;
          if (grp == 0) {
            // queries [0, 96): three chunks of 32, packed at [0, 48)
            uint32_t v[32], w[32], v2[32], w2[32], pk[16], dk[16];
            tmem_ld_32x32(ts0, v); tmem_ld_32x32(td0, w); tmem_ld_wait();
#define W7_CHUNK_DONE(CH, LAST)                                                                                  \
            tmem_st_wait();                                                                                      \
            if (LAST) fence_proxy_async();                                                                       \
            tc_fence_before();                                                                                   \
            __syncwarp();                                                                                        \
            if (lane == 0) mbar_arrive(&chunk_ready[CH]);
#define W7_BWD_EMIT(C0, PV, PW)                                                                                  \
            w7_bwd_chunk<C0, 32>(PV, PW, tbj, pk, dk);                                                           \
            tmem_st_32x16(ts0 + (C0) / 2, pk); tmem_st_32x16(td0 + (C0) / 2, dk);                                 \
            if (valid) {                                                                                         \
              uint8_t* chunk = ds_row + ((C0) >> 6) * 16384;                                                     \
              _Pragma("unroll") for (int q = 0; q < 4; ++q)                                                      \
                *reinterpret_cast<uint4*>(chunk + (((((C0) >> 5) & 1) * 4 + q) ^ rsw) * 16) =                    \
                    make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);                          \
            }
            tmem_ld_32x32(ts0 + 32, v2); tmem_ld_32x32(td0 + 32, w2);
            W7_BWD_EMIT(0, v, w)
            W7_CHUNK_DONE(0, false)
            tmem_ld_wait();
            tmem_ld_32x32(ts0 + 64, v); tmem_ld_32x32(td0 + 64, w);
            W7_BWD_EMIT(32, v2, w2)
            W7_CHUNK_DONE(1, false)
            tmem_ld_wait();
            W7_BWD_EMIT(64, v, w)
            W7_CHUNK_DONE(2, true)
          } else if (SEQ > W7_SPLIT + 2) {
            // queries [96, 196): three chunks of 32 and a tail of 4, packed at [96, 146); zero the packed pads [146, 152)
            uint32_t v[32], w[32], v2[32], w2[32], pk[16], dk[16];
            tmem_ld_32x32(ts0 + 96, v); tmem_ld_32x32(td0 + 96, w); tmem_ld_wait();
#define W7_BWD_EMIT1(C0, PV, PW)                                                                                 \
            w7_bwd_chunk<C0, 32>(PV, PW, tbj, pk, dk);                                                           \
            tmem_st_32x16(ts0 + 96 + ((C0) - 96) / 2, pk); tmem_st_32x16(td0 + 96 + ((C0) - 96) / 2, dk);         \
            if (valid) {                                                                                         \
              uint8_t* chunk = ds_row + ((C0) >> 6) * 16384;                                                     \
              _Pragma("unroll") for (int q = 0; q < 4; ++q)                                                      \
                *reinterpret_cast<uint4*>(chunk + (((((C0) >> 5) & 1) * 4 + q) ^ rsw) * 16) =                    \
                    make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);                          \
            }
            tmem_ld_32x32(ts0 + 128, v2); tmem_ld_32x32(td0 + 128, w2);
            W7_BWD_EMIT1(96, v, w)
            W7_CHUNK_DONE(3, false)
            tmem_ld_wait();
            tmem_ld_32x32(ts0 + 160, v); tmem_ld_32x32(td0 + 160, w);
            W7_BWD_EMIT1(128, v2, w2)
            W7_CHUNK_DONE(4, false)
            tmem_ld_wait();
            uint32_t v4[4], w4[4];
            tmem_ld_32x4(ts0 + 192, v4); tmem_ld_32x4(td0 + 192, w4);
            W7_BWD_EMIT1(160, v, w)
            W7_CHUNK_DONE(5, false)
            tmem_ld_wait();
            w7_bwd_chunk<192, 4>(v4, w4, tbj, pk, dk);
            {
              const uint32_t zp[8] = {pk[0], pk[1], 0, 0, 0, 0, 0, 0}, zd[8] = {dk[0], dk[1], 0, 0, 0, 0, 0, 0};
              tmem_st_32x8(ts0 + 96 + 48, zp); tmem_st_32x8(td0 + 96 + 48, zd);     // packed columns [144, 152)
            }
            if (valid) {
              *reinterpret_cast<uint4*>(ds_row + 3 * 16384 + ((0 ^ rsw) * 16)) = make_uint4(dk[0], dk[1], 0, 0);
            }
            W7_CHUNK_DONE(6, true)
          } else {
            // SEQ == 98: group 1 takes the last two queries; packed at [96/2 .. ) does not apply -> single segment
            uint32_t v4[2], w4[2], pk[16], dk[16];
            tmem_ld_32x2(ts0 + 96, v4); tmem_ld_32x2(td0 + 96, w4); tmem_ld_wait();
            w7_bwd_chunk<96, 2>(v4, w4, tbj, pk, dk);
            {
              const uint32_t zp[8] = {pk[0], 0, 0, 0, 0, 0, 0, 0}, zd[8] = {dk[0], 0, 0, 0, 0, 0, 0, 0};
              tmem_st_32x8(ts0 + 96, zp); tmem_st_32x8(td0 + 96, zd);               // packed columns [96, 104) <- queries [96, 112)
            }
            if (valid) {
              *reinterpret_cast<uint4*>(ds_row + 1 * 16384 + (((4 + 0) ^ rsw) * 16)) = make_uint4(dk[0], 0, 0, 0);
            }
            W7_CHUNK_DONE(3, true)
          }
        }

        mbar_wait(dvk_done, tt & 1);
        tc_fence_after();
        if (warp_active) {
          // dV rows by warp group 0, dK rows by warp group 1
          uint32_t o[32];
          tmem_ld_32x32(taddr + (grp ? a.col_dk : a.col_dv), o);
          tmem_ld_wait();
          if (valid) {
            // dK picked up the log2e-free scores' gradient directly: dS^T already is d/ds
            uint4* g = reinterpret_cast<uint4*>(a.dqkv + ((long long)b * SEQ + j) * (3 * C) + (grp ? 1 : 2) * C + h * W7_HD);
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
          mbar_wait(mma2_done, tt & 1);
          tc_fence_after();
          for (int mq = grp; mq < a.n_mq; mq += 2) {
            const int i = mq * 128 + r;
            if (mq * 128 + quarter * 32 < SEQ) {
              uint32_t oq[32];
              tmem_ld_32x32(taddr + a.col_dq + mq * W7_HD, oq);
              tmem_ld_wait();
              if (i < SEQ) {
                uint4* gq = reinterpret_cast<uint4*>(a.dqkv + ((long long)b * SEQ + i) * (3 * C) + h * W7_HD);
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

// =================================================================================================================
// Backward, second generation (seq = 196): the same math as attn_w7_bwd_kernel, re-pipelined around the measured cost
// of tcgen05.mma -- about 98 cycles per instruction for any N <= 128 (tools/probes/mma_probe.cu) -- which makes the
// 48 MMAs of a key tile (~5.9 k cycles) the bound, not the exponentials (~1.6 k).  The first version serialises
// "all of S^T and dP^T -> all threads -> all of dV/dK/dQ" because the two fp32 score tiles fill tensor memory; here
// the queries are cut into chunks of 96 that cycle through two small TMEM buffers,
//     buf b: S^T chunk at [192 b, 192 b + 96), dP^T chunk at [192 b + 96, 192 b + 192);   dV [384,416)  dK [416,448)  dQ [448,512)
// so the tensor pipe computes the scores of chunk c+1 and the dV/dK steps of chunk c-1 while the threads
// exponentiate chunk c, and no accumulator aliases live data.  Chunks: queries [0,96) -> buf 0, [96,192) -> buf 1,
// [192,208) -> buf 0 (4 real queries + zero padding).  Inside a chunk warp group g owns columns [48 g, 48 g + 48)
// and packs P^T / dS^T in place at [48 g, 48 g + 24) of the S^T / dP^T halves (3 K-steps of 16 queries each).
// =================================================================================================================
constexpr int W7B2_CH = 96;            // queries per chunk
constexpr int W7B2_BUF = 192;          // TMEM columns per buffer (S^T chunk + dP^T chunk)
constexpr int W7B2_DV = 384, W7B2_DK = 416, W7B2_DQ = 448;

// dS^T of 8 consecutive queries starting at I8 (multiple of 8) -> shared tile (64-query chunks, 128-byte swizzle)
template <int I8>
CLV_DEVICE void w7_ds_store(uint8_t* ds_row, int rsw, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(ds_row + (I8 >> 6) * 16384 + ((((I8 >> 3) & 7) ^ rsw) * 16)) = make_uint4(a, b, c, d);
}

// 48 query columns [I0, I0 + 48) of one key row: TMEM (S^T at ts, dP^T at td, column 0 = query I0) -> P^T / dS^T packed in place
// (operands of the dV / dK steps); the dS^T values are returned in dk / dk2 for the shared tile the dQ products read.
template <int I0>
CLV_DEVICE void w7_bwd2_compute(uint32_t ts, uint32_t td, const float* tbj, uint32_t (&dk)[16], uint32_t (&dk2)[8]) {
  uint32_t v[32], w[32], v2[16], w2[16], pk[16];
  tmem_ld_32x32(ts, v); tmem_ld_32x32(td, w); tmem_ld_wait();
  tmem_ld_32x16(ts + 32, v2); tmem_ld_32x16(td + 32, w2);
  w7_bwd_chunk<I0, 32>(v, w, tbj, pk, dk);
  tmem_st_32x16(ts, pk); tmem_st_32x16(td, dk);
  tmem_ld_wait();
  uint32_t pk2[8];
  w7_bwd_chunk<I0 + 32, 16>(v2, w2, tbj, pk2, dk2);
  tmem_st_32x8(ts + 16, pk2); tmem_st_32x8(td + 16, dk2);
}
template <int I0>
CLV_DEVICE void w7_bwd2_store(uint8_t* ds_row, int rsw, const uint32_t (&dk)[16], const uint32_t (&dk2)[8]) {
  w7_ds_store<I0>(ds_row, rsw, dk[0], dk[1], dk[2], dk[3]);
  w7_ds_store<I0 + 8>(ds_row, rsw, dk[4], dk[5], dk[6], dk[7]);
  w7_ds_store<I0 + 16>(ds_row, rsw, dk[8], dk[9], dk[10], dk[11]);
  w7_ds_store<I0 + 24>(ds_row, rsw, dk[12], dk[13], dk[14], dk[15]);
  w7_ds_store<I0 + 32>(ds_row, rsw, dk2[0], dk2[1], dk2[2], dk2[3]);
  w7_ds_store<I0 + 40>(ds_row, rsw, dk2[4], dk2[5], dk2[6], dk2[7]);
}
template <int I0>
CLV_DEVICE void w7_bwd2_body(uint32_t ts, uint32_t td, const float* tbj, uint8_t* ds_row, int rsw, bool valid) {
  uint32_t dk[16], dk2[8];
  w7_bwd2_compute<I0>(ts, td, tbj, dk, dk2);
  if (valid) w7_bwd2_store<I0>(ds_row, rsw, dk, dk2);
}

// KSEQ = 196: unit = (window, head).  KSEQ = 392 (the full (8,7,7) window of 16-frame clips and of BASELINE config c2):
// unit = (window, head, query half); a unit holds 196 queries (the TMEM budget above is unchanged) and walks all four
// key tiles.  dQ of the half is complete; dK / dV are partial sums over the half's queries: half 0 writes them to dqkv,
// half 1 to a.dkv_part, and attn_w7_dkv_combine_kernel adds the two.  The bias gather only needs the half's code
// offset (196 tokens = 4 slabs -> 4 * 169) added to the per-thread table base.
template <int KSEQ>
__global__ void __launch_bounds__(W7_BWD_THREADS, 1)
attn_w7_bwd2_kernel(const __grid_constant__ CUtensorMap tm_q_full, const __grid_constant__ CUtensorMap tm_kv_tile,
                    const __grid_constant__ CUtensorMap tm_do_full, const __grid_constant__ CUtensorMap tm_e,
                    const __grid_constant__ CUtensorMap tm_kx, const __grid_constant__ CUtensorMap tm_ds, W7BwdArgs a) {
  constexpr int SEQ = 196, NQ = 208;
  constexpr int NKT = KSEQ / W7_TILE, NH = KSEQ / SEQ;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: [Q0 dO0 e0 | Q1 dO1 e1] [K0 V0 kx0 | K1 V1 kx1] [vx] [dS^T tile] table barriers   (as attn_w7_bwd_kernel)
  const int unit_bytes = 2 * a.qb_bytes + a.eb_bytes;
  constexpr int tile_bytes = 2 * 8192 + 4096;
  uint8_t* sQdO = smem;
  uint8_t* sKV = sQdO + 2 * unit_bytes;
  uint8_t* sVx = sKV + 2 * tile_bytes;
  uint8_t* sDS = sVx + 4096;
  constexpr int ds_bytes = 4 * 16384;
  float* sTable = reinterpret_cast<float*>(sDS + ds_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTable + a.table_ld);
  uint64_t* qdo_full = bars;        // [2]
  uint64_t* qdo_empty = bars + 2;   // [2]
  uint64_t* kv_full = bars + 4;     // [2]
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* s_ready = bars + 8;     // [2] scores of the chunk in buffer b are in tensor memory
  uint64_t* p_ready = bars + 10;    // [2] packed operands of the chunk in buffer b are published (8 warps)
  uint64_t* dvk_done = bars + 12;
  uint64_t* mma2_done = bars + 13;
  uint64_t* acc_free = bars + 14;
  uint64_t* dq_free = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.heads * W7_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q_full); tma_prefetch_desc(&tm_kv_tile); tma_prefetch_desc(&tm_do_full); tma_prefetch_desc(&tm_e);
    if (a.has_kx) tma_prefetch_desc(&tm_kx);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      mbar_init(&s_ready[s], 1); mbar_init(&p_ready[s], 8);
    }
    mbar_init(dvk_done, 1); mbar_init(mma2_done, 1); mbar_init(acc_free, 8); mbar_init(dq_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp >= 2) {
    const int tid = threadIdx.x - 64;
    for (int x = tid; x < ds_bytes / 16; x += 256) reinterpret_cast<uint4*>(sDS)[x] = make_uint4(0, 0, 0, 0);
    const uint32_t one2 = 0x3F803F80u;
    for (int x = tid; x < 128; x += 256) {
      uint4* rowp = reinterpret_cast<uint4*>(sVx + x * 32);
      const int sw = (x >> 2) & 1;
      rowp[sw] = make_uint4(0, one2, 0, 0);
      rowp[sw ^ 1] = make_uint4(0, 0, 0, 0);
      if (!a.has_kx) {
        for (int s = 0; s < 2; ++s) {
          uint4* kp = reinterpret_cast<uint4*>(sKV + s * tile_bytes + 2 * 8192 + x * 32);
          kp[sw] = make_uint4(one2, 0, 0, 0);
          kp[sw ^ 1] = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long u_begin, u_end;
  w7_unit_range(a.units, u_begin, u_end);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0, tt = 0;
      const uint64_t pol_stream = l2_policy_evict_first();
      W7Idx ix; ix.init(u_begin, NH, a.batch);
      for (long long u = u_begin; u < u_end; ++u, ++it, ix.next(NH, a.batch)) {
        const int qh = ix.t, b = ix.b, h = ix.h;
        const int us = it & 1;
        mbar_wait(&qdo_empty[us], ((it >> 1) & 1) ^ 1);
        uint8_t* sQ = sQdO + us * unit_bytes;
        uint8_t* sDO = sQ + a.qb_bytes;
        uint8_t* sE = sDO + a.qb_bytes;
        mbar_expect_tx(&qdo_full[us], 2 * NQ * W7_ROWB + NQ * W7_XROWB);
        const int row0 = b * KSEQ;
        if (a.l2_hint) {
          tma_load_2d_hint(sQ, &tm_q_full, &qdo_full[us], h * W7_HD, row0 + qh * SEQ, pol_stream);
          tma_load_2d_hint(sDO, &tm_do_full, &qdo_full[us], h * W7_HD, row0 + qh * SEQ, pol_stream);
        } else {
          tma_load_2d(sQ, &tm_q_full, &qdo_full[us], h * W7_HD, row0 + qh * SEQ);
          tma_load_2d(sDO, &tm_do_full, &qdo_full[us], h * W7_HD, row0 + qh * SEQ);
        }
        tma_load_2d(sE, &tm_e, &qdo_full[us], 0, (int)(((long long)b * a.heads + h) * KSEQ + qh * SEQ));
        for (int t = 0; t < NKT; ++t, ++tt) {
          const int ts = tt & 1;
          mbar_wait(&kv_empty[ts], ((tt >> 1) & 1) ^ 1);
          uint8_t* sK = sKV + ts * tile_bytes;
          mbar_expect_tx(&kv_full[ts], 2 * 8192 + (a.has_kx ? 4096 : 0));
          if (a.l2_hint) {
            tma_load_2d_hint(sK, &tm_kv_tile, &kv_full[ts], C + h * W7_HD, row0 + t * W7_TILE, pol_stream);
            tma_load_2d_hint(sK + 8192, &tm_kv_tile, &kv_full[ts], 2 * C + h * W7_HD, row0 + t * W7_TILE, pol_stream);
          } else {
            tma_load_2d(sK, &tm_kv_tile, &kv_full[ts], C + h * W7_HD, row0 + t * W7_TILE);
            tma_load_2d(sK + 8192, &tm_kv_tile, &kv_full[ts], 2 * C + h * W7_HD, row0 + t * W7_TILE);
          }
          if (a.has_kx) tma_load_2d(sK + 2 * 8192, &tm_kx, &kv_full[ts], 0, (b % a.nwin) * KSEQ + t * W7_TILE);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_ts = make_idesc_bf16(128, W7_HD, 0, 1);
      const uint32_t idesc_dq = make_idesc_bf16(128, W7_HD, 1, 1);
      const uint32_t idesc_c96 = make_idesc_bf16(128, 96, 0, 0), idesc_c16 = make_idesc_bf16(128, 16, 0, 0);
      const uint64_t desc_vx = make_smem_desc(smem_u32(sVx), 16, 256, 6);
      uint32_t nuse0 = 0, nuse1 = 0;          // completed uses of buffer 0 / 1 (barrier phases)
      const long long G = (u_end - u_begin) * NKT;      // key tiles this CTA walks (NKT per unit)
      // scores of one chunk of tile g: S^T = [K_t | kx] [Q_c | e_c]^T and dP^T = [V_t | vx] [dO_c | e_c]^T into buffer `buf`
      auto mma1 = [&](long long g, int c, int buf) {
        const uint32_t q_addr = smem_u32(sQdO + ((g / NKT) & 1) * unit_bytes);
        const uint32_t do_addr = q_addr + a.qb_bytes, e_addr = do_addr + a.qb_bytes;
        const uint32_t k_addr = smem_u32(sKV + (g & 1) * tile_bytes), v_addr = k_addr + 8192;
        const uint64_t desc_kx = make_smem_desc(v_addr + 8192, 16, 256, 6);
        const uint32_t idesc = c < 2 ? idesc_c96 : idesc_c16;
        const uint32_t ds_ = tmem_base + buf * W7B2_BUF, dp_ = ds_ + W7B2_CH;
        const uint32_t qo = c * W7B2_CH * W7_ROWB, eo = c * W7B2_CH * W7_XROWB;
        const uint64_t desc_e = make_smem_desc(e_addr + eo, 16, 256, 6);
#pragma unroll
        for (int k = 0; k < W7_HD / 16; ++k)
          umma_bf16_ss(ds_, make_smem_desc(k_addr + k * 32, 16, 512, 4), make_smem_desc(q_addr + qo + k * 32, 16, 512, 4), idesc, k > 0);
        umma_bf16_ss(ds_, desc_kx, desc_e, idesc, 1);
#pragma unroll
        for (int k = 0; k < W7_HD / 16; ++k)
          umma_bf16_ss(dp_, make_smem_desc(v_addr + k * 32, 16, 512, 4), make_smem_desc(do_addr + qo + k * 32, 16, 512, 4), idesc, k > 0);
        umma_bf16_ss(dp_, desc_vx, desc_e, idesc, 1);
        umma_commit(&s_ready[buf]);
      };
      // operands of tile g have landed -> score MMAs of its first two chunks
      auto scores01 = [&](long long g) {
        const uint32_t it = (uint32_t)(g / NKT);
        if (g % NKT == 0) mbar_wait(&qdo_full[it & 1], (it >> 1) & 1);
        mbar_wait(&kv_full[g & 1], (uint32_t)(g >> 1) & 1);
        tc_fence_after();
        mma1(g, 0, 0);
        mma1(g, 1, 1);
      };
      // Tile order.  early == 0: scores(g), dV/dK(g), dQ(g).  early == 1: the score MMAs of tile g + 1 are issued BEFORE the dQ
      // products of tile g, so the softmax warps exponentiate the first chunk of the next tile while the tensor pipe works
      // through the 16 dQ instructions (they used to idle there, and the pipe idled at the start of every tile).
      if (a.early && G > 0) scores01(0);
      W7Idx mx; mx.init(u_begin, NH, a.batch);        // (query half, window, head) of the unit of tile g
      const int h_first = mx.h;
      for (long long g = 0; g < G; ++g) {
        const uint32_t it = (uint32_t)(g / NKT);
        const int t = (int)(g % NKT);
        const int us = it & 1, ts = (int)(g & 1);
        const uint32_t q_addr = smem_u32(sQdO + us * unit_bytes);
        const uint32_t do_addr = q_addr + a.qb_bytes;
        const uint32_t k_addr = smem_u32(sKV + ts * tile_bytes);
        if (!a.early) scores01(g);
        // dV_t += P^T_c dO_c ; dK_t += dS^T_c Q_c : K-steps of 16 queries; packed operands of warp group g at [48 g, 48 g + 24)
        auto mma_dvk = [&](int c, int buf, uint32_t& acc) {
          const uint32_t ps = tmem_base + buf * W7B2_BUF, pd = ps + W7B2_CH;
          const int nsteps = c < 2 ? 6 : 1;
          for (int s = 0; s < nsteps; ++s) {
            const uint32_t pc = (s < 3 ? 0 : 48) + (s % 3) * 8;
            const uint32_t ro = (c * W7B2_CH + s * 16) * W7_ROWB;     // 16 query rows = 1024 bytes
            umma_bf16_ts(tmem_base + W7B2_DV, ps + pc, make_smem_desc(do_addr + ro, 16, 512, 4), idesc_ts, acc);
            umma_bf16_ts(tmem_base + W7B2_DK, pd + pc, make_smem_desc(q_addr + ro, 16, 512, 4), idesc_ts, acc);
            acc = 1;
          }
        };
        uint32_t acc = 0;
        mbar_wait(&p_ready[0], nuse0 & 1); ++nuse0;
        mbar_wait(acc_free, (uint32_t)(g & 1) ^ 1);   // previous tile's dV / dK have been read
        tc_fence_after();
        mma_dvk(0, 0, acc);
        mma1(g, 2, 0);                                // executes after the dV/dK steps that read buffer 0
        mbar_wait(&p_ready[1], nuse1 & 1); ++nuse1;
        tc_fence_after();
        mma_dvk(1, 1, acc);
        mbar_wait(&p_ready[0], nuse0 & 1); ++nuse0;
        tc_fence_after();
        mma_dvk(2, 0, acc);
        umma_commit(dvk_done);
        if (a.early && g + 1 < G) scores01(g + 1);
        if (a.dump_ds == 2) {
          const int buf = ((int)blockIdx.x * a.ds_spans + (mx.h - h_first)) * NH + mx.t;
          const int grow = buf * KSEQ + t * W7_TILE;
          if (a.l2_hint) {
            const uint64_t pol_keep = l2_policy_evict_last();
            for (int q = 0; q < 4; ++q) tma_reduce_add_2d_hint(&tm_ds, sDS + q * 16384, q * 64, grow, pol_keep);
          } else {
            for (int q = 0; q < 4; ++q) tma_reduce_add_2d(&tm_ds, sDS + q * 16384, q * 64, grow);
          }
          tma_store_commit();
        } else if (a.dump_ds) {
          const int grow = (int)(mx.t * a.ds_half_rows + ((long long)mx.b * a.heads + mx.h) * KSEQ) + t * W7_TILE;
          for (int q = 0; q < 4; ++q) tma_store_2d(&tm_ds, sDS + q * 16384, q * 64, grow);
          tma_store_commit();
        }
        if (t == 0) mbar_wait(dq_free, (it & 1) ^ 1);     // the previous unit's dQ accumulators have been read out
        const uint32_t ds_addr = smem_u32(sDS);
        for (int mq = 0; mq < 2; ++mq)           // dQ[mq] += dS K_t   (98 keys of this tile: K steps of 16 up to 112; rows >= 98 of
          for (int ks = 0; ks < 7; ++ks)         //  dS^T are zero, so the step over keys [112, 128) would add nothing)
            umma_bf16_ss(tmem_base + W7B2_DQ + mq * W7_HD, make_smem_desc(ds_addr + mq * 2 * 16384 + ks * 2048, 16384, 1024, 2),
                         make_smem_desc(k_addr + ks * 1024, 16, 512, 4), idesc_dq, (t > 0 || ks > 0) ? 1u : 0u);
        if (a.dump_ds) tma_store_wait_read();     // (the dQ products above were only issued; the stores finish reading first)
        umma_commit(mma2_done);
        umma_commit(&kv_empty[ts]);
        if (t == NKT - 1) { umma_commit(&qdo_empty[us]); mx.next(NH, a.batch); }
      }
      if (a.dump_ds) tma_store_wait_all();
    }
  } else {
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool valid = r < W7_TILE;
    uint8_t* ds_row = sDS + (r >> 3) * 1024 + (r & 7) * 128;
    const int rsw = r & 7;
    int cur_h = -1;
    uint32_t it = 0, tt = 0;
    uint32_t nuse0 = 0, nuse1 = 0;
    const uint32_t total_tiles = (uint32_t)((u_end - u_begin) * NKT);
    bool pend_dq = false;                      // early mode: dQ of the unit that ended with the previous tile still to be read out
    int pend_b = 0, pend_qh = 0, pend_h = 0;
    W7Idx ix; ix.init(u_begin, NH, a.batch);
    for (long long u = u_begin; u < u_end; ++u, ++it, ix.next(NH, a.batch)) {
      const int qh = ix.t, b = ix.b, h = ix.h;
      if (h != cur_h) {
        named_bar_sync(1, 256);
        const float4* src = reinterpret_cast<const float4*>(a.table_t + (long long)h * a.table_ld);
        for (int x = tid; x < a.table_ld / 4; x += 256) reinterpret_cast<float4*>(sTable)[x] = __ldg(src + x);
        cur_h = h;
        named_bar_sync(1, 256);
      }
      for (int t = 0; t < NKT; ++t, ++tt) {
        const int j = t * W7_TILE + (valid ? r : quarter * 32);     // lanes beyond the tile mirror their warp's first row (broadcast bias loads)
        const float* tbj = sTable + (a.code_off + qh * 4 * W7_PH - ((j / 49) * W7_PH + ((j % 49) / 7) * W7_PW + (j % 7)));
#define W7B2_PUBLISH(BUF)                                                          \
        tmem_st_wait();                                                            \
        fence_proxy_async();                                                       \
        tc_fence_before();                                                         \
        __syncwarp();                                                              \
        if (lane == 0) mbar_arrive(&p_ready[BUF]);
        // dQ of a whole unit (all key tiles accumulated) -> dqkv; query row i = grp*128 + r
        auto dq_readout = [&](int b_, int qh_, int h_) {
          tc_fence_after();
          const int i = grp * 128 + r;
          if (grp * 128 + quarter * 32 < SEQ) {
            uint32_t oq[32];
            tmem_ld_32x32(taddr + W7B2_DQ + grp * W7_HD, oq);
            tmem_ld_wait();
            if (i < SEQ) {
              uint4* gq = reinterpret_cast<uint4*>(a.dqkv + ((long long)b_ * KSEQ + qh_ * SEQ + i) * (3 * C) + h_ * W7_HD);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                gq[q] = make_uint4(pack_bf16(__uint_as_float(oq[q * 8]) * a.q_scale, __uint_as_float(oq[q * 8 + 1]) * a.q_scale),
                                   pack_bf16(__uint_as_float(oq[q * 8 + 2]) * a.q_scale, __uint_as_float(oq[q * 8 + 3]) * a.q_scale),
                                   pack_bf16(__uint_as_float(oq[q * 8 + 4]) * a.q_scale, __uint_as_float(oq[q * 8 + 5]) * a.q_scale),
                                   pack_bf16(__uint_as_float(oq[q * 8 + 6]) * a.q_scale, __uint_as_float(oq[q * 8 + 7]) * a.q_scale));
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_free);
        };
        // ---- chunk 0: queries [0, 96) in buffer 0
        mbar_wait(&s_ready[0], nuse0 & 1); ++nuse0;
        tc_fence_after();
        if (!a.early) {
          // the previous tile's dQ products (which read the shared dS^T tile) are ordered before this tile's first
          // score MMA, so s_ready also says the tile may be overwritten
          if (grp == 0) w7_bwd2_body<0>(taddr, taddr + W7B2_CH, tbj, ds_row, rsw, valid);
          else w7_bwd2_body<48>(taddr + 48, taddr + W7B2_CH + 48, tbj, ds_row, rsw, valid);
          W7B2_PUBLISH(0)
        } else {
          // early mode: these scores were issued BEFORE the previous tile's dQ products.  Exponentiate and publish the packed
          // operands first (that is what the dV / dK steps wait for), then wait until the dQ products -- and the dS^T dump --
          // of the previous tile have finished reading the shared dS^T tile before overwriting it; the dQ read-out of a unit
          // that ended with the previous tile happens here too.
          uint32_t dk[16], dk2[8];
          if (grp == 0) w7_bwd2_compute<0>(taddr, taddr + W7B2_CH, tbj, dk, dk2);
          else w7_bwd2_compute<48>(taddr + 48, taddr + W7B2_CH + 48, tbj, dk, dk2);
          W7B2_PUBLISH(0)
          if (tt > 0) mbar_wait(mma2_done, (tt - 1) & 1);
          if (pend_dq) { dq_readout(pend_b, pend_qh, pend_h); pend_dq = false; }
          if (valid) {
            if (grp == 0) w7_bwd2_store<0>(ds_row, rsw, dk, dk2);
            else w7_bwd2_store<48>(ds_row, rsw, dk, dk2);
          }
        }
        // ---- chunk 1: queries [96, 192) in buffer 1
        mbar_wait(&s_ready[1], nuse1 & 1); ++nuse1;
        tc_fence_after();
        if (grp == 0) w7_bwd2_body<96>(taddr + W7B2_BUF, taddr + W7B2_BUF + W7B2_CH, tbj, ds_row, rsw, valid);
        else w7_bwd2_body<144>(taddr + W7B2_BUF + 48, taddr + W7B2_BUF + W7B2_CH + 48, tbj, ds_row, rsw, valid);
        W7B2_PUBLISH(1)
        // ---- chunk 2: queries [192, 196) (+ zero padding up to 208) in buffer 0, warp group 0 only
        mbar_wait(&s_ready[0], nuse0 & 1); ++nuse0;
        tc_fence_after();
        if (grp == 0) {
          uint32_t v4[4], w4[4], pk[2], dk[2];
          tmem_ld_32x4(taddr, v4); tmem_ld_32x4(taddr + W7B2_CH, w4); tmem_ld_wait();
          w7_bwd_chunk<192, 4>(v4, w4, tbj, pk, dk);
          const uint32_t zp[8] = {pk[0], pk[1], 0, 0, 0, 0, 0, 0}, zd[8] = {dk[0], dk[1], 0, 0, 0, 0, 0, 0};
          tmem_st_32x8(taddr, zp); tmem_st_32x8(taddr + W7B2_CH, zd);
          if (valid) w7_ds_store<192>(ds_row, rsw, dk[0], dk[1], 0, 0);
        }
        W7B2_PUBLISH(0)

        mbar_wait(dvk_done, tt & 1);
        tc_fence_after();
        {
          // dV rows by warp group 0, dK rows by warp group 1
          uint32_t o[32];
          tmem_ld_32x32(taddr + (grp ? W7B2_DK : W7B2_DV), o);
          tmem_ld_wait();
          if (valid) {
            uint4* g = (NH > 1 && qh == 1)
                ? reinterpret_cast<uint4*>(a.dkv_part + ((long long)b * KSEQ + j) * (2 * C) + (grp ? 0 : 1) * C + h * W7_HD)
                : reinterpret_cast<uint4*>(a.dqkv + ((long long)b * KSEQ + j) * (3 * C) + (grp ? 1 : 2) * C + h * W7_HD);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              g[q] = make_uint4(pack_bf16(__uint_as_float(o[q * 8]), __uint_as_float(o[q * 8 + 1])),
                                pack_bf16(__uint_as_float(o[q * 8 + 2]), __uint_as_float(o[q * 8 + 3])),
                                pack_bf16(__uint_as_float(o[q * 8 + 4]), __uint_as_float(o[q * 8 + 5])),
                                pack_bf16(__uint_as_float(o[q * 8 + 6]), __uint_as_float(o[q * 8 + 7])));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
        if (t == NKT - 1) {
          if (a.early && tt + 1 < total_tiles) {       // read out behind the next tile's first chunk (see chunk 0 above)
            pend_dq = true; pend_b = b; pend_qh = qh; pend_h = h;
          } else {
            mbar_wait(mma2_done, tt & 1);
            dq_readout(b, qh, h);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// dTable[(code_i - code_j + off), h] += sum_b dS^T[b, h, j, i]   (bf16 dS^T, fp32 accumulation).
// One thread per 8 consecutive queries of one key row (16-byte loads), 256-thread blocks over the flattened
// (key row, query octet) index; grid = (position blocks, heads, batch splits); static 7x7 codes.
// seq = key rows per (window, head); nqv = valid query columns (== seq, or 196 per query half of a 392-token window, whose
// code offset the caller folds into code_off).
__global__ void __launch_bounds__(256) attn_w7_dbias_kernel(const __nv_bfloat16* ds, int batch, int heads, int seq, int ld,
                                                            int nqv, int code_off, float* dtable) {
  const int oct = ld >> 3;
  const int pos = blockIdx.x * 256 + threadIdx.x;
  if (pos >= seq * oct) return;
  const int j = pos / oct, i0 = (pos - j * oct) * 8;
  const int h = blockIdx.y;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)heads * seq * ld;
  const __nv_bfloat16* p = ds + ((long long)h * seq + j) * ld + i0;
  const int per = (batch + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * per, b1 = min(batch, b0 + per);
#pragma unroll 8
  for (int b = b0; b < b1; ++b) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p + b * stride));
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
    acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
  }
  const int cj = (j / 49) * W7_SH + ((j % 49) / 7) * W7_SW + (j % 7);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int i = i0 + e;
    if (i < nqv) {
      const int ci = (i / 49) * W7_SH + ((i % 49) / 7) * W7_SW + (i % 7);
      atomicAdd(dtable + (long long)(ci - cj + code_off) * heads + h, acc[e]);
    }
  }
}

// Same reduction from the per-CTA accumulation buffers of the dump_ds == 2 mode: buffer (c, slot, qh) holds the sum of
// dS^T over the units of head h_first(c) + slot that CTA c of the backward kernel processed (its contiguous unit range is
// recomputed here from the same formula), so a thread walks the <= grid_bwd CTAs and adds the buffers of ITS head.
__global__ void __launch_bounds__(256) attn_w7_dbias_acc_kernel(const __nv_bfloat16* ds, int grid_bwd, long long units, int batch,
                                                                int heads, int nh, int spans, int kseq, int ld, int nqv,
                                                                int code_off, float* dtable) {
  const int oct = ld >> 3;
  const int pos = blockIdx.x * 256 + threadIdx.x;
  const int h = blockIdx.y, qh = blockIdx.z;
  // the buffers that hold head h: found once per block (the backward kernel's unit ranges are recomputed from its formula)
  __shared__ int hit[256];
  __shared__ int nhit;
  if (threadIdx.x == 0) nhit = 0;
  __syncthreads();
  {
    const long long base = units / grid_bwd, rem = units % grid_bwd;
    for (int c = threadIdx.x; c < grid_bwd; c += 256) {
      const long long u0 = c * base + (c < rem ? c : rem), u1 = u0 + base + (c < rem ? 1 : 0);
      if (u1 <= u0) continue;
      const int h0 = (int)((u0 / nh) / batch), h1 = (int)(((u1 - 1) / nh) / batch);
      if (h >= h0 && h <= h1) hit[atomicAdd(&nhit, 1)] = (int)((((long long)c * spans + (h - h0)) * nh) + qh);
    }
  }
  __syncthreads();
  if (pos >= kseq * oct) return;
  const int j = pos / oct, i0 = (pos - j * oct) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < nhit; ++k) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(ds + ((long long)hit[k] * kseq + j) * ld + i0));
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
    acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
  }
  const int cj = (j / 49) * W7_SH + ((j % 49) / 7) * W7_SW + (j % 7);
  const int off = code_off + qh * 4 * W7_SH;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int i = i0 + e;
    if (i < nqv && acc[e] != 0.f) {
      const int ci = (i / 49) * W7_SH + ((i % 49) / 7) * W7_SW + (i % 7);
      atomicAdd(dtable + (long long)(ci - cj + off) * heads + h, acc[e]);
    }
  }
}

// dqkv[row, C .. 3C) += part[row, 0 .. 2C)   (dK | dV partial sums of the second query half; 8 bf16 per thread)
__global__ void __launch_bounds__(256) attn_w7_dkv_combine_kernel(__nv_bfloat16* dqkv, const __nv_bfloat16* part, long long rows, int C) {
  const int per_row = 2 * C / 8;
  const long long total = rows * per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
    const long long row = x / per_row; const int c = (int)(x % per_row) * 8;
    uint4* dst = reinterpret_cast<uint4*>(dqkv + row * 3 * C + C + c);
    const uint4 p = __ldg(reinterpret_cast<const uint4*>(part + row * 2 * C + c));
    uint4 d = *dst;
    const float2 a0 = unpack_bf16(d.x), a1 = unpack_bf16(d.y), a2 = unpack_bf16(d.z), a3 = unpack_bf16(d.w);
    const float2 b0 = unpack_bf16(p.x), b1 = unpack_bf16(p.y), b2 = unpack_bf16(p.z), b3 = unpack_bf16(p.w);
    d.x = pack_bf16(a0.x + b0.x, a0.y + b0.y); d.y = pack_bf16(a1.x + b1.x, a1.y + b1.y);
    d.z = pack_bf16(a2.x + b2.x, a2.y + b2.y); d.w = pack_bf16(a3.x + b3.x, a3.y + b3.y);
    *dst = d;
  }
}

static int w7_check(const clv_attn_w7_desc_t* d, const char* who, int max_wd) {
  CLV_REQUIRE(d != nullptr, "%s: null descriptor", who);
  CLV_REQUIRE(d->batch > 0 && d->heads > 0 && d->wd >= 2 && d->wd <= max_wd && d->wd % 2 == 0,
              "%s: wd must be even in [2, %d] (got %d)", who, max_wd, d->wd);
  CLV_REQUIRE(d->bias_table && d->cfg_wd >= d->wd && d->table_len == (2 * d->cfg_wd - 1) * W7_SH,
              "%s: bias table must be the (2*cfg_wd-1)*169 table of a (cfg_wd,7,7) window (len %d, cfg_wd %d)", who,
              d->table_len, d->cfg_wd);
  CLV_REQUIRE((d->q_ext == nullptr) == (d->k_ext == nullptr) && (!d->q_ext || d->nwin > 0), "%s: q_ext/k_ext/nwin go together", who);
  CLV_REQUIRE((long long)d->batch * d->wd * 49 < 2000000000LL, "%s: too many rows", who);
  return 0;
}

}  // namespace clv

using namespace clv;

// staged table: 2 wd - 1 depth offsets at stride PH, the last one cut after its 13 x 13 block; floats per head, multiple of 4
static int w7_table_ld(const clv_attn_w7_desc_t* d) { return (2 * (d->wd - 1) * W7_PH + 12 * W7_PW + 13 + 3) & ~3; }
static long long w7_table_bytes(const clv_attn_w7_desc_t* d) {
  return ((long long)w7_table_ld(d) * d->heads * 4 + 255) / 256 * 256;
}

extern "C" long long clv_attention_w7_fwd_workspace_bytes(const clv_attn_w7_desc_t* d) { return d ? w7_table_bytes(d) : 0; }

extern "C" int clv_attention_w7_fwd(const clv_attn_w7_desc_t* d, const void* qkv, void* out, float* lse, void* workspace,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = w7_check(d, "attention_w7_fwd", 8)) return rc;
  CLV_REQUIRE(qkv && out && lse && workspace, "attention_w7_fwd: null pointer");
  CLV_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "attention_w7_fwd: workspace must be 16-byte aligned");
  W7FwdArgs a{};
  a.batch = d->batch; a.heads = d->heads; a.seq = 49 * d->wd;
  // query tiles of 128 rows use every TMEM lane: 7 instead of 8 warp-passes per 196-token window (-4 % where two CTAs share an
  // SM and the softmax is partly throughput-bound; at 392 tokens -- one CTA per SM, latency-bound -- the ragged fourth tile
  // costs 0.8 %, so the 98-row tiles stay there; w7_fwd_qtile = 0 / 1 forces 98 / 128; tools/ab_w7_fwd_qtile.py)
  a.q_rows = tunable(TUNE_W7_FWD_QTILE, a.seq <= 196 ? 1 : 0) == 1 ? 128 : W7_TILE;
  a.n_qt = (a.seq + a.q_rows - 1) / a.q_rows;
  a.nmma = (a.seq + 15) / 16 * 16;
  a.n0 = a.nmma <= 256 ? a.nmma : 208;
  a.kb_bytes = (a.nmma * W7_ROWB + 1023) / 1024 * 1024;
  a.kx_bytes = (a.nmma * W7_XROWB + 1023) / 1024 * 1024;
  a.tmem_cols = a.nmma + W7_HD <= 256 ? 256 : 512;
  a.col_o = a.nmma;
  a.units = (long long)d->batch * d->heads * a.n_qt;
  a.out = reinterpret_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.table_len = d->table_len; a.table_ld = w7_table_ld(d);
  {
    float* tt = reinterpret_cast<float*>(workspace);
    const int n = a.table_ld * d->heads;
    attn_w7_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d->bias_table, tt, a.table_ld, d->heads, d->wd, d->cfg_wd);
    if (int rc = after_launch("attn_w7_table_kernel")) return rc;
    a.table_t = tt;
  }
  a.code_off = (d->wd - 1) * W7_PH + 6 * W7_PW + 6;            // staged-table offset of (dz, dy, dx) = 0
  a.has_ext = d->q_ext != nullptr; a.nwin = d->nwin > 0 ? d->nwin : 1;
  const long long rows = (long long)d->batch * a.seq;
  const long long ld = 3LL * d->heads * W7_HD;
  const int n1 = a.nmma - a.n0;
  CUtensorMap tq, tkv0, tkv1, tqx, tkx0, tkx1;
  if (int rc = make_tmap_bf16_2d(&tq, qkv, ld, rows, ld, W7_HD, 128, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tkv0, qkv, ld, rows, ld, W7_HD, a.n0, 64)) return rc;
  tkv1 = tkv0; tqx = tq; tkx0 = tq; tkx1 = tq;
  if (n1 > 0) { if (int rc = make_tmap_bf16_2d(&tkv1, qkv, ld, rows, ld, W7_HD, n1, 64)) return rc; }
  if (a.has_ext) {
    const long long xrows = (long long)a.nwin * a.seq;
    if (int rc = make_tmap_bf16_2d(&tqx, d->q_ext, 16, xrows, 16, 16, 128, 32)) return rc;
    if (int rc = make_tmap_bf16_2d(&tkx0, d->k_ext, 16, xrows, 16, 16, a.n0, 32)) return rc;
    tkx1 = tkx0;
    if (n1 > 0) { if (int rc = make_tmap_bf16_2d(&tkx1, d->k_ext, 16, xrows, 16, 16, n1, 32)) return rc; }
  }
  const size_t stage = 8192 + 2 * (size_t)a.kb_bytes + (a.has_ext ? 4096 + (size_t)a.kx_bytes : 0);
  // second-generation forward (query tiles cut at multiples of 32 rows, eight softmax warps per CTA, rows split between two
  // threads with a max / sum exchange): it was +9 % at 392 tokens when written, but the first-generation kernel has since
  // gained more from the elect.sync role entry and the barrier wait hints and is now 3-10 % faster at every size
  // (tools/ab_w7_fwd2.py, profiles/r02x_attn_microbench_fwd_generations.jsonl), so it is the default again; w7_fwd2 = 1
  // selects the second generation
  const long long fwd2 = tunable(TUNE_W7_FWD2, 0);
  // early issue of the next tile's score MMAs + split P V products: together -2 % at 196 tokens, -3.6 % at 392 (one CTA per SM,
  // nothing else hides the MMA latency); each alone is neutral at 196 (tools/ab_w7_fwd_early.py, profiles/r04b_*)
  a.early = (int)tunable(TUNE_W7_FWD_EARLY, 1);
  if (fwd2 == 1) {
    a.n_qt = (a.seq + 127) / 128;
    a.tile_rows = a.n_qt == 1 ? a.seq : (a.seq / a.n_qt) / 32 * 32;
    a.units = (long long)d->batch * d->heads * a.n_qt;
    const size_t smem2 = 1024 + 2 * stage + (size_t)a.table_ld * 4 + 2048 + 9 * 8 + 16;
    CLV_REQUIRE(smem2 <= 227 * 1024, "attention_w7_fwd: %zu bytes of shared memory needed", smem2);
    auto kern = d->wd == 2 ? attn_w7_fwd2_kernel<98> : d->wd == 4 ? attn_w7_fwd2_kernel<196>
              : d->wd == 6 ? attn_w7_fwd2_kernel<294> : attn_w7_fwd2_kernel<392>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)smem2)) return rc;
    const int per_sm2 = (a.tmem_cols == 256 && smem2 <= 113 * 1024) ? 2 : 1;
    const int grid2 = (int)std::min<long long>(a.units, (long long)num_sms() * per_sm2);
    kern<<<grid2, W7_FWD2_THREADS, smem2, stream>>>(tq, tkv0, tkv1, tqx, tkx0, tkx1, a);
    return after_launch("attn_w7_fwd2_kernel");
  }
  const size_t smem = 1024 + 2 * stage + (size_t)a.table_ld * 4 + 10 * 8 + 16;
  CLV_REQUIRE(smem <= 227 * 1024, "attention_w7_fwd: %zu bytes of shared memory needed", smem);
  const int per_sm = (a.tmem_cols == 256 && smem <= 113 * 1024) ? 2 : 1;
  const int grid = (int)std::min<long long>(a.units, (long long)num_sms() * per_sm);
  {   // w7_fwd_pvsplit: 0 off, 1 (default) split after half of the key bodies (never for a single body)
    const int n_body = a.seq / W7_TILE;
    a.split_body = (tunable(TUNE_W7_FWD_PVSPLIT, 1) == 1 && n_body >= 2) ? n_body / 2 : 0;
  }
  a.dbg = reinterpret_cast<long long*>(tunable(TUNE_W7_FWD_DBG, 0));        // device buffer of grid * 8 int64 (tools only)
  if (a.dbg) {
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_w7_fwd_kernel<true>), (int)smem)) return rc;
    attn_w7_fwd_kernel<true><<<grid, W7_FWD_THREADS, smem, stream>>>(tq, tkv0, tkv1, tqx, tkx0, tkx1, a);
    return after_launch("attn_w7_fwd_kernel<dbg>");
  }
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_w7_fwd_kernel<false>), (int)smem)) return rc;
  attn_w7_fwd_kernel<false><<<grid, W7_FWD_THREADS, smem, stream>>>(tq, tkv0, tkv1, tqx, tkx0, tkx1, a);
  return after_launch("attn_w7_fwd_kernel");
}

extern "C" long long clv_attention_w7_bwd_workspace_bytes(const clv_attn_w7_desc_t* d, int with_dbias) {
  if (!d) return 0;
  const long long seq = 49LL * d->wd;
  const long long nq = (seq + 15) / 16 * 16;
  long long bytes = w7_table_bytes(d) + ((long long)d->batch * d->heads * seq * 32 + 255) / 256 * 256 + 1024;   // table^T, e rows (16 bf16)
  if (d->wd == 8) {             // query-half units: dK | dV partial of half 1 (bf16 [rows, 2C]), dS^T of both halves [2][b*h*392, 208]
    bytes += ((long long)d->batch * seq * 2 * d->heads * W7_HD * 2 + 255) / 256 * 256;
    if (with_dbias) bytes += 2LL * d->batch * d->heads * seq * 208 * 2 + 256;
    return bytes;
  }
  if (with_dbias) bytes += (long long)d->batch * d->heads * seq * nq * 2 + 256;                 // bf16 dS^T
  return bytes;
}

extern "C" int clv_attention_w7_bwd(const clv_attn_w7_desc_t* d, const void* qkv, const void* out, const void* dout,
                                    const float* lse, void* dqkv, float q_scale, float* dbias_table, void* workspace,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = w7_check(d, "attention_w7_bwd", 8)) return rc;
  CLV_REQUIRE(d->wd != 6, "attention_w7_bwd: wd must be 2, 4 or 8 (got 6)");
  CLV_REQUIRE(qkv && out && dout && lse && dqkv && workspace, "attention_w7_bwd: null pointer");
  CLV_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "attention_w7_bwd: workspace must be 256-byte aligned");
  const bool halves = d->wd == 8;        // 392 keys: units are (window, head, query half) of the 196-query kernel
  W7BwdArgs a{};
  a.batch = d->batch; a.heads = d->heads; a.seq = 49 * d->wd; a.n_kt = d->wd / 2;
  a.nq = halves ? 208 : (a.seq + 15) / 16 * 16;
  a.qb_bytes = (a.nq * W7_ROWB + 1023) / 1024 * 1024;
  a.eb_bytes = (a.nq * W7_XROWB + 1023) / 1024 * 1024;
  a.n_mq = (a.nq + 127) / 128;
  a.col_dp = a.nq;
  a.col_dq = 2 * a.nq;
  const int need_cols = 2 * a.nq + a.n_mq * W7_HD;
  a.tmem_cols = need_cols <= 256 ? 256 : 512;
  CLV_REQUIRE(need_cols <= 512, "attention_w7_bwd: %d TMEM columns needed", need_cols);
  if (a.seq == 98) {                     // packed at [0,48) + [96,104): columns [56,88) of each region are dead after chunk 2
    a.col_dv = 56; a.col_dk = a.col_dp + 56;
  } else {                               // see the accumulation order in the kernel
    a.col_dv = 480; a.col_dk = a.col_dp + 64;
  }
  a.units = (long long)d->batch * d->heads * (halves ? 2 : 1);
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv); a.q_scale = q_scale;
  a.table_len = d->table_len; a.table_ld = w7_table_ld(d);
  {
    float* tt = reinterpret_cast<float*>(workspace);
    const int n = a.table_ld * d->heads;
    attn_w7_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d->bias_table, tt, a.table_ld, d->heads, d->wd, d->cfg_wd);
    if (int rc = after_launch("attn_w7_table_kernel")) return rc;
    a.table_t = tt;
  }
  a.code_off = (d->wd - 1) * W7_PH + 6 * W7_PW + 6;            // staged-table offset of (dz, dy, dx) = 0
  a.has_kx = d->k_ext != nullptr; a.nwin = d->nwin > 0 ? d->nwin : 1;
  a.pipe = (int)tunable(TUNE_W7_PIPE, 1);
  const long long rows = (long long)d->batch * a.seq;
  const long long t_bytes = w7_table_bytes(d);
  __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + t_bytes);
  const long long e_bytes = ((long long)d->batch * d->heads * a.seq * 32 + 255) / 256 * 256 + 1024;
  const long long part_bytes = halves ? ((long long)rows * 2 * d->heads * W7_HD * 2 + 255) / 256 * 256 : 0;
  a.dkv_part = halves ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + t_bytes + e_bytes) : nullptr;
  a.ds_half_rows = rows * d->heads;
  __nv_bfloat16* ds_out = dbias_table ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(workspace) + t_bytes + e_bytes + part_bytes) : nullptr;
  a.dump_ds = ds_out != nullptr;
  {
    const long long n = rows * d->heads;
    attn_w7_prep_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), lse,
        reinterpret_cast<const __nv_bfloat16*>(d->q_ext), a.nwin, e, rows, d->heads, a.seq);
    if (int rc = after_launch("attn_w7_prep_kernel")) return rc;
  }
  const long long ld = 3LL * d->heads * W7_HD, ldo = (long long)d->heads * W7_HD;
  CUtensorMap tfull, ttile, tdo, te, tkx, tds;
  if (int rc = make_tmap_bf16_2d(&tfull, qkv, ld, rows, ld, W7_HD, a.nq, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&ttile, qkv, ld, rows, ld, W7_HD, 128, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tdo, dout, ldo, rows, ldo, W7_HD, a.nq, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&te, e, 16, rows * d->heads, 16, 16, a.nq, 32)) return rc;
  tkx = te; tds = te;
  // bias-table gradient: either dump every unit's dS^T (batch * heads * seq * nq bf16, read back by attn_w7_dbias_kernel) or,
  // for the chunk-ring kernels, let TMA ADD the tiles into per-CTA buffers that stay in L2 (dump_ds == 2)
  const int gen2 = (int)tunable(TUNE_W7_BWD2, 1);
  a.early = (int)tunable(TUNE_W7_BWD_EARLY, 1);
  const bool ring = halves || (a.seq == 196 && gen2);
  const int nh = halves ? 2 : 1;
  const int grid = (int)std::min<long long>(a.units, (long long)num_sms());
  long long nbuf = 0;
  if (ds_out && ring) {
    // acc_mode: 0 never, 1 when the dump would be large (it then costs more HBM traffic than the L2 adds), 2 always
    const int acc_mode = (int)tunable(TUNE_W7_DBIAS_ACC, 1);
    const long long acc_min_bytes = tunable(TUNE_W7_DBIAS_ACC_MIN_MB, 128) << 20;
    int spans = 1;
    const long long base = a.units / grid, rem = a.units % grid;
    for (int c = 0; c < grid; ++c) {
      const long long u0 = c * base + (c < rem ? c : rem), u1 = u0 + base + (c < rem ? 1 : 0);
      if (u1 > u0) spans = std::max(spans, (int)(((u1 - 1) / nh) / d->batch - (u0 / nh) / d->batch) + 1);
    }
    nbuf = (long long)grid * spans * nh;
    const long long dump_bytes = rows * d->heads * nh * a.nq * 2;
    if (acc_mode && (acc_mode == 2 || dump_bytes >= acc_min_bytes) && nbuf * a.seq <= rows * d->heads * nh) {   // fits in the dump's space
      a.dump_ds = 2;
      a.ds_spans = spans;
      a.l2_hint = (int)tunable(TUNE_W7_L2_HINT, 1);
      CLV_CHECK_CUDA(cudaMemsetAsync(ds_out, 0, (size_t)nbuf * a.seq * a.nq * 2, stream));
    } else {
      nbuf = 0;
    }
  }
  if (ds_out) {
    const long long ds_rows = a.dump_ds == 2 ? nbuf * a.seq : rows * d->heads * nh;
    if (int rc = make_tmap_bf16_2d(&tds, ds_out, a.nq, ds_rows, a.nq, 64, W7_TILE, 128)) return rc;
  }
  if (a.has_kx) { if (int rc = make_tmap_bf16_2d(&tkx, d->k_ext, 16, (long long)a.nwin * a.seq, 16, 16, 128, 32)) return rc; }
  const size_t smem = 1024 + 2 * (2 * (size_t)a.qb_bytes + a.eb_bytes) + 2 * (2 * 8192 + 4096) + 4096 + (size_t)a.n_mq * 2 * 16384 +
                      (size_t)a.table_ld * 4 + 22 * 8 + 16;
  CLV_REQUIRE(smem <= 227 * 1024, "attention_w7_bwd: %zu bytes of shared memory needed", smem);
  auto kern = halves ? attn_w7_bwd2_kernel<392>
                     : (a.seq == 196 ? (gen2 ? attn_w7_bwd2_kernel<196> : attn_w7_bwd_kernel<196>) : attn_w7_bwd_kernel<98>);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)smem)) return rc;
  kern<<<grid, W7_BWD_THREADS, smem, stream>>>(tfull, ttile, tdo, te, tkx, tds, a);
  if (int rc = after_launch("attn_w7_bwd_kernel")) return rc;
  if (halves) {
    const long long work = rows * (2 * d->heads * W7_HD / 8);
    const int g = (int)std::min<long long>((work + 255) / 256, (long long)num_sms() * 16);
    attn_w7_dkv_combine_kernel<<<g, 256, 0, stream>>>(a.dqkv, a.dkv_part, rows, d->heads * W7_HD);
    if (int rc = after_launch("attn_w7_dkv_combine_kernel")) return rc;
  }
  const int code_off_ref = (d->cfg_wd - 1) * W7_SH + 6 * W7_SW + 6;     // the gradient is scattered in the reference's table layout
  if (dbias_table && a.dump_ds == 2) {
    dim3 g((a.seq * (a.nq / 8) + 255) / 256, d->heads, nh);
    attn_w7_dbias_acc_kernel<<<g, 256, 0, stream>>>(ds_out, grid, a.units, d->batch, d->heads, nh, a.ds_spans, a.seq, a.nq,
                                                    halves ? 196 : a.seq, code_off_ref, dbias_table);
    if (int rc = after_launch("attn_w7_dbias_acc_kernel")) return rc;
  } else if (dbias_table) {
    const int pos_blocks = (a.seq * (a.nq / 8) + 255) / 256;
    const int zsplit = std::max(1, std::min(std::min(64, d->batch / 8), (8 * num_sms()) / (pos_blocks * d->heads) + 1));
    dim3 g(pos_blocks, d->heads, zsplit);
    for (int qh = 0; qh < (halves ? 2 : 1); ++qh) {
      attn_w7_dbias_kernel<<<g, 256, 0, stream>>>(ds_out + (long long)qh * a.ds_half_rows * a.nq, d->batch, d->heads, a.seq, a.nq,
                                                 halves ? 196 : a.seq, code_off_ref + qh * 4 * W7_SH, dbias_table);
      if (int rc = after_launch("attn_w7_dbias_kernel")) return rc;
    }
  }
  return 0;
}
