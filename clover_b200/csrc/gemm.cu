// Persistent warp-specialised bf16 GEMM on tcgen05 / TMEM / TMA (sm_100a).
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )         fp32 accumulation in tensor memory
//
// Both operands may be K-major (row-major [rows, K]) or MN-major (stored transposed, [K, rows]),
// which covers forward (X W^T), dgrad (dY W) and wgrad (dY^T X) of every nn.Linear on the Clover
// hot path (reference: swin_transformer_3d.py:376,398,263,266,542; HF BertSelfAttention /
// BertSelfOutput / BertIntermediate / BertOutput denses; heads/ssl_head.py projections) without
// materialising a transposed copy.  Fused epilogues: bias, q-scale on a column prefix, erf-GELU
// (optionally also storing the pre-activation), multiply by GELU'(pre) (fc2 dgrad), residual add
// (fp32 or bf16), window-reverse row scatter (window_reverse + roll back, :471-474), bf16 or fp32
// output, and split-K with fp32 atomic accumulation for weight gradients.
//
// Roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-17 = epilogue (four column
// quarters x four TMEM lane quarters: TMEM -> registers -> 256-bit global stores; bias staged in shared memory), warps
// 18-19 = row sums of the MN-major A tiles (bias gradients; busy only in the RS instantiations).  Two TMEM accumulator
// stages let the epilogue of tile i overlap the mainloop of tile i+1.
// Tiles 128x256x64 (4 smem stages) when N is a multiple of 256 and the grid still fills, else 128x128x64 (6 stages).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "clover_b200.h"

namespace clv {

constexpr int BM = 128, BK = 64;
constexpr int GEMM_EPI_WARPS = 16;    // four column quarters x four lane quarters
constexpr int GEMM_RS_WARPS = 2;       // row-sum warps (only busy in the RS kernels)
// warp 0 TMA, warp 1 MMA, warps 2-17 epilogue, (RS kernels only) warps 18-19 row sums: 576 threads leave 112 registers per
// thread to the epilogue of the ordinary kernels, 640 threads (96 registers) only where the row-sum warps exist
template <bool RS> constexpr int gemm_threads() { return 64 + 32 * GEMM_EPI_WARPS + (RS ? 32 * GEMM_RS_WARPS : 0); }

// RS ("row sums"): the weight-gradient GEMMs dW = dY^T X also deliver the bias gradient sum_t dY[t, m] = the row sums of
// their (MN-major) A operand.  Two extra warps read every A tile of the n_idx == 0 tiles out of shared memory behind the
// TMA barrier (16-byte loads through the 128-byte swizzle), keep 8 fp32 partial sums per lane across the K loop and leave
// with 64 atomics per warp and tile; they are a third arriver on the stage's "empty" barrier.  The tensor pipe, the operand
// tiles and the accumulator layout are untouched (an extra N = 16 MMA per K step, or 16 extra B columns of ones, cost
// 6-33 % of these GEMMs), and the separate column-sum pass over dY (5.5 ms per step) disappears.
template <int BN, bool RS, bool BOX = false> struct GemmCfg {
  static constexpr int STAGES = BOX ? 3 : (BN == 128 ? 6 : 4);   // BOX: one pipeline stage makes room for the epilogue's row boxes
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = GEMM_EPI_WARPS * 2048;     // per epilogue warp: 32 rows x 64 B staged for TMA stores
  // shared memory of the kernels without / with the TMA-store stage: the per-lane store path wants the ~28 KB of L1 that
  // 200 KB of shared memory leave (its 32-byte row pieces merge into full lines there; 230 KB cost it 15-25 % on the
  // store-bound K = 128 shapes), so only the TMA-store instantiations pay for the stage
  static constexpr int SMEM_BASE = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int SMEM_TS = SMEM_BASE + STG_BYTES;
  static constexpr int BOX_BYTES = GEMM_EPI_WARPS * 4096;     // per epilogue warp: 32 rows x 128 B (64 bf16 columns)
  static constexpr int SMEM_BOX = SMEM_BASE + BOX_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;   // two accumulator stages
};

struct GemmEpi {
  const float* bias;
  const void* residual;
  void* out;
  __nv_bfloat16* out_pre;
  const __nv_bfloat16* gelu_pre;
  long long ld_res, ld_out, ld_pre, ld_gpre;
  int residual_bf16, out_bf16, act, atomic_out;
  int scale_cols;
  float scale;
  int use_row_map;
  int vec32;                // every pointer / pitch 32-byte aligned and N % 32 == 0: 256-bit global accesses
  int spec;                 // host: EpiSpec specialisation this epilogue matches exactly (0 = none)
  int atomic_vec;           // split-K / accumulate output rows are 16-byte aligned: 4-wide fp32 reductions
  int tma_out;              // bit 0: `out` leaves through TMA stores (tensor map tma_out), bit 1: `out_pre` too (tma_pre)
  const float* row_scale;   // per row-group factor on (acc + bias) before the residual (DropPath), or nullptr
  long long row_scale_rows;
  float* rowsum;            // [M] += sum_k A[m, k] (fp32 atomics), or nullptr: see GemmCfg (RS kernels, MN-major A only)
  WindowGeom geom;
};

CLV_DEVICE void ld256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
CLV_DEVICE void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
CLV_DEVICE void st256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// NE (16 or 32) bf16 / fp32 of one row <-> registers; `wide` selects the 256-bit path
template <int NE>
CLV_DEVICE void load_row(const void* base, int is_bf16, bool wide, int ncols, float (&f)[NE]) {
  if (is_bf16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base);
    if (wide) {
#pragma unroll
      for (int h = 0; h < NE / 16; ++h) {
        uint32_t r[8];
        ld256(p + h * 16, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float2 t = unpack_bf16(r[j]); f[h * 16 + 2 * j] = t.x; f[h * 16 + 2 * j + 1] = t.y; }
      }
    } else {
#pragma unroll
      for (int q = 0; q < NE / 8; ++q)
        if (q * 8 < ncols) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + q);
          const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
          f[q * 8] = a.x; f[q * 8 + 1] = a.y; f[q * 8 + 2] = b.x; f[q * 8 + 3] = b.y;
          f[q * 8 + 4] = c.x; f[q * 8 + 5] = c.y; f[q * 8 + 6] = d.x; f[q * 8 + 7] = d.y;
        }
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base);
    if (wide) {
#pragma unroll
      for (int h = 0; h < NE / 8; ++h) {
        uint32_t r[8];
        ld256(p + h * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[h * 8 + j] = __uint_as_float(r[j]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < NE / 4; ++q)
        if (q * 4 < ncols) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
          f[q * 4] = t.x; f[q * 4 + 1] = t.y; f[q * 4 + 2] = t.z; f[q * 4 + 3] = t.w;
        }
    }
  }
}
template <int NE>
CLV_DEVICE void store_row(void* base, int is_bf16, bool wide, int ncols, const float (&v)[NE]) {
  if (is_bf16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base);
    if (wide) {
#pragma unroll
      for (int h = 0; h < NE / 16; ++h) {
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = pack_bf16(v[h * 16 + 2 * j], v[h * 16 + 2 * j + 1]);
        st256(p + h * 16, r);
      }
    } else {
#pragma unroll
      for (int q = 0; q < NE / 8; ++q)
        if (q * 8 < ncols)
          reinterpret_cast<uint4*>(p)[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                                      pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
    }
  } else {
    float* p = reinterpret_cast<float*>(base);
    if (wide) {
#pragma unroll
      for (int h = 0; h < NE / 8; ++h) {
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(v[h * 8 + j]);
        st256(p + h * 8, r);
      }
    } else {
#pragma unroll
      for (int q = 0; q < NE / 4; ++q)
        if (q * 4 < ncols) reinterpret_cast<float4*>(p)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    }
  }
}

// Compile-time knowledge about the epilogue of the two hottest GEMM families.  The generic kernel decides every option at run
// time (a dozen uniform branches, parameter loads and predicated tails per 16-column step: ncu r02m counts 24 executed
// instructions per output element on fc1, 10 on a bias-only GEMM); the specialisations fold those decisions away.
//   SPEC 0  generic (all flags read from GemmEpi)
//   SPEC 1  fc1 of the Swin / BERT MLP: + bias, tanh-fit GELU, bf16 `out` and bf16 pre-activation copy, both through TMA stores
//   SPEC 2  fc2 dgrad: x GELU'(pre) with a bf16 `pre` row, bf16 out, per-lane 256-bit stores
//   SPEC 3 / 4 / 5 / 6  qkv, plain dgrad, fc2 forward, proj (see below)
// value -1: decided at run time
template <int SPEC> struct EpiSpec {
  static constexpr int bias = -1, act = -1, out_pre = -1, gelu_pre = -1, residual = -1, row_scale = -1, scale = -1, atomic = -1,
                       row_map = -1, wide = -1, out_bf16 = -1, dual = -1, pipe = 0, box = 0;
};
template <> struct EpiSpec<1> {
  static constexpr int bias = 1, act = 1, out_pre = 1, gelu_pre = 0, residual = 0, row_scale = 0, scale = 0, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 1, pipe = 0, box = 0;
};
template <> struct EpiSpec<2> {
  static constexpr int bias = 0, act = 0, out_pre = 0, gelu_pre = 1, residual = 0, row_scale = 0, scale = 0, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 0, pipe = 0, box = 0;
};
template <> struct EpiSpec<3> {      // qkv: + bias, q columns scaled (run-time column count), bf16 out, per-lane stores
  static constexpr int bias = 1, act = 0, out_pre = 0, gelu_pre = 0, residual = 0, row_scale = 0, scale = -1, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 0, pipe = 1, box = 0;
};
template <> struct EpiSpec<4> {      // activation gradients (dgrad): accumulator -> bf16, nothing else
  static constexpr int bias = 0, act = 0, out_pre = 0, gelu_pre = 0, residual = 0, row_scale = 0, scale = 0, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 0, pipe = 1, box = 0;
};
template <> struct EpiSpec<5> {      // fc2 forward: + bias, (DropPath row scale), + residual, fp32 out through TMA stores
  static constexpr int bias = 1, act = 0, out_pre = 0, gelu_pre = 0, residual = 1, row_scale = -1, scale = 0, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 0, dual = 0, pipe = 0, box = 0;
};
template <> struct EpiSpec<6> {      // proj: + bias, (DropPath row scale), + residual, fp32 out scattered through window_reverse
  static constexpr int bias = 1, act = 0, out_pre = 0, gelu_pre = 0, residual = 1, row_scale = -1, scale = 0, atomic = 0,
                       row_map = 1, wide = 1, out_bf16 = 0, dual = 0, pipe = 0, box = 0;
};
template <> struct EpiSpec<7> {      // fc2 dgrad with 256-column tiles: the pre-activation rows arrive, and the results leave, as
  static constexpr int bias = 0, act = 0, out_pre = 0, gelu_pre = 1, residual = 0, row_scale = 0, scale = 0, atomic = 0,   // TMA boxes
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 0, pipe = 0, box = 1;
};
template <> struct EpiSpec<8> {      // qkv (bias, q-scale) with 256-column tiles, results leave as TMA boxes
  static constexpr int bias = 1, act = 0, out_pre = 0, gelu_pre = 0, residual = 0, row_scale = 0, scale = -1, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 0, pipe = 0, box = 1;
};
template <> struct EpiSpec<9> {      // plain dgrad with 256-column tiles, results leave as TMA boxes
  static constexpr int bias = 0, act = 0, out_pre = 0, gelu_pre = 0, residual = 0, row_scale = 0, scale = 0, atomic = 0,
                       row_map = 0, wide = 1, out_bf16 = 1, dual = 0, pipe = 0, box = 1;
};
#define EPI_IS(field, runtime) (S::field < 0 ? (runtime) : (S::field != 0))

// Element math of one 16-column epilogue step: accumulator -> v (final fp32 values) and pk (packed bf16 pre-activation when
// act == 1 && out_pre).  Returns false when the lane has nothing to do (row beyond M, columns beyond N).
template <int EC, int SPEC>
CLV_DEVICE bool gemm_chunk_math(const GemmEpi& ep, uint32_t taddr, int n0, int N, long long row, long long drow, bool row_ok,
                          float rscale, float (&v)[EC], uint32_t (&pk)[EC / 2], int& ncols) {
  using S = EpiSpec<SPEC>;
  static_assert(EC == 16, "one 16-column step");
  uint32_t r[EC];
  tmem_ld_32x16(taddr, r);
  // Specialised epilogues have registers to spare (64-80 of 96): what the step reads from global memory does not depend on
  // the accumulator, so it is requested BEFORE waiting for the tensor-memory load and the two latencies overlap.  (The
  // generic kernel sits at the 96-register cap; hoisting there spills and costs 30 %.)
  constexpr bool HOIST = SPEC != 0;
  [[maybe_unused]] float4 hb[EC / 4];
  [[maybe_unused]] uint32_t ha[8], ha2[8];
  if constexpr (HOIST) {
    if (row_ok && n0 < N) {
      if constexpr (S::bias == 1) {
#pragma unroll
        for (int q = 0; q < EC / 4; ++q) hb[q] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + q);
      }
      if constexpr (S::gelu_pre == 1) ld256(ep.gelu_pre + row * ep.ld_gpre + n0, ha);
      if constexpr (S::residual == 1) {
        if (ep.residual_bf16) {
          ld256(reinterpret_cast<const __nv_bfloat16*>(ep.residual) + drow * ep.ld_res + n0, ha);
        } else {
          const float* p = reinterpret_cast<const float*>(ep.residual) + drow * ep.ld_res + n0;
          ld256(p, ha);
          ld256(p + 8, ha2);
        }
      }
    }
  }
  tmem_ld_wait();

  if (!row_ok || n0 >= N) return false;
#pragma unroll
  for (int j = 0; j < EC; ++j) v[j] = __uint_as_float(r[j]);
  const bool wide = EPI_IS(wide, ep.vec32 != 0);    // implies ncols == EC
  ncols = wide ? EC : min(EC, N - n0);               // multiple of 8 (N % 8 == 0)
  if (EPI_IS(atomic, ep.atomic_out)) return true;
  if (EPI_IS(bias, ep.bias != nullptr)) {
#pragma unroll
    for (int q = 0; q < EC / 4; ++q) {
      if (wide || q * 4 < ncols) {    // N % 8 == 0 and n0 % 16 == 0: whole float4 groups are in range (bias is 16-byte aligned)
        float4 b4;
        if constexpr (HOIST && S::bias == 1) b4 = hb[q];
        else b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + q);
        v[q * 4] += b4.x; v[q * 4 + 1] += b4.y; v[q * 4 + 2] += b4.z; v[q * 4 + 3] += b4.w;
      }
    }
  }
  if (EPI_IS(scale, ep.scale_cols > n0)) {
#pragma unroll
    for (int j = 0; j < EC; ++j)
      if (n0 + j < ep.scale_cols) v[j] *= ep.scale;
  }
  if (EPI_IS(act, ep.act == 1)) {
    if (EPI_IS(out_pre, ep.out_pre != nullptr)) {
#pragma unroll
      for (int j = 0; j < EC / 2; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < EC; ++j) v[j] = gelu_fit(v[j]);
  }
  if (EPI_IS(gelu_pre, ep.gelu_pre != nullptr)) {
    if constexpr (HOIST && S::gelu_pre == 1) {
#pragma unroll
      for (int j = 0; j < EC / 2; ++j) {
        const float2 t = unpack_bf16(ha[j]);
        v[2 * j] *= gelu_fit_grad(t.x); v[2 * j + 1] *= gelu_fit_grad(t.y);
      }
    } else {
      float g[EC];
      load_row<EC>(ep.gelu_pre + row * ep.ld_gpre + n0, 1, wide, ncols, g);
#pragma unroll
      for (int j = 0; j < EC; ++j) v[j] *= gelu_fit_grad(g[j]);
    }
  }
  if (EPI_IS(row_scale, ep.row_scale != nullptr)) {
#pragma unroll
    for (int j = 0; j < EC; ++j) v[j] *= rscale;
  }
  if constexpr (HOIST && S::residual == 1) {
    if (ep.residual_bf16) {
#pragma unroll
      for (int j = 0; j < EC / 2; ++j) { const float2 t = unpack_bf16(ha[j]); v[2 * j] += t.x; v[2 * j + 1] += t.y; }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[j] += __uint_as_float(ha[j]); v[8 + j] += __uint_as_float(ha2[j]); }
    }
  } else if (EPI_IS(residual, ep.residual != nullptr)) {
    float g[EC];
    if (ep.residual_bf16)
      load_row<EC>(reinterpret_cast<const __nv_bfloat16*>(ep.residual) + drow * ep.ld_res + n0, 1, wide, ncols, g);
    else
      load_row<EC>(reinterpret_cast<const float*>(ep.residual) + drow * ep.ld_res + n0, 0, wide, ncols, g);
#pragma unroll
    for (int j = 0; j < EC; ++j) v[j] += g[j];
  }
  return true;
}

// tile t -> (K split, n tile, m tile), t = (m * num_n + n) * k_splits + split.  32-bit unsigned divisions (the launcher checks
// the tile count): the 64-bit forms cost two ~300-cycle subroutine calls per tile in every role, on the critical path of the
// epilogue warps of the K <= 512 GEMMs.
CLV_DEVICE void gemm_tile_coords(uint32_t t, int k_splits, int num_n, int& split, int& n_idx, int& m_idx) {
  uint32_t mn = t;
  split = 0;
  if (k_splits != 1) { mn = t / (uint32_t)k_splits; split = (int)(t - mn * (uint32_t)k_splits); }
  const uint32_t m = mn / (uint32_t)num_n;
  m_idx = (int)m;
  n_idx = (int)(mn - m * (uint32_t)num_n);
}

template <int A_MN, int B_MN, int BN, bool RS, bool TS, int SPEC>
__global__ void __launch_bounds__(gemm_threads<RS>(), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_pre,
                 int M, int N, int K, int k_splits, GemmEpi ep) {
  using S = EpiSpec<SPEC>;
  constexpr bool BOX = S::box == 1;
  using Cfg = GemmCfg<BN, RS, BOX>;
  static_assert(!BOX || (BN == 256 && !TS && !RS), "the box epilogue is written for 128 x 256 tiles");
  static_assert(!RS || A_MN == 1, "row sums are implemented for the MN-major A operand of the weight gradients");
  constexpr int STAGES = Cfg::STAGES, A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (an integer round-trip would demote every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sStage = smem + STAGES * STAGE_BYTES;                                 // TS: [16 warps][2 KB], 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + (TS ? Cfg::STG_BYTES : (BOX ? Cfg::BOX_BYTES : 0)));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* box_bar = tempty_bar + 3;                                            // BOX: one per epilogue warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per_split = (num_kb + k_splits - 1) / k_splits;
  const long long num_tiles = (long long)num_m * num_n * k_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (TS) tma_prefetch_desc(&tma_out);
    if (TS && (ep.tma_out & 2)) tma_prefetch_desc(&tma_pre);
    if (BOX) {
      tma_prefetch_desc(&tma_out);
      if (S::gelu_pre == 1) tma_prefetch_desc(&tma_pre);
      for (int w = 0; w < GEMM_EPI_WARPS; ++w) mbar_init(&box_bar[w], 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RS ? 1 + GEMM_RS_WARPS : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int split, n_idx, m_idx;
        gemm_tile_coords((uint32_t)t, k_splits, num_n, split, n_idx, m_idx);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(num_kb, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (A_MN) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], m_idx * BM, kb * BK);
            tma_load_2d(sa + A_BYTES / 2, &tma_a, &full_bar[stage], m_idx * BM + 64, kb * BK);
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_idx * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (BK * 128), &tma_b, &full_bar[stage], n_idx * BN + c * 64, kb * BK);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c)
              tma_load_2d(sb + c * (128 * BK * 2), &tma_b, &full_bar[stage], kb * BK, n_idx * BN + c * 128);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int split = k_splits == 1 ? 0 : (int)((uint32_t)t % (uint32_t)k_splits);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(num_kb, kb0 + kb_per_split);
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart (a 256-row B tile is
            // two stacked 128-row boxes, still 1024 B per group).  MN-major: 16 k-rows = 2 groups of 8 rows x 128 B
            // = 2048 B; 64-element MN chunks BK*128 B apart.
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, BK * 128, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, BK * 128, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 2 + GEMM_EPI_WARPS) {
    // ---------------- row sums of the A tiles (RS kernels): warp w owns the 64-row half w of the tile ----------------
    if (RS) {
      const int half = warp - (2 + GEMM_EPI_WARPS);          // which 64 x BK box of the A tile (TMA loads two of them)
      const int unit = lane & 7, kq = lane >> 3;             // 16-byte unit (8 rows m) of the 128-byte line; k rows kq, kq+4, ...
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int split, n_idx, m_idx;
        gemm_tile_coords((uint32_t)t, k_splits, num_n, split, n_idx, m_idx);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(num_kb, kb0 + kb_per_split);
        // the num_n tiles that share this A panel split its K blocks between them (kb % num_n == n_idx), so no tile's
        // row-sum warps have more than 1 / num_n of the panel to read while its MMAs run
        const bool mine = ep.rowsum != nullptr;
        float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (mine && (kb % num_n) == n_idx) {
            const uint8_t* sa = smem + stage * STAGE_BYTES + half * (A_BYTES / 2);
#pragma unroll
            for (int k = kq; k < BK; k += 4) {
              const uint4 u = *reinterpret_cast<const uint4*>(sa + k * 128 + ((unit ^ (k & 7)) << 4));
              const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
              acc8[0] += a0.x; acc8[1] += a0.y; acc8[2] += a1.x; acc8[3] += a1.y;
              acc8[4] += a2.x; acc8[5] += a2.y; acc8[6] += a3.x; acc8[7] += a3.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], 8);
            acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], 16);
          }
          if (kq == 0) {
            const int row = m_idx * BM + half * 64 + unit * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (row + j < M) atomicAdd(ep.rowsum + row + j, acc8[j]);
          }
        }
      }
    }
  } else {
    // ---------------- epilogue: 16 warps; TMEM lanes (warp % 4) * 32 ... +31, column quarter (warp - 2) / 4 -------------
    // (issue-bound for K <= 512: four warps per scheduler hide the dependent-latency stalls of the element math)
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int CPW = BN / (GEMM_EPI_WARPS / 4);   // columns per warp
    constexpr int EC = 16;                           // columns per epilogue step (register budget: 576 threads)
    constexpr int CHUNKS = CPW / EC;
    uint32_t it = 0;
    [[maybe_unused]] uint32_t box_phase = 0;
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int split_unused, n_idx, m_idx;
      gemm_tile_coords((uint32_t)t, k_splits, num_n, split_unused, n_idx, m_idx);
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      // bias: every lane of a warp needs the SAME 16 values per step -> four broadcast 128-bit loads straight from L1 / L2
      // (one wavefront each); staging the tile's bias in shared memory needed a 512-thread barrier per tile that cost 15 %
      // of a K = 512 GEMM (tools/epi_probe.py)
      // BOX: this warp's 32 rows x 64 columns of the pre-activation tensor are requested as ONE swizzled TMA box while the tile's
      // MMAs are still running (the per-lane path reads them as 32-byte row pieces: 32 L1 wavefronts per instruction, the busiest
      // unit of this epilogue in ncu r02m); the results overwrite the box in place and leave as one TMA store.
      [[maybe_unused]] uint8_t* box = nullptr;
      [[maybe_unused]] bool box_live = false;
      if constexpr (BOX) {
        box = sStage + (warp - 2) * 4096;
        const int bcol = n_idx * BN + chalf * CPW, brow = m_idx * BM + quarter * 32;
        box_live = bcol < N;
        if (lane == 0 && box_live) {
          tma_store_wait_read();                       // the previous tile's store has finished reading the box
          if constexpr (S::gelu_pre == 1) {
            mbar_expect_tx(&box_bar[warp - 2], 4096);
            tma_load_2d(box, &tma_pre, &box_bar[warp - 2], bcol, brow);
          }
        }
        if constexpr (S::gelu_pre != 1) __syncwarp();  // every lane may overwrite the box after lane 0's wait
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const long long row = (long long)m_idx * BM + quarter * 32 + lane;
      long long drow = row;
      if (EPI_IS(row_map, ep.use_row_map) && row < M) drow = window_row_to_src(ep.geom, row);
      const bool row_ok = row < M && drow >= 0;
      const float rscale = (EPI_IS(row_scale, ep.row_scale != nullptr) && row < M) ? __ldg(ep.row_scale + row / ep.row_scale_rows) : 1.0f;
      const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + chalf * CPW;
      if constexpr (TS) {
        // ---- outputs leave through TMA stores: the row-per-lane layout of a TMEM load makes every per-lane global store
        // touch 32 different lines (32 L1 wavefronts per instruction -- the LSU data pipe was the busiest unit of the
        // K <= 512 GEMMs, ncu r01c: 65 %); instead each lane drops its 16-byte pieces into the warp's swizzled 2 KB stage
        // (4 wavefronts per instruction) and one lane hands the 32-row x 64-byte box to the TMA unit.
        // Per 16-column step the warp's stage holds [32 rows][32 B] of bf16 `out` (+ [32 rows][32 B] of `out_pre` at +1 KB) or
        // [32 rows][64 B] of fp32 `out`; 16-byte pieces are XOR-swizzled exactly like the tensor map (SWIZZLE_32B / 64B).
        uint8_t* stg = sStage + (warp - 2) * 2048;
        const int box_row = m_idx * BM + quarter * 32;
        const bool dual = EPI_IS(dual, (ep.tma_out & 2) != 0);
        const bool out_bf16 = EPI_IS(out_bf16, ep.out_bf16 != 0);
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          float v[EC]; uint32_t pk[EC / 2]; int ncols;
          const int n0 = n_idx * BN + chalf * CPW + c * EC;
          const bool live = gemm_chunk_math<EC, SPEC>(ep, tacc + c * EC, n0, N, row, drow, row_ok, rscale, v, pk, ncols);
          // a single bf16 output needs 1 KB per step: the two halves of the stage alternate and only the step before the
          // previous one must have left shared memory; dual / fp32 outputs fill the whole stage every step
          const bool two_buf = out_bf16 && !dual;
          uint8_t* buf = stg + ((two_buf && (c & 1)) ? 1024 : 0);
          if (lane == 0) { if (two_buf) tma_store_wait_read_1(); else tma_store_wait_read(); }
          __syncwarp();
          if (live) {
            if (out_bf16) {
              uint8_t* myrow = buf + lane * 32;
              const int sw = (lane >> 2) & 1;
              *reinterpret_cast<uint4*>(myrow + ((0 ^ sw) << 4)) =
                  make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              *reinterpret_cast<uint4*>(myrow + ((1 ^ sw) << 4)) =
                  make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
              if (dual) {
                *reinterpret_cast<uint4*>(myrow + 1024 + ((0 ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(myrow + 1024 + ((1 ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            } else {
              uint8_t* myrow = stg + lane * 64;
              const int sw = (lane >> 1) & 3;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(myrow + ((q ^ sw) << 4)) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
            if (EPI_IS(act, ep.act == 1) && EPI_IS(out_pre, ep.out_pre != nullptr) && !dual) {
              uint4* pp = reinterpret_cast<uint4*>(ep.out_pre + row * ep.ld_pre + n0);
              pp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              pp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && n0 < N) {
            tma_store_2d(&tma_out, buf, n0, box_row);
            if (dual) tma_store_2d(&tma_pre, stg + 1024, n0, box_row);
            tma_store_commit();
          }
        }
      } else if constexpr (BOX) {
        if (box_live) {
          if constexpr (S::gelu_pre == 1) {
            mbar_wait(&box_bar[warp - 2], box_phase);
            box_phase ^= 1;
          }
          uint8_t* myrow = box + lane * 128;
          const int sw = lane & 7;
          const int nb = n_idx * BN + chalf * CPW;
#pragma unroll 1
          for (int c = 0; c < CHUNKS; ++c) {
            uint32_t r[EC];
            tmem_ld_32x16(tacc + c * EC, r);
            uint4* q0 = reinterpret_cast<uint4*>(myrow + (((2 * c) ^ sw) << 4));
            uint4* q1 = reinterpret_cast<uint4*>(myrow + (((2 * c + 1) ^ sw) << 4));
            uint32_t o[8];
            if constexpr (S::gelu_pre == 1) {
              const uint4 p0 = *q0, p1 = *q1;
              tmem_ld_wait();
              const uint32_t pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 t = unpack_bf16(pw[j]);
                o[j] = pack_bf16(__uint_as_float(r[2 * j]) * gelu_fit_grad(t.x), __uint_as_float(r[2 * j + 1]) * gelu_fit_grad(t.y));
              }
            } else {
              const int n0 = nb + c * EC;
              [[maybe_unused]] float4 hb[EC / 4];
              if constexpr (S::bias == 1) {
#pragma unroll
                for (int q = 0; q < EC / 4; ++q) hb[q] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + q);
              }
              tmem_ld_wait();
              float v[EC];
#pragma unroll
              for (int j = 0; j < EC; ++j) v[j] = __uint_as_float(r[j]);
              if constexpr (S::bias == 1) {
#pragma unroll
                for (int q = 0; q < EC / 4; ++q) {
                  v[q * 4] += hb[q].x; v[q * 4 + 1] += hb[q].y; v[q * 4 + 2] += hb[q].z; v[q * 4 + 3] += hb[q].w;
                }
              }
              if (EPI_IS(scale, ep.scale_cols > n0)) {
#pragma unroll
                for (int j = 0; j < EC; ++j)
                  if (n0 + j < ep.scale_cols) v[j] *= ep.scale;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
            }
            *q0 = make_uint4(o[0], o[1], o[2], o[3]);
            *q1 = make_uint4(o[4], o[5], o[6], o[7]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tma_out, box, nb, m_idx * BM + quarter * 32);
            tma_store_commit();
          }
        }
      } else if constexpr (S::pipe == 1) {
        // bias (+ q-scale) / plain bf16 epilogues need so few registers that the NEXT step's accumulator load can be in flight
        // while the current step is converted and stored: tcgen05.wait::ld waits for every outstanding load, so the next load
        // is issued right after the wait and overlaps the math / stores of the current step
        uint32_t rb[2][EC];
        tmem_ld_32x16(tacc, rb[0]);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int n0 = n_idx * BN + chalf * CPW + c * EC;
          const bool live = row_ok && n0 < N;
          [[maybe_unused]] float4 hb[EC / 4];
          if constexpr (S::bias == 1) {
            if (live) {
#pragma unroll
              for (int q = 0; q < EC / 4; ++q) hb[q] = __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + q);
            }
          }
          tmem_ld_wait();
          if (c + 1 < CHUNKS) tmem_ld_32x16(tacc + (c + 1) * EC, rb[(c + 1) & 1]);
          if (!live) continue;
          float v[EC];
#pragma unroll
          for (int j = 0; j < EC; ++j) v[j] = __uint_as_float(rb[c & 1][j]);
          if constexpr (S::bias == 1) {
#pragma unroll
            for (int q = 0; q < EC / 4; ++q) {
              v[q * 4] += hb[q].x; v[q * 4 + 1] += hb[q].y; v[q * 4 + 2] += hb[q].z; v[q * 4 + 3] += hb[q].w;
            }
          }
          if (EPI_IS(scale, ep.scale_cols > n0)) {
#pragma unroll
            for (int j = 0; j < EC; ++j)
              if (n0 + j < ep.scale_cols) v[j] *= ep.scale;
          }
          store_row<EC>(reinterpret_cast<__nv_bfloat16*>(ep.out) + row * ep.ld_out + n0, 1, true, EC, v);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          float v[EC]; uint32_t pk[EC / 2]; int n0, ncols;
          n0 = n_idx * BN + chalf * CPW + c * EC;
          if (!gemm_chunk_math<EC, SPEC>(ep, tacc + c * EC, n0, N, row, drow, row_ok, rscale, v, pk, ncols)) continue;
          const bool wide = EPI_IS(wide, ep.vec32 != 0);
          if (EPI_IS(atomic, ep.atomic_out)) {
            float* o = reinterpret_cast<float*>(ep.out) + drow * ep.ld_out + n0;
            if (ep.atomic_vec) {          // 16-byte aligned rows: one 4-wide reduction per four columns (a quarter of the L2 atomics)
#pragma unroll
              for (int q = 0; q < EC / 4; ++q)
                if (q * 4 < ncols) red_add_v4(o + q * 4, v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < EC; ++j)
                if (j < ncols) atomicAdd(o + j, v[j]);
            }
            continue;
          }
          if (EPI_IS(act, ep.act == 1) && EPI_IS(out_pre, ep.out_pre != nullptr)) {
            uint4* pp = reinterpret_cast<uint4*>(ep.out_pre + row * ep.ld_pre + n0);
            pp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (ncols > 8) pp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (EPI_IS(out_bf16, ep.out_bf16 != 0))
            store_row<EC>(reinterpret_cast<__nv_bfloat16*>(ep.out) + drow * ep.ld_out + n0, 1, wide, ncols, v);
          else
            store_row<EC>(reinterpret_cast<float*>(ep.out) + drow * ep.ld_out + n0, 0, wide, ncols, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if ((TS || BOX) && lane == 0) tma_store_wait_all();      // the stage must outlive the bulk reads; writes drain here
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D tensor map of 2-byte (bf16) or 4-byte (fp32) elements: inner dimension `inner` (contiguous), outer `outer`, row pitch ld elements.
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, long long inner, long long outer, long long ld,
                 int box_inner, int box_outer, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  CLV_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CLV_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "make_tmap_2d: 2- or 4-byte elements only");
  CLV_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * elem_bytes) % 16 == 0,
              "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ptr=%p ld=%lld)", ptr, ld);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CLV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r, inner,
              outer, ld);
  return 0;
}
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld,
                      int box_inner, int box_outer, int swizzle_bytes) {
  return make_tmap_2d(map, ptr, 2, inner, outer, ld, box_inner, box_outer, swizzle_bytes);
}

template <int A_MN, int B_MN, int BN, bool RS = false, bool TS = false, int SPEC = 0>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tp, int M, int N, int K,
                       int k_splits, const GemmEpi& ep, cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<A_MN, B_MN, BN, RS, TS, SPEC>;
  constexpr bool BOX = EpiSpec<SPEC>::box == 1;
  constexpr int SMEM = BOX ? GemmCfg<BN, RS, true>::SMEM_BOX : (TS ? GemmCfg<BN, RS>::SMEM_TS : GemmCfg<BN, RS>::SMEM_BASE);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), SMEM)) return rc;
  const long long tiles = (long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN) * k_splits;
  CLV_REQUIRE(tiles < (1LL << 31), "clv_gemm_bf16: %lld tiles exceed the 32-bit tile index of the kernel", tiles);
  const int grid = (int)std::min<long long>(tiles, num_sms());
  kern<<<grid, gemm_threads<RS>(), SMEM, stream>>>(ta, tb, to, tp, M, N, K, k_splits, ep);
  return after_launch("gemm_bf16_kernel launch");
}

template <int BN>
static int dispatch_gemm_rowsum(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                                const CUtensorMap& tp, int M, int N, int K, int k_splits, const GemmEpi& ep, cudaStream_t stream) {
  CLV_REQUIRE(a_mn, "clv_gemm_bf16: rowsum needs an MN-major A operand (a weight gradient dY^T X)");
  if (b_mn) return launch_gemm<1, 1, BN, true>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  return launch_gemm<1, 0, BN, true>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
}

template <int BN>
static int dispatch_gemm(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tp,
                         int M, int N, int K, int k_splits, const GemmEpi& ep, cudaStream_t stream) {
  // compile-time epilogues of the two hottest families (EpiSpec); anything else takes the generic kernel
  if (ep.spec == 1 && !a_mn && !b_mn) return launch_gemm<0, 0, BN, false, true, 1>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if constexpr (BN == 256) {
    if (ep.spec == 7 && !a_mn && b_mn) return launch_gemm<0, 1, 256, false, false, 7>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
    if (ep.spec == 8 && !a_mn && !b_mn) return launch_gemm<0, 0, 256, false, false, 8>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
    if (ep.spec == 9 && !a_mn && b_mn) return launch_gemm<0, 1, 256, false, false, 9>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  }
  if (ep.spec == 8) return launch_gemm<0, 0, BN, false, false, 3>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (ep.spec == 9) return launch_gemm<0, 1, BN, false, false, 4>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if ((ep.spec == 2 || ep.spec == 7) && !a_mn && b_mn) return launch_gemm<0, 1, BN, false, false, 2>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (ep.spec == 3 && !a_mn && !b_mn) return launch_gemm<0, 0, BN, false, false, 3>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (ep.spec == 4 && !a_mn && b_mn) return launch_gemm<0, 1, BN, false, false, 4>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (ep.spec == 5 && !a_mn && !b_mn) return launch_gemm<0, 0, BN, false, true, 5>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (ep.spec == 6 && !a_mn && !b_mn) return launch_gemm<0, 0, BN, false, false, 6>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (ep.tma_out & 1) {
    if (a_mn && b_mn) return launch_gemm<1, 1, BN, false, true>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
    if (a_mn) return launch_gemm<1, 0, BN, false, true>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
    if (b_mn) return launch_gemm<0, 1, BN, false, true>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
    return launch_gemm<0, 0, BN, false, true>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  }
  if (a_mn && b_mn) return launch_gemm<1, 1, BN>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (a_mn) return launch_gemm<1, 0, BN>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (b_mn) return launch_gemm<0, 1, BN>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  return launch_gemm<0, 0, BN>(ta, tb, to, tp, M, N, K, k_splits, ep, stream);
}

}  // namespace clv

using namespace clv;

extern "C" int clv_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                             int b_mn_major, int M, int N, int K, const clv_gemm_epilogue_t* e, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CLV_REQUIRE(A && B && e && e->out, "clv_gemm_bf16: null pointer");
  CLV_REQUIRE(M > 0 && N > 0 && K > 0, "clv_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  CLV_REQUIRE(N % 8 == 0, "clv_gemm_bf16: N must be a multiple of 8 (got %d)", N);
  CLV_REQUIRE(!e->bias || (reinterpret_cast<uintptr_t>(e->bias) & 15) == 0, "clv_gemm_bf16: bias must be 16-byte aligned");
  // 128x256 tiles when N fills them (less smem traffic per MAC); 128x128 otherwise
  const long long tiles256 = (long long)((M + BM - 1) / BM) * ((N + 255) / 256);
  const long long min_units256 = tunable(TUNE_GEMM_BN256_MIN_UNITS, 2LL * num_sms());   // the 2-waves rule
  // weight gradients (split-K, K blocks per tile in the hundreds) are L2-bound: always take the wide tile there
  const bool long_k = (e->k_splits > 1 || e->accumulate) && K / (e->k_splits > 0 ? e->k_splits : 1) >= 4096;
  const bool bn256 = (N % 256 == 0) && (tiles256 * (e->k_splits > 0 ? e->k_splits : 1) >= min_units256 || long_k);
  CUtensorMap ta, tb;
  int rc;
  if (a_mn_major) rc = make_tmap_bf16_2d(&ta, A, M, K, lda, 64, BK);
  else rc = make_tmap_bf16_2d(&ta, A, K, M, lda, BK, BM);
  if (rc) return rc;
  if (b_mn_major) rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, 64, BK);
  else rc = make_tmap_bf16_2d(&tb, B, K, N, ldb, BK, 128);
  if (rc) return rc;

  GemmEpi ep{};
  ep.bias = e->bias;
  ep.residual = e->residual;
  ep.out = e->out;
  ep.out_pre = reinterpret_cast<__nv_bfloat16*>(e->out_pre);
  ep.gelu_pre = reinterpret_cast<const __nv_bfloat16*>(e->gelu_pre);
  ep.ld_res = e->ld_residual; ep.ld_out = e->ld_out; ep.ld_pre = e->ld_pre; ep.ld_gpre = e->ld_gelu_pre;
  ep.residual_bf16 = e->residual_is_bf16; ep.out_bf16 = e->out_is_bf16; ep.act = e->act;
  ep.scale_cols = e->scale_cols; ep.scale = e->scale;
  ep.use_row_map = e->window != nullptr;
  ep.row_scale = e->row_scale;
  ep.row_scale_rows = e->row_scale_rows > 0 ? e->row_scale_rows : 1;
  ep.rowsum = e->rowsum;
  CLV_REQUIRE(!e->row_scale || e->row_scale_rows > 0, "clv_gemm_bf16: row_scale needs row_scale_rows > 0");
  int k_splits = e->k_splits > 0 ? e->k_splits : 1;
  const int num_kb = (K + BK - 1) / BK;
  if (k_splits > num_kb) k_splits = num_kb;
  {  // every split must own at least one k-block (the kernel's barrier protocol assumes it)
    const int per = (num_kb + k_splits - 1) / k_splits;
    k_splits = (num_kb + per - 1) / per;
  }
  ep.atomic_out = k_splits > 1 || e->accumulate;
  ep.atomic_vec = ep.atomic_out && (reinterpret_cast<uintptr_t>(e->out) & 15) == 0 && e->ld_out % 4 == 0;
  if (ep.atomic_out) {
    CLV_REQUIRE(!e->out_is_bf16 && !e->bias && !e->residual && !e->act && !e->gelu_pre && !e->window && !e->row_scale,
                "clv_gemm_bf16: split-K / accumulate supports plain fp32 output only");
    if (!e->accumulate)
      CLV_CHECK_CUDA(cudaMemset2DAsync(e->out, e->ld_out * 4, 0, (size_t)N * 4, M, stream));
  }
  if (e->window) {
    const clv_window_geom_t* w = e->window;
    WindowGeom& g = ep.geom;
    g.B = w->B; g.D = w->D; g.H = w->H; g.W = w->W; g.wd = w->wd; g.wh = w->wh; g.ww = w->ww;
    g.sd = w->sd; g.sh = w->sh; g.sw = w->sw;
    g.Dp = (w->D + w->wd - 1) / w->wd * w->wd; g.Hp = (w->H + w->wh - 1) / w->wh * w->wh;
    g.Wp = (w->W + w->ww - 1) / w->ww * w->ww;
    g.nD = g.Dp / g.wd; g.nH = g.Hp / g.wh; g.nW = g.Wp / g.ww; g.N = g.wd * g.wh * g.ww; g.nWin = g.nD * g.nH * g.nW;
    CLV_REQUIRE((long long)g.B * g.nWin * g.N == M, "clv_gemm_bf16: window geometry does not match M=%d", M);
  }
  {
    auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    bool ok = (N % 32 == 0) && al32(e->out) && (e->ld_out * (e->out_is_bf16 ? 2 : 4)) % 32 == 0;
    if (e->residual) ok = ok && al32(e->residual) && (e->ld_residual * (e->residual_is_bf16 ? 2 : 4)) % 32 == 0;
    if (e->out_pre) ok = ok && al32(e->out_pre) && (e->ld_pre * 2) % 32 == 0;
    if (e->gelu_pre) ok = ok && al32(e->gelu_pre) && (e->ld_gelu_pre * 2) % 32 == 0;
    ep.vec32 = ok ? 1 : 0;
  }
  // TMA-store epilogue: dense (unscattered) bf16 / fp32 outputs with 32-byte aligned rows and N % 32 == 0, so a 16-column
  // box is either fully inside N or fully outside; rows beyond M are clipped by the tensor map
  CUtensorMap to = ta, tp = ta;
  // Measured (tools/epi_probe.py, profiles/r02f_epilogue_probe_*): -7 % for the dual-output fc1 epilogue and for fp32 outputs;
  // a SINGLE bf16 output is as fast (K = 512) or faster (K = 128, store-bound: 32-byte-wide boxes cost the TMA unit more than
  // 256-bit LSU stores) on the per-lane path, which therefore keeps it.  gemm_tma_store = 2 forces TMA stores everywhere, 0 off.
  const long long ts_mode = tunable(TUNE_GEMM_TMA_STORE, 1);
  const bool ts_want = ts_mode == 2 || (ts_mode == 1 && (!e->out_is_bf16 || (e->out_pre && e->act)));
  if (ep.vec32 && !ep.atomic_out && !e->window && !e->rowsum && ts_want) {
    if (int rc2 = make_tmap_2d(&to, e->out, e->out_is_bf16 ? 2 : 4, N, M, e->ld_out, 16, 32, e->out_is_bf16 ? 32 : 64)) return rc2;
    ep.tma_out = 1;
    if (e->out_pre && e->act && e->out_is_bf16) {
      if (int rc2 = make_tmap_2d(&tp, e->out_pre, 2, N, M, e->ld_pre, 16, 32, 32)) return rc2;
      ep.tma_out |= 2;
    }
  }
  if (tunable(TUNE_GEMM_SPEC, 1) && ep.vec32 && !ep.atomic_out && !e->rowsum) {
    const bool plain = !e->window && !e->residual && !e->row_scale;
    if (plain && e->out_is_bf16 && e->scale_cols <= 0 && ep.tma_out == 3 && e->bias && e->act == 1 && e->out_pre && !e->gelu_pre) ep.spec = 1;
    else if (plain && e->out_is_bf16 && e->scale_cols <= 0 && ep.tma_out == 0 && !e->bias && !e->act && !e->out_pre && e->gelu_pre) ep.spec = 2;
    else if (plain && e->out_is_bf16 && ep.tma_out == 0 && e->bias && !e->act && !e->out_pre && !e->gelu_pre) ep.spec = 3;
    else if (plain && e->out_is_bf16 && e->scale_cols <= 0 && ep.tma_out == 0 && !e->bias && !e->act && !e->out_pre && !e->gelu_pre) ep.spec = 4;
    else if (!e->out_is_bf16 && e->bias && e->residual && !e->act && !e->out_pre && !e->gelu_pre && e->scale_cols <= 0) {
      if (!e->window && ep.tma_out == 1) ep.spec = 5;
      else if (e->window && ep.tma_out == 0) ep.spec = 6;
    }
  }
  if (ep.spec == 2 && bn256 && !a_mn_major && b_mn_major && tunable(TUNE_GEMM_BOX, 1)) {
    // fc2 dgrad on 128 x 256 tiles: pre-activation rows in / results out as 32-row x 64-column TMA boxes (SWIZZLE_128B)
    if (int rc2 = make_tmap_2d(&to, e->out, 2, N, M, e->ld_out, 64, 32, 128)) return rc2;
    if (int rc2 = make_tmap_2d(&tp, e->gelu_pre, 2, N, M, e->ld_gelu_pre, 64, 32, 128)) return rc2;
    ep.spec = 7;
  } else if ((ep.spec == 3 || ep.spec == 4) && bn256 && !a_mn_major && (ep.spec == 3 ? !b_mn_major : b_mn_major) &&
             tunable(TUNE_GEMM_BOX, 1) >= 2) {
    // bias-only (qkv) / plain (dgrad) epilogues on 128 x 256 tiles: results leave as 32-row x 64-column TMA boxes
    if (int rc2 = make_tmap_2d(&to, e->out, 2, N, M, e->ld_out, 64, 32, 128)) return rc2;
    ep.spec = ep.spec == 3 ? 8 : 9;
  }
  if (e->rowsum) return bn256 ? dispatch_gemm_rowsum<256>(a_mn_major, b_mn_major, ta, tb, to, tp, M, N, K, k_splits, ep, stream)
                              : dispatch_gemm_rowsum<128>(a_mn_major, b_mn_major, ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  if (bn256) return dispatch_gemm<256>(a_mn_major, b_mn_major, ta, tb, to, tp, M, N, K, k_splits, ep, stream);
  return dispatch_gemm<128>(a_mn_major, b_mn_major, ta, tb, to, tp, M, N, K, k_splits, ep, stream);
}
