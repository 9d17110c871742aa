// Shared device helpers for the clover_b200 kernels (sm_100a only).
// PTX wrappers for mbarrier / TMA / tcgen05, warp reductions, bf16 packing and the closed-form
// window index maps of the Video Swin backbone (reference
// mmaction/models/backbones/swin_transformer_3d.py:271-299,460,474).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CLV_DEVICE __device__ __forceinline__

namespace clv {

// ------------------------------------------------------------------------------------------
// error slot (thread-local, read through clv_last_error())
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
// call right after every kernel launch: counts it (clv_launch_count) and reports launch errors
int after_launch(const char* what);

#define CLV_CHECK_CUDA(expr)                                   \
  do {                                                         \
    int _e = ::clv::check_cuda((expr), #expr);                 \
    if (_e) return _e;                                         \
  } while (0)

#define CLV_REQUIRE(cond, ...)                                 \
  do {                                                         \
    if (!(cond)) {                                             \
      ::clv::set_error(__VA_ARGS__);                           \
      return 1;                                                \
    }                                                          \
  } while (0)

int num_sms();                                            // of the current device (cached per device)
int ensure_dynamic_smem(const void* kernel, int bytes);   // cudaFuncSetAttribute once per (kernel, device)
enum Tunable { TUNE_GEMM_BN256_MIN_UNITS = 0, TUNE_W7_PIPE, TUNE_W7_BWD2, TUNE_W7_DBIAS_ACC, TUNE_W7_DBIAS_ACC_MIN_MB, TUNE_W7_FWD2, TUNE_GEMM_TMA_STORE, TUNE_W7_L2_HINT, TUNE_GEMM_SPEC, TUNE_W7_BWD_EARLY, TUNE_GEMM_BOX, TUNE_W7_FWD_EARLY, TUNE_W7_FWD_DBG, TUNE_W7_FWD_PVSPLIT, TUNE_W7_FWD_QTILE, TUNE_COUNT };
long long tunable(int id, long long dflt);                // clv_set_tunable overrides (tools only); no getenv in the library
// D[b,h,i] = <dO_i, O_i> (attention backward preparation), defined in attention.cu
int launch_attn_bwd_prep(const void* out, const void* dout, float* dsum, long long rows, int heads, int hd, int seq,
                         cudaStream_t stream);
// 2-D bf16 tensor map (inner contiguous dimension first).  swizzle_bytes: 128, 64, 32 or 0.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld, int box_inner,
                      int box_outer, int swizzle_bytes = 128);
// same for 2- or 4-byte elements (elem_bytes 2 = bf16, 4 = fp32); ld in elements
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, long long inner, long long outer, long long ld, int box_inner,
                 int box_outer, int swizzle_bytes);

// ------------------------------------------------------------------------------------------
// small math
// ------------------------------------------------------------------------------------------
CLV_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CLV_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Branch-free erf (Abramowitz & Stegun 7.1.26, |abs err| < 5e-7 incl. the MUFU approximations) that also
// returns exp(-u^2): exact-erf GELU (nn.GELU default / HF "gelu") and its derivative need both, and the epilogues
// that evaluate them are instruction-bound, so erff()'s range branches are avoided.
CLV_DEVICE void erf_exp(float u, float& erf_u, float& exp_mu2) {
  const float au = fabsf(u);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, au, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-au * au * 1.4426950408889634f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  erf_u = copysignf(fmaf(-p, e, 1.0f), u);
  exp_mu2 = e;
}
CLV_DEVICE float gelu_erf(float x) {
  float er, e;
  erf_exp(x * 0.70710678118654752440f, er, e);
  return 0.5f * x * (1.0f + er);
}
// d/dx [0.5 x (1+erf(x/sqrt2))] = 0.5(1+erf(x/sqrt2)) + x * exp(-x^2/2)/sqrt(2pi)
CLV_DEVICE float gelu_erf_grad(float x) {
  float er, e;
  erf_exp(x * 0.70710678118654752440f, er, e);
  return fmaf(x * 0.39894228040143267794f, e, 0.5f * (1.0f + er));
}
// GEMM-epilogue GELU: Phi(x) = 0.5 (1 + tanh(P(x))) with the odd polynomial P fitted to atanh(erf(x / sqrt 2)) on
// |x| <= 6 (|gelu - exact erf GELU| < 2.6e-5 before the MUFU.TANH approximation, whose 2^-11 relative error stays
// below the bf16 rounding of the stored activation).  x^2 is clamped at 36 so that P stays monotone: beyond |x| = 6
// P = 1.68 x and tanh saturates, as erf does.  9 issued instructions per element against 17 for the A&S erf above:
// the fc1 / fc2-dgrad epilogues (K = 128..512) are issue-bound, not tensor-bound.
constexpr float GELU_A0 = 7.97507884e-01f, GELU_A1 = 3.70056460e-02f, GELU_A2 = -3.51516788e-04f;
CLV_DEVICE float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
CLV_DEVICE float gelu_fit(float x) {
  const float x2 = fminf(x * x, 36.0f);
  const float t = fmaf(x2, fmaf(x2, GELU_A2, GELU_A1), GELU_A0);
  const float th = tanh_approx(x * t);
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
// d/dx of the SAME fitted function gelu_fit(x) = 0.5 x (1 + tanh(u)), u = x P(x^2):
//   0.5 (1 + tanh u) + 0.5 x (1 - tanh^2 u) u'(x),   u'(x) = A0 + 3 A1 x^2 + 5 A2 x^4   (0 beyond the |x| = 6 clamp, where
// 1 - tanh^2 u underflows anyway).  One MUFU (tanh) instead of two (tanh + ex2 for the exact phi(x)): the fc2-dgrad epilogue
// that evaluates it on 2 x 10^8 elements per launch is XU / issue-bound.  |error| vs the exact erf-GELU derivative < 4e-4
// (tests/test_kernels_gpu.py::test_gemm_epilogues), below the bf16 rounding of the gradient it multiplies.
CLV_DEVICE float gelu_fit_grad(float x) {
  const float x2 = fminf(x * x, 36.0f);
  const float t = fmaf(x2, fmaf(x2, GELU_A2, GELU_A1), GELU_A0);
  const float du = fmaf(x2, fmaf(x2, 5.0f * GELU_A2, 3.0f * GELU_A1), GELU_A0);
  const float th = tanh_approx(x * t);
  const float hs = fmaf(-th, th, 1.0f) * (0.5f * x);          // 0.5 x sech^2(u)
  return fmaf(hs, du, fmaf(0.5f, th, 0.5f));
}
// ------------------------------------------------------------------------------------------
// counter-based random stream of the dropout kernels: 32 bits = high word of splitmix64(seed * phi + idx).
// keep iff bits >= p * 2^32.  Restated in numpy by tests/rng_ref.py (bit-exact check).
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rand_u32(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t <= 0.0 ? 0u : (t >= 4294967295.0 ? 4294967295u : (uint32_t)t);
}
CLV_DEVICE float keep_scale(unsigned long long seed, unsigned long long idx, uint32_t thresh, float inv_keep) {
  return rand_u32(seed, idx) >= thresh ? inv_keep : 0.f;
}
CLV_DEVICE uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
CLV_DEVICE float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ------------------------------------------------------------------------------------------
// window geometry: fused cyclic roll + window_partition and its inverse, with zero padding.
// Frame (Dp,Hp,Wp) is the padded extent; (D,H,W) the real one.
// ------------------------------------------------------------------------------------------
struct WindowGeom {
  int B, D, H, W;        // real extent
  int Dp, Hp, Wp;        // padded to window multiples
  int wd, wh, ww;        // clamped window
  int sd, sh, sw;        // clamped shift
  int nD, nH, nW;        // windows per axis
  int N;                 // tokens per window
  int nWin;              // windows per clip
};

// window-order row r -> source spatial row (or -1 when the token lies in the zero padding).
// Row counts fit 32 bits (checked on the host): 64-bit div/mod would cost more than the rest of a row.
CLV_DEVICE long long window_row_to_src(const WindowGeom& g, long long r64) {
  const unsigned r = (unsigned)r64;
  const unsigned n = r % (unsigned)g.N, gi = r / (unsigned)g.N;
  const unsigned win = gi % (unsigned)g.nWin, b = gi / (unsigned)g.nWin;
  const int bw = win % g.nW, bh = (win / g.nW) % g.nH, bd = win / (g.nW * g.nH);
  const int lw = n % g.ww, lh = (n / g.ww) % g.wh, ld = n / (g.ww * g.wh);
  int d = bd * g.wd + ld + g.sd; if (d >= g.Dp) d -= g.Dp;
  int h = bh * g.wh + lh + g.sh; if (h >= g.Hp) h -= g.Hp;
  int w = bw * g.ww + lw + g.sw; if (w >= g.Wp) w -= g.Wp;
  if (d >= g.D || h >= g.H || w >= g.W) return -1;
  return (long long)(((b * (unsigned)g.D + d) * (unsigned)g.H + h) * (unsigned)g.W + w);
}

// source spatial row s -> window-order row (always valid: every real token lives in one window).
CLV_DEVICE long long src_row_to_window(const WindowGeom& g, long long s64) {
  const unsigned s = (unsigned)s64;
  const int w = s % (unsigned)g.W; unsigned t = s / (unsigned)g.W;
  const int h = t % (unsigned)g.H; t /= (unsigned)g.H;
  const int d = t % (unsigned)g.D; const unsigned b = t / (unsigned)g.D;
  int d2 = d - g.sd; if (d2 < 0) d2 += g.Dp;
  int h2 = h - g.sh; if (h2 < 0) h2 += g.Hp;
  int w2 = w - g.sw; if (w2 < 0) w2 += g.Wp;
  const int bd = d2 / g.wd, ld = d2 % g.wd;
  const int bh = h2 / g.wh, lh = h2 % g.wh;
  const int bw = w2 / g.ww, lw = w2 % g.ww;
  const unsigned win = ((b * g.nD + bd) * g.nH + bh) * g.nW + bw;
  return (long long)(win * (unsigned)g.N + (ld * g.wh + lh) * g.ww + lw);
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers
// ------------------------------------------------------------------------------------------
CLV_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

CLV_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CLV_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CLV_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
CLV_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CLV_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CLV_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes or `hint_ns` elapse, so a waiting
// role (TMA producer, MMA issuer, softmax / epilogue warps between tiles) costs no issue slots -- the un-hinted form returns
// after ~100 cycles and the retry loop around it was a third of all instructions issued by the attention kernels (ncu r01c).
CLV_DEVICE bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (sticky launch error) instead of hanging the GPU.
CLV_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if ((++n & 0x3F) == 0 && clock64() - t0 > 8000000000LL) __trap();
  }
}

CLV_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled TMA load global -> shared, completion on mbarrier (bytes).
CLV_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- TMA store (shared -> global), bulk-group completion ----------------------------------------------------------
CLV_DEVICE void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// shared -> global with an element-wise add performed at the destination (L2): bf16 tiles of several units accumulate
// into one small buffer instead of being written out one by one
CLV_DEVICE void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
// L2 eviction-priority policies for the bulk copies: streamed operands leave first, accumulation buffers stay
CLV_DEVICE uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
CLV_DEVICE uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
CLV_DEVICE void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
CLV_DEVICE void tma_reduce_add_2d_hint(const CUtensorMap* map, const void* smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
CLV_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
CLV_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
CLV_DEVICE void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }   // all but the newest group
CLV_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

CLV_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CLV_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Warp-collective.  Writes the TMEM base address to *smem_slot.
CLV_DEVICE void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
CLV_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues.
CLV_DEVICE void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier when all previously issued tcgen05.mma of this thread have completed.
CLV_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread = lane/row).
CLV_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
CLV_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B).
CLV_DEVICE uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 inputs and fp32 accumulation
// (cute::UMMA::InstrDescriptor): c_format[4,6)=1, a_format[7,10)=1, b_format[10,13)=1,
// a_major bit15, b_major bit16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Generic descriptor: layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B, 0 = none.
CLV_DEVICE uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// D[tmem] (+)= A[tmem] * B[smem desc]  (A: 128 lanes x K bf16 packed two per 32-bit column)
CLV_DEVICE void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

CLV_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
CLV_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
CLV_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// named barrier among a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
CLV_DEVICE void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

CLV_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace clv
