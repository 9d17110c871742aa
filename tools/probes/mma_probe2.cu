// Timing probe 2: is the ~98-cycle floor of a small tcgen05.mma (M=128, K=16, N<=128) (a) the latency of a DEPENDENT accumulation
// chain, (b) a per-instruction cost of the tensor pipe, or (c) the issue rate of ONE thread?
//   lean issue loop, `nacc` independent accumulators cycled round-robin, `nwarps` issuing warps (each its own accumulators).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../clover_b200/csrc -I../../include mma_probe2.cu -o mma_probe2 -lcuda
#include <cstdio>
#include "common.cuh"
using namespace clv;

template <int NACC>
__global__ void probe(int N, int reps, int a_tmem, int nwarps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) mbar_init(bar + w, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < nwarps) {
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t da = make_smem_desc(a_addr, 16, 512, 4), db = make_smem_desc(b_addr, 16, 512, 4);
    // accumulators of warp w: columns [w*128 + k*32) (N <= 32) -- only N <= 32 is probed with several accumulators
    const uint32_t d0 = tb + warp * 128;
    for (int round = 0; round < 2; ++round) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t d = d0 + (k % NACC) * 32;
          if (a_tmem) umma_bf16_ts(d, tb + 384 + k * 8, db + k * 64, idesc, 1);
          else umma_bf16_ss(d, da + k * 2, db + k * 64, idesc, 1);
        }
      }
      umma_commit(bar + warp);
      mbar_wait(bar + warp, round & 1);
      out[warp * 2 + round] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int NACC>
void run(int N, int a_tmem, int nwarps, long long* d) {
  const int reps = 512;
  cudaFuncSetAttribute(probe<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  probe<NACC><<<1, 128, 70000>>>(N, reps, a_tmem, nwarps, d);
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  long long worst = 0;
  for (int w = 0; w < nwarps; ++w) worst = h[w * 2 + 1] > worst ? h[w * 2 + 1] : worst;
  printf("N=%3d A=%s nacc=%d issuing warps=%d : %6.1f clk per MMA per warp, %6.1f clk per MMA overall (%s)\n", N, a_tmem ? "tmem" : "smem",
         NACC, nwarps, (double)worst / reps, (double)worst / (reps * nwarps), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int N : {16, 32})
      for (int nw : {1, 2, 4}) {
        run<1>(N, a_tmem, nw, d); run<2>(N, a_tmem, nw, d); run<4>(N, a_tmem, nw, d);
      }
  for (int N : {64, 96, 128}) { run<1>(N, 0, 1, d); run<1>(N, 0, 2, d); run<1>(N, 1, 1, d); }
  return 0;
}
