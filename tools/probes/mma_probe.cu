// Timing probe: cycles per tcgen05.mma (kind::f16, bf16, M=128, K=16) as a function of N and of the A-operand source.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../clover_b200/csrc -I../../include mma_probe.cu -o mma_probe -lcuda
#include <cstdio>
#include "common.cuh"
using namespace clv;

__global__ void probe(int N, int reps, int a_tmem, int M, long long* out, int nacc = 1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
    const uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
    for (int round = 0; round < 2; ++round) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t dcol = (r % nacc) * 64;
        if (a_tmem) umma_bf16_ts(tb + dcol, tb + 256, make_smem_desc(b_addr, 16, 512, 4), idesc, r >= nacc);
        else umma_bf16_ss(tb + dcol, make_smem_desc(a_addr, 16, 512, 4), make_smem_desc(b_addr, 16, 512, 4), idesc, r >= nacc);
      }
      umma_commit(bar);
      mbar_wait(bar, round & 1);
      const long long t1 = clock64();
      out[round] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

__global__ void probe_lean(int N, int reps, int a_tmem, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t da = make_smem_desc(a_addr, 16, 512, 4), db = make_smem_desc(b_addr, 16, 512, 4);
    for (int round = 0; round < 2; ++round) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (a_tmem) umma_bf16_ts(tb, tb + 256 + k * 8, db + k * 64, idesc, 1);
          else umma_bf16_ss(tb, da + k * 2, db + k * 64, idesc, 1);
        }
      }
      umma_commit(bar);
      mbar_wait(bar, round & 1);
      out[round] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int reps = 256;
  for (int M : {128, 64})
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
      for (int N : {16, 32, 64, 96, 128, 208, 256}) {
        probe<<<1, 128, 70000>>>(N, reps, a_tmem, M, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("M=%3d A=%s N=%3d  %6.1f clk/MMA (%s)\n", M, a_tmem ? "tmem" : "smem", N, (double)h[1] / reps, cudaGetErrorString(e));
      }
  for (int nacc : {1, 2, 4})
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem) {
      probe<<<1, 128, 70000>>>(32, reps, a_tmem, 128, d, nacc);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("independent accumulators: nacc=%d A=%s N=32  %6.1f clk/MMA\n", nacc, a_tmem ? "tmem" : "smem", (double)h[1] / reps);
    }
  cudaFuncSetAttribute(probe_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int N : {32, 64, 128, 208, 256}) {
      probe_lean<<<1, 128, 70000>>>(N, reps, a_tmem, d);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("lean issue: A=%s N=%3d  %6.1f clk/MMA\n", a_tmem ? "tmem" : "smem", N, (double)h[1] / reps);
    }
  // 1 rep latency
  probe<<<1, 128, 70000>>>(208, 1, 0, 128, d);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("single MMA N=208 issue->commit->wait latency: %lld clk\n", h[1]);
  probe<<<1, 128, 70000>>>(32, 1, 1, 128, d);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("single MMA N=32 (A tmem) latency: %lld clk\n", h[1]);
  return 0;
}
