// Development probe: tcgen05.mma issue / execution rates of one SM-resident CTA under different per-stage protocols.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../fullbatchtraining_b200/csrc -o /tmp/mi mma_issue.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "fb_common.cuh"
using namespace fb;

// mode bits: 1 = commit to a barrier after every group, 2 = also try_wait on an already completed barrier per group,
//            4 = tcgen05.fence::after per group, 8 = producer-style handshake with a second warp (full/empty ring)
template <int N, int GROUP, int PAD = 0>
__global__ void __launch_bounds__(192, 1) k_issue(long long* out, int groups, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bars[20];
  uint64_t* done_bar = &bars[0];
  uint64_t* dummy_bar = &bars[1];
  uint64_t* ready_bar = &bars[2];
  uint64_t* full_bar = &bars[4];
  uint64_t* empty_bar = &bars[12];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 20; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (threadIdx.x == 0) mbar_arrive(ready_bar);  // completes phase 0 of ready_bar: try_wait(ready, 0) succeeds at once
  __syncthreads();
  constexpr int STAGES = 4;
  if (warp == 0 && (mode & 8)) {
    int s = 0;
    uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      mbar_wait(&empty_bar[s], phase ^ 1, 1);
      if (elect_one()) mbar_arrive(&full_bar[s]);
      __syncwarp();
      if (++s == STAGES) { s = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t smem0 = smem_u32(smem);
    const long long t0 = clock64();
    int s = 0;
    uint32_t phase = 0;
    for (int g = 0; g < groups; ++g) {
      if (mode & 8) mbar_wait(&full_bar[s], phase, 2);
      if (mode & 2) mbar_wait(ready_bar, 0, 3);
      if (mode & 4) tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = smem_desc_lo(smem0 + s * 49152, 16);
        const uint32_t b_lo = a_lo + (32768 >> 4);
#pragma unroll
        for (int i = 0; i < GROUP; ++i) {
          tc_mma_bf16_lohi(tmem_base, a_lo + ((((i / 4) % 2) * 16384 + (i % 4) * 32) >> 4), b_lo + (((i % 4) * 32) >> 4),
                           hi, hi, idesc, 1u);
          if (PAD > 0) {  // PAD dependent integer instructions between two MMAs (does the pipe run ahead of them?)
            uint32_t x = i;
#pragma unroll
            for (int j = 0; j < PAD; ++j) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(tmem_base));
            if (x == 0x12345678u) out[7] = x;
          }
        }
        if (mode & 8) tc_commit(&empty_bar[s]);
        else if (mode & 1) tc_commit(dummy_bar);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; phase ^= 1; }
    }
    const long long t1 = clock64();
    if (elect_one()) tc_commit(done_bar);
    __syncwarp();
    mbar_wait(done_bar, 0, 9);
    const long long t2 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Two issuing warps alternate groups ("stages"); a turn token (mbarrier per warp) keeps the MMA order fixed.  Each warp
// does its per-group overhead (waits, elect, descriptor setup) while the OTHER warp's MMAs run.
template <int N, int GROUP>
__global__ void __launch_bounds__(192, 1) k_issue2(long long* out, int groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bars[8];
  uint64_t* done_bar = &bars[0];
  uint64_t* turn = &bars[2];  // turn[w]: completed once per group handed TO warp w
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 1 || warp == 2) {
    const int me = warp - 1;
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = smem_desc_hi_sw128(1024);
    const uint32_t smem0 = smem_u32(smem);
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int g = me; g < groups; g += 2) {
      const int s = g & 3;
      const uint32_t a_lo = smem_desc_lo(smem0 + s * 49152, 16);
      const uint32_t b_lo = a_lo + (32768 >> 4);
      if (g > 0) mbar_wait(&turn[me], phase, 5);  // previous group (other warp) has been issued
      if (g > 0 || me == 1) phase ^= (g > 0);
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < GROUP; ++i)
          tc_mma_bf16_lohi(tmem_base, a_lo + ((((i / 4) % 2) * 16384 + (i % 4) * 32) >> 4), b_lo + (((i % 4) * 32) >> 4),
                           hi, hi, idesc, 1u);
        mbar_arrive(&turn[me ^ 1]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (me == ((groups - 1) & 1)) {
      if (elect_one()) tc_commit(done_bar);
      __syncwarp();
      mbar_wait(done_bar, 0, 9);
      const long long t2 = clock64();
      if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int N, int GROUP>
void run2(int total_mmas, long long* d_out) {
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k_issue2<N, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    k_issue2<N, GROUP><<<148, 192, smem>>>(d_out, total_mmas / GROUP);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("two issuing warps N=%d group=%2d: issue %7.1f clk/MMA, complete %7.1f clk/MMA  (%s)\n", N, GROUP,
         double(h[0]) / total_mmas, double(h[1]) / total_mmas, cudaGetErrorString(cudaGetLastError()));
}

template <int N, int GROUP, int PAD = 0>
void run(const char* name, int total_mmas, long long* d_out) {
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k_issue<N, GROUP, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = total_mmas / GROUP;
  for (int mode : {0, 8}) {
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
      k_issue<N, GROUP, PAD><<<148, 192, smem>>>(d_out, groups, mode);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s N=%d group=%2d pad=%2d mode=%2d: issue %7.1f clk/MMA, complete %7.1f clk/MMA  (%s)\n", name, N, GROUP, PAD,
           mode, double(h[0]) / total_mmas, double(h[1]) / total_mmas, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<128, 4>("bf16", 2048, d_out);
  run<128, 8>("bf16", 2048, d_out);
  run<128, 16>("bf16", 2048, d_out);
  run<128, 32>("bf16", 2048, d_out);
  run2<128, 4>(2048, d_out);
  run2<128, 8>(2048, d_out);
  run2<128, 16>(2048, d_out);
  run<64, 8>("bf16", 2048, d_out);
  run<128, 8, 4>("bf16", 2048, d_out);
  run<128, 8, 8>("bf16", 2048, d_out);
  run<128, 8, 16>("bf16", 2048, d_out);
  run<128, 8, 32>("bf16", 2048, d_out);

  return 0;
}
