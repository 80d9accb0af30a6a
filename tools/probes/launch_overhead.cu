// Development probe: per-launch cost of kernels as a function of dynamic shared memory, TMEM allocation and of
// alternating shared-memory carve-outs between consecutive launches.   nvcc -arch=sm_100a -o /tmp/lo launch_overhead.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void k_small(float* p) {
  if (p && threadIdx.x == 0 && blockIdx.x == 1000000) p[0] = 1.f;
}
__global__ void __launch_bounds__(192, 1) k_big(float* p) {
  extern __shared__ uint8_t smem[];
  if (p && threadIdx.x == 0 && blockIdx.x == 1000000) p[0] = smem[0];
}
__global__ void __launch_bounds__(192, 1) k_big_tmem(float* p, int cols) {
  extern __shared__ uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 1) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  if (p && threadIdx.x == 0 && blockIdx.x == 1000000) p[0] = smem[0];
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// medium kernel: 296 blocks x 512 threads, a little static smem (like the BatchNorm kernels)
__global__ void __launch_bounds__(512, 2) k_bn_like(float* p) {
  __shared__ float s[2048];
  s[threadIdx.x] = 0.f;
  __syncthreads();
  if (p && threadIdx.x == 0 && blockIdx.x == 1000000) p[0] = s[1];
}

// PDL variants: griddepcontrol.launch_dependents early, griddepcontrol.wait before touching memory
__global__ void __launch_bounds__(512, 2) k_bn_like_pdl(float* p, int spin) {
  __shared__ float s[2048];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  s[threadIdx.x] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (p && threadIdx.x == 0 && blockIdx.x == 1000000) p[0] = s[1];
}

template <class F>
float time_us(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 20; ++i) f(i);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f(i);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / reps;
}

int main() {
  const int big = 220 * 1024;
  cudaFuncSetAttribute(k_big, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_big_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  const int reps = 2000;
  printf("small only               %.2f us/launch\n", time_us([&](int) { k_small<<<148, 192>>>(nullptr); }, reps));
  printf("bn-like only             %.2f us/launch\n", time_us([&](int) { k_bn_like<<<296, 512>>>(nullptr); }, reps));
  printf("big smem only            %.2f us/launch\n", time_us([&](int) { k_big<<<148, 192, big>>>(nullptr); }, reps));
  printf("big smem + tmem 512      %.2f us/launch\n",
         time_us([&](int) { k_big_tmem<<<148, 192, big>>>(nullptr, 512); }, reps));
  printf("big smem + tmem 256      %.2f us/launch\n",
         time_us([&](int) { k_big_tmem<<<148, 192, big>>>(nullptr, 256); }, reps));
  printf("alternate big / small    %.2f us/launch\n", time_us([&](int i) {
           if (i & 1) k_small<<<148, 192>>>(nullptr); else k_big<<<148, 192, big>>>(nullptr);
         }, reps));
  printf("alternate big+tmem / bn  %.2f us/launch\n", time_us([&](int i) {
           if (i & 1) k_bn_like<<<296, 512>>>(nullptr); else k_big_tmem<<<148, 192, big>>>(nullptr, 512);
         }, reps));
  // same alternation, but the small kernels ask for the maximum shared-memory carve-out too
  cudaFuncSetAttribute(k_small, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_bn_like, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  printf("alternate big / small (carve-out 100)    %.2f us/launch\n", time_us([&](int i) {
           if (i & 1) k_small<<<148, 192>>>(nullptr); else k_big<<<148, 192, big>>>(nullptr);
         }, reps));
  printf("alternate big+tmem / bn (carve-out 100)  %.2f us/launch\n", time_us([&](int i) {
           if (i & 1) k_bn_like<<<296, 512>>>(nullptr); else k_big_tmem<<<148, 192, big>>>(nullptr, 512);
         }, reps));
  printf("bn-like only (carve-out 100)             %.2f us/launch\n",
         time_us([&](int) { k_bn_like<<<296, 512>>>(nullptr); }, reps));
  // in a CUDA graph (as the engine runs): 100 alternating nodes
  for (int variant = 0; variant < 2; ++variant) {
    cudaFuncSetAttribute(k_bn_like, cudaFuncAttributePreferredSharedMemoryCarveout, variant ? 100 : -1);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 100; ++i) {
      if (i & 1) k_bn_like<<<296, 512, 0, st>>>(nullptr); else k_big_tmem<<<148, 192, big, st>>>(nullptr, 512);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    float us = time_us([&](int) { cudaGraphLaunch(ge, st); }, 50) / 100;
    cudaStreamSynchronize(st);
    printf("graph of 100 alternating nodes, bn carve-out %s: %.2f us/node\n", variant ? "100" : "default", us);
  }
  // programmatic dependent launch inside a graph: 100 kernels of ~5 us each, with and without the attribute
  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 100; ++i) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(296);
      cfg.blockDim = dim3(512);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, k_bn_like_pdl, (float*)nullptr, 10000);
    }
    cudaStreamEndCapture(st, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    float us = time_us([&](int) { cudaGraphLaunch(ge, st); }, 50) / 100;
    cudaStreamSynchronize(st);
    printf("graph of 100 x ~5us kernels, PDL %d: %.2f us/node (instantiate: %s)\n", pdl, us, cudaGetErrorString(e));
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
