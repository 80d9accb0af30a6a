// Bandwidth-bound layer kernels of the full-batch step (sm_100a): BatchNorm (train mode) statistics / apply / backward
// fused with ReLU and the residual add, AvgPool2d(2), stem im2col, and the pooled-linear-cross-entropy head.
// All of them are coalesced, 16-byte vectorised streaming kernels; reductions are two-stage and deterministic
// (fixed partition, fixed summation order), so repeated runs are bit-identical
// (cf. measure_floating_point_accuracy.py / fullbatch/training/training.py:429-600).
//
// Reference call sites replaced: torch.nn.BatchNorm2d / ReLU(inplace) / `out += identity`
// (fullbatch/models/resnets.py:71,207-230,296-316), AvgPool2d (resnets.py:149), AdaptiveAvgPool2d + Linear
// (resnets.py:106-107), LabelSmoothCrossEntropyLoss (fullbatch/models/modules.py:96-101), accuracy count
// (fullbatch/training/training.py:80) and their autograd backward.
#include <cooperative_groups.h>

#include "../../include/fullbatch_b200.h"
#include "fb_common.cuh"

namespace fb {

constexpr int kMaxChunks = 296;  // 2 reduction blocks per SM
typedef __nv_bfloat16 bf16;

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load8_bf16(const bf16* p, float (&v)[8]) {
  const bf16x8 t = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(t.v[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// split 8 fp32 values into hi/lo bf16 and store (lo optional)
__device__ __forceinline__ void store8_split(bf16* hi, bf16* lo, long long off, const float (&v)[8]) {
  bf16x8 h, l;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bf16 h0, l0, h1, l1;
    split_bf16(v[2 * i], h0, l0);
    split_bf16(v[2 * i + 1], h1, l1);
    h.v[i] = __halves2bfloat162(h0, h1);
    l.v[i] = __halves2bfloat162(l0, l1);
  }
  *reinterpret_cast<bf16x8*>(hi + off) = h;
  if (lo) *reinterpret_cast<bf16x8*>(lo + off) = l;
}
__device__ __forceinline__ void store8_bf16(bf16* dst, long long off, const float (&v)[8]) {
  bf16x8 h;
#pragma unroll
  for (int i = 0; i < 4; ++i) h.v[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<bf16x8*>(dst + off) = h;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-channel column reductions over a [P][C] fp32 matrix: shared skeleton for BN statistics and BN backward.
// Block = 256 threads = TX float4-columns x TY rows; grid = (chunks, column slabs) -> partial[chunk][2][C].
// The LAST block to finish (atomic ticket) reduces the partials in a fixed order and finalises, so the whole reduction
// is one launch and still deterministic.
// ---------------------------------------------------------------------------------------------------------------
struct BnFinalize {
  long long P;
  int chunks;
  // forward (statistics)
  float *mean_out, *rstd_out, *running_mean, *running_var;
  float momentum, eps;
  // backward
  float *coef, *dgamma, *dbeta;
};

// Second stage: block = 8 channels x 32 lanes; lane l sums chunks l, l+32, ... (independent loads in flight), lanes
// are combined by a fixed shuffle tree -> deterministic.
template <bool BWD>
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int C, BnFinalize f) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int k = lane; k < f.chunks; k += 32) {
      s1 += partial[(long long)k * 2 * C + c];
      s2 += partial[(long long)k * 2 * C + C + c];
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane != 0 || c >= C) return;
  if (!BWD) {
    const double m = s1 / double(f.P);
    double var = s2 / double(f.P) - m * m;
    var = var < 0.0 ? 0.0 : var;
    f.mean_out[c] = float(m);
    f.rstd_out[c] = float(1.0 / sqrt(var + double(f.eps)));
    if (f.running_mean) {
      const double unbiased = f.P > 1 ? var * double(f.P) / double(f.P - 1) : var;
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * float(m);
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * float(unbiased);
    }
  } else {
    f.dbeta[c] = float(s1);
    f.dgamma[c] = float(s2);
    f.coef[c] = float(s1 / double(f.P));
    f.coef[C + c] = float(s2 / double(f.P));
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ y, const float* __restrict__ dA,
                                                        const float* __restrict__ dA2,
                                                        const bf16* __restrict__ mask, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, long long P, int C, int TX,
                                                        int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float4 red[2][256];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const int c = (blockIdx.y * TX + tx) * 4;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(P, r0 + rows_per_chunk);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  float4 mu = s1, rs = s1;
  if (BWD) {
    mu = *reinterpret_cast<const float4*>(mean + c);
    rs = *reinterpret_cast<const float4*>(rstd + c);
  }
#pragma unroll 8
  for (long long r = r0 + ty; r < r1; r += TY) {
    const float4 v = *reinterpret_cast<const float4*>(y + r * C + c);
    if (!BWD) {
      s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
      s2.x += v.x * v.x; s2.y += v.y * v.y; s2.z += v.z * v.z; s2.w += v.w * v.w;
    } else {
      float4 d = *reinterpret_cast<const float4*>(dA + r * C + c);
      if (dA2) {
        const float4 d2 = *reinterpret_cast<const float4*>(dA2 + r * C + c);
        d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
      }
      if (mask) {
        const uint2 m = *reinterpret_cast<const uint2*>(mask + r * C + c);
        const float2 m01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.x));
        const float2 m23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.y));
        d.x = m01.x > 0.f ? d.x : 0.f;
        d.y = m01.y > 0.f ? d.y : 0.f;
        d.z = m23.x > 0.f ? d.z : 0.f;
        d.w = m23.y > 0.f ? d.w : 0.f;
      }
      s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
      s2.x += d.x * (v.x - mu.x) * rs.x;
      s2.y += d.y * (v.y - mu.y) * rs.y;
      s2.z += d.z * (v.z - mu.z) * rs.z;
      s2.w += d.w * (v.w - mu.w) * rs.w;
    }
  }
  red[0][threadIdx.x] = s1;
  red[1][threadIdx.x] = s2;
  __syncthreads();
  if (ty == 0) {
    for (int j = 1; j < TY; ++j) {
      const float4 a = red[0][j * TX + tx], b = red[1][j * TX + tx];
      s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
      s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
    }
    float* dst = partial + (long long)blockIdx.x * 2 * C;
    *reinterpret_cast<float4*>(dst + c) = s1;
    *reinterpret_cast<float4*>(dst + C + c) = s2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BN apply (+ second normalised branch, + residual, + ReLU) -> bf16 hi/lo
// ---------------------------------------------------------------------------------------------------------------
// Per-thread channel parameters are loop invariant: the grid stride (gridDim.x * 2048 elements) is a multiple of C for
// every channel count of the ResNet family (C | 2048), so they are loaded once per thread.
struct BnAffine {
  float mu[8], scale[8], shift[8];
  __device__ __forceinline__ void load(const float* mean, const float* rstd, const float* gamma, const float* beta,
                                       int c) {
    float rs[8], ga[8];
    load8(mean + c, mu);
    load8(rstd + c, rs);
    load8(gamma + c, ga);
    load8(beta + c, shift);
#pragma unroll
    for (int j = 0; j < 8; ++j) scale[j] = rs[j] * ga[j];
  }
};

__global__ void __launch_bounds__(256) bn_apply_kernel(fb_bn_apply_args a) {
  const long long total8 = a.P * a.C / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool invariant = (stride * 8) % a.C == 0;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  BnAffine p1, p2;
  if (i < total8) {
    const int c = int((i * 8) % a.C);
    p1.load(a.mean, a.rstd, a.gamma, a.beta, c);
    if (a.y2) p2.load(a.mean2, a.rstd2, a.gamma2, a.beta2, c);
  }
  for (; i < total8; i += stride) {
    const long long off = i * 8;
    if (!invariant) {
      const int c = int(off % a.C);
      p1.load(a.mean, a.rstd, a.gamma, a.beta, c);
      if (a.y2) p2.load(a.mean2, a.rstd2, a.gamma2, a.beta2, c);
    }
    float y[8], o[8];
    load8(a.y + off, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (y[j] - p1.mu[j]) * p1.scale[j] + p1.shift[j];
    if (a.y2) {
      load8(a.y2 + off, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += (y[j] - p2.mu[j]) * p2.scale[j] + p2.shift[j];
    }
    if (a.res_hi) {
      float rh[8];
      load8_bf16(static_cast<const bf16*>(a.res_hi) + off, rh);
      if (a.res_lo) {
        float rl[8];
        load8_bf16(static_cast<const bf16*>(a.res_lo) + off, rl);
#pragma unroll
        for (int j = 0; j < 8; ++j) rh[j] += rl[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += rh[j];
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    store8_split(static_cast<bf16*>(a.out_hi), static_cast<bf16*>(a.out_lo), off, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BN backward apply: dy = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) -> bf16; optional dz (fp32) output
// ---------------------------------------------------------------------------------------------------------------
struct BnBwdCoef {
  float mu[8], rs[8], grs[8], c1[8], c2[8];
  __device__ __forceinline__ void load(const fb_bn_bwd_args& a, const float* coef, int c) {
    float ga[8];
    load8(a.mean + c, mu);
    load8(a.rstd + c, rs);
    load8(a.gamma + c, ga);
    load8(coef + c, c1);
    load8(coef + a.C + c, c2);
#pragma unroll
    for (int j = 0; j < 8; ++j) grs[j] = ga[j] * rs[j];
  }
};

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(fb_bn_bwd_args a, const float* __restrict__ coef) {
  const long long total8 = a.P * a.C / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool invariant = (stride * 8) % a.C == 0;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  BnBwdCoef p;
  if (i < total8) p.load(a, coef, int((i * 8) % a.C));
  for (; i < total8; i += stride) {
    const long long off = i * 8;
    if (!invariant) p.load(a, coef, int(off % a.C));
    float d[8], y[8], o[8];
    load8(a.dA + off, d);
    if (a.dA2) {
      float d2[8];
      load8(a.dA2 + off, d2);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] += d2[j];
    }
    if (a.mask_hi) {
      float m[8];
      load8_bf16(static_cast<const bf16*>(a.mask_hi) + off, m);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = m[j] > 0.f ? d[j] : 0.f;
    }
    load8(a.y + off, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xhat = (y[j] - p.mu[j]) * p.rs[j];
      o[j] = p.grs[j] * (d[j] - p.c1[j] - xhat * p.c2[j]);
    }
    store8_bf16(static_cast<bf16*>(a.dy_bf16), off, o);
    if (a.dz_out) {
      if (a.dz_accumulate) {
        float e[8];
        load8(a.dz_out + off, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += e[j];
      }
      store8(a.dz_out + off, d);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused BatchNorm kernels: statistics -> finalize -> apply in ONE persistent launch with two grid-wide barriers.
// All blocks are co-resident (grid <= 2 per SM, checked on the host), so a spin barrier on a global counter is safe; it
// is bounded and traps instead of hanging.  Every block applies to the same rows it reduced, so the second pass over
// Y / dA / mask is served by L2 (the largest ResNet-18 tensors are 33 MB, L2 is 126 MB) instead of HBM, and the two
// extra launches per BatchNorm disappear.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
      if (clock64() - t0 > 4000000000LL) {
        printf("[fb] grid barrier timeout: block %d target %u have %u\n", blockIdx.x, target, *counter);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}
// after the last barrier: the last block to leave resets the counters for the next launch
__device__ __forceinline__ void grid_barrier_release(unsigned int* counters) {
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counters + 1, 1u);
    if (prev == gridDim.x - 1) {
      counters[0] = 0u;
      counters[1] = 0u;
      __threadfence();
    }
  }
}

struct BnFusedFwdArgs {
  fb_bn_apply_args ap;           // y, gamma, beta, (y2, gamma2, beta2), residual, relu, P, C, outputs; mean/rstd = outputs
  float *mean, *rstd, *mean2, *rstd2;
  float *running_mean, *running_var, *running_mean2, *running_var2;
  float momentum, eps;
  float* partial;                // [grid][2 branches][2][C]
  unsigned int* counters;        // [2], zero on entry
  int rows_per_block;
  // statistics already reduced per CTA by the producing convolution's epilogue (fb_conv_gemm / fb_conv3x3 stats_out):
  // [ext_rows][2][C] per branch; when given, phase 1 and the first barrier are skipped
  const float *ext0, *ext1;
  int ext_rows0, ext_rows1;
};

// column sums of one branch over rows [r0, r1): partial[2][C] of this block
template <bool BWD>
__device__ __forceinline__ void block_column_sums(const float* __restrict__ y, const float* __restrict__ dA,
                                                  const float* __restrict__ dA2, const bf16* __restrict__ mask,
                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                  long long r0, long long r1, int C, float* __restrict__ out,
                                                  float4 (*red)[256]) {
  const int c4 = C / 4;
  const int TX = c4 < 256 ? c4 : 256;
  const int TY = 256 / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  for (int slab = 0; slab < c4 / TX; ++slab) {
    const int c = (slab * TX + tx) * 4;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, mu = s1, rs = s1;
    if (BWD) {
      mu = *reinterpret_cast<const float4*>(mean + c);
      rs = *reinterpret_cast<const float4*>(rstd + c);
    }
#pragma unroll 4
    for (long long r = r0 + ty; r < r1; r += TY) {
      const float4 v = *reinterpret_cast<const float4*>(y + r * C + c);
      if (!BWD) {
        s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
        s2.x += v.x * v.x; s2.y += v.y * v.y; s2.z += v.z * v.z; s2.w += v.w * v.w;
      } else {
        float4 d = *reinterpret_cast<const float4*>(dA + r * C + c);
        if (dA2) {
          const float4 d2 = *reinterpret_cast<const float4*>(dA2 + r * C + c);
          d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
        }
        if (mask) {
          const uint2 m = *reinterpret_cast<const uint2*>(mask + r * C + c);
          const float2 m01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.x));
          const float2 m23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.y));
          d.x = m01.x > 0.f ? d.x : 0.f;
          d.y = m01.y > 0.f ? d.y : 0.f;
          d.z = m23.x > 0.f ? d.z : 0.f;
          d.w = m23.y > 0.f ? d.w : 0.f;
        }
        s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
        s2.x += d.x * (v.x - mu.x) * rs.x;
        s2.y += d.y * (v.y - mu.y) * rs.y;
        s2.z += d.z * (v.z - mu.z) * rs.z;
        s2.w += d.w * (v.w - mu.w) * rs.w;
      }
    }
    __syncthreads();
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (ty == 0) {
      for (int j = 1; j < TY; ++j) {
        const float4 a = red[0][j * TX + tx], b = red[1][j * TX + tx];
        s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
        s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
      }
      *reinterpret_cast<float4*>(out + c) = s1;
      *reinterpret_cast<float4*>(out + C + c) = s2;
    }
  }
}

// fixed-order reduction of the per-block partials of channel c by one warp (lane = block index mod 32)
__device__ __forceinline__ void warp_reduce_partials(const float* __restrict__ partial, long long block_stride, int nblk,
                                                     int C, int c, int lane, double& s1, double& s2) {
  s1 = 0.0;
  s2 = 0.0;
#pragma unroll 4
  for (int k = lane; k < nblk; k += 32) {
    s1 += __ldcg(partial + (long long)k * block_stride + c);
    s2 += __ldcg(partial + (long long)k * block_stride + C + c);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
}

__global__ void __launch_bounds__(256, 2) bn_fwd_fused_kernel(BnFusedFwdArgs a) {
  griddep_wait();
  __shared__ float4 red[2][256];
  const fb_bn_apply_args& ap = a.ap;
  const int C = ap.C;
  const long long P = ap.P;
  const int branches = ap.y2 ? 2 : 1;
  const long long r0 = (long long)blockIdx.x * a.rows_per_block;
  const long long r1 = min(P, r0 + a.rows_per_block);
  const long long block_stride = (long long)branches * 2 * C;
  float* mine = a.partial + blockIdx.x * block_stride;
  const bool external = a.ext0 != nullptr;
  unsigned int barrier_target = gridDim.x;
  if (!external) {
    // ---- phase 1: per-block column sums
    block_column_sums<false>(ap.y, nullptr, nullptr, nullptr, nullptr, nullptr, r0, r1, C, mine, red);
    if (ap.y2)
      block_column_sums<false>(ap.y2, nullptr, nullptr, nullptr, nullptr, nullptr, r0, r1, C, mine + 2 * C, red);
    grid_barrier(a.counters, gridDim.x);
    barrier_target = 2 * gridDim.x;
  }
  // ---- finalize: one warp per channel, spread over the blocks
  {
    const int lane = threadIdx.x & 31;
    const int total = branches * C;
    for (int item = blockIdx.x * 8 + (threadIdx.x >> 5); item < total; item += gridDim.x * 8) {
      const int br = item / C, c = item % C;
      double s1, s2;
      if (external)
        warp_reduce_partials(br ? a.ext1 : a.ext0, 2LL * C, br ? a.ext_rows1 : a.ext_rows0, C, c, lane, s1, s2);
      else
        warp_reduce_partials(a.partial + br * 2 * C, block_stride, gridDim.x, C, c, lane, s1, s2);
      if (lane == 0) {
        const double m = s1 / double(P);
        double var = s2 / double(P) - m * m;
        var = var < 0.0 ? 0.0 : var;
        (br ? a.mean2 : a.mean)[c] = float(m);
        (br ? a.rstd2 : a.rstd)[c] = float(1.0 / sqrt(var + double(a.eps)));
        float* rm = br ? a.running_mean2 : a.running_mean;
        float* rv = br ? a.running_var2 : a.running_var;
        if (rm) {
          const double unbiased = P > 1 ? var * double(P) / double(P - 1) : var;
          rm[c] = (1.f - a.momentum) * rm[c] + a.momentum * float(m);
          rv[c] = (1.f - a.momentum) * rv[c] + a.momentum * float(unbiased);
        }
      }
    }
  }
  grid_barrier(a.counters, barrier_target);
  grid_barrier_release(a.counters);
  // ---- phase 2: normalise the rows this block reduced (L2 hits)
  const long long e0 = r0 * C / 8, e1 = r1 * C / 8;
  const bool invariant = (256 * 8) % C == 0;
  long long i = e0 + threadIdx.x;
  BnAffine p1, p2;
  if (i < e1 && invariant) {
    const int c = int((i * 8) % C);
    p1.load(a.mean, a.rstd, ap.gamma, ap.beta, c);
    if (ap.y2) p2.load(a.mean2, a.rstd2, ap.gamma2, ap.beta2, c);
  }
  for (; i < e1; i += 256) {
    const long long off = i * 8;
    if (!invariant) {
      const int c = int(off % C);
      p1.load(a.mean, a.rstd, ap.gamma, ap.beta, c);
      if (ap.y2) p2.load(a.mean2, a.rstd2, ap.gamma2, ap.beta2, c);
    }
    float y[8], o[8];
    load8(ap.y + off, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (y[j] - p1.mu[j]) * p1.scale[j] + p1.shift[j];
    if (ap.y2) {
      load8(ap.y2 + off, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += (y[j] - p2.mu[j]) * p2.scale[j] + p2.shift[j];
    }
    if (ap.res_hi) {
      float rh[8];
      load8_bf16(static_cast<const bf16*>(ap.res_hi) + off, rh);
      if (ap.res_lo) {
        float rl[8];
        load8_bf16(static_cast<const bf16*>(ap.res_lo) + off, rl);
#pragma unroll
        for (int j = 0; j < 8; ++j) rh[j] += rl[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += rh[j];
    }
    if (ap.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    store8_split(static_cast<bf16*>(ap.out_hi), static_cast<bf16*>(ap.out_lo), off, o);
  }
}

struct BnFusedBwdArgs {
  fb_bn_bwd_args bw;       // dA, dA2, mask, y, mean, rstd, gamma, P, C, ws, dgamma, dbeta, dy, dz_out, (stats, stats_rows)
  float* partial;          // [grid][2][C]
  float* coef;             // [2][C]
  unsigned int* counters;  // [2]
  int rows_per_block;
};

__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(BnFusedBwdArgs a) {
  griddep_wait();
  __shared__ float4 red[2][256];
  const fb_bn_bwd_args& bw = a.bw;
  const int C = bw.C;
  const long long P = bw.P;
  const long long r0 = (long long)blockIdx.x * a.rows_per_block;
  const long long r1 = min(P, r0 + a.rows_per_block);
  const long long block_stride = 2LL * C;
  // statistics already reduced per CTA by the dgrad that produced dA (fb_conv_gemm_args.bwd_y): no first pass, one barrier
  const bool external = bw.stats != nullptr;
  unsigned int barrier_target = gridDim.x;
  if (!external) {
    block_column_sums<true>(bw.y, bw.dA, bw.dA2, static_cast<const bf16*>(bw.mask_hi), bw.mean, bw.rstd, r0, r1, C,
                            a.partial + blockIdx.x * block_stride, red);
    grid_barrier(a.counters, gridDim.x);
    barrier_target = 2 * gridDim.x;
  }
  {
    const int lane = threadIdx.x & 31;
    for (int c = blockIdx.x * 8 + (threadIdx.x >> 5); c < C; c += gridDim.x * 8) {
      double s1, s2;
      if (external)
        warp_reduce_partials(bw.stats, 2LL * C, bw.stats_rows, C, c, lane, s1, s2);
      else
        warp_reduce_partials(a.partial, block_stride, gridDim.x, C, c, lane, s1, s2);
      if (lane == 0) {
        bw.dbeta[c] = float(s1);
        bw.dgamma[c] = float(s2);
        a.coef[c] = float(s1 / double(P));
        a.coef[C + c] = float(s2 / double(P));
      }
    }
  }
  grid_barrier(a.counters, barrier_target);
  grid_barrier_release(a.counters);
  // the rows are walked BACKWARDS: phase 1 read them front to back, so the tail of this block's slice is what L2 still
  // holds when the whole tensor set (up to 117 MB on the 32x32 stage) does not fit
  const long long e0 = r0 * C / 8, e1 = r1 * C / 8;
  const bool invariant = (256 * 8) % C == 0;
  const long long first = e0 + threadIdx.x;
  long long i = first < e1 ? first + ((e1 - 1 - first) / 256) * 256 : e1;
  BnBwdCoef p;
  if (first < e1 && invariant) p.load(bw, a.coef, int((first * 8) % C));
  for (; first < e1 && i >= first; i -= 256) {
    const long long off = i * 8;
    if (!invariant) p.load(bw, a.coef, int(off % C));
    float d[8], y[8], o[8];
    load8(bw.dA + off, d);
    if (bw.dA2) {
      float d2[8];
      load8(bw.dA2 + off, d2);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] += d2[j];
    }
    if (bw.mask_hi) {
      float m[8];
      load8_bf16(static_cast<const bf16*>(bw.mask_hi) + off, m);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = m[j] > 0.f ? d[j] : 0.f;
    }
    load8(bw.y + off, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xhat = (y[j] - p.mu[j]) * p.rs[j];
      o[j] = p.grs[j] * (d[j] - p.c1[j] - xhat * p.c2[j]);
    }
    store8_bf16(static_cast<bf16*>(bw.dy_bf16), off, o);
    if (bw.dz_out) store8(bw.dz_out + off, d);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Small feature maps (8x8, 4x4, ...: the whole [P][C] problem is a few MB): variant WITHOUT grid barriers, opt-in (see
// sliced_geometry for the measurements).  The tensor is cut into channel SLICES of `sw` channels; a slice is owned by
// `cs` CTAs that split the rows.
//   forward  (statistics come from the conv epilogue): no synchronisation at all -- every CTA reduces the partial rows
//            of its own slice redundantly (fixed order), the CTA of row group 0 publishes mean / rstd / running stats;
//   backward: the cs CTAs of a slice form a thread-block CLUSTER and combine their column sums through distributed
//            shared memory (fixed rank order -> deterministic) between two hardware cluster barriers.
// Thread layout: TX = sw/4 threads per row (one float4 each), TY = 256/TX rows per iteration.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSliceMax = 32;   // channels per slice (sw in {8, 16, 32})
constexpr int kSliceCtas = 8;   // CTAs (= cluster size in the backward kernel) per slice

__device__ __forceinline__ void store4_split(bf16* hi, bf16* lo, long long off, const float4 v) {
  bf16 h0, l0, h1, l1, h2, l2, h3, l3;
  split_bf16(v.x, h0, l0);
  split_bf16(v.y, h1, l1);
  split_bf16(v.z, h2, l2);
  split_bf16(v.w, h3, l3);
  __nv_bfloat162 hh[2] = {__halves2bfloat162(h0, h1), __halves2bfloat162(h2, h3)};
  *reinterpret_cast<uint2*>(hi + off) = *reinterpret_cast<uint2*>(hh);
  if (lo) {
    __nv_bfloat162 ll[2] = {__halves2bfloat162(l0, l1), __halves2bfloat162(l2, l3)};
    *reinterpret_cast<uint2*>(lo + off) = *reinterpret_cast<uint2*>(ll);
  }
}
__device__ __forceinline__ float4 load4_bf16(const bf16* p) {
  const uint2 m = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&m.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

__global__ void __launch_bounds__(256) bn_fwd_sliced_kernel(BnFusedFwdArgs a, int sw, int rows_per_cta) {
  griddep_wait();
  __shared__ float s_mu[2][kSliceMax], s_scale[2][kSliceMax], s_shift[2][kSliceMax];
  const fb_bn_apply_args& ap = a.ap;
  const int C = ap.C;
  const long long P = ap.P;
  const int branches = ap.y2 ? 2 : 1;
  const int c_base = blockIdx.y * sw;
  {  // per-CTA finalize of this slice from the conv epilogue's partial rows: one warp per (branch, channel)
    const int lane = threadIdx.x & 31;
    for (int item = threadIdx.x >> 5; item < branches * sw; item += 8) {
      const int br = item / sw, cl = item % sw, c = c_base + cl;
      double s1, s2;
      warp_reduce_partials(br ? a.ext1 : a.ext0, 2LL * C, br ? a.ext_rows1 : a.ext_rows0, C, c, lane, s1, s2);
      if (lane == 0) {
        const double m = s1 / double(P);
        double var = s2 / double(P) - m * m;
        var = var < 0.0 ? 0.0 : var;
        const float mean = float(m), rstd = float(1.0 / sqrt(var + double(a.eps)));
        const float ga = (br ? ap.gamma2 : ap.gamma)[c], be = (br ? ap.beta2 : ap.beta)[c];
        s_mu[br][cl] = mean;
        s_scale[br][cl] = rstd * ga;
        s_shift[br][cl] = be;
        if (blockIdx.x == 0) {
          (br ? a.mean2 : a.mean)[c] = mean;
          (br ? a.rstd2 : a.rstd)[c] = rstd;
          float* rm = br ? a.running_mean2 : a.running_mean;
          float* rv = br ? a.running_var2 : a.running_var;
          if (rm) {
            const double unbiased = P > 1 ? var * double(P) / double(P - 1) : var;
            rm[c] = (1.f - a.momentum) * rm[c] + a.momentum * mean;
            rv[c] = (1.f - a.momentum) * rv[c] + a.momentum * float(unbiased);
          }
        }
      }
    }
  }
  __syncthreads();
  const int TX = sw / 4, TY = 256 / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int cl = tx * 4, c = c_base + cl;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(P, r0 + rows_per_cta);
  float mu[2][4], sc[2][4], sh[2][4];
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mu[b][j] = s_mu[b][cl + j];
      sc[b][j] = s_scale[b][cl + j];
      sh[b][j] = s_shift[b][cl + j];
    }
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += TY) {
    const long long off = r * C + c;
    const float4 y = *reinterpret_cast<const float4*>(ap.y + off);
    float4 o = make_float4((y.x - mu[0][0]) * sc[0][0] + sh[0][0], (y.y - mu[0][1]) * sc[0][1] + sh[0][1],
                           (y.z - mu[0][2]) * sc[0][2] + sh[0][2], (y.w - mu[0][3]) * sc[0][3] + sh[0][3]);
    if (ap.y2) {
      const float4 z = *reinterpret_cast<const float4*>(ap.y2 + off);
      o.x += (z.x - mu[1][0]) * sc[1][0] + sh[1][0];
      o.y += (z.y - mu[1][1]) * sc[1][1] + sh[1][1];
      o.z += (z.z - mu[1][2]) * sc[1][2] + sh[1][2];
      o.w += (z.w - mu[1][3]) * sc[1][3] + sh[1][3];
    }
    if (ap.res_hi) {
      float4 rh = load4_bf16(static_cast<const bf16*>(ap.res_hi) + off);
      if (ap.res_lo) {
        const float4 rl = load4_bf16(static_cast<const bf16*>(ap.res_lo) + off);
        rh.x += rl.x; rh.y += rl.y; rh.z += rl.z; rh.w += rl.w;
      }
      o.x += rh.x; o.y += rh.y; o.z += rh.z; o.w += rh.w;
    }
    if (ap.relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    store4_split(static_cast<bf16*>(ap.out_hi), static_cast<bf16*>(ap.out_lo), off, o);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_cluster_kernel(BnFusedBwdArgs a, int sw, int rows_per_cta) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  griddep_wait();
  __shared__ float4 red[2][256];
  __shared__ float part[2][kSliceMax];  // this CTA's column sums, read by the other CTAs of the cluster
  __shared__ float coef[2][kSliceMax];
  const fb_bn_bwd_args& bw = a.bw;
  const int C = bw.C;
  const long long P = bw.P;
  const unsigned rank = cluster.block_rank(), cs = cluster.num_blocks();
  const int c_base = blockIdx.y * sw;
  const int TX = sw / 4, TY = 256 / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int cl = tx * 4, c = c_base + cl;
  const long long r0 = (long long)rank * rows_per_cta;
  const long long r1 = min(P, r0 + rows_per_cta);
  const float4 mu = *reinterpret_cast<const float4*>(bw.mean + c);
  const float4 rs = *reinterpret_cast<const float4*>(bw.rstd + c);
  const bf16* mask = static_cast<const bf16*>(bw.mask_hi);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += TY) {
    const long long off = r * C + c;
    const float4 v = *reinterpret_cast<const float4*>(bw.y + off);
    float4 d = *reinterpret_cast<const float4*>(bw.dA + off);
    if (bw.dA2) {
      const float4 d2 = *reinterpret_cast<const float4*>(bw.dA2 + off);
      d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
    }
    if (mask) {
      const float4 m = load4_bf16(mask + off);
      d.x = m.x > 0.f ? d.x : 0.f;
      d.y = m.y > 0.f ? d.y : 0.f;
      d.z = m.z > 0.f ? d.z : 0.f;
      d.w = m.w > 0.f ? d.w : 0.f;
    }
    s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
    s2.x += d.x * (v.x - mu.x) * rs.x;
    s2.y += d.y * (v.y - mu.y) * rs.y;
    s2.z += d.z * (v.z - mu.z) * rs.z;
    s2.w += d.w * (v.w - mu.w) * rs.w;
  }
  red[0][threadIdx.x] = s1;
  red[1][threadIdx.x] = s2;
  __syncthreads();
  if (ty == 0) {
    for (int j = 1; j < TY; ++j) {
      const float4 p1 = red[0][j * TX + tx], p2 = red[1][j * TX + tx];
      s1.x += p1.x; s1.y += p1.y; s1.z += p1.z; s1.w += p1.w;
      s2.x += p2.x; s2.y += p2.y; s2.z += p2.z; s2.w += p2.w;
    }
    *reinterpret_cast<float4*>(&part[0][cl]) = s1;
    *reinterpret_cast<float4*>(&part[1][cl]) = s2;
  }
  cluster.sync();
  if (threadIdx.x < sw) {  // every CTA sums the cluster's partials in rank order: identical totals everywhere
    double t1 = 0.0, t2 = 0.0;
    for (unsigned k = 0; k < cs; ++k) {
      const float* rp = cluster.map_shared_rank(&part[0][0], k);
      t1 += rp[threadIdx.x];
      t2 += rp[kSliceMax + threadIdx.x];
    }
    coef[0][threadIdx.x] = float(t1 / double(P));
    coef[1][threadIdx.x] = float(t2 / double(P));
    if (rank == 0) {
      bw.dbeta[c_base + threadIdx.x] = float(t1);
      bw.dgamma[c_base + threadIdx.x] = float(t2);
    }
  }
  cluster.sync();  // all remote reads of `part` are done (no CTA may exit before), coef visible to the block
  const float4 ga = *reinterpret_cast<const float4*>(bw.gamma + c);
  const float4 grs = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
  const float4 c1 = *reinterpret_cast<const float4*>(&coef[0][cl]);
  const float4 c2 = *reinterpret_cast<const float4*>(&coef[1][cl]);
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += TY) {
    const long long off = r * C + c;
    const float4 v = *reinterpret_cast<const float4*>(bw.y + off);
    float4 d = *reinterpret_cast<const float4*>(bw.dA + off);
    if (bw.dA2) {
      const float4 d2 = *reinterpret_cast<const float4*>(bw.dA2 + off);
      d.x += d2.x; d.y += d2.y; d.z += d2.z; d.w += d2.w;
    }
    if (mask) {
      const float4 m = load4_bf16(mask + off);
      d.x = m.x > 0.f ? d.x : 0.f;
      d.y = m.y > 0.f ? d.y : 0.f;
      d.z = m.z > 0.f ? d.z : 0.f;
      d.w = m.w > 0.f ? d.w : 0.f;
    }
    float4 o;
    o.x = grs.x * (d.x - c1.x - (v.x - mu.x) * rs.x * c2.x);
    o.y = grs.y * (d.y - c1.y - (v.y - mu.y) * rs.y * c2.y);
    o.z = grs.z * (d.z - c1.z - (v.z - mu.z) * rs.z * c2.z);
    o.w = grs.w * (d.w - c1.w - (v.w - mu.w) * rs.w * c2.w);
    __nv_bfloat162 ob[2] = {__floats2bfloat162_rn(o.x, o.y), __floats2bfloat162_rn(o.z, o.w)};
    *reinterpret_cast<uint2*>(static_cast<bf16*>(bw.dy_bf16) + off) = *reinterpret_cast<uint2*>(ob);
    if (bw.dz_out) *reinterpret_cast<float4*>(bw.dz_out + off) = d;
  }
}

// slice width / rows per CTA of the small-map kernels, or false if the problem should use the grid-barrier kernels
static bool sliced_geometry(long long P, int C, int& sw, int& rows_per_cta) {
  // opt-in (FB_BN_SLICED_MAX = largest P*C in elements): measured on B200 (tools/bn_timing.py) the cluster kernel wins
  // only with L2-warm operands and 128-byte slice rows (4x4 maps: 9.2 vs 12.9 us); in the step Y and the ReLU mask
  // come from HBM and the two variants are within 1 us, narrower slices (64 / 32-byte rows) are 1.3-2.4x slower.
  const char* e = getenv("FB_BN_SLICED_MAX");
  const long long max_elems = e ? atoll(e) : 0;
  if (P * C > max_elems || C % 8 != 0) return false;
  sw = C < kSliceMax ? C : kSliceMax;
  while (sw > 8 && C % sw != 0) sw /= 2;
  if (C % sw != 0) return false;
  const int TY = 256 / (sw / 4);
  long long rows = (P + kSliceCtas - 1) / kSliceCtas;
  rows = (rows + TY - 1) / TY * TY;
  rows_per_cta = int(rows);
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// AvgPool2d(2)
// ---------------------------------------------------------------------------------------------------------------
__global__ void avgpool2_fwd_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, int n, int h, int w,
                                    int c, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
  griddep_wait();
  griddep_launch();
  const int ho = h / 2, wo = w / 2;
  const long long total8 = (long long)n * ho * wo * c / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const long long off = i * 8;
    const int cc = int(off % c);
    long long pix = off / c;
    const int x = int(pix % wo);
    pix /= wo;
    const int yy = int(pix % ho);
    const int img = int(pix / ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long src = (((long long)img * h + (2 * yy + dy)) * w + (2 * x + dx)) * c + cc;
        float v[8];
        load8_bf16(in_hi + src, v);
        if (in_lo) {
          float l[8];
          load8_bf16(in_lo + src, l);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += l[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    store8_split(out_hi, out_lo, off, acc);
  }
}

__global__ void avgpool2_bwd_kernel(const float* __restrict__ dP, int n, int h, int w, int c, float* __restrict__ dX,
                                    int accumulate) {
  griddep_wait();
  griddep_launch();
  const int ho = h / 2, wo = w / 2;
  const long long total4 = (long long)n * h * w * c / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long off = i * 4;
    const int cc = int(off % c);
    long long pix = off / c;
    const int x = int(pix % w);
    pix /= w;
    const int yy = int(pix % h);
    const int img = int(pix / h);
    const float4 g = *reinterpret_cast<const float4*>(dP + (((long long)img * ho + yy / 2) * wo + x / 2) * c + cc);
    float4 o = make_float4(0.25f * g.x, 0.25f * g.y, 0.25f * g.z, 0.25f * g.w);
    if (accumulate) {
      const float4 e = *reinterpret_cast<const float4*>(dX + off);
      o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
    }
    *reinterpret_cast<float4*>(dX + off) = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// stem im2col: x NCHW fp32 -> 3x3/pad-1 patches [n*1024][64] bf16 hi/lo, column = ci*9 + kh*3 + kw
// ---------------------------------------------------------------------------------------------------------------
__global__ void stem_im2col_kernel(const float* __restrict__ x, const long long* __restrict__ labels,
                                   const long long* __restrict__ perm, const int* __restrict__ first_dev,
                                   long long first, int n, bf16* __restrict__ p_hi, bf16* __restrict__ p_lo,
                                   long long* __restrict__ labels_out) {
  griddep_wait();
  griddep_launch();
  if (first_dev) first += (long long)(*first_dev) * n;
  const long long total = (long long)n * 1024 * 8;  // 8 groups of 8 columns per pixel
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int grp = int(i & 7);
    const long long pix = i >> 3;
    const int w = int(pix & 31), h = int((pix >> 5) & 31);
    const int img = int(pix >> 10);
    const long long src_img = perm ? perm[first + img] : first + img;
    const float* xi = x + src_img * 3072;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = grp * 8 + j;
      float val = 0.f;
      if (k < 27) {
        const int ci = k / 9, kh = (k % 9) / 3, kw = k % 3;
        const int hh = h + kh - 1, ww = w + kw - 1;
        if (hh >= 0 && hh < 32 && ww >= 0 && ww < 32) val = xi[ci * 1024 + hh * 32 + ww];
      }
      v[j] = val;
    }
    store8_split(p_hi, p_lo, pix * 64 + grp * 8, v);
    if (labels_out && grp == 0 && (pix & 1023) == 0) labels_out[img] = labels[src_img];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Same, fused with the reference's training augmentation for CIFAR (config/data/CIFAR10.yaml:22-26 applied by
// torchvision in data_preparation.py:173-200): RandomCrop(32, padding=4) -> RandomHorizontalFlip -> ToTensor ->
// Normalize(mean, std), evaluated on the fly from a device-resident uint8 HWC dataset.  aug[pos] = (dx, dy, flip, -)
// holds this epoch's draws for the sample at position pos of the (optionally permuted) order: crop offsets 0..8 inside
// the zero-padded 40x40 image and the flip bit; padded pixels are black BEFORE normalisation, exactly like torchvision.
// ---------------------------------------------------------------------------------------------------------------
struct AugNorm {
  float mean[3], inv_std[3];
};

__global__ void stem_im2col_u8aug_kernel(const uint8_t* __restrict__ x, const long long* __restrict__ labels,
                                         const long long* __restrict__ perm, const int* __restrict__ first_dev,
                                         long long first, int n, const char4* __restrict__ aug, AugNorm nrm,
                                         bf16* __restrict__ p_hi, bf16* __restrict__ p_lo,
                                         long long* __restrict__ labels_out) {
  griddep_wait();
  griddep_launch();
  if (first_dev) first += (long long)(*first_dev) * n;
  const long long total = (long long)n * 1024 * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int grp = int(i & 7);
    const long long pix = i >> 3;
    const int w = int(pix & 31), h = int((pix >> 5) & 31);
    const int img = int(pix >> 10);
    const long long pos = first + img;
    const long long src_img = perm ? perm[pos] : pos;
    const char4 a = aug ? aug[pos] : make_char4(4, 4, 0, 0);
    const uint8_t* xi = x + src_img * 3072;  // HWC
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = grp * 8 + j;
      float val = 0.f;
      if (k < 27) {
        const int ci = k / 9, kh = (k % 9) / 3, kw = k % 3;
        const int hh = h + kh - 1, ww = w + kw - 1;  // pixel of the AUGMENTED image; outside -> conv zero padding
        if (hh >= 0 && hh < 32 && ww >= 0 && ww < 32) {
          const int wc = a.z ? 31 - ww : ww;          // flip acts on the cropped image
          const int sh = hh + a.y - 4, sw = wc + a.x - 4;  // source pixel in the un-padded original
          const float raw = (sh >= 0 && sh < 32 && sw >= 0 && sw < 32) ? float(xi[(sh * 32 + sw) * 3 + ci]) : 0.f;
          val = (raw * (1.f / 255.f) - nrm.mean[ci]) * nrm.inv_std[ci];
        }
      }
      v[j] = val;
    }
    store8_split(p_hi, p_lo, pix * 64 + grp * 8, v);
    if (labels_out && grp == 0 && (pix & 1023) == 0) labels_out[img] = labels[src_img];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// head: global average pool -> linear -> label-smoothed cross entropy (+accuracy) and backward
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 16;

// grid = n, block = 128
__global__ void __launch_bounds__(128) head_fwd_kernel(const bf16* __restrict__ a_hi, const bf16* __restrict__ a_lo,
                                                       int n, int hw, int c, const float* __restrict__ fc_w,
                                                       const float* __restrict__ fc_b,
                                                       const long long* __restrict__ labels, int classes,
                                                       float smoothing, float* __restrict__ pooled,
                                                       float* __restrict__ dlogits, float* __restrict__ loss_n,
                                                       float* __restrict__ correct_n) {
  griddep_wait();
  griddep_launch();
  extern __shared__ float sp[];  // c floats + classes logits
  float* logits = sp + c;
  const int img = blockIdx.x;
  const float inv = 1.f / float(hw);
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < hw; ++p) {
      const long long o = ((long long)img * hw + p) * c + ch;
      s += __bfloat162float(a_hi[o]) + (a_lo ? __bfloat162float(a_lo[o]) : 0.f);
    }
    s *= inv;
    sp[ch] = s;
    pooled[(long long)img * c + ch] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cls = warp; cls < classes; cls += 4) {
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) s += sp[ch] * fc_w[(long long)cls * c + ch];
    s = warp_sum(s);
    if (lane == 0) logits[cls] = s + fc_b[cls];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int label = int(labels[img]);
    float mx = logits[0];
    int arg = 0;
    for (int k = 1; k < classes; ++k)
      if (logits[k] > mx) {
        mx = logits[k];
        arg = k;
      }
    float se = 0.f;
    for (int k = 0; k < classes; ++k) se += expf(logits[k] - mx);
    const float lse = mx + logf(se);
    const float w_off = smoothing / float(classes - 1), w_on = 1.f - smoothing;
    float loss = 0.f;
    for (int k = 0; k < classes; ++k) {
      const float logp = logits[k] - lse;
      const float wk = (k == label) ? w_on : w_off;
      loss -= wk * logp;
      // d/dz_k of -sum_j w_j logp_j = softmax_k * sum_j w_j - w_k; mean over the microbatch -> / n
      const float wsum = w_on + w_off * float(classes - 1);
      dlogits[img * kMaxClasses + k] = (expf(logp) * wsum - wk) / float(n);
    }
    loss_n[img] = loss;
    correct_n[img] = (arg == label) ? 1.f : 0.f;
  }
}

// grid = (c/128, n): dA[n][p][ch] = (sum_k dlogits[n][k] * W[k][ch]) / hw
__global__ void __launch_bounds__(128) head_bwd_act_kernel(const float* __restrict__ dlogits,
                                                           const float* __restrict__ fc_w, int hw, int c, int classes,
                                                           float* __restrict__ dA) {
  griddep_wait();
  griddep_launch();
  const int ch = blockIdx.x * 128 + threadIdx.x;
  const int img = blockIdx.y;
  if (ch >= c) return;
  float s = 0.f;
  for (int k = 0; k < classes; ++k) s += dlogits[img * kMaxClasses + k] * fc_w[(long long)k * c + ch];
  s /= float(hw);
  for (int p = 0; p < hw; ++p) dA[((long long)img * hw + p) * c + ch] = s;
}

// grid = c/32 (+ block 0 also reduces bias grad, loss, accuracy); block = 32 channels x 8 sample lanes
__global__ void __launch_bounds__(256) head_bwd_param_kernel(const float* __restrict__ dlogits,
                                                             const float* __restrict__ pooled,
                                                             const float* __restrict__ loss_n,
                                                             const float* __restrict__ correct_n, int n, int c,
                                                             int classes, float* __restrict__ d_fcw,
                                                             float* __restrict__ d_fcb, float* __restrict__ scal,
                                                             int loss_slot, int correct_slot) {
  griddep_wait();
  griddep_launch();
  __shared__ float red[8][kMaxClasses][33];
  const int cl = threadIdx.x & 31, lane_n = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  float acc[kMaxClasses];
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) acc[k] = 0.f;
  if (ch < c) {
    for (int i = lane_n; i < n; i += 8) {
      const float pv = pooled[(long long)i * c + ch];
      const float4* dl = reinterpret_cast<const float4*>(dlogits + i * kMaxClasses);
#pragma unroll
      for (int q = 0; q < kMaxClasses / 4; ++q) {
        const float4 d = dl[q];
        acc[4 * q] += d.x * pv;
        acc[4 * q + 1] += d.y * pv;
        acc[4 * q + 2] += d.z * pv;
        acc[4 * q + 3] += d.w * pv;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) red[lane_n][k][cl] = acc[k];
  __syncthreads();
  // 256 threads -> (class, channel) pairs of this block: 16 x 32 = 512 outputs, two per thread
  for (int o = threadIdx.x; o < kMaxClasses * 32; o += 256) {
    const int k = o >> 5, cc = o & 31;
    if (k < classes && blockIdx.x * 32 + cc < c) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += red[j][k][cc];
      d_fcw[(long long)k * c + blockIdx.x * 32 + cc] = sum;
    }
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < classes) {
      float sum = 0.f;
      for (int i = 0; i < n; ++i) sum += dlogits[i * kMaxClasses + threadIdx.x];
      d_fcb[threadIdx.x] = sum;
    } else if (threadIdx.x == 32) {
      double sum = 0.0;
      for (int i = 0; i < n; ++i) sum += loss_n[i];
      scal[loss_slot] += float(sum / double(n));
    } else if (threadIdx.x == 64) {
      float sum = 0.f;
      for (int i = 0; i < n; ++i) sum += correct_n[i];
      scal[correct_slot] += sum;
    }
  }
}

static int reduce_geometry(long long P, int C, int& TX, int& slabs, int& chunks, int& rows_per_chunk) {
  if (C % 4 != 0) return FB_ERR_UNSUPPORTED;
  const int c4 = C / 4;
  TX = c4 < 256 ? c4 : 256;
  if (256 % TX != 0 || c4 % TX != 0) return FB_ERR_UNSUPPORTED;
  slabs = c4 / TX;
  const int TY = 256 / TX;
  long long want = (P + TY * 4 - 1) / (TY * 4);  // >= 4 rows per thread
  if (want < 1) want = 1;
  chunks = int(want < kMaxChunks ? want : kMaxChunks);
  rows_per_chunk = int((P + chunks - 1) / chunks);
  chunks = int((P + rows_per_chunk - 1) / rows_per_chunk);
  return 0;
}

static int stream_grid(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return int(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace fb

using namespace fb;

extern "C" int fb_bn_stats(const float* y, int64_t P, int C, float* ws, float* mean, float* rstd, float* running_mean,
                           float* running_var, float momentum, float eps, void* stream) {
  FB_REQUIRE(y && ws && mean && rstd && P > 0, "fb_bn_stats: bad arguments");
  int TX, slabs, chunks, rpc;
  if (reduce_geometry(P, C, TX, slabs, chunks, rpc)) {
    set_error("fb_bn_stats: unsupported channel count %d", C);
    return FB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BnFinalize fin = {};
  fin.P = P;
  fin.chunks = chunks;
  fin.mean_out = mean;
  fin.rstd_out = rstd;
  fin.running_mean = running_mean;
  fin.running_var = running_var;
  fin.momentum = momentum;
  fin.eps = eps;
  bn_reduce_kernel<false><<<dim3(chunks, slabs), 256, 0, st>>>(y, nullptr, nullptr, nullptr, nullptr, nullptr, P, C, TX,
                                                               rpc, ws);
  bn_finalize_kernel<false><<<(C + 7) / 8, 256, 0, st>>>(ws, C, fin);
  FB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fb_bn_apply(const fb_bn_apply_args* a, void* stream) {
  FB_REQUIRE(a && a->y && a->mean && a->rstd && a->gamma && a->beta && a->out_hi, "fb_bn_apply: null pointer");
  FB_REQUIRE(a->C % 8 == 0 && a->P > 0, "fb_bn_apply: C must be a multiple of 8");
  bn_apply_kernel<<<stream_grid(a->P * a->C / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  FB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fb_bn_bwd(const fb_bn_bwd_args* a, void* stream) {
  FB_REQUIRE(a && a->dA && a->y && a->mean && a->rstd && a->gamma && a->ws && a->dgamma && a->dbeta && a->dy_bf16,
             "fb_bn_bwd: null pointer");
  FB_REQUIRE(a->C % 8 == 0 && a->P > 0, "fb_bn_bwd: C must be a multiple of 8");
  int TX, slabs, chunks, rpc;
  if (reduce_geometry(a->P, a->C, TX, slabs, chunks, rpc)) {
    set_error("fb_bn_bwd: unsupported channel count %d", a->C);
    return FB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* coef = a->ws + (long long)2 * a->C * kMaxChunks;
  BnFinalize fin = {};
  fin.P = a->P;
  fin.chunks = chunks;
  fin.coef = coef;
  fin.dgamma = a->dgamma;
  fin.dbeta = a->dbeta;
  bn_reduce_kernel<true><<<dim3(chunks, slabs), 256, 0, st>>>(a->y, a->dA, a->dA2, static_cast<const bf16*>(a->mask_hi),
                                                              a->mean, a->rstd, a->P, a->C, TX, rpc, a->ws);
  bn_finalize_kernel<true><<<(a->C + 7) / 8, 256, 0, st>>>(a->ws, a->C, fin);
  bn_bwd_apply_kernel<<<stream_grid(a->P * a->C / 8), 256, 0, st>>>(*a, coef);
  FB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fb_avgpool2_fwd(const void* in_hi, const void* in_lo, int n, int h, int w, int c, void* out_hi,
                               void* out_lo, void* stream) {
  FB_REQUIRE(in_hi && out_hi && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "fb_avgpool2_fwd: bad arguments");
  FB_CUDA(launch_pdl(avgpool2_fwd_kernel, dim3(stream_grid((long long)n * h * w * c / 32)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), static_cast<const bf16*>(in_hi), static_cast<const bf16*>(in_lo),
                     n, h, w, c, static_cast<bf16*>(out_hi), static_cast<bf16*>(out_lo)));
  return 0;
}

extern "C" int fb_avgpool2_bwd(const float* dP, int n, int h, int w, int c, float* dX, int accumulate, void* stream) {
  FB_REQUIRE(dP && dX && h % 2 == 0 && w % 2 == 0 && c % 4 == 0, "fb_avgpool2_bwd: bad arguments");
  FB_CUDA(launch_pdl(avgpool2_bwd_kernel, dim3(stream_grid((long long)n * h * w * c / 4)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), dP, n, h, w, c, dX, accumulate));
  return 0;
}

extern "C" int fb_stem_im2col(const float* x, const int64_t* labels, const int64_t* perm, const int32_t* first_dev,
                              int64_t first, int n, void* patches_hi, void* patches_lo, int64_t* labels_out,
                              void* stream) {
  FB_REQUIRE(x && patches_hi && n > 0, "fb_stem_im2col: bad arguments");
  FB_REQUIRE(!labels_out || labels, "fb_stem_im2col: labels_out needs labels");
  FB_CUDA(launch_pdl(stem_im2col_kernel, dim3(stream_grid((long long)n * 1024 * 8)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), x, reinterpret_cast<const long long*>(labels),
                     reinterpret_cast<const long long*>(perm), first_dev, (long long)first, n,
                     static_cast<bf16*>(patches_hi), static_cast<bf16*>(patches_lo),
                     reinterpret_cast<long long*>(labels_out)));
  return 0;
}

extern "C" int fb_head_fwd_bwd(const void* a_hi, const void* a_lo, int n, int hw, int c, const float* fc_w,
                               const float* fc_b, const int64_t* labels, int classes, float smoothing, float* ws,
                               float* scal, int loss_slot, int correct_slot, float* d_fcw, float* d_fcb, float* dA,
                               void* stream) {
  FB_REQUIRE(a_hi && fc_w && fc_b && labels && ws && scal && d_fcw && d_fcb && dA, "fb_head_fwd_bwd: null pointer");
  if (classes > kMaxClasses || classes < 2) {
    set_error("fb_head_fwd_bwd: classes %d not supported (max %d)", classes, kMaxClasses);
    return FB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* pooled = ws;
  float* dlogits = pooled + (long long)n * c;
  float* loss_n = dlogits + (long long)n * kMaxClasses;
  float* correct_n = loss_n + n;
  FB_CUDA(launch_pdl(head_fwd_kernel, dim3(n), dim3(128), (c + kMaxClasses) * sizeof(float), st,
                     static_cast<const bf16*>(a_hi), static_cast<const bf16*>(a_lo), n, hw, c, fc_w, fc_b,
                     reinterpret_cast<const long long*>(labels), classes, smoothing, pooled, dlogits, loss_n, correct_n));
  FB_CUDA(launch_pdl(head_bwd_act_kernel, dim3((c + 127) / 128, n), dim3(128), 0, st, (const float*)dlogits, fc_w, hw, c,
                     classes, dA));
  FB_CUDA(launch_pdl(head_bwd_param_kernel, dim3((c + 31) / 32), dim3(256), 0, st, (const float*)dlogits,
                     (const float*)pooled, (const float*)loss_n, (const float*)correct_n, n, c, classes, d_fcw, d_fcb, scal,
                     loss_slot, correct_slot));
  return 0;
}

static int fused_geometry(long long P, int C, int& grid, int& rows_per_block) {
  if (C % 8 != 0) return FB_ERR_UNSUPPORTED;
  const int c4 = C / 4;
  const int TX = c4 < 256 ? c4 : 256;
  if (256 % TX != 0 || c4 % TX != 0) return FB_ERR_UNSUPPORTED;
  const int TY = 256 / TX;
  // every block must own whole groups of TY rows and an element range that keeps the channel offset thread-invariant:
  // rows_per_block is a multiple of lcm(TY, 2048 / C)
  int unit = TY;
  const int inv = (2048 % C == 0) ? 2048 / C : 1;
  if (inv > unit) unit = inv;
  long long units = (P + unit - 1) / unit;
  long long g = units < 2 * kNumSMs ? units : 2 * kNumSMs;
  long long upb = (units + g - 1) / g;
  rows_per_block = int(upb * unit);
  grid = int((P + rows_per_block - 1) / rows_per_block);
  return 0;
}

extern "C" int fb_bn_fwd_fused(const fb_bn_apply_args* ap, float* mean2_out, float* rstd2_out, float* running_mean,
                               float* running_var, float* running_mean2, float* running_var2, float momentum,
                               float eps, float* ws, const float* stats, int stats_rows, const float* stats2,
                               int stats_rows2, void* stream) {
  FB_REQUIRE(ap && ap->y && ap->mean && ap->rstd && ap->gamma && ap->beta && ap->out_hi && ws,
             "fb_bn_fwd_fused: null pointer");
  FB_REQUIRE(!ap->y2 || (mean2_out && rstd2_out && ap->gamma2 && ap->beta2), "fb_bn_fwd_fused: second branch incomplete");
  int grid, rpb;
  if (fused_geometry(ap->P, ap->C, grid, rpb)) {
    set_error("fb_bn_fwd_fused: unsupported channel count %d", ap->C);
    return FB_ERR_UNSUPPORTED;
  }
  BnFusedFwdArgs a;
  a.ap = *ap;
  a.mean = const_cast<float*>(ap->mean);
  a.rstd = const_cast<float*>(ap->rstd);
  a.mean2 = mean2_out;
  a.rstd2 = rstd2_out;
  a.ap.mean2 = mean2_out;
  a.ap.rstd2 = rstd2_out;
  a.running_mean = running_mean;
  a.running_var = running_var;
  a.running_mean2 = running_mean2;
  a.running_var2 = running_var2;
  a.momentum = momentum;
  a.eps = eps;
  a.counters = reinterpret_cast<unsigned int*>(ws);
  a.partial = ws + 4;
  a.rows_per_block = rpb;
  FB_REQUIRE(!stats || stats_rows > 0, "fb_bn_fwd_fused: stats_rows must be positive");
  FB_REQUIRE(!stats || !ap->y2 || (stats2 && stats_rows2 > 0),
             "fb_bn_fwd_fused: epilogue statistics must be given for both branches or for none");
  a.ext0 = stats;
  a.ext_rows0 = stats_rows;
  a.ext1 = stats2;
  a.ext_rows1 = stats_rows2;
  int sw, rows_per_cta;
  if (stats && sliced_geometry(ap->P, ap->C, sw, rows_per_cta)) {
    FB_CUDA(launch_pdl(bn_fwd_sliced_kernel, dim3(kSliceCtas, ap->C / sw), dim3(256), 0,
                       static_cast<cudaStream_t>(stream), a, sw, rows_per_cta));
    return 0;
  }
  FB_CUDA(launch_pdl(bn_fwd_fused_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), a));
  return 0;
}

extern "C" int fb_bn_bwd_fused(const fb_bn_bwd_args* bw, void* stream) {
  FB_REQUIRE(bw && bw->dA && bw->y && bw->mean && bw->rstd && bw->gamma && bw->ws && bw->dgamma && bw->dbeta &&
                 bw->dy_bf16,
             "fb_bn_bwd_fused: null pointer");
  FB_REQUIRE(!bw->dz_accumulate, "fb_bn_bwd_fused: dz_accumulate is not supported");
  FB_REQUIRE(!bw->stats || (bw->stats_rows > 0 && !bw->dA2),
             "fb_bn_bwd_fused: epilogue statistics need stats_rows > 0 and a single gradient addend");
  int grid, rpb;
  if (fused_geometry(bw->P, bw->C, grid, rpb)) {
    set_error("fb_bn_bwd_fused: unsupported channel count %d", bw->C);
    return FB_ERR_UNSUPPORTED;
  }
  BnFusedBwdArgs a;
  a.bw = *bw;
  a.counters = reinterpret_cast<unsigned int*>(bw->ws);
  a.partial = bw->ws + 4;
  a.coef = a.partial + (long long)2 * bw->C * 2 * kNumSMs;
  a.rows_per_block = rpb;
  int sw, rows_per_cta;
  if (sliced_geometry(bw->P, bw->C, sw, rows_per_cta)) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kSliceCtas, bw->C / sw);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kSliceCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    int n_attr = 1;
    if (pdl_enabled()) {
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      n_attr = 2;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    FB_CUDA(cudaLaunchKernelEx(&cfg, bn_bwd_cluster_kernel, a, sw, rows_per_cta));
    return 0;
  }
  FB_CUDA(launch_pdl(bn_bwd_fused_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), a));
  return 0;
}

extern "C" int fb_stem_im2col_u8aug(const uint8_t* x_hwc, const int64_t* labels, const int64_t* perm,
                                    const int32_t* first_dev, int64_t first, int n, const int8_t* aug,
                                    const float* mean3, const float* std3, void* patches_hi, void* patches_lo,
                                    int64_t* labels_out, void* stream) {
  FB_REQUIRE(x_hwc && patches_hi && mean3 && std3 && n > 0, "fb_stem_im2col_u8aug: bad arguments");
  FB_REQUIRE(!labels_out || labels, "fb_stem_im2col_u8aug: labels_out needs labels");
  AugNorm nrm;
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = mean3[c];
    nrm.inv_std[c] = 1.f / std3[c];
  }
  FB_CUDA(launch_pdl(stem_im2col_u8aug_kernel, dim3(stream_grid((long long)n * 1024 * 8)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), x_hwc, reinterpret_cast<const long long*>(labels),
                     reinterpret_cast<const long long*>(perm), first_dev, (long long)first, n,
                     reinterpret_cast<const char4*>(aug), nrm, static_cast<bf16*>(patches_hi),
                     static_cast<bf16*>(patches_lo), reinterpret_cast<long long*>(labels_out)));
  return 0;
}
