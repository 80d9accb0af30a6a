// Bandwidth-bound layer kernels of the full-batch step (sm_100a): BatchNorm (train mode) apply / backward fused with
// ReLU and the residual add, AvgPool2d(2), stem im2col, the pooled-linear-cross-entropy head and the running-stat EMA.
// All of them are coalesced, 16-byte vectorised streaming kernels over `ng` microbatch groups per launch; reductions
// are two-stage and deterministic (fixed partition that only depends on ONE group's problem, fixed summation order), so
// repeated runs are bit-identical (cf. measure_floating_point_accuracy.py / fullbatch/training/training.py:429-600) and
// results do not depend on how many groups share a launch.  No kernel spins on a grid-wide barrier: cross-block
// reductions finish in the last block to arrive (atomic ticket), the apply pass is a separate launch.
//
// Reference call sites replaced: torch.nn.BatchNorm2d / ReLU(inplace) / `out += identity`
// (fullbatch/models/resnets.py:71,207-230,296-316), AvgPool2d (resnets.py:149), AdaptiveAvgPool2d + Linear
// (resnets.py:106-107), LabelSmoothCrossEntropyLoss (fullbatch/models/modules.py:96-101), accuracy count
// (fullbatch/training/training.py:80) and their autograd backward.
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
// Stand-alone BatchNorm statistics over a [P][C] fp32 matrix (fb_bn_stats): block = 256 threads = TX float4-columns x
// TY rows; grid = (chunks, column slabs) -> partial[chunk][2][C]; a second launch reduces the chunks in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
struct BnFinalize {
  long long P;
  int chunks;
  float *mean_out, *rstd_out, *running_mean, *running_var;
  float momentum, eps;
};

// block = 8 channels x 32 lanes; lane l sums chunks l, l+32, ... (independent loads in flight), lanes are combined by
// a fixed shuffle tree -> deterministic.
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
}

__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ y, long long P, int C, int TX,
                                                        int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float4 red[2][256];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const int c = (blockIdx.y * TX + tx) * 4;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(P, r0 + rows_per_chunk);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll 8
  for (long long r = r0 + ty; r < r1; r += TY) {
    const float4 v = *reinterpret_cast<const float4*>(y + r * C + c);
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    s2.x += v.x * v.x; s2.y += v.y * v.y; s2.z += v.z * v.z; s2.w += v.w * v.w;
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
// BN apply (+ second normalised branch, + residual, + ReLU) -> bf16 hi/lo, grid = (blocks, groups)
// ---------------------------------------------------------------------------------------------------------------
// The per-channel affine parameters (mean, rstd*gamma, beta) of the group live in shared memory; every thread streams
// items of 8 consecutive channels, TWO items per loop round with all loads issued before the arithmetic, so that
// ~3 blocks x 256 threads x 128-192 bytes are in flight per SM (the kernel is a pure HBM stream).
struct BnItem {
  float y[8], y2[8], r[8];
};

template <bool DUAL, bool RES>
__device__ __forceinline__ void bn_item_load(const fb_bn_apply_args& a, long long off, BnItem& it) {
  load8(a.y + off, it.y);
  if (DUAL) load8(a.y2 + off, it.y2);
  if (RES) {
    load8_bf16(static_cast<const bf16*>(a.res_hi) + off, it.r);
    if (a.res_lo) {
      float rl[8];
      load8_bf16(static_cast<const bf16*>(a.res_lo) + off, rl);
#pragma unroll
      for (int j = 0; j < 8; ++j) it.r[j] += rl[j];
    }
  }
}

// Shared-memory layout of the per-channel parameters: one record per channel OCTET (the 8 channels of an item), K
// arrays of 8 floats each, padded to a record stride that keeps the 16-byte reads of 8 consecutive lanes on different
// banks (28 / 52 floats for K = 3 / 6).
__device__ __forceinline__ void load8_smem(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ int channel_of(long long e8, int C, int cmask) {
  return cmask ? int(e8 & cmask) : int(e8 % C);
}

template <bool DUAL, bool RES>
__device__ __forceinline__ void bn_item_store(const fb_bn_apply_args& a, long long off, int c, const float* sp,
                                              const BnItem& it) {
  constexpr int kRec = DUAL ? 52 : 28;
  const float* rec = sp + (c >> 3) * kRec;
  float mu[8], sc[8], sh[8], o[8];
  load8_smem(rec, mu);
  load8_smem(rec + 8, sc);
  load8_smem(rec + 16, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (it.y[j] - mu[j]) * sc[j] + sh[j];
  if (DUAL) {
    load8_smem(rec + 24, mu);
    load8_smem(rec + 32, sc);
    load8_smem(rec + 40, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += (it.y2[j] - mu[j]) * sc[j] + sh[j];
  }
  if (RES) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += it.r[j];
  }
  if (a.relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
  }
  store8_split(static_cast<bf16*>(a.out_hi), static_cast<bf16*>(a.out_lo), off, o);
  if (a.mask_out) {  // ReLU mask of the item as one byte (read by the backward kernels instead of the bf16 plane)
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) m |= (o[j] > 0.f ? 1u : 0u) << j;
    a.mask_out[off >> 3] = (uint8_t)m;
  }
}

template <bool DUAL, bool RES>
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(const __grid_constant__ fb_bn_apply_args a) {
  extern __shared__ __align__(16) float sp[];  // records of [mean | rstd*gamma | beta] (x2 with the second branch)
  constexpr int kRec = DUAL ? 52 : 28;
  griddep_wait();
  griddep_launch();
  // reverse: start with the last elements of the last group = what the producer of y (the convolution) wrote last
  const int g = a.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int C = a.C;
  const int cmask = (C & (C - 1)) == 0 ? C - 1 : 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long long pc = (long long)g * a.param_gstride + c;
    float* rec = sp + (c >> 3) * kRec + (c & 7);
    rec[0] = a.mean[(long long)g * C + c];
    rec[8] = a.rstd[(long long)g * C + c] * a.gamma[pc];
    rec[16] = a.beta[pc];
    if (DUAL) {
      rec[24] = a.mean2[(long long)g * C + c];
      rec[32] = a.rstd2[(long long)g * C + c] * a.gamma2[pc];
      rec[40] = a.beta2[pc];
    }
  }
  __syncthreads();
  const long long total8 = a.P * C / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long gbase = (long long)g * a.P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < total8;
    const long long e1 = a.reverse ? total8 - 1 - i : i;
    const long long e2 = a.reverse ? total8 - 1 - i2 : i2;
    BnItem it1, it2;
    bn_item_load<DUAL, RES>(a, gbase + e1 * 8, it1);
    if (two) bn_item_load<DUAL, RES>(a, gbase + e2 * 8, it2);
    bn_item_store<DUAL, RES>(a, gbase + e1 * 8, channel_of(e1 * 8, C, cmask), sp, it1);
    if (two) bn_item_store<DUAL, RES>(a, gbase + e2 * 8, channel_of(e2 * 8, C, cmask), sp, it2);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BN backward: dy = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) -> bf16; optional dz (fp32) output.
//   launch 1 (bn_bwd_reduce_kernel, grid = chunks x column slabs x groups): per-chunk column sums of dz and dz*xhat;
//            the LAST block of a group (atomic ticket) sums the chunks in a fixed order -> dgamma, dbeta, coef;
//   launch 2 (bn_bwd_apply_kernel, grid = blocks x groups): streaming apply.
// The reduce walks the tensors in the opposite direction of the apply, so that the apply starts on what the reduce
// touched last (L2), and the pair starts where the producer of dA ended when `reverse` says so.
// ---------------------------------------------------------------------------------------------------------------
struct BnBwdGeom {
  int TX, slabs, chunks, rows_per_chunk;
};

struct BnBwdK {
  fb_bn_bwd_args a;
  BnBwdGeom geo;
  float* partial;          // [ng][chunks][2][C]
  float* coef;             // [ng][2][C]
  unsigned int* tickets;   // [ng]
  // 1: the reduce pass stores dz = mask * (dA + dA2) to dz_out (it has the value in registers anyway) and the apply pass
  // reads dz back instead of dA, dA2 and the mask: one fp32 read per element less for BatchNorms with two gradient
  // addends and an identity shortcut (30 -> 26 bytes of DRAM traffic per element); same arithmetic, same bits
  int dz_from_reduce;
};

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(BnBwdK k) {
  griddep_wait();
  griddep_launch();
  __shared__ float4 red[2][256];
  __shared__ unsigned int s_last;
  const fb_bn_bwd_args& a = k.a;
  const int C = a.C;
  const long long P = a.P;
  const int g = a.reverse ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int TX = k.geo.TX, TY = 256 / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int chunk = a.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int c = (blockIdx.y * TX + tx) * 4;
  const long long gbase = (long long)g * P * C;
  const float* y = a.y + gbase;
  const float* dA = a.dA + gbase;
  const float* dA2 = a.dA2 ? a.dA2 + gbase : nullptr;
  const bf16* mask = (a.mask_hi && !a.mask_bits) ? static_cast<const bf16*>(a.mask_hi) + gbase : nullptr;
  const uint8_t* bits = a.mask_bits ? a.mask_bits + (gbase >> 3) : nullptr;
  const int bit_shift = c & 4;
  float* dz_out = k.dz_from_reduce ? a.dz_out + gbase : nullptr;
  const long long r0 = (long long)chunk * k.geo.rows_per_chunk;
  const long long r1 = min(P, r0 + k.geo.rows_per_chunk);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const float4 mu = *reinterpret_cast<const float4*>(a.mean + (long long)g * C + c);
  const float4 rs = *reinterpret_cast<const float4*>(a.rstd + (long long)g * C + c);
  // four rows per round with ALL loads issued before the arithmetic (the kernel is a pure stream: bytes in flight per
  // thread decide its speed); rows are accumulated in the same order as a plain loop would
  constexpr int kRows = 4;
  for (long long r = r0 + ty; r < r1; r += (long long)kRows * TY) {
    float4 v[kRows], d[kRows], d2[kRows];
    uint2 mk[kRows];
    unsigned mb[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const long long rr = r + (long long)u * TY;
      if (rr < r1) {
        const long long o = rr * C + c;
        v[u] = *reinterpret_cast<const float4*>(y + o);
        d[u] = *reinterpret_cast<const float4*>(dA + o);
        if (dA2) d2[u] = *reinterpret_cast<const float4*>(dA2 + o);
        if (bits) mb[u] = bits[o >> 3];
        if (mask) mk[u] = *reinterpret_cast<const uint2*>(mask + o);
      }
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const long long rr = r + (long long)u * TY;
      if (rr < r1) {
        float4 dd = d[u];
        if (dA2) {
          dd.x += d2[u].x; dd.y += d2[u].y; dd.z += d2[u].z; dd.w += d2[u].w;
        }
        if (bits) {
          const unsigned m = mb[u] >> bit_shift;
          dd.x = (m & 1u) ? dd.x : 0.f;
          dd.y = (m & 2u) ? dd.y : 0.f;
          dd.z = (m & 4u) ? dd.z : 0.f;
          dd.w = (m & 8u) ? dd.w : 0.f;
        }
        if (mask) {
          const float2 m01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mk[u].x));
          const float2 m23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mk[u].y));
          dd.x = m01.x > 0.f ? dd.x : 0.f;
          dd.y = m01.y > 0.f ? dd.y : 0.f;
          dd.z = m23.x > 0.f ? dd.z : 0.f;
          dd.w = m23.y > 0.f ? dd.w : 0.f;
        }
        if (dz_out) *reinterpret_cast<float4*>(dz_out + rr * C + c) = dd;
        s1.x += dd.x; s1.y += dd.y; s1.z += dd.z; s1.w += dd.w;
        s2.x += dd.x * (v[u].x - mu.x) * rs.x;
        s2.y += dd.y * (v[u].y - mu.y) * rs.y;
        s2.z += dd.z * (v[u].z - mu.z) * rs.z;
        s2.w += dd.w * (v[u].w - mu.w) * rs.w;
      }
    }
  }
  red[0][threadIdx.x] = s1;
  red[1][threadIdx.x] = s2;
  __syncthreads();
  float* part_g = k.partial + (long long)g * k.geo.chunks * 2 * C;
  if (ty == 0) {
    for (int j = 1; j < TY; ++j) {
      const float4 p1 = red[0][j * TX + tx], p2 = red[1][j * TX + tx];
      s1.x += p1.x; s1.y += p1.y; s1.z += p1.z; s1.w += p1.w;
      s2.x += p2.x; s2.y += p2.y; s2.z += p2.z; s2.w += p2.w;
    }
    float* dst = part_g + (long long)chunk * 2 * C;
    *reinterpret_cast<float4*>(dst + c) = s1;
    *reinterpret_cast<float4*>(dst + C + c) = s2;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int prev = atomicAdd(k.tickets + g, 1u);
    const bool last = prev == total - 1;
    if (last) k.tickets[g] = 0u;  // every block of this group has arrived: ready for the next launch
    __threadfence();
    s_last = last ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last block of the group: sum the chunks (row slices in parallel, then the slices in order)
  double* scratch = reinterpret_cast<double*>(&red[0][0]);  // [256][4] doubles = sizeof(red)
  const int n_items = C / 2;                                 // (sum | dot) x float4 column
  const int per = n_items < 256 ? n_items : 256;
  const int S = 256 / per;
  const int t = threadIdx.x;
  for (int base = 0; base < n_items; base += per) {
    const int item = base + t % per, slice = t / per;
    const int which = item / (C / 4), c4 = item % (C / 4);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (slice < S && item < n_items) {
      const float* src = part_g + (long long)which * C + c4 * 4;
#pragma unroll 4
      for (int ch = slice; ch < k.geo.chunks; ch += S) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (long long)ch * 2 * C));
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      }
    }
    __syncthreads();
    scratch[t * 4 + 0] = a0; scratch[t * 4 + 1] = a1; scratch[t * 4 + 2] = a2; scratch[t * 4 + 3] = a3;
    __syncthreads();
    if (t < per && item < n_items) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += scratch[(s * per + t) * 4 + j];
      float* coef = k.coef + (long long)g * 2 * C;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = c4 * 4 + j;
        if (which == 0) {
          a.dbeta[(long long)g * a.grad_gstride + ch] = float(v[j]);
          coef[ch] = float(v[j] / double(P));
        } else {
          a.dgamma[(long long)g * a.grad_gstride + ch] = float(v[j]);
          coef[C + ch] = float(v[j] / double(P));
        }
      }
    }
  }
}

// Streaming apply: per-channel coefficients (mean, rstd, gamma*rstd, mean(dz), mean(dz*xhat)) of the group in shared
// memory, two items of 8 channels per loop round with all loads issued first (same structure as bn_apply_kernel).
struct BnBwdItem {
  float d[8], y[8];
};

template <bool ADD2, bool MASK>
__device__ __forceinline__ void bn_bwd_item_load(const fb_bn_bwd_args& a, long long off, BnBwdItem& it, bool from_dz) {
  if (from_dz) {  // (instantiated with ADD2 = MASK = false) dz was stored by the reduce pass
    load8(a.dz_out + off, it.d);
    load8(a.y + off, it.y);
    return;
  }
  load8(a.dA + off, it.d);
  if (ADD2) {
    float d2[8];
    load8(a.dA2 + off, d2);
#pragma unroll
    for (int j = 0; j < 8; ++j) it.d[j] += d2[j];
  }
  if (MASK) {
    if (a.mask_bits) {
      const unsigned m = a.mask_bits[off >> 3];
#pragma unroll
      for (int j = 0; j < 8; ++j) it.d[j] = ((m >> j) & 1u) ? it.d[j] : 0.f;
    } else {
      float m[8];
      load8_bf16(static_cast<const bf16*>(a.mask_hi) + off, m);
#pragma unroll
      for (int j = 0; j < 8; ++j) it.d[j] = m[j] > 0.f ? it.d[j] : 0.f;
    }
  }
  load8(a.y + off, it.y);
}

__device__ __forceinline__ void bn_bwd_item_store(const fb_bn_bwd_args& a, long long off, int c, const float* sp,
                                                  const BnBwdItem& it, bool from_dz) {
  const float* rec = sp + (c >> 3) * 44;
  float mu[8], rs[8], grs[8], c1[8], c2[8], o[8];
  load8_smem(rec, mu);
  load8_smem(rec + 8, rs);
  load8_smem(rec + 16, grs);
  load8_smem(rec + 24, c1);
  load8_smem(rec + 32, c2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float xhat = (it.y[j] - mu[j]) * rs[j];
    o[j] = grs[j] * (it.d[j] - c1[j] - xhat * c2[j]);
  }
  store8_bf16(static_cast<bf16*>(a.dy_bf16), off, o);
  if (a.dz_out && !from_dz) store8(a.dz_out + off, it.d);
}

template <bool ADD2, bool MASK>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_kernel(const __grid_constant__ BnBwdK k) {
  // records per channel octet: [mean | rstd | gamma*rstd | mean(dz) | mean(dz*xhat)] x 8, stride 44 floats
  extern __shared__ __align__(16) float sp[];
  griddep_wait();
  griddep_launch();
  const fb_bn_bwd_args& a = k.a;
  const bool backwards = !a.reverse;  // opposite direction of the reduce pass (groups included)
  const int g = backwards ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int C = a.C;
  const int cmask = (C & (C - 1)) == 0 ? C - 1 : 0;
  const float* coef = k.coef + (long long)g * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float rs = a.rstd[(long long)g * C + c];
    float* rec = sp + (c >> 3) * 44 + (c & 7);
    rec[0] = a.mean[(long long)g * C + c];
    rec[8] = rs;
    rec[16] = a.gamma[(long long)g * a.param_gstride + c] * rs;
    rec[24] = coef[c];
    rec[32] = coef[C + c];
  }
  __syncthreads();
  const long long total8 = a.P * C / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long gbase = (long long)g * a.P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < total8;
    const long long e1 = backwards ? total8 - 1 - i : i;
    const long long e2 = backwards ? total8 - 1 - i2 : i2;
    BnBwdItem it1, it2;
    const bool from_dz = k.dz_from_reduce != 0;
    bn_bwd_item_load<ADD2, MASK>(a, gbase + e1 * 8, it1, from_dz);
    if (two) bn_bwd_item_load<ADD2, MASK>(a, gbase + e2 * 8, it2, from_dz);
    bn_bwd_item_store(a, gbase + e1 * 8, channel_of(e1 * 8, C, cmask), sp, it1, from_dz);
    if (two) bn_bwd_item_store(a, gbase + e2 * 8, channel_of(e2 * 8, C, cmask), sp, it2, from_dz);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Running-stat EMA of every BatchNorm layer, in the reference's order (group = microbatch, pass 1 then the FD passes)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_ema_multi_kernel(const fb_bn_ema_entry* __restrict__ table, int n,
                                                           int total_channels, int n_passes, int ng, float momentum) {
  griddep_wait();
  griddep_launch();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_channels) return;
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last entry with c_start <= t
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].c_start <= t) lo = mid; else hi = mid - 1;
  }
  const fb_bn_ema_entry e = table[lo];
  const int c = t - e.c_start;
  float rm = e.running_mean[c], rv = e.running_var[c];
  for (int g = 0; g < ng; ++g)
    for (int p = 0; p < n_passes; ++p) {
      const float* b = e.batch + (long long)p * e.pass_stride + (long long)g * 2 * e.C;
      rm = (1.f - momentum) * rm + momentum * b[c];
      rv = (1.f - momentum) * rv + momentum * b[e.C + c];
    }
  e.running_mean[c] = rm;
  e.running_var[c] = rv;
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
                                   long long first, int cursor_stride, int n, bf16* __restrict__ p_hi,
                                   bf16* __restrict__ p_lo, long long* __restrict__ labels_out) {
  griddep_wait();
  griddep_launch();
  if (first_dev) first += (long long)(*first_dev) * cursor_stride;
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
                                         long long first, int cursor_stride, int n, const char4* __restrict__ aug,
                                         AugNorm nrm, bf16* __restrict__ p_hi, bf16* __restrict__ p_lo,
                                         long long* __restrict__ labels_out) {
  griddep_wait();
  griddep_launch();
  if (first_dev) first += (long long)(*first_dev) * cursor_stride;
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
// head: global average pool -> linear -> label-smoothed cross entropy (+accuracy) and backward, ng groups of n images
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 16;

// grid = ng*n, block = 128
__global__ void __launch_bounds__(128) head_fwd_kernel(const bf16* __restrict__ a_hi, const bf16* __restrict__ a_lo,
                                                       int n, int hw, int c, const float* __restrict__ fc_w_base,
                                                       const float* __restrict__ fc_b_base, long long param_gstride,
                                                       const long long* __restrict__ labels, int classes,
                                                       float smoothing, float* __restrict__ pooled,
                                                       float* __restrict__ dlogits, float* __restrict__ loss_n,
                                                       float* __restrict__ correct_n) {
  griddep_wait();
  griddep_launch();
  extern __shared__ float sp[];  // c floats + classes logits
  float* logits = sp + c;
  const int img = blockIdx.x;
  const int g = img / n;
  const float* fc_w = fc_w_base + (long long)g * param_gstride;
  const float* fc_b = fc_b_base + (long long)g * param_gstride;
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
      dlogits[(long long)img * kMaxClasses + k] = (expf(logp) * wsum - wk) / float(n);
    }
    loss_n[img] = loss;
    correct_n[img] = (arg == label) ? 1.f : 0.f;
  }
}

// grid = (c/128, ng*n): dA[img][p][ch] = (sum_k dlogits[img][k] * W_g[k][ch]) / hw
__global__ void __launch_bounds__(128) head_bwd_act_kernel(const float* __restrict__ dlogits,
                                                           const float* __restrict__ fc_w_base,
                                                           long long param_gstride, int n, int hw, int c, int classes,
                                                           float* __restrict__ dA) {
  griddep_wait();
  griddep_launch();
  const int ch = blockIdx.x * 128 + threadIdx.x;
  const int img = blockIdx.y;
  if (ch >= c) return;
  const float* fc_w = fc_w_base + (long long)(img / n) * param_gstride;
  float s = 0.f;
  for (int k = 0; k < classes; ++k) s += dlogits[(long long)img * kMaxClasses + k] * fc_w[(long long)k * c + ch];
  s /= float(hw);
  for (int p = 0; p < hw; ++p) dA[((long long)img * hw + p) * c + ch] = s;
}

// grid = (c/32, ng) (block x == 0 also reduces bias grad, loss, accuracy of its group); block = 32 channels x 8 sample lanes
__global__ void __launch_bounds__(256) head_bwd_param_kernel(const float* __restrict__ dlogits,
                                                             const float* __restrict__ pooled,
                                                             const float* __restrict__ loss_n,
                                                             const float* __restrict__ correct_n, int n, int c,
                                                             int classes, float* __restrict__ d_fcw_base,
                                                             float* __restrict__ d_fcb_base, long long grad_gstride,
                                                             float* __restrict__ scal, int loss_base,
                                                             int correct_base) {
  griddep_wait();
  griddep_launch();
  __shared__ float red[8][kMaxClasses][33];
  const int g = blockIdx.y;
  const int i0 = g * n;
  float* d_fcw = d_fcw_base + (long long)g * grad_gstride;
  float* d_fcb = d_fcb_base + (long long)g * grad_gstride;
  const int cl = threadIdx.x & 31, lane_n = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  float acc[kMaxClasses];
#pragma unroll
  for (int k = 0; k < kMaxClasses; ++k) acc[k] = 0.f;
  if (ch < c) {
    for (int i = i0 + lane_n; i < i0 + n; i += 8) {
      const float pv = pooled[(long long)i * c + ch];
      const float4* dl = reinterpret_cast<const float4*>(dlogits + (long long)i * kMaxClasses);
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
      for (int i = i0; i < i0 + n; ++i) sum += dlogits[(long long)i * kMaxClasses + threadIdx.x];
      d_fcb[threadIdx.x] = sum;
    } else if (threadIdx.x == 32) {
      double sum = 0.0;
      for (int i = i0; i < i0 + n; ++i) sum += loss_n[i];
      scal[loss_base + g] = float(sum / double(n));
    } else if (threadIdx.x == 64) {
      float sum = 0.f;
      for (int i = i0; i < i0 + n; ++i) sum += correct_n[i];
      scal[correct_base + g] = sum;
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

// BN backward: ~2 blocks per SM over all groups of the launch policy; the partition only depends on (P, C, policy)
static int bwd_geometry(long long P, int C, int policy_groups, BnBwdGeom& geo) {
  if (C % 8 != 0) return FB_ERR_UNSUPPORTED;
  const int c4 = C / 4;
  geo.TX = c4 < 256 ? c4 : 256;
  if (256 % geo.TX != 0 || c4 % geo.TX != 0) return FB_ERR_UNSUPPORTED;
  geo.slabs = c4 / geo.TX;
  const int TY = 256 / geo.TX;
  if (policy_groups < 1) policy_groups = 1;
  long long target = (4LL * kNumSMs) / ((long long)policy_groups * geo.slabs);
  if (target < 1) target = 1;
  long long most = P / (TY * 4);  // >= 4 rows per thread
  if (most < 1) most = 1;
  long long chunks = target < most ? target : most;
  long long rpc = (P + chunks - 1) / chunks;
  rpc = (rpc + TY - 1) / TY * TY;
  geo.rows_per_chunk = int(rpc);
  geo.chunks = int((P + rpc - 1) / rpc);
  return 0;
}

// grid.x of a streaming kernel launched as (grid.x, groups): enough blocks for ~8 resident blocks per SM over all
// groups, but every block runs at least ~8 loop rounds (the per-block prologue -- channel parameters into shared memory
// -- is two dependent memory latencies)
static int stream_grid(long long work_items, int groups = 1) {
  long long blocks = (work_items + 256 * 8 - 1) / (256 * 8);
  long long cap = ((long long)kNumSMs * 8 + groups - 1) / (groups > 0 ? groups : 1);
  if (cap < 1) cap = 1;
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
  bn_reduce_kernel<<<dim3(chunks, slabs), 256, 0, st>>>(y, P, C, TX, rpc, ws);
  bn_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(ws, C, fin);
  FB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fb_bn_apply(const fb_bn_apply_args* a, void* stream) {
  FB_REQUIRE(a && a->y && a->mean && a->rstd && a->gamma && a->beta && a->out_hi, "fb_bn_apply: null pointer");
  FB_REQUIRE(a->C % 8 == 0 && a->P > 0, "fb_bn_apply: C must be a multiple of 8");
  FB_REQUIRE(!a->y2 || (a->mean2 && a->rstd2 && a->gamma2 && a->beta2), "fb_bn_apply: second branch incomplete");
  fb_bn_apply_args k = *a;
  if (k.ng <= 0) k.ng = 1;
  FB_REQUIRE(k.ng <= FB_MAX_GROUPS, "fb_bn_apply: at most %d groups", FB_MAX_GROUPS);
  const size_t smem = size_t(k.y2 ? 52 : 28) * (k.C / 8) * sizeof(float);
  FB_REQUIRE(smem <= 96 * 1024, "fb_bn_apply: at most %d channels", k.y2 ? 3776 : 7008);
  static bool configured = false;
  if (!configured) {  // wide dual-branch layers need more than the default 48 KB of dynamic shared memory
    FB_CUDA(cudaFuncSetAttribute(bn_apply_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    FB_CUDA(cudaFuncSetAttribute(bn_apply_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    FB_CUDA(cudaFuncSetAttribute(bn_apply_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  const dim3 grid(stream_grid(k.P * k.C / 16, k.ng), k.ng);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (k.y2) {
    FB_REQUIRE(!k.res_hi, "fb_bn_apply: a second normalised branch and an identity residual exclude each other");
    FB_CUDA(launch_pdl(bn_apply_kernel<true, false>, grid, dim3(256), smem, st, k));
  } else if (k.res_hi) {
    FB_CUDA(launch_pdl(bn_apply_kernel<false, true>, grid, dim3(256), smem, st, k));
  } else {
    FB_CUDA(launch_pdl(bn_apply_kernel<false, false>, grid, dim3(256), smem, st, k));
  }
  return 0;
}

extern "C" int fb_bn_bwd_chunks(int64_t P, int C, int policy_groups) {
  BnBwdGeom geo;
  if (bwd_geometry(P, C, policy_groups, geo)) return -1;
  return geo.chunks;
}

extern "C" int fb_bn_bwd(const fb_bn_bwd_args* a, void* stream) {
  FB_REQUIRE(a && a->dA && a->y && a->mean && a->rstd && a->gamma && a->ws && a->dgamma && a->dbeta && a->dy_bf16,
             "fb_bn_bwd: null pointer");
  FB_REQUIRE(a->C % 8 == 0 && a->P > 0, "fb_bn_bwd: C must be a multiple of 8");
  BnBwdK k;
  k.a = *a;
  if (k.a.ng <= 0) k.a.ng = 1;
  FB_REQUIRE(k.a.ng <= FB_MAX_GROUPS, "fb_bn_bwd: at most %d groups", FB_MAX_GROUPS);
  if (bwd_geometry(a->P, a->C, a->policy_groups > 0 ? a->policy_groups : k.a.ng, k.geo)) {
    set_error("fb_bn_bwd: unsupported channel count %d", a->C);
    return FB_ERR_UNSUPPORTED;
  }
  k.tickets = reinterpret_cast<unsigned int*>(a->ws);
  k.partial = a->ws + 16;
  k.coef = k.partial + (long long)k.a.ng * k.geo.chunks * 2 * a->C;
  {
    static const int mode = [] {
      const char* e = getenv("FB_DZ_FROM_REDUCE");  // 0: the apply pass recomputes dz; 2: also with a single addend
      return e ? atoi(e) : 1;
    }();
    // dz_out must not alias an input another thread still reads: it is only ever read/written element-wise by the thread
    // that owns the element, so aliasing dA / dA2 would be fine too
    k.dz_from_reduce = (mode && a->dz_out && (a->dA2 || mode > 1)) ? 1 : 0;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FB_CUDA(launch_pdl(bn_bwd_reduce_kernel, dim3(k.geo.chunks, k.geo.slabs, k.a.ng), dim3(256), 0, st, k));
  const size_t smem = size_t(44) * (a->C / 8) * sizeof(float);
  FB_REQUIRE(smem <= 48 * 1024, "fb_bn_bwd: at most 2232 channels");
  const dim3 grid(stream_grid(a->P * a->C / 16, k.a.ng), k.a.ng);
  const bool masked = a->mask_hi || a->mask_bits;
  if (k.dz_from_reduce)
    FB_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false>, grid, dim3(256), smem, st, k));
  else if (a->dA2 && masked)
    FB_CUDA(launch_pdl(bn_bwd_apply_kernel<true, true>, grid, dim3(256), smem, st, k));
  else if (a->dA2)
    FB_CUDA(launch_pdl(bn_bwd_apply_kernel<true, false>, grid, dim3(256), smem, st, k));
  else if (masked)
    FB_CUDA(launch_pdl(bn_bwd_apply_kernel<false, true>, grid, dim3(256), smem, st, k));
  else
    FB_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false>, grid, dim3(256), smem, st, k));
  return 0;
}

extern "C" int fb_bn_ema_multi(const fb_bn_ema_entry* table_dev, int n_entries, int total_channels, int n_passes, int ng,
                               float momentum, void* stream) {
  FB_REQUIRE(table_dev && n_entries > 0 && total_channels > 0 && n_passes >= 1 && ng >= 1 && ng <= FB_MAX_GROUPS,
             "fb_bn_ema_multi: bad arguments");
  FB_CUDA(launch_pdl(bn_ema_multi_kernel, dim3((total_channels + 255) / 256), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), table_dev, n_entries, total_channels, n_passes, ng, momentum));
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
                              int64_t first, int cursor_stride, int n, void* patches_hi, void* patches_lo,
                              int64_t* labels_out, void* stream) {
  FB_REQUIRE(x && patches_hi && n > 0, "fb_stem_im2col: bad arguments");
  FB_REQUIRE(!labels_out || labels, "fb_stem_im2col: labels_out needs labels");
  FB_CUDA(launch_pdl(stem_im2col_kernel, dim3(stream_grid((long long)n * 1024 * 8)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), x, reinterpret_cast<const long long*>(labels),
                     reinterpret_cast<const long long*>(perm), first_dev, (long long)first, cursor_stride, n,
                     static_cast<bf16*>(patches_hi), static_cast<bf16*>(patches_lo),
                     reinterpret_cast<long long*>(labels_out)));
  return 0;
}

extern "C" int fb_stem_im2col_u8aug(const uint8_t* x_hwc, const int64_t* labels, const int64_t* perm,
                                    const int32_t* first_dev, int64_t first, int cursor_stride, int n,
                                    const int8_t* aug, const float* mean3, const float* std3, void* patches_hi,
                                    void* patches_lo, int64_t* labels_out, void* stream) {
  FB_REQUIRE(x_hwc && patches_hi && mean3 && std3 && n > 0, "fb_stem_im2col_u8aug: bad arguments");
  FB_REQUIRE(!labels_out || labels, "fb_stem_im2col_u8aug: labels_out needs labels");
  AugNorm nrm;
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = mean3[c];
    nrm.inv_std[c] = 1.f / std3[c];
  }
  FB_CUDA(launch_pdl(stem_im2col_u8aug_kernel, dim3(stream_grid((long long)n * 1024 * 8)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream), x_hwc, reinterpret_cast<const long long*>(labels),
                     reinterpret_cast<const long long*>(perm), first_dev, (long long)first, cursor_stride, n,
                     reinterpret_cast<const char4*>(aug), nrm, static_cast<bf16*>(patches_hi),
                     static_cast<bf16*>(patches_lo), reinterpret_cast<long long*>(labels_out)));
  return 0;
}

extern "C" int fb_head_fwd_bwd(const void* a_hi, const void* a_lo, int n, int hw, int c, const float* fc_w,
                               const float* fc_b, const int64_t* labels, int classes, float smoothing, float* ws,
                               float* scal, int loss_base, int correct_base, float* d_fcw, float* d_fcb, float* dA,
                               int ng, int64_t param_gstride, int64_t grad_gstride, void* stream) {
  FB_REQUIRE(a_hi && fc_w && fc_b && labels && ws && scal && d_fcw && d_fcb && dA, "fb_head_fwd_bwd: null pointer");
  if (classes > kMaxClasses || classes < 2) {
    set_error("fb_head_fwd_bwd: classes %d not supported (max %d)", classes, kMaxClasses);
    return FB_ERR_UNSUPPORTED;
  }
  if (ng <= 0) ng = 1;
  FB_REQUIRE(ng <= FB_MAX_GROUPS, "fb_head_fwd_bwd: at most %d groups", FB_MAX_GROUPS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nt = (long long)ng * n;
  float* pooled = ws;
  float* dlogits = pooled + nt * c;
  float* loss_n = dlogits + nt * kMaxClasses;
  float* correct_n = loss_n + nt;
  FB_CUDA(launch_pdl(head_fwd_kernel, dim3((unsigned)nt), dim3(128), (c + kMaxClasses) * sizeof(float), st,
                     static_cast<const bf16*>(a_hi), static_cast<const bf16*>(a_lo), n, hw, c, fc_w, fc_b,
                     (long long)param_gstride, reinterpret_cast<const long long*>(labels), classes, smoothing, pooled,
                     dlogits, loss_n, correct_n));
  FB_CUDA(launch_pdl(head_bwd_act_kernel, dim3((c + 127) / 128, (unsigned)nt), dim3(128), 0, st, (const float*)dlogits,
                     fc_w, (long long)param_gstride, n, hw, c, classes, dA));
  FB_CUDA(launch_pdl(head_bwd_param_kernel, dim3((c + 31) / 32, ng), dim3(256), 0, st, (const float*)dlogits,
                     (const float*)pooled, (const float*)loss_n, (const float*)correct_n, n, c, classes, d_fcw, d_fcb,
                     (long long)grad_gstride, scal, loss_base, correct_base));
  return 0;
}
