// Flat-buffer ("multi-tensor") kernels of the full-batch step (sm_100a).
//
// All parameters / gradients live in persistent flat fp32 buffers in model.parameters() order (the order of
// fullbatch/training/utils.py:34), one gradient buffer per microbatch group of a launch, so every list-of-tensors
// `torch._foreach_*` sweep of the reference becomes one vectorised, coalesced pass over `ng` groups:
//   fb_flat_sqnorm    : sum g^2 per group, eps_n per group      (training.py:162, modules.py:223: 62 pow/sum kernels)
//   fb_perturb_ranges : theta' = theta + eps_n*v for the non-conv parameters (modules.py:215-226; conv weights are
//                       perturbed inside fb_weight_prep_multi, theta' is never materialised for them)
//   fb_fd_combine     : g += cf*(g2-g)/eps_n ; avg += (g-avg)/k, groups in loader order  (modules.py:232-240 +
//                       training.py:45-47,165-168)
// eps_n and the norms stay on the device (the reference syncs the host once per microbatch through alpha=eps_n).
// Conv weight gradients are in the kernels' native [co][tap][ci] layout; fb_flat_relayout converts at the boundary.
#include "../../include/fullbatch_b200.h"
#include "fb_common.cuh"

namespace fb {

constexpr int kSqBlocks = 1024;

// grid = (kSqBlocks, ng)
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* __restrict__ x, long long x_gstride,
                                                             const float* __restrict__ y, float a, float b,
                                                             long long n, double* __restrict__ partial) {
  griddep_wait();
  griddep_launch();
  __shared__ double red[8];
  const int g = blockIdx.y;
  const float* xg = x + (long long)g * x_gstride;
  float acc = 0.f;
  if (!y && a == 1.f) {
    const long long n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(xg);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
      const float4 v = x4[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < int(n - n4 * 4)) {
      const float v = xg[n4 * 4 + threadIdx.x];
      acc += v * v;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
      const float v = a * xg[i] + (y ? b * y[i] : 0.f);
      acc += v * v;
    }
  }
  double d = warp_sum(double(acc));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[(long long)g * gridDim.x + blockIdx.x] = s;
  }
}

// grid = ng
__global__ void __launch_bounds__(256) sqnorm_final_kernel(const double* __restrict__ partial, int nblocks,
                                                           float* __restrict__ scal, int slot_base,
                                                           float* __restrict__ norms_out,
                                                           const int* __restrict__ cursor, int eps_mode, float bs,
                                                           float eps, int eps_base) {
  griddep_wait();
  griddep_launch();
  __shared__ double red[8];
  const int g = blockIdx.x;
  double d = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) d += partial[(long long)g * nblocks + i];
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    const float n2 = float(s);
    scal[slot_base + g] = n2;
    if (norms_out) norms_out[(cursor ? *cursor : 0) + g] = n2;
    if (eps_mode == 1) scal[eps_base + g] = eps / sqrtf(bs * bs * n2);  // modules.py:223 with v = bs*g
    if (eps_mode == 2) scal[eps_base + g] = eps / sqrtf(n2);
  }
}

// grid = (blocks, ng); thread -> element of the concatenated ranges (binary search over the first-thread column)
__global__ void __launch_bounds__(256) perturb_ranges_kernel(const float* __restrict__ theta,
                                                             const float* __restrict__ grad, long long grad_gstride,
                                                             const float* __restrict__ pre,
                                                             const long long* __restrict__ ranges, int n_ranges,
                                                             long long total, float bs, float acc, float scale,
                                                             const float* __restrict__ scal, int eps_base,
                                                             float* __restrict__ theta_p, long long theta_p_gstride) {
  griddep_wait();
  griddep_launch();
  const int g = blockIdx.y;
  const float step = scale * scal[eps_base + g];
  const float* gg = grad + (long long)g * grad_gstride;
  float* out = theta_p + (long long)g * theta_p_gstride;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_ranges - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (ranges[3 * mid + 2] <= t) lo = mid; else hi = mid - 1;
    }
    const long long i = ranges[3 * lo] + (t - ranges[3 * lo + 2]);
    float v = bs * gg[i];
    if (pre && acc != 0.f) v += acc * pre[i];
    out[i] = theta[i] + step * v;
  }
}

// The two sweeps below read every group's gradient once and keep the running mean in registers.  Per element the groups
// are combined in loader order (a sequential chain, as in the reference), but the LOADS of four groups are issued
// together (16-byte vectors), so enough bytes are in flight to stream at HBM rate.
constexpr int kGroupChunk = 4;

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float& comp(float4& v, int j) { return (&v.x)[j]; }

template <bool MINUS>
__global__ void __launch_bounds__(256) fd_combine_kernel(float* __restrict__ grad, const float* __restrict__ g_plus,
                                                         const float* __restrict__ g_minus, long long gstride,
                                                         float* __restrict__ avg, long long n, int ng,
                                                         const float* __restrict__ scal, int eps_base, int cf_slot,
                                                         const int* __restrict__ cursor, int write_g) {
  griddep_wait();
  griddep_launch();
  __shared__ float s_eps[FB_MAX_GROUPS], s_inv[FB_MAX_GROUPS];
  const float cf = scal[cf_slot];
  if (threadIdx.x < FB_MAX_GROUPS) {
    const int count0 = cursor ? *cursor : 0;
    s_eps[threadIdx.x] = (int)threadIdx.x < ng ? scal[eps_base + threadIdx.x] : 1.f;
    s_inv[threadIdx.x] = float(1.0 / double(count0 + (int)threadIdx.x + 1));
  }
  __syncthreads();
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = avg ? ldg4(avg + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g0 = 0; g0 < ng; g0 += kGroupChunk) {
      float4 gr[kGroupChunk], gp[kGroupChunk], gm[kGroupChunk];
#pragma unroll
      for (int j = 0; j < kGroupChunk; ++j) {
        if (g0 + j < ng) {
          const long long o = (long long)(g0 + j) * gstride + 4 * i;
          gr[j] = ldg4(grad + o);
          gp[j] = ldg4(g_plus + o);
          if (MINUS) gm[j] = ldg4(g_minus + o);
        }
      }
#pragma unroll
      for (int j = 0; j < kGroupChunk; ++j) {
        if (g0 + j < ng) {
          const float eps_n = s_eps[g0 + j], inv = s_inv[g0 + j];
          float4 r;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float g = comp(gr[j], c);
            const float base = MINUS ? comp(gm[j], c) : g;
            const float h = (comp(gp[j], c) - base) / eps_n;  // modules.py:232-234 / :292-293
            const float v = g + cf * h;                        // modules.py:240 / :299
            comp(r, c) = v;
            comp(a, c) = comp(a, c) + (v - comp(a, c)) * inv;  // training.py:45-47
          }
          if (write_g) *reinterpret_cast<float4*>(grad + (long long)(g0 + j) * gstride + 4 * i) = r;
        }
      }
    }
    if (avg) *reinterpret_cast<float4*>(avg + 4 * i) = a;
  }
  // tail (n % 4 elements)
  if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n - n4 * 4)) {
    const long long i = n4 * 4 + threadIdx.x;
    float a = avg ? avg[i] : 0.f;
    for (int g = 0; g < ng; ++g) {
      const long long o = (long long)g * gstride + i;
      float gr = grad[o];
      const float base = MINUS ? g_minus[o] : gr;
      const float h = (g_plus[o] - base) / s_eps[g];
      gr = gr + cf * h;
      if (write_g) grad[o] = gr;
      a = a + (gr - a) * s_inv[g];
    }
    if (avg) avg[i] = a;
  }
}

__global__ void __launch_bounds__(256) mean_accumulate_kernel(float* __restrict__ grad, long long gstride,
                                                              float* __restrict__ avg, long long n, int ng,
                                                              const int* __restrict__ cursor, float* __restrict__ scal,
                                                              int norm_base, float clip, int clipped_slot) {
  griddep_wait();
  griddep_launch();
  __shared__ float s_coef[FB_MAX_GROUPS], s_inv[FB_MAX_GROUPS];
  __shared__ int s_clipped[FB_MAX_GROUPS];
  if (threadIdx.x < FB_MAX_GROUPS) {
    const int g = threadIdx.x;
    const int count0 = cursor ? *cursor : 0;
    float coef = 1.f;
    int clipped = 0;
    if (g < ng && clip > 0.f) {
      const float norm = sqrtf(scal[norm_base + g]);
      if (norm > clip) {  // training/utils.py:4-19
        coef = clip / (norm + 1e-6f);
        clipped = 1;
      }
    }
    s_coef[g] = coef;
    s_clipped[g] = clipped;
    s_inv[g] = float(1.0 / double(count0 + g + 1));
  }
  __syncthreads();
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = ldg4(avg + 4 * i);
    for (int g0 = 0; g0 < ng; g0 += kGroupChunk) {
      float4 gr[kGroupChunk];
#pragma unroll
      for (int j = 0; j < kGroupChunk; ++j)
        if (g0 + j < ng) gr[j] = ldg4(grad + (long long)(g0 + j) * gstride + 4 * i);
#pragma unroll
      for (int j = 0; j < kGroupChunk; ++j) {
        if (g0 + j < ng) {
          const float coef = s_coef[g0 + j], inv = s_inv[g0 + j];
          float4 r;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float v = comp(gr[j], c) * coef;
            comp(r, c) = v;
            comp(a, c) = comp(a, c) + (v - comp(a, c)) * inv;
          }
          // the reference scales the microbatch gradient in place
          if (s_clipped[g0 + j]) *reinterpret_cast<float4*>(grad + (long long)(g0 + j) * gstride + 4 * i) = r;
        }
      }
    }
    *reinterpret_cast<float4*>(avg + 4 * i) = a;
  }
  if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n - n4 * 4)) {
    const long long i = n4 * 4 + threadIdx.x;
    float a = avg[i];
    for (int g = 0; g < ng; ++g) {
      const long long o = (long long)g * gstride + i;
      const float v = grad[o] * s_coef[g];
      if (s_clipped[g]) grad[o] = v;
      a = a + (v - a) * s_inv[g];
    }
    avg[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int n_clipped = 0;
    for (int g = 0; g < ng; ++g) n_clipped += s_clipped[g];
    if (n_clipped) scal[clipped_slot] += float(n_clipped);
  }
}

__global__ void group_finish_kernel(int* cursor, int ng, int cursor_step, const float* scal, float* totals, int loss_slot,
                                    int correct_slot, int loss_base, int correct_base) {
  griddep_wait();
  griddep_launch();
  float loss = totals[loss_slot], correct = totals[correct_slot];
  for (int g = 0; g < ng; ++g) {  // training.py:172-173, loader order
    loss += scal[loss_base + g];
    correct += scal[correct_base + g];
  }
  totals[loss_slot] = loss;
  totals[correct_slot] = correct;
  *cursor += cursor_step;
}

__global__ void __launch_bounds__(256) flat_scale_kernel(float* __restrict__ x, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= alpha;
}

// One block per (3x3 conv, co): [ci][tap] <-> [tap][ci] through shared memory; a leading grid-stride copy moves
// everything else.  table: (offset, cout, cin, taps, first block) per conv.
__global__ void __launch_bounds__(256) flat_relayout_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            long long n, const long long* __restrict__ table,
                                                            int n_entries, int conv_blocks, int to_native) {
  extern __shared__ float tile[];
  if ((int)blockIdx.x >= conv_blocks) {  // the remaining blocks copy everything that is not a permuted conv weight
    const long long nb = gridDim.x - conv_blocks;
    for (long long i = (long long)(blockIdx.x - conv_blocks) * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
      // elements of a permuted conv are written by that conv's own blocks
      int lo = 0, hi = n_entries - 1;
      bool in_conv = false;
      if (n_entries > 0 && i >= table[0]) {
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (table[5 * mid] <= i) lo = mid; else hi = mid - 1;
        }
        in_conv = i < table[5 * lo] + table[5 * lo + 1] * table[5 * lo + 2] * table[5 * lo + 3];
      }
      if (!in_conv) dst[i] = src[i];
    }
    return;
  }
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[5 * mid + 4] <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const long long off = table[5 * lo];
  const int cin = int(table[5 * lo + 2]), taps = int(table[5 * lo + 3]);
  const int co = int(blockIdx.x - table[5 * lo + 4]);
  const long long base = off + (long long)co * cin * taps;
  const int len = cin * taps;
  for (int i = threadIdx.x; i < len; i += blockDim.x) tile[i] = src[base + i];
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    // to_native: dst[tap][ci] = src[ci][tap]; else dst[ci][tap] = src[tap][ci]
    int s;
    if (to_native) {
      const int tap = i / cin, ci = i % cin;
      s = ci * taps + tap;
    } else {
      const int ci = i / taps, tap = i % taps;
      s = tap * cin + ci;
    }
    dst[base + i] = tile[s];
  }
}

static int flat_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 8;
  return int(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace fb

using namespace fb;

extern "C" int fb_flat_sqnorm(const float* x, int64_t x_gstride, const float* y, float a, float b, int64_t n, int ng,
                              double* ws, float* scal, int slot_base, float* norms_out, const int32_t* cursor,
                              int eps_mode, float bs, float eps, int eps_base, void* stream) {
  FB_REQUIRE(x && ws && scal && n > 0 && ng >= 1 && ng <= FB_MAX_GROUPS, "fb_flat_sqnorm: bad arguments");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && x_gstride % 4 == 0,
             "fb_flat_sqnorm: x and its group stride must be 16-byte aligned");
  FB_REQUIRE(eps_mode >= 0 && eps_mode <= 2, "fb_flat_sqnorm: eps_mode must be 0, 1 or 2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FB_CUDA(launch_pdl(sqnorm_partial_kernel, dim3(kSqBlocks, ng), dim3(256), 0, st, x, (long long)x_gstride, y, a, b,
                     (long long)n, ws));
  FB_CUDA(launch_pdl(sqnorm_final_kernel, dim3(ng), dim3(256), 0, st, (const double*)ws, kSqBlocks, scal, slot_base,
                     norms_out, (const int*)cursor, eps_mode, bs, eps, eps_base));
  return 0;
}

extern "C" int fb_perturb_ranges(const float* theta, const float* grad, int64_t grad_gstride, const float* pre,
                                 const int64_t* ranges_dev, int n_ranges, int64_t total, float bs, float acc,
                                 float scale, const float* scal, int eps_base, float* theta_p, int64_t theta_p_gstride,
                                 int ng, void* stream) {
  FB_REQUIRE(theta && grad && ranges_dev && scal && theta_p && n_ranges > 0 && total > 0 && ng >= 1 &&
                 ng <= FB_MAX_GROUPS,
             "fb_perturb_ranges: bad arguments");
  FB_CUDA(launch_pdl(perturb_ranges_kernel, dim3(flat_grid(total), ng), dim3(256), 0, static_cast<cudaStream_t>(stream),
                     theta, grad, (long long)grad_gstride, pre, reinterpret_cast<const long long*>(ranges_dev), n_ranges,
                     (long long)total, bs, acc, scale, scal, eps_base, theta_p, (long long)theta_p_gstride));
  return 0;
}

extern "C" int fb_fd_combine(float* grad, const float* g_plus, const float* g_minus, int64_t gstride, float* avg,
                             int64_t n, int ng, const float* scal, int eps_base, int cf_slot, const int32_t* cursor,
                             int write_g, void* stream) {
  FB_REQUIRE(grad && g_plus && scal && n > 0 && ng >= 1 && ng <= FB_MAX_GROUPS && cf_slot >= 0,
             "fb_fd_combine: bad arguments");
  FB_REQUIRE(((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(g_plus) |
               reinterpret_cast<uintptr_t>(g_minus) | reinterpret_cast<uintptr_t>(avg)) & 15) == 0 && gstride % 4 == 0,
             "fb_fd_combine: buffers and the group stride must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(flat_grid(n / 4 + 1));
  if (g_minus)
    FB_CUDA(launch_pdl(fd_combine_kernel<true>, grid, dim3(256), 0, st, grad, g_plus, g_minus, (long long)gstride, avg,
                       (long long)n, ng, scal, eps_base, cf_slot, (const int*)cursor, write_g));
  else
    FB_CUDA(launch_pdl(fd_combine_kernel<false>, grid, dim3(256), 0, st, grad, g_plus, g_minus, (long long)gstride, avg,
                       (long long)n, ng, scal, eps_base, cf_slot, (const int*)cursor, write_g));
  return 0;
}

extern "C" int fb_mean_accumulate(float* grad, int64_t gstride, float* avg, int64_t n, int ng, const int32_t* cursor,
                                  float* scal, int norm_base, float clip, int clipped_slot, void* stream) {
  FB_REQUIRE(grad && avg && n > 0 && ng >= 1 && ng <= FB_MAX_GROUPS, "fb_mean_accumulate: bad arguments");
  FB_REQUIRE(clip <= 0.f || scal, "fb_mean_accumulate: clipping needs scal");
  FB_REQUIRE(((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(avg)) & 15) == 0 && gstride % 4 == 0,
             "fb_mean_accumulate: buffers and the group stride must be 16-byte aligned");
  FB_CUDA(launch_pdl(mean_accumulate_kernel, dim3(flat_grid(n / 4 + 1)), dim3(256), 0, static_cast<cudaStream_t>(stream), grad,
                     (long long)gstride, avg, (long long)n, ng, (const int*)cursor, scal, norm_base, clip,
                     clipped_slot));
  return 0;
}

extern "C" int fb_group_finish(int32_t* cursor, int ng, int cursor_step, const float* scal, float* totals, int loss_slot,
                               int correct_slot, int loss_base, int correct_base, void* stream) {
  FB_REQUIRE(cursor && scal && totals && ng >= 1 && ng <= FB_MAX_GROUPS && cursor_step >= 0,
             "fb_group_finish: bad arguments");
  FB_CUDA(launch_pdl(group_finish_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), (int*)cursor, ng,
                     cursor_step, scal, totals, loss_slot, correct_slot, loss_base, correct_base));
  return 0;
}

extern "C" int fb_flat_scale(float* x, int64_t n, float alpha, void* stream) {
  FB_REQUIRE(x && n > 0, "fb_flat_scale: bad arguments");
  flat_scale_kernel<<<flat_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, alpha);
  FB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fb_flat_relayout(const float* src, float* dst, int64_t n, const int64_t* table_dev, int n_entries,
                                int total_blocks, int to_native, void* stream) {
  FB_REQUIRE(src && dst && src != dst && n > 0 && (n_entries == 0 || table_dev) && total_blocks >= 0,
             "fb_flat_relayout: bad arguments");
  const int copy_blocks = flat_grid(n);
  const size_t smem = size_t(512) * 9 * sizeof(float);  // cin <= 512 for 3x3 convs of the ResNet family
  flat_relayout_kernel<<<total_blocks + copy_blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      src, dst, (long long)n, reinterpret_cast<const long long*>(table_dev), n_entries, total_blocks, to_native);
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The step right after the path: global-norm clip (training.py:198-211) + torch.optim.SGD with
// weight decay / momentum / dampening / Nesterov (optimizers.py:25-28) + sum theta^2 for _record_stats (training.py:92)
// as ONE sweep over the flat buffers.  The clip coefficient is derived on the device from scal[norm_slot] (= |g|^2
// from fb_flat_sqnorm), so the optimizer step needs no host synchronisation.
//   d = g * coef + wd * theta ; buf = first ? d : momentum * buf + (1 - dampening) * d ; d = nesterov ? d + momentum*buf
//   : buf ; theta -= lr * d            (torch/optim/sgd.py, _single_tensor_sgd, same operation order)
// ---------------------------------------------------------------------------------------------------------------
namespace fb {
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ theta, float* __restrict__ g,
                                                       float* __restrict__ buf, long long n,
                                                       const float* __restrict__ scal, int norm_slot, float clip,
                                                       float lr, float momentum, float dampening, float wd,
                                                       int nesterov, int first, int write_clipped_grad,
                                                       double* __restrict__ partial) {
  __shared__ double red[8];
  float coef = 1.f;
  if (clip > 0.f) {
    const float norm = sqrtf(scal[norm_slot]);
    if (norm > clip) coef = clip / (norm + 1e-6f);  // training.py:207
  }
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = theta[i];
    float d = g[i] * coef;
    if (write_clipped_grad) g[i] = d;  // param.grad is clipped in place by the reference
    if (wd != 0.f) d = d + wd * t;
    if (momentum != 0.f) {
      float b = first ? d : momentum * buf[i] + (1.f - dampening) * d;
      buf[i] = b;
      d = nesterov ? d + momentum * b : b;
    }
    t = t - lr * d;
    theta[i] = t;
    acc += t * t;
  }
  double s = warp_sum(double(acc));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    partial[blockIdx.x] = tot;
  }
}
}  // namespace fb

extern "C" int fb_sgd_step(float* theta, float* grad, float* momentum_buf, int64_t n, float* scal, int norm_slot,
                           float clip, float lr, float momentum, float dampening, float weight_decay, int nesterov,
                           int first_step, int write_clipped_grad, double* ws, int param_norm_slot, void* stream) {
  FB_REQUIRE(theta && grad && scal && ws && n > 0, "fb_sgd_step: bad arguments");
  FB_REQUIRE(momentum == 0.f || momentum_buf, "fb_sgd_step: momentum needs a buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = flat_grid(n) < kSqBlocks ? flat_grid(n) : kSqBlocks;
  sgd_step_kernel<<<blocks, 256, 0, st>>>(theta, grad, momentum_buf, n, scal, norm_slot, clip, lr, momentum, dampening,
                                          weight_decay, nesterov, first_step, write_clipped_grad, ws);
  sqnorm_final_kernel<<<1, 256, 0, st>>>(ws, blocks, scal, param_norm_slot, nullptr, nullptr, 0, 0.f, 0.f, 0);
  FB_CUDA(cudaGetLastError());
  return 0;
}
