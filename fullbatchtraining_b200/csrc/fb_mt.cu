// Flat-buffer ("multi-tensor") kernels of the full-batch step (sm_100a).
//
// All parameters / gradients live in persistent flat fp32 buffers in model.parameters() order (the order of
// fullbatch/training/utils.py:34), so every list-of-tensors `torch._foreach_*` sweep of the reference becomes one
// vectorised, coalesced pass:
//   fb_flat_sqnorm   : sum g^2                         (training.py:162, modules.py:223: 62 pow/sum kernels + stack)
//   fb_fd_perturb    : eps_n, theta' = theta + eps_n*bs*g   (modules.py:215-226: clone + mul + add_)
//   fb_fd_combine    : g += cf*(g2-g)/eps_n ; avg += (g-avg)/k   (modules.py:232-240 + training.py:45-47,165-168)
// eps_n and the norms stay on the device (the reference syncs the host once per microbatch through alpha=eps_n).
#include "../../include/fullbatch_b200.h"
#include "fb_common.cuh"

namespace fb {

constexpr int kSqBlocks = 1024;

__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* __restrict__ x, long long n,
                                                             double* __restrict__ partial) {
  griddep_wait();
  griddep_launch();
  __shared__ double red[8];
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < int(n - n4 * 4)) {
    const float v = x[n4 * 4 + threadIdx.x];
    acc += v * v;
  }
  double d = warp_sum(double(acc));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) sqnorm_final_kernel(const double* __restrict__ partial, int nblocks,
                                                           float* __restrict__ scal, int slot,
                                                           float* __restrict__ norms_out,
                                                           const int* __restrict__ cursor) {
  griddep_wait();
  griddep_launch();
  __shared__ double red[8];
  double d = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) d += partial[i];
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    scal[slot] = float(s);
    if (norms_out) norms_out[cursor ? *cursor : 0] = float(s);
  }
}

__global__ void __launch_bounds__(256) fd_perturb_kernel(const float* __restrict__ theta, const float* __restrict__ g,
                                                         long long n, float bs, float eps, float* __restrict__ scal,
                                                         int sq_slot, int eps_slot, float* __restrict__ theta_p) {
  griddep_wait();
  griddep_launch();
  const float n2 = scal[sq_slot];
  // modules.py:223: eps / sqrt(sum (bs*g)^2)
  const float eps_n = eps / sqrtf(bs * bs * n2);
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[eps_slot] = eps_n;
  const long long n4 = n / 4;
  const float4* t4 = reinterpret_cast<const float4*>(theta);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* o4 = reinterpret_cast<float4*>(theta_p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 t = t4[i], v = g4[i];
    o4[i] = make_float4(t.x + eps_n * (bs * v.x), t.y + eps_n * (bs * v.y), t.z + eps_n * (bs * v.z),
                        t.w + eps_n * (bs * v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < int(n - n4 * 4)) {
    const long long i = n4 * 4 + threadIdx.x;
    theta_p[i] = theta[i] + eps_n * (bs * g[i]);
  }
}

template <bool REG>
__global__ void __launch_bounds__(256) fd_combine_kernel(float* __restrict__ g, const float* __restrict__ g2,
                                                         float* __restrict__ avg, long long n,
                                                         const float* __restrict__ scal, int eps_slot, float cf,
                                                         int cf_slot, const int* __restrict__ cursor, int count0,
                                                         int write_g) {
  griddep_wait();
  griddep_launch();
  const float eps_n = REG ? scal[eps_slot] : 1.f;
  if (REG && cf_slot >= 0) cf = scal[cf_slot];
  const int count = count0 + (cursor ? *cursor : 0) + 1;
  const float inv = float(1.0 / double(count));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gr = g[i];
    if (REG) {
      const float h = (g2[i] - gr) / eps_n;  // modules.py:232-234
      gr = gr + cf * h;                       // modules.py:240
      if (write_g) g[i] = gr;
    }
    if (avg) {
      const float a = avg[i];
      avg[i] = a + (gr - a) * inv;  // training.py:45-47
    }
  }
}

__global__ void cursor_add_kernel(int* cursor, int delta) {
  griddep_wait();
  griddep_launch();
  *cursor += delta;
}

__global__ void __launch_bounds__(256) flat_scale_kernel(float* __restrict__ x, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= alpha;
}

static int flat_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 8;
  return int(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace fb

using namespace fb;

extern "C" int fb_flat_sqnorm(const float* x, int64_t n, double* ws, float* scal, int slot, float* norms_out,
                              const int32_t* cursor, void* stream) {
  FB_REQUIRE(x && ws && scal && n > 0, "fb_flat_sqnorm: bad arguments");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "fb_flat_sqnorm: x must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FB_CUDA(launch_pdl(sqnorm_partial_kernel, dim3(kSqBlocks), dim3(256), 0, st, x, (long long)n, ws));
  FB_CUDA(launch_pdl(sqnorm_final_kernel, dim3(1), dim3(256), 0, st, (const double*)ws, kSqBlocks, scal, slot, norms_out,
                     (const int*)cursor));
  return 0;
}

extern "C" int fb_fd_perturb(const float* theta, const float* g, int64_t n, float block_strength, float eps, float* scal,
                             int sq_slot, int eps_slot, float* theta_p, void* stream) {
  FB_REQUIRE(theta && g && scal && theta_p && n > 0, "fb_fd_perturb: bad arguments");
  FB_REQUIRE(((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(g) |
               reinterpret_cast<uintptr_t>(theta_p)) & 15) == 0,
             "fb_fd_perturb: buffers must be 16-byte aligned");
  FB_CUDA(launch_pdl(fd_perturb_kernel, dim3(flat_grid(n / 4 + 1)), dim3(256), 0, static_cast<cudaStream_t>(stream), theta,
                     g, (long long)n, block_strength, eps, scal, sq_slot, eps_slot, theta_p));
  return 0;
}

extern "C" int fb_fd_combine(float* g, const float* g2, float* avg, int64_t n, const float* scal, int eps_slot, float cf,
                             int cf_slot, const int32_t* cursor, int32_t count0, int write_g, void* stream) {
  FB_REQUIRE(g && g2 && scal && n > 0, "fb_fd_combine: bad arguments");
  FB_CUDA(launch_pdl(fd_combine_kernel<true>, dim3(flat_grid(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), g, g2,
                     avg, (long long)n, scal, eps_slot, cf, cf_slot, (const int*)cursor, (int)count0, write_g));
  return 0;
}

extern "C" int fb_mean_accumulate(const float* g, float* avg, int64_t n, const int32_t* cursor, int32_t count0,
                                  void* stream) {
  FB_REQUIRE(g && avg && n > 0, "fb_mean_accumulate: bad arguments");
  FB_CUDA(launch_pdl(fd_combine_kernel<false>, dim3(flat_grid(n)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                     const_cast<float*>(g), (const float*)nullptr, avg, (long long)n, (const float*)nullptr, 0, 0.f, -1,
                     (const int*)cursor, (int)count0, 0));
  return 0;
}

extern "C" int fb_cursor_add(int32_t* cursor, int32_t delta, void* stream) {
  FB_REQUIRE(cursor, "fb_cursor_add: null pointer");
  FB_CUDA(launch_pdl(cursor_add_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), (int*)cursor, (int)delta));
  return 0;
}

extern "C" int fb_flat_scale(float* x, int64_t n, float alpha, void* stream) {
  FB_REQUIRE(x && n > 0, "fb_flat_scale: bad arguments");
  flat_scale_kernel<<<flat_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, alpha);
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

extern "C" int fb_sgd_step(float* theta, float* grad, float* momentum_buf, int64_t n, float* scal, int norm_slot,
                           float clip, float lr, float momentum, float dampening, float weight_decay, int nesterov,
                           int first_step, int write_clipped_grad, double* ws, int param_norm_slot, void* stream) {
  FB_REQUIRE(theta && grad && scal && ws && n > 0, "fb_sgd_step: bad arguments");
  FB_REQUIRE(momentum == 0.f || momentum_buf, "fb_sgd_step: momentum needs a buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = flat_grid(n) < kSqBlocks ? flat_grid(n) : kSqBlocks;
  sgd_step_kernel<<<blocks, 256, 0, st>>>(theta, grad, momentum_buf, n, scal, norm_slot, clip, lr, momentum, dampening,
                                          weight_decay, nesterov, first_step, write_clipped_grad, ws);
  sqnorm_final_kernel<<<1, 256, 0, st>>>(ws, blocks, scal, param_norm_slot, nullptr, nullptr);
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Generalised variants for the rest of the hyp.grad_reg surface:
//   acc_strength  : v = bs*g + acc*pre_grads                       (modules.py:217-221, training.py:128-142)
//   central diffs : theta +- 0.5*eps_n*v, vhp = (g+ - g-)/eps_n    (modules.py:266-300)
//   batch_clip    : per-microbatch L2 clip before the running mean (training/utils.py:4-19, training.py:166-168)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqnorm_axpby_partial_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ y, float a, float b,
                                                                   long long n, double* __restrict__ partial) {
  __shared__ double red[8];
  griddep_wait();
  griddep_launch();
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = a * x[i] + (y ? b * y[i] : 0.f);
    acc += v * v;
  }
  double d = warp_sum(double(acc));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) fd_perturb_ex_kernel(const float* __restrict__ theta, const float* __restrict__ g,
                                                            const float* __restrict__ pre, long long n, float bs,
                                                            float acc, float eps, float scale, float* __restrict__ scal,
                                                            int vsq_slot, int eps_slot, float* __restrict__ theta_p) {
  griddep_wait();
  griddep_launch();
  const float eps_n = eps / sqrtf(scal[vsq_slot]);  // scal[vsq_slot] = sum v^2
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[eps_slot] = eps_n;
  const float step = scale * eps_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = bs * g[i] + (pre ? acc * pre[i] : 0.f);
    theta_p[i] = theta[i] + step * v;
  }
}

__global__ void __launch_bounds__(256) fd_combine_ex_kernel(float* __restrict__ g, const float* __restrict__ g_plus,
                                                            const float* __restrict__ g_minus, float* __restrict__ avg,
                                                            long long n, const float* __restrict__ scal, int eps_slot,
                                                            float cf, int cf_slot, const int* __restrict__ cursor,
                                                            int count0, int write_g) {
  griddep_wait();
  griddep_launch();
  const float eps_n = scal[eps_slot];
  if (cf_slot >= 0) cf = scal[cf_slot];
  const int count = count0 + (cursor ? *cursor : 0) + 1;
  const float inv = float(1.0 / double(count));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float h = (g_plus[i] - g_minus[i]) / eps_n;  // modules.py:292-293
    const float gr = g[i] + cf * h;                    // modules.py:299
    if (write_g) g[i] = gr;
    if (avg) {
      const float a = avg[i];
      avg[i] = a + (gr - a) * inv;
    }
  }
}

// avg += (coef*g - avg)/count with coef = clip/(norm+1e-6) if norm > clip (norm = sqrt(scal[norm_slot]));
// scal[clipped_slot] += 1 when the microbatch was clipped; g is scaled in place like the reference does.
__global__ void __launch_bounds__(256) mean_accumulate_clip_kernel(float* __restrict__ g, float* __restrict__ avg,
                                                                   long long n, const int* __restrict__ cursor,
                                                                   int count0, float* __restrict__ scal, int norm_slot,
                                                                   float clip, int clipped_slot) {
  griddep_wait();
  griddep_launch();
  const float norm = sqrtf(scal[norm_slot]);
  const bool clipped = norm > clip;
  const float coef = clipped ? clip / (norm + 1e-6f) : 1.f;
  const int count = count0 + (cursor ? *cursor : 0) + 1;
  const float inv = float(1.0 / double(count));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * coef;
    if (clipped) g[i] = gr;
    const float a = avg[i];
    avg[i] = a + (gr - a) * inv;
  }
  if (clipped && blockIdx.x == 0 && threadIdx.x == 0) scal[clipped_slot] += 1.f;
}

extern "C" int fb_flat_sqnorm_axpby(const float* x, const float* y, float a, float b, int64_t n, double* ws, float* scal,
                                    int slot, void* stream) {
  FB_REQUIRE(x && ws && scal && n > 0, "fb_flat_sqnorm_axpby: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FB_CUDA(launch_pdl(sqnorm_axpby_partial_kernel, dim3(kSqBlocks), dim3(256), 0, st, x, y, a, b, (long long)n, ws));
  FB_CUDA(launch_pdl(sqnorm_final_kernel, dim3(1), dim3(256), 0, st, (const double*)ws, kSqBlocks, scal, slot,
                     (float*)nullptr, (const int*)nullptr));
  return 0;
}

extern "C" int fb_fd_perturb_ex(const float* theta, const float* g, const float* pre, int64_t n, float block_strength,
                                float acc_strength, float eps, float scale, float* scal, int vsq_slot, int eps_slot,
                                float* theta_p, void* stream) {
  FB_REQUIRE(theta && g && scal && theta_p && n > 0, "fb_fd_perturb_ex: bad arguments");
  FB_CUDA(launch_pdl(fd_perturb_ex_kernel, dim3(flat_grid(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), theta, g,
                     pre, (long long)n, block_strength, acc_strength, eps, scale, scal, vsq_slot, eps_slot, theta_p));
  return 0;
}

extern "C" int fb_fd_combine_ex(float* g, const float* g_plus, const float* g_minus, float* avg, int64_t n,
                                const float* scal, int eps_slot, float cf, int cf_slot, const int32_t* cursor,
                                int32_t count0, int write_g, void* stream) {
  FB_REQUIRE(g && g_plus && g_minus && scal && n > 0, "fb_fd_combine_ex: bad arguments");
  FB_CUDA(launch_pdl(fd_combine_ex_kernel, dim3(flat_grid(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), g, g_plus,
                     g_minus, avg, (long long)n, scal, eps_slot, cf, cf_slot, (const int*)cursor, (int)count0, write_g));
  return 0;
}

extern "C" int fb_mean_accumulate_clip(float* g, float* avg, int64_t n, const int32_t* cursor, int32_t count0,
                                       float* scal, int norm_slot, float clip, int clipped_slot, void* stream) {
  FB_REQUIRE(g && avg && scal && n > 0 && clip > 0.f, "fb_mean_accumulate_clip: bad arguments");
  FB_CUDA(launch_pdl(mean_accumulate_clip_kernel, dim3(flat_grid(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), g,
                     avg, (long long)n, (const int*)cursor, (int)count0, scal, norm_slot, clip, clipped_slot));
  return 0;
}
