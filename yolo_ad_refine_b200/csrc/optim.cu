// Parameter plumbing and the optimizer of the training path:
//   * table-driven permutes between the reference's parameter layouts (the fp32 master copy, torch state-dict layout) and the kernel
//     layouts (conv weights [cout][tap][cin] in the activation dtype, their dgrad twins, padded biases, depthwise [tap][c]), and the reverse
//     accumulation of packed weight gradients into the torch-layout gradient arena;
//   * global gradient norm, fused clip + SGD(nesterov) + weight decay over the flat arenas (engine/trainer.py:580-588, :784-808),
//     ModelEMA update (utils/torch_utils.py:530-541).
#include "common.cuh"

namespace {

template <typename T>
__global__ void permute_pack_kernel(const yad_permute_entry* __restrict__ tab, const float* __restrict__ src, T* __restrict__ dst_t,
                                    float* __restrict__ dst_f) {
  pdl_sync();
  const yad_permute_entry e = tab[blockIdx.y];
  const int64_t total = e.p0 * e.n1 * e.p2;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = it % e.p2, t = (it / e.p2) % e.n1, i = it / (e.p2 * e.n1);
    const int64_t ts = e.flip ? e.n1 - 1 - t : t;
    const float v = (i < e.n0 && j < e.n2) ? src[e.src_off + i * e.s0 + ts * e.s1 + j * e.s2] : 0.f;
    if (e.dst_f32)
      dst_f[e.dst_off + it] = v;
    else
      st1(dst_t + e.dst_off + it, v);
  }
}

// grads_torch[src index] += packed_f32[dst index]
__global__ void permute_unpack_kernel(const yad_permute_entry* __restrict__ tab, const float* __restrict__ packed, float* __restrict__ grads) {
  pdl_sync();
  const yad_permute_entry e = tab[blockIdx.y];
  const int64_t total = e.n0 * e.n1 * e.n2;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = it % e.n2, t = (it / e.n2) % e.n1, i = it / (e.n2 * e.n1);
    const int64_t ts = e.flip ? e.n1 - 1 - t : t;
    grads[e.src_off + i * e.s0 + ts * e.s1 + j * e.s2] += packed[e.dst_off + (i * e.n1 + t) * e.p2 + j];
  }
}

__global__ void sqnorm_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
  pdl_sync();
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fmaf(x[i], x[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, (double)s);
}

// torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.SGD(nesterov=True) with per-group lr / weight decay
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, const uint8_t* __restrict__ group,
                                int64_t n, float lr0, float lr1, float lr2, float wd0, float wd1, float wd2, float momentum, float max_norm,
                                const double* __restrict__ norm_sq, int first) {
  pdl_sync();
  float coef = 1.0f;
  if (max_norm > 0.f && norm_sq) {
    const float nrm = (float)sqrt(*norm_sq);
    coef = fminf(max_norm / (nrm + 1e-6f), 1.0f);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int gr = group[i];
    if (gr > 2) continue;  // frozen
    const float lr = gr == 0 ? lr0 : (gr == 1 ? lr1 : lr2), wd = gr == 0 ? wd0 : (gr == 1 ? wd1 : wd2);
    const float w = p[i];
    float d = g[i] * coef + wd * w;
    const float b = first ? d : momentum * mom[i] + d;
    mom[i] = b;
    d = d + momentum * b;
    p[i] = w - lr * d;
  }
}

// torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.AdamW (decoupled weight decay) with per-group lr / weight decay:
// p *= 1 - lr wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adamw_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                  const uint8_t* __restrict__ group, int64_t n, float lr0, float lr1, float lr2, float wd0, float wd1, float wd2,
                                  float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float max_norm, const double* __restrict__ norm_sq) {
  pdl_sync();
  float coef = 1.0f;
  if (max_norm > 0.f && norm_sq) {
    const float nrm = (float)sqrt(*norm_sq);
    coef = fminf(max_norm / (nrm + 1e-6f), 1.0f);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int gr = group[i];
    if (gr > 2) continue;  // frozen
    const float lr = gr == 0 ? lr0 : (gr == 1 ? lr1 : lr2), wd = gr == 0 ? wd0 : (gr == 1 ? wd1 : wd2);
    const float gi = g[i] * coef;
    float w = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = w;
  }
}

__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ p, int64_t n, float d) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    ema[i] = d * ema[i] + (1.0f - d) * p[i];
}

// ---- the same three kernels with every per-step scalar read from DEVICE memory, so that one captured CUDA graph serves all steps of a run
// (learning-rate warm-up, Adam bias correction and the EMA decay ramp change from step to step).  hyper[13] = lr0 lr1 lr2 | wd0 wd1 wd2 |
// momentum or beta1 | beta2 | eps | 1 - beta1^t | sqrt(1 - beta2^t) | max_norm | ema decay
__global__ void sgd_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, const uint8_t* __restrict__ group,
                                    int64_t n, const float* __restrict__ hyper, const double* __restrict__ norm_sq) {
  pdl_sync();
  const float lr0 = hyper[0], lr1 = hyper[1], lr2 = hyper[2], wd0 = hyper[3], wd1 = hyper[4], wd2 = hyper[5], momentum = hyper[6], max_norm = hyper[11];
  float coef = 1.0f;
  if (max_norm > 0.f && norm_sq) {
    const float nrm = (float)sqrt(*norm_sq);
    coef = fminf(max_norm / (nrm + 1e-6f), 1.0f);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int gr = group[i];
    if (gr > 2) continue;  // frozen
    const float lr = gr == 0 ? lr0 : (gr == 1 ? lr1 : lr2), wd = gr == 0 ? wd0 : (gr == 1 ? wd1 : wd2);
    const float w = p[i];
    float d = g[i] * coef + wd * w;
    const float b = momentum * mom[i] + d;  // the momentum arena starts at zero: the first step's buffer is d, as torch.optim.SGD initialises it
    mom[i] = b;
    d = d + momentum * b;
    p[i] = w - lr * d;
  }
}
__global__ void adamw_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                      const uint8_t* __restrict__ group, int64_t n, const float* __restrict__ hyper, const double* __restrict__ norm_sq) {
  pdl_sync();
  const float lr0 = hyper[0], lr1 = hyper[1], lr2 = hyper[2], wd0 = hyper[3], wd1 = hyper[4], wd2 = hyper[5], beta1 = hyper[6], beta2 = hyper[7],
              eps = hyper[8], bc1 = hyper[9], bc2_sqrt = hyper[10], max_norm = hyper[11];
  float coef = 1.0f;
  if (max_norm > 0.f && norm_sq) {
    const float nrm = (float)sqrt(*norm_sq);
    coef = fminf(max_norm / (nrm + 1e-6f), 1.0f);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int gr = group[i];
    if (gr > 2) continue;
    const float lr = gr == 0 ? lr0 : (gr == 1 ? lr1 : lr2), wd = gr == 0 ? wd0 : (gr == 1 ? wd1 : wd2);
    const float gi = g[i] * coef;
    float w = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi, vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = w - (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}
__global__ void ema_dev_kernel(float* __restrict__ ema, const float* __restrict__ p, int64_t n, const float* __restrict__ d_dev) {
  pdl_sync();
  const float d = *d_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    ema[i] = d * ema[i] + (1.0f - d) * p[i];
}

int blocks_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

extern "C" {

int yad_permute_pack(const yad_permute_entry* table_dev, int n_entries, int64_t max_elems, const float* src, void* dst_t, float* dst_f32, int dtype,
                     void* stream) {
  if (n_entries <= 0) return 0;
  int gx = (int)((max_elems + 255) / 256);
  gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
  dim3 grid(gx, n_entries);
  YAD_DISPATCH_DTYPE(dtype, YAD_LAUNCH(permute_pack_kernel<T>, grid, 256, 0, (cudaStream_t)stream, table_dev, src, (T*)dst_t, dst_f32);)
  YAD_LAUNCH_CHECK("permute_pack");
  return 0;
}

int yad_permute_unpack(const yad_permute_entry* table_dev, int n_entries, int64_t max_elems, const float* packed, float* grads, void* stream) {
  if (n_entries <= 0) return 0;
  int gx = (int)((max_elems + 255) / 256);
  gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
  dim3 grid(gx, n_entries);
  YAD_LAUNCH(permute_unpack_kernel, grid, 256, 0, (cudaStream_t)stream, table_dev, packed, grads);
  YAD_LAUNCH_CHECK("permute_unpack");
  return 0;
}

int yad_sqnorm(const float* x, int64_t n, double* out, void* stream) {
  YAD_LAUNCH(sqnorm_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, x, n, out);
  YAD_LAUNCH_CHECK("sqnorm");
  return 0;
}

int yad_sgd_step(float* params, const float* grads, float* momentum_buf, const uint8_t* group, int64_t n, const float* lr3_host,
                 const float* wd3_host, float momentum, float max_norm, const double* norm_sq, int first_step, void* stream) {
  YAD_LAUNCH(sgd_step_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, params, grads, momentum_buf, group, n, lr3_host[0], lr3_host[1], lr3_host[2],
                                                                    wd3_host[0], wd3_host[1], wd3_host[2], momentum, max_norm, norm_sq, first_step);
  YAD_LAUNCH_CHECK("sgd_step");
  return 0;
}

int yad_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* group, int64_t n, const float* lr3_host,
                   const float* wd3_host, float beta1, float beta2, float eps, int step, float max_norm, const double* norm_sq, void* stream) {
  YAD_CHECK(step >= 1, "adamw_step: step counts from 1");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  YAD_LAUNCH(adamw_step_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, group, n, lr3_host[0], lr3_host[1], lr3_host[2],
                                                                      wd3_host[0], wd3_host[1], wd3_host[2], beta1, beta2, eps, bc1, bc2_sqrt, max_norm,
                                                                      norm_sq);
  YAD_LAUNCH_CHECK("adamw_step");
  return 0;
}

int yad_sgd_step_dev(float* params, const float* grads, float* momentum_buf, const uint8_t* group, int64_t n, const float* hyper_dev, const double* norm_sq,
                     void* stream) {
  YAD_LAUNCH(sgd_step_dev_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, params, grads, momentum_buf, group, n, hyper_dev, norm_sq);
  YAD_LAUNCH_CHECK("sgd_step_dev");
  return 0;
}

int yad_adamw_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* group, int64_t n, const float* hyper_dev,
                       const double* norm_sq, void* stream) {
  YAD_LAUNCH(adamw_step_dev_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, group, n, hyper_dev, norm_sq);
  YAD_LAUNCH_CHECK("adamw_step_dev");
  return 0;
}

int yad_ema_update_dev(float* ema, const float* params, int64_t n, const float* decay_dev, void* stream) {
  YAD_LAUNCH(ema_dev_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, ema, params, n, decay_dev);
  YAD_LAUNCH_CHECK("ema_update_dev");
  return 0;
}

int yad_ema_update(float* ema, const float* params, int64_t n, float decay, void* stream) {
  YAD_LAUNCH(ema_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, ema, params, n, decay);
  YAD_LAUNCH_CHECK("ema_update");
  return 0;
}

}  // extern "C"
