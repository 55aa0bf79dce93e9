// Novograd + cosine-annealing-with-warm-restarts learning-rate schedule, fused over the step runtime's FLAT parameter /
// gradient buffers (runtime.ParamBank).  Replaces scheduler/novograd.py:75-145 (Optimizer.step: per parameter tensor a
// grad.norm(), a host-syncing `if exp_avg_sq == 0`, and ~6 elementwise launches -- 100-145 tensors per step) and
// scheduler/cosine_annearing_with_warmup.py:53-89 (host-side schedule) with THREE launches per step and no host
// round trip, so the optimizer lives inside the step's CUDA graph:
//
//   novograd_norms    one CTA per chunk (<= 4096 consecutive elements of ONE parameter tensor): sum of squares of the
//                     gradient, fp64 RED into norms[param]                                   (novograd.py:113)
//   novograd_moments  one thread per parameter tensor: v = (v == 0) ? |g|^2 : b2*v + (1-b2)*|g|^2, denom = sqrt(v)+eps
//                     (:115-126); thread 0 also publishes the learning rate of THIS step and advances the schedule
//                     state exactly like CosineAnnealingWarmupRestarts.step(epoch=None) (:66-72, 84-89)
//   novograd_update   one CTA per chunk: g' = g/denom + wd*p (:128-130) [* (1-b1) with grad_averaging :131-132],
//                     m = b1*m + g' (:133), p -= lr*m (:143); the bf16 shadow of p used by the tensor-core kernels is
//                     refreshed in the same pass (the step then needs no separate cast launch)
//
// amsgrad / luc (layer-wise update clipping) are not used by the reference's training script (train.py:46) and are
// rejected by the Python wrapper.
#include "common.cuh"

namespace lasr {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
novograd_norms_kernel(const float* __restrict__ grads, const int32_t* __restrict__ chunk_off,
                      const int32_t* __restrict__ chunk_len, const int32_t* __restrict__ chunk_param,
                      double* __restrict__ norms) {
  const int ch = blockIdx.x;
  const float* g = grads + chunk_off[ch];
  const int len = chunk_len[ch];
  float acc = 0.f;
  // chunk offsets are multiples of 4 elements (bank alignment 64), tails handled scalar
  const int n4 = len >> 2;
  for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc = fmaf(v.x, v.x, acc);
    acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc);
    acc = fmaf(v.w, v.w, acc);
  }
  for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) acc = fmaf(g[i], g[i], acc);
  __shared__ double part[OPT_THREADS / 32];
  double d = static_cast<double>(warp_sum(acc));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += part[w];
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(norms + chunk_param[ch]), "d"(s) : "memory");
  }
}

__device__ __forceinline__ double sched_lr(const lasr_lr_sched_t& s) {
  // get_lr(): scheduler/cosine_annearing_with_warmup.py:53-62 (base_lr == min_lr after init_lr, :47-51)
  if (s.step_in_cycle == -1) return s.min_lr;
  if (s.step_in_cycle < s.warmup_steps)
    return (s.max_lr - s.min_lr) * static_cast<double>(s.step_in_cycle) / static_cast<double>(s.warmup_steps) + s.min_lr;
  const double kPi = 3.141592653589793;
  return s.min_lr + (s.max_lr - s.min_lr) *
                        (1.0 + cos(kPi * static_cast<double>(s.step_in_cycle - s.warmup_steps) /
                                   static_cast<double>(s.cur_cycle_steps - s.warmup_steps))) / 2.0;
}

__global__ void novograd_moments_kernel(const double* __restrict__ norms, float* __restrict__ exp_avg_sq,
                                        float* __restrict__ denom, int num_params, float beta2, float eps,
                                        lasr_lr_sched_t* __restrict__ sched, float* __restrict__ lr_use,
                                        float fixed_lr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < num_params) {
    const float n2 = static_cast<float>(norms[p]);
    const float v0 = exp_avg_sq[p];
    const float v = (v0 == 0.f) ? n2 : v0 * beta2 + (1.0f - beta2) * n2;  // :115-118 (mul_ then add_ with alpha)
    exp_avg_sq[p] = v;
    denom[p] = sqrtf(v) + eps;  // :126
  }
  if (p == 0) {
    if (sched == nullptr) {
      *lr_use = fixed_lr;
    } else {
      lasr_lr_sched_t s = *sched;
      *lr_use = static_cast<float>(s.lr);  // the optimizer step uses the lr set by the PREVIOUS scheduler.step()
      // scheduler.step(epoch=None): :66-72
      s.last_epoch += 1;
      s.step_in_cycle += 1;
      if (s.step_in_cycle >= s.cur_cycle_steps) {
        s.cycle += 1;
        s.step_in_cycle -= s.cur_cycle_steps;
        s.cur_cycle_steps =
            static_cast<int>(static_cast<double>(s.cur_cycle_steps - s.warmup_steps) * s.cycle_mult) + s.warmup_steps;
      }
      s.max_lr = s.base_max_lr * pow(s.gamma, static_cast<double>(s.cycle));  // :86
      s.lr = sched_lr(s);
      *sched = s;
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
novograd_update_kernel(float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ exp_avg,
                       __nv_bfloat16* __restrict__ shadow, const int32_t* __restrict__ chunk_off,
                       const int32_t* __restrict__ chunk_len, const int32_t* __restrict__ chunk_param,
                       const float* __restrict__ denom, const float* __restrict__ lr_use, float beta1,
                       float weight_decay, int grad_averaging) {
  const int ch = blockIdx.x;
  const int off = chunk_off[ch];
  const int len = chunk_len[ch];
  const float den = denom[chunk_param[ch]];
  const float lr = *lr_use;
  const float ga = grad_averaging ? (1.0f - beta1) : 1.0f;
  auto upd = [&](float p, float g, float m, float& p_out, float& m_out) {
    float gg = g / den;                                  // grad.div_(denom)              :128
    if (weight_decay != 0.f) gg = gg + weight_decay * p;  // grad.add_(p, alpha=wd)        :129-130
    if (grad_averaging) gg *= ga;                         //                               :131-132
    m_out = m * beta1 + gg;                               // exp_avg.mul_(beta1).add_(grad) :133
    p_out = p - lr * m_out;                               // p.add_(exp_avg, alpha=-lr)    :143
  };
  const int n4 = len >> 2;
  for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
    float4 p = reinterpret_cast<float4*>(params + off)[i];
    const float4 g = reinterpret_cast<const float4*>(grads + off)[i];
    float4 m = reinterpret_cast<float4*>(exp_avg + off)[i];
    upd(p.x, g.x, m.x, p.x, m.x);
    upd(p.y, g.y, m.y, p.y, m.y);
    upd(p.z, g.z, m.z, p.z, m.z);
    upd(p.w, g.w, m.w, p.w, m.w);
    reinterpret_cast<float4*>(params + off)[i] = p;
    reinterpret_cast<float4*>(exp_avg + off)[i] = m;
    if (shadow != nullptr) {
      uint2 u;
      u.x = f32x2_to_bf16x2(p.x, p.y);
      u.y = f32x2_to_bf16x2(p.z, p.w);
      reinterpret_cast<uint2*>(shadow + off)[i] = u;
    }
  }
  for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) {
    float p, m;
    upd(params[off + i], grads[off + i], exp_avg[off + i], p, m);
    params[off + i] = p;
    exp_avg[off + i] = m;
    if (shadow != nullptr) shadow[off + i] = __float2bfloat16_rn(p);
  }
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_novograd_step(float* params, const float* grads, float* exp_avg, void* shadow_bf16, const int32_t* chunk_off,
                       const int32_t* chunk_len, const int32_t* chunk_param, int num_chunks, double* norms,
                       float* exp_avg_sq, float* denom, int num_params, lasr_lr_sched_t* sched, float* lr_use,
                       float lr, float beta1, float beta2, float eps, float weight_decay, int grad_averaging,
                       lasr_stream_t stream) {
  if (num_chunks <= 0 || num_params <= 0 || params == nullptr || grads == nullptr || exp_avg == nullptr ||
      norms == nullptr || exp_avg_sq == nullptr || denom == nullptr || lr_use == nullptr)
    return LASR_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg)) & 15)
    return LASR_ERR_ALIGNMENT;
  novograd_norms_kernel<<<num_chunks, OPT_THREADS, 0, stream>>>(grads, chunk_off, chunk_len, chunk_param, norms);
  LASR_CHECK_LAUNCH();
  novograd_moments_kernel<<<cdiv(num_params, 128), 128, 0, stream>>>(norms, exp_avg_sq, denom, num_params, beta2, eps,
                                                                     sched, lr_use, lr);
  LASR_CHECK_LAUNCH();
  novograd_update_kernel<<<num_chunks, OPT_THREADS, 0, stream>>>(
      params, grads, exp_avg, static_cast<__nv_bfloat16*>(shadow_bf16), chunk_off, chunk_len, chunk_param, denom, lr_use,
      beta1, weight_decay, grad_averaging);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // extern "C"
