// BatchNorm1d(eps=1e-3) + ReLU + residual-add (+ squeeze-excitation gate) as memory-bound passes over
// channels-last activations.  Replaces nn.BatchNorm1d / nn.ReLU / the residual add at
// models/QuartNet.py:24,35-37,64,75-77,147-148 and the SE scaling at models/QuartNetContextSE.py:23,55.
//
// Forward (training): the pointwise-GEMM epilogue already produced per-32-row partial sum / sum-of-squares;
//   bn_finalize folds them (fp64) into mean / invstd / folded scale+shift and updates the running stats;
//   bn_apply_act_fwd is ONE pass: out = act(scale1*y + shift1 [*gate] [+ scale2*r + shift2]).
// Backward: bn_act_bwd_reduce (one pass: sum g, sum g*y, sum g*r per utterance chunk), bn_bwd_finalize (per-channel
//   coefficients + dgamma/dbeta), bn_act_bwd_apply (one pass: dy, dr with the MaskCNN gradient mask).
// Every thread moves 8 channels (16 B of bf16 / 32 B of fp32) per row.
#include "common.cuh"

namespace lasr {

template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = bf16x2_to_f32x2(u.x), b = bf16x2_to_f32x2(u.y), c = bf16x2_to_f32x2(u.z), d = bf16x2_to_f32x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = f32x2_to_bf16x2(v[0], v[1]);
    u.y = f32x2_to_bf16x2(v[2], v[3]);
    u.z = f32x2_to_bf16x2(v[4], v[5]);
    u.w = f32x2_to_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) { Vec8<float>::load(p, v); }

// ------------------------------------------------------------------------------------------------
// statistics finalize: grid = C/32 blocks of (32 channels x 8 group slices)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ stats, int groups, int C, double count, float eps, float momentum,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mean_o,
                   float* __restrict__ invstd_o, float* __restrict__ scale_o, float* __restrict__ shift_o,
                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  __shared__ double ss[8][32], sq[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int g = slice; g < groups; g += 8) {
      s += static_cast<double>(stats[(static_cast<size_t>(g) * 2 + 0) * C + c]);
      q += static_cast<double>(stats[(static_cast<size_t>(g) * 2 + 1) * C + c]);
    }
  }
  ss[slice][lane] = s;
  sq[slice][lane] = q;
  __syncthreads();
  if (slice == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      s += ss[i][lane];
      q += sq[i][lane];
    }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + static_cast<double>(eps));
    mean_o[c] = static_cast<float>(mean);
    invstd_o[c] = static_cast<float>(invstd);
    const double g = gamma[c], b = beta[c];
    scale_o[c] = static_cast<float>(g * invstd);
    shift_o[c] = static_cast<float>(b - mean * g * invstd);
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  }
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - rm[c] * s;
  }
}

// ------------------------------------------------------------------------------------------------
// forward apply
// ------------------------------------------------------------------------------------------------
template <typename T, bool HAS_R, bool HAS_GATE>
__global__ void __launch_bounds__(256)
bn_apply_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale1, const float* __restrict__ shift1,
                    const T* __restrict__ r, const float* __restrict__ scale2, const float* __restrict__ shift2,
                    const float* __restrict__ gate, T* __restrict__ out, long long total_vec, int CV, int C, int T_len,
                    int act) {
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total_vec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = v / CV;
    const int c = static_cast<int>(v - row * CV) * 8;
    const size_t off = static_cast<size_t>(row) * C + c;
    float a[8], s1[8], b1[8], o[8];
    Vec8<T>::load(y + off, a);
    load8f(scale1 + c, s1);
    load8f(shift1 + c, b1);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(a[i], s1[i], b1[i]);
    if constexpr (HAS_GATE) {
      const int n = static_cast<int>(row / T_len);
      float g[8];
      load8f(gate + static_cast<size_t>(n) * C + c, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] *= g[i];
    }
    if constexpr (HAS_R) {
      float rr[8], s2[8], b2[8];
      Vec8<T>::load(r + off, rr);
      load8f(scale2 + c, s2);
      load8f(shift2 + c, b2);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += fmaf(rr[i], s2[i], b2[i]);
    }
    if (act == LASR_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
    }
    Vec8<T>::store(out + off, o);
  }
}

// ------------------------------------------------------------------------------------------------
// per-(utterance, channel) sum over time (SE squeeze): sums[n, c] = sum_t y[n, t, c]
// grid (N, chunks); atomics into a zeroed buffer
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
sum_over_time_kernel(const T* __restrict__ y, float* __restrict__ sums, int T_len, int C, int rows_per_chunk) {
  const int n = blockIdx.x;
  const int CV = C / 8;
  const int rows_par = 256 / CV;
  const int tr = threadIdx.x / CV, cv = threadIdx.x - tr * CV;
  if (tr >= rows_par) return;
  const int t0 = blockIdx.y * rows_per_chunk;
  const int t1 = min(T_len, t0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int t = t0 + tr; t < t1; t += rows_par) {
    float a[8];
    Vec8<T>::load(y + (static_cast<size_t>(n) * T_len + t) * C + cv * 8, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += a[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(sums + static_cast<size_t>(n) * C + cv * 8 + i, acc[i]);
}

// ------------------------------------------------------------------------------------------------
// backward reduce: one block per (utterance n, chunk k); partials[(n*chunks + k), s, c], s in 0..2:
//   s=0: sum g, s=1: sum g*y, s=2: sum g*r        with g = dout * (act ? out > 0 : 1)
// ------------------------------------------------------------------------------------------------
template <typename T, bool HAS_R>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ y,
                     const T* __restrict__ r, float* __restrict__ partials, int T_len, int C, int chunks,
                     int rows_per_chunk, int act) {
  extern __shared__ float red[];  // [rows_par][3][C]
  const int n = blockIdx.x / chunks, k = blockIdx.x - n * chunks;
  const int CV = C / 8;
  const int rows_par = 256 / CV;
  const int tr = threadIdx.x / CV, cv = threadIdx.x - tr * CV;
  const int t0 = k * rows_per_chunk;
  const int t1 = min(T_len, t0 + rows_per_chunk);
  float sg[8], sgy[8], sgr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sg[i] = sgy[i] = sgr[i] = 0.f;
  if (tr < rows_par) {
    for (int t = t0 + tr; t < t1; t += rows_par) {
      const size_t off = (static_cast<size_t>(n) * T_len + t) * C + cv * 8;
      float g[8], o[8], yy[8];
      Vec8<T>::load(dout + off, g);
      Vec8<T>::load(y + off, yy);
      if (act == LASR_ACT_RELU) {
        Vec8<T>::load(out + off, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sg[i] += g[i];
        sgy[i] = fmaf(g[i], yy[i], sgy[i]);
      }
      if constexpr (HAS_R) {
        float rr[8];
        Vec8<T>::load(r + off, rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) sgr[i] = fmaf(g[i], rr[i], sgr[i]);
      }
    }
    float* dst = red + static_cast<size_t>(tr) * 3 * C + cv * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dst[i] = sg[i];
      dst[C + i] = sgy[i];
      dst[2 * C + i] = sgr[i];
    }
  }
  __syncthreads();
  float* pdst = partials + static_cast<size_t>(blockIdx.x) * 3 * C;
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    float s = 0.f;
    for (int j = 0; j < rows_par; ++j) s += red[static_cast<size_t>(j) * 3 * C + i];
    pdst[i] = s;
  }
}

// per-channel finalize: sums partial slots idx_g / idx_gx over all groups
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int groups, int nslots, int C, double count, int idx_g,
                       int idx_gx, const float* __restrict__ gamma, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ coef) {
  __shared__ double s0[8][32], s1[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double a = 0.0, b = 0.0;
  if (c < C) {
    for (int g = slice; g < groups; g += 8) {
      a += static_cast<double>(partials[(static_cast<size_t>(g) * nslots + idx_g) * C + c]);
      b += static_cast<double>(partials[(static_cast<size_t>(g) * nslots + idx_gx) * C + c]);
    }
  }
  s0[slice][lane] = a;
  s1[slice][lane] = b;
  __syncthreads();
  if (slice == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      a += s0[i][lane];
      b += s1[i][lane];
    }
    const double mu = mean[c], is = invstd[c], ga = gamma[c];
    const double dga = is * (b - mu * a);
    if (dgamma != nullptr) dgamma[c] += static_cast<float>(dga);
    if (dbeta != nullptr) dbeta[c] += static_cast<float>(a);
    const double c0 = ga * is;
    const double c1 = -ga * is * is * dga / count;
    const double c2 = -ga * is * a / count - c1 * mu;
    coef[c] = static_cast<float>(c0);
    coef[C + c] = static_cast<float>(c1);
    coef[2 * C + c] = static_cast<float>(c2);
  }
}

// backward apply: dy = mask(coef1[0]*(g*gate + extra) + coef1[1]*y + coef1[2]); dr = coef2[0]*g + coef2[1]*r + coef2[2]
template <typename T, bool HAS_R, bool HAS_GATE>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ y,
                    const T* __restrict__ r, const float* __restrict__ gate, const float* __restrict__ extra,
                    const float* __restrict__ coef1, const float* __restrict__ coef2,
                    const int32_t* __restrict__ lengths, int T_len, T* __restrict__ dy, T* __restrict__ dr,
                    long long total_vec, int CV, int C, int act) {
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total_vec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = v / CV;
    const int c = static_cast<int>(v - row * CV) * 8;
    const size_t off = static_cast<size_t>(row) * C + c;
    const int n = static_cast<int>(row / T_len);
    const int t = static_cast<int>(row - static_cast<long long>(n) * T_len);
    float g[8], yy[8];
    Vec8<T>::load(dout + off, g);
    if (act == LASR_ACT_RELU) {
      float o[8];
      Vec8<T>::load(out + off, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
    }
    if constexpr (HAS_R) {
      float rr[8], a0[8], a1[8], a2[8], d[8];
      Vec8<T>::load(r + off, rr);
      load8f(coef2 + c, a0);
      load8f(coef2 + C + c, a1);
      load8f(coef2 + 2 * C + c, a2);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = fmaf(a0[i], g[i], fmaf(a1[i], rr[i], a2[i]));
      Vec8<T>::store(dr + off, d);
    }
    const bool keep = lengths == nullptr || t < lengths[n];
    float d[8];
    if (keep) {
      Vec8<T>::load(y + off, yy);
      if constexpr (HAS_GATE) {
        float gt[8], ex[8];
        load8f(gate + static_cast<size_t>(n) * C + c, gt);
        load8f(extra + static_cast<size_t>(n) * C + c, ex);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(g[i], gt[i], ex[i]);
      }
      float a0[8], a1[8], a2[8];
      load8f(coef1 + c, a0);
      load8f(coef1 + C + c, a1);
      load8f(coef1 + 2 * C + c, a2);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = fmaf(a0[i], g[i], fmaf(a1[i], yy[i], a2[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = 0.f;
    }
    Vec8<T>::store(dy + off, d);
  }
}

static inline int ew_grid(long long total_vec) {
  long long b = (total_vec + 255) / 256;
  const long long cap = 8LL * kNumSMs;
  return static_cast<int>(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_bn_finalize(const float* stats, int groups, int C, int count, float eps, float momentum, const float* gamma,
                     const float* beta, float* mean, float* invstd, float* scale, float* shift, float* running_mean,
                     float* running_var, lasr_stream_t stream) {
  if (groups <= 0 || C <= 0 || count <= 0) return LASR_ERR_BAD_SHAPE;
  bn_finalize_kernel<<<cdiv(C, 32), 256, 0, stream>>>(stats, groups, C, static_cast<double>(count), eps, momentum,
                                                      gamma, beta, mean, invstd, scale, shift, running_mean,
                                                      running_var);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, float* scale, float* shift, int C, lasr_stream_t stream) {
  if (C <= 0) return LASR_ERR_BAD_SHAPE;
  bn_eval_coeffs_kernel<<<cdiv(C, 256), 256, 0, stream>>>(gamma, beta, running_mean, running_var, eps, scale, shift,
                                                          C);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_sum_over_time(const void* y, float* sums, int N, int T, int C, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || (C % 8) || C > 2048) return LASR_ERR_BAD_SHAPE;
  cudaError_t e = cudaMemsetAsync(sums, 0, static_cast<size_t>(N) * C * sizeof(float), stream);
  if (e != cudaSuccess) {
    lasr_set_cuda_error(e);
    return LASR_ERR_CUDA;
  }
  int chunks = cdiv(4 * kNumSMs, N);
  if (chunks > cdiv(T, 16)) chunks = cdiv(T, 16);
  const int rows_per_chunk = cdiv(T, chunks);
  chunks = cdiv(T, rows_per_chunk);
  dim3 grid(N, chunks);
  if (dtype == LASR_F32)
    sum_over_time_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(y), sums, T, C, rows_per_chunk);
  else if (dtype == LASR_BF16)
    sum_over_time_kernel<__nv_bfloat16>
        <<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), sums, T, C, rows_per_chunk);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

#define LASR_BN_FWD_LAUNCH(TT)                                                                                       \
  do {                                                                                                               \
    const TT* yy = static_cast<const TT*>(y);                                                                        \
    const TT* rr = static_cast<const TT*>(r);                                                                        \
    TT* oo = static_cast<TT*>(out);                                                                                  \
    if (r != nullptr && gate != nullptr)                                                                             \
      bn_apply_fwd_kernel<TT, true, true>                                                                            \
          <<<grid, 256, 0, stream>>>(yy, scale1, shift1, rr, scale2, shift2, gate, oo, total, CV, C, T, act);        \
    else if (r != nullptr)                                                                                           \
      bn_apply_fwd_kernel<TT, true, false>                                                                           \
          <<<grid, 256, 0, stream>>>(yy, scale1, shift1, rr, scale2, shift2, gate, oo, total, CV, C, T, act);        \
    else if (gate != nullptr)                                                                                        \
      bn_apply_fwd_kernel<TT, false, true>                                                                           \
          <<<grid, 256, 0, stream>>>(yy, scale1, shift1, rr, scale2, shift2, gate, oo, total, CV, C, T, act);        \
    else                                                                                                             \
      bn_apply_fwd_kernel<TT, false, false>                                                                          \
          <<<grid, 256, 0, stream>>>(yy, scale1, shift1, rr, scale2, shift2, gate, oo, total, CV, C, T, act);        \
  } while (0)

int lasr_bn_apply_act_fwd(const void* y, const float* scale1, const float* shift1, const void* r, const float* scale2,
                          const float* shift2, const float* gate, void* out, int M, int C, int T, int act, int dtype,
                          lasr_stream_t stream) {
  if (M <= 0 || C <= 0 || (C % 8)) return LASR_ERR_BAD_SHAPE;
  if (gate != nullptr && T <= 0) return LASR_ERR_BAD_SHAPE;
  const int CV = C / 8;
  const long long total = static_cast<long long>(M) * CV;
  const int grid = ew_grid(total);
  if (dtype == LASR_F32)
    LASR_BN_FWD_LAUNCH(float);
  else if (dtype == LASR_BF16)
    LASR_BN_FWD_LAUNCH(__nv_bfloat16);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_bn_bwd_chunks(int N, int T) {
  int chunks = cdiv(4 * kNumSMs, N);
  const int max_chunks = cdiv(T, 8);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows = cdiv(T, chunks);
  return cdiv(T, rows);
}

int lasr_bn_act_bwd_reduce(const void* dout, const void* out, const void* y, const void* r, float* partials, int N,
                           int T, int C, int chunks, int act, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || (C % 8) || C > 2048 || chunks <= 0) return LASR_ERR_BAD_SHAPE;
  const int rows_per_chunk = cdiv(T, chunks);
  const int CV = C / 8;
  const int rows_par = 256 / CV;
  const int smem = rows_par * 3 * C * static_cast<int>(sizeof(float));
  const int grid = N * chunks;
#define LASR_BN_RED_LAUNCH(TT)                                                                                     \
  do {                                                                                                             \
    if (r != nullptr)                                                                                              \
      bn_bwd_reduce_kernel<TT, true><<<grid, 256, smem, stream>>>(                                                 \
          static_cast<const TT*>(dout), static_cast<const TT*>(out), static_cast<const TT*>(y),                    \
          static_cast<const TT*>(r), partials, T, C, chunks, rows_per_chunk, act);                                 \
    else                                                                                                           \
      bn_bwd_reduce_kernel<TT, false><<<grid, 256, smem, stream>>>(                                                \
          static_cast<const TT*>(dout), static_cast<const TT*>(out), static_cast<const TT*>(y),                    \
          static_cast<const TT*>(r), partials, T, C, chunks, rows_per_chunk, act);                                 \
  } while (0)
  if (dtype == LASR_F32)
    LASR_BN_RED_LAUNCH(float);
  else if (dtype == LASR_BF16)
    LASR_BN_RED_LAUNCH(__nv_bfloat16);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_bn_bwd_finalize(const float* partials, int groups, int nslots, int C, int count, int idx_g, int idx_gx,
                         const float* gamma, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                         float* coef, lasr_stream_t stream) {
  if (groups <= 0 || C <= 0 || count <= 0 || idx_g >= nslots || idx_gx >= nslots) return LASR_ERR_BAD_SHAPE;
  bn_bwd_finalize_kernel<<<cdiv(C, 32), 256, 0, stream>>>(partials, groups, nslots, C, static_cast<double>(count),
                                                          idx_g, idx_gx, gamma, mean, invstd, dgamma, dbeta, coef);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_bn_act_bwd_apply(const void* dout, const void* out, const void* y, const void* r, const float* gate,
                          const float* extra, const float* coef1, const float* coef2, const int32_t* lengths, int T,
                          void* dy, void* dr, int M, int C, int act, int dtype, lasr_stream_t stream) {
  if (M <= 0 || C <= 0 || (C % 8) || T <= 0) return LASR_ERR_BAD_SHAPE;
  if ((r != nullptr) != (dr != nullptr) || (r != nullptr) != (coef2 != nullptr)) return LASR_ERR_BAD_SHAPE;
  if ((gate != nullptr) != (extra != nullptr)) return LASR_ERR_BAD_SHAPE;
  const int CV = C / 8;
  const long long total = static_cast<long long>(M) * CV;
  const int grid = ew_grid(total);
#define LASR_BN_BAPPLY_LAUNCH(TT)                                                                                   \
  do {                                                                                                              \
    const TT* a0 = static_cast<const TT*>(dout);                                                                    \
    const TT* a1 = static_cast<const TT*>(out);                                                                     \
    const TT* a2 = static_cast<const TT*>(y);                                                                       \
    const TT* a3 = static_cast<const TT*>(r);                                                                       \
    TT* o0 = static_cast<TT*>(dy);                                                                                  \
    TT* o1 = static_cast<TT*>(dr);                                                                                  \
    if (r != nullptr && gate != nullptr)                                                                            \
      bn_bwd_apply_kernel<TT, true, true><<<grid, 256, 0, stream>>>(a0, a1, a2, a3, gate, extra, coef1, coef2,      \
                                                                    lengths, T, o0, o1, total, CV, C, act);         \
    else if (r != nullptr)                                                                                          \
      bn_bwd_apply_kernel<TT, true, false><<<grid, 256, 0, stream>>>(a0, a1, a2, a3, gate, extra, coef1, coef2,     \
                                                                     lengths, T, o0, o1, total, CV, C, act);        \
    else if (gate != nullptr)                                                                                       \
      bn_bwd_apply_kernel<TT, false, true><<<grid, 256, 0, stream>>>(a0, a1, a2, a3, gate, extra, coef1, coef2,     \
                                                                     lengths, T, o0, o1, total, CV, C, act);        \
    else                                                                                                            \
      bn_bwd_apply_kernel<TT, false, false><<<grid, 256, 0, stream>>>(a0, a1, a2, a3, gate, extra, coef1, coef2,    \
                                                                      lengths, T, o0, o1, total, CV, C, act);       \
  } while (0)
  if (dtype == LASR_F32)
    LASR_BN_BAPPLY_LAUNCH(float);
  else if (dtype == LASR_BF16)
    LASR_BN_BAPPLY_LAUNCH(__nv_bfloat16);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // extern "C"
