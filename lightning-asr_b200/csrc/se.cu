// Squeeze-excitation (models/QuartNetContextSE.py:8-23, applied at :55 between BatchNorm and ReLU).
// The squeeze mean over time of the BatchNorm output is an affine function of the per-(utterance, channel) time
// sum of the PRE-BN tensor: mean_t(BN(y)) = scale*(sum_t y)/T + shift, so no extra pass over the BN output exists.
// Everything here works on [N, C] tensors: one CTA per utterance.
#include "common.cuh"

namespace lasr {

// s = scale*Sy/T + shift ; h = relu(W1 s) ; gate = sigmoid(W2 h)
__global__ void __launch_bounds__(256)
se_fwd_kernel(const float* __restrict__ sums, const float* __restrict__ scale, const float* __restrict__ shift,
              float inv_T, const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ s_out,
              float* __restrict__ hidden, float* __restrict__ gate, int C, int Cr) {
  extern __shared__ float sm[];  // s[C], h[Cr]
  float* s = sm;
  float* h = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = fmaf(scale[c], sums[static_cast<size_t>(n) * C + c] * inv_T, shift[c]);
    s[c] = v;
    s_out[static_cast<size_t>(n) * C + c] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cr; j += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(w1[static_cast<size_t>(j) * C + c], s[c], a);
    a = warp_sum(a);
    if (lane == 0) {
      a = fmaxf(a, 0.f);
      h[j] = a;
      hidden[static_cast<size_t>(n) * Cr + j] = a;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < Cr; ++j) a = fmaf(w2[static_cast<size_t>(c) * Cr + j], h[j], a);
    gate[static_cast<size_t>(n) * C + c] = 1.f / (1.f + expf(-a));
  }
}

// dgate[c] = scale*Sgy + shift*Sg (time sums of g*y and g for this utterance, folded from the chunk partials)
// -> ds ; extra = ds / T ; dW1 += , dW2 +=
__global__ void __launch_bounds__(256)
se_bwd_kernel(const float* __restrict__ partials, int chunks, const float* __restrict__ scale,
              const float* __restrict__ shift, float inv_T, const float* __restrict__ w1, const float* __restrict__ w2,
              const float* __restrict__ s_in, const float* __restrict__ hidden, const float* __restrict__ gate,
              float* __restrict__ extra, float* __restrict__ dw1, float* __restrict__ dw2, int C, int Cr) {
  extern __shared__ float sm[];  // dz[C], s[C], h[Cr], dh[Cr]
  float* dz = sm;
  float* s = sm + C;
  float* h = sm + 2 * C;
  float* dh = h + Cr;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sg = 0.f, sgy = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float* p = partials + static_cast<size_t>(n * chunks + k) * 3 * C;
      sg += p[c];
      sgy += p[C + c];
    }
    const float dgate = scale[c] * sgy + shift[c] * sg;
    const float g = gate[static_cast<size_t>(n) * C + c];
    dz[c] = dgate * g * (1.f - g);
    s[c] = s_in[static_cast<size_t>(n) * C + c];
  }
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) h[j] = hidden[static_cast<size_t>(n) * Cr + j];
  __syncthreads();
  // dW2[c, j] += dz[c] * h[j]
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    const int c = i / Cr, j = i - c * Cr;
    atomicAdd(dw2 + i, dz[c] * h[j]);
  }
  // dh[j] = (h[j] > 0) * sum_c W2[c, j] dz[c]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cr; j += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(w2[static_cast<size_t>(c) * Cr + j], dz[c], a);
    a = warp_sum(a);
    if (lane == 0) dh[j] = h[j] > 0.f ? a : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    const int j = i / C, c = i - j * C;
    atomicAdd(dw1 + i, dh[j] * s[c]);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < Cr; ++j) a = fmaf(w1[static_cast<size_t>(j) * C + c], dh[j], a);
    extra[static_cast<size_t>(n) * C + c] = a * inv_T;
  }
}

// BN backward finalize for the gated branch: dz1[n,t,c] = g*gate[n,c] + extra[n,c]
//   sum dz1   = sum_n gate*Sg + T*extra ;  sum dz1*y = sum_n gate*Sgy + extra*Sy
__global__ void __launch_bounds__(256)
se_bn_bwd_finalize_kernel(const float* __restrict__ partials, int N, int chunks, int C, int T_len, double count,
                          const float* __restrict__ gate, const float* __restrict__ extra,
                          const float* __restrict__ sums_y, const float* __restrict__ gamma,
                          const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ coef) {
  // blockDim = (32 channels, 8 utterance lanes): the per-utterance terms are independent, so 8 lanes walk the batch in
  // parallel (a single thread per channel chained 5 dependent-latency loads x N utterances: ~60 us at N = 64) and a
  // shared-memory reduction folds them
  __shared__ double red_a[8][33], red_b[8][33];
  const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double a = 0.0, b = 0.0;
  if (c < C) {
#pragma unroll 2
    for (int n = nl; n < N; n += 8) {
      double sg = 0.0, sgy = 0.0;
      for (int k = 0; k < chunks; ++k) {
        const float* p = partials + static_cast<size_t>(n * chunks + k) * 3 * C;
        sg += p[c];
        sgy += p[C + c];
      }
      const double gt = gate[static_cast<size_t>(n) * C + c], ex = extra[static_cast<size_t>(n) * C + c];
      a += gt * sg + ex * T_len;
      b += gt * sgy + ex * sums_y[static_cast<size_t>(n) * C + c];
    }
  }
  red_a[nl][cl] = a;
  red_b[nl][cl] = b;
  __syncthreads();
  if (nl != 0 || c >= C) return;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    a += red_a[i][cl];
    b += red_b[i][cl];
  }
  const double mu = mean[c], is = invstd[c], ga = gamma[c];
  const double dga = is * (b - mu * a);
  dgamma[c] += static_cast<float>(dga);
  dbeta[c] += static_cast<float>(a);
  const double c0 = ga * is;
  const double c1 = -ga * is * is * dga / count;
  const double c2 = -ga * is * a / count - c1 * mu;
  coef[c] = static_cast<float>(c0);
  coef[C + c] = static_cast<float>(c1);
  coef[2 * C + c] = static_cast<float>(c2);
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_se_excite_fwd(const float* sums, const float* scale, const float* shift, int T, const float* w1,
                       const float* w2, float* s, float* hidden, float* gate, int N, int C, int Cr,
                       lasr_stream_t stream) {
  if (N <= 0 || C <= 0 || Cr <= 0 || T <= 0) return LASR_ERR_BAD_SHAPE;
  const int smem = (C + Cr) * static_cast<int>(sizeof(float));
  se_fwd_kernel<<<N, 256, smem, stream>>>(sums, scale, shift, 1.f / static_cast<float>(T), w1, w2, s, hidden, gate, C,
                                          Cr);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_se_excite_bwd(const float* partials, int chunks, const float* scale, const float* shift, int T,
                       const float* w1, const float* w2, const float* s, const float* hidden, const float* gate,
                       float* extra, float* dw1, float* dw2, int N, int C, int Cr, lasr_stream_t stream) {
  if (N <= 0 || C <= 0 || Cr <= 0 || T <= 0 || chunks <= 0) return LASR_ERR_BAD_SHAPE;
  const int smem = (2 * C + 2 * Cr) * static_cast<int>(sizeof(float));
  se_bwd_kernel<<<N, 256, smem, stream>>>(partials, chunks, scale, shift, 1.f / static_cast<float>(T), w1, w2, s,
                                          hidden, gate, extra, dw1, dw2, C, Cr);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_se_bn_bwd_finalize(const float* partials, int N, int chunks, int C, int T, const float* gate,
                            const float* extra, const float* sums_y, const float* gamma, const float* mean,
                            const float* invstd, float* dgamma, float* dbeta, float* coef, lasr_stream_t stream) {
  if (N <= 0 || C <= 0 || T <= 0 || chunks <= 0) return LASR_ERR_BAD_SHAPE;
  se_bn_bwd_finalize_kernel<<<cdiv(C, 32), 256, 0, stream>>>(partials, N, chunks, C, T,
                                                              static_cast<double>(N) * T, gate, extra, sums_y, gamma,
                                                              mean, invstd, dgamma, dbeta, coef);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // extern "C"
