// Squeeze-excitation (models/QuartNetContextSE.py:8-23, applied at :55 between BatchNorm and ReLU).
// The squeeze mean over time of the BatchNorm output is an affine function of the per-(utterance, channel) time
// sum of the PRE-BN tensor: mean_t(BN(y)) = scale*(sum_t y)/T + shift, so no extra pass over the BN output exists.
// Everything here works on [N, C] tensors: one CTA per utterance.
#include "common.cuh"

namespace lasr {

// s = scale*Sy/T + shift ; h = relu(W1 s) ; gate = sigmoid(W2 h)
// One CTA of 16 warps per utterance.  The MLP is 2 C Cr MACs per utterance -- nothing -- so the kernel is the sum of its
// dependent L2 round trips: the first version walked W1 one row per warp and trip and W2 one scalar per thread and
// trip (24 us under ncu at C = 512, Cr = 64).  Here W1 rows are read two per warp and trip in 16-byte vectors, and a
// thread reads its whole W2 row in 16-byte vectors with four loads in flight.
constexpr int SE_THREADS = 512;
__device__ __forceinline__ bool se_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(SE_THREADS)
se_fwd_kernel(const float* __restrict__ sums, const float* __restrict__ scale, const float* __restrict__ shift,
              float inv_T, const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ s_out,
              float* __restrict__ hidden, float* __restrict__ gate, int C, int Cr) {
  extern __shared__ __align__(16) float sm[];  // s[C], h[Cr]
  float* s = sm;
  float* h = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = fmaf(scale[c], sums[static_cast<size_t>(n) * C + c] * inv_T, shift[c]);
    s[c] = v;
    s_out[static_cast<size_t>(n) * C + c] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool vec1 = (C & 3) == 0 && se_al16(w1);
  for (int j0 = 2 * warp; j0 < Cr; j0 += 2 * nw) {
    const int j1 = min(j0 + 1, Cr - 1);
    const float* r0 = w1 + static_cast<size_t>(j0) * C;
    const float* r1 = w1 + static_cast<size_t>(j1) * C;
    float a0 = 0.f, a1 = 0.f;
    if (vec1) {
      for (int c = lane * 4; c < C; c += 128) {
        const float4 u = *reinterpret_cast<const float4*>(r0 + c), v = *reinterpret_cast<const float4*>(r1 + c);
        const float4 x = *reinterpret_cast<const float4*>(s + c);
        a0 = fmaf(u.x, x.x, fmaf(u.y, x.y, fmaf(u.z, x.z, fmaf(u.w, x.w, a0))));
        a1 = fmaf(v.x, x.x, fmaf(v.y, x.y, fmaf(v.z, x.z, fmaf(v.w, x.w, a1))));
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        a0 = fmaf(r0[c], s[c], a0);
        a1 = fmaf(r1[c], s[c], a1);
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      a0 = fmaxf(a0, 0.f);
      h[j0] = a0;
      hidden[static_cast<size_t>(n) * Cr + j0] = a0;
      if (j0 + 1 < Cr) {
        a1 = fmaxf(a1, 0.f);
        h[j0 + 1] = a1;
        hidden[static_cast<size_t>(n) * Cr + j0 + 1] = a1;
      }
    }
  }
  __syncthreads();
  const bool vec2 = (Cr & 3) == 0 && se_al16(w2);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* wr = w2 + static_cast<size_t>(c) * Cr;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (vec2) {
#pragma unroll 4
      for (int j = 0; j < Cr; j += 4) {
        const float4 u = *reinterpret_cast<const float4*>(wr + j);
        a0 = fmaf(u.x, h[j], a0);
        a1 = fmaf(u.y, h[j + 1], a1);
        a2 = fmaf(u.z, h[j + 2], a2);
        a3 = fmaf(u.w, h[j + 3], a3);
      }
    } else {
      for (int j = 0; j < Cr; ++j) a0 = fmaf(wr[j], h[j], a0);
    }
    const float a = (a0 + a1) + (a2 + a3);
    gate[static_cast<size_t>(n) * C + c] = 1.f / (1.f + expf(-a));
  }
}

// dgate[c] = scale*Sgy + shift*Sg (time sums of g*y and g for this utterance, folded from the chunk partials)
// -> dz = dgate g (1 - g) -> dh = relu'(h) W2^T dz -> ds = W1^T dh ; extra = ds / T.
// dz [N, C] and dh [N, Cr] are left in `ws` for se_wgrad_kernel: the first version added dz h^T and dh s^T to dW2 / dW1
// with one global atomic per (utterance, element) -- 4 M atomics on 64 K addresses per launch at N = 64, C = 512.
__global__ void __launch_bounds__(SE_THREADS)
se_bwd_kernel(const float* __restrict__ partials, int chunks, const float* __restrict__ scale,
              const float* __restrict__ shift, float inv_T, const float* __restrict__ w1, const float* __restrict__ w2,
              const float* __restrict__ hidden, const float* __restrict__ gate, float* __restrict__ extra,
              float* __restrict__ ws_dz, float* __restrict__ ws_dh, int C, int Cr) {
  extern __shared__ __align__(16) float sm[];  // dz[C], h[Cr], dh[Cr]
  float* dz = sm;
  float* h = sm + C;
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
    const float v = dgate * g * (1.f - g);
    dz[c] = v;
    ws_dz[static_cast<size_t>(n) * C + c] = v;
  }
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    h[j] = hidden[static_cast<size_t>(n) * Cr + j];
    dh[j] = 0.f;
  }
  __syncthreads();
  // dh[j] = sum_c W2[c, j] dz[c]: a warp walks rows c of W2 (coalesced over j), four rows in flight, and keeps partial
  // sums for j = lane + 32 k; the warps' partials meet in shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int jb = 0; jb < Cr; jb += 128) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int c = warp; c < C; c += nw) {
      const float* wr = w2 + static_cast<size_t>(c) * Cr + jb;
      const float z = dz[c];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = jb + lane + 32 * k;
        if (j < Cr) acc[k] = fmaf(wr[lane + 32 * k], z, acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = jb + lane + 32 * k;
      if (j < Cr) atomicAdd(&dh[j], acc[k]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    const float v = h[j] > 0.f ? dh[j] : 0.f;
    dh[j] = v;
    ws_dh[static_cast<size_t>(n) * Cr + j] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
#pragma unroll 4
    for (; j + 1 < Cr; j += 2) {
      a0 = fmaf(w1[static_cast<size_t>(j) * C + c], dh[j], a0);
      a1 = fmaf(w1[static_cast<size_t>(j + 1) * C + c], dh[j + 1], a1);
    }
    if (j < Cr) a0 = fmaf(w1[static_cast<size_t>(j) * C + c], dh[j], a0);
    extra[static_cast<size_t>(n) * C + c] = (a0 + a1) * inv_T;
  }
}

// dW2[c, j] += sum_n dz[n, c] h[n, j] ; dW1[j, c] += sum_n dh[n, j] s[n, c]: one thread, one element, no atomics
__global__ void __launch_bounds__(256)
se_wgrad_kernel(const float* __restrict__ ws_dz, const float* __restrict__ ws_dh, const float* __restrict__ s_in,
                const float* __restrict__ hidden, float* __restrict__ dw1, float* __restrict__ dw2, int N, int C,
                int Cr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = C * Cr;
  if (i >= 2 * total) return;
  float a0 = 0.f, a1 = 0.f;
  if (i < total) {
    const int c = i / Cr, j = i - c * Cr;
    int n = 0;
#pragma unroll 8
    for (; n + 1 < N; n += 2) {
      a0 = fmaf(ws_dz[static_cast<size_t>(n) * C + c], hidden[static_cast<size_t>(n) * Cr + j], a0);
      a1 = fmaf(ws_dz[static_cast<size_t>(n + 1) * C + c], hidden[static_cast<size_t>(n + 1) * Cr + j], a1);
    }
    if (n < N) a0 = fmaf(ws_dz[static_cast<size_t>(n) * C + c], hidden[static_cast<size_t>(n) * Cr + j], a0);
    dw2[i] += a0 + a1;
  } else {
    const int i2 = i - total;
    const int j = i2 / C, c = i2 - j * C;
    int n = 0;
#pragma unroll 8
    for (; n + 1 < N; n += 2) {
      a0 = fmaf(ws_dh[static_cast<size_t>(n) * Cr + j], s_in[static_cast<size_t>(n) * C + c], a0);
      a1 = fmaf(ws_dh[static_cast<size_t>(n + 1) * Cr + j], s_in[static_cast<size_t>(n + 1) * C + c], a1);
    }
    if (n < N) a0 = fmaf(ws_dh[static_cast<size_t>(n) * Cr + j], s_in[static_cast<size_t>(n) * C + c], a0);
    dw1[i2] += a0 + a1;
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
  if (smem > 48 * 1024) return LASR_ERR_UNSUPPORTED;
  se_fwd_kernel<<<N, SE_THREADS, smem, stream>>>(sums, scale, shift, 1.f / static_cast<float>(T), w1, w2, s, hidden,
                                                 gate, C, Cr);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_se_excite_bwd(const float* partials, int chunks, const float* scale, const float* shift, int T,
                       const float* w1, const float* w2, const float* s, const float* hidden, const float* gate,
                       float* extra, float* dw1, float* dw2, float* ws, int N, int C, int Cr, lasr_stream_t stream) {
  if (N <= 0 || C <= 0 || Cr <= 0 || T <= 0 || chunks <= 0 || ws == nullptr) return LASR_ERR_BAD_SHAPE;
  const int smem = (C + 2 * Cr) * static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) return LASR_ERR_UNSUPPORTED;
  float* ws_dz = ws;
  float* ws_dh = ws + static_cast<size_t>(N) * C;
  se_bwd_kernel<<<N, SE_THREADS, smem, stream>>>(partials, chunks, scale, shift, 1.f / static_cast<float>(T), w1, w2,
                                                 hidden, gate, extra, ws_dz, ws_dh, C, Cr);
  LASR_CHECK_LAUNCH();
  se_wgrad_kernel<<<cdiv(2 * C * Cr, 256), 256, 0, stream>>>(ws_dz, ws_dh, s, hidden, dw1, dw2, N, C, Cr);
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
