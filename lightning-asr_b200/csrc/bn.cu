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
#include "dw_common.cuh"

#include <cstdlib>

namespace lasr {

template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  // Raw = the bits as loaded: kernels issue the loads of several vectors first and unpack afterwards, so every thread
  // keeps 4-8 independent 16-byte requests in flight (the passes are pure HBM streams)
  using Raw = uint4;
  static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[8]) {
    float2 a = bf16x2_to_f32x2(u.x), b = bf16x2_to_f32x2(u.y), c = bf16x2_to_f32x2(u.z), d = bf16x2_to_f32x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) { unpack(ldraw(p), v); }
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
  struct Raw {
    float4 a, b;
  };
  static __device__ __forceinline__ Raw ldraw(const float* p) {
    Raw r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) { unpack(ldraw(p), v); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) { Vec8<float>::load(p, v); }
// 8 consecutive floats from shared memory as two 128-bit loads
__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) { Vec8<float>::load(p, v); }

// ------------------------------------------------------------------------------------------------
// Dropout (nn.Dropout(p=drop_rate), models/QuartNet.py:27,38,149) fused into the passes.  The keep mask is one byte per
// element ([M, C], 1 = keep); the forward pass either READS it (mode 1: a mask supplied by the caller, the parity hook of
// SURVEY.md 7.2-8) or GENERATES it with Philox4x32-10 keyed by (seed, vector index) and WRITES it (mode 2) for the
// backward passes.  Kept elements are scaled by 1/(1-p) like torch.
// ------------------------------------------------------------------------------------------------
struct DropArgs {
  uint8_t* mask;
  int mode;  // 0 off, 1 read, 2 generate + write
  float scale;
  uint32_t thresh;  // keep iff 16-bit uniform >= thresh  (thresh = round(p * 65536))
  unsigned long long seed;
  const unsigned long long* seed_dev;  // nullable: a device-resident step counter added to the seed (CUDA-graph replays
                                       // must not repeat the mask, and kernel arguments are frozen at capture)
};
// the 8 keep bytes of vector v (8 consecutive channels of one row)
__device__ __forceinline__ uint2 drop_generate(const DropArgs& d, long long v) {
  const unsigned long long seed = d.seed + (d.seed_dev != nullptr ? __ldg(d.seed_dev) : 0ull);
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(v), static_cast<uint32_t>(v >> 32), 0u, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t k[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t u16 = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu;
    k[i >> 2] |= (u16 >= d.thresh ? 1u : 0u) << (8 * (i & 3));
  }
  return make_uint2(k[0], k[1]);
}
__device__ __forceinline__ void drop_factors(const uint2 m, float scale, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = (((i < 4 ? m.x : m.y) >> (8 * (i & 3))) & 0xffu) ? scale : 0.f;
}

// ------------------------------------------------------------------------------------------------
// ReLU gate bits.  The backward passes only need the SIGN of the forward output (g = dout * [out > 0]); reading the
// whole bf16 tensor for it is 2 of the 10 streams of the two backward passes.  The forward pass can leave one byte per
// (frame, 8 channels) instead -- bit i = out[n, t, 8 cv + i] > 0 -- laid out [N][ceil(T / 8)][C / 8][8 frames] so that a
// thread that owns 8 consecutive frames of a channel vector writes one 8-byte word.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t relu_bits_index(int n, int t, int cv, int Tb, int CV) {
  return ((static_cast<size_t>(n) * Tb + (t >> 3)) * CV + cv) * 8 + (t & 7);
}
__device__ __forceinline__ uint32_t relu_byte(const float (&o)[8]) {
  uint32_t b = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) b |= (o[i] > 0.f ? 1u : 0u) << i;
  return b;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm coefficients.  The pointwise-GEMM epilogue leaves the batch sum / sum of squares of every channel in a
// double [2, C] buffer (RED.f64); there is NO separate finalize launch: every CTA of the apply pass folds them into
// scale = gamma * invstd, shift = beta - mean * scale in its prologue (C <= 1024 channels, a few hundred ns), and
// CTA 0 also stores mean / invstd for the backward and updates the running statistics (momentum, unbiased variance)
// and num_batches_tracked.  Eval mode (sums == NULL) reads the running statistics instead.
// ------------------------------------------------------------------------------------------------
// Raw inputs of one channel's coefficients: loaded up front (one memory round trip for both BatchNorms of a pass)
struct BnRaw {
  double s1, s2;
  float gamma, beta;
};
__device__ __forceinline__ BnRaw bn_load_raw(const lasr_bn_t& bn, int c, int C) {
  BnRaw r;
  if (bn.sums != nullptr) {
    r.s1 = bn.sums[c];
    r.s2 = bn.sums[C + c];
  } else {
    r.s1 = bn.running_mean[c];
    r.s2 = bn.running_var[c];
  }
  r.gamma = bn.gamma[c];
  r.beta = bn.beta[c];
  return r;
}
// no fp64 divisions or square roots (software sequences of ~50 dependent instructions each): 1/count comes from the
// host, 1/sqrt is rsqrt (<= 1 ulp in double, the results are rounded to fp32 anyway)
__device__ __forceinline__ void bn_coeffs_raw(const BnRaw& r, bool batch, double inv_count, float eps, float& scale,
                                              float& shift, float& mean_f, float& invstd_f, double& var_out) {
  double mean, var;
  if (batch) {
    mean = r.s1 * inv_count;
    var = r.s2 * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
  } else {
    mean = r.s1;
    var = r.s2;
  }
  const double invstd = rsqrt(var + static_cast<double>(eps));
  const double g = r.gamma, b = r.beta;
  scale = static_cast<float>(g * invstd);
  shift = static_cast<float>(b - mean * g * invstd);
  mean_f = static_cast<float>(mean);
  invstd_f = static_cast<float>(invstd);
  var_out = var;
}
__device__ __forceinline__ void bn_coeffs(const lasr_bn_t& bn, int c, int C, double count, float eps, float& scale,
                                          float& shift, float& mean_f, float& invstd_f, double& var_out) {
  bn_coeffs_raw(bn_load_raw(bn, c, C), bn.sums != nullptr, 1.0 / count, eps, scale, shift, mean_f, invstd_f, var_out);
}

__device__ __forceinline__ void bn_side_effects(const lasr_bn_t& bn, int c, double count, float momentum, float mean_f,
                                                float invstd_f, double var) {
  if (bn.sums == nullptr) return;
  if (bn.save_mean != nullptr) bn.save_mean[c] = mean_f;
  if (bn.save_invstd != nullptr) bn.save_invstd[c] = invstd_f;
  if (bn.running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    bn.running_mean[c] = static_cast<float>((1.0 - momentum) * bn.running_mean[c] + momentum * mean_f);
    bn.running_var[c] = static_cast<float>((1.0 - momentum) * bn.running_var[c] + momentum * unbiased);
  }
  if (c == 0 && bn.num_batches_tracked != nullptr) *bn.num_batches_tracked += 1;
}

// standalone coefficient kernel (used by the squeeze-excitation path, which needs scale / shift before the apply
// pass, and by tests): writes scale / shift [C] (+ the training side effects when side_effects != 0)
__global__ void bn_coeffs_kernel(const lasr_bn_t bn, int C, double count, float eps, float momentum,
                                 float* __restrict__ scale, float* __restrict__ shift, int side_effects) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc, sh, mf, isf;
  double var;
  bn_coeffs(bn, c, C, count, eps, sc, sh, mf, isf, var);
  scale[c] = sc;
  shift[c] = sh;
  if (side_effects) bn_side_effects(bn, c, count, momentum, mf, isf, var);
}

// ------------------------------------------------------------------------------------------------
// forward apply: out = act( (scale1*y + shift1) [* gate[n,c]] [+ scale2*r + shift2] )
// persistent CTAs of 512 threads; smem: scale1, shift1, scale2, shift2 [C] each
// ------------------------------------------------------------------------------------------------
template <typename T, bool HAS_R, bool HAS_GATE, bool HAS_DROP>
__global__ void __launch_bounds__(256, 2)
bn_apply_fwd_kernel(const T* __restrict__ y, const lasr_bn_t bn1, const T* __restrict__ r, const lasr_bn_t bn2,
                    const float* __restrict__ gate, T* __restrict__ out, long long total_vec, int CV, int C, int T_len,
                    double count, float eps, float momentum, int act, int side_effects, const DropArgs drop,
                    uint8_t* __restrict__ relu_bits) {
  pdl_launch_dependents();
  pdl_wait();  // the statistics come from the GEMM right before this pass
  extern __shared__ float coef_s[];  // [4][C]
  float* s_scale1 = coef_s;
  float* s_shift1 = coef_s + C;
  float* s_scale2 = coef_s + 2 * C;
  float* s_shift2 = coef_s + 3 * C;
  const double inv_count = 1.0 / count;  // one division per thread, overlapped with the loads below
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const BnRaw r1 = bn_load_raw(bn1, c, C);
    BnRaw r2{};
    if constexpr (HAS_R) r2 = bn_load_raw(bn2, c, C);
    float sc, sh, mf, isf;
    double var;
    bn_coeffs_raw(r1, bn1.sums != nullptr, inv_count, eps, sc, sh, mf, isf, var);
    s_scale1[c] = sc;
    s_shift1[c] = sh;
    if (blockIdx.x == 0 && side_effects) bn_side_effects(bn1, c, count, momentum, mf, isf, var);
    if constexpr (HAS_R) {
      bn_coeffs_raw(r2, bn2.sums != nullptr, inv_count, eps, sc, sh, mf, isf, var);
      s_scale2[c] = sc;
      s_shift2[c] = sh;
      if (blockIdx.x == 0 && side_effects) bn_side_effects(bn2, c, count, momentum, mf, isf, var);
    }
  }
  __syncthreads();
  // Every thread owns ONE 8-channel vector position (cv) for its whole life and walks rows: its coefficients live in
  // registers.  (Round 1 re-read them from shared memory for every vector -- 2-6 LDS.128 per 16 bytes of activation, which
  // made the pass L1TEX-bound at 85 % while HBM idled at ~50 %, ncu r1f -- and paid a 64-bit division per vector.)
  using Raw = typename Vec8<T>::Raw;
  constexpr int U = sizeof(T) == 2 ? 4 : 2;  // rows per thread and trip, all loads issued before the first use
  const int rpb = static_cast<int>(blockDim.x) / CV;  // rows per CTA and trip
  const int r_in = static_cast<int>(threadIdx.x) / CV;
  const int cv = static_cast<int>(threadIdx.x) - r_in * CV;
  if (r_in >= rpb) return;
  const int c = cv * 8;
  const int rows = static_cast<int>(total_vec / CV);
  float sc1[8], sh1[8], sc2[8], sh2[8];
  lds8(s_scale1 + c, sc1);
  lds8(s_shift1 + c, sh1);
  if constexpr (HAS_R) {
    lds8(s_scale2 + c, sc2);
    lds8(s_shift2 + c, sh2);
  }
  const int row_stride = static_cast<int>(gridDim.x) * rpb;
  for (int row0 = static_cast<int>(blockIdx.x) * rpb + r_in; row0 < rows; row0 += U * row_stride) {
    Raw ya[U], ra[U];
    uint2 ma[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = row0 + u * row_stride;
      if (row < rows) {
        const long long v = static_cast<long long>(row) * CV + cv;
        ya[u] = Vec8<T>::ldraw(y + v * 8);  // element offset row*C + c == 8*v
        if constexpr (HAS_R) ra[u] = Vec8<T>::ldraw(r + v * 8);
        if constexpr (HAS_DROP) {
          if (drop.mode == 1) ma[u] = *reinterpret_cast<const uint2*>(drop.mask + v * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = row0 + u * row_stride;
      if (row >= rows) break;
      const long long v = static_cast<long long>(row) * CV + cv;
      float a[8], o[8];
      Vec8<T>::unpack(ya[u], a);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(a[i], sc1[i], sh1[i]);
      if constexpr (HAS_GATE) {
        const int n = row / T_len;
        float g[8];
        load8f(gate + static_cast<size_t>(n) * C + c, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= g[i];
      }
      if constexpr (HAS_DROP) {
        // dropout sits after BN [/ SE / ReLU] and BEFORE the residual add (models/QuartNet.py:38,76); without a
        // residual relu(x)*m == relu(x*m) because m >= 0
        if (drop.mode == 2) {
          ma[u] = drop_generate(drop, v);
          *reinterpret_cast<uint2*>(drop.mask + v * 8) = ma[u];
        }
        float f[8];
        drop_factors(ma[u], drop.scale, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= f[i];
      }
      if constexpr (HAS_R) {
        float rr[8];
        Vec8<T>::unpack(ra[u], rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += fmaf(rr[i], sc2[i], sh2[i]);
      }
      if (act == LASR_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
      }
      Vec8<T>::store(out + v * 8, o);
      if (relu_bits != nullptr) {
        const int n = row / T_len;
        relu_bits[relu_bits_index(n, row - n * T_len, cv, (T_len + 7) >> 3, CV)] = static_cast<uint8_t>(relu_byte(o));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward apply that ALSO writes the channel-major "series" companion of its output (bf16 only): the operand format of
// the TMA-fed depthwise kernels (dwconv_cm.cu; layout contract in include/lasr.h, "channel-major series").
//   out  [N*T, C]  channels-last, as bn_apply_fwd_kernel
//   outT [C][N][S] frame t of utterance n at position off + t of row (c, n); positions [0, off) and [off + T, S) zero;
//                  16-byte groups of 8 positions stored at group index g ^ ((g >> 3) & 1): the tensor cores' 32-byte
//                  swizzle applied in GLOBAL memory, so that an unswizzled TMA copy of 128-position-aligned blocks lands
//                  as a SWIZZLE_32B operand.
// A depthwise conv consumes per-channel time series; in a channels-last tensor they do not exist, and gathering them
// inside the conv kernel (one L2 request per 32 useful bytes, a transposition on the critical path of a 4-item pipeline)
// is what kept that kernel at 20 % of its HBM bound in round 1.  This pass is a pure stream with every thread owning an
// 8-frame x 8-channel block: 8 row loads of 16 B per operand (a quarter-warp covers a full 128-byte line), the
// arithmetic, 8 row stores, then the 8x8 transposition with byte permutes and 8 more 16-byte stores, one per channel
// row of outT (four neighbouring lanes fill a 64-byte run).  The extra cost is one more write stream in a pass that
// already moves three.
// ------------------------------------------------------------------------------------------------
template <bool HAS_R, bool HAS_GATE>
__global__ void __launch_bounds__(256, 2)
bn_apply_fwd_cm_kernel(const __nv_bfloat16* __restrict__ y, const lasr_bn_t bn1, const __nv_bfloat16* __restrict__ r,
                       const lasr_bn_t bn2, const float* __restrict__ gate, __nv_bfloat16* __restrict__ out,
                       __nv_bfloat16* __restrict__ outT, int N, int T_len, int C, int S, int off, double count, float eps,
                       float momentum, int act, int side_effects, uint8_t* __restrict__ relu_bits,
                       const float* __restrict__ dw_w, __nv_bfloat16* __restrict__ toep,
                       __nv_bfloat16* __restrict__ toep_flip, int dwK, int dwKS, int dw_delta) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float coef_s[];  // [4][C]
  float* s_scale1 = coef_s;
  float* s_shift1 = coef_s + C;
  float* s_scale2 = coef_s + 2 * C;
  float* s_shift2 = coef_s + 3 * C;
  const double inv_count = 1.0 / count;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const BnRaw r1 = bn_load_raw(bn1, c, C);
    BnRaw r2{};
    if constexpr (HAS_R) r2 = bn_load_raw(bn2, c, C);
    float sc, sh, mf, isf;
    double var;
    bn_coeffs_raw(r1, bn1.sums != nullptr, inv_count, eps, sc, sh, mf, isf, var);
    s_scale1[c] = sc;
    s_shift1[c] = sh;
    if (blockIdx.x == 0 && side_effects) bn_side_effects(bn1, c, count, momentum, mf, isf, var);
    if constexpr (HAS_R) {
      bn_coeffs_raw(r2, bn2.sums != nullptr, inv_count, eps, sc, sh, mf, isf, var);
      s_scale2[c] = sc;
      s_shift2[c] = sh;
      if (blockIdx.x == 0 && side_effects) bn_side_effects(bn2, c, count, momentum, mf, isf, var);
    }
  }
  __syncthreads();
  // Toeplitz factors of the depthwise conv that will read outT (forward taps, and reversed: its data gradient), one
  // 16-byte chunk per thread and trip: built here, once, instead of in the prologue of every depthwise launch
  if (toep != nullptr) {
    const int chunks_c = 2 * dwKS;
    const long long total = 2ll * C * chunks_c;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < total;
         q += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int flip = q >= static_cast<long long>(C) * chunks_c ? 1 : 0;
      const long long qq = q - (flip ? static_cast<long long>(C) * chunks_c : 0);
      const int c = static_cast<int>(qq / chunks_c);
      const uint4 v = toeplitz_chunk(dw_w + static_cast<size_t>(c) * dwK, dwK, dw_delta, flip, static_cast<int>(qq - static_cast<long long>(c) * chunks_c));
      *reinterpret_cast<uint4*>((flip ? toep_flip : toep) + qq * 8) = v;
    }
  }
  // warp unit = 4 position groups (32 positions) x 8 channel vectors (64 channels); lane = (group, vector)
  const int lane = threadIdx.x & 31;
  const int gl = lane & 3, cvl = lane >> 2;
  const int CVO = C / 64;         // channel octets
  const int quads = S / 32;       // per utterance row
  const long long units = static_cast<long long>(N) * quads * CVO;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long wstride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const bool relu = act == LASR_ACT_RELU;
  for (long long u = warp0; u < units; u += wstride) {
    const int cvo = static_cast<int>(u % CVO);
    const long long rest = u / CVO;
    const int quad = static_cast<int>(rest % quads);
    const int n = static_cast<int>(rest / quads);
    const int c = (cvo * 8 + cvl) * 8;
    const int g = quad * 4 + gl;         // position group of this lane
    const int f0 = g * 8 - off;          // first frame of the group (off % 8 == 0: all 8 frames share validity but the last)
    const size_t row0 = static_cast<size_t>(n) * T_len;
    uint4 ya[8], ra[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = f0 + i;
      ya[i] = make_uint4(0u, 0u, 0u, 0u);
      if constexpr (HAS_R) ra[i] = make_uint4(0u, 0u, 0u, 0u);
      if (f >= 0 && f < T_len) {
        const size_t e = (row0 + f) * C + c;
        ya[i] = *reinterpret_cast<const uint4*>(y + e);
        if constexpr (HAS_R) ra[i] = *reinterpret_cast<const uint4*>(r + e);
      }
    }
    float sc1[8], sh1[8], sc2[8], sh2[8], gt[8];
    lds8(s_scale1 + c, sc1);
    lds8(s_shift1 + c, sh1);
    if constexpr (HAS_R) {
      lds8(s_scale2 + c, sc2);
      lds8(s_shift2 + c, sh2);
    }
    if constexpr (HAS_GATE) load8f(gate + static_cast<size_t>(n) * C + c, gt);
    uint4 o[8];
    uint32_t rb[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = f0 + i;
      float a[8], v[8];
      Vec8<__nv_bfloat16>::unpack(ya[i], a);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(a[k], sc1[k], sh1[k]);
      if constexpr (HAS_GATE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= gt[k];
      }
      if constexpr (HAS_R) {
        float rr[8];
        Vec8<__nv_bfloat16>::unpack(ra[i], rr);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += fmaf(rr[k], sc2[k], sh2[k]);
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
      }
      o[i].x = f32x2_to_bf16x2(v[0], v[1]);
      o[i].y = f32x2_to_bf16x2(v[2], v[3]);
      o[i].z = f32x2_to_bf16x2(v[4], v[5]);
      o[i].w = f32x2_to_bf16x2(v[6], v[7]);
      if (f >= 0 && f < T_len) {
        *reinterpret_cast<uint4*>(out + (row0 + f) * C + c) = o[i];
        rb[i >> 2] |= relu_byte(v) << (8 * (i & 3));
      } else {
        o[i] = make_uint4(0u, 0u, 0u, 0u);  // the conv's zero padding
      }
    }
    if (relu_bits != nullptr && f0 >= 0 && f0 < T_len)
      *reinterpret_cast<uint2*>(relu_bits + relu_bits_index(n, f0, c >> 3, (T_len + 7) >> 3, C >> 3)) = make_uint2(rb[0], rb[1]);
    // 8 frames x 8 channels -> 8 channels x 8 frames; group g is stored at its swizzled index
    const int gs = g ^ ((g >> 3) & 1);
    __nv_bfloat16* dst = outT + (static_cast<size_t>(c) * N + n) * S + static_cast<size_t>(gs) * 8;
    const size_t cstride = static_cast<size_t>(N) * S;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const uint32_t lo = (&o[2 * m].x)[q >> 1], hi = (&o[2 * m + 1].x)[q >> 1];
        w[m] = __byte_perm(lo, hi, (q & 1) ? 0x7632 : 0x5410);
      }
      *reinterpret_cast<uint4*>(dst + q * cstride) = make_uint4(w[0], w[1], w[2], w[3]);
    }
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
  __shared__ float red[256 * 8];  // the CTA's row lanes meet here: one global atomic per (CTA, channel), not per thread
  const bool active = tr < rows_par;
  const int t0 = blockIdx.y * rows_per_chunk;
  const int t1 = min(T_len, t0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const T* yp = y + static_cast<size_t>(n) * T_len * C + cv * 8;
  int t = active ? t0 + tr : t1;
  // four rows in flight per thread
  for (; t + 3 * rows_par < t1; t += 4 * rows_par) {
    float a[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) Vec8<T>::load(yp + static_cast<size_t>(t + k * rows_par) * C, a[k]);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += (a[0][i] + a[1][i]) + (a[2][i] + a[3][i]);
  }
  for (; t < t1; t += rows_par) {
    float a[8];
    Vec8<T>::load(yp + static_cast<size_t>(t) * C, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += a[i];
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(tr * CV + cv) * 8 + i] = acc[i];
  }
  __syncthreads();
  // thread c < C folds the row lanes of channel c (consecutive threads, consecutive channels)
  for (int c = threadIdx.x; c < C; c += 256) {
    float v = 0.f;
    for (int r = 0; r < rows_par; ++r) v += red[r * C + c];
    atomicAdd(sums + static_cast<size_t>(n) * C + c, v);
  }
}

// ------------------------------------------------------------------------------------------------
// backward reduce: one CTA per (utterance n, time chunk k).  With g = dout * (act ? out > 0 : 1):
//   totals[0][c] += sum g, totals[1][c] += sum g*y, totals[2][c] += sum g*r      (double, RED.f64; pre-zeroed)
//   per_n[n][0][c] += sum_t g, per_n[n][1][c] += sum_t g*y                       (float, only when per_n != NULL: SE)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

template <typename T, bool HAS_R, bool HAS_DROP>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ y,
                     const T* __restrict__ r, double* __restrict__ totals, float* __restrict__ per_n, int T_len, int C,
                     int chunks, int rows_per_chunk, int act, const uint8_t* __restrict__ drop_mask, float drop_scale,
                     const uint8_t* __restrict__ relu_bits, int T_true) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int NS = HAS_DROP ? 4 : 3;  // slots: sum g, sum g1*y, sum g*r [, sum g1]; g1 = g * dropout factor
  extern __shared__ float red[];  // [rows_par][NS][C]
  const int n = blockIdx.x / chunks, k = blockIdx.x - n * chunks;
  const int CV = C / 8;
  const int rows_par = 256 / CV;
  const int tr = threadIdx.x / CV, cv = threadIdx.x - tr * CV;
  const int t0 = k * rows_per_chunk;
  const int t1 = min(T_len, t0 + rows_per_chunk);
  float sg[8], sgy[8], sgr[8], sg1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sg[i] = sgy[i] = sgr[i] = sg1[i] = 0.f;
  if (tr < rows_par) {
    using Raw = typename Vec8<T>::Raw;
    constexpr int U = sizeof(T) == 2 ? 4 : 2;  // rows per trip: up to 16 independent 16-byte loads in flight per thread
    const T* base_g = dout + static_cast<size_t>(n) * T_len * C + cv * 8;
    const T* base_y = y + static_cast<size_t>(n) * T_len * C + cv * 8;
    const T* base_o = out + static_cast<size_t>(n) * T_len * C + cv * 8;
    const T* base_r = HAS_R ? r + static_cast<size_t>(n) * T_len * C + cv * 8 : nullptr;
    const uint8_t* base_m = HAS_DROP ? drop_mask + static_cast<size_t>(n) * T_len * C + cv * 8 : nullptr;
    const bool relu = act == LASR_ACT_RELU;
    const int Tb = (T_true + 7) >> 3;
    for (int tt = t0 + tr; tt < t1; tt += U * rows_par) {
      Raw gr[U], yr[U], orr[U], rrr[U];
      uint2 mr[U];
      uint32_t ob[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = tt + u * rows_par;
        if (t < t1) {
          const size_t off = static_cast<size_t>(t) * C;
          gr[u] = Vec8<T>::ldraw(base_g + off);
          yr[u] = Vec8<T>::ldraw(base_y + off);
          if (relu) {
            if (relu_bits != nullptr) {
              const int rowg = n * T_len + t;  // (the rows may have been flattened into one "utterance")
              const int nn = rowg / T_true;
              ob[u] = relu_bits[relu_bits_index(nn, rowg - nn * T_true, cv, Tb, CV)];
            } else {
              orr[u] = Vec8<T>::ldraw(base_o + off);
            }
          }
          if constexpr (HAS_R) rrr[u] = Vec8<T>::ldraw(base_r + off);
          if constexpr (HAS_DROP) mr[u] = *reinterpret_cast<const uint2*>(base_m + off);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = tt + u * rows_par;
        if (t >= t1) break;
        float g[8], o[8], yy[8], rr[8];
        Vec8<T>::unpack(gr[u], g);
        Vec8<T>::unpack(yr[u], yy);
        if (relu) {
          if (relu_bits != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = ((ob[u] >> i) & 1u) ? g[i] : 0.f;
          } else {
            Vec8<T>::unpack(orr[u], o);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
          }
        }
        if constexpr (HAS_DROP) {
          // the gated / normalised branch sees g1 = g * m / (1-p); the residual branch sees g
          float f[8];
          drop_factors(mr[u], drop_scale, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float g1 = g[i] * f[i];
            sg[i] += g[i];
            sg1[i] += g1;
            sgy[i] = fmaf(g1, yy[i], sgy[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sg[i] += g[i];
            sgy[i] = fmaf(g[i], yy[i], sgy[i]);
          }
        }
        if constexpr (HAS_R) {
          Vec8<T>::unpack(rrr[u], rr);
#pragma unroll
          for (int i = 0; i < 8; ++i) sgr[i] = fmaf(g[i], rr[i], sgr[i]);
        }
      }
    }
    float* dst = red + static_cast<size_t>(tr) * NS * C + cv * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dst[i] = sg[i];
      dst[C + i] = sgy[i];
      dst[2 * C + i] = sgr[i];
      if constexpr (HAS_DROP) dst[3 * C + i] = sg1[i];
    }
  }
  __syncthreads();
  const int nred = HAS_DROP ? 4 * C : (HAS_R ? 3 * C : 2 * C);
  for (int i = threadIdx.x; i < nred; i += 256) {
    float s = 0.f;
    for (int j = 0; j < rows_par; ++j) s += red[static_cast<size_t>(j) * NS * C + i];
    red_add_f64(totals + i, static_cast<double>(s));
    if (per_n != nullptr) {
      // the SE branch's per-utterance sums are those of its own upstream gradient g1: slot 0 <- sum g1, slot 1 <- sum g1*y
      if constexpr (HAS_DROP) {
        if (i >= 3 * C) atomicAdd(per_n + static_cast<size_t>(n) * 3 * C + (i - 3 * C), s);
        else if (i >= C && i < 2 * C) atomicAdd(per_n + static_cast<size_t>(n) * 3 * C + i, s);
      } else {
        if (i < 2 * C) atomicAdd(per_n + static_cast<size_t>(n) * 3 * C + i, s);
      }
    }
  }
}

// coefficients of the BatchNorm input gradient: d(input) = c0*g + c1*x + c2, from a = sum g and b = sum g*x
__device__ __forceinline__ void bn_bwd_coef(double a, double b, double count, float gamma, float mean, float invstd,
                                            float& c0, float& c1, float& c2, float& dgamma, float& dbeta) {
  const double mu = mean, is = invstd, ga = gamma;
  const double inv_count = 1.0 / count;
  const double dga = is * (b - mu * a);
  const double k0 = ga * is;
  const double k1 = -ga * is * is * dga * inv_count;
  const double k2 = -ga * is * a * inv_count - k1 * mu;
  c0 = static_cast<float>(k0);
  c1 = static_cast<float>(k1);
  c2 = static_cast<float>(k2);
  dgamma = static_cast<float>(dga);
  dbeta = static_cast<float>(a);
}

// standalone variant (tests / tools): coef [3, C] from totals slots (0, slot_gx)
__global__ void bn_bwd_coef_kernel(const double* __restrict__ totals, int C, double count, int slot_gx,
                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ invstd, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float c0, c1, c2, dg, db;
  bn_bwd_coef(totals[c], totals[slot_gx * C + c], count, gamma[c], mean[c], invstd[c], c0, c1, c2, dg, db);
  coef[c] = c0;
  coef[C + c] = c1;
  coef[2 * C + c] = c2;
  if (dgamma != nullptr) dgamma[c] += dg;
  if (dbeta != nullptr) dbeta[c] += db;
}

// backward apply: dy = mask(coef1[0]*(g*gate + extra) + coef1[1]*y + coef1[2]); dr = coef2[0]*g + coef2[1]*r + coef2[2]
// coefficients come from `totals` (folded in the prologue; CTA 0 accumulates dgamma / dbeta) unless coef1_in is given
// (the squeeze-excitation branch, whose upstream gradient is not g).
struct BnBwdSide {
  const float* gamma;
  const float* mean;
  const float* invstd;
  float* dgamma;
  float* dbeta;
};

template <typename T, bool HAS_R, bool HAS_GATE, bool HAS_DROP>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ y,
                    const T* __restrict__ r, const float* __restrict__ gate, const float* __restrict__ extra,
                    const double* __restrict__ totals, const float* __restrict__ coef1_in, const BnBwdSide bn1,
                    const BnBwdSide bn2, double count, const int32_t* __restrict__ lengths, int T_len,
                    T* __restrict__ dy, T* __restrict__ dr, long long total_vec, int CV, int C, int act,
                    const uint8_t* __restrict__ drop_mask, float drop_scale, const uint8_t* __restrict__ relu_bits) {
  pdl_launch_dependents();
  pdl_wait();  // totals come from the reduce pass right before
  extern __shared__ float coef_s[];  // coef1 [3][C], coef2 [3][C]
  float* k1 = coef_s;
  float* k2 = coef_s + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float c0, c1, c2, dg, db;
    // all global loads of this channel first: one round trip instead of one per BatchNorm
    double t_a = 0.0, t_a1 = 0.0, t_b1 = 0.0, t_b2 = 0.0;
    float g1 = 0.f, m1 = 0.f, i1 = 0.f, g2 = 0.f, m2 = 0.f, i2 = 0.f;
    if (totals != nullptr) {
      t_a = totals[c];
      t_a1 = HAS_DROP ? totals[3 * C + c] : t_a;  // sum of the normalised branch's own upstream gradient
    }
    if (coef1_in == nullptr) {
      t_b1 = totals[C + c];
      g1 = bn1.gamma[c];
      m1 = bn1.mean[c];
      i1 = bn1.invstd[c];
    }
    if constexpr (HAS_R) {
      t_b2 = totals[2 * C + c];
      g2 = bn2.gamma[c];
      m2 = bn2.mean[c];
      i2 = bn2.invstd[c];
    }
    if (coef1_in != nullptr) {
      k1[c] = coef1_in[c];
      k1[C + c] = coef1_in[C + c];
      k1[2 * C + c] = coef1_in[2 * C + c];
    } else {
      bn_bwd_coef(t_a1, t_b1, count, g1, m1, i1, c0, c1, c2, dg, db);
      k1[c] = c0;
      k1[C + c] = c1;
      k1[2 * C + c] = c2;
      if (blockIdx.x == 0) {
        if (bn1.dgamma != nullptr) bn1.dgamma[c] += dg;
        if (bn1.dbeta != nullptr) bn1.dbeta[c] += db;
      }
    }
    if constexpr (HAS_R) {
      bn_bwd_coef(t_a, t_b2, count, g2, m2, i2, c0, c1, c2, dg, db);
      k2[c] = c0;
      k2[C + c] = c1;
      k2[2 * C + c] = c2;
      if (blockIdx.x == 0) {
        if (bn2.dgamma != nullptr) bn2.dgamma[c] += dg;
        if (bn2.dbeta != nullptr) bn2.dbeta[c] += db;
      }
    }
  }
  __syncthreads();
  // one 8-channel vector position per thread for its whole life, coefficients in registers (see bn_apply_fwd_kernel)
  using Raw = typename Vec8<T>::Raw;
  constexpr int U = sizeof(T) == 2 ? 2 : 1;
  const bool relu = act == LASR_ACT_RELU;
  const int rpb = static_cast<int>(blockDim.x) / CV;
  const int r_in = static_cast<int>(threadIdx.x) / CV;
  const int cv = static_cast<int>(threadIdx.x) - r_in * CV;
  if (r_in >= rpb) return;
  const int c = cv * 8;
  const int rows = static_cast<int>(total_vec / CV);
  float a0[8], a1[8], a2[8], b0[8], b1[8], b2[8];
  lds8(k1 + c, a0);
  lds8(k1 + C + c, a1);
  lds8(k1 + 2 * C + c, a2);
  if constexpr (HAS_R) {
    lds8(k2 + c, b0);
    lds8(k2 + C + c, b1);
    lds8(k2 + 2 * C + c, b2);
  }
  const int row_stride = static_cast<int>(gridDim.x) * rpb;
  for (int row0 = static_cast<int>(blockIdx.x) * rpb + r_in; row0 < rows; row0 += U * row_stride) {
    Raw gr[U], orr[U], yr[U], rrr[U];
    uint2 mr[U];
    uint32_t ob[U];
    bool keep_u[U];
    int n_u[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = row0 + u * row_stride;
      keep_u[u] = false;
      if (row < rows) {
        const long long v = static_cast<long long>(row) * CV + cv;
        n_u[u] = row / T_len;
        const int t = row - n_u[u] * T_len;
        keep_u[u] = lengths == nullptr || t < lengths[n_u[u]];
        gr[u] = Vec8<T>::ldraw(dout + v * 8);
        if (relu) {
          if (relu_bits != nullptr)
            ob[u] = relu_bits[relu_bits_index(n_u[u], t, cv, (T_len + 7) >> 3, CV)];
          else
            orr[u] = Vec8<T>::ldraw(out + v * 8);
        }
        if constexpr (HAS_R) rrr[u] = Vec8<T>::ldraw(r + v * 8);
        if (keep_u[u]) yr[u] = Vec8<T>::ldraw(y + v * 8);
        if constexpr (HAS_DROP) mr[u] = *reinterpret_cast<const uint2*>(drop_mask + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = row0 + u * row_stride;
      if (row >= rows) break;
      const long long v = static_cast<long long>(row) * CV + cv;
      const int n = n_u[u];
      const bool keep = keep_u[u];
      float g[8];
      Vec8<T>::unpack(gr[u], g);
      if (relu) {
        if (relu_bits != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = ((ob[u] >> i) & 1u) ? g[i] : 0.f;
        } else {
          float o[8];
          Vec8<T>::unpack(orr[u], o);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
        }
      }
      if constexpr (HAS_R) {
        float rr[8], d[8];
        Vec8<T>::unpack(rrr[u], rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(b0[i], g[i], fmaf(b1[i], rr[i], b2[i]));
        Vec8<T>::store(dr + v * 8, d);
      }
      float d[8];
      if (keep) {
        float yy[8];
        Vec8<T>::unpack(yr[u], yy);
        if constexpr (HAS_DROP) {
          float f[8];
          drop_factors(mr[u], drop_scale, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] *= f[i];
        }
        if constexpr (HAS_GATE) {
          float gt[8], ex[8];
          load8f(gate + static_cast<size_t>(n) * C + c, gt);
          load8f(extra + static_cast<size_t>(n) * C + c, ex);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = fmaf(g[i], gt[i], ex[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(a0[i], g[i], fmaf(a1[i], yy[i], a2[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 0.f;
      }
      Vec8<T>::store(dy + v * 8, d);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Ring variants (round 2).  ncu on the passes above: DRAM 41-45 %, issue slots 42-49 %, and of the ~8.5 cycles between
// two issues of a warp 4 are `long_scoreboard` -- a thread loads a trip's vectors into registers, waits, computes,
// stores, and only then asks for the next trip: the memory pipe drains once per trip, and the 124 registers that hold
// the raw vectors + coefficients cap the SM at 16 warps, too few to cover it.  Here every thread streams ITS OWN vectors
// through a private slot ring in shared memory with cp.async (16 B per request, 3 trips ahead): the requests of trips
// i+1 .. i+3 are in flight while trip i is unpacked (no CTA barrier -- a thread only ever reads what it asked for
// itself, `cp.async.wait_group` is enough), ~110-150 KB per SM permanently in flight instead of a burst per trip.
// The hot bf16 training variant of the backward APPLY pass only (no SE gate, no dropout, ReLU from the gate bits);
// everything else keeps the register kernels.  Same arithmetic in the same order: bit-identical outputs; LASR_BN_RING=0
// turns it off.  Measured (cold L2, N = 32, T' = 801, minus the 6.2 us event overhead): C = 512 28.6 -> 25.5 us
// (5.1 TB/s of algorithmic bytes), C = 256 16.3 -> 14.3 us.  The same ring under the backward REDUCE pass measured
// equal-to-slower (22.5 us for 80 MB at C = 512 either way: that pass already keeps 16 loads per thread in flight and is
// paced by its launch ramp and the fp64 reduction tail) and was dropped; the forward pass issues 16 loads per thread
// per unit as well.
// ------------------------------------------------------------------------------------------------
constexpr int RING_D = 4;  // trips in the ring (3 in flight + the one being consumed)
constexpr int RING_U = 2;  // rows per trip and thread
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
// L2 eviction hints: a stream that nobody reads again after this pass should not push the tensors the NEXT kernel needs
// out of the 126 MB L2 (the backward apply pass reads dout, y, r for the last time and writes dy, dr for the two GEMMs
// that follow)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_16_hint(uint32_t smem_addr, const void* gsrc, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gsrc), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
static bool bn_ring_enabled() {
  static const bool on = !(getenv("LASR_BN_RING") != nullptr && atoi(getenv("LASR_BN_RING")) == 0);
  return on;
}

template <bool HAS_R>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_ring_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                         const __nv_bfloat16* __restrict__ r, const double* __restrict__ totals,
                         const float* __restrict__ coef1_in, const BnBwdSide bn1, const BnBwdSide bn2, double count,
                         const int32_t* __restrict__ lengths, int T_len, __nv_bfloat16* __restrict__ dy,
                         __nv_bfloat16* __restrict__ dr, long long total_vec, int CV, int C,
                         const uint8_t* __restrict__ relu_bits, int dead_hint) {
  using T = __nv_bfloat16;
  pdl_launch_dependents();
  pdl_wait();  // totals come from the reduce pass right before
  extern __shared__ float coef_s[];  // coef1 [3][C], coef2 [3][C], then the slot ring
  float* k1 = coef_s;
  float* k2 = coef_s + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float c0, c1, c2, dg, db;
    double t_a = totals[c], t_b1 = 0.0, t_b2 = 0.0;
    float g1 = 0.f, m1 = 0.f, i1 = 0.f, g2 = 0.f, m2 = 0.f, i2 = 0.f;
    if (coef1_in == nullptr) {
      t_b1 = totals[C + c];
      g1 = bn1.gamma[c];
      m1 = bn1.mean[c];
      i1 = bn1.invstd[c];
    }
    if constexpr (HAS_R) {
      t_b2 = totals[2 * C + c];
      g2 = bn2.gamma[c];
      m2 = bn2.mean[c];
      i2 = bn2.invstd[c];
    }
    if (coef1_in != nullptr) {
      k1[c] = coef1_in[c];
      k1[C + c] = coef1_in[C + c];
      k1[2 * C + c] = coef1_in[2 * C + c];
    } else {
      bn_bwd_coef(t_a, t_b1, count, g1, m1, i1, c0, c1, c2, dg, db);
      k1[c] = c0;
      k1[C + c] = c1;
      k1[2 * C + c] = c2;
      if (blockIdx.x == 0) {
        if (bn1.dgamma != nullptr) bn1.dgamma[c] += dg;
        if (bn1.dbeta != nullptr) bn1.dbeta[c] += db;
      }
    }
    if constexpr (HAS_R) {
      bn_bwd_coef(t_a, t_b2, count, g2, m2, i2, c0, c1, c2, dg, db);
      k2[c] = c0;
      k2[C + c] = c1;
      k2[2 * C + c] = c2;
      if (blockIdx.x == 0) {
        if (bn2.dgamma != nullptr) bn2.dgamma[c] += dg;
        if (bn2.dbeta != nullptr) bn2.dbeta[c] += db;
      }
    }
  }
  __syncthreads();
  const int rpb = static_cast<int>(blockDim.x) / CV;
  const int r_in = static_cast<int>(threadIdx.x) / CV;
  const int cv = static_cast<int>(threadIdx.x) - r_in * CV;
  if (r_in >= rpb) return;
  const int c = cv * 8;
  const int rows = static_cast<int>(total_vec / CV);
  float a0[8], a1[8], a2[8], b0[8], b1[8], b2[8];
  lds8(k1 + c, a0);
  lds8(k1 + C + c, a1);
  lds8(k1 + 2 * C + c, a2);
  if constexpr (HAS_R) {
    lds8(k2 + c, b0);
    lds8(k2 + C + c, b1);
    lds8(k2 + 2 * C + c, b2);
  }
  constexpr int K = HAS_R ? 3 : 2;  // vectors per row: dout, y [, r]
  // slot (stage, u, k) of this thread: consecutive threads 16 B apart (conflict-free 128-bit accesses)
  const uint32_t ring = smem_u32(coef_s + 6 * C) + 16u * threadIdx.x;
  auto slot = [&](int st, int u, int k) { return ring + 16u * 256u * static_cast<uint32_t>((st * RING_U + u) * K + k); };
  const int row_stride = static_cast<int>(gridDim.x) * rpb;
  const int first = static_cast<int>(blockIdx.x) * rpb + r_in;
  const int Tb = (T_len + 7) >> 3;
  // per-stage bookkeeping in registers (the loop below is unrolled over the ring: compile-time indices)
  uint32_t ob[RING_D][RING_U];
  int len_n[RING_D][RING_U], t_of[RING_D][RING_U];
  const uint64_t dead = l2_policy_evict_first();
  auto issue = [&](int trip, int st) {
#pragma unroll
    for (int u = 0; u < RING_U; ++u) {
      const int row = first + (trip * RING_U + u) * row_stride;
      if (row < rows) {
        const long long v = static_cast<long long>(row) * CV + cv;
        const int n = row / T_len;
        const int t = row - n * T_len;
        if (dead_hint) {
          cp_async_16_hint(slot(st, u, 0), dout + v * 8, dead);
          cp_async_16_hint(slot(st, u, 1), y + v * 8, dead);  // (also for padded frames: `keep` is known later)
          if constexpr (HAS_R) cp_async_16_hint(slot(st, u, 2), r + v * 8, dead);
        } else {
          cp_async_16(slot(st, u, 0), dout + v * 8);
          cp_async_16(slot(st, u, 1), y + v * 8);
          if constexpr (HAS_R) cp_async_16(slot(st, u, 2), r + v * 8);
        }
        ob[st][u] = relu_bits != nullptr ? relu_bits[relu_bits_index(n, t, cv, Tb, CV)] : 0xffu;
        len_n[st][u] = lengths != nullptr ? lengths[n] : T_len;
        t_of[st][u] = t;
      }
    }
    cp_async_commit_group();
  };
  auto consume = [&](int trip, int st) {
#pragma unroll
    for (int u = 0; u < RING_U; ++u) {
      const int row = first + (trip * RING_U + u) * row_stride;
      if (row >= rows) break;
      const long long v = static_cast<long long>(row) * CV + cv;
      float g[8];
      Vec8<T>::unpack(lds_u4(slot(st, u, 0)), g);
      const uint32_t bits = ob[st][u];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = ((bits >> i) & 1u) ? g[i] : 0.f;
      if constexpr (HAS_R) {
        float rr[8], d[8];
        Vec8<T>::unpack(lds_u4(slot(st, u, 2)), rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(b0[i], g[i], fmaf(b1[i], rr[i], b2[i]));
        Vec8<T>::store(dr + v * 8, d);
      }
      float d[8];
      if (t_of[st][u] < len_n[st][u]) {
        float yy[8];
        Vec8<T>::unpack(lds_u4(slot(st, u, 1)), yy);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(a0[i], g[i], fmaf(a1[i], yy[i], a2[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 0.f;
      }
      Vec8<T>::store(dy + v * 8, d);
    }
  };
  const int per_trip = RING_U * row_stride;
  const int ntrips = first < rows ? (rows - first + per_trip - 1) / per_trip : 0;
#pragma unroll
  for (int s = 0; s < RING_D - 1; ++s) {
    if (s < ntrips)
      issue(s, s);
    else
      cp_async_commit_group();
  }
#pragma unroll 1
  for (int i = 0; i < ntrips; i += RING_D) {
#pragma unroll
    for (int d = 0; d < RING_D; ++d) {
      const int trip = i + d;
      if (trip < ntrips) {
        if (trip + RING_D - 1 < ntrips)
          issue(trip + RING_D - 1, (d + RING_D - 1) % RING_D);
        else
          cp_async_commit_group();
        cp_async_wait_group<RING_D - 1>();
        consume(trip, d);
      }
    }
  }
}

constexpr int BN_THREADS = 256;  // 2 CTAs / SM x 256 threads x <= 128 registers: coefficients + 8 vectors in flight per thread
static inline int bn_block(int CV) { return CV >= BN_THREADS ? CV : BN_THREADS / CV * CV; }
static inline int persistent_grid(long long total_vec, int threads) {
  long long b = (total_vec + threads - 1) / threads;
  const long long cap = 2LL * sm_budget();
  return static_cast<int>(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <typename TT>
static int bn_fwd_launch(const void* y, const lasr_bn_t& b1, const void* r, const lasr_bn_t& b2, const float* gate,
                         void* out, long long total, int CV, int C, int T, double count, float eps, float momentum,
                         int act, int side_effects, const DropArgs& drop, uint8_t* relu_bits, cudaStream_t stream) {
  const int threads = bn_block(CV);
  const int grid = persistent_grid(total, threads);
  const int smem = 4 * C * static_cast<int>(sizeof(float));
  const TT* yy = static_cast<const TT*>(y);
  const TT* rr = static_cast<const TT*>(r);
  TT* oo = static_cast<TT*>(out);
  cudaError_t le;
#define LASR_BN_FWD(R, G, D)                                                                                      \
  le = launch_pdl(4, bn_apply_fwd_kernel<TT, R, G, D>, dim3(grid), dim3(threads), smem, stream, yy, b1, rr, b2, gate, oo, \
                  total, CV, C, T, count, eps, momentum, act, side_effects, drop, relu_bits)
  const int sel = (r != nullptr ? 4 : 0) | (gate != nullptr ? 2 : 0) | (drop.mode != 0 ? 1 : 0);
  switch (sel) {
    case 0: LASR_BN_FWD(false, false, false); break;
    case 1: LASR_BN_FWD(false, false, true); break;
    case 2: LASR_BN_FWD(false, true, false); break;
    case 3: LASR_BN_FWD(false, true, true); break;
    case 4: LASR_BN_FWD(true, false, false); break;
    case 5: LASR_BN_FWD(true, false, true); break;
    case 6: LASR_BN_FWD(true, true, false); break;
    default: LASR_BN_FWD(true, true, true); break;
  }
#undef LASR_BN_FWD
  LASR_CHECK_PDL(le);
  return LASR_OK;
}

template <typename TT>
static int bn_bwd_launch(const void* dout, const void* out, const void* y, const void* r, const float* gate,
                         const float* extra, const double* totals, const float* coef1, const BnBwdSide& s1,
                         const BnBwdSide& s2, double count, const int32_t* lengths, int T, void* dy, void* dr,
                         long long total, int CV, int C, int act, const uint8_t* drop_mask, float drop_scale,
                         const uint8_t* relu_bits, cudaStream_t stream) {
  const int threads = bn_block(CV);
  const int grid = persistent_grid(total, threads);
  const int smem = 6 * C * static_cast<int>(sizeof(float));
  const TT* a0 = static_cast<const TT*>(dout);
  const TT* a1 = static_cast<const TT*>(out);
  const TT* a2 = static_cast<const TT*>(y);
  const TT* a3 = static_cast<const TT*>(r);
  TT* o0 = static_cast<TT*>(dy);
  TT* o1 = static_cast<TT*>(dr);
  cudaError_t le;
  if constexpr (sizeof(TT) == 2) {
    // the hot training variant streams through a per-thread cp.async ring (bn_bwd_apply_ring_kernel)
    const bool relu_ok = act == LASR_ACT_RELU ? relu_bits != nullptr : true;
    const int ring_smem = smem + RING_D * RING_U * (r != nullptr ? 3 : 2) * 256 * 16;
    if (bn_ring_enabled() && gate == nullptr && drop_mask == nullptr && relu_ok && totals != nullptr && threads <= 256 &&
        ring_smem <= 113 * 1024) {
      const uint8_t* bits = act == LASR_ACT_RELU ? relu_bits : nullptr;
      static const int hint = getenv("LASR_L2_HINTS") != nullptr ? atoi(getenv("LASR_L2_HINTS")) : 1;
      static bool configured = false;
      if (!configured) {
        cudaFuncSetAttribute(bn_bwd_apply_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        cudaFuncSetAttribute(bn_bwd_apply_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        configured = true;
      }
      if (r != nullptr)
        le = launch_pdl(4, bn_bwd_apply_ring_kernel<true>, dim3(grid), dim3(threads), ring_smem, stream, a0, a2, a3, totals,
                        coef1, s1, s2, count, lengths, T, o0, o1, total, CV, C, bits, hint);
      else
        le = launch_pdl(4, bn_bwd_apply_ring_kernel<false>, dim3(grid), dim3(threads), ring_smem, stream, a0, a2, a3,
                        totals, coef1, s1, s2, count, lengths, T, o0, o1, total, CV, C, bits, hint);
      LASR_CHECK_PDL(le);
      return LASR_OK;
    }
  }
#define LASR_BN_BWD(R, G, D)                                                                                       \
  le = launch_pdl(4, bn_bwd_apply_kernel<TT, R, G, D>, dim3(grid), dim3(threads), smem, stream, a0, a1, a2, a3, gate,   \
                  extra, totals, coef1, s1, s2, count, lengths, T, o0, o1, total, CV, C, act, drop_mask, drop_scale,   \
                  relu_bits)
  const int sel = (r != nullptr ? 4 : 0) | (gate != nullptr ? 2 : 0) | (drop_mask != nullptr ? 1 : 0);
  switch (sel) {
    case 0: LASR_BN_BWD(false, false, false); break;
    case 1: LASR_BN_BWD(false, false, true); break;
    case 2: LASR_BN_BWD(false, true, false); break;
    case 3: LASR_BN_BWD(false, true, true); break;
    case 4: LASR_BN_BWD(true, false, false); break;
    case 5: LASR_BN_BWD(true, false, true); break;
    case 6: LASR_BN_BWD(true, true, false); break;
    default: LASR_BN_BWD(true, true, true); break;
  }
#undef LASR_BN_BWD
  LASR_CHECK_PDL(le);
  return LASR_OK;
}

// host view of lasr_dropout_t
static int drop_args(const lasr_dropout_t* d, DropArgs& a) {
  a = DropArgs{};
  if (d == nullptr || d->mode == 0 || d->p <= 0.f) return LASR_OK;
  if (d->mask == nullptr || d->p >= 1.f || d->mode < 0 || d->mode > 2) return LASR_ERR_BAD_SHAPE;
  a.mask = d->mask;
  a.mode = d->mode;
  a.scale = 1.f / (1.f - d->p);
  a.thresh = static_cast<uint32_t>(d->p * 65536.f + 0.5f);
  a.seed = d->seed;
  a.seed_dev = reinterpret_cast<const unsigned long long*>(d->seed_dev);
  return LASR_OK;
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_bn_coeffs(const lasr_bn_t* bn, int C, int count, float eps, float momentum, float* scale, float* shift,
                   int side_effects, lasr_stream_t stream) {
  if (bn == nullptr || C <= 0 || count <= 0) return LASR_ERR_BAD_SHAPE;
  bn_coeffs_kernel<<<cdiv(C, 128), 128, 0, stream>>>(*bn, C, static_cast<double>(count), eps, momentum, scale, shift,
                                                      side_effects);
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
  int chunks = cdiv(4 * sm_budget(), N);
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

int lasr_bn_apply_act_fwd(const void* y, const lasr_bn_t* bn1, const void* r, const lasr_bn_t* bn2, const float* gate,
                          void* out, int M, int C, int T, int count, float eps, float momentum, int act,
                          int side_effects, const lasr_dropout_t* drop, int dtype, uint8_t* relu_bits,
                          lasr_stream_t stream) {
  if (M <= 0 || C <= 0 || (C % 8) || C > 2048 || count <= 0 || bn1 == nullptr) return LASR_ERR_BAD_SHAPE;
  if (relu_bits != nullptr && (T <= 0 || (M % T) != 0)) return LASR_ERR_BAD_SHAPE;
  DropArgs da;
  if (int rc = drop_args(drop, da)) return rc;
  if ((r != nullptr) != (bn2 != nullptr)) return LASR_ERR_BAD_SHAPE;
  if (gate != nullptr && T <= 0) return LASR_ERR_BAD_SHAPE;
  const int CV = C / 8;
  const long long total = static_cast<long long>(M) * CV;
  const lasr_bn_t none{};
  const lasr_bn_t& b2 = bn2 ? *bn2 : none;
  if (dtype == LASR_F32)
    return bn_fwd_launch<float>(y, *bn1, r, b2, gate, out, total, CV, C, T, count, eps, momentum, act, side_effects,
                                da, relu_bits, stream);
  if (dtype == LASR_BF16)
    return bn_fwd_launch<__nv_bfloat16>(y, *bn1, r, b2, gate, out, total, CV, C, T, count, eps, momentum, act,
                                        side_effects, da, relu_bits, stream);
  return LASR_ERR_BAD_DTYPE;
}

int lasr_bn_apply_act_fwd_cm(const void* y, const lasr_bn_t* bn1, const void* r, const lasr_bn_t* bn2, const float* gate,
                             void* out, void* outT, int N, int T, int C, int S, int off, float eps, float momentum,
                             int act, int side_effects, uint8_t* relu_bits, const float* dw_w, void* toep,
                             void* toep_flip, int dw_K, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || (C % 64) || C > 2048 || bn1 == nullptr || outT == nullptr) return LASR_ERR_BAD_SHAPE;
  if ((r != nullptr) != (bn2 != nullptr)) return LASR_ERR_BAD_SHAPE;
  if (S <= 0 || (S % 128) || off < 0 || (off % 8) || off + T > S) return LASR_ERR_BAD_SHAPE;
  if ((toep != nullptr) != (toep_flip != nullptr) || (toep != nullptr && (dw_w == nullptr || dw_K < 3))) return LASR_ERR_BAD_SHAPE;
  const int dwKS = toep != nullptr ? lasr_cm_ks_host(dw_K) : 0;
  const int dw_delta = toep != nullptr ? off - dw_K / 2 : 0;
  __nv_bfloat16* tp = static_cast<__nv_bfloat16*>(toep);
  __nv_bfloat16* tpf = static_cast<__nv_bfloat16*>(toep_flip);
  const lasr_bn_t none{};
  const lasr_bn_t& b2 = bn2 ? *bn2 : none;
  const long long units = static_cast<long long>(N) * (S / 32) * (C / 64);
  long long ctas = (units + 7) / 8;
  if (ctas > 2 * sm_budget()) ctas = 2 * sm_budget();
  const int smem = 4 * C * static_cast<int>(sizeof(float));
  const double count = static_cast<double>(N) * T;
  const __nv_bfloat16* yy = static_cast<const __nv_bfloat16*>(y);
  const __nv_bfloat16* rr = static_cast<const __nv_bfloat16*>(r);
  __nv_bfloat16* oo = static_cast<__nv_bfloat16*>(out);
  __nv_bfloat16* ot = static_cast<__nv_bfloat16*>(outT);
  cudaError_t le;
#define LASR_BN_CM(R, G)                                                                                           \
  le = launch_pdl(4, bn_apply_fwd_cm_kernel<R, G>, dim3(static_cast<unsigned>(ctas)), dim3(256), smem, stream, yy, *bn1, \
                  rr, b2, gate, oo, ot, N, T, C, S, off, count, eps, momentum, act, side_effects, relu_bits, dw_w, tp, tpf, \
                  dw_K, dwKS, dw_delta)
  if (r != nullptr && gate != nullptr) LASR_BN_CM(true, true);
  else if (r != nullptr) LASR_BN_CM(true, false);
  else if (gate != nullptr) LASR_BN_CM(false, true);
  else LASR_BN_CM(false, false);
#undef LASR_BN_CM
  LASR_CHECK_PDL(le);
  return LASR_OK;
}

int lasr_bn_bwd_chunks(int N, int T) {
  int chunks = cdiv(4 * sm_budget(), N);
  const int max_chunks = cdiv(T, 8);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows = cdiv(T, chunks);
  return cdiv(T, rows);
}

int lasr_bn_act_bwd_reduce(const void* dout, const void* out, const void* y, const void* r, double* totals,
                           float* per_n, int N, int T, int C, int act, const lasr_dropout_t* drop, int dtype,
                           const uint8_t* relu_bits, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || (C % 8) || C > 2048 || totals == nullptr) return LASR_ERR_BAD_SHAPE;
  if (act == LASR_ACT_RELU && out == nullptr && relu_bits == nullptr) return LASR_ERR_BAD_SHAPE;
  const int T_true = T;
  DropArgs da;
  if (int rc = drop_args(drop, da)) return rc;
  const uint8_t* dmask = da.mode ? da.mask : nullptr;
  const float dscale = da.scale;
  // Without per-utterance sums (no SE) the rows are one flat [N*T, C] matrix: two fat CTAs per SM instead of 4 thin
  // ones per SM and utterance halve the fp64 REDs that all land on the same 3*C addresses.
  if (per_n == nullptr) {
    const long long rows = static_cast<long long>(N) * T;
    if (rows < (1ll << 30)) {
      T = static_cast<int>(rows);
      N = 1;
    }
  }
  int chunks = lasr_bn_bwd_chunks(N, T);
  if (N == 1) {
    static const int mult = getenv("LASR_BN_RED_CTAS") ? atoi(getenv("LASR_BN_RED_CTAS")) : 2;
    chunks = mult * sm_budget();
    if (chunks > cdiv(T, 8)) chunks = cdiv(T, 8);
    chunks = cdiv(T, cdiv(T, chunks));
  }
  const int rows_per_chunk = cdiv(T, chunks);
  const int CV = C / 8;
  const int rows_par = 256 / CV;
  const int smem = rows_par * (dmask ? 4 : 3) * C * static_cast<int>(sizeof(float));
  const int grid = N * chunks;
  cudaError_t le = cudaSuccess;
#define LASR_BN_RED_ONE(TT, R, D)                                                                                  \
  le = launch_pdl(4, bn_bwd_reduce_kernel<TT, R, D>, dim3(grid), dim3(256), smem, stream,                          \
                  static_cast<const TT*>(dout), static_cast<const TT*>(out), static_cast<const TT*>(y),            \
                  static_cast<const TT*>(r), totals, per_n, T, C, chunks, rows_per_chunk, act, dmask, dscale,      \
                  relu_bits, T_true)
#define LASR_BN_RED_LAUNCH(TT)                                                                                     \
  do {                                                                                                             \
    if (r != nullptr && dmask != nullptr) LASR_BN_RED_ONE(TT, true, true);                                         \
    else if (r != nullptr) LASR_BN_RED_ONE(TT, true, false);                                                       \
    else if (dmask != nullptr) LASR_BN_RED_ONE(TT, false, true);                                                   \
    else LASR_BN_RED_ONE(TT, false, false);                                                                        \
  } while (0)
  if (dtype == LASR_F32)
    LASR_BN_RED_LAUNCH(float);
  else if (dtype == LASR_BF16)
    LASR_BN_RED_LAUNCH(__nv_bfloat16);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_PDL(le);
  return LASR_OK;
}

int lasr_bn_bwd_coef(const double* totals, int C, int count, int slot_gx, const float* gamma, const float* mean,
                     const float* invstd, float* dgamma, float* dbeta, float* coef, lasr_stream_t stream) {
  if (C <= 0 || count <= 0 || slot_gx < 1 || slot_gx > 2 || totals == nullptr) return LASR_ERR_BAD_SHAPE;
  bn_bwd_coef_kernel<<<cdiv(C, 128), 128, 0, stream>>>(totals, C, static_cast<double>(count), slot_gx, gamma, mean,
                                                        invstd, dgamma, dbeta, coef);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_bn_act_bwd_apply(const void* dout, const void* out, const void* y, const void* r, const float* gate,
                          const float* extra, const double* totals, const float* coef1, const lasr_bn_bwd_t* bn1,
                          const lasr_bn_bwd_t* bn2, int count, const int32_t* lengths, int T, void* dy, void* dr,
                          int M, int C, int act, const lasr_dropout_t* drop, int dtype, const uint8_t* relu_bits,
                          lasr_stream_t stream) {
  if (M <= 0 || C <= 0 || (C % 8) || C > 2048 || T <= 0 || count <= 0) return LASR_ERR_BAD_SHAPE;
  if (act == LASR_ACT_RELU && out == nullptr && relu_bits == nullptr) return LASR_ERR_BAD_SHAPE;
  if (relu_bits != nullptr && (M % T) != 0) return LASR_ERR_BAD_SHAPE;
  DropArgs da;
  if (int rc = drop_args(drop, da)) return rc;
  const uint8_t* dmask = da.mode ? da.mask : nullptr;
  if ((r != nullptr) != (dr != nullptr) || (r != nullptr) != (bn2 != nullptr)) return LASR_ERR_BAD_SHAPE;
  if ((gate != nullptr) != (extra != nullptr)) return LASR_ERR_BAD_SHAPE;
  if (coef1 == nullptr && (bn1 == nullptr || totals == nullptr)) return LASR_ERR_BAD_SHAPE;
  if (r != nullptr && totals == nullptr) return LASR_ERR_BAD_SHAPE;
  const int CV = C / 8;
  const long long total = static_cast<long long>(M) * CV;
  BnBwdSide s1{}, s2{};
  if (bn1) s1 = BnBwdSide{bn1->gamma, bn1->mean, bn1->invstd, bn1->dgamma, bn1->dbeta};
  if (bn2) s2 = BnBwdSide{bn2->gamma, bn2->mean, bn2->invstd, bn2->dgamma, bn2->dbeta};
  if (dtype == LASR_F32)
    return bn_bwd_launch<float>(dout, out, y, r, gate, extra, totals, coef1, s1, s2, count, lengths, T, dy, dr, total,
                                CV, C, act, dmask, da.scale, relu_bits, stream);
  if (dtype == LASR_BF16)
    return bn_bwd_launch<__nv_bfloat16>(dout, out, y, r, gate, extra, totals, coef1, s1, s2, count, lengths, T, dy,
                                        dr, total, CV, C, act, dmask, da.scale, relu_bits, stream);
  return LASR_ERR_BAD_DTYPE;
}

}  // extern "C"
