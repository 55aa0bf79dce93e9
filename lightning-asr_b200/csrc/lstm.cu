// Context BiLSTM (256 -> 2 x 40, one layer) recurrence, forward and backward, variable lengths, no host sync.
// Replaces models/QuartNetContext.py:157,171-173,186-199: `length = (T' * percents).int().cpu()` ->
// pack_padded_sequence(enforce_sorted=False) -> nn.LSTM(256, 40, bidirectional) -> pad_packed_sequence.
//
// Split of the work:
//   * the input projections of all frames and both directions, pre[n, t, d*160 + g*40 + u] = W_ih x_t + b_ih + b_hh,
//     are ONE pointwise-conv GEMM (256 -> 320, gemm_tc.cu / gemm_simt.cu) issued by the caller; likewise dW_ih and dx
//     in the backward are the regular weight-gradient / data-gradient GEMMs over dpre;
//   * this file is the sequential part.  One CTA per (utterance, direction), 160 threads = one per gate row
//     (PyTorch order i, f, g, o).  The 160 x 40 recurrent matrix lives in registers (40 per thread), h_{t-1} in shared
//     memory (broadcast reads), the cell state in the registers of the first 40 threads.  Two barriers per frame.
//     The forward direction walks t = 0 .. len-1, the reverse direction t = len-1 .. 0; frames >= len produce zeros,
//     exactly what pad_packed_sequence returns.  `pre` rows are prefetched 8 frames ahead through registers.
//   * backward: same CTA shape walking the frames in the opposite order.  Thread u < 40 turns dh into the four
//     pre-activation gate gradients of unit u; then every thread j adds dgate_j * h_prev[k] to its 40 register
//     accumulators of dW_hh[j, :] (one atomicAdd pass at the end) and the threads, regrouped as (unit k, gate q),
//     reduce dh_prev[k] = sum_j W_hh[j, k] dgate_j with two shuffles.
#include "common.cuh"

#include <cstdlib>

namespace lasr {

constexpr int LS_H = 40;          // hidden units per direction
constexpr int LS_G = 4 * LS_H;    // gate rows per direction
constexpr int LS_PF = 8;          // frames of prefetch distance

// full-accuracy libm versions: the recurrence is latency-bound by its two barriers per frame, not by these, and
// whole-network gradient parity (train-mode BatchNorm amplifies rounding ~20x) wants every ulp
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }

// pre [N, T, 2*160] (T_), whh [2, 160, 40] fp32, lengths [N] (nullable: all T)
// out  [N, T, 80]  (T_): h of (direction d, unit u) at column d*40 + u; zero for t >= len
// gates [N, T, 2, 40] float4 (i, f, g, o after the non-linearities), cells [N, T, 2, 40] fp32: kept for the backward
template <typename T_>
__global__ void __launch_bounds__(LS_G)
bilstm_fwd_kernel(const T_* __restrict__ pre, const float* __restrict__ whh, const int32_t* __restrict__ lengths,
                  T_* __restrict__ out, float4* __restrict__ gates, float* __restrict__ cells, int T) {
  const int n = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  __shared__ float h_s[2][LS_H];
  __shared__ float g_s[LS_G];
  float w[LS_H];
#pragma unroll
  for (int k = 0; k < LS_H; ++k) w[k] = whh[(static_cast<size_t>(d) * LS_G + j) * LS_H + k];
  int len = lengths != nullptr ? lengths[n] : T;
  len = max(0, min(len, T));
  if (j < LS_H) h_s[0][j] = 0.f;
  float c = 0.f;
  const int gate_type = j / LS_H;
  const T_* pre_n = pre + static_cast<size_t>(n) * T * (2 * LS_G) + d * LS_G + j;
  auto frame = [&](int s) { return d == 0 ? s : len - 1 - s; };
  float pf[LS_PF];
#pragma unroll
  for (int i = 0; i < LS_PF; ++i) pf[i] = (i < len) ? to_f32<T_>(pre_n[static_cast<size_t>(frame(i)) * (2 * LS_G)]) : 0.f;
  __syncthreads();
  int cur = 0;
  for (int s0 = 0; s0 < len; s0 += LS_PF) {
#pragma unroll
    for (int i = 0; i < LS_PF; ++i) {
      const int s = s0 + i;
      if (s >= len) break;
      float acc0 = pf[i], acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      // refill this slot for frame s + LS_PF: the load has LS_PF frames of recurrence to land
      if (s + LS_PF < len) pf[i] = to_f32<T_>(pre_n[static_cast<size_t>(frame(s + LS_PF)) * (2 * LS_G)]);
      const float* hp = h_s[cur];
#pragma unroll
      for (int k = 0; k < LS_H; k += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hp + k);
        acc0 = fmaf(w[k], hv.x, acc0);
        acc1 = fmaf(w[k + 1], hv.y, acc1);
        acc2 = fmaf(w[k + 2], hv.z, acc2);
        acc3 = fmaf(w[k + 3], hv.w, acc3);
      }
      const float a = (acc0 + acc1) + (acc2 + acc3);
      g_s[j] = gate_type == 2 ? tanhf_(a) : sigmoidf_(a);
      __syncthreads();
      if (j < LS_H) {
        const float gi = g_s[j], gf = g_s[LS_H + j], gg = g_s[2 * LS_H + j], go = g_s[3 * LS_H + j];
        c = fmaf(gf, c, gi * gg);
        const float h = go * tanhf_(c);
        h_s[cur ^ 1][j] = h;
        const int t = frame(s);
        const size_t row = static_cast<size_t>(n) * T + t;
        out[row * (2 * LS_H) + d * LS_H + j] = from_f32<T_>(h);
        gates[(row * 2 + d) * LS_H + j] = make_float4(gi, gf, gg, go);
        cells[(row * 2 + d) * LS_H + j] = c;
      }
      __syncthreads();
      cur ^= 1;
    }
  }
  // pad_packed_sequence: zeros after the utterance's last frame
  if (j < LS_H)
    for (int t = len; t < T; ++t) out[(static_cast<size_t>(n) * T + t) * (2 * LS_H) + d * LS_H + j] = from_f32<T_>(0.f);
}

// dout [N, T, 80] (T_) gradient of `out`; out / gates / cells from the forward
// dpre [N, T, 320] (T_) out: gradient of the pre-activations (zero for t >= len)
// dwhh [2, 160, 40] fp32: ACCUMULATED (atomicAdd)
template <typename T_>
__global__ void __launch_bounds__(LS_G)
bilstm_bwd_kernel(const T_* __restrict__ dout, const T_* __restrict__ out, const float4* __restrict__ gates,
                  const float* __restrict__ cells, const float* __restrict__ whh, const int32_t* __restrict__ lengths,
                  T_* __restrict__ dpre, float* __restrict__ dwhh, int T) {
  const int n = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  __shared__ float dg_s[LS_G];
  __shared__ float hprev_s[LS_H];
  __shared__ float dhrec_s[LS_H];
  // regrouped view for the dh_prev reduction: thread = (unit k, gate q); it owns W_hh[q*40 + jj, k], jj < 40
  const int k_own = j >> 2, q_own = j & 3;
  float wt[LS_H];
#pragma unroll
  for (int jj = 0; jj < LS_H; ++jj)
    wt[jj] = whh[(static_cast<size_t>(d) * LS_G + q_own * LS_H + jj) * LS_H + k_own];
  float acc[LS_H];
#pragma unroll
  for (int k = 0; k < LS_H; ++k) acc[k] = 0.f;
  int len = lengths != nullptr ? lengths[n] : T;
  len = max(0, min(len, T));
  if (j < LS_H) dhrec_s[j] = 0.f;
  float dc_carry = 0.f;
  __syncthreads();
  // recurrence order of direction d was: forward t = 0..len-1, reverse t = len-1..0.  Walk it backwards; position s
  // of the recurrence is frame t(s), its predecessor (whose h / c fed it) is position s-1.
  auto frame = [&](int s) { return d == 0 ? s : len - 1 - s; };
  struct Ld {
    float4 g4;
    float ct, cp, hp, dh;
  };
  auto load = [&](int s) {
    Ld v;
    const size_t row = static_cast<size_t>(n) * T + frame(s);
    v.g4 = gates[(row * 2 + d) * LS_H + j];
    v.ct = cells[(row * 2 + d) * LS_H + j];
    v.dh = to_f32<T_>(dout[row * (2 * LS_H) + d * LS_H + j]);
    v.cp = 0.f;
    v.hp = 0.f;
    if (s > 0) {
      const size_t rowp = static_cast<size_t>(n) * T + frame(s - 1);
      v.cp = cells[(rowp * 2 + d) * LS_H + j];
      v.hp = to_f32<T_>(out[rowp * (2 * LS_H) + d * LS_H + j]);
    }
    return v;
  };
  constexpr int PF = 4;  // positions of prefetch distance (unit threads only)
  Ld q[PF];
  if (j < LS_H) {
#pragma unroll
    for (int i = 0; i < PF; ++i)
      if (len - 1 - i >= 0) q[i] = load(len - 1 - i);
  }
  for (int s0 = len - 1; s0 >= 0; s0 -= PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int s = s0 - i;
      if (s < 0) break;
      if (j < LS_H) {
        const Ld v = q[i];
        if (s - PF >= 0) q[i] = load(s - PF);
        const float4 g4 = v.g4;
        const float dh = v.dh + dhrec_s[j];
        const float tc = tanhf_(v.ct);
        const float d_o = dh * tc;
        const float dc = fmaf(dh * g4.w, 1.f - tc * tc, dc_carry);
        const float d_i = dc * g4.z, d_g = dc * g4.x, d_f = dc * v.cp;
        dc_carry = dc * g4.y;
        const float pi = d_i * g4.x * (1.f - g4.x);
        const float pf_ = d_f * g4.y * (1.f - g4.y);
        const float pg = d_g * (1.f - g4.z * g4.z);
        const float po = d_o * g4.w * (1.f - g4.w);
        dg_s[j] = pi;
        dg_s[LS_H + j] = pf_;
        dg_s[2 * LS_H + j] = pg;
        dg_s[3 * LS_H + j] = po;
        hprev_s[j] = v.hp;
        T_* dp = dpre + (static_cast<size_t>(n) * T + frame(s)) * (2 * LS_G) + d * LS_G + j;
        dp[0] = from_f32<T_>(pi);
        dp[LS_H] = from_f32<T_>(pf_);
        dp[2 * LS_H] = from_f32<T_>(pg);
        dp[3 * LS_H] = from_f32<T_>(po);
      }
      __syncthreads();
      {
        const float dgj = dg_s[j];
#pragma unroll
        for (int k = 0; k < LS_H; k += 4) {
          const float4 hv = *reinterpret_cast<const float4*>(hprev_s + k);
          acc[k] = fmaf(dgj, hv.x, acc[k]);
          acc[k + 1] = fmaf(dgj, hv.y, acc[k + 1]);
          acc[k + 2] = fmaf(dgj, hv.z, acc[k + 2]);
          acc[k + 3] = fmaf(dgj, hv.w, acc[k + 3]);
        }
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        const float* dq = dg_s + q_own * LS_H;
#pragma unroll
        for (int jj = 0; jj < LS_H; jj += 4) {
          const float4 dv = *reinterpret_cast<const float4*>(dq + jj);
          p0 = fmaf(wt[jj], dv.x, p0);
          p1 = fmaf(wt[jj + 1], dv.y, p1);
          p2 = fmaf(wt[jj + 2], dv.z, p2);
          p3 = fmaf(wt[jj + 3], dv.w, p3);
        }
        float part = (p0 + p1) + (p2 + p3);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (q_own == 0) dhrec_s[k_own] = part;  // its readers (the unit threads) finished before the barrier above
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int k = 0; k < LS_H; ++k) atomicAdd(dwhh + (static_cast<size_t>(d) * LS_G + j) * LS_H + k, acc[k]);
  // frames past the end: no gradient
  for (int t = len; t < T; ++t) dpre[(static_cast<size_t>(n) * T + t) * (2 * LS_G) + d * LS_G + j] = from_f32<T_>(0.f);
}

// ------------------------------------------------------------------------------------------------
// Second generation (round 2): ONE barrier per frame.  The kernels above are paced by two CTA barriers per frame and, in
// the backward, by a phase that only 40 of the 160 threads execute (the tanh / gate-gradient chain) while the others
// wait: 0.54 / 1.05 us per frame.  Here thread j = 4 u + q owns gate q of unit u, so the four gates of a unit sit in
// four adjacent lanes:
//   forward   every thread computes its gate's pre-activation (the same 40-term dot product in the same order), the
//             quad exchanges the four activations with shuffles and all four lanes update the unit's cell state
//             redundantly (no g_s round trip, no second barrier); lane q = 0 publishes h for the next frame;
//   backward  all four lanes of a quad run the unit's gate-gradient chain redundantly and keep their own gate's
//             gradient; after the frame's one barrier every thread adds its row's outer-product term to dW_hh and the
//             regrouped reduction of dh_prev (thread = (unit k, gate q), two xor-shuffles) leaves dh_prev[k] in all
//             four lanes of unit k's quad -- exactly where the next frame's chain needs it, in registers.
// The exchange buffers are double-buffered, which is what makes one barrier enough.  Same operations in the same order
// per value: outputs, saved gates / cells and all gradients are bit-identical to the kernels above
// (tests/test_kernels_gpu.py::test_bilstm_one_barrier_kernels_are_bit_identical); LASR_LSTM_V1=1 selects those, 2 these.
// ------------------------------------------------------------------------------------------------
template <typename T_>
__global__ void __launch_bounds__(LS_G)
bilstm_fwd2_kernel(const T_* __restrict__ pre, const float* __restrict__ whh, const int32_t* __restrict__ lengths,
                   T_* __restrict__ out, float4* __restrict__ gates, float* __restrict__ cells, int T) {
  const int n = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  const int u = j >> 2, q = j & 3;
  const int row_j = q * LS_H + u;  // this thread's gate row in PyTorch's (i, f, g, o) order
  __shared__ __align__(16) float h_s[2][LS_H];
  float w[LS_H];
#pragma unroll
  for (int k = 0; k < LS_H; ++k) w[k] = whh[(static_cast<size_t>(d) * LS_G + row_j) * LS_H + k];
  int len = lengths != nullptr ? lengths[n] : T;
  len = max(0, min(len, T));
  if (j < LS_H) h_s[0][j] = 0.f;
  float c = 0.f;
  const T_* pre_n = pre + static_cast<size_t>(n) * T * (2 * LS_G) + d * LS_G + row_j;
  auto frame = [&](int s) { return d == 0 ? s : len - 1 - s; };
  float pf[LS_PF];
#pragma unroll
  for (int i = 0; i < LS_PF; ++i) pf[i] = (i < len) ? to_f32<T_>(pre_n[static_cast<size_t>(frame(i)) * (2 * LS_G)]) : 0.f;
  __syncthreads();
  const int quad = (j & 31) & ~3;  // first lane of this unit's quad
  int cur = 0;
  for (int s0 = 0; s0 < len; s0 += LS_PF) {
#pragma unroll
    for (int i = 0; i < LS_PF; ++i) {
      const int s = s0 + i;
      if (s >= len) break;
      float acc0 = pf[i], acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      if (s + LS_PF < len) pf[i] = to_f32<T_>(pre_n[static_cast<size_t>(frame(s + LS_PF)) * (2 * LS_G)]);
      const float* hp = h_s[cur];
#pragma unroll
      for (int k = 0; k < LS_H; k += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hp + k);
        acc0 = fmaf(w[k], hv.x, acc0);
        acc1 = fmaf(w[k + 1], hv.y, acc1);
        acc2 = fmaf(w[k + 2], hv.z, acc2);
        acc3 = fmaf(w[k + 3], hv.w, acc3);
      }
      const float a = (acc0 + acc1) + (acc2 + acc3);
      const float act = q == 2 ? tanhf_(a) : sigmoidf_(a);
      const float gi = __shfl_sync(0xffffffffu, act, quad);
      const float gf = __shfl_sync(0xffffffffu, act, quad + 1);
      const float gg = __shfl_sync(0xffffffffu, act, quad + 2);
      const float go = __shfl_sync(0xffffffffu, act, quad + 3);
      c = fmaf(gf, c, gi * gg);
      const float h = go * tanhf_(c);
      if (q == 0) {
        h_s[cur ^ 1][u] = h;
        const int t = frame(s);
        const size_t row = static_cast<size_t>(n) * T + t;
        out[row * (2 * LS_H) + d * LS_H + u] = from_f32<T_>(h);
        gates[(row * 2 + d) * LS_H + u] = make_float4(gi, gf, gg, go);
        cells[(row * 2 + d) * LS_H + u] = c;
      }
      __syncthreads();
      cur ^= 1;
    }
  }
  // pad_packed_sequence: zeros after the utterance's last frame
  if (j < LS_H)
    for (int t = len; t < T; ++t) out[(static_cast<size_t>(n) * T + t) * (2 * LS_H) + d * LS_H + j] = from_f32<T_>(0.f);
}

template <typename T_>
__global__ void __launch_bounds__(LS_G)
bilstm_bwd2_kernel(const T_* __restrict__ dout, const T_* __restrict__ out, const float4* __restrict__ gates,
                   const float* __restrict__ cells, const float* __restrict__ whh, const int32_t* __restrict__ lengths,
                   T_* __restrict__ dpre, float* __restrict__ dwhh, int T) {
  const int n = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  const int u = j >> 2, q = j & 3;     // phase 1 / dW view: gate q of unit u (row q*40 + u)
  const int row_j = q * LS_H + u;
  __shared__ __align__(16) float dg_s[2][LS_G];
  __shared__ __align__(16) float hprev_s[2][LS_H];
  // dh_prev view: thread = (unit k = u, gate q); it owns W_hh[q*40 + jj, u], jj < 40
  float wt[LS_H];
#pragma unroll
  for (int jj = 0; jj < LS_H; ++jj) wt[jj] = whh[(static_cast<size_t>(d) * LS_G + q * LS_H + jj) * LS_H + u];
  float acc[LS_H];
#pragma unroll
  for (int k = 0; k < LS_H; ++k) acc[k] = 0.f;
  int len = lengths != nullptr ? lengths[n] : T;
  len = max(0, min(len, T));
  float dc_carry = 0.f, dh_rec = 0.f;
  auto frame = [&](int s) { return d == 0 ? s : len - 1 - s; };
  struct Ld {
    float4 g4;
    float ct, cp, hp, dh;
  };
  auto load = [&](int s) {  // the unit's saved state: the four lanes of a quad read the same addresses
    Ld v;
    const size_t row = static_cast<size_t>(n) * T + frame(s);
    v.g4 = gates[(row * 2 + d) * LS_H + u];
    v.ct = cells[(row * 2 + d) * LS_H + u];
    v.dh = to_f32<T_>(dout[row * (2 * LS_H) + d * LS_H + u]);
    v.cp = 0.f;
    v.hp = 0.f;
    if (s > 0) {
      const size_t rowp = static_cast<size_t>(n) * T + frame(s - 1);
      v.cp = cells[(rowp * 2 + d) * LS_H + u];
      v.hp = to_f32<T_>(out[rowp * (2 * LS_H) + d * LS_H + u]);
    }
    return v;
  };
  constexpr int PF = 4;  // positions of prefetch distance
  Ld ring[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i)
    if (len - 1 - i >= 0) ring[i] = load(len - 1 - i);
  int cur = 0;
  for (int s0 = len - 1; s0 >= 0; s0 -= PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int s = s0 - i;
      if (s < 0) break;
      const Ld v = ring[i];
      if (s - PF >= 0) ring[i] = load(s - PF);
      const float4 g4 = v.g4;
      const float dh = v.dh + dh_rec;
      const float tc = tanhf_(v.ct);
      const float d_o = dh * tc;
      const float dc = fmaf(dh * g4.w, 1.f - tc * tc, dc_carry);
      const float d_i = dc * g4.z, d_g = dc * g4.x, d_f = dc * v.cp;
      dc_carry = dc * g4.y;
      const float pi = d_i * g4.x * (1.f - g4.x);
      const float pf_ = d_f * g4.y * (1.f - g4.y);
      const float pg = d_g * (1.f - g4.z * g4.z);
      const float po = d_o * g4.w * (1.f - g4.w);
      const float dgj = q == 0 ? pi : (q == 1 ? pf_ : (q == 2 ? pg : po));  // this thread's row
      dg_s[cur][row_j] = dgj;
      if (q == 0) hprev_s[cur][u] = v.hp;
      dpre[(static_cast<size_t>(n) * T + frame(s)) * (2 * LS_G) + d * LS_G + row_j] = from_f32<T_>(dgj);
      __syncthreads();
      {
        const float* hp = hprev_s[cur];
#pragma unroll
        for (int k = 0; k < LS_H; k += 4) {
          const float4 hv = *reinterpret_cast<const float4*>(hp + k);
          acc[k] = fmaf(dgj, hv.x, acc[k]);
          acc[k + 1] = fmaf(dgj, hv.y, acc[k + 1]);
          acc[k + 2] = fmaf(dgj, hv.z, acc[k + 2]);
          acc[k + 3] = fmaf(dgj, hv.w, acc[k + 3]);
        }
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        const float* dq = dg_s[cur] + q * LS_H;
#pragma unroll
        for (int jj = 0; jj < LS_H; jj += 4) {
          const float4 dv = *reinterpret_cast<const float4*>(dq + jj);
          p0 = fmaf(wt[jj], dv.x, p0);
          p1 = fmaf(wt[jj + 1], dv.y, p1);
          p2 = fmaf(wt[jj + 2], dv.z, p2);
          p3 = fmaf(wt[jj + 3], dv.w, p3);
        }
        float part = (p0 + p1) + (p2 + p3);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        dh_rec = part;  // dh_prev of unit u, in all four lanes of its quad
      }
      cur ^= 1;  // the next frame fills the other buffers: nobody can still be reading them (one barrier behind)
    }
  }
#pragma unroll
  for (int k = 0; k < LS_H; ++k) atomicAdd(dwhh + (static_cast<size_t>(d) * LS_G + row_j) * LS_H + k, acc[k]);
  // frames past the end: no gradient
  for (int t = len; t < T; ++t) dpre[(static_cast<size_t>(n) * T + t) * (2 * LS_G) + d * LS_G + j] = from_f32<T_>(0.f);
}


// ------------------------------------------------------------------------------------------------
// Third generation: the same thread layout and single barrier as above, with the per-frame instruction stream of a warp
// cut in half.  A CTA is 5 warps on 4 schedulers and every frame is one dependent chain, so the frame time is the length
// of a warp's instruction stream (measured: generation 1 and 2 both 0.74 / 1.15 us per frame at 185 / 245 instructions),
// not the barrier count.  What went:
//   * the dot products run as packed fma.rn.f32x2 (FFMA2: two IEEE fp32 FMAs per instruction, same values as the scalar
//     chains: 20 instead of 40 issue slots per 40-term product);
//   * FAST (the bf16 path): sigmoid = rcp(1 + ex2(-x log2 e)), tanh(x) = 2 sigmoid(2x) - 1, branch-free for all four
//     gate types (5 instructions; abs. error ~2e-7, three decimal orders below the bf16 rounding of `pre` and `out`).
//     libm's tanhf / IEEE division cost ~60 instructions per frame and diverge inside a quad.  fp32 keeps libm;
//   * no global load is consumed inside the recurrence: the operands of the next 16 positions (forward: the `pre` rows;
//     backward: gates, cells, dout and the previous h, ~1 KB per position) are copied to shared memory with cp.async
//     while the current 16 are walked.  Loads issued several frames ahead through registers share scoreboards, so
//     every frame waited for the YOUNGEST load's full latency (probe with L1-resident rows: fwd 260 -> 188 us, bwd 659 ->
//     258 us); rows are addressed with 32-bit element indices advanced by +-1 per frame;
//   * backward: the cell state of the previous position is the next row of the staged chunk, each lane derives only ITS
//     gate's gradient -- dgate = A * (B * D) with B, D selected per lane off the critical chain, A = dc (i, f, g) or
//     dh tanh(c) (o) -- instead of all four; the operands of position p + 1 are fetched from the stage right behind
//     position p's barrier.  Chunks are walked in pairs of positions (static exchange-buffer parity); an odd tail runs
//     one virtual frame that stores nothing.
// Forward pre-activations are bit-identical to the kernels above; the backward re-associates a few products (1 ulp).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
// sc = 1: sigmoid(x); sc = 2: tanh(x).  km = -sc log2(e), kb = 1 - sc
__device__ __forceinline__ float act_fast(float x, float km, float sc, float kb) {
  return fmaf(rcp_approx(1.f + ex2_approx(x * km)), sc, kb);
}
__device__ __forceinline__ float tanh_fast(float x) { return act_fast(x, -2.f * kLog2e, 2.f, -1.f); }

// branch-free select on a lane constant (the compiler turns nested ?: on per-lane values into divergent branches)
__device__ __forceinline__ float selp(float a, float b, int take_a) {
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"(take_a));
  return r;
}

constexpr int LS_CH = 16;  // positions of the recurrence per staged chunk (even: exchange buffers alternate by parity)
__device__ __forceinline__ void ls_cp16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(sdst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void ls_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ls_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T_, bool FAST>
__global__ void __launch_bounds__(LS_G)
bilstm_fwd3_kernel(const T_* __restrict__ pre, const float* __restrict__ whh, const int32_t* __restrict__ lengths,
                   T_* __restrict__ out, float4* __restrict__ gates, float* __restrict__ cells, int T) {
  static_assert(LS_CH % 2 == 0, "the exchange buffer index is the unrolled position's parity");
  const int n = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  const int u = j >> 2, q = j & 3;
  const int row_j = q * LS_H + u;  // this thread's gate row in PyTorch's (i, f, g, o) order
  __shared__ __align__(16) float h_s[2][LS_H];
  __shared__ __align__(16) T_ pre_s[2][LS_CH][LS_G];  // two chunks of this direction's input projections
  float2 w2[LS_H / 2];
#pragma unroll
  for (int k = 0; k < LS_H / 2; ++k)
    w2[k] = *reinterpret_cast<const float2*>(whh + (static_cast<size_t>(d) * LS_G + row_j) * LS_H + 2 * k);
  int len = lengths != nullptr ? lengths[n] : T;
  len = max(0, min(len, T));
  if (j < LS_H) h_s[0][j] = 0.f;
  float c = 0.f;
  const float sc = q == 2 ? 2.f : 1.f, km = -sc * kLog2e, kb = 1.f - sc;
  const int dir = d == 0 ? 1 : -1;
  const int row0 = n * T + (d == 0 ? 0 : len - 1);  // row of the recurrence's first position; the host checked the range
  // chunk ch of the walk = positions ch*LS_CH ..: their `pre` rows (160 values of this direction each) go to shared
  // memory with cp.async while the previous chunk is being walked -- a global load consumed inside the recurrence costs
  // its full latency every frame (the loads of several frames share a scoreboard: measured 260 vs 188 us)
  constexpr int V = LS_G * sizeof(T_) / 16;  // 16-byte vectors per position
  auto issue = [&](int ch) {
    const int cnt = min(LS_CH, len - ch * LS_CH);
    char* dst = reinterpret_cast<char*>(&pre_s[ch & 1][0][0]);
    for (int e = j; e < cnt * V; e += LS_G) {
      const int p = e / V, v = e - p * V;
      const int row = row0 + dir * (ch * LS_CH + p);
      ls_cp16(dst + e * 16, reinterpret_cast<const char*>(pre + (static_cast<size_t>(row) * 2 + d) * LS_G) + v * 16);
    }
  };
  const int nch = (len + LS_CH - 1) / LS_CH;
  if (nch > 0) issue(0);
  ls_cp_commit();
  const int s80 = dir * (2 * LS_H);
  int i80 = row0 * (2 * LS_H) + d * LS_H + u;  // element index of the current row in out / gates / cells
  const int quad = (j & 31) & ~3;              // first lane of this unit's quad
  for (int ch = 0; ch < nch; ++ch) {
    ls_cp_wait_all();
    __syncthreads();  // chunk ch has landed for everybody, and everybody is done with the stage chunk ch + 1 goes to
    if (ch + 1 < nch) issue(ch + 1);
    ls_cp_commit();
    const int cnt = min(LS_CH, len - ch * LS_CH);
    const T_* ps = &pre_s[ch & 1][0][row_j];
#pragma unroll 1
    for (int p0 = 0; p0 < cnt; p0 += 2) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool valid = p0 + i < cnt;  // an odd chunk ends with a virtual frame: nothing is stored
        float2 a01 = make_float2(to_f32<T_>(ps[(p0 + i) * LS_G]), 0.f), a23 = make_float2(0.f, 0.f);
        const float* hp = h_s[i];
#pragma unroll
        for (int k = 0; k < LS_H; k += 4) {
          const float4 hv = *reinterpret_cast<const float4*>(hp + k);
          a01 = ffma2(w2[k / 2], make_float2(hv.x, hv.y), a01);
          a23 = ffma2(w2[k / 2 + 1], make_float2(hv.z, hv.w), a23);
        }
        const float a = (a01.x + a01.y) + (a23.x + a23.y);
        float act;
        if (FAST)
          act = act_fast(a, km, sc, kb);
        else
          act = q == 2 ? tanhf_(a) : sigmoidf_(a);
        const float gi = __shfl_sync(0xffffffffu, act, quad);
        const float gf = __shfl_sync(0xffffffffu, act, quad + 1);
        const float gg = __shfl_sync(0xffffffffu, act, quad + 2);
        const float go = __shfl_sync(0xffffffffu, act, quad + 3);
        c = fmaf(gf, c, gi * gg);
        if (q == 0 && valid) {
          const float h = go * (FAST ? tanh_fast(c) : tanhf_(c));
          h_s[i ^ 1][u] = h;
          out[i80] = from_f32<T_>(h);
          gates[i80] = make_float4(gi, gf, gg, go);
          cells[i80] = c;
        }
        i80 += s80;
        __syncthreads();
      }
    }
  }
  // pad_packed_sequence: zeros after the utterance's last frame
  if (j < LS_H)
    for (int t = len; t < T; ++t) out[(static_cast<size_t>(n) * T + t) * (2 * LS_H) + d * LS_H + j] = from_f32<T_>(0.f);
}

// one staged chunk of the backward walk: the saved state of LS_CH positions of one (utterance, direction)
template <typename T_>
struct LstmBwdStage {
  float4 g4[LS_CH][LS_H];     // gates after the non-linearities
  float ct[LS_CH + 1][LS_H];  // cell states; row p + 1 = the position before p (its c feeds the forget gate's gradient)
  T_ dh[LS_CH][LS_H];         // upstream gradient of h
  T_ hp[LS_CH][LS_H];         // h of the position before (the recurrent input)
};

template <typename T_, bool FAST>
__global__ void __launch_bounds__(LS_G, 1)
bilstm_bwd3_kernel(const T_* __restrict__ dout, const T_* __restrict__ out, const float4* __restrict__ gates,
                   const float* __restrict__ cells, const float* __restrict__ whh, const int32_t* __restrict__ lengths,
                   T_* __restrict__ dpre, float* __restrict__ dwhh, int T) {
  const int n = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  const int u = j >> 2, q = j & 3;  // gate q of unit u (row q*40 + u); for the dh_prev reduction: (unit k = u, gate q)
  const int row_j = q * LS_H + u;
  __shared__ __align__(16) LstmBwdStage<T_> st_s[2];
  __shared__ __align__(16) float dg_s[2][LS_G];
  __shared__ __align__(16) float hprev_s[2][LS_H];
  float2 wt2[LS_H / 2];  // W_hh[q*40 + jj, u], jj < 40
#pragma unroll
  for (int jj = 0; jj < LS_H / 2; ++jj) {
    wt2[jj].x = whh[(static_cast<size_t>(d) * LS_G + q * LS_H + 2 * jj) * LS_H + u];
    wt2[jj].y = whh[(static_cast<size_t>(d) * LS_G + q * LS_H + 2 * jj + 1) * LS_H + u];
  }
  float2 acc2[LS_H / 2];
#pragma unroll
  for (int k = 0; k < LS_H / 2; ++k) acc2[k] = make_float2(0.f, 0.f);
  int len = lengths != nullptr ? lengths[n] : T;
  len = max(0, min(len, T));
  float dc_carry = 0.f, dh_rec = 0.f;
  const int dir = d == 0 ? 1 : -1;
  const int base = n * T + (d == 0 ? 0 : len - 1);  // row of recurrence position s: base + dir * s
  // chunk ch of the walk = positions s_hi = len - 1 - ch*LS_CH downwards.  Per position: 40 + 10 + VD + VD 16-byte vectors
  // (gates, cells, dout, out of the position before), plus one more row of cells behind the chunk's last position
  constexpr int VD = LS_H * sizeof(T_) / 16, OPS = LS_H + LS_H / 4 + 2 * VD;
  auto issue = [&](int ch) {
    LstmBwdStage<T_>& S = st_s[ch & 1];
    const int s_hi = len - 1 - ch * LS_CH;
    const int cnt = min(LS_CH, s_hi + 1);
    for (int e = j; e < cnt * OPS; e += LS_G) {
      const int p = e / OPS, o = e - p * OPS;
      const int s = s_hi - p;
      const size_t r80 = static_cast<size_t>(base + dir * s) * (2 * LS_H) + d * LS_H;  // the (row, direction) block
      if (o < LS_H) {
        ls_cp16(&S.g4[p][o], gates + r80 + o);
      } else if (o < LS_H + LS_H / 4) {
        ls_cp16(&S.ct[p][(o - LS_H) * 4], cells + r80 + (o - LS_H) * 4);
      } else if (o < LS_H + LS_H / 4 + VD) {
        const int v = o - (LS_H + LS_H / 4);
        ls_cp16(reinterpret_cast<char*>(&S.dh[p][0]) + v * 16, reinterpret_cast<const char*>(dout + r80) + v * 16);
      } else if (s > 0) {
        const int v = o - (LS_H + LS_H / 4 + VD);
        const size_t r80p = static_cast<size_t>(base + dir * (s - 1)) * (2 * LS_H) + d * LS_H;
        ls_cp16(reinterpret_cast<char*>(&S.hp[p][0]) + v * 16, reinterpret_cast<const char*>(out + r80p) + v * 16);
      }
    }
    if (j < LS_H / 4 && s_hi - cnt >= 0) {
      const size_t r80 = static_cast<size_t>(base + dir * (s_hi - cnt)) * (2 * LS_H) + d * LS_H;
      ls_cp16(&S.ct[cnt][j * 4], cells + r80 + j * 4);
    }
  };
  const int nch = (len + LS_CH - 1) / LS_CH;
  if (nch > 0) issue(0);
  ls_cp_commit();
  const int spre = -dir * (2 * LS_G);
  int ipre = (base + dir * (len - 1)) * (2 * LS_G) + d * LS_G + row_j;  // the row being processed, in dpre
  const int is_q0 = q == 0, is_q1 = q == 1, is_q2 = q == 2, is_q3 = q == 3;
  for (int ch = 0; ch < nch; ++ch) {
    ls_cp_wait_all();
    __syncthreads();  // chunk ch has landed for everybody, and everybody is done with the stage chunk ch + 1 goes to
    if (ch + 1 < nch) issue(ch + 1);
    ls_cp_commit();
    const LstmBwdStage<T_>& S = st_s[ch & 1];
    const int s_hi = len - 1 - ch * LS_CH;
    const int cnt = min(LS_CH, s_hi + 1);
    // operands of the next position, fetched from the stage one frame ahead (behind the previous frame's barrier)
    float4 g4 = S.g4[0][u];
    float ct = S.ct[0][u], cpn = S.ct[1][u], dhv = to_f32<T_>(S.dh[0][u]), hpv = to_f32<T_>(S.hp[0][u]);
#pragma unroll 1
    for (int p0 = 0; p0 < cnt; p0 += 2) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int p = p0 + i, s = s_hi - p;
        const bool valid = p < cnt;  // an odd last chunk ends with a virtual frame (s = -1): zero gate gradient, no store
        const float cp = s > 0 ? cpn : 0.f;  // c of position s - 1
        // off the recurrent chain: this lane's gate value S, its partner factor B and the activation derivative D
        const float Sg = selp(g4.x, selp(g4.y, selp(g4.z, g4.w, is_q2), is_q1), is_q0);
        const float B = selp(g4.z, selp(cp, selp(g4.x, 1.f, is_q2), is_q1), is_q0);
        const float D = fmaf(-Sg, Sg, selp(1.f, Sg, is_q2));  // 1 - g^2 (tanh) or s - s^2 (sigmoid)
        const float BD = B * D;
        const float tc = FAST ? tanh_fast(ct) : tanhf_(ct);
        const float k1 = g4.w * fmaf(-tc, tc, 1.f);
        // the chain
        const float dh = dhv + dh_rec;
        const float dc = fmaf(dh, k1, dc_carry);
        dc_carry = dc * g4.y;
        const float dgj = valid ? selp(dh * tc, dc, is_q3) * BD : 0.f;
        dg_s[i][row_j] = dgj;
        if (q == 0) hprev_s[i][u] = s > 0 ? hpv : 0.f;
        if (valid) dpre[ipre] = from_f32<T_>(dgj);
        ipre += spre;
        __syncthreads();
        {
          const int pn = min(p + 1, LS_CH - 1);
          g4 = S.g4[pn][u];
          ct = S.ct[pn][u];
          cpn = S.ct[pn + 1][u];
          dhv = to_f32<T_>(S.dh[pn][u]);
          hpv = to_f32<T_>(S.hp[pn][u]);
        }
        {
          const float* hp = hprev_s[i];
          const float2 dg2 = make_float2(dgj, dgj);
#pragma unroll
          for (int k = 0; k < LS_H; k += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hp + k);
            acc2[k / 2] = ffma2(dg2, make_float2(hv.x, hv.y), acc2[k / 2]);
            acc2[k / 2 + 1] = ffma2(dg2, make_float2(hv.z, hv.w), acc2[k / 2 + 1]);
          }
          float2 p01 = make_float2(0.f, 0.f), p23 = make_float2(0.f, 0.f);
          const float* dq = dg_s[i] + q * LS_H;
#pragma unroll
          for (int jj = 0; jj < LS_H; jj += 4) {
            const float4 dv = *reinterpret_cast<const float4*>(dq + jj);
            p01 = ffma2(wt2[jj / 2], make_float2(dv.x, dv.y), p01);
            p23 = ffma2(wt2[jj / 2 + 1], make_float2(dv.z, dv.w), p23);
          }
          float part = (p01.x + p01.y) + (p23.x + p23.y);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          dh_rec = part;  // dh_prev of unit u, in all four lanes of its quad
        }
        // the next frame fills the other pair of exchange buffers: nobody can still be reading them (one barrier behind)
      }
    }
  }
#pragma unroll
  for (int k = 0; k < LS_H / 2; ++k) {
    atomicAdd(dwhh + (static_cast<size_t>(d) * LS_G + row_j) * LS_H + 2 * k, acc2[k].x);
    atomicAdd(dwhh + (static_cast<size_t>(d) * LS_G + row_j) * LS_H + 2 * k + 1, acc2[k].y);
  }
  // frames past the end: no gradient
  for (int t = len; t < T; ++t) dpre[(static_cast<size_t>(n) * T + t) * (2 * LS_G) + d * LS_G + j] = from_f32<T_>(0.f);
}

// which generation runs: LASR_LSTM_V1 = 1 / 2 / 3 forces one (tests, A/B timing); default: bf16 -> 3 (fast activations),
// fp32 -> 2 (the exact-parity mode keeps libm activations and the unre-associated backward).  Generation 3 indexes rows
// with 32-bit element offsets and copies 16-byte vectors: batches beyond 2^31 / 320 rows or operands that are not
// 16-byte aligned fall back to generation 2.
static int lstm_generation(int dtype, int N, int T, bool aligned16) {
  const char* e = getenv("LASR_LSTM_V1");
  int g = e != nullptr ? atoi(e) : 0;
  if (g < 1 || g > 3) g = dtype == LASR_BF16 ? 3 : 2;
  if (g == 3 && (!aligned16 || static_cast<long long>(N) * T * (2 * LS_G) >= (1LL << 31))) g = 2;
  return g;
}
static bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_bilstm_fwd(const void* pre, const float* whh, const int32_t* lengths, void* out, float* gates, float* cells,
                    int N, int T, int hidden, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || pre == nullptr || whh == nullptr || out == nullptr || gates == nullptr || cells == nullptr)
    return LASR_ERR_BAD_SHAPE;
  if (hidden != LS_H) return LASR_ERR_UNSUPPORTED;  // the reference's only configuration (QuartNetContext.py:157)
  if (reinterpret_cast<uintptr_t>(gates) & 15) return LASR_ERR_ALIGNMENT;
  dim3 grid(N, 2);
  const int gen = lstm_generation(dtype, N, T, aligned16(pre));
  float4* g4 = reinterpret_cast<float4*>(gates);
  if (dtype == LASR_F32) {
    const float* p = static_cast<const float*>(pre);
    float* o = static_cast<float*>(out);
    if (gen == 1)
      bilstm_fwd_kernel<float><<<grid, LS_G, 0, stream>>>(p, whh, lengths, o, g4, cells, T);
    else if (gen == 2)
      bilstm_fwd2_kernel<float><<<grid, LS_G, 0, stream>>>(p, whh, lengths, o, g4, cells, T);
    else
      bilstm_fwd3_kernel<float, false><<<grid, LS_G, 0, stream>>>(p, whh, lengths, o, g4, cells, T);
  } else if (dtype == LASR_BF16) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(pre);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
    if (gen == 1)
      bilstm_fwd_kernel<__nv_bfloat16><<<grid, LS_G, 0, stream>>>(p, whh, lengths, o, g4, cells, T);
    else if (gen == 2)
      bilstm_fwd2_kernel<__nv_bfloat16><<<grid, LS_G, 0, stream>>>(p, whh, lengths, o, g4, cells, T);
    else
      bilstm_fwd3_kernel<__nv_bfloat16, true><<<grid, LS_G, 0, stream>>>(p, whh, lengths, o, g4, cells, T);
  } else {
    return LASR_ERR_BAD_DTYPE;
  }
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_bilstm_bwd(const void* dout, const void* out, const float* gates, const float* cells, const float* whh,
                    const int32_t* lengths, void* dpre, float* dwhh, int N, int T, int hidden, int dtype,
                    lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || dout == nullptr || out == nullptr || gates == nullptr || cells == nullptr ||
      whh == nullptr || dpre == nullptr || dwhh == nullptr)
    return LASR_ERR_BAD_SHAPE;
  if (hidden != LS_H) return LASR_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(gates) & 15) return LASR_ERR_ALIGNMENT;
  dim3 grid(N, 2);
  const int gen = lstm_generation(dtype, N, T, aligned16(dout, out, cells, gates));
  const float4* g4 = reinterpret_cast<const float4*>(gates);
  if (dtype == LASR_F32) {
    const float* a = static_cast<const float*>(dout);
    const float* b = static_cast<const float*>(out);
    float* o = static_cast<float*>(dpre);
    if (gen == 1)
      bilstm_bwd_kernel<float><<<grid, LS_G, 0, stream>>>(a, b, g4, cells, whh, lengths, o, dwhh, T);
    else if (gen == 2)
      bilstm_bwd2_kernel<float><<<grid, LS_G, 0, stream>>>(a, b, g4, cells, whh, lengths, o, dwhh, T);
    else
      bilstm_bwd3_kernel<float, false><<<grid, LS_G, 0, stream>>>(a, b, g4, cells, whh, lengths, o, dwhh, T);
  } else if (dtype == LASR_BF16) {
    const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(dout);
    const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(out);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(dpre);
    if (gen == 1)
      bilstm_bwd_kernel<__nv_bfloat16><<<grid, LS_G, 0, stream>>>(a, b, g4, cells, whh, lengths, o, dwhh, T);
    else if (gen == 2)
      bilstm_bwd2_kernel<__nv_bfloat16><<<grid, LS_G, 0, stream>>>(a, b, g4, cells, whh, lengths, o, dwhh, T);
    else
      bilstm_bwd3_kernel<__nv_bfloat16, true><<<grid, LS_G, 0, stream>>>(a, b, g4, cells, whh, lengths, o, dwhh, T);
  } else {
    return LASR_ERR_BAD_DTYPE;
  }
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // extern "C"
