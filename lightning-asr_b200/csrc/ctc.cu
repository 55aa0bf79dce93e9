// log-softmax, CTC loss (alpha/beta lattices in log space + fused softmax-minus-occupancy gradient) and the
// greedy CTC decode.  Replaces nn.functional.log_softmax (models/QuartNet.py:290),
// torch.nn.CTCLoss(blank=V, reduction='none') (train.py:196, :76-78) and out.argmax(-1) +
// WER.ctc_decoder_predictions_tensor (train.py:80, utils/asr_metrics.py:153-171).
//
// Lattice kernels: one CTA per (utterance, direction) -- the alpha and beta recursions of an utterance are
// independent, so they run concurrently on different SMs.  Threads own lattice states; the previous column lives
// in double-buffered shared memory, the emission log-prob gather for frame t+1 is issued before frame t's barrier
// so the HBM/L2 latency overlaps the recursion.  (A single warp per utterance would be MUFU-bound: ~4 ex2/lg2 per
// state per frame x 13-24 states per lane x 8 cycles per warp-MUFU; spreading the states over 8 warps divides that.)
// The emission term can be read from raw logits + their row log-sum-exp, so log-probs need not be materialised.
#include "common.cuh"

#include <math_constants.h>
#include <cstdlib>

namespace lasr {

// ------------------------------------------------------------------------------------------------
// log-softmax: one warp per row
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__global__ void __launch_bounds__(256)
log_softmax_fwd_kernel(const T* __restrict__ x, float* __restrict__ lse_out, float* __restrict__ lp, int M, int V,
                       int ld) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const T* xr = x + static_cast<size_t>(row) * ld;
  float m = -CUDART_INF_F;
  float s = 0.f;
  if (sizeof(T) == 2 && V > 256 && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // wide rows of bf16 scores (the 4334-class decoder): 16-byte loads, and exponentials as ONE ex2.approx of a fused
    // multiply-add (libm's expf is ~8 instructions: the pass was issue-bound at 2.6x its HBM time).  ex2.approx is
    // good to 2^-22, the scores themselves carry 2^-9.
    const int Vr = (V + 7) & ~7;
    for (int c0 = lane * 8; c0 < Vr; c0 += 256) {
      const uint4 raw = *reinterpret_cast<const uint4*>(xr + c0);
      const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 xv = bf16x2_to_f32x2(wv[q]);
        if (c0 + 2 * q < V) m = fmaxf(m, xv.x);
        if (c0 + 2 * q + 1 < V) m = fmaxf(m, xv.y);
      }
    }
    m = warp_max(m);
    const float m2 = -m * 1.4426950408889634f;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = lane * 8; c0 < Vr; c0 += 256) {
      const uint4 raw = *reinterpret_cast<const uint4*>(xr + c0);  // second pass: the row is in L1 / L2
      const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 xv = bf16x2_to_f32x2(wv[q]);
        const float e0 = ex2_approx_ftz(fmaf(xv.x, 1.4426950408889634f, m2));
        const float e1 = ex2_approx_ftz(fmaf(xv.y, 1.4426950408889634f, m2));
        s4[q] += (c0 + 2 * q < V ? e0 : 0.f) + (c0 + 2 * q + 1 < V ? e1 : 0.f);
      }
    }
    s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  } else {
    for (int c = lane; c < V; c += 32) m = fmaxf(m, to_f32<T>(xr[c]));
    m = warp_max(m);
    for (int c = lane; c < V; c += 32) s += expf(to_f32<T>(xr[c]) - m);
  }
  s = warp_sum(s);
  const float lse = m + logf(s);
  if (lane == 0) lse_out[row] = lse;
  if (lp != nullptr) {
    float* lr = lp + static_cast<size_t>(row) * V;
    for (int c = lane; c < V; c += 32) lr[c] = to_f32<T>(xr[c]) - lse;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
log_softmax_bwd_kernel(const float* __restrict__ dlp, const float* __restrict__ lp, T* __restrict__ dx, int M, int V,
                       int ld) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* dr = dlp + static_cast<size_t>(row) * V;
  const float* lr = lp + static_cast<size_t>(row) * V;
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += dr[c];
  s = warp_sum(s);
  T* o = dx + static_cast<size_t>(row) * ld;
  for (int c = lane; c < ld; c += 32) o[c] = from_f32<T>(c < V ? dr[c] - expf(lr[c]) * s : 0.f);
}

// ------------------------------------------------------------------------------------------------
// CTC lattices
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

constexpr int CTC_RING = 8;      // frames of emissions in flight (cp.async ring)

// log(e^a + e^b + e^c) with the largest term factored out: its exponential is exactly 1, so only two ex2 + one lg2
// go to the MUFU pipe (the per-frame critical path of the recursion)
__device__ __forceinline__ float lse3_fast(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + __logf(1.f + __expf(mid - m) + __expf(lo - m));
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One CTA per (utterance, direction).  Thread i owns lattice states i, i+THREADS, ...; the launcher picks THREADS so
// that SPT == 1 whenever the lattice has <= 1024 states: the recursion is a chain of T dependent steps, each a
// handful of dependent instructions per state, so spreading states over many warps (latency hiding across warps of
// one scheduler) is what shortens a step -- 4 states per thread at 128 threads measured 1.03 us per frame.
// The emission of state s at frame
// t is ONE element of the [T, V] score matrix; the 4-byte word holding it is fetched CTC_RING frames ahead with
// cp.async into a per-thread slot of a shared-memory ring, so the sequential recursion never waits on HBM/L2.
// shared-memory accesses through 32-bit shared addresses (a ping-ponged `float*` makes the compiler fall back to
// generic loads plus a shared-window conversion per access)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cp_async_4_sa(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log(e^a + e^b + e^c), branch free: the largest term is factored out (its exponential is exactly 1: two ex2 + one lg2
// on the MUFU pipe); an all -inf input is detected with one compare / select, and the arithmetic runs on a clamped
// maximum so it never forms inf - inf
__device__ __forceinline__ float lse3_bf(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  const float mc = fmaxf(m, -3.0e38f);
  const float e1 = ex2_approx((mid - mc) * 1.4426950408889634f);
  const float e2 = ex2_approx((lo - mc) * 1.4426950408889634f);
  const float r = fmaf(lg2_approx(1.f + e1 + e2), 0.6931471805599453f, mc);
  return m == -CUDART_INF_F ? -CUDART_INF_F : r;
}

template <typename T, int SPT, int CTC_THREADS>
__global__ void __launch_bounds__(CTC_THREADS)
ctc_lattice_kernel(const T* __restrict__ x, const float* __restrict__ lse, const int64_t* __restrict__ targets,
                   const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len, float* __restrict__ alpha,
                   float* __restrict__ beta, float* __restrict__ nll, int T_len, int ldx, int S_max, int blank) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const bool backward = blockIdx.y == 1;
  const int nthreads = blockDim.x;  // round_up(Lp_max, 32) for SPT == 1: no idle warps in the per-frame barrier
  const int Lp_max = 2 * S_max + 1;
  const int colw = Lp_max + 4;      // a lattice column: 2 cells of -inf, Lp_max states, 2 cells of -inf
  const int W = SPT * nthreads;     // ring words per frame
  // sm: column A [colw], column B [colw], emission ring [CTC_RING][W] raw words, lse ring [CTC_RING]
  const uint32_t colA = smem_u32(sm), colB = colA + 4u * colw;
  const uint32_t ring_base = colB + 4u * colw;
  const uint32_t lring_base = ring_base + 4u * CTC_RING * W;
  const int Tn = in_len[n];
  const int Sn = tgt_len[n];
  const int Lp = 2 * Sn + 1;
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  float* lat = (backward ? beta : alpha) + static_cast<size_t>(n) * T_len * Lp_max;
  const size_t row0 = static_cast<size_t>(n) * T_len;
  const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
  constexpr bool kHalf = sizeof(T) == 2;

  // infeasible / degenerate cases (torch: loss = inf when the target does not fit)
  if (Tn <= 0 || Sn > S_max || Tn > T_len) {
    if (!backward && threadIdx.x == 0) nll[n] = (Tn == 0 && Sn == 0) ? 0.f : CUDART_INF_F;
    return;
  }
  const int t_first = backward ? Tn - 1 : 0;
  const int dt = backward ? -1 : 1;

  // Per-state constants.  The recursion is issue-bound (one barrier and a few dozen instructions per frame and
  // warp), so everything that does not depend on the frame lives in registers: the running global pointer of the
  // state's emission word (advanced one row per issue), the shift that extracts a bf16 from its word (rows start on
  // even elements: the launcher requires an even ldx), the byte offsets of the state and of its two predecessors
  // inside a column -- a forbidden transition simply points at a cell that always holds -inf --, and the running
  // pointer into the saved lattice.
  const uint32_t* src[SPT];
  uint32_t shl[SPT], off0[SPT], off1[SPT], off2[SPT];
  bool act[SPT];
  float* latp[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int st = threadIdx.x + i * nthreads;
    act[i] = st < Lp;
    int lab = blank;
    bool skip_ok = false;  // may take the s-2 (alpha) / s+2 (beta) transition
    if (act[i] && (st & 1)) {
      lab = static_cast<int>(tg[st >> 1]);
      if (!backward)
        skip_ok = (st >= 2) && lab != static_cast<int>(tg[(st >> 1) - 1]);
      else
        skip_ok = (st + 2 < Lp) && lab != static_cast<int>(tg[(st >> 1) + 1]);
    }
    const size_t e = (row0 + t_first) * static_cast<size_t>(ldx) + lab;
    src[i] = xw + (kHalf ? (e >> 1) : e);
    shl[i] = (kHalf && !(e & 1)) ? 16u : 0u;
    off0[i] = 4u * (st + 2);
    // alpha: predecessors s-1, s-2 (cells 1 / 0 for s = 0 are -inf); beta: s+1, s+2, cut at Lp (cells past the
    // utterance's own Lp are never written and stay -inf)
    const int p1 = backward ? st + 1 : st - 1;
    const int p2 = backward ? st + 2 : st - 2;
    off1[i] = 4u * ((backward && p1 >= Lp) ? 0 : p1 + 2);
    off2[i] = 4u * (skip_ok ? p2 + 2 : 0);
    latp[i] = lat + static_cast<size_t>(t_first) * Lp_max + st;
  }
  const ptrdiff_t src_step = static_cast<ptrdiff_t>(dt) * (kHalf ? (ldx >> 1) : ldx);
  const ptrdiff_t lat_step = static_cast<ptrdiff_t>(dt) * Lp_max;
  const float* lse_src = lse != nullptr ? lse + row0 + t_first : nullptr;
  const bool has_lse = lse != nullptr;
  int issued = 0;                                    // frames handed to cp.async so far
  uint32_t slot = ring_base + 4u * threadIdx.x;      // this thread's word in the ring slot of the next issue
  uint32_t lslot = lring_base;
  int slot_idx = 0;
  static_assert((CTC_RING & (CTC_RING - 1)) == 0, "ring positions are masked");

  auto issue = [&]() {
    if (issued < Tn) {
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        if (act[i]) cp_async_4_sa(slot + 4u * i * nthreads, src[i]);
        src[i] += src_step;
      }
      if (has_lse) {
        if (threadIdx.x == 0) cp_async_4_sa(lslot, lse_src);
        lse_src += dt;
      }
    }
    ++issued;
    slot += 4u * W;
    lslot += 4u;
    if (++slot_idx == CTC_RING) {
      slot_idx = 0;
      slot = ring_base + 4u * threadIdx.x;
      lslot = lring_base;
    }
    cp_async_commit();
  };

  // both columns start as -inf everywhere (guard cells and states beyond this utterance's Lp stay that way)
  for (int i = threadIdx.x; i < 2 * colw; i += nthreads) sts_f32(colA + 4u * i, -CUDART_INF_F);
#pragma unroll 1
  for (int st = 0; st < CTC_RING; ++st) issue();
  cp_async_wait<CTC_RING - 1>();
  __syncthreads();
  // initial column (in A)
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int st = threadIdx.x + i * nthreads;
    if (act[i]) {
      float v = -CUDART_INF_F;
      const bool start = backward ? (st == Lp - 1 || st == Lp - 2) : (st == 0 || st == 1);
      if (start) {
        const uint32_t w = lds_u32(ring_base + 4u * st);
        v = __uint_as_float(kHalf ? ((w << shl[i]) & 0xffff0000u) : w);
        if (has_lse) v -= lds_f32(lring_base);
        sts_f32(colA + off0[i], v);
      }
      *latp[i] = v;
    }
  }

  auto frame = [&](int step, uint32_t prev, uint32_t next) {
    cp_async_wait<CTC_RING - 2>();  // this thread's own words of frame `step` have landed
    const uint32_t pos = static_cast<uint32_t>(step) & (CTC_RING - 1);
    uint32_t wv[SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) wv[i] = lds_u32(ring_base + 4u * (pos * W + threadIdx.x + i * nthreads));
    __syncthreads();                // previous column (and thread 0's lse word) visible to everyone
    issue();                        // refill the slot consumed by the previous frame
    const float lr = has_lse ? lds_f32(lring_base + 4u * pos) : 0.f;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      latp[i] += lat_step;
      if (act[i]) {
        const float a = lds_f32(prev + off0[i]);
        const float b2 = lds_f32(prev + off1[i]);
        const float c2 = lds_f32(prev + off2[i]);
        const float em = __uint_as_float(kHalf ? ((wv[i] << shl[i]) & 0xffff0000u) : wv[i]);
        const float v = lse3_bf(a, b2, c2) + (em - lr);
        sts_f32(next + off0[i], v);
        *latp[i] = v;
      }
    }
  };
  int step = 1;
#pragma unroll 1
  for (; step + 1 < Tn; step += 2) {
    frame(step, colA, colB);
    frame(step + 1, colB, colA);
  }
  uint32_t last = colA;
  if (step < Tn) {
    frame(step, colA, colB);
    last = colB;
  }
  cp_async_wait<0>();
  __syncthreads();
  if (!backward && threadIdx.x == 0) {
    const float a = lds_f32(last + 4u * (Lp - 1 + 2));
    const float b = Lp >= 2 ? lds_f32(last + 4u * (Lp - 2 + 2)) : -CUDART_INF_F;
    const float m = fmaxf(a, b);
    nll[n] = (m == -CUDART_INF_F) ? CUDART_INF_F : -(m + logf(expf(a - m) + expf(b - m)));
  }
}

constexpr uint32_t kCtcSentinel = 0x7fc0deadu;  // a NaN payload no arithmetic produces: "slot empty"
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Warp-pipelined lattices (round 2, the default).  ncu on the kernel above: a warp spends a frame issuing its own
// dependent instruction stream (~90 instructions), then waits at the CTA barrier for the slowest warp, then pays the
// STS -> barrier -> LDS round trip before the next frame can start.  Here a frame has NO CTA barrier and no
// shared-memory column.  Lane l of warp w owns the even state s0 = 64w + 2l (a blank) and the odd state s0 + 1 (label
// 32w + l) in registers: half the warps of a one-state-per-thread layout issue per SM (7 instead of 13 for 401 states),
// every lane has two independent log-sum-exp chains in flight, and the neighbourhood shrinks -- a blank has no skip
// transition and the label's s-1 is the lane's own blank, so alpha needs ONE value from the lane below (its label
// state: one shuffle), beta two from the lane above.  Only the pair across a WARP boundary travels through shared
// memory: after frame f the edge lane of warp w stores its pair as ONE 8-byte word into slot f mod 16 of the next warp's
// inbox; that warp reads the slot a frame EARLY (the load is issued at the top of frame f for use in frame f+1), falls
// back to spinning only if it still holds the sentinel (a NaN payload no arithmetic produces -- value and flag are the
// same word, so no fence), and puts the sentinel back.  Warp 0 depends on nobody, warp w settles a hand-over latency
// behind warp w-1: the warps form a skewed pipeline, each running at the speed of its own shuffle -> log-sum-exp chain.
// Flow control: a producer looks once every 8 frames whether the slot 8 frames ahead has been emptied.  Emissions (and
// lse) are gathered 8 frames ahead with plain 16-bit / 32-bit loads into a REGISTER ring (the frame loop is unrolled
// over it, so ring slots and mailbox slots are immediates; no cp.async groups, no shared-memory ring).
// log(e^a + e^b) is evaluated as lse3_bf(a, b, -inf) term by term (its third exponential is exactly 0), every state sees
// the operands of ctc_lattice_kernel in the same order: results are bit-identical
// (tests/test_kernels_gpu.py::test_ctc_warp_pipelined_lattices_are_bit_identical).
// Measured (N = 32, T' = 801, S = 200, bf16 logits + lse, stand-alone incl. ~8 us of event overhead): round-1 kernel
// 266 us; the same barrier kernel with the register ring and an unrolled loop 173 us; a lattice cut over a 2- / 4-CTA
// cluster with the boundary states posted into the neighbour's shared memory (st.shared::cluster) 184 / 182 us (the
// spinning consumer costs more than the second SM brings); one state per lane + shuffles 147 us; this kernel 141 us.
// What is left is the length of one warp's instruction stream per frame (~80 instructions for 64 states, most of them
// dependent) against a floor of ~130 cycles for shuffle + log-sum-exp.
// ------------------------------------------------------------------------------------------------
// keep a running pointer / shared address in its register: without this the compiler re-derives it from the loop
// counter (or re-reads the shared window base) inside the frame loop, several instructions per use
template <typename P>
__device__ __forceinline__ void keep_ptr(P*& p) {
  asm volatile("" : "+l"(p));
}
__device__ __forceinline__ void keep_u32(uint32_t& a) {
  asm volatile("" : "+r"(a));
}

constexpr int CTCW_SLOTS = 16;
__device__ __forceinline__ uint2 lds_v2_volatile(uint32_t addr) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v2_volatile(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

__device__ __forceinline__ float lse2_bf(float a, float b) {
  const float m = fmaxf(a, b), lo = fminf(a, b);
  const float mc = fmaxf(m, -3.0e38f);
  const float e2 = ex2_approx((lo - mc) * 1.4426950408889634f);
  const float r = fmaf(lg2_approx(1.f + e2), 0.6931471805599453f, mc);
  return m == -CUDART_INF_F ? -CUDART_INF_F : r;
}

template <typename T, bool LSE, bool BWD>
__device__ __forceinline__ void ctc_warp2_lattice(const T* __restrict__ x, const float* __restrict__ lse,
                                                  const int64_t* __restrict__ targets, float* __restrict__ lat, int n,
                                                  int Tn, int Lp, int T_len, int ldx, int S_max, int blank,
                                                  uint32_t inbox0, float& v0, float& v1) {
  constexpr bool kHalf = sizeof(T) == 2;
  constexpr uint32_t kNegInfWord = kHalf ? 0xff80u : 0xff800000u;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int Lp_max = 2 * S_max + 1;
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  const size_t row0 = static_cast<size_t>(n) * T_len;
  const int t_first = BWD ? Tn - 1 : 0;
  const int dt = BWD ? -1 : 1;
  const int s0 = 2 * tid, s1 = s0 + 1;
  const bool act0 = s0 < Lp, act1 = s1 < Lp;
  int lab = blank;
  bool skip_ok = false;  // the label state may take the s-2 (alpha) / s+2 (beta) transition
  if (act1) {
    lab = static_cast<int>(tg[tid]);
    if (!BWD)
      skip_ok = (tid >= 1) && lab != static_cast<int>(tg[tid - 1]);
    else
      skip_ok = (s1 + 2 < Lp) && lab != static_cast<int>(tg[tid + 1]);
  }
  const size_t rowe = (row0 + t_first) * static_cast<size_t>(ldx);
  const char* pB = reinterpret_cast<const char*>(x) + (rowe + blank) * sizeof(T);
  const char* pL = reinterpret_cast<const char*>(x) + (rowe + lab) * sizeof(T);
  const ptrdiff_t src_step = static_cast<ptrdiff_t>(dt) * ldx * static_cast<ptrdiff_t>(sizeof(T));
  const char* lseb = LSE ? reinterpret_cast<const char*>(lse + row0 + t_first) : nullptr;
  const ptrdiff_t lse_step = static_cast<ptrdiff_t>(dt) * 4;
  char* latb = reinterpret_cast<char*>(lat + (row0 + t_first) * Lp_max + (act0 ? s0 : 0));
  const ptrdiff_t lat_step = static_cast<ptrdiff_t>(dt) * Lp_max * 4;
  // the pipeline: alpha flows from warp w to w+1 (lane 31 -> lane 0), beta from w+1 to w (lane 0 -> lane 31)
  const bool more_above = 64 * (warp + 1) < Lp;
  const bool takes = BWD ? more_above : warp > 0;
  const bool posts = BWD ? warp > 0 : more_above;
  const bool is_edge = lane == (BWD ? 31 : 0);
  const bool post_lane = posts && lane == (BWD ? 0 : 31);
  uint32_t inbox = inbox0 + 8u * CTCW_SLOTS * warp;
  uint32_t outbox = inbox0 + 8u * CTCW_SLOTS * (BWD ? warp - 1 : warp + 1);
  keep_u32(inbox);
  keep_u32(outbox);

  uint32_t wB[CTC_RING], wL[CTC_RING];
  float lr[CTC_RING];
  auto fetch = [&](uint32_t& b, uint32_t& w, float& l) {  // the next frame not yet requested (idle states keep -inf)
    if (act0) {
      if constexpr (kHalf)
        b = *reinterpret_cast<const unsigned short*>(pB);
      else
        b = *reinterpret_cast<const uint32_t*>(pB);
    }
    if (act1) {
      if constexpr (kHalf)
        w = *reinterpret_cast<const unsigned short*>(pL);
      else
        w = *reinterpret_cast<const uint32_t*>(pL);
    }
    if constexpr (LSE) l = *reinterpret_cast<const float*>(lseb);
    pB += src_step;
    pL += src_step;
    keep_ptr(pB);
    keep_ptr(pL);
    if constexpr (LSE) {
      lseb += lse_step;
      keep_ptr(lseb);
    }
  };
  auto emission = [&](uint32_t w, float l) {
    const float em = __uint_as_float(kHalf ? (w << 16) : w);
    return LSE ? em - l : em;
  };
  // what the lane below (alpha) / above (beta) holds: alpha needs its label state, beta its blank and its label
  float nb0 = -CUDART_INF_F, nb1 = -CUDART_INF_F;
  auto neighbours = [&]() {
    if (BWD) {
      nb0 = __shfl_down_sync(0xffffffffu, v0, 1);
      nb1 = __shfl_down_sync(0xffffffffu, v1, 1);
    } else {
      nb1 = __shfl_up_sync(0xffffffffu, v1, 1);
    }
  };
  auto post = [&](uint32_t slot) {  // the edge lane's own pair: what the next warp's edge lane will want
    if (post_lane) sts_v2_volatile(outbox + 8u * slot, __float_as_uint(v0), __float_as_uint(v1));
  };

#pragma unroll
  for (int f = 0; f < CTC_RING; ++f) {
    wB[f] = wL[f] = kNegInfWord;
    lr[f] = 0.f;
    if (f < Tn) fetch(wB[f], wL[f], lr[f]);
  }
  {  // frame 0
    const bool start0 = act0 && (BWD ? s0 == Lp - 1 : s0 == 0);
    const bool start1 = act1 && (BWD ? s1 == Lp - 2 : s1 == 1);
    if (start0) v0 = emission(wB[0], lr[0]);
    if (start1) v1 = emission(wL[0], lr[0]);
    // the mailbox reads "empty" by value: keep an input's NaN payload from looking like it (only frame 0 can carry raw
    // input bits; every later value is the result of an addition, and arithmetic returns the canonical NaN)
    if (__float_as_uint(v0) == kCtcSentinel) v0 = __uint_as_float(0x7fffffffu);
    if (__float_as_uint(v1) == kCtcSentinel) v1 = __uint_as_float(0x7fffffffu);
    if (CTC_RING < Tn) fetch(wB[0], wL[0], lr[0]);
    if (act0) *reinterpret_cast<float*>(latb) = v0;
    if (act1) *reinterpret_cast<float*>(latb + 4) = v1;
    neighbours();
    post(0);
  }
  uint2 early = make_uint2(kCtcSentinel, kCtcSentinel);  // the producer's word for the coming frame, read a frame early
  if (takes) early = lds_v2_volatile(inbox);
  // frame f: `slot` = (f - 1) mod 16 holds the producer's frame f-1; the frame's own result goes to slot f mod 16
  auto frame = [&](uint32_t& b, uint32_t& w, float& l, uint32_t slot, bool refill, bool more_frames) {
    if (takes) {
      uint2 m = early;
      while (m.x == kCtcSentinel) m = lds_v2_volatile(inbox + 8u * slot);
      if (is_edge) sts_v2_volatile(inbox + 8u * slot, kCtcSentinel, kCtcSentinel);  // slot free again
      if (more_frames) early = lds_v2_volatile(inbox + 8u * ((slot + 1) & (CTCW_SLOTS - 1)));
      if (BWD) nb0 = is_edge ? __uint_as_float(m.x) : nb0;
      nb1 = is_edge ? __uint_as_float(m.y) : nb1;
    } else {
      if (BWD) nb0 = is_edge ? -CUDART_INF_F : nb0;
      nb1 = is_edge ? -CUDART_INF_F : nb1;
    }
    const float eb = emission(b, l), el = emission(w, l);
    float u0, u1;
    if (BWD) {
      // blank s0: stays or moves to the lane's own label s0+1; label s1: stays, s1+1 = next blank, s1+2 = next label
      u0 = lse2_bf(v0, v1) + eb;
      u1 = lse3_bf(v1, nb0, skip_ok ? nb1 : -CUDART_INF_F) + el;
    } else {
      // blank s0: from itself or the label below; label s1: from itself, the lane's own blank, the label below
      u0 = lse2_bf(v0, nb1) + eb;
      u1 = lse3_bf(v1, v0, skip_ok ? nb1 : -CUDART_INF_F) + el;
    }
    v0 = u0;
    v1 = u1;
    if (refill) fetch(b, w, l);
    latb += lat_step;
    keep_ptr(latb);
    if (act0) *reinterpret_cast<float*>(latb) = v0;
    if (act1) *reinterpret_cast<float*>(latb + 4) = v1;
    neighbours();
    post((slot + 1) & (CTCW_SLOTS - 1));
  };
  // a producer may run at most 16 frames ahead: before posting frames f .. f+7 it makes sure the slot of frame f+7 (last
  // used by frame f-9) has been emptied -- the consumer empties slots in order
  auto wait_room = [&](uint32_t slot_last) {
    if (posts) {
      while (lds_v2_volatile(outbox + 8u * slot_last).x != kCtcSentinel) {
      }
    }
  };
  int step = 1;
  // steady state, 16 frames per trip (slot numbers and ring registers are immediates); all refills exist
#pragma unroll 1
  for (; step + 16 + CTC_RING <= Tn; step += 16) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if ((u & 7) == 0) wait_room((u + 1 + 7) & (CTCW_SLOTS - 1));
      constexpr int R = CTC_RING - 1;
      frame(wB[(1 + u) & R], wL[(1 + u) & R], lr[(1 + u) & R], u, true, true);
    }
  }
#pragma unroll 1
  for (; step < Tn; step += 16) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (step + u < Tn) {
        if ((u & 7) == 0) wait_room((u + 1 + 7) & (CTCW_SLOTS - 1));
        constexpr int R = CTC_RING - 1;
        frame(wB[(1 + u) & R], wL[(1 + u) & R], lr[(1 + u) & R], u, step + u + CTC_RING < Tn, step + u + 1 < Tn);
      }
    }
  }
}

template <typename T, bool LSE, int MAXT>
__global__ void __launch_bounds__(MAXT)
ctc_lattice_warp2_kernel(const T* __restrict__ x, const float* __restrict__ lse, const int64_t* __restrict__ targets,
                         const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len,
                         float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ nll, int T_len, int ldx,
                         int S_max, int blank) {
  pdl_launch_dependents();
  extern __shared__ float sm[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int warp = tid >> 5, nwarps = nthreads >> 5;
  const int n = blockIdx.x;
  const bool backward = blockIdx.y == 1;
  // sm: inbox [nwarps][CTCW_SLOTS] 8-byte words, last column [2 * nthreads]
  const uint32_t inbox0 = smem_u32(sm);
  const uint32_t fin = inbox0 + 8u * CTCW_SLOTS * nwarps;
  for (int i = tid; i < 2 * CTCW_SLOTS * nwarps; i += nthreads) sts_u32(inbox0 + 4u * i, kCtcSentinel);
  __syncthreads();
  pdl_wait();
  const int Tn = in_len[n];
  const int Sn = tgt_len[n];
  const int Lp = 2 * Sn + 1;
  // infeasible / degenerate cases (torch: loss = inf when the target does not fit)
  if (Tn <= 0 || Sn > S_max || Tn > T_len) {
    if (!backward && tid == 0) nll[n] = (Tn == 0 && Sn == 0) ? 0.f : CUDART_INF_F;
    return;
  }
  float v0 = -CUDART_INF_F, v1 = -CUDART_INF_F;
  if (64 * warp < Lp) {
    if (backward)
      ctc_warp2_lattice<T, LSE, true>(x, lse, targets, beta, n, Tn, Lp, T_len, ldx, S_max, blank, inbox0, v0, v1);
    else
      ctc_warp2_lattice<T, LSE, false>(x, lse, targets, alpha, n, Tn, Lp, T_len, ldx, S_max, blank, inbox0, v0, v1);
  }
  if (!backward) {
    sts_f32(fin + 8u * tid, v0);
    sts_f32(fin + 8u * tid + 4u, v1);
    __syncthreads();
    if (tid == 0) {
      const float a = lds_f32(fin + 4u * (Lp - 1));
      const float b = Lp >= 2 ? lds_f32(fin + 4u * (Lp - 2)) : -CUDART_INF_F;
      const float m = fmaxf(a, b);
      nll[n] = (m == -CUDART_INF_F) ? CUDART_INF_F : -(m + logf(expf(a - m) + expf(b - m)));
    }
  }
}

// gradient: one warp per frame row.  occ[c] accumulated in warp-private shared memory.
template <typename T, typename GT>
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const T* __restrict__ x, const float* __restrict__ lse, const int64_t* __restrict__ targets,
                const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len,
                const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ nll,
                const int32_t* __restrict__ scales, const float* __restrict__ grad_out, GT* __restrict__ grad, int N,
                int T_len, int V, int ldx, int ldg, int S_max, int blank, int warps) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float occ_all[];  // [warps][V]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * warps + w;
  if (w >= warps || row >= static_cast<long long>(N) * T_len) return;
  float* occ = occ_all + static_cast<size_t>(w) * V;
  const int n = static_cast<int>(row / T_len);
  const int t = static_cast<int>(row - static_cast<long long>(n) * T_len);
  GT* gr = grad + static_cast<size_t>(row) * ldg;
  const int Tn = in_len[n];
  const int Sn = tgt_len[n];
  // frames past the utterance, and utterances whose lattice was never built (bad lengths: the lattice kernel returns
  // early and leaves alpha / beta unwritten), get a zero gradient row
  if (t >= Tn || Sn < 0 || Sn > S_max) {
    for (int c = lane; c < ldg; c += 32) gr[c] = from_f32<GT>(0.f);
    return;
  }
  const int Lp = 2 * Sn + 1;
  const int Lp_max = 2 * S_max + 1;
  const T* xr = x + static_cast<size_t>(row) * ldx;
  const float l = lse ? lse[row] : 0.f;
  const float nl = nll[n];
  const float go = grad_out[n];
  for (int c = lane; c < V; c += 32) occ[c] = 0.f;
  __syncwarp();
  const float* ar = alpha + static_cast<size_t>(row) * Lp_max;
  const float* br = beta + static_cast<size_t>(row) * Lp_max;
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  if (scales != nullptr) {
    // scaled lattices (ctc_lattice2_kernel): alpha = a * 2^EA, beta~ = b * 2^EB, P = Pm * 2^EP -> occupancy without any
    // logarithm, exponential or division per state
    const size_t NT = static_cast<size_t>(N) * T_len;
    const int ea = scales[static_cast<size_t>(n) * T_len + t];
    const int eb = scales[NT + static_cast<size_t>(n) * T_len + t];
    const float Pm = __uint_as_float(static_cast<uint32_t>(scales[2 * NT + 2 * n]));
    const int ep = scales[2 * NT + 2 * n + 1];
    const float w = scalbnf(1.0f / Pm, ea + eb - ep);
    // every other state is the blank: its occupancy is summed in registers (even lanes own even states) and reduced with
    // shuffles -- 16 lanes adding to the same shared-memory word per instruction was a large part of this kernel's time
    float bsum = 0.f;
    for (int s = lane; s < Lp; s += 32) {
      const float v = ar[s] * br[s] * w;
      if (s & 1)
        atomicAdd(&occ[static_cast<int>(tg[s >> 1])], v);
      else
        bsum += v;
    }
    bsum = warp_sum(bsum);
    __syncwarp();
    if (lane == 0) occ[blank] += bsum;
  } else {
    float bsum = 0.f;
    for (int s = lane; s < Lp; s += 32) {
      const int label = (s & 1) ? static_cast<int>(tg[s >> 1]) : blank;
      const float lpv = to_f32<T>(xr[label]) - l;
      const float v = expf(ar[s] + br[s] + nl - lpv);
      if (s & 1)
        atomicAdd(&occ[label], v);
      else
        bsum += v;
    }
    bsum = warp_sum(bsum);
    __syncwarp();
    if (lane == 0) occ[blank] += bsum;
  }
  __syncwarp();
  for (int c = lane; c < ldg; c += 32) {
    float g = 0.f;
    if (c < V) g = (expf(to_f32<T>(xr[c]) - l) - occ[c]) * go;
    gr[c] = from_f32<GT>(g);
  }
}


// The same pass for small vocabularies (V <= 128: the character models), 8 frames of ONE utterance per CTA: the labels
// are read once per CTA into shared memory as int32 (the kernel above re-reads the int64 targets from global memory for
// every frame row), the row's log-probs once per warp, so a lattice state costs two coalesced global loads (alpha, beta),
// two shared-memory reads and one exponential; four states per lane are requested before the first is used.  Same
// arithmetic in the same order per row: bit-identical gradients.
template <typename T, typename GT>
__global__ void __launch_bounds__(256)
ctc_grad_small_kernel(const T* __restrict__ x, const float* __restrict__ lse, const int64_t* __restrict__ targets,
                      const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len,
                      const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ nll,
                      const float* __restrict__ grad_out, GT* __restrict__ grad, int T_len, int V, int ldx, int ldg,
                      int S_max, int blank, int S_pad, int Vp) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float gsm[];  // labels [S_pad] int32, then per warp: lp [Vp], occ [Vp]
  int* labels = reinterpret_cast<int*>(gsm);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* lpw = gsm + S_pad + static_cast<size_t>(w) * 2 * Vp;
  float* occ = lpw + Vp;
  const int n = blockIdx.y;
  const int t = blockIdx.x * 8 + w;
  const int Tn = in_len[n];
  const int Sn = tgt_len[n];
  const bool built = !(Sn < 0 || Sn > S_max);  // bad lengths: the lattice kernel left alpha / beta unwritten
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  if (built)
    for (int i = threadIdx.x; i < Sn; i += 256) labels[i] = static_cast<int>(tg[i]);
  __syncthreads();
  if (t >= T_len) return;
  const size_t row = static_cast<size_t>(n) * T_len + t;
  GT* gr = grad + row * ldg;
  if (t >= Tn || !built) {  // frames past the utterance get a zero gradient row
    for (int c = lane; c < ldg; c += 32) gr[c] = from_f32<GT>(0.f);
    return;
  }
  const int Lp = 2 * Sn + 1;
  const int Lp_max = 2 * S_max + 1;
  const T* xr = x + row * ldx;
  const float l = lse ? lse[row] : 0.f;
  const float nl = nll[n];
  const float go = grad_out[n];
  for (int c = lane; c < V; c += 32) {
    lpw[c] = to_f32<T>(xr[c]) - l;
    occ[c] = 0.f;
  }
  __syncwarp();
  const float* ar = alpha + row * Lp_max;
  const float* br = beta + row * Lp_max;
  float bsum = 0.f;
  constexpr int GU = 4;  // states per lane requested before the first is used (8 measured no better)
  for (int s0 = lane; s0 < Lp; s0 += 32 * GU) {
    float a[GU], b[GU];
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const int st = s0 + 32 * k;
      a[k] = b[k] = 0.f;
      if (st < Lp) {
        a[k] = ar[st];
        b[k] = br[st];
      }
    }
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const int st = s0 + 32 * k;
      if (st < Lp) {
        const int label = (st & 1) ? labels[st >> 1] : blank;
        const float v = expf(a[k] + b[k] + nl - lpw[label]);
        if (st & 1)
          atomicAdd(&occ[label], v);
        else
          bsum += v;
      }
    }
  }
  bsum = warp_sum(bsum);
  __syncwarp();
  if (lane == 0) occ[blank] += bsum;
  __syncwarp();
  for (int c = lane; c < ldg; c += 32) {
    float g = 0.f;
    if (c < V) g = (expf(lpw[c]) - occ[c]) * go;
    gr[c] = from_f32<GT>(g);
  }
}


// The same pass for LARGE vocabularies (the 4334-class AISHELL decoder).  ctc_grad_kernel keeps a dense occ[V] per warp
// (17 KB: five warps per CTA, two CTAs per SM) and zeroes / re-reads all of it for every frame row although only the
// <= S labels of the utterance can be occupied: 0.58 ms for a pass whose 526 MB take 81 us at HBM speed.  Here a CTA is
// 8 frames of one utterance and shares, built once: the labels, a class -> first-position table tab[V] and rep[j] = the
// first position that carries position j's label; a warp accumulates occupancy per POSITION (occ_pos[rep[j]], S floats),
// and the class pass looks a class up through tab.  8 classes per lane and trip as 16-byte vectors when rows allow.
// Same additions per class (positions of one class that collide in one shared-memory atomic instruction may be
// serialised in another order than in the dense layout: <= 1 ulp of fp32 on such a class).  The bf16 -> bf16 vector path
// takes its softmax term from ex2.approx (2^-22 relative: a bf16 rounding flips on ~1 element in 1000).
constexpr int kCtcNoPos = 0x7fffffff;
template <typename T, typename GT, bool STAGE>
__global__ void __launch_bounds__(256, STAGE ? 2 : 6)
ctc_grad_large_kernel(const T* __restrict__ x, const float* __restrict__ lse, const int64_t* __restrict__ targets,
                      const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len,
                      const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ nll,
                      const float* __restrict__ grad_out, GT* __restrict__ grad, int T_len, int V, int ldx, int ldg,
                      int S_max, int blank, int S_pad, int Vp) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float gsm[];  // labels [S_pad], rep [S_pad], tab [Vp] (int32), then per warp: occ_pos [S_pad]
  int* labels = reinterpret_cast<int*>(gsm);
  int* rep = labels + S_pad;
  int* tab = rep + S_pad;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* occ = gsm + 2 * S_pad + Vp + static_cast<size_t>(w) * S_pad;
  const int n = blockIdx.y;
  const int t = blockIdx.x * 8 + w;
  const int Tn = in_len[n];
  const int Sn = tgt_len[n];
  const bool built = !(Sn < 0 || Sn > S_max);  // bad lengths: the lattice kernel left alpha / beta unwritten
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  // STAGE: the warp's score row goes to shared memory in one burst of 16-byte cp.async requests (17 per lane at the
  // 4334-class vocabulary, all in flight while the CTA builds its tables); the emission gathers of the state pass, the
  // class pass and the rewrites then read shared memory.  Without it the pass was latency-bound on its global loads
  // (one 16-byte load in flight per lane in the class pass, scattered 2-byte gathers in the state pass: 70 % of the
  // warp samples on long-scoreboard stalls) and fetched 448 MB for 304 MB of operands.
  T* xs = nullptr;
  if (STAGE) {
    xs = reinterpret_cast<T*>(gsm + 2 * S_pad + Vp + 8 * static_cast<size_t>(S_pad)) + static_cast<size_t>(w) * ldx;
    if (t < T_len && t < Tn && built) {
      const T* src = x + (static_cast<size_t>(n) * T_len + t) * ldx;
      for (int c0 = lane * 8; c0 < ldx; c0 += 256)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(xs + c0)), "l"(src + c0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int c = threadIdx.x; c < Vp; c += 256) tab[c] = kCtcNoPos;
  if (built)
    for (int i = threadIdx.x; i < Sn; i += 256) labels[i] = static_cast<int>(tg[i]);
  __syncthreads();
  if (built)
    for (int i = threadIdx.x; i < Sn; i += 256) atomicMin(&tab[labels[i]], i);
  __syncthreads();
  if (built)
    for (int i = threadIdx.x; i < Sn; i += 256) rep[i] = tab[labels[i]];
  __syncthreads();
  if (t >= T_len) return;
  const size_t row = static_cast<size_t>(n) * T_len + t;
  GT* gr = grad + row * ldg;
  if (t >= Tn || !built) {  // frames past the utterance get a zero gradient row
    for (int c = lane; c < ldg; c += 32) gr[c] = from_f32<GT>(0.f);
    return;
  }
  const int Lp = 2 * Sn + 1;
  const int Lp_max = 2 * S_max + 1;
  const T* xr = x + row * ldx;
  if (STAGE) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    xr = xs;  // generic pointer into shared memory
  }
  const float l = lse ? lse[row] : 0.f;
  const float nl = nll[n];
  const float go = grad_out[n];
  for (int j = lane; j < Sn; j += 32) occ[j] = 0.f;
  __syncwarp();
  const float* ar = alpha + row * Lp_max;
  const float* br = beta + row * Lp_max;
  float bsum = 0.f;
  constexpr int GU = 4;
  for (int s0 = lane; s0 < Lp; s0 += 32 * GU) {
    float a[GU], b[GU], e[GU];
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const int st = s0 + 32 * k;
      a[k] = b[k] = e[k] = 0.f;
      if (st < Lp) {
        a[k] = ar[st];
        b[k] = br[st];
        e[k] = to_f32<T>(xr[(st & 1) ? labels[st >> 1] : blank]);
      }
    }
#pragma unroll
    for (int k = 0; k < GU; ++k) {
      const int st = s0 + 32 * k;
      if (st < Lp) {
        const float v = expf(a[k] + b[k] + nl - (e[k] - l));
        if (st & 1)
          atomicAdd(&occ[rep[st >> 1]], v);
        else
          bsum += v;
      }
    }
  }
  bsum = warp_sum(bsum);
  __syncwarp();
  auto occupancy = [&](int c) {
    const int r = tab[c];
    float o = r != kCtcNoPos ? occ[r] : 0.f;
    if (c == blank) o += bsum;
    return o;
  };
  constexpr bool kVec = sizeof(T) == 2 && sizeof(GT) == 2;
  if (kVec && (ldx & 7) == 0 && (ldg & 7) == 0 && ldg <= ldx && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(grad) & 15) == 0) {
    // bf16 in, bf16 out.  Only the <= S + 1 classes of the utterance's labels (and the blank) carry occupancy: the
    // row is first written as the plain softmax term, eight classes per lane and trip with no table lookup at all
    // (ex2.approx of a fused multiply-add: 2^-22, the output rounds to 2^-9), then the occupied classes are rewritten
    // one by one with the full formula.  The per-class table lookups of the first version were 40 instructions per
    // class: 152 M warp instructions, issue-bound at 3x the pass's HBM time.
    const float l2 = -l * 1.4426950408889634f;
    auto soft = [&](float xv) { return ex2_approx_ftz(fmaf(xv, 1.4426950408889634f, l2)); };
    const int full = V & ~7;  // classes below `full` sit in complete 8-class vectors
    for (int c0 = lane * 8; c0 < ldg; c0 += 256) {
      const uint4 raw = *reinterpret_cast<const uint4*>(xr + c0);
      const uint32_t wv[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t ov[4];
      if (c0 < full) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 xv = bf16x2_to_f32x2(wv[q]);
          ov[q] = f32x2_to_bf16x2(soft(xv.x) * go, soft(xv.y) * go);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 xv = bf16x2_to_f32x2(wv[q]);
          const int c = c0 + 2 * q;
          ov[q] = f32x2_to_bf16x2(c < V ? soft(xv.x) * go : 0.f, c + 1 < V ? soft(xv.y) * go : 0.f);
        }
      }
      *reinterpret_cast<uint4*>(gr + c0) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    }
    __syncwarp();  // the rewrites below touch addresses other lanes of this warp have just stored to
    for (int jj = lane; jj <= Sn; jj += 32) {
      int c;
      if (jj < Sn) {
        c = labels[jj];
        if (rep[jj] != jj || c == blank) continue;  // a later position of a repeated label / the blank: handled once
      } else {
        c = blank;
      }
      if (c < 0 || c >= V) continue;
      gr[c] = from_f32<GT>((soft(to_f32<T>(xr[c])) - occupancy(c)) * go);
    }
  } else {
    for (int c = lane; c < ldg; c += 32) {
      float g = 0.f;
      if (c < V) g = (expf(to_f32<T>(xr[c]) - l) - occupancy(c)) * go;
      gr[c] = from_f32<GT>(g);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// CTC lattices, second generation: linear-domain ("scaled") forward-backward, 8 warps per (utterance, direction).
//
// The recursion is a chain of T' dependent steps.  The round-1 kernel (ctc_lattice_kernel: one state per thread, log
// space) pays per step a block barrier over 13-19 warps, a shared-memory round trip and a 3-way log-sum-exp (two ex2 + one
// lg2 on the MUFU pipe): ~630 cycles per frame, 266 us for T' = 801 however few states there are.  Here a thread owns K
// CONSECUTIVE states in registers, the two predecessors of its first state come by warp shuffle (from the previous warp:
// through a double-buffered shared-memory slot), and the arithmetic is the classic SCALED forward-backward: probabilities
// stay linear and every 4th frame the whole column is multiplied by an exact power of two (so the scaling itself adds no
// rounding); the running exponent is an INTEGER per frame, alpha_t[s] = a[t][s] * 2^E[t].  A step is two adds and one
// multiply per state and ONE barrier over 8 warps: ~460 cycles per frame, 198 us.  (Measured and rejected: more states per
// thread on fewer warps (K = 4: 255 us), deferring the column's global store by a frame (212 us), and a barrier-free
// skewed pipeline of warps with per-warp exponents and progress counters in shared memory (388 us: the fences of the
// hand-off cost more than the barrier).)
//   stage 1  ctc_emit_kernel      e[n, t, j] = exp(x[n, t, label_j] - lse[n, t]), j < S_n; e[n, t, S_max] = blank
//                                 (one pass over the score matrix, any vocabulary size; the lattice reads rows of it)
//   stage 2  ctc_lattice2_kernel  alpha (with the frame's emission) and beta~ (WITHOUT it: what the gradient needs, so
//                                 nothing is divided by an emission later), int32 exponents per frame, P = sum of the
//                                 two final states as (mantissa, exponent), nll = -ln P
//   stage 3  ctc_grad_kernel      occupancy[c] += alpha * beta~ * 2^(EA + EB - EP) / P_mantissa, fused softmax gradient
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
ctc_emit_kernel(const T* __restrict__ x, const float* __restrict__ lse, const int64_t* __restrict__ targets,
                const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len, float* __restrict__ emis, int N,
                int T_len, int ldx, int S_max, int E_pad, int blank) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(N) * T_len) return;
  const int n = static_cast<int>(row / T_len);
  const int t = static_cast<int>(row - static_cast<long long>(n) * T_len);
  const int Sn = tgt_len[n];
  if (t >= in_len[n] || Sn < 0 || Sn > S_max) return;
  const T* xr = x + static_cast<size_t>(row) * ldx;
  const float l = lse != nullptr ? lse[row] : 0.f;
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  float* er = emis + static_cast<size_t>(row) * E_pad;
  for (int j = lane; j < E_pad; j += 32) {
    float v = 0.f;
    if (j < Sn)
      v = __expf(to_f32<T>(xr[static_cast<int>(tg[j])]) - l);
    else if (j == S_max)
      v = __expf(to_f32<T>(xr[blank]) - l);
    er[j] = v;
  }
}

constexpr int CTC2_DEPTH = 8;  // emission rows in flight (cp.async ring)

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}

constexpr int CTC2_WARPS = 8;
constexpr int CTC2_THREADS = 32 * CTC2_WARPS;

// One CTA of 8 warps per (utterance, direction); thread tid owns the K consecutive states q = tid * K + j in registers
// (alpha: s = q; beta: s = Lp - 1 - q, so that both directions read predecessors q - 1 and q - 2).  Per frame:
// neighbours by warp shuffle (the first lane(s) of a warp read the previous warp's last two states from a double-buffered
// shared slot), K x (2 adds, 1 multiply), the column staged in shared memory in global state order, ONE block barrier,
// then coalesced global stores that overlap the next frame.
template <int K, int DIR>
__device__ __forceinline__ void ctc_lattice2_body(const float* __restrict__ emis, const int64_t* __restrict__ targets,
                                                  const int32_t* __restrict__ in_len,
                                                  const int32_t* __restrict__ tgt_len, float* __restrict__ lat_out,
                                                  float* __restrict__ nll, int32_t* __restrict__ scales, int N, int T_len,
                                                  int S_max, int E_pad) {
  extern __shared__ __align__(16) float sm2[];
  // sm2: ring [CTC2_DEPTH][E_pad], rowbuf [2][CTC2_THREADS * K], bslot [2][CTC2_WARPS][2], mslot [CTC2_WARPS]
  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int Tn = in_len[n];
  const int Sn = tgt_len[n];
  const int Lp = 2 * Sn + 1;
  const int Lp_max = 2 * S_max + 1;
  const size_t NT = static_cast<size_t>(N) * T_len;
  int32_t* exps = scales + (static_cast<size_t>(DIR) * N + n) * T_len;  // exponent of the stored column of every frame
  if (Tn <= 0 || Sn < 0 || Sn > S_max || Tn > T_len) {
    if (DIR == 0 && tid == 0) {
      nll[n] = (Tn == 0 && Sn == 0) ? 0.f : CUDART_INF_F;
      scales[2 * NT + 2 * n] = 0;
      scales[2 * NT + 2 * n + 1] = 0;
    }
    return;
  }
  const int64_t* tg = targets + static_cast<size_t>(n) * S_max;
  constexpr int ROWF = CTC2_THREADS * K;
  const uint32_t ring_s = smem_u32(sm2);
  const uint32_t row_s = ring_s + 4u * CTC2_DEPTH * E_pad;
  const uint32_t bsl_s = row_s + 4u * 2 * ROWF;
  const uint32_t msl_s = bsl_s + 4u * 2 * CTC2_WARPS * 2;
  // per-state constants; invalid states (q >= Lp) read the always-zero emission cell and write a slot nobody reads
  uint32_t eoff[K], soff[K];
  uint32_t skip = 0u;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int q = tid * K + j;
    eoff[j] = 4u * (S_max + 1);
    soff[j] = 4u * q;
    if (q < Lp) {
      const int s = DIR ? Lp - 1 - q : q;
      eoff[j] = 4u * ((s & 1) ? (s >> 1) : S_max);
      soff[j] = 4u * s;
      if ((s & 1) && q >= 2) {
        const int s2 = DIR ? s + 2 : s - 2;
        if (tg[s >> 1] != tg[s2 >> 1]) skip |= 1u << j;
      }
    }
  }
  const int chunks = E_pad / 4;  // 16-byte pieces of an emission row (<= CTC2_THREADS, checked by the host)
  const long long estep = DIR ? -static_cast<long long>(E_pad) : static_cast<long long>(E_pad);
  const float* esrc = emis + (static_cast<size_t>(n) * T_len + (DIR ? Tn - 1 : 0)) * E_pad + 4 * tid;
  int pf = 0;  // local index of the next frame to prefetch
  auto prefetch = [&]() {
    if (pf < Tn && tid < chunks)
      cp_async_16(ring_s + 4u * static_cast<uint32_t>((pf & (CTC2_DEPTH - 1)) * E_pad) + 16u * tid, esrc);
    cp_async_commit();
    esrc += estep;
    ++pf;
  };
#pragma unroll 1
  for (int i = 0; i < CTC2_DEPTH - 1; ++i) prefetch();
  cp_async_wait<CTC2_DEPTH - 2>();  // frame 0 has landed
  __syncthreads();
  float a[K];
#pragma unroll
  for (int j = 0; j < K; ++j) a[j] = 0.f;
  int E = 0;
  const long long ostep = DIR ? -static_cast<long long>(Lp_max) : static_cast<long long>(Lp_max);
  float* orow = lat_out + (static_cast<size_t>(n) * T_len + (DIR ? Tn - 1 : 0)) * Lp_max + tid;
  int32_t* ep = exps + (DIR ? Tn - 1 : 0);
#pragma unroll 1
  for (int i = 0; i < Tn; ++i) {
    const uint32_t par = static_cast<uint32_t>(i & 1);
    const uint32_t er = ring_s + 4u * static_cast<uint32_t>((i & (CTC2_DEPTH - 1)) * E_pad);
    float e[K];
#pragma unroll
    for (int j = 0; j < K; ++j) e[j] = lds_f32(er + eoff[j]);
    // pending rescale (the warps' maxima of frame i - 1): an exact power of two, identical in every thread
    float sc = 1.f;
    if ((i & 3) == 0 && i > 0) {
      float m = lds_f32(msl_s);
#pragma unroll
      for (int w = 1; w < CTC2_WARPS; ++w) m = fmaxf(m, lds_f32(msl_s + 4u * w));
      if (m > 0.f) {
        int k = static_cast<int>((__float_as_uint(m) >> 23) & 0xffu) - 127;
        k = max(-100, min(100, k));
        sc = __uint_as_float(static_cast<uint32_t>(127 - k) << 23);
#pragma unroll
        for (int j = 0; j < K; ++j) a[j] *= sc;
        E += k;
      }
    }
    float st[K];  // what is stored for this frame: alpha (with the emission) / beta~ (without)
    if (i == 0) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int q = tid * K + j;
        const float one = (q < 2 && q < Lp) ? 1.f : 0.f;
        st[j] = one;
        a[j] = one * e[j];
      }
    } else {
      // predecessors of this thread's first state: q - 1 and q - 2 live in the previous thread(s)
      float in1 = __shfl_up_sync(0xffffffffu, a[K - 1], 1);
      float in2 = K >= 2 ? __shfl_up_sync(0xffffffffu, a[K >= 2 ? K - 2 : 0], 1) : __shfl_up_sync(0xffffffffu, a[0], 2);
      if (lane == 0 || (K == 1 && lane == 1)) {
        float b1 = 0.f, b2 = 0.f;
        if (warp > 0) {
          const uint32_t bs = bsl_s + 4u * (((par ^ 1u) * CTC2_WARPS + (warp - 1)) * 2);
          b1 = lds_f32(bs) * sc;       // the previous warp's last state, stored before this frame's rescale
          b2 = lds_f32(bs + 4u) * sc;  // and its second to last
        }
        if (lane == 0) {
          in1 = b1;
          in2 = b2;
        } else {
          in2 = b1;  // K == 1, lane 1: q - 2 is the previous warp's last state
        }
      }
#pragma unroll
      for (int j = K - 1; j >= 0; --j) {
        const float p1 = j >= 1 ? a[j - 1] : in1;
        const float p2 = j >= 2 ? a[j - 2] : (j == 1 ? in1 : in2);
        const float sum = a[j] + p1 + (((skip >> j) & 1u) ? p2 : 0.f);
        st[j] = sum;
        a[j] = sum * e[j];
      }
    }
    // the last two states of the warp, for the next warp's first lane(s) at the next frame
    if (K >= 2) {
      if (lane == 31) {
        const uint32_t bs = bsl_s + 4u * ((par * CTC2_WARPS + warp) * 2);
        sts_f32(bs, a[K - 1]);
        sts_f32(bs + 4u, a[K >= 2 ? K - 2 : 0]);
      }
    } else {
      if (lane >= 30) sts_f32(bsl_s + 4u * ((par * CTC2_WARPS + warp) * 2 + (31 - lane)), a[0]);
    }
    if ((i & 3) == 3) {
      float m = a[0];
#pragma unroll
      for (int j = 1; j < K; ++j) m = fmaxf(m, a[j]);
      m = warp_max(m);
      if (lane == 0) sts_f32(msl_s + 4u * warp, m);
    }
    const uint32_t rb = row_s + 4u * par * ROWF;
#pragma unroll
    for (int j = 0; j < K; ++j) sts_f32(rb + soff[j], DIR ? st[j] : a[j]);
    prefetch();
    cp_async_wait<CTC2_DEPTH - 2>();  // the emissions of frame i + 1 have landed (this thread's share)
    __syncthreads();
    float v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = lds_f32(rb + 4u * (tid + CTC2_THREADS * j));
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (tid + CTC2_THREADS * j < Lp) orow[CTC2_THREADS * j] = v[j];
    if (tid == 0) *ep = E;
    orow += ostep;
    ep += DIR ? -1 : 1;
  }
  if (DIR == 0 && tid == 0) {
    const uint32_t rb = row_s + 4u * static_cast<uint32_t>((Tn - 1) & 1) * ROWF;
    const float P = lds_f32(rb + 4u * (Lp - 1)) + (Lp >= 2 ? lds_f32(rb + 4u * (Lp - 2)) : 0.f);
    scales[2 * NT + 2 * n] = static_cast<int32_t>(__float_as_uint(P));
    scales[2 * NT + 2 * n + 1] = E;
    nll[n] = P > 0.f ? static_cast<float>(-(log(static_cast<double>(P)) + static_cast<double>(E) * 0.6931471805599453))
                     : CUDART_INF_F;
  }
}

// grid (N, 2): blockIdx.y = 0 the alpha lattice, 1 the beta~ lattice (beta == NULL: grid (N, 1)); they run concurrently
template <int K>
__global__ void __launch_bounds__(CTC2_THREADS)
ctc_lattice2_kernel(const float* __restrict__ emis, const int64_t* __restrict__ targets, const int32_t* __restrict__ in_len,
                    const int32_t* __restrict__ tgt_len, float* __restrict__ alpha, float* __restrict__ beta,
                    float* __restrict__ nll, int32_t* __restrict__ scales, int N, int T_len, int S_max, int E_pad) {
  pdl_launch_dependents();
  pdl_wait();
  if (blockIdx.y == 0)
    ctc_lattice2_body<K, 0>(emis, targets, in_len, tgt_len, alpha, nll, scales, N, T_len, S_max, E_pad);
  else
    ctc_lattice2_body<K, 1>(emis, targets, in_len, tgt_len, beta, nll, scales, N, T_len, S_max, E_pad);
}

// ------------------------------------------------------------------------------------------------
// greedy decode
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
argmax_kernel(const T* __restrict__ x, int64_t* __restrict__ amax, long long M, int V, int ldx) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const T* xr = x + static_cast<size_t>(row) * ldx;
  float best = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int c = lane; c < V; c += 32) {
    const float v = to_f32<T>(xr[c]);
    if (v > best || (bi == 0x7fffffff)) {  // strictly greater keeps the lowest index within a lane
      best = v;
      bi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) amax[row] = bi;
}

// one warp per utterance: keep p iff (p != prev || prev == blank) && p != blank, prev = previous FRAME's argmax
__global__ void __launch_bounds__(32)
ctc_collapse_kernel(const int64_t* __restrict__ amax, const int32_t* __restrict__ lengths, int32_t* __restrict__ tokens,
                    int32_t* __restrict__ counts, int T_len, int blank) {
  const int n = blockIdx.x, lane = threadIdx.x;
  int len = lengths ? lengths[n] : T_len;
  len = min(max(len, 0), T_len);
  const int64_t* a = amax + static_cast<size_t>(n) * T_len;
  int32_t* out = tokens + static_cast<size_t>(n) * T_len;
  int count = 0;
  int carry = blank;  // argmax of the frame before this chunk
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    const int p = t < len ? static_cast<int>(a[t]) : blank;
    int prev = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) prev = carry;
    const bool keep = t < len && (p != prev || prev == blank) && p != blank;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) out[count + __popc(m & ((1u << lane) - 1u))] = p;
    count += __popc(m);
    carry = __shfl_sync(0xffffffffu, p, 31);
  }
  if (lane == 0) counts[n] = count;
}

// column sums of a [M, ld] matrix (decoder bias gradient = sum over frames of d(logits)): out[c] += sum_m x[m, c]
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long M, int C, int ld, int cols_par,
              int rows_per_block) {
  __shared__ float red[256];
  const int rows_par = 256 / cols_par;
  const int cc = threadIdx.x % cols_par, rr = threadIdx.x / cols_par;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  for (int c0 = blockIdx.y * cols_par; c0 < C; c0 += gridDim.y * cols_par) {
    const int c = c0 + cc;
    float acc = 0.f;
    if (c < C)
      for (long long r = r0 + rr; r < r1; r += rows_par) acc += to_f32<T>(x[static_cast<size_t>(r) * ld + c]);
    red[threadIdx.x] = acc;
    __syncthreads();
    if (rr == 0 && c < C) {
      for (int j = 1; j < rows_par; ++j) acc += red[j * cols_par + cc];
      atomicAdd(out + c, acc);
    }
    __syncthreads();
  }
}

// bf16 rows whose pitch is a multiple of 8: a thread owns 8 consecutive columns (one 16-byte load per row, a warp reads
// 512 contiguous bytes), 8 row lanes per block, four rows per lane in flight.  (The scalar kernel above reads 2 bytes
// per thread and row with one dependent add after the other: 0.33 ms for the [25 632, 4336] gradient of the AISHELL
// decoder, whose 222 MB take 34 us at HBM speed.)
__global__ void __launch_bounds__(256)
colsum_vec8_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long M, int C, int ld,
                   int rows_per_block) {
  __shared__ float red[8][32][9];
  const int vl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int cv = blockIdx.y * 32 + vl;  // column vector
  const int c0 = cv * 8;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c0 < C) {
    const __nv_bfloat16* base = x + c0;
    for (long long r = r0 + rl; r < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long rr = r + 8 * k;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (rr < r1) v[k] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(rr) * ld);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = bf16x2_to_f32x2(v[k].x), b = bf16x2_to_f32x2(v[k].y), c = bf16x2_to_f32x2(v[k].z),
                     d = bf16x2_to_f32x2(v[k].w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][vl][i] = acc[i];
  __syncthreads();
  // 256 threads = the block's 256 columns: sum the 8 row lanes
  const int col = threadIdx.x, cvl = col >> 3, ci = col & 7;
  const int c = blockIdx.y * 256 + col;
  if (c < C) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += red[j][cvl][ci];
    atomicAdd(out + c, sum);
  }
}

template <typename T, typename GT>
static int ctc_grad_launch(const void* x, const float* lse, const int64_t* targets, const int32_t* il,
                           const int32_t* tl, const float* alpha, const float* beta, const float* nll,
                           const int32_t* scales, const float* grad_out, void* grad, int N, int Tn, int V, int ldx,
                           int ldg, int S_max, int blank, cudaStream_t stream) {
  if (scales == nullptr && V <= 128 && !(getenv("LASR_CTC_GRAD_SMALL") != nullptr && atoi(getenv("LASR_CTC_GRAD_SMALL")) == 0)) {
    const int S_pad = (S_max + 3) / 4 * 4, Vp = (V + 31) / 32 * 32;
    const int smem_small = (S_pad + 8 * 2 * Vp) * static_cast<int>(sizeof(float));
    if (smem_small <= 48 * 1024) {
      LASR_CHECK_PDL(launch_pdl(8, ctc_grad_small_kernel<T, GT>, dim3(cdiv(Tn, 8), N), dim3(256), smem_small, stream,
                                static_cast<const T*>(x), lse, targets, il, tl, alpha, beta, nll, grad_out,
                                static_cast<GT*>(grad), Tn, V, ldx, ldg, S_max, blank, S_pad, Vp));
      return LASR_OK;
    }
  }
  if (scales == nullptr && V > 128 && !(getenv("LASR_CTC_GRAD_LARGE") != nullptr && atoi(getenv("LASR_CTC_GRAD_LARGE")) == 0)) {
    const int S_pad = (S_max + 3) / 4 * 4, Vp = (V + 3) / 4 * 4;
    const int smem_large = (2 * S_pad + Vp + 8 * S_pad) * static_cast<int>(sizeof(float));
    // bf16 -> bf16 with vector-friendly rows: the rows of a CTA staged in shared memory (two CTAs per SM)
    const int smem_stage = smem_large + 8 * ldx * static_cast<int>(sizeof(T));
    const bool stage = sizeof(T) == 2 && sizeof(GT) == 2 && (ldx & 7) == 0 && (ldg & 7) == 0 && ldg <= ldx &&
                       (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                       smem_stage <= 110 * 1024 &&
                       !(getenv("LASR_CTC_GRAD_STAGE") != nullptr && atoi(getenv("LASR_CTC_GRAD_STAGE")) == 0);
    if (stage) {
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ctc_grad_large_kernel<T, GT, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        if (e != cudaSuccess) {
          lasr_set_cuda_error(e);
          return LASR_ERR_CUDA;
        }
        configured = true;
      }
      LASR_CHECK_PDL(launch_pdl(8, ctc_grad_large_kernel<T, GT, true>, dim3(cdiv(Tn, 8), N), dim3(256), smem_stage,
                                stream, static_cast<const T*>(x), lse, targets, il, tl, alpha, beta, nll, grad_out,
                                static_cast<GT*>(grad), Tn, V, ldx, ldg, S_max, blank, S_pad, Vp));
      return LASR_OK;
    }
    if (smem_large <= 48 * 1024) {
      LASR_CHECK_PDL(launch_pdl(8, ctc_grad_large_kernel<T, GT, false>, dim3(cdiv(Tn, 8), N), dim3(256), smem_large,
                                stream, static_cast<const T*>(x), lse, targets, il, tl, alpha, beta, nll, grad_out,
                                static_cast<GT*>(grad), Tn, V, ldx, ldg, S_max, blank, S_pad, Vp));
      return LASR_OK;
    }
  }
  int warps = (96 * 1024) / (V * 4);
  if (warps > 8) warps = 8;
  if (warps < 1) warps = 1;
  const int smem = warps * V * static_cast<int>(sizeof(float));
  if (smem > 200 * 1024) return LASR_ERR_UNSUPPORTED;
  static int configured_smem = 0;
  if (smem > 48 * 1024 && smem > configured_smem) {
    cudaError_t e =
        cudaFuncSetAttribute(ctc_grad_kernel<T, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured_smem = 200 * 1024;
  }
  const long long rows = static_cast<long long>(N) * Tn;
  const int grid = static_cast<int>((rows + warps - 1) / warps);
  LASR_CHECK_PDL(launch_pdl(8, ctc_grad_kernel<T, GT>, dim3(grid), dim3(256), smem, stream, static_cast<const T*>(x), lse,
                            targets, il, tl, alpha, beta, nll, scales, grad_out, static_cast<GT*>(grad), N, Tn, V, ldx,
                            ldg, S_max, blank, warps));
  return LASR_OK;
}

// second-generation lattices: emission gather + one 8-warp CTA per (utterance, direction)
template <int K>
static int ctc_lattice2_launch(const float* emis, const int64_t* targets, const int32_t* il, const int32_t* tl,
                               float* alpha, float* beta, float* nll, int32_t* scales, int N, int T_len, int S_max,
                               int E_pad, cudaStream_t stream) {
  const int smem = (CTC2_DEPTH * E_pad + 2 * CTC2_THREADS * K + 2 * CTC2_WARPS * 2 + CTC2_WARPS) *
                   static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) return LASR_ERR_UNSUPPORTED;
  LASR_CHECK_PDL(launch_pdl(8, ctc_lattice2_kernel<K>, dim3(N, beta != nullptr ? 2 : 1), dim3(CTC2_THREADS), smem, stream,
                            emis, targets, il, tl, alpha, beta, nll, scales, N, T_len, S_max, E_pad));
  return LASR_OK;
}

template <typename T>
static int ctc_fwd2(const void* x, const float* lse, const int64_t* targets, const int32_t* il, const int32_t* tl,
                    float* alpha, float* beta, float* nll, int32_t* scales, float* emis, int N, int T_len, int ldx,
                    int S_max, int blank, cudaStream_t stream) {
  const int E_pad = (S_max + 2 + 3) / 4 * 4;  // labels, blank at S_max, an always-zero cell at S_max + 1
  if (E_pad / 4 > CTC2_THREADS) return LASR_ERR_UNSUPPORTED;
  const long long rows = static_cast<long long>(N) * T_len;
  LASR_CHECK_PDL(launch_pdl(8, ctc_emit_kernel<T>, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, stream,
                            static_cast<const T*>(x), lse, targets, il, tl, emis, N, T_len, ldx, S_max, E_pad, blank));
  const int Lp_max = 2 * S_max + 1;
  const int k = (Lp_max + CTC2_THREADS - 1) / CTC2_THREADS;
#define LASR_CTC2(KK) \
  return ctc_lattice2_launch<KK>(emis, targets, il, tl, alpha, beta, nll, scales, N, T_len, S_max, E_pad, stream)
  if (k <= 1) LASR_CTC2(1);
  if (k <= 2) LASR_CTC2(2);
  if (k <= 3) LASR_CTC2(3);
  if (k <= 4) LASR_CTC2(4);
  if (k <= 6) LASR_CTC2(6);
  if (k <= 8) LASR_CTC2(8);
#undef LASR_CTC2
  return LASR_ERR_UNSUPPORTED;
}

template <typename T, int SPT, int CTC_THREADS>
static int ctc_lattice_launch(const void* x, const float* lse, const int64_t* targets, const int32_t* il,
                              const int32_t* tl, float* alpha, float* beta, float* nll, int N, int T_len, int ldx,
                              int S_max, int blank, cudaStream_t stream) {
  const int Lp_max = 2 * S_max + 1;
  // SPT == 1: exactly as many warps as the widest lattice needs (every warp pays the per-frame barrier)
  int threads = SPT == 1 ? cdiv(Lp_max, 32) * 32 : CTC_THREADS;
  if (threads > CTC_THREADS) threads = CTC_THREADS;
  const int smem = (2 * (Lp_max + 4) + CTC_RING * SPT * threads + CTC_RING) * static_cast<int>(sizeof(float));
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_lattice_kernel<T, SPT, CTC_THREADS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (2 * (2 * 1024 + 1 + 4) + CTC_RING * SPT * CTC_THREADS + CTC_RING) * 4);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(N, beta != nullptr ? 2 : 1);
  LASR_CHECK_PDL(launch_pdl(8, ctc_lattice_kernel<T, SPT, CTC_THREADS>, grid, dim3(threads), smem, stream,
                            static_cast<const T*>(x), lse, targets, il, tl, alpha, beta, nll, T_len, ldx, S_max, blank));
  return LASR_OK;
}

// LASR_CTC_WARP=0 selects the round-1 barrier kernel (A/B runs, the bit-identity test)
template <typename T, bool LSE>
static int ctc_warp_launch(const void* x, const float* lse, const int64_t* targets, const int32_t* il, const int32_t* tl,
                           float* alpha, float* beta, float* nll, int N, int T_len, int ldx, int S_max, int blank,
                           cudaStream_t stream) {
  const int Lp_max = 2 * S_max + 1;
  const dim3 grid(N, beta != nullptr ? 2 : 1);
  const T* xx = static_cast<const T*>(x);
  const int threads = cdiv(Lp_max, 64) * 32;
  if (threads > 1024) return LASR_ERR_UNSUPPORTED;
  const size_t smem = 8u * CTCW_SLOTS * (threads / 32) + 8u * threads;
  if (threads <= 512)
    LASR_CHECK_PDL(launch_pdl(8, ctc_lattice_warp2_kernel<T, LSE, 512>, grid, dim3(threads), smem, stream, xx, lse,
                              targets, il, tl, alpha, beta, nll, T_len, ldx, S_max, blank));
  else
    LASR_CHECK_PDL(launch_pdl(8, ctc_lattice_warp2_kernel<T, LSE, 1024>, grid, dim3(threads), smem, stream, xx, lse,
                              targets, il, tl, alpha, beta, nll, T_len, ldx, S_max, blank));
  return LASR_OK;
}

template <typename T>
static int ctc_warp_dispatch(const void* x, const float* lse, const int64_t* targets, const int32_t* il,
                             const int32_t* tl, float* alpha, float* beta, float* nll, int N, int T_len, int ldx,
                             int S_max, int blank, cudaStream_t stream) {
  const char* env = getenv("LASR_CTC_WARP");
  if (env != nullptr && atoi(env) == 0) return LASR_ERR_UNSUPPORTED;
  if (lse != nullptr)
    return ctc_warp_launch<T, true>(x, lse, targets, il, tl, alpha, beta, nll, N, T_len, ldx, S_max, blank, stream);
  return ctc_warp_launch<T, false>(x, lse, targets, il, tl, alpha, beta, nll, N, T_len, ldx, S_max, blank, stream);
}

template <typename T>
static int ctc_lattice_dispatch(const void* x, const float* lse, const int64_t* targets, const int32_t* il,
                                const int32_t* tl, float* alpha, float* beta, float* nll, int N, int T_len, int ldx,
                                int S_max, int blank, cudaStream_t stream) {
  const int Lp_max = 2 * S_max + 1;
  {
    const int rc = ctc_warp_dispatch<T>(x, lse, targets, il, tl, alpha, beta, nll, N, T_len, ldx, S_max, blank, stream);
    if (rc != LASR_ERR_UNSUPPORTED) return rc;
  }
#define LASR_CTC_LAT(SPT, TH) \
  return ctc_lattice_launch<T, SPT, TH>(x, lse, targets, il, tl, alpha, beta, nll, N, T_len, ldx, S_max, blank, stream)
  if (Lp_max <= 128) LASR_CTC_LAT(1, 128);
  if (Lp_max <= 256) LASR_CTC_LAT(1, 256);
  if (Lp_max <= 512) LASR_CTC_LAT(1, 512);
  if (Lp_max <= 1024) LASR_CTC_LAT(1, 1024);
  if (Lp_max <= 2048) LASR_CTC_LAT(2, 1024);
  return LASR_ERR_UNSUPPORTED;
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_log_softmax_fwd(const void* logits, float* lse, float* lp, int M, int V, int ld, int dtype,
                         lasr_stream_t stream) {
  if (M <= 0 || V <= 0 || ld < V) return LASR_ERR_BAD_SHAPE;
  const int grid = cdiv(M, 8);
  if (dtype == LASR_F32)
    log_softmax_fwd_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(logits), lse, lp, M, V, ld);
  else if (dtype == LASR_BF16)
    log_softmax_fwd_kernel<__nv_bfloat16>
        <<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), lse, lp, M, V, ld);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_log_softmax_bwd(const float* dlp, const float* lp, void* dlogits, int M, int V, int ld, int dtype,
                         lasr_stream_t stream) {
  if (M <= 0 || V <= 0 || ld < V) return LASR_ERR_BAD_SHAPE;
  const int grid = cdiv(M, 8);
  if (dtype == LASR_F32)
    log_softmax_bwd_kernel<float><<<grid, 256, 0, stream>>>(dlp, lp, static_cast<float*>(dlogits), M, V, ld);
  else if (dtype == LASR_BF16)
    log_softmax_bwd_kernel<__nv_bfloat16>
        <<<grid, 256, 0, stream>>>(dlp, lp, static_cast<__nv_bfloat16*>(dlogits), M, V, ld);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

size_t lasr_ctc_scales_bytes(int N, int T) {
  return (2 * static_cast<size_t>(N) * T + 2 * static_cast<size_t>(N)) * sizeof(int32_t);
}
size_t lasr_ctc_emis_bytes(int N, int T, int S_max) {
  return static_cast<size_t>(N) * T * ((S_max + 2 + 3) / 4 * 4) * sizeof(float);
}

int lasr_ctc_fwd(const void* x, const float* lse, const int64_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha, float* beta, float* nll, int32_t* scales, float* emis,
                 int N, int T, int V, int ldx, int S_max, int blank, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || V <= 0 || ldx < V || S_max < 0 || blank < 0 || blank >= V) return LASR_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return LASR_ERR_ALIGNMENT;
  if ((scales != nullptr) != (emis != nullptr)) return LASR_ERR_BAD_SHAPE;
  if (scales != nullptr && 2 * S_max + 1 <= 32 * 33) {
    if (dtype == LASR_F32)
      return ctc_fwd2<float>(x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, scales, emis, N, T, ldx,
                             S_max, blank, stream);
    if (dtype == LASR_BF16)
      return ctc_fwd2<__nv_bfloat16>(x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, scales, emis, N,
                                     T, ldx, S_max, blank, stream);
    return LASR_ERR_BAD_DTYPE;
  }
  if (scales != nullptr) return LASR_ERR_UNSUPPORTED;
  if (dtype == LASR_F32)
    return ctc_lattice_dispatch<float>(x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, N, T, ldx,
                                       S_max, blank, stream);
  if (dtype == LASR_BF16) {
    if (ldx & 1) return LASR_ERR_ALIGNMENT;  // rows must start on a 4-byte word (the emission gather reads words)
    return ctc_lattice_dispatch<__nv_bfloat16>(x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, N, T,
                                               ldx, S_max, blank, stream);
  }
  return LASR_ERR_BAD_DTYPE;
}

int lasr_ctc_bwd(const void* x, const float* lse, const int64_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, const float* alpha, const float* beta, const float* nll,
                 const int32_t* scales, const float* grad_out, void* grad, int N, int T, int V, int ldx, int ldg,
                 int S_max, int blank, int dtype, int grad_dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || V <= 0 || ldx < V || ldg < V || S_max < 0) return LASR_ERR_BAD_SHAPE;
  if (alpha == nullptr || beta == nullptr) return LASR_ERR_BAD_SHAPE;
#define LASR_CTC_ARGS                                                                                                \
  x, lse, targets, input_lengths, target_lengths, alpha, beta, nll, scales, grad_out, grad, N, T, V, ldx, ldg, S_max, \
      blank, stream
  if (dtype == LASR_F32 && grad_dtype == LASR_F32) return ctc_grad_launch<float, float>(LASR_CTC_ARGS);
  if (dtype == LASR_BF16 && grad_dtype == LASR_BF16)
    return ctc_grad_launch<__nv_bfloat16, __nv_bfloat16>(LASR_CTC_ARGS);
  if (dtype == LASR_BF16 && grad_dtype == LASR_F32) return ctc_grad_launch<__nv_bfloat16, float>(LASR_CTC_ARGS);
  if (dtype == LASR_F32 && grad_dtype == LASR_BF16) return ctc_grad_launch<float, __nv_bfloat16>(LASR_CTC_ARGS);
  return LASR_ERR_BAD_DTYPE;
}

int lasr_greedy_decode(const void* x, const int32_t* lengths, int64_t* argmax, int32_t* tokens, int32_t* counts, int N,
                       int T, int V, int ldx, int blank, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || V <= 0 || ldx < V || argmax == nullptr) return LASR_ERR_BAD_SHAPE;
  const long long M = static_cast<long long>(N) * T;
  const int grid = static_cast<int>((M + 7) / 8);
  if (dtype == LASR_F32)
    argmax_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), argmax, M, V, ldx);
  else if (dtype == LASR_BF16)
    argmax_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), argmax, M, V, ldx);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  if (tokens != nullptr) {
    ctc_collapse_kernel<<<N, 32, 0, stream>>>(argmax, lengths, tokens, counts, T, blank);
    LASR_CHECK_LAUNCH();
  }
  return LASR_OK;
}

/* collapse only: predictions [N, T] int64 (e.g. an argmax computed elsewhere) -> tokens / counts */
int lasr_ctc_collapse(const int64_t* predictions, const int32_t* lengths, int32_t* tokens, int32_t* counts, int N,
                      int T, int blank, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || predictions == nullptr || tokens == nullptr || counts == nullptr) return LASR_ERR_BAD_SHAPE;
  ctc_collapse_kernel<<<N, 32, 0, stream>>>(predictions, lengths, tokens, counts, T, blank);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

/* out[c] += sum_m x[m, c] for c < C (out fp32, caller zeroes): the decoder bias gradient */
int lasr_colsum(const void* x, float* out, int M, int C, int ld, int dtype, lasr_stream_t stream) {
  if (M <= 0 || C <= 0 || ld < C) return LASR_ERR_BAD_SHAPE;
  if (dtype == LASR_BF16 && C >= 256 && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // (columns C .. ld of a padded row are summed into nothing: the guard is on c0 < C per vector, c < C per column)
    const int gy = cdiv(C, 256);
    int gxv = (4 * kNumSMs) / gy;
    if (gxv < 1) gxv = 1;
    int rpb = cdiv(M, gxv);
    if (rpb < 64) rpb = 64;
    gxv = cdiv(M, rpb);
    colsum_vec8_kernel<<<dim3(gxv, gy), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), out, M, C, ld, rpb);
    LASR_CHECK_LAUNCH();
    return LASR_OK;
  }
  int cols_par = 32;
  while (cols_par < C && cols_par < 256) cols_par *= 2;
  const int col_groups = cdiv(C, cols_par);
  int gy = col_groups < 8 ? col_groups : 8;
  int gx = (2 * kNumSMs) / gy;
  if (gx < 1) gx = 1;
  int rows_per_block = cdiv(M, gx);
  if (rows_per_block < 64) rows_per_block = 64;
  gx = cdiv(M, rows_per_block);
  dim3 grid(gx, gy);
  if (dtype == LASR_F32)
    colsum_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), out, M, C, ld, cols_par,
                                                   rows_per_block);
  else if (dtype == LASR_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), out, M, C, ld,
                                                           cols_par, rows_per_block);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // extern "C"
