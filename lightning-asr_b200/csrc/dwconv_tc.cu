// Depthwise Conv1d on the tensor cores (bf16, stride 1): forward, data gradient (flipped taps, fused residual addend)
// and weight gradient.  Replaces nn.Conv1d(C, C, k, groups=C) at models/QuartNet.py:14-21,30 and its autograd backward.
//
// Why: with k = 33..87 taps the op costs k FMAs per element -- on the fp32 pipe (dwconv.cu, packed FFMA2) that is
// 22 us for a 512-channel layer at 100 % pipe utilisation against 8 us of HBM time.  A depthwise conv is a banded
// Toeplitz product per channel, y_c = Toep(w_c) z_c, and tcgen05 can consume BOTH factors without materialising them:
//
//   * the channel's time series z_c (bf16, frames contiguous) is a legal K-major, no-swizzle A operand of the HANKEL
//     matrix A[m, k] = z_c[8 m + k]: canonical layout ((8,m),(8,2)) : ((16 B, SBO), (2 B, LBO)) with SBO = 128 B and
//     LBO = 16 B -- the core matrices overlap, row m simply starts 8 frames after row m-1.  128 rows x 8 frames =
//     1024 output frames per MMA, K-steps advance the start address by 32 B (16 frames);
//   * the Toeplitz factor B[t, s] = w_c[s - t] (t < 8 output frames of a row, s < KS = 16*ceil((k+7)/16) inputs) is
//     KS/8 core matrices of 128 B per channel, built once per CTA.
//   D[m, t] = sum_s z_c[8 m + s] w_c[s - t] = y_c[8 m + t].  N = 16 is the smallest legal N at M = 128, columns 8..15
//   are ignored.  One MMA (K = 16) costs the 4 KB shared-memory read of A (32 cycles): 6-10 outputs / cycle / SM, i.e.
//   about the HBM time of the layer, against 1.5 outputs / cycle for FFMA2.
//
// Activations are channels-last, so the per-channel series do not exist in memory: producer warps load 8 frames x 8
// channels per thread (16-byte loads, a 16-lane group covers both halves of every 32-byte sector), transpose the 8x8
// bf16 block with byte permutes and store 16-byte chunks of 8 consecutive frames into the channel's series.  Out of
// range frames (the conv zero padding, utterance edges) are stored as zeros.
//
// CTA = 16 channels, persistent over (utterance, 1024-frame chunk) items; warps 0-3 epilogue (TMEM -> bf16 -> 32-byte
// global stores, + addend), warp 4 MMA issuer, warp 5 TMEM allocator, warps 6-9 producers; series and accumulators
// double buffered, so load/transpose, MMA and store of consecutive items overlap.
//
// wgrad: dw_c[j] = sum_tau dy_c[tau] z_c[tau + j].  With windows of 8 frames in the K dimension both operands are again
// plain series: A[t', w] = z_c[8 w + t'] and B[w, t] = dy_c[8 w + t] are MN-major no-swizzle operands (LBO = 128 B
// between groups of 8 windows, SBO = 16 B between groups of 8 t'), D[t', t] += sum_w z_c[8w + t'] dy_c[8w + t] and
// dw_c[j] = sum_{t<8} D[t + j, t].  The accumulator stays in TMEM over all items of the CTA; one RED pass at the end.
#include "dw_common.cuh"

#include <cstdlib>

namespace lasr {

constexpr int DT_CHUNK = 1024;   // output frames per item
constexpr int DT_PROD_WARPS = 8;  // two groups of 4: group g fills series stage g
constexpr int DT_THREADS = 32 * (6 + DT_PROD_WARPS);
constexpr int DT_STAGES = 4;     // series buffers of the forward kernel

struct DwTcParams {
  const __nv_bfloat16* x;
  const float* w;
  __nv_bfloat16* y;
  const __nv_bfloat16* addend;
  const __nv_bfloat16* dy;  // wgrad only
  float* dw;                // wgrad only
  int N, T, C, K, KS, flip;
  int num_cg, t_chunks, items_per_cg, ctas_per_cg;
  int ZL;  // series length in frames (multiple of 8)
  unsigned long long* trace;  // debug timeline (tools/trace_dw.py), normally NULL
  int exp;                    // debug: descriptor experiments (timing only, results are garbage when != 0)
  int w_early;                // taps may be read before griddepcontrol.wait (lasr_set_early_param_loads)
};

__device__ __forceinline__ unsigned long long dt_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DT_TRACE(slot, idx)                                                                          \
  do {                                                                                               \
    if (p.trace != nullptr && (idx) < 16) p.trace[(blockIdx.x * 8 + (slot)) * 16 + (idx)] = dt_gtimer(); \
  } while (0)
static unsigned long long* g_dt_trace = nullptr;
extern "C" void lasr_debug_set_dw_trace(unsigned long long* buf) { g_dt_trace = buf; }
unsigned long long* dw_trace_buffer() { return g_dt_trace; }



// global [frames, C] -> per-channel series in shared memory.  series[c][sigma] (bf16, pitch ZL) for sigma in
// [0, ZL): frame f = f_base + sigma of utterance rows `src` (out of [0, T): zero).  Called by the 4 warps of a producer
// group (pw = 0..3); a warp handles 128 frames x 16 channels per pass and keeps the loads of ALL its passes (<= 3,
// ZL <= 1536) in flight before transposing: the step is latency-bound otherwise.
__device__ __forceinline__ void load_series(const __nv_bfloat16* __restrict__ src, int C, int T, int f_base, int ZL,
                                            uint8_t* series, int pw, int lane) {
  const int h = lane >> 4;    // channel half: channels 8h .. 8h+7
  const int b = lane & 15;    // 8-frame block within the pass
  uint4 r[3][8];
#pragma unroll
  for (int ps = 0; ps < 3; ++ps) {
    const int sigma = (pw + 4 * ps) * 128 + 8 * b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = f_base + sigma + i;
      r[ps][i] = make_uint4(0u, 0u, 0u, 0u);
      if (sigma < ZL && f >= 0 && f < T)
        r[ps][i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(f) * C + 8 * h));
    }
  }
#pragma unroll
  for (int ps = 0; ps < 3; ++ps) {
    const int sigma = (pw + 4 * ps) * 128 + 8 * b;
    if (sigma >= ZL) continue;
    // 8 frames x 8 channels -> 8 channels x 8 frames
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t o[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const uint32_t a = (&r[ps][2 * m].x)[q >> 1], bb = (&r[ps][2 * m + 1].x)[q >> 1];
        o[m] = __byte_perm(a, bb, (q & 1) ? 0x7632 : 0x5410);
      }
      *reinterpret_cast<uint4*>(series + (static_cast<size_t>(8 * h + q) * ZL + sigma) * 2) =
          make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

__device__ __forceinline__ void dwconv_tc_fwd_body(const DwTcParams& p, const int bid) {
  pdl_launch_dependents();  // the Toeplitz build / TMEM allocation below overlap the previous kernel's tail; taps are
                            // parameters (written behind a full stream barrier), see common.cuh
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  const int cores = p.KS / 8;
  const int series_bytes = DT_CG * p.ZL * 2;
  uint8_t* s_toep = smem;                                   // [16][cores][128 B]
  uint8_t* s_ser = s_toep + DT_CG * cores * 128;            // [DT_STAGES][16][ZL] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ser + DT_STAGES * series_bytes);
  uint64_t* full_bar = bars;                         // [DT_STAGES] producers -> MMA
  uint64_t* empty_bar = bars + DT_STAGES;            // [DT_STAGES] MMA -> producers
  uint64_t* tmem_full_bar = bars + 2 * DT_STAGES;    // [2] MMA -> epilogue
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] epilogue -> MMA
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;
  const int P = p.K / 2;

  if (!p.w_early) pdl_wait();
  // Toeplitz cores: core kk, row r (= output frame t of the window), element e: w[8 kk + e - r]
  for (int i = threadIdx.x; i < DT_CG * cores * 64; i += DT_THREADS) {
    const int c = i / (cores * 64);
    const int rem = i - c * cores * 64;
    const int kk = rem >> 6, r = (rem >> 3) & 7, e = rem & 7;
    const int j = 8 * kk + e - r;
    float v = 0.f;
    if (j >= 0 && j < p.K && c0 + c < p.C) v = p.w[static_cast<size_t>(c0 + c) * p.K + (p.flip ? p.K - 1 - j : j)];
    reinterpret_cast<__nv_bfloat16*>(s_toep)[i] = __float2bfloat16_rn(v);
  }
  if (warp_idx == 4 && lane == 0) {
    for (int s = 0; s < DT_STAGES; ++s) {
      mbar_init(&full_bar[s], 4);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    mbar_fence_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // Toeplitz cores were written through the generic proxy, the MMA reads them asynchronously
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp_idx >= 6) {
    // ===================== producers: two groups of 4 warps take alternate items; with DT_STAGES = 4 series
    // buffers a group loads item it+2 while the MMAs of item it+1 (the other group's) are still running ==========
    const int pw = (warp_idx - 6) & 3;
    const int grp = (warp_idx - 6) >> 2;
    int it = grp;
    for (int idx = first + grp * p.ctas_per_cg; idx < p.items_per_cg; idx += 2 * p.ctas_per_cg, it += 2) {
      const int stage = it % DT_STAGES;
      const uint32_t phase = (it / DT_STAGES) & 1;
      const int n = idx / p.t_chunks, tc = idx - n * p.t_chunks;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (pw == 0 && lane == 0) DT_TRACE(0, it);
      load_series(p.x + static_cast<size_t>(n) * p.T * p.C + c0, p.C, p.T, tc * DT_CHUNK - P, p.ZL,
                  s_ser + stage * series_bytes, pw, lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (pw == 0 && lane == 0) DT_TRACE(1, it);
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp_idx == 4) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 0, 0);
      const int ksteps = p.KS / 16;
      int it = 0;
      for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
        const int stage = it % DT_STAGES;
        const uint32_t phase = (it / DT_STAGES) & 1;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        if (leader) DT_TRACE(2, it);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) DT_TRACE(3, it);
        // The issuing thread, not the tensor pipe, paces these small MMAs (N = 16): keep the loop to two 64-bit adds
        // per instruction -- descriptors advance by 32 B (A: 16 frames) and 256 B (B: 2 cores) in their 16-byte address
        // field.
        uint64_t da_c = umma_desc_none(smem_u32(s_ser + stage * series_bytes), 16, 128);
        uint64_t db_c = umma_desc_none(smem_u32(s_toep), 128, 128);
        uint32_t tmem_d = tmem_base + acc * 256;
        const uint64_t da_step = static_cast<uint64_t>(p.ZL * 2 >> 4), db_step = static_cast<uint64_t>(cores * 128 >> 4);
#pragma unroll 1
        for (int c = 0; c < DT_CG; ++c) {
          uint64_t da = da_c, db = db_c;
          if (leader) umma_bf16_first(tmem_d, da, db, idesc);
#pragma unroll 1
          for (int kc = 1; kc < ksteps; ++kc) {
            da += 2;
            db += 16;
            if (leader) umma_bf16_acc(tmem_d, da, db, idesc);
          }
          da_c += da_step;
          db_c += db_step;
          tmem_d += 16;
        }
        if (leader) {
          umma_commit(&empty_bar[stage]);
          umma_commit(&tmem_full_bar[acc]);
          DT_TRACE(4, it);
        }
        __syncwarp();
      }
    }
  } else if (warp_idx < 4) {
    // ===================== epilogue =====================
    const int row = warp_idx * 32 + lane;  // window: output frames 8 row .. 8 row + 7 of the chunk
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      const int n = idx / p.t_chunks, tc = idx - n * p.t_chunks;
      const int f0 = tc * DT_CHUNK + 8 * row;
      const size_t off0 = (static_cast<size_t>(n) * p.T + f0) * p.C + c0;
      // the residual-branch gradient this kernel adds (dgrad): fetched before waiting on the accumulator
      uint32_t add[2][4][8];
      if (p.addend != nullptr) {
#pragma unroll
        for (int th = 0; th < 2; ++th)
#pragma unroll
          for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int c = 0; c < 8; ++c) add[th][t][c] = 0u;
            if (f0 + 4 * th + t < p.T) ldg_v8(p.addend + off0 + static_cast<size_t>(4 * th + t) * p.C, add[th][t]);
          }
      }
      mbar_wait(&tmem_full_bar[stage], phase);
      tc_fence_after();
      if (threadIdx.x == 0) DT_TRACE(5, it);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16) + stage * 256;
#pragma unroll
      for (int th = 0; th < 2; ++th) {  // 4 frames x 16 channels at a time: one 32-byte store per frame
        uint32_t v[DT_CG][4];
#pragma unroll
        for (int c = 0; c < DT_CG; ++c) tmem_ld_32x32_x4(taddr + c * 16 + 4 * th, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int f = f0 + 4 * th + t;
          if (f < p.T) {
            const size_t off = off0 + static_cast<size_t>(4 * th + t) * p.C;
            uint32_t u[8];
            if (p.addend != nullptr) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float2 av = bf16x2_to_f32x2(add[th][t][c]);
                u[c] = f32x2_to_bf16x2(__uint_as_float(v[2 * c][t]) + av.x, __uint_as_float(v[2 * c + 1][t]) + av.y);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                u[c] = f32x2_to_bf16x2(__uint_as_float(v[2 * c][t]), __uint_as_float(v[2 * c + 1][t]));
            }
            stg_v8(p.y + off, u);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 0) DT_TRACE(6, it);
      if (lane == 0) mbar_arrive(&tmem_empty_bar[stage]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dwconv_tc_wgrad_body(const DwTcParams& p, const int bid) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  // per stage: x series [16][ZL] (ZL = 1024 + 128 frames incl. halo) and dy series [16][1024]
  const int xs_bytes = DT_CG * p.ZL * 2;
  const int dys_bytes = DT_CG * DT_CHUNK * 2;
  const int stage_bytes = xs_bytes + dys_bytes;
  uint8_t* s_ser = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ser + 2 * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 2;
  uint64_t* done_bar = bars + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;
  const int P = p.K / 2;

  if (warp_idx == 4 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 4);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  const bool has_items = first < p.items_per_cg;

  if (warp_idx >= 6) {
    const int pw = (warp_idx - 6) & 3;
    const int grp = (warp_idx - 6) >> 2;
    int it = grp;
    for (int idx = first + grp * p.ctas_per_cg; idx < p.items_per_cg; idx += 2 * p.ctas_per_cg, it += 2) {
      const int stage = grp;
      const uint32_t phase = (it >> 1) & 1;
      const int n = idx / p.t_chunks, tc = idx - n * p.t_chunks;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* base = s_ser + stage * stage_bytes;
      const size_t uoff = static_cast<size_t>(n) * p.T * p.C + c0;
      load_series(p.x + uoff, p.C, p.T, tc * DT_CHUNK - P, p.ZL, base, pw, lane);
      load_series(p.dy + uoff, p.C, p.T, tc * DT_CHUNK, DT_CHUNK, base + xs_bytes, pw, lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp_idx == 4) {
    {
      // D[t' (M = 128), t (N = 16)] += sum_w z[8w + t'] dy[8w + t]: A and B both MN-major
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 1, 1);
      int it = 0;
      for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
        const int stage = it & 1;
        const uint32_t phase = (it >> 1) & 1;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t xs = smem_u32(s_ser + stage * stage_bytes);
        uint64_t da_c = umma_desc_none(xs, 128, 16);
        uint64_t db_c = umma_desc_none(xs + xs_bytes, 128, 16);
        uint32_t tmem_d = tmem_base;
        const uint64_t da_step = static_cast<uint64_t>(p.ZL * 2 >> 4);
#pragma unroll 1
        for (int c = 0; c < DT_CG; ++c) {
          // 1024 frames = 128 windows of 8 = 8 K-steps of 16 windows (256 B of either series per step)
          uint64_t da = da_c, db = db_c;
          if (leader) {
            if (it == 0)
              umma_bf16_first(tmem_d, da, db, idesc);
            else
              umma_bf16_acc(tmem_d, da, db, idesc);
          }
#pragma unroll
          for (int kc = 1; kc < DT_CHUNK / 128; ++kc) {
            da += 16;
            db += 16;
            if (leader) umma_bf16_acc(tmem_d, da, db, idesc);
          }
          da_c += da_step;
          db_c += DT_CHUNK * 2 >> 4;
          tmem_d += 16;
        }
        if (leader) umma_commit(&empty_bar[stage]);
        __syncwarp();
      }
      if (leader) umma_commit(done_bar);
    }
  } else if (warp_idx < 4) {
    if (has_items) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int tp = warp_idx * 32 + lane;  // t'
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < DT_CG; ++c) {
        uint32_t v[8];
        tmem_ld_32x32_x8(taddr + c * 16, v);
        tmem_ld_wait();
        if (c0 + c < p.C) {
          // dw[j] += D[t + j, t]: this thread holds row t' = t + j
          float* dst = p.dw + static_cast<size_t>(c0 + c) * p.K;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int j = tp - t;
            if (j >= 0 && j < p.K) atomicAdd(dst + j, __uint_as_float(v[t]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

__global__ void __launch_bounds__(DT_THREADS, 1) dwconv_tc_fwd_kernel(const DwTcParams p) {
  dwconv_tc_fwd_body(p, static_cast<int>(blockIdx.x));
}
__global__ void __launch_bounds__(DT_THREADS, 1) dwconv_tc_wgrad_kernel(const DwTcParams p) {
  dwconv_tc_wgrad_body(p, static_cast<int>(blockIdx.x));
}
// Backward of one depthwise layer in ONE launch: CTAs [0, split) compute the data gradient (flipped taps + residual
// addend), the rest the weight gradient.  Both read the same upstream gradient; each half walks twice as many items per
// CTA as a full-grid launch would, which amortises the load -> MMA -> store pipeline fill that dominates these kernels.
__global__ void __launch_bounds__(DT_THREADS, 1)
dwconv_tc_bwd_kernel(const DwTcParams pd, const DwTcParams pw, const int split) {
  if (static_cast<int>(blockIdx.x) < split)
    dwconv_tc_fwd_body(pd, static_cast<int>(blockIdx.x));
  else
    dwconv_tc_wgrad_body(pw, static_cast<int>(blockIdx.x) - split);
}

// ------------------------------------------------------------------------------------------------
// Forward / data gradient, 16-frame rows (LASR_DW16): halves the MMA count per output frame.
//
// In the kernel above a row of the A operand is 16 bytes = 8 frames (the row pitch inside a no-swizzle core matrix), so
// only 8 of the N = 16 accumulator columns carry results.  With the 32-byte swizzle the K-major atom is 8 rows x 32 B:
// row m of A starts at frame 16 m of the channel's series, one K-step (16 bf16) is exactly one row, the K-step j
// operand is the same series viewed 32 j bytes later (rows still overlap: a Hankel matrix), and D[m, t] = y[16 m + t]
// fills all 16 columns: one M = 128 MMA now yields 2048 output frames.  The swizzle is a function of the absolute
// shared-memory address (16-byte chunk bit ^= address bit 7), so the producers store the series once, swizzled, and
// every shifted view of it is consistent.  An item is TWO slots of 1024 series frames (64 rows each): a slot holds
// SL = 1024 - KS valid output frames of one (utterance, chunk); rows past SL / 16 overhang into the next slot's data
// and are discarded.  At T' = 801 a whole utterance is one slot, an item covers two utterances.
// ------------------------------------------------------------------------------------------------
constexpr int D16_SLOTF = 1024;                 // series frames per slot
constexpr int D16_ROWB = 2 * D16_SLOTF * 2;     // bytes of one channel's series in a stage (2 slots)
constexpr int D16_STAGE = DT_CG * D16_ROWB;     // 64 KB


// one slot: global [frames, C] -> series[c][slot frames 0..1023] swizzled; out of [0, T) frames are zero
__device__ __forceinline__ void load_slot_sw32(const __nv_bfloat16* __restrict__ src, int C, int T, int f_base,
                                               uint8_t* rows, int pw, int lane) {
  const int h = lane >> 4;
  const int b = lane & 15;
  uint4 r[2][8];
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int sigma = (pw + 4 * ps) * 128 + 8 * b;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = f_base + sigma + i;
      r[ps][i] = make_uint4(0u, 0u, 0u, 0u);
      if (src != nullptr && f >= 0 && f < T)
        r[ps][i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(f) * C + 8 * h));
    }
  }
#pragma unroll
  for (int ps = 0; ps < 2; ++ps) {
    const int sigma = (pw + 4 * ps) * 128 + 8 * b;
    const uint32_t L = static_cast<uint32_t>(sigma) * 2u;          // logical byte offset inside the slot's row
    const uint32_t Ls = L ^ (((L >> 7) & 1u) << 4);               // 32-byte swizzle (slot rows are 256-byte aligned)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t o[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const uint32_t a = (&r[ps][2 * m].x)[q >> 1], bb = (&r[ps][2 * m + 1].x)[q >> 1];
        o[m] = __byte_perm(a, bb, (q & 1) ? 0x7632 : 0x5410);
      }
      *reinterpret_cast<uint4*>(rows + static_cast<size_t>(8 * h + q) * D16_ROWB + Ls) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// one 128-frame block of a slot (8 warps cover the slot's 1024 series frames in one trip, 32 registers of loads in flight)
__device__ __forceinline__ void load_slot_sw32_blk(const __nv_bfloat16* __restrict__ src, int C, int T, int f_base,
                                                   uint8_t* rows, int blk, int lane) {
  const int h = lane >> 4;
  const int b = lane & 15;
  const int sigma = blk * 128 + 8 * b;
  uint4 r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = f_base + sigma + i;
    r[i] = make_uint4(0u, 0u, 0u, 0u);
    if (src != nullptr && f >= 0 && f < T)
      r[i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(f) * C + 8 * h));
  }
  const uint32_t L = static_cast<uint32_t>(sigma) * 2u;
  const uint32_t Ls = L ^ (((L >> 7) & 1u) << 4);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t o[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const uint32_t a = (&r[2 * m].x)[q >> 1], bb = (&r[2 * m + 1].x)[q >> 1];
      o[m] = __byte_perm(a, bb, (q & 1) ? 0x7632 : 0x5410);
    }
    *reinterpret_cast<uint4*>(rows + static_cast<size_t>(8 * h + q) * D16_ROWB + Ls) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

struct Dw16Params {
  const __nv_bfloat16* x;
  const float* w;
  __nv_bfloat16* y;
  const __nv_bfloat16* addend;
  int N, T, C, K, KS, flip;
  int SL, rows_valid, t_chunks, slots_per_cg, items_per_cg, num_cg, ctas_per_cg, stages;
  int w_early;
  unsigned long long* trace;  // debug timeline (tools/trace_dw.py), normally NULL
};

__global__ void __launch_bounds__(DT_THREADS, 1) dwconv_tc16_fwd_kernel(const Dw16Params p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int toep_bytes_c = p.KS * 32;                       // per channel: KS/8 x 2 cores of 128 B
  uint8_t* s_ser = smem;                                    // [stages][16][2 slots x 1024 frames] bf16, swizzled
  uint8_t* s_toep = s_ser + p.stages * D16_STAGE;           // [16][KS/8][2][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_toep + DT_CG * toep_bytes_c);
  uint64_t* full_bar = bars;          // [4]
  uint64_t* empty_bar = bars + 4;     // [4]
  uint64_t* tmem_full_bar = bars + 8;
  uint64_t* tmem_empty_bar = bars + 10;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = blockIdx.x % p.num_cg;
  const int first = blockIdx.x / p.num_cg;
  const int c0 = cg * DT_CG;
  const int P = p.K / 2;

  if (!p.w_early) pdl_wait();
  // Toeplitz factor B[t, s] = w[s - t], t < 16 output frames of a row, s < KS series frames: K-major no-swizzle cores
  // [s / 8][t / 8], element (t % 8, s % 8)
  for (int i = threadIdx.x; i < DT_CG * p.KS * 16; i += DT_THREADS) {
    const int c = i / (p.KS * 16);
    const int rem = i - c * p.KS * 16;
    const int core = rem >> 6, r = (rem >> 3) & 7, e = rem & 7;
    const int kk = core >> 1, tn = core & 1;
    const int t = 8 * tn + r, sidx = 8 * kk + e;
    const int j = sidx - t;
    float v = 0.f;
    if (j >= 0 && j < p.K && c0 + c < p.C) v = p.w[static_cast<size_t>(c0 + c) * p.K + (p.flip ? p.K - 1 - j : j)];
    reinterpret_cast<__nv_bfloat16*>(s_toep)[i] = __float2bfloat16_rn(v);
  }
  if (warp_idx == 4 && lane == 0) {
    for (int st = 0; st < 4; ++st) {
      mbar_init(&full_bar[st], 4);
      mbar_init(&empty_bar[st], 1);
    }
    for (int st = 0; st < 2; ++st) {
      mbar_init(&tmem_full_bar[st], 1);
      mbar_init(&tmem_empty_bar[st], 4);
    }
    mbar_fence_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp_idx >= 6) {
    // ===================== producers: two groups of 4 warps take alternate items =====================
    const int pw = (warp_idx - 6) & 3;
    const int grp = (warp_idx - 6) >> 2;
    int it = grp;
    for (int idx = first + grp * p.ctas_per_cg; idx < p.items_per_cg; idx += 2 * p.ctas_per_cg, it += 2) {
      const int stage = it % p.stages;
      const uint32_t phase = (it / p.stages) & 1;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
#pragma unroll 1
      for (int sl = 0; sl < 2; ++sl) {
        const int slot = 2 * idx + sl;
        const __nv_bfloat16* src = nullptr;
        int f_base = 0;
        if (slot < p.slots_per_cg) {
          const int n = slot / p.t_chunks, tc = slot - n * p.t_chunks;
          src = p.x + static_cast<size_t>(n) * p.T * p.C + c0;
          f_base = tc * p.SL - P;
        }
        load_slot_sw32(src, p.C, p.T, f_base, s_ser + stage * D16_STAGE + sl * (D16_SLOTF * 2), pw, lane);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp_idx == 4) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 0, 0);
    const int ksteps = p.KS / 16;
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int stage = it % p.stages;
      const uint32_t phase = (it / p.stages) & 1;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      uint64_t da_c = umma_desc_sw32(smem_u32(s_ser + stage * D16_STAGE), 256);
      uint64_t db_c = umma_desc_none(smem_u32(s_toep), 256, 128);
      uint32_t tmem_d = tmem_base + acc * 256;
      const uint64_t da_step = static_cast<uint64_t>(D16_ROWB >> 4), db_step = static_cast<uint64_t>(toep_bytes_c >> 4);
#pragma unroll 1
      for (int c = 0; c < DT_CG; ++c) {
        uint64_t da = da_c, db = db_c;
        if (leader) umma_bf16_first(tmem_d, da, db, idesc);
#pragma unroll 1
        for (int kc = 1; kc < ksteps; ++kc) {
          da += 2;    // 32 B: the series 16 frames later
          db += 32;   // 512 B: two K-cores (x 2 N-cores)
          if (leader) umma_bf16_acc(tmem_d, da, db, idesc);
        }
        da_c += da_step;
        db_c += db_step;
        tmem_d += 16;
      }
      if (leader) {
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
    }
  } else if (warp_idx < 4) {
    // ===================== epilogue: row = 16 output frames x 16 channels =====================
    const int row = warp_idx * 32 + lane;
    const int sl = row >> 6, mr = row & 63;
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int acc = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      const int slot = 2 * idx + sl;
      int n = 0, tc = 0;
      bool live = slot < p.slots_per_cg && mr < p.rows_valid;
      if (live) {
        n = slot / p.t_chunks;
        tc = slot - n * p.t_chunks;
      }
      const int f0 = tc * p.SL + 16 * mr;
      live = live && f0 < p.T;
      const size_t off0 = (static_cast<size_t>(n) * p.T + f0) * p.C + c0;
      mbar_wait(&tmem_full_bar[acc], phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16) + acc * 256;
#pragma unroll 1
      for (int th = 0; th < 4; ++th) {  // 4 frames x 16 channels at a time
        uint32_t add[4][8];
        if (p.addend != nullptr) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int c = 0; c < 8; ++c) add[t][c] = 0u;
            if (live && f0 + 4 * th + t < p.T) ldg_v8(p.addend + off0 + static_cast<size_t>(4 * th + t) * p.C, add[t]);
          }
        }
        uint32_t v[DT_CG][4];
#pragma unroll
        for (int c = 0; c < DT_CG; ++c) tmem_ld_32x32_x4(taddr + c * 16 + 4 * th, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int f = f0 + 4 * th + t;
          if (live && f < p.T) {
            const size_t off = off0 + static_cast<size_t>(4 * th + t) * p.C;
            uint32_t u[8];
            if (p.addend != nullptr) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float2 av = bf16x2_to_f32x2(add[t][c]);
                u[c] = f32x2_to_bf16x2(__uint_as_float(v[2 * c][t]) + av.x, __uint_as_float(v[2 * c + 1][t]) + av.y);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                u[c] = f32x2_to_bf16x2(__uint_as_float(v[2 * c][t]), __uint_as_float(v[2 * c + 1][t]));
            }
            stg_v8(p.y + off, u);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// 16-frame rows, second generation (default for stride-1 bf16 layers): same operands and MMAs as dwconv_tc16_fwd_kernel,
// restructured around what the timeline (tools/trace_dw.py) and tools/tma_probe.cu showed:
//   * the kernel is bound by how fast ONE SM can move 32-byte column slices of a channels-last matrix: ~16-20 B/clk/SM
//     through LDG, TMA loads and STG alike (one L2 request per 32 useful bytes), i.e. ~1.3 us per (16 channels x 801
//     frames) slot for its read + write, against ~0.7 us of MMA time -- so everything else has to hide behind the slices;
//   * FOUR producer groups of 4 warps, one per (stage, slot): the gathers of all four utterance slots a CTA can hold
//     (2 stages x 2 slots) are in flight at once (the round-1 kernels loaded one slot per group and item: 2-8 exposed
//     gather latencies per CTA);
//   * the Toeplitz cores are built from a shared-memory copy of the taps, one 16-byte chunk per thread and trip (the
//     round-1 loop built them element by element with an integer division and a 2-byte store each: 3.5-5.5 us of
//     exposed prologue per launch, measured);
//   * the epilogue moves 2 frames x 16 channels per trip (32 + 16 live registers), which is what lets 22 warps
//     (704 threads) share the register file; the residual-gradient addend (data-gradient launches only: template
//     parameter) is requested one trip ahead.
// ------------------------------------------------------------------------------------------------
constexpr int D2_GROUPS = 4;
constexpr int D2_THREADS = 32 * (6 + 4 * D2_GROUPS);  // 704


template <bool HAS_ADDEND>
__device__ __forceinline__ void dwconv_tc16v2_fwd_body(const Dw16Params& p, const int bid) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) DT_TRACE(7, 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int toep_bytes_c = p.KS * 32;                       // per channel: KS/8 x 2 cores of 128 B
  uint8_t* s_ser = smem;                                    // [2 stages][16][2 slots x 1024 frames] bf16, swizzled
  uint8_t* s_toep = s_ser + 2 * D16_STAGE;                  // [16][KS/8][2][128 B]
  float* s_w = reinterpret_cast<float*>(s_toep + DT_CG * toep_bytes_c);  // [16][DT_MAX_KS] fp32 taps of this channel group
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + DT_CG * DT_MAX_KS);
  uint64_t* full_bar = bars;          // [2] producers (8 warps) -> MMA
  uint64_t* empty_bar = bars + 2;     // [2] MMA -> producers
  uint64_t* tmem_full_bar = bars + 4;
  uint64_t* tmem_empty_bar = bars + 6;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;
  const int P = p.K / 2;

  if (!p.w_early) pdl_wait();
  for (int i = threadIdx.x; i < DT_CG * p.K; i += D2_THREADS) {
    const int c = i / p.K, j = i - c * p.K;
    s_w[c * DT_MAX_KS + j] = c0 + c < p.C ? p.w[static_cast<size_t>(c0 + c) * p.K + (p.flip ? p.K - 1 - j : j)] : 0.f;
  }
  if (warp_idx == 4 && lane == 0) {
    for (int st = 0; st < 2; ++st) {
      mbar_init(&full_bar[st], 16);
      mbar_init(&empty_bar[st], 1);
      mbar_init(&tmem_full_bar[st], 1);
      mbar_init(&tmem_empty_bar[st], 4);
    }
    mbar_fence_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  __syncthreads();
  // Toeplitz factor B[t, s] = w[s - t], t < 16 output frames of a row, s < KS series frames: K-major no-swizzle cores
  // [s / 8][t / 8], element (t % 8, s % 8).  One 16-byte chunk = the 8 elements of (channel, core, row).
  {
    const int chunks_c = 2 * p.KS;  // per channel: KS/8 x 2 cores x 8 rows
    for (int q = threadIdx.x; q < DT_CG * chunks_c; q += D2_THREADS) {
      const int c = q / chunks_c;
      const int rem = q - c * chunks_c;
      const int core = rem >> 3, r = rem & 7;
      const int j0 = 8 * (core >> 1) - (8 * (core & 1) + r);
      const float* wc = s_w + c * DT_MAX_KS;
      uint32_t o[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int j = j0 + 2 * m;
        const float a = (j >= 0 && j < p.K) ? wc[j] : 0.f;
        const float b = (j + 1 >= 0 && j + 1 < p.K) ? wc[j + 1] : 0.f;
        o[m] = f32x2_to_bf16x2(a, b);
      }
      *reinterpret_cast<uint4*>(s_toep + static_cast<size_t>(q) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) DT_TRACE(7, 1);
  pdl_wait();

  if (warp_idx >= 6) {
    // ===================== producers: 8 warps per slot, ALL 16 on the same item =====================
    // (the first item of a CTA is complete after one gather of 51 KB instead of after the gathers of everything the CTA
    // owns; the next item's gather then overlaps the MMAs and the epilogue of this one)
    const int blk = (warp_idx - 6) & 7;
    const int sl = (warp_idx - 6) >> 3;
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (blk == 0 && lane == 0) DT_TRACE(0, 2 * it + sl);
      const int slot = 2 * idx + sl;
      const __nv_bfloat16* src = nullptr;
      int f_base = 0;
      if (slot < p.slots_per_cg) {
        const int n = slot / p.t_chunks, tc = slot - n * p.t_chunks;
        src = p.x + static_cast<size_t>(n) * p.T * p.C + c0;
        f_base = tc * p.SL - P;
      }
      load_slot_sw32_blk(src, p.C, p.T, f_base, s_ser + stage * D16_STAGE + sl * (D16_SLOTF * 2), blk, lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (blk == 0 && lane == 0) DT_TRACE(1, 2 * it + sl);
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp_idx == 4) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 0, 0);
    const int ksteps = p.KS / 16;
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[stage], phase ^ 1u);
      if (leader) DT_TRACE(2, it);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (leader) DT_TRACE(3, it);
      uint64_t da_c = umma_desc_sw32(smem_u32(s_ser + stage * D16_STAGE), 256);
      uint64_t db_c = umma_desc_none(smem_u32(s_toep), 256, 128);
      uint32_t tmem_d = tmem_base + stage * 256;
      const uint64_t da_step = static_cast<uint64_t>(D16_ROWB >> 4), db_step = static_cast<uint64_t>(toep_bytes_c >> 4);
#pragma unroll 1
      for (int c = 0; c < DT_CG; ++c) {
        uint64_t da = da_c, db = db_c;
        if (leader) umma_bf16_first(tmem_d, da, db, idesc);
#pragma unroll 1
        for (int kc = 1; kc < ksteps; ++kc) {
          da += 2;    // 32 B: the series 16 frames later
          db += 32;   // 512 B: two K-cores (x 2 N-cores)
          if (leader) umma_bf16_acc(tmem_d, da, db, idesc);
        }
        da_c += da_step;
        db_c += db_step;
        tmem_d += 16;
      }
      if (leader) {
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[stage]);
        DT_TRACE(4, it);
      }
      __syncwarp();
    }
  } else if (warp_idx < 4) {
    // ===================== epilogue: row = 16 output frames x 16 channels, 2 frames per trip =====================
    const int row = warp_idx * 32 + lane;
    const int sl = row >> 6, mr = row & 63;
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int acc = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      const int slot = 2 * idx + sl;
      int n = 0, tc = 0;
      bool live = slot < p.slots_per_cg && mr < p.rows_valid;
      if (live) {
        n = slot / p.t_chunks;
        tc = slot - n * p.t_chunks;
      }
      const int f0 = tc * p.SL + 16 * mr;
      live = live && f0 < p.T;
      const size_t off0 = (static_cast<size_t>(n) * p.T + f0) * p.C + c0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16) + acc * 256;
      uint32_t nxt[2][8];  // HAS_ADDEND: the residual-branch gradient rows of the next trip
      if constexpr (HAS_ADDEND) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int c = 0; c < 8; ++c) nxt[t][c] = 0u;
          if (live && f0 + t < p.T) ldg_v8(p.addend + off0 + static_cast<size_t>(t) * p.C, nxt[t]);
        }
      }
      mbar_wait(&tmem_full_bar[acc], phase);
      tc_fence_after();
      if (threadIdx.x == 0) DT_TRACE(5, it);
#pragma unroll 1
      for (int th = 0; th < 8; ++th) {
        uint32_t v[DT_CG][2];
#pragma unroll
        for (int c = 0; c < DT_CG; ++c) tmem_ld_32x32_x2(taddr + c * 16 + 2 * th, v[c]);
        uint32_t add[2][8];
        if constexpr (HAS_ADDEND) {
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < 8; ++c) add[t][c] = nxt[t][c];
          if (th + 1 < 8) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
              for (int c = 0; c < 8; ++c) nxt[t][c] = 0u;
              if (live && f0 + 2 * th + 2 + t < p.T)
                ldg_v8(p.addend + off0 + static_cast<size_t>(2 * th + 2 + t) * p.C, nxt[t]);
            }
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int f = f0 + 2 * th + t;
          if (live && f < p.T) {
            const size_t off = off0 + static_cast<size_t>(2 * th + t) * p.C;
            uint32_t u[8];
            if constexpr (HAS_ADDEND) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float2 av = bf16x2_to_f32x2(add[t][c]);
                u[c] = f32x2_to_bf16x2(__uint_as_float(v[2 * c][t]) + av.x, __uint_as_float(v[2 * c + 1][t]) + av.y);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                u[c] = f32x2_to_bf16x2(__uint_as_float(v[2 * c][t]), __uint_as_float(v[2 * c + 1][t]));
            }
            stg_v8(p.y + off, u);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 0) DT_TRACE(6, it);
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  if (threadIdx.x == 0) DT_TRACE(7, 3);

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient, second generation: the MMAs of dwconv_tc_wgrad_body, FOUR producer groups -- group g fills
// (stage g / 2, operand g % 2: the x series with its halo / the dy series), so both operands of both stages are gathered
// concurrently instead of one after the other by the same group.  One 128-frame block per warp and trip (32 registers).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_series_block(const __nv_bfloat16* __restrict__ src, int C, int T, int f_base, int ZL,
                                                  uint8_t* series, int blk, int lane) {
  const int h = lane >> 4;
  const int b = lane & 15;
  const int sigma = blk * 128 + 8 * b;
  if (sigma >= ZL) return;
  uint4 r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = f_base + sigma + i;
    r[i] = make_uint4(0u, 0u, 0u, 0u);
    if (f >= 0 && f < T) r[i] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(f) * C + 8 * h));
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t o[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const uint32_t a = (&r[2 * m].x)[q >> 1], bb = (&r[2 * m + 1].x)[q >> 1];
      o[m] = __byte_perm(a, bb, (q & 1) ? 0x7632 : 0x5410);
    }
    *reinterpret_cast<uint4*>(series + (static_cast<size_t>(8 * h + q) * ZL + sigma) * 2) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __forceinline__ void dwconv_tc_wgrad2_body(const DwTcParams& p, const int bid) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((128u - (raw_addr & 127u)) & 127u);
  const int xs_bytes = DT_CG * p.ZL * 2;
  const int dys_bytes = DT_CG * DT_CHUNK * 2;
  const int stage_bytes = xs_bytes + dys_bytes;
  uint8_t* s_ser = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ser + 2 * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 2;
  uint64_t* done_bar = bars + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;
  const int P = p.K / 2;

  if (warp_idx == 4 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 8);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  const bool has_items = first < p.items_per_cg;

  if (warp_idx >= 6) {
    const int pw = (warp_idx - 6) & 3;
    const int grp = (warp_idx - 6) >> 2;
    const int stage = grp >> 1, which = grp & 1;
    int it = stage;
    for (int idx = first + stage * p.ctas_per_cg; idx < p.items_per_cg; idx += 2 * p.ctas_per_cg, it += 2) {
      const uint32_t phase = (it >> 1) & 1;
      const int n = idx / p.t_chunks, tc = idx - n * p.t_chunks;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* base = s_ser + stage * stage_bytes;
      const size_t uoff = static_cast<size_t>(n) * p.T * p.C + c0;
      if (which == 0) {
#pragma unroll 1
        for (int blk = pw; blk * 128 < p.ZL; blk += 4)
          load_series_block(p.x + uoff, p.C, p.T, tc * DT_CHUNK - P, p.ZL, base, blk, lane);
      } else {
#pragma unroll 1
        for (int blk = pw; blk * 128 < DT_CHUNK; blk += 4)
          load_series_block(p.dy + uoff, p.C, p.T, tc * DT_CHUNK, DT_CHUNK, base + xs_bytes, blk, lane);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp_idx == 4) {
    // D[t' (M = 128), t (N = 16)] += sum_w z[8w + t'] dy[8w + t]: A and B both MN-major
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 1, 1);
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t xs = smem_u32(s_ser + stage * stage_bytes);
      uint64_t da_c = umma_desc_none(xs, 128, 16);
      uint64_t db_c = umma_desc_none(xs + xs_bytes, 128, 16);
      uint32_t tmem_d = tmem_base;
      const uint64_t da_step = static_cast<uint64_t>(p.ZL * 2 >> 4);
#pragma unroll 1
      for (int c = 0; c < DT_CG; ++c) {
        uint64_t da = da_c, db = db_c;
        if (leader) {
          if (it == 0)
            umma_bf16_first(tmem_d, da, db, idesc);
          else
            umma_bf16_acc(tmem_d, da, db, idesc);
        }
#pragma unroll
        for (int kc = 1; kc < DT_CHUNK / 128; ++kc) {
          da += 16;
          db += 16;
          if (leader) umma_bf16_acc(tmem_d, da, db, idesc);
        }
        da_c += da_step;
        db_c += DT_CHUNK * 2 >> 4;
        tmem_d += 16;
      }
      if (leader) umma_commit(&empty_bar[stage]);
      __syncwarp();
    }
    if (leader) umma_commit(done_bar);
  } else if (warp_idx < 4) {
    if (has_items) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int tp = warp_idx * 32 + lane;  // t'
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < DT_CG; ++c) {
        uint32_t v[8];
        tmem_ld_32x32_x8(taddr + c * 16, v);
        tmem_ld_wait();
        if (c0 + c < p.C) {
          float* dst = p.dw + static_cast<size_t>(c0 + c) * p.K;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int j = tp - t;
            if (j >= 0 && j < p.K) atomicAdd(dst + j, __uint_as_float(v[t]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <bool HAS_ADDEND>
__global__ void __launch_bounds__(D2_THREADS, 1) dwconv_tc16v2_fwd_kernel(const Dw16Params p) {
  dwconv_tc16v2_fwd_body<HAS_ADDEND>(p, static_cast<int>(blockIdx.x));
}
__global__ void __launch_bounds__(D2_THREADS, 1) dwconv_tc_wgrad2_kernel(const DwTcParams p) {
  dwconv_tc_wgrad2_body(p, static_cast<int>(blockIdx.x));
}
// Backward of one depthwise layer in ONE launch (second generation bodies): CTAs [0, split) compute the data gradient
// (flipped taps + residual addend, 16-frame rows), the rest the weight gradient.  Each half walks twice as many items per
// CTA as a launch of its own would, so the fixed cost of a launch (ramp, prologue, tail: ~4 us) is paid once per layer
// and the two halves' slice traffic interleaves on every SM pair.
template <bool HAS_ADDEND>
__global__ void __launch_bounds__(D2_THREADS, 1)
dwconv_tc_bwd2_kernel(const Dw16Params pd, const DwTcParams pw, const int split) {
  if (static_cast<int>(blockIdx.x) < split)
    dwconv_tc16v2_fwd_body<HAS_ADDEND>(pd, static_cast<int>(blockIdx.x));
  else
    dwconv_tc_wgrad2_body(pw, static_cast<int>(blockIdx.x) - split);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static void dt_schedule(DwTcParams& p, int sms = 0) {
  if (sms <= 0) sms = sm_budget();
  p.num_cg = cdiv(p.C, DT_CG);
  p.t_chunks = cdiv(p.T, DT_CHUNK);
  p.items_per_cg = p.N * p.t_chunks;
  int per = sms / p.num_cg;
  if (per < 1) per = 1;
  if (per > p.items_per_cg) per = p.items_per_cg;
  const int rounds = cdiv(p.items_per_cg, per);
  p.ctas_per_cg = cdiv(p.items_per_cg, rounds);
}

int dwconv_tc_supported(int C, int K, int stride) {
  static const bool off = getenv("LASR_DWCONV_FFMA") != nullptr;  // A/B switch: force the fp32-pipe kernels
  return !off && stride == 1 && (C % DT_CG) == 0 && K >= 3 && (K & 1) && K + 7 <= DT_MAX_KS;
}

static void dw16_params(Dw16Params& p, const void* x, const float* w, void* y, const void* addend, int N, int T, int C,
                        int K, int flip, int sms) {
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.w = w;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.N = N;
  p.T = T;
  p.C = C;
  p.K = K;
  p.KS = cdiv(K + 15, 16) * 16;
  p.flip = flip;
  p.w_early = early_param_loads() ? 1 : 0;
  p.trace = g_dt_trace;
  p.SL = (D16_SLOTF - p.KS) / 16 * 16;
  p.rows_valid = p.SL / 16;
  p.t_chunks = cdiv(T, p.SL);
  p.slots_per_cg = N * p.t_chunks;
  p.items_per_cg = cdiv(p.slots_per_cg, 2);
  p.num_cg = cdiv(C, DT_CG);
  int per = sms / p.num_cg;
  if (per < 1) per = 1;
  if (per > p.items_per_cg) per = p.items_per_cg;
  const int rounds = cdiv(p.items_per_cg, per);
  p.ctas_per_cg = cdiv(p.items_per_cg, rounds);
  p.stages = 2;
}

static int dwconv_tc16_fwd(const void* x, const float* w, void* y, const void* addend, int N, int T, int C, int K,
                           int flip, cudaStream_t stream) {
  Dw16Params p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.w = w;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.N = N;
  p.T = T;
  p.C = C;
  p.K = K;
  p.KS = cdiv(K + 15, 16) * 16;
  p.flip = flip;
  p.w_early = early_param_loads() ? 1 : 0;
  p.trace = g_dt_trace;
  p.SL = (D16_SLOTF - p.KS) / 16 * 16;
  p.rows_valid = p.SL / 16;
  p.t_chunks = cdiv(T, p.SL);
  p.slots_per_cg = N * p.t_chunks;
  p.items_per_cg = cdiv(p.slots_per_cg, 2);
  p.num_cg = cdiv(C, DT_CG);
  int per = sm_budget() / p.num_cg;
  if (per < 1) per = 1;
  if (per > p.items_per_cg) per = p.items_per_cg;
  const int rounds = cdiv(p.items_per_cg, per);
  p.ctas_per_cg = cdiv(p.items_per_cg, rounds);
  const int toep = DT_CG * p.KS * 32;
  static const int v2 = getenv("LASR_DW16_V2") ? atoi(getenv("LASR_DW16_V2")) : 1;
  if (v2) {
    p.stages = 2;
    const int smem2 = 1024 + 2 * D16_STAGE + toep + DT_CG * DT_MAX_KS * 4 + 256;
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(dwconv_tc16v2_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           232448);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dwconv_tc16v2_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e != cudaSuccess) {
        lasr_set_cuda_error(e);
        return LASR_ERR_CUDA;
      }
      configured2 = true;
    }
    const dim3 grid(p.num_cg * p.ctas_per_cg);
    if (addend != nullptr)
      LASR_CHECK_PDL(launch_pdl(2, dwconv_tc16v2_fwd_kernel<true>, grid, dim3(D2_THREADS), smem2, stream, p));
    else
      LASR_CHECK_PDL(launch_pdl(2, dwconv_tc16v2_fwd_kernel<false>, grid, dim3(D2_THREADS), smem2, stream, p));
    return LASR_OK;
  }
  int stages = (232448 - 1024 - 256 - toep) / D16_STAGE;
  if (stages > 3) stages = 3;
  if (stages < 2) return LASR_ERR_UNSUPPORTED;
  p.stages = stages;
  const int smem = 1024 + stages * D16_STAGE + toep + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_tc16_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  LASR_CHECK_PDL(launch_pdl(2, dwconv_tc16_fwd_kernel, dim3(p.num_cg * p.ctas_per_cg), dim3(DT_THREADS), smem, stream, p));
  return LASR_OK;
}

int dwconv_tc_fwd(const void* x, const float* w, void* y, const void* addend, int N, int T, int C, int K, int flip,
                  cudaStream_t stream) {
  static const int dw16 = getenv("LASR_DW16") ? atoi(getenv("LASR_DW16")) : 1;
  if (dw16 && K + 15 <= 112) {
    const int rc16 = dwconv_tc16_fwd(x, w, y, addend, N, T, C, K, flip, stream);
    if (rc16 != LASR_ERR_UNSUPPORTED) return rc16;
  }
  DwTcParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.w = w;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.N = N;
  p.T = T;
  p.C = C;
  p.K = K;
  p.KS = cdiv(K + 7, 16) * 16;
  p.flip = flip;
  p.trace = g_dt_trace;
  p.exp = getenv("LASR_DW_EXP") ? atoi(getenv("LASR_DW_EXP")) : 0;
  p.w_early = early_param_loads() ? 1 : 0;
  p.ZL = DT_CHUNK + p.KS - 8;
  dt_schedule(p);
  int smem = 128 + DT_CG * (p.KS / 8) * 128 + DT_STAGES * DT_CG * p.ZL * 2 + 128;
  if (smem < 116 * 1024) smem = 116 * 1024;  // one CTA per SM: each allocates all 512 TMEM columns
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  LASR_CHECK_PDL(launch_pdl(2, dwconv_tc_fwd_kernel, dim3(p.num_cg * p.ctas_per_cg), dim3(DT_THREADS), smem, stream, p));
  return LASR_OK;
}

int dwconv_tc_wgrad(const void* x, const void* dy, float* dw, int N, int T, int C, int K, cudaStream_t stream) {
  DwTcParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.dw = dw;
  p.N = N;
  p.T = T;
  p.C = C;
  p.K = K;
  p.KS = 0;
  // rows t' = t + j <= 7 + K - 1 < 128; the x series must cover frames up to 8*127 + 127 of the chunk: 1024 + 128
  p.ZL = DT_CHUNK + 128;
  dt_schedule(p);
  const int smem = 128 + 2 * (DT_CG * p.ZL * 2 + DT_CG * DT_CHUNK * 2) + 128;
  static const int wg2 = getenv("LASR_DW_WGRAD2") ? atoi(getenv("LASR_DW_WGRAD2")) : 1;
  if (wg2) {
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e =
          cudaFuncSetAttribute(dwconv_tc_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e != cudaSuccess) {
        lasr_set_cuda_error(e);
        return LASR_ERR_CUDA;
      }
      configured2 = true;
    }
    LASR_CHECK_PDL(
        launch_pdl(2, dwconv_tc_wgrad2_kernel, dim3(p.num_cg * p.ctas_per_cg), dim3(D2_THREADS), smem, stream, p));
    return LASR_OK;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(dwconv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  LASR_CHECK_PDL(launch_pdl(2, dwconv_tc_wgrad_kernel, dim3(p.num_cg * p.ctas_per_cg), dim3(DT_THREADS), smem, stream, p));
  return LASR_OK;
}


// dx = corr(dy, flipped taps) + addend   and   dw += sum dy * shifted x   in one launch (dwconv_tc_bwd_kernel)
int dwconv_tc_bwd(const void* x, const void* dy, const float* w, const void* addend, void* dx, float* dw, int N, int T,
                  int C, int K, cudaStream_t stream) {
  // one launch for both gradients (8-frame-row bodies) only when asked for: the default data-gradient kernel is the
  // 16-frame-row one (dwconv_tc16v2_fwd_kernel), launched on its own next to the weight-gradient kernel
  static const bool dw16 = getenv("LASR_DW16") ? atoi(getenv("LASR_DW16")) != 0 : true;
  static const bool grouped = getenv("LASR_DW_GROUPED") ? atoi(getenv("LASR_DW_GROUPED")) != 0 : true;
  if (!grouped || g_dt_trace != nullptr) return LASR_ERR_UNSUPPORTED;
  if (dw16 && K + 15 <= 112) {
    // second-generation bodies: 16-frame-row data gradient + 4-group weight gradient behind one blockIdx split
    Dw16Params pd{};
    dw16_params(pd, dy, w, dx, addend, N, T, C, K, 1, sm_budget() / 2);
    DwTcParams pw{};
    pw.x = static_cast<const __nv_bfloat16*>(x);
    pw.dy = static_cast<const __nv_bfloat16*>(dy);
    pw.dw = dw;
    pw.N = N;
    pw.T = T;
    pw.C = C;
    pw.K = K;
    pw.KS = 0;
    pw.ZL = DT_CHUNK + 128;
    dt_schedule(pw, sm_budget() / 2);
    const int smem_d = 1024 + 2 * D16_STAGE + DT_CG * pd.KS * 32 + DT_CG * DT_MAX_KS * 4 + 256;
    const int smem_w = 128 + 2 * (DT_CG * pw.ZL * 2 + DT_CG * DT_CHUNK * 2) + 128;
    const int smem = smem_d > smem_w ? smem_d : smem_w;
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e =
          cudaFuncSetAttribute(dwconv_tc_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dwconv_tc_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      if (e != cudaSuccess) {
        lasr_set_cuda_error(e);
        return LASR_ERR_CUDA;
      }
      configured2 = true;
    }
    const int split = pd.num_cg * pd.ctas_per_cg;
    const dim3 grid(split + pw.num_cg * pw.ctas_per_cg);
    if (addend != nullptr)
      LASR_CHECK_PDL(launch_pdl(2, dwconv_tc_bwd2_kernel<true>, grid, dim3(D2_THREADS), smem, stream, pd, pw, split));
    else
      LASR_CHECK_PDL(launch_pdl(2, dwconv_tc_bwd2_kernel<false>, grid, dim3(D2_THREADS), smem, stream, pd, pw, split));
    return LASR_OK;
  }
  DwTcParams pd{};
  pd.x = static_cast<const __nv_bfloat16*>(dy);
  pd.w = w;
  pd.y = static_cast<__nv_bfloat16*>(dx);
  pd.addend = static_cast<const __nv_bfloat16*>(addend);
  pd.N = N;
  pd.T = T;
  pd.C = C;
  pd.K = K;
  pd.KS = cdiv(K + 7, 16) * 16;
  pd.flip = 1;
  pd.w_early = early_param_loads() ? 1 : 0;
  pd.ZL = DT_CHUNK + pd.KS - 8;
  dt_schedule(pd, sm_budget() / 2);
  DwTcParams pw{};
  pw.x = static_cast<const __nv_bfloat16*>(x);
  pw.dy = static_cast<const __nv_bfloat16*>(dy);
  pw.dw = dw;
  pw.N = N;
  pw.T = T;
  pw.C = C;
  pw.K = K;
  pw.KS = 0;
  pw.ZL = DT_CHUNK + 128;
  dt_schedule(pw, sm_budget() / 2);
  int smem_d = 128 + DT_CG * (pd.KS / 8) * 128 + DT_STAGES * DT_CG * pd.ZL * 2 + 128;
  const int smem_w = 128 + 2 * (DT_CG * pw.ZL * 2 + DT_CG * DT_CHUNK * 2) + 128;
  int smem = smem_d > smem_w ? smem_d : smem_w;
  if (smem < 116 * 1024) smem = 116 * 1024;  // one CTA per SM (the data-gradient half allocates all 512 TMEM columns)
  if (smem > 200 * 1024) return LASR_ERR_UNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  const int split = pd.num_cg * pd.ctas_per_cg;
  const int grid = split + pw.num_cg * pw.ctas_per_cg;
  LASR_CHECK_PDL(launch_pdl(2, dwconv_tc_bwd_kernel, dim3(grid), dim3(DT_THREADS), smem, stream, pd, pw, split));
  return LASR_OK;
}

}  // namespace lasr
