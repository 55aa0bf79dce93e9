// Depthwise Conv1d on the tensor cores, fed by TMA from CHANNEL-MAJOR series (bf16, stride 1): forward, data gradient
// and weight gradient of nn.Conv1d(C, C, k, groups=C) at models/QuartNet.py:14-21,30.
//
// dwconv_tc.cu computes the same Hankel x Toeplitz products (see its header) from channels-last activations, and is
// bound by getting the per-channel time series out of a [frames, C] matrix: producer warps gather 16-byte pieces (one
// L2 request per 32 useful bytes), transpose them with byte permutes and only then can the first MMA start, on a CTA that
// owns 2-8 items in all.  Here the series already exist in memory: the pass that PRODUCES a depthwise conv's input (the
// BatchNorm apply pass, bn_apply_fwd_cm_kernel) or its upstream gradient (the pointwise data-gradient GEMM with swapped
// operands, gemm_tc.cu) writes a channel-major companion in exactly the shared-memory image the MMA wants:
//
//   xT [C][N][S] bf16   frame t of utterance n at position off + t of row (c, n), off = round_up(K/2, 8) (the conv's
//                       left zero padding, rounded to a 16-byte group; the remainder delta = off - K/2 is folded into
//                       the Toeplitz factor), zeros before and after; S % 128 == 0; 16-byte groups stored at
//                       g ^ ((g >> 3) & 1) -- the SWIZZLE_32B pattern as a function of the position, which equals
//                       the pattern as a function of the shared-memory address when 128-position blocks are copied to
//                       256-byte aligned destinations.
//
// One cp.async.bulk.tensor.5d per item moves 16 channels x 2 slots x 1024 positions (64 KB, 256-byte rows) into
//   series[16 channels][2 slots][1024]  =  the K-major SWIZZLE_32B A operand of dwconv_tc16v2 (row m = 16 positions,
//   64 rows per slot; a slot's first 56 rows = 896 outputs are valid, the others overhang),
// issued by one thread; nothing else touches the input.  Warp roles: 0-7 epilogue (TMEM -> bf16 -> channels-last
// 32-byte stores, two warps per TMEM lane quadrant taking 8 of a row's 16 frames each), 8 MMA issuer, 9 TMEM allocator +
// TMA producer.
#include "dw_common.cuh"

#include <cstdlib>

namespace lasr {

constexpr int CM_SLOTF = 1024;                 // series positions per slot
constexpr int CM_SL = 896;                     // valid outputs per slot (7 blocks of 128)
constexpr int CM_ROWB = 2 * CM_SLOTF * 2;      // bytes of one channel's series in a stage (2 slots)
constexpr int CM_STAGE = DT_CG * CM_ROWB;      // 64 KB
constexpr int CM_THREADS = 32 * 11;  // 8 epilogue warps, 2 MMA issuers, 1 TMA producer / TMEM allocator

struct DwCmParams {
  const float* w;
  __nv_bfloat16* y;
  const __nv_bfloat16* addend;
  int S, off;  // series pitch / frame offset (ADD == 2: the addend series)
  int N, T, C, K, KS, flip, delta;
  int t_chunks, pair_utt, pairs_per_utt;
  int items_per_cg, num_cg, ctas_per_cg, stages, w_early;
  int issuers;                // MMA issuer warps (1 or 2: 8 channels each)
  const __nv_bfloat16* toep;  // nullable: the Toeplitz factors of all channels, prebuilt by the pass that wrote the series
                              // ([C][2 KS chunks of 16 B], lasr_bn_apply_act_fwd_cm): one bulk copy instead of the build
  unsigned long long* trace;  // debug timeline (tools/trace_dw.py), normally NULL
};

int lasr_cm_offset_host(int K) { return (K / 2 + 7) / 8 * 8; }
int lasr_cm_ks_host(int K) { return cdiv(K + 15 + lasr_cm_offset_host(K) - K / 2, 16) * 16; }
int lasr_cm_pitch_host(int T, int K) {
  const int off = lasr_cm_offset_host(K);
  const int t_chunks = cdiv(T, CM_SL);
  if (t_chunks == 1) return cdiv(off + T + 48, 128) * 128;
  return CM_SL * (cdiv(t_chunks, 2) * 2 - 1) + CM_SLOTF;
}

// ADD: 0 no addend, 1 channels-last addend [N, T, C], 2 addend as a series tensor laid out like the input
template <int ADD>
__device__ __forceinline__ void dwconv_cm_fwd_body(const CUtensorMap* __restrict__ tmap, const DwCmParams& p, const int bid) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int toep_bytes_c = p.KS * 32;                       // per channel: KS/8 x 2 cores of 128 B
  uint8_t* s_ser = smem;                                    // [stages][16][2 slots x 1024] bf16, swizzled as stored
  uint8_t* s_toep = s_ser + p.stages * CM_STAGE;            // [16][KS/8][2][128 B]
  float* s_w = reinterpret_cast<float*>(s_toep + DT_CG * toep_bytes_c);  // [16][DT_MAX_KS] fp32 taps of the channel group
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + DT_CG * DT_MAX_KS);
  uint64_t* full_bar = bars;          // [4] TMA -> MMA
  uint64_t* empty_bar = bars + 4;     // [4] MMA -> TMA
  uint64_t* tmem_full_bar = bars + 8;   // [2]
  uint64_t* tmem_empty_bar = bars + 10; // [2]
  uint64_t* toep_bar = bars + 12;       // prebuilt Toeplitz factors landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;

  constexpr int W_TMA = 10;
  if (threadIdx.x == 0) DW_TRACE(p.trace, 7, 0);
  if (warp_idx == W_TMA && lane == 0) {
    tma_prefetch_desc(tmap);
    for (int st = 0; st < 4; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], p.issuers);
    }
    for (int st = 0; st < 2; ++st) {
      mbar_init(&tmem_full_bar[st], p.issuers);
      mbar_init(&tmem_empty_bar[st], 8);
    }
    mbar_init(toep_bar, 1);
    mbar_fence_init();
    // the series of the first items are requested NOW: they land while the other warps build the Toeplitz factor
    pdl_wait();
    if (p.toep != nullptr) {
      mbar_arrive_expect_tx(toep_bar, static_cast<uint32_t>(DT_CG * toep_bytes_c));
      bulk_load(s_toep, p.toep + static_cast<size_t>(c0) * p.KS * 16, static_cast<uint32_t>(DT_CG * toep_bytes_c), toep_bar);
    }
    int it = 0;
    for (int idx = first; idx < p.items_per_cg && it < p.stages; idx += p.ctas_per_cg, ++it) {
      mbar_arrive_expect_tx(&full_bar[it], CM_STAGE);
      int n, tc;
      if (p.pair_utt) {
        n = 2 * idx;
        tc = 0;
      } else {
        n = idx / p.pairs_per_utt;
        tc = 2 * (idx - n * p.pairs_per_utt);
      }
      tma_load_5d(s_ser + it * CM_STAGE, tmap, &full_bar[it], 0, 0, tc, n, c0);
    }
  }
  if (warp_idx == W_TMA) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  if (p.toep == nullptr) {
    if (!p.w_early) pdl_wait();
    for (int i = threadIdx.x; i < DT_CG * p.K; i += CM_THREADS) {
      const int c = i / p.K, j = i - c * p.K;
      s_w[c * DT_MAX_KS + j] = c0 + c < p.C ? p.w[static_cast<size_t>(c0 + c) * p.K + j] : 0.f;
    }
    __syncthreads();
    // Toeplitz factor B[t, s] = w[s - t - delta] (dw_common.cuh), one 16-byte chunk per thread and trip
    const int chunks_c = 2 * p.KS;  // per channel: KS/8 x 2 cores x 8 rows
    for (int q = threadIdx.x; q < DT_CG * chunks_c; q += CM_THREADS) {
      const int c = q / chunks_c;
      *reinterpret_cast<uint4*>(s_toep + static_cast<size_t>(q) * 16) =
          toeplitz_chunk(s_w + c * DT_MAX_KS, p.K, p.delta, p.flip, q - c * chunks_c);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) DW_TRACE(p.trace, 7, 1);
  pdl_wait();

  if (warp_idx == W_TMA) {
    // ===================== TMA producer: one box = the whole item (the first p.stages items are already in flight) ====
    if (lane == 0) {
      int it = 0;
      for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
        if (it < p.stages) continue;
        const int stage = it % p.stages;
        const uint32_t phase = (it / p.stages) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], CM_STAGE);
        int n, tc;
        if (p.pair_utt) {
          n = 2 * idx;
          tc = 0;
        } else {
          n = idx / p.pairs_per_utt;
          tc = 2 * (idx - n * p.pairs_per_utt);
        }
        DW_TRACE(p.trace, 0, it);
        tma_load_5d(s_ser + stage * CM_STAGE, tmap, &full_bar[stage], 0, 0, tc, n, c0);
      }
    }
  } else if (warp_idx >= 8) {
    // ===================== MMA issuers: warp 8 + w takes channels [w * 16 / issuers, (w + 1) * 16 / issuers) ==========
    // (one thread issuing M128 x N16 x K16 MMAs back to back is paced by its own instruction stream, not by the tensor
    // pipe's 32-cycle read of the A operand: two issuing threads keep the pipe fed)
    const int wi = warp_idx - 8;
    if (wi < p.issuers) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 0, 0);
      const int ksteps = p.KS / 16;
      const int cpi = DT_CG / p.issuers;
      if (p.toep != nullptr) mbar_wait(toep_bar, 0);
      int it = 0;
      for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = (it / p.stages) & 1;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        if (leader && wi == 0) DW_TRACE(p.trace, 2, it);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader && wi == 0) DW_TRACE(p.trace, 3, it);
        const uint64_t da_step = static_cast<uint64_t>(CM_ROWB >> 4), db_step = static_cast<uint64_t>(toep_bytes_c >> 4);
        uint64_t da_c = umma_desc_sw32(smem_u32(s_ser + stage * CM_STAGE), 256) + da_step * (wi * cpi);
        uint64_t db_c = umma_desc_none(smem_u32(s_toep), 256, 128) + db_step * (wi * cpi);
        uint32_t tmem_d = tmem_base + acc * 256 + 16 * (wi * cpi);
#pragma unroll 1
        for (int c = 0; c < cpi; ++c) {
          uint64_t da = da_c, db = db_c;
          if (leader) umma_bf16_first(tmem_d, da, db, idesc);
#pragma unroll 1
          for (int kc = 1; kc < ksteps; ++kc) {
            da += 2;    // 32 B: the series 16 positions later
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
          if (wi == 0) DW_TRACE(p.trace, 4, it);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: row = 16 output frames x 16 channels; warp w and w + 4 share a lane quadrant and
    // take frames 0-7 / 8-15 of every row; 2 frames per trip =====================
    const int quad = warp_idx & 3, half = warp_idx >> 2;
    const int row = quad * 32 + lane;
    const int sl = row >> 6, mr = row & 63;
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int acc = it & 1;
      const uint32_t phase = (it >> 1) & 1;
      int n, tc;
      if (p.pair_utt) {
        n = 2 * idx + sl;
        tc = 0;
      } else {
        n = idx / p.pairs_per_utt;
        tc = 2 * (idx - n * p.pairs_per_utt) + sl;
      }
      const int f0 = tc * CM_SL + 16 * mr + 8 * half;
      const bool live = n < p.N && tc < p.t_chunks && mr < CM_SL / 16 && f0 < p.T;
      const size_t off0 = (static_cast<size_t>(live ? n : 0) * p.T + (live ? f0 : 0)) * p.C + c0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + 8 * half;
      uint32_t nxt[2][8];  // ADD == 1: the residual-branch gradient rows of the next trip
      uint4 adds[ADD == 2 ? DT_CG : 1];  // ADD == 2: this thread's 8 frames of every channel (one 16-byte group each)
      if constexpr (ADD == 1) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int c = 0; c < 8; ++c) nxt[t][c] = 0u;
          if (live && f0 + t < p.T) ldg_v8(p.addend + off0 + static_cast<size_t>(t) * p.C, nxt[t]);
        }
      }
      if constexpr (ADD == 2) {
        const int g = (p.off + (live ? f0 : 0)) >> 3;
        const int gs = g ^ ((g >> 3) & 1);
        const __nv_bfloat16* ap =
            p.addend + (static_cast<size_t>(c0) * p.N + (live ? n : 0)) * p.S + static_cast<size_t>(gs) * 8;
        const size_t cstride = static_cast<size_t>(p.N) * p.S;
#pragma unroll
        for (int c = 0; c < DT_CG; ++c) {
          adds[c] = make_uint4(0u, 0u, 0u, 0u);
          if (live) adds[c] = __ldg(reinterpret_cast<const uint4*>(ap + c * cstride));
        }
      }
      mbar_wait(&tmem_full_bar[acc], phase);
      tc_fence_after();
      if (threadIdx.x == 0) DW_TRACE(p.trace, 5, it);
#pragma unroll
      for (int th = 0; th < 4; ++th) {
        uint32_t v[DT_CG][2];
#pragma unroll
        for (int c = 0; c < DT_CG; ++c) tmem_ld_32x32_x2(taddr + c * 16 + 2 * th, v[c]);
        uint32_t add[2][8];
        if constexpr (ADD == 1) {
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < 8; ++c) add[t][c] = nxt[t][c];
          if (th + 1 < 4) {
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
        if constexpr (ADD == 2) {
          // word th of a channel's group = frames (2 th, 2 th + 1)
#pragma unroll
          for (int c = 0; c < DT_CG; ++c) {
            const float2 av = bf16x2_to_f32x2((&adds[c].x)[th]);
            v[c][0] = __float_as_uint(__uint_as_float(v[c][0]) + av.x);
            v[c][1] = __float_as_uint(__uint_as_float(v[c][1]) + av.y);
          }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int f = f0 + 2 * th + t;
          if (live && f < p.T) {
            const size_t off = off0 + static_cast<size_t>(2 * th + t) * p.C;
            uint32_t u[8];
            if constexpr (ADD == 1) {
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
      if (threadIdx.x == 0) DW_TRACE(p.trace, 6, it);
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  if (threadIdx.x == 0) DW_TRACE(p.trace, 7, 2);

  tc_fence_before();
  __syncthreads();
  if (warp_idx == W_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int ADD>
__global__ void __launch_bounds__(CM_THREADS, 1)
dwconv_cm_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const DwCmParams p) {
  dwconv_cm_fwd_body<ADD>(&tmap, p, static_cast<int>(blockIdx.x));
}

// ------------------------------------------------------------------------------------------------
// weight gradient from two series tensors: dw_c[j] = sum_f dy_c[f] x_c[f + j - K/2].
// With X[p] = x[p - off] and DY[p] = dy[p - off] (the stored series), windows of 16 positions in the K dimension and
//   A[t', w] = X[b - sh + 16 w + t']   (M = 128 lags, MN-major: t' contiguous)        sh = round_up(K/2, 16)
//   B[w, t]  = DY[b + 16 w + t]        (N = 16, MN-major)
// D[t', t] += sum_w A[t', w] B[w, t] collects sum_p DY[p] X[p + t' - t - sh] over the chunk, i.e. dw[j] = sum_t D[t + j -
// K/2 + sh, t].  Both operands are the series as they lie in shared memory (SWIZZLE_32B MN-major atoms = 16 positions x 8
// windows = 256 contiguous bytes; the atoms of consecutive t' blocks overlap: a Hankel matrix again), one MMA covers 256
// positions, an item (16 channels x 1024 positions of one utterance) is 64 MMAs and two TMA boxes (X with one extra
// 128-position block on either side, DY).  The accumulator stays in TMEM over all items of the CTA; the diagonals are
// summed through shared memory and leave as one atomic per tap and CTA.
// ------------------------------------------------------------------------------------------------
constexpr int WG_CHUNK = 1024;                       // positions per item
constexpr int WG_XPOS = WG_CHUNK + 256;              // X positions loaded per item (one block before, one after)
constexpr int WG_XB = DT_CG * WG_XPOS * 2;           // 40 KB
constexpr int WG_DB = DT_CG * WG_CHUNK * 2;          // 32 KB
constexpr int WG_STAGE = WG_XB + WG_DB;              // 72 KB

struct DwCmWgParams {
  float* dw;
  int N, C, K, sh, chunks, items_per_cg, num_cg, ctas_per_cg, stages;
  unsigned long long* trace;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void dwconv_cm_wgrad_body(const CUtensorMap* __restrict__ tm_x, const CUtensorMap* __restrict__ tm_d,
                                                     const DwCmWgParams& p, const int bid) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* s_ser = smem;  // [stages][X 16 x 1280 | DY 16 x 1024] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ser + p.stages * WG_STAGE);
  uint64_t* full_bar = bars;       // [4]
  uint64_t* empty_bar = bars + 4;  // [4]
  uint64_t* done_bar = bars + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_dw = reinterpret_cast<float*>(s_ser);  // [16][128][17] accumulator copy: re-uses the stages after the last MMA

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;
  constexpr int W_TMA = 10, W_MMA = 8;

  if (warp_idx == W_TMA && lane == 0) {
    tma_prefetch_desc(tm_x);
    tma_prefetch_desc(tm_d);
    for (int st = 0; st < 4; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp_idx == W_TMA) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  const bool has_items = first < p.items_per_cg;
  if (threadIdx.x == 0) {
    DW_TRACE(p.trace, 7, 0);
    DW_TRACE(p.trace, 7, 1);
  }

  if (warp_idx == W_TMA) {
    if (lane == 0) {
      int it = 0;
      for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = (it / p.stages) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        DW_TRACE(p.trace, 0, it);
        mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE);
        const int n = idx / p.chunks, q = idx - n * p.chunks;
        uint8_t* base = s_ser + stage * WG_STAGE;
        tma_load_4d(base, tm_x, &full_bar[stage], 0, 8 * q - 1, n, c0);
        tma_load_4d(base + WG_XB, tm_d, &full_bar[stage], 0, 8 * q, n, c0);
      }
    }
  } else if (warp_idx == W_MMA) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(DT_ROWS, 16, 1, 1);
    int it = 0;
    for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
      const int stage = it % p.stages;
      const uint32_t phase = (it / p.stages) & 1;
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (leader) DW_TRACE(p.trace, 3, it);
      const uint32_t xs = smem_u32(s_ser + stage * WG_STAGE);
      // A starts sh positions before the chunk (the X box starts one 128-position block before it)
      uint64_t da_c = umma_desc_sw32_mn(xs + (128 - p.sh) * 2, 32, 256);
      uint64_t db_c = umma_desc_sw32_mn(xs + WG_XB, 32, 256);
      uint32_t tmem_d = tmem_base;
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
        for (int kc = 1; kc < WG_CHUNK / 256; ++kc) {
          da += 32;  // 512 B: 16 windows of 16 positions
          db += 32;
          if (leader) umma_bf16_acc(tmem_d, da, db, idesc);
        }
        da_c += WG_XPOS * 2 >> 4;
        db_c += WG_CHUNK * 2 >> 4;
        tmem_d += 16;
      }
      if (leader) {
        umma_commit(&empty_bar[stage]);
        DW_TRACE(p.trace, 4, it);
      }
      __syncwarp();
    }
    if (leader) umma_commit(done_bar);
  }
  // ===================== epilogue: accumulators -> shared memory (warps 0-7, two per TMEM lane quadrant), then every
  // thread folds the diagonals dw[c][j] = sum_t D_c[t + j - e, t] and adds them to the gradient =====================
  if (has_items) {
    mbar_wait(done_bar, 0);  // every MMA has retired: the series buffers are free
    tc_fence_after();
    if (threadIdx.x == 0) DW_TRACE(p.trace, 5, 0);
    if (warp_idx < 8) {
      const int quad = warp_idx & 3, chalf = warp_idx >> 2;
      const int tp = quad * 32 + lane;  // t'
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int c = chalf * 8; c < chalf * 8 + 8; ++c) {
        uint32_t v[16];
        tmem_ld_32x32_x16(taddr + c * 16, v);
        tmem_ld_wait();
        float* dst = s_dw + (c * 128 + tp) * 17;
#pragma unroll
        for (int t = 0; t < 16; ++t) dst[t] = __uint_as_float(v[t]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (has_items) {
    const int e = p.K / 2 - p.sh;  // j = t' - t + e
    for (int i = threadIdx.x; i < DT_CG * p.K; i += CM_THREADS) {
      const int c = i / p.K, j = i - c * p.K;
      const float* src = s_dw + c * 128 * 17;
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int tp = j - e + t;
        if (tp >= 0 && tp < 128) acc += src[tp * 17 + t];
      }
      if (c0 + c < p.C) atomicAdd(p.dw + static_cast<size_t>(c0) * p.K + i, acc);
    }
  }
  if (threadIdx.x == 0) DW_TRACE(p.trace, 7, 2);
  if (warp_idx == W_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

__global__ void __launch_bounds__(CM_THREADS, 1)
dwconv_cm_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_d,
                       const DwCmWgParams p) {
  dwconv_cm_wgrad_body(&tm_x, &tm_d, p, static_cast<int>(blockIdx.x));
}

// Backward of one depthwise layer in ONE launch: CTAs [0, split) compute the data gradient (flipped taps + the residual
// branch's gradient as addend), the rest the weight gradient; both read the upstream-gradient series.
template <int ADD>
__global__ void __launch_bounds__(CM_THREADS, 1)
dwconv_cm_bwd_kernel(const __grid_constant__ CUtensorMap tm_items, const __grid_constant__ CUtensorMap tm_x,
                     const __grid_constant__ CUtensorMap tm_d, const DwCmParams pd, const DwCmWgParams pw, const int split) {
  if (static_cast<int>(blockIdx.x) < split)
    dwconv_cm_fwd_body<ADD>(&tm_items, pd, static_cast<int>(blockIdx.x));
  else
    dwconv_cm_wgrad_body(&tm_x, &tm_d, pw, static_cast<int>(blockIdx.x) - split);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
int dwconv_cm_supported(int C, int K) {
  return (C % DT_CG) == 0 && K >= 3 && (K & 1) && K + 15 + 7 <= DT_MAX_KS;
}

// tensor map of a channel-major series tensor for the forward / data-gradient items: {128, blocks, slot, utterance,
// channel}, box = one item (16 channels x 2 slots x 8 blocks x 128 positions)
static int cm_item_tmap(CUtensorMap* tm, const void* xT, int N, int C, int S, int t_chunks, int pair_utt) {
  const uint64_t dims[5] = {128, static_cast<uint64_t>(t_chunks == 1 ? S / 128 : 8), static_cast<uint64_t>(t_chunks),
                            static_cast<uint64_t>(N), static_cast<uint64_t>(C)};
  const uint64_t strides[4] = {256, static_cast<uint64_t>(CM_SL) * 2, static_cast<uint64_t>(S) * 2,
                               static_cast<uint64_t>(N) * S * 2};
  const uint32_t box[5] = {128, 8, pair_utt ? 1u : 2u, pair_utt ? 2u : 1u, DT_CG};
  return make_tmap_nd_bf16(tm, xT, 5, dims, strides, box, false);
}

static void cm_params(DwCmParams& p, const float* w, void* y, const void* addend, int N, int T, int C, int K, int flip,
                      int sms, const void* toep = nullptr) {
  p.w = w;
  p.toep = static_cast<const __nv_bfloat16*>(toep);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.N = N;
  p.T = T;
  p.C = C;
  p.K = K;
  p.flip = flip;
  p.delta = lasr_cm_offset_host(K) - K / 2;
  p.KS = lasr_cm_ks_host(K);
  p.w_early = early_param_loads() ? 1 : 0;
  static const int issuers = getenv("LASR_CM_ISSUERS") ? atoi(getenv("LASR_CM_ISSUERS")) : 2;
  p.issuers = issuers == 1 ? 1 : 2;
  p.trace = dw_trace_buffer();
  p.t_chunks = cdiv(T, CM_SL);
  p.pair_utt = p.t_chunks == 1 ? 1 : 0;
  p.pairs_per_utt = cdiv(p.t_chunks, 2);
  p.items_per_cg = p.pair_utt ? cdiv(N, 2) : N * p.pairs_per_utt;
  p.num_cg = cdiv(C, DT_CG);
  int per = sms / p.num_cg;
  if (per < 1) per = 1;
  if (per > p.items_per_cg) per = p.items_per_cg;
  const int rounds = cdiv(p.items_per_cg, per);
  p.ctas_per_cg = cdiv(p.items_per_cg, rounds);
  const int fixed = 1024 + DT_CG * p.KS * 32 + DT_CG * DT_MAX_KS * 4 + 256;
  int stages = (232448 - fixed) / CM_STAGE;
  if (stages > 3) stages = 3;
  if (stages > rounds) stages = rounds < 2 ? 2 : rounds;
  p.stages = stages;
}

static int cm_configure() {
  static bool configured = false;
  if (configured) return LASR_OK;
  cudaError_t e = cudaFuncSetAttribute(dwconv_cm_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_cm_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_cm_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_cm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_cm_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_cm_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_cm_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) {
    lasr_set_cuda_error(e);
    return LASR_ERR_CUDA;
  }
  configured = true;
  return LASR_OK;
}

static int cm_fwd_smem(const DwCmParams& p) {
  return 1024 + p.stages * CM_STAGE + DT_CG * p.KS * 32 + DT_CG * DT_MAX_KS * 4 + 256;
}

// addend: channels-last [N, T, C]; addendT: a series tensor laid out like xT (at most one of them)
int dwconv_cm_fwd(const void* xT, const float* w, void* y, const void* addend, const void* addendT, const void* toep,
                  int N, int T, int C, int K, int S, int flip, cudaStream_t stream) {
  if (!dwconv_cm_supported(C, K) || S != lasr_cm_pitch_host(T, K) || (addend != nullptr && addendT != nullptr))
    return LASR_ERR_BAD_SHAPE;
  DwCmParams p{};
  cm_params(p, w, y, addendT != nullptr ? addendT : addend, N, T, C, K, flip, sm_budget(), toep);
  p.S = S;
  p.off = lasr_cm_offset_host(K);
  CUtensorMap tm;
  if (int rc = cm_item_tmap(&tm, xT, N, C, S, p.t_chunks, p.pair_utt)) return rc;
  if (int rc = cm_configure()) return rc;
  const int smem = cm_fwd_smem(p);
  const dim3 grid(p.num_cg * p.ctas_per_cg);
  if (addendT != nullptr)
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_fwd_kernel<2>, grid, dim3(CM_THREADS), smem, stream, tm, p));
  else if (addend != nullptr)
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_fwd_kernel<1>, grid, dim3(CM_THREADS), smem, stream, tm, p));
  else
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_fwd_kernel<0>, grid, dim3(CM_THREADS), smem, stream, tm, p));
  return LASR_OK;
}

// {128, blocks per utterance row, utterance, channel}: box = 16 channels x `blocks` 128-position blocks of one utterance
static int cm_row_tmap(CUtensorMap* tm, const void* xT, int N, int C, int S, int blocks) {
  const uint64_t dims[4] = {128, static_cast<uint64_t>(S / 128), static_cast<uint64_t>(N), static_cast<uint64_t>(C)};
  const uint64_t strides[3] = {256, static_cast<uint64_t>(S) * 2, static_cast<uint64_t>(N) * S * 2};
  const uint32_t box[4] = {128, static_cast<uint32_t>(blocks), 1, DT_CG};
  return make_tmap_nd_bf16(tm, xT, 4, dims, strides, box, false);
}

static void cm_wg_params(DwCmWgParams& p, float* dw, int N, int C, int K, int S, int sms) {
  p.dw = dw;
  p.N = N;
  p.C = C;
  p.K = K;
  p.sh = (K / 2 + 15) / 16 * 16;
  p.chunks = cdiv(S, WG_CHUNK);
  p.items_per_cg = N * p.chunks;
  p.num_cg = cdiv(C, DT_CG);
  int per = sms / p.num_cg;
  if (per < 1) per = 1;
  if (per > p.items_per_cg) per = p.items_per_cg;
  const int rounds = cdiv(p.items_per_cg, per);
  p.ctas_per_cg = cdiv(p.items_per_cg, rounds);
  p.stages = rounds < 3 ? (rounds < 2 ? 2 : rounds) : 3;
  p.trace = dw_trace_buffer();
}

static int cm_wg_smem(const DwCmWgParams& p) { return 1024 + p.stages * WG_STAGE + 256; }

int dwconv_cm_wgrad(const void* xT, const void* dyT, float* dw, int N, int T, int C, int K, int S, cudaStream_t stream) {
  if (!dwconv_cm_supported(C, K) || S != lasr_cm_pitch_host(T, K)) return LASR_ERR_BAD_SHAPE;
  DwCmWgParams p{};
  cm_wg_params(p, dw, N, C, K, S, sm_budget());
  CUtensorMap tx, td;
  if (int rc = cm_row_tmap(&tx, xT, N, C, S, WG_XPOS / 128)) return rc;
  if (int rc = cm_row_tmap(&td, dyT, N, C, S, WG_CHUNK / 128)) return rc;
  if (int rc = cm_configure()) return rc;
  LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_wgrad_kernel, dim3(p.num_cg * p.ctas_per_cg), dim3(CM_THREADS), cm_wg_smem(p),
                            stream, tx, td, p));
  return LASR_OK;
}

// dx = corr(dy, flipped taps) + addend   and   dw += sum dy * shifted x   in one launch
int dwconv_cm_bwd(const void* xT, const void* dyT, const float* w, const void* addend, const void* addendT,
                  const void* toep_flip, void* dx, float* dw, int N, int T, int C, int K, int S, cudaStream_t stream) {
  if (!dwconv_cm_supported(C, K) || S != lasr_cm_pitch_host(T, K) || (addend != nullptr && addendT != nullptr))
    return LASR_ERR_BAD_SHAPE;
  // SMs of the data-gradient half.  Half and half is the measured optimum at the config-2 shapes: the data gradient is
  // paced by its epilogue (addend + 32-byte stores), the weight gradient by its MMAs, and a split by MMA counts (tried:
  // 2 + 7 CTAs per channel group at C = 256) made the launch 40 % slower
  static const int forced_share = getenv("LASR_CM_DGRAD_SMS") ? atoi(getenv("LASR_CM_DGRAD_SMS")) : 0;
  const int share = forced_share > 0 ? forced_share : sm_budget() / 2;
  DwCmParams pd{};
  cm_params(pd, w, dx, addendT != nullptr ? addendT : addend, N, T, C, K, 1, share, toep_flip);
  pd.S = S;
  pd.off = lasr_cm_offset_host(K);
  DwCmWgParams pw{};
  cm_wg_params(pw, dw, N, C, K, S, sm_budget() - share);
  CUtensorMap ti, tx, td;
  if (int rc = cm_item_tmap(&ti, dyT, N, C, S, pd.t_chunks, pd.pair_utt)) return rc;
  if (int rc = cm_row_tmap(&tx, xT, N, C, S, WG_XPOS / 128)) return rc;
  if (int rc = cm_row_tmap(&td, dyT, N, C, S, WG_CHUNK / 128)) return rc;
  if (int rc = cm_configure()) return rc;
  const int smem_d = cm_fwd_smem(pd), smem_w = cm_wg_smem(pw);
  const int smem = smem_d > smem_w ? smem_d : smem_w;
  const int split = pd.num_cg * pd.ctas_per_cg;
  const dim3 grid(split + pw.num_cg * pw.ctas_per_cg);
  if (addendT != nullptr)
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_bwd_kernel<2>, grid, dim3(CM_THREADS), smem, stream, ti, tx, td, pd, pw, split));
  else if (addend != nullptr)
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_bwd_kernel<1>, grid, dim3(CM_THREADS), smem, stream, ti, tx, td, pd, pw, split));
  else
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_bwd_kernel<0>, grid, dim3(CM_THREADS), smem, stream, ti, tx, td, pd, pw, split));
  return LASR_OK;
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_cm_offset(int K) { return lasr_cm_offset_host(K); }
int lasr_cm_pitch(int T, int K) { return lasr_cm_pitch_host(T, K); }

int lasr_cm_ks(int K) { return lasr_cm_ks_host(K); }

int lasr_dwconv1d_fwd_cm(const void* xT, const float* w, void* y, const void* addend, const void* addendT,
                         const void* toep, int N, int T, int C, int K, int S, int flip, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || xT == nullptr || y == nullptr || w == nullptr) return LASR_ERR_BAD_SHAPE;
  return dwconv_cm_fwd(xT, w, y, addend, addendT, toep, N, T, C, K, S, flip, stream);
}

int lasr_dwconv1d_wgrad_cm(const void* xT, const void* dyT, float* dw, int N, int T, int C, int K, int S,
                           lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || xT == nullptr || dyT == nullptr || dw == nullptr) return LASR_ERR_BAD_SHAPE;
  return dwconv_cm_wgrad(xT, dyT, dw, N, T, C, K, S, stream);
}

int lasr_dwconv1d_bwd_cm(const void* xT, const void* dyT, const float* w, const void* addend, const void* addendT,
                         const void* toep_flip, void* dx, float* dw, int N, int T, int C, int K, int S,
                         lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || xT == nullptr || dyT == nullptr || w == nullptr || dx == nullptr || dw == nullptr)
    return LASR_ERR_BAD_SHAPE;
  return dwconv_cm_bwd(xT, dyT, w, addend, addendT, toep_flip, dx, dw, N, T, C, K, S, stream);
}

}  // extern "C"
