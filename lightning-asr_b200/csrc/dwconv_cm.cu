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
constexpr int CM_THREADS = 32 * 10;

struct DwCmParams {
  const float* w;
  __nv_bfloat16* y;
  const __nv_bfloat16* addend;
  int N, T, C, K, KS, flip, delta;
  int t_chunks, pair_utt, pairs_per_utt;
  int items_per_cg, num_cg, ctas_per_cg, stages, w_early;
};

int lasr_cm_offset_host(int K) { return (K / 2 + 7) / 8 * 8; }
int lasr_cm_pitch_host(int T, int K) {
  const int off = lasr_cm_offset_host(K);
  const int t_chunks = cdiv(T, CM_SL);
  if (t_chunks == 1) return cdiv(off + T + 48, 128) * 128;
  return CM_SL * (cdiv(t_chunks, 2) * 2 - 1) + CM_SLOTF;
}

template <bool HAS_ADDEND>
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
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cg = bid % p.num_cg;
  const int first = bid / p.num_cg;
  const int c0 = cg * DT_CG;

  if (warp_idx == 9 && lane == 0) tma_prefetch_desc(tmap);
  if (!p.w_early) pdl_wait();
  for (int i = threadIdx.x; i < DT_CG * p.K; i += CM_THREADS) {
    const int c = i / p.K, j = i - c * p.K;
    s_w[c * DT_MAX_KS + j] = c0 + c < p.C ? p.w[static_cast<size_t>(c0 + c) * p.K + (p.flip ? p.K - 1 - j : j)] : 0.f;
  }
  if (warp_idx == 8 && lane == 0) {
    for (int st = 0; st < 4; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], 1);
    }
    for (int st = 0; st < 2; ++st) {
      mbar_init(&tmem_full_bar[st], 1);
      mbar_init(&tmem_empty_bar[st], 8);
    }
    mbar_fence_init();
  }
  if (warp_idx == 9) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  __syncthreads();
  // Toeplitz factor B[t, s] = w[s - t - delta], t < 16 output frames of a row, s < KS series positions: K-major
  // no-swizzle cores [s / 8][t / 8], element (t % 8, s % 8).  One 16-byte chunk = the 8 elements of (channel, core, row).
  {
    const int chunks_c = 2 * p.KS;  // per channel: KS/8 x 2 cores x 8 rows
    for (int q = threadIdx.x; q < DT_CG * chunks_c; q += CM_THREADS) {
      const int c = q / chunks_c;
      const int rem = q - c * chunks_c;
      const int core = rem >> 3, r = rem & 7;
      const int j0 = 8 * (core >> 1) - (8 * (core & 1) + r) - p.delta;
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
  pdl_wait();

  if (warp_idx == 9) {
    // ===================== TMA producer: one box = the whole item =====================
    if (lane == 0) {
      int it = 0;
      for (int idx = first; idx < p.items_per_cg; idx += p.ctas_per_cg, ++it) {
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
        tma_load_5d(s_ser + stage * CM_STAGE, tmap, &full_bar[stage], 0, 0, tc, n, c0);
      }
    }
  } else if (warp_idx == 8) {
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
      uint64_t da_c = umma_desc_sw32(smem_u32(s_ser + stage * CM_STAGE), 256);
      uint64_t db_c = umma_desc_none(smem_u32(s_toep), 256, 128);
      uint32_t tmem_d = tmem_base + acc * 256;
      const uint64_t da_step = static_cast<uint64_t>(CM_ROWB >> 4), db_step = static_cast<uint64_t>(toep_bytes_c >> 4);
#pragma unroll 1
      for (int c = 0; c < DT_CG; ++c) {
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
      }
      __syncwarp();
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
#pragma unroll 1
      for (int th = 0; th < 4; ++th) {
        uint32_t v[DT_CG][2];
#pragma unroll
        for (int c = 0; c < DT_CG; ++c) tmem_ld_32x32_x2(taddr + c * 16 + 2 * th, v[c]);
        uint32_t add[2][8];
        if constexpr (HAS_ADDEND) {
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
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool HAS_ADDEND>
__global__ void __launch_bounds__(CM_THREADS, 1)
dwconv_cm_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const DwCmParams p) {
  dwconv_cm_fwd_body<HAS_ADDEND>(&tmap, p, static_cast<int>(blockIdx.x));
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
                      int sms) {
  p.w = w;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.N = N;
  p.T = T;
  p.C = C;
  p.K = K;
  p.flip = flip;
  p.delta = lasr_cm_offset_host(K) - K / 2;
  p.KS = cdiv(K + 15 + p.delta, 16) * 16;
  p.w_early = early_param_loads() ? 1 : 0;
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

int dwconv_cm_fwd(const void* xT, const float* w, void* y, const void* addend, int N, int T, int C, int K, int S, int flip,
                  cudaStream_t stream) {
  if (!dwconv_cm_supported(C, K) || S != lasr_cm_pitch_host(T, K)) return LASR_ERR_BAD_SHAPE;
  DwCmParams p{};
  cm_params(p, w, y, addend, N, T, C, K, flip, kNumSMs);
  CUtensorMap tm;
  if (int rc = cm_item_tmap(&tm, xT, N, C, S, p.t_chunks, p.pair_utt)) return rc;
  const int smem = 1024 + p.stages * CM_STAGE + DT_CG * p.KS * 32 + DT_CG * DT_MAX_KS * 4 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_cm_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dwconv_cm_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  const dim3 grid(p.num_cg * p.ctas_per_cg);
  if (addend != nullptr)
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_fwd_kernel<true>, grid, dim3(CM_THREADS), smem, stream, tm, p));
  else
    LASR_CHECK_PDL(launch_pdl(2, dwconv_cm_fwd_kernel<false>, grid, dim3(CM_THREADS), smem, stream, tm, p));
  return LASR_OK;
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_cm_offset(int K) { return lasr_cm_offset_host(K); }
int lasr_cm_pitch(int T, int K) { return lasr_cm_pitch_host(T, K); }

int lasr_dwconv1d_fwd_cm(const void* xT, const float* w, void* y, const void* addend, int N, int T, int C, int K, int S,
                         int flip, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || xT == nullptr || y == nullptr || w == nullptr) return LASR_ERR_BAD_SHAPE;
  return dwconv_cm_fwd(xT, w, y, addend, N, T, C, K, S, flip, stream);
}

}  // extern "C"
