// Pointwise (1x1) convolution as a tcgen05 / TMEM GEMM for sm_100a.
//
//   D[M, N] (+)= A[M, K] * B[N, K]^T          bf16 operands, fp32 accumulation in tensor memory
//
// K-major mode (forward / data-gradient):  A = activations [frames, Cin] (channels-last, so K is the
//   contiguous axis), B = weights [Cout, Cin].  Both operands are TMA-loaded as 128-byte-swizzled
//   {64 x rows} boxes and consumed by tcgen05.mma through K-major shared-memory descriptors.
// MN-major mode (weight-gradient):  dW[Cout, Cin] = dY[frames, Cout]^T * X[frames, Cin]; the reduction
//   runs over frames, so both operands are "MN-major" (the M / N axis is the contiguous one).  They are
//   loaded as {64 channels x 64 frames} swizzled boxes and described to the MMA unit with MN-major
//   descriptors; the frame axis is split over CTAs and partial products are reduced with fp32 vector RED.
//
// Structure: persistent, warp-specialised CTA of 256 threads, one CTA per SM.
//   warp 0   TMA producer (one elected lane)      smem ring of STAGES x {A 16 KB, B BN*128 B}
//   warp 1   MMA issuer   (one elected lane)      4 x tcgen05.mma (K=16) per 64-wide k-block
//   warp 2   TMEM allocator / deallocator         2 accumulator buffers of BN columns each
//   warps 4-7 epilogue: tcgen05.ld 32x32b -> registers -> bias / MaskCNN row mask / BatchNorm partial
//            statistics (warp butterfly transpose-reduce) -> bf16|fp32 global stores, overlapped with the
//            next tile's MMAs through the double-buffered accumulator.
//
// Replaces: nn.Conv1d(Cin, Cout, kernel_size=1) call sites models/QuartNet.py:22-23,31 / :62-63 / :145-146 / :275
// and MaskCNN (:309-321) + the statistics pass of nn.BatchNorm1d (:24,35) which are folded into the epilogue.
#include "common.cuh"

#include <cstdlib>

namespace lasr {

struct GemmTcParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int k_splits, kb_per_split;
  void* out;
  int ldc;
  int out_f32;
  int vec_ok;
  const float* bias;
  const int32_t* lengths;
  int T;
  double* stats;  // [2, N] column sum / sum of squares, accumulated with RED.f64 (caller zeroes)
  int groups;     // 1, or 2: two independent problems of identical shape share one launch (second set of tensor maps)
};

template <int BN>
struct GemmTcCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // weight-gradient epilogue: per epilogue warp two 32 x 32 fp32 staging tiles handed to the TMA reduce-add unit
  static constexpr int RED_STG_BYTES = 4 * 2 * 4096;
  static constexpr int SMEM_BYTES_RED = SMEM_BYTES + RED_STG_BYTES;
};

// warp-level transpose-reduce: on entry lane r holds f[0..31] (row r, 32 columns); on exit every lane j
// returns sum over the 32 rows of column j.  31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_column_sums(const float (&f)[32], int lane) {
  float a[16];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float send = hi ? f[i] : f[i + 16];
      float keep = hi ? f[i + 16] : f[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  float b[8];
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? a[i] : a[i + 8];
      float keep = hi ? a[i + 8] : a[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  float c[4];
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? b[i] : b[i + 4];
      float keep = hi ? b[i + 4] : b[i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  float d[2];
  {
    const bool hi = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? c[i] : c[i + 2];
      float keep = hi ? c[i + 2] : c[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  const bool hi = lane & 1;
  float send = hi ? d[0] : d[1];
  float keep = hi ? d[1] : d[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// EPI: 0 = store (bias / mask / stats), 1 = fp32 RED accumulate (split-K weight gradient)
template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_a2,
               const __grid_constant__ CUtensorMap tma_b2, const __grid_constant__ CUtensorMap tma_c2,
               const GemmTcParams p) {
  using Cfg = GemmTcCfg<BN>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr int A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;

  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte aligned bases
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int RED_STG = (EPI == 1) ? Cfg::RED_STG_BYTES : 0;
  uint8_t* s_red = smem + STAGES * STAGE_BYTES;  // [4 warps][2][32 rows x 128 B], 1024-byte aligned (EPI == 1)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + RED_STG);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (EPI == 1) tma_prefetch_desc(&tma_c);
    if (p.groups > 1) {
      tma_prefetch_desc(&tma_a2);
      tma_prefetch_desc(&tma_b2);
      tma_prefetch_desc(&tma_c2);
    }
  }
  if (warp_idx == 1 && lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 4);
    mbar_init(&tmem_empty_bar[1], 4);
    mbar_fence_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // prologue above overlapped the previous kernel; operands and outputs are touched only from here on

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  const int group_tiles = tiles_mn * p.k_splits;       // tiles of one problem
  const int num_tiles = group_tiles * (p.groups > 1 ? 2 : 1);

  if (warp_idx == 0) {
    // ===================== TMA producer (converged warp, elected lane issues) =====================
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int grp = tile >= group_tiles ? 1 : 0;
        const int tg = tile - grp * group_tiles;
        const CUtensorMap* ma = grp ? &tma_a2 : &tma_a;
        const CUtensorMap* mb = grp ? &tma_b2 : &tma_b;
        const int split = tg / tiles_mn;
        const int mn = tg - split * tiles_mn;
        const int m_blk = mn / p.num_n_blocks;
        const int n_blk = mn - m_blk * p.num_n_blocks;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d(sa, ma, &full_bar[stage], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c)
                tma_load_2d(sa + c * (64 * BK * 2), ma, &full_bar[stage], m_blk * BM + c * 64, kb * BK);
            }
            if constexpr (!B_MN) {
              tma_load_2d(sb, mb, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d(sb + c * (64 * BK * 2), mb, &full_bar[stage], n_blk * BN + c * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer: the whole warp stays converged (uniform loops, descriptors in uniform
    // registers), one elected lane issues =====================
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int local_tile = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
        const int tg = tile >= group_tiles ? tile - group_tiles : tile;
        const int split = tg / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        const int acc = local_tile & 1;
        const uint32_t acc_phase = (local_tile >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major, SW128: 8-row groups 1024 B apart; 16 bf16 of K = 32 B inside the swizzle atom.
            // MN-major, SW128: 64-channel chunks 8192 B apart (LBO), 8-row (K) groups 1024 B apart (SBO);
            // 16 rows of K = 2048 B.
            const uint64_t da = A_MN ? umma_desc_sw128(sa + k * 2048, 64 * BK * 2, 1024)
                                     : umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, 64 * BK * 2, 1024)
                                     : umma_desc_sw128(sb + k * 32, 16, 1024);
            if (leader) umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (leader) umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (leader) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int ew = warp_idx - 4;  // == warp_idx % 4: the TMEM lane quarter this warp may access
    int local_tile = 0;
    int nred = 0;  // staging tiles this warp has handed to the TMA reduce unit (EPI == 1)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
      const int grp = tile >= group_tiles ? 1 : 0;
      const int tg = tile - grp * group_tiles;
      const CUtensorMap* mc = grp ? &tma_c2 : &tma_c;
      const int split = tg / tiles_mn;
      const int mn = tg - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int acc = local_tile & 1;
      const uint32_t acc_phase = (local_tile >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();

      const int row = m_blk * BM + ew * 32 + lane;
      const bool row_ok = row < p.M;
      bool keep = true;
      if (EPI == 0 && p.lengths != nullptr && row_ok) {
        const int n = row / p.T;
        keep = (row - n * p.T) < p.lengths[n];
      }
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int col0 = n_blk * BN + ch * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(taddr0 + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);

        if constexpr (EPI == 0) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += (col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
          }
          if (!keep || !row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
          if (p.stats != nullptr) {
            const float s = warp_column_sums(f, lane);
            float q[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) q[j] = f[j] * f[j];
            const float ss = warp_column_sums(q, lane);
            if (col0 + lane < p.N) {  // one 128-byte RED transaction per warp and statistic
              red_add_f64(p.stats + col0 + lane, static_cast<double>(s));
              red_add_f64(p.stats + p.N + col0 + lane, static_cast<double>(ss));
            }
          }
          if (row_ok) {
            const bool full = (col0 + 32 <= p.N) && p.vec_ok;
            if (p.out_f32) {
              float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldc + col0;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) dst[j] = f[j];
              }
            } else {
              __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldc + col0;
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 u;
                  u.x = f32x2_to_bf16x2(f[j], f[j + 1]);
                  u.y = f32x2_to_bf16x2(f[j + 2], f[j + 3]);
                  u.z = f32x2_to_bf16x2(f[j + 4], f[j + 5]);
                  u.w = f32x2_to_bf16x2(f[j + 6], f[j + 7]);
                  *reinterpret_cast<uint4*>(dst + j) = u;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) dst[j] = __float2bfloat16_rn(f[j]);
              }
            }
          }
        } else {
          if (p.vec_ok) {
            // split-K reduction through the TMA unit: the warp's 32 x 32 fp32 sub-tile goes to a 128B-swizzled staging
            // buffer and ONE bulk reduce-add covers it with full 128-byte lines (a per-thread RED.v4 touches 32
            // half-used sectors per instruction and 150-300 CTAs walk the same addresses in the same order); rows
            // and columns outside [M, N) are clipped by the tensor map
            uint8_t* buf = s_red + ew * 8192 + (nred & 1) * 4096;
            if (nred >= 2) {
              if (lane == 0) tma_store_wait_read<1>();  // the reduce issued from this buffer two chunks ago has read it
              __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                  make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(mc, buf, col0, m_blk * BM + ew * 32);
              tma_store_commit();
            }
            ++nred;
          } else if (row_ok) {
            float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(dst + j, f[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (EPI == 1 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight-stationary forward / data-gradient kernel (bf16 out).
//
// The 1x1 convs of this network are short-K GEMMs (K = 64..512) over ~25 k frames: with a streamed B operand every
// 128-frame tile re-loads the whole weight slice (2x the bytes of the activation tile itself) and the L2 -> SM path,
// not the tensor pipe, sets the pace.  Here a CTA owns ONE 128-column slice of the output for its whole life:
//   - its weight slice [128 x K] (<= 128 KB) is TMA-loaded once and stays in shared memory;
//   - only activation tiles stream through the TMA ring (CTAs that share a row block run adjacently -> L2 hits);
//   - the epilogue warps pack the accumulator to bf16 into a 128B-swizzled staging tile that ONE thread hands to
//     the TMA store unit (full-line, asynchronous stores);
//   - four more warps read the same staging tile column-wise and keep the BatchNorm sum / sum of squares of the
//     CTA's 128 columns in registers across all its tiles: one fp64 RED per column and CTA at the very end
//     (the per-tile RED.f64 of the streamed kernel serialised ~800 atomics on every statistics address).
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-7 epilogue (TMEM -> staging), 8-11 statistics.
// ------------------------------------------------------------------------------------------------
struct GemmWsParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int ctas_per_n;
  int stages;
  int cluster;  // CTAs per cluster = num_n_blocks when the activation tiles are multicast, else 1
  int w_early;  // weight slice may be fetched before griddepcontrol.wait (lasr_set_early_param_loads)
  // fused inference epilogue (lasr_pwconv_fwd_fused): y = act(mask(acc) + bias + residual)
  const __nv_bfloat16* residual;  // [M, ld_res] or NULL
  int ld_res;
  int relu;
  int bias_after_mask;  // 1: masked rows keep bias (+ residual): BatchNorm folded into w / bias sees mask(pw) = 0
  // grouped launch: two problems of identical shape (a block's pointwise and residual conv); CTAs >= group_ctas work
  // on problem 2 with the second set of tensor maps and its own mask / statistics pointers
  int groups;
  int group_ctas;
  const int32_t* lengths2;
  double* stats2;
  const float* bias;
  const int32_t* lengths;
  int T;
  double* stats;
  unsigned long long* trace;  // debug timeline (tools/trace_gemm.py), normally NULL
  // TR (series output, data gradients only): tiles are 128-POSITION blocks of the channel-major series layout
  // (include/lasr.h): block b = (utterance b / tr_bpu, block b % tr_bpu) covers frames [128 j - tr_off, +128) of the
  // utterance (rows outside [0, T) are zero-filled by TMA), and the result is written transposed, [channel][position]
  int tr_bpu, tr_off, tr_S, tr_blocks;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define WS_TRACE(slot, idx)                                                                       \
  do {                                                                                            \
    if (p.trace != nullptr && (idx) < 16) p.trace[(blockIdx.x * 8 + (slot)) * 16 + (idx)] = gtimer(); \
  } while (0)

constexpr int WS_BM = 128, WS_BN = 128, WS_BK = 64;
constexpr int WS_A_BYTES = WS_BM * WS_BK * 2;        // 16 KB per ring stage
constexpr int WS_WKB_BYTES = WS_BN * WS_BK * 2;      // 16 KB of the weight slice per k-block
constexpr int WS_STG_BYTES = WS_BM * WS_BN * 2;      // 32 KB staging tile (two 64-column sub-tiles)
constexpr int WS_MAX_STAGES = 8;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool B_MN, bool TR = false>
__global__ void __launch_bounds__(384, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_a2,
                const __grid_constant__ CUtensorMap tma_b2, const __grid_constant__ CUtensorMap tma_c2,
                const GemmWsParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* s_w = smem;                                        // [num_k_blocks][16 KB]
  uint8_t* s_a = s_w + p.num_k_blocks * WS_WKB_BYTES;         // [stages][16 KB]
  uint8_t* s_stg = s_a + p.stages * WS_A_BYTES;               // [2][128 rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stg + WS_STG_BYTES);
  uint64_t* empty_bar = full_bar + WS_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WS_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = (p.groups > 1 && static_cast<int>(blockIdx.x) >= p.group_ctas) ? 1 : 0;
  const int bid = static_cast<int>(blockIdx.x) - grp * p.group_ctas;
  const CUtensorMap* ma = grp ? &tma_a2 : &tma_a;
  const CUtensorMap* mb = grp ? &tma_b2 : &tma_b;
  const CUtensorMap* mc = grp ? &tma_c2 : &tma_c;
  const int32_t* lengths_g = grp ? p.lengths2 : p.lengths;
  double* stats_g = grp ? p.stats2 : p.stats;
  const int n_blk = bid % p.num_n_blocks;
  const int m_first = bid / p.num_n_blocks;
  if (threadIdx.x == 0) WS_TRACE(0, 0);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(ma);
    tma_prefetch_desc(mb);
    tma_prefetch_desc(mc);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], p.cluster);  // every CTA of the cluster must have consumed the slot
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 4);
    mbar_init(&tmem_empty_bar[1], 4);
    mbar_init(w_bar, 1);
    mbar_fence_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, 2 * WS_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();  // peers' barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cluster) - 1u);
  const int slice_rows = WS_BM / p.cluster;
  const int crank = p.cluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  if (threadIdx.x == 0) WS_TRACE(0, 1);
  // Programmatic dependent launch: everything above, and the weight-slice load below, overlaps the tail of the previous
  // kernel.  Weights are safe to read early: parameters and their bf16 shadows are only written by kernels that never
  // trigger their dependents early (cast_weight, the optimizer), i.e. behind a full stream-order barrier.
  if (warp_idx != 0) pdl_wait();

  if (warp_idx == 0) {
    // ===================== TMA producer (converged warp, elected lane issues) =====================
    {
      const bool leader = elect_one();
      if (!p.w_early) pdl_wait();
      // the CTA's weight slice, once
      if (leader) mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(p.num_k_blocks) * WS_WKB_BYTES);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        uint8_t* dst = s_w + kb * WS_WKB_BYTES;
        if (leader) {
          if constexpr (!B_MN) {
            tma_load_2d(dst, mb, w_bar, kb * WS_BK, n_blk * WS_BN);
          } else {
#pragma unroll
            for (int c = 0; c < WS_BN / 64; ++c)
              tma_load_2d(dst + c * (64 * WS_BK * 2), mb, w_bar, n_blk * WS_BN + c * 64, kb * WS_BK);
          }
        }
      }
      __syncwarp();
      pdl_wait();  // activations come from the previous kernel
      int stage = 0;
      uint32_t phase = 0;
      for (int m_blk = m_first; m_blk < p.num_m_blocks; m_blk += p.ctas_per_n) {
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], WS_A_BYTES);
            if constexpr (TR) {
              const int un = m_blk / p.tr_bpu, uj = m_blk - un * p.tr_bpu;
              tma_load_3d(s_a + stage * WS_A_BYTES, ma, &full_bar[stage], kb * WS_BK, uj * WS_BM - p.tr_off, un);
            } else if (p.cluster == 1) {
              tma_load_2d(s_a + stage * WS_A_BYTES, ma, &full_bar[stage], kb * WS_BK, m_blk * WS_BM);
            } else {
              // this CTA fetches its slice of rows and multicasts it; the peers deliver the other slices
              tma_load_2d_mc(s_a + stage * WS_A_BYTES + crank * slice_rows * 128, ma, &full_bar[stage],
                             kb * WS_BK, m_blk * WS_BM + crank * slice_rows, mc_mask);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (converged warp, elected lane) =====================
    {
      const bool leader = elect_one();
      // TR: operands swapped -- A = the weight slice (MN-major), B = the activation tile (K-major): D[channel, frame]
      constexpr uint32_t idesc = TR ? umma_idesc_bf16(WS_BN, WS_BM, 1, 0) : umma_idesc_bf16(WS_BM, WS_BN, 0, B_MN ? 1 : 0);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      if (leader) WS_TRACE(0, 2);
      int stage = 0;
      uint32_t phase = 0;
      int local_tile = 0;
      for (int m_blk = m_first; m_blk < p.num_m_blocks; m_blk += p.ctas_per_n, ++local_tile) {
        const int acc = local_tile & 1;
        const uint32_t acc_phase = (local_tile >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        if (leader) WS_TRACE(1, local_tile);
        const uint32_t tmem_d = tmem_base + acc * WS_BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0 && leader) WS_TRACE(2, local_tile);
          const uint32_t sa = smem_u32(s_a + stage * WS_A_BYTES);
          const uint32_t sb = smem_u32(s_w + kb * WS_WKB_BYTES);
#pragma unroll
          for (int k = 0; k < WS_BK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, 64 * WS_BK * 2, 1024)
                                     : umma_desc_sw128(sb + k * 32, 16, 1024);
            if (leader) {
              if constexpr (TR)
                umma_bf16(tmem_d, db, da, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (leader) {
            if (p.cluster == 1)
              umma_commit(&empty_bar[stage]);
            else
              umma_commit_mc(&empty_bar[stage], mc_mask);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (leader) {
          umma_commit(&tmem_full_bar[acc]);
          WS_TRACE(3, local_tile);
        }
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    // ===================== epilogue: TMEM -> bf16 staging tile -> TMA store =====================
    const int ew = warp_idx - 4;
    const int r = ew * 32 + lane;  // row inside the tile
    int local_tile = 0;
    for (int m_blk = m_first; m_blk < p.num_m_blocks; m_blk += p.ctas_per_n, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = (local_tile >> 1) & 1;
      const int row = m_blk * WS_BM + r;
      bool keep = row < p.M;
      if (keep && lengths_g != nullptr) {
        const int n = row / p.T;
        keep = (row - n * p.T) < lengths_g[n];
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (threadIdx.x == 128) WS_TRACE(4, local_tile);
      // staging tile free: the previous TMA store has read it and the statistics warps are done with it
      if (threadIdx.x == 128) tma_store_wait_read<0>();
      named_bar_sync(1, 256);
      if (threadIdx.x == 128) WS_TRACE(5, local_tile);
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * WS_BN;
      if constexpr (TR) {
        // lane = channel, columns = the tile's 128 positions: 8 consecutive positions are one 16-byte group, stored at its
        // pre-swizzled group index (g ^ ((g >> 3) & 1), include/lasr.h) inside the 128-byte-swizzled staging tile
#pragma unroll
        for (int ch = 0; ch < WS_BM / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(taddr0 + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
            u.y = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
            u.z = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
            u.w = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
            const int gl = ch * 4 + q;                 // group inside the 128-position block
            const int glp = gl ^ ((gl >> 3) & 1);
            uint8_t* sub = s_stg + (glp >> 3) * (WS_BM * 128) + r * 128;
            *reinterpret_cast<uint4*>(sub + (((glp & 7) ^ (r & 7)) << 4)) = u;
          }
        }
      } else {
#pragma unroll
      for (int ch = 0; ch < WS_BN / 32; ++ch) {
        const int col0 = n_blk * WS_BN + ch * 32;
        // residual row segment of the fused inference epilogue: requested before the accumulator load so that its
        // latency hides under tcgen05.ld
        uint4 rr[4];
        const bool use_res = p.residual != nullptr && row < p.M && col0 + 32 <= p.N;
        if (use_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ld_res + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) rr[q] = __ldg(rp + q);
        }
        uint32_t v[32];
        tmem_ld_32x32(taddr0 + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias_after_mask && !keep) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.f;
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += (col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
        }
        if (!p.bias_after_mask && !keep) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.f;
        }
        if (use_res) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = rr[q];
            const float2 a0 = bf16x2_to_f32x2(u.x), a1 = bf16x2_to_f32x2(u.y), a2 = bf16x2_to_f32x2(u.z),
                         a3 = bf16x2_to_f32x2(u.w);
            f[8 * q + 0] += a0.x; f[8 * q + 1] += a0.y; f[8 * q + 2] += a1.x; f[8 * q + 3] += a1.y;
            f[8 * q + 4] += a2.x; f[8 * q + 5] += a2.y; f[8 * q + 6] += a3.x; f[8 * q + 7] += a3.y;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        uint8_t* sub = s_stg + (ch >> 1) * (WS_BM * 128) + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = f32x2_to_bf16x2(f[8 * q + 0], f[8 * q + 1]);
          u.y = f32x2_to_bf16x2(f[8 * q + 2], f[8 * q + 3]);
          u.z = f32x2_to_bf16x2(f[8 * q + 4], f[8 * q + 5]);
          u.w = f32x2_to_bf16x2(f[8 * q + 6], f[8 * q + 7]);
          const int c16 = (ch & 1) * 4 + q;  // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(sub + ((c16 ^ (r & 7)) << 4)) = u;
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      fence_proxy_async_smem();
      named_bar_sync(2, 256);  // staging tile complete
      if (threadIdx.x == 128) WS_TRACE(6, local_tile);
      if (threadIdx.x == 128) {
        if constexpr (TR) {
          const int un = m_blk / p.tr_bpu, uj = m_blk - un * p.tr_bpu;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_store_2d(mc, s_stg + j * (WS_BM * 128), un * p.tr_S + uj * WS_BM + j * 64, n_blk * WS_BN);
        } else {
#pragma unroll
          for (int j = 0; j < WS_BN / 64; ++j)
            tma_store_2d(mc, s_stg + j * (WS_BM * 128), n_blk * WS_BN + j * 64, m_blk * WS_BM);
        }
        tma_store_commit();
      }
    }
    if (threadIdx.x == 128) tma_store_wait<0>();
    if (threadIdx.x == 128) WS_TRACE(0, 3);
  } else if (warp_idx >= 8) {
    // ===================== BatchNorm statistics from the staging tile =====================
    const int sw = warp_idx - 8;
    const int sub = sw & 1;          // 64-column sub-tile
    const int r0 = (sw >> 1) * 64;   // row half
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const bool want = stats_g != nullptr;
    for (int m_blk = m_first; m_blk < p.num_m_blocks; m_blk += p.ctas_per_n) {
      named_bar_sync(1, 256);
      named_bar_sync(2, 256);
      if (want) {
        const uint8_t* base = s_stg + sub * (WS_BM * 128);
#pragma unroll 8
        for (int rr = 0; rr < 64; ++rr) {
          const int r = r0 + rr;
          const uint32_t w =
              *reinterpret_cast<const uint32_t*>(base + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
          const float2 v = bf16x2_to_f32x2(w);
          s0 += v.x;
          s1 += v.y;
          q0 = fmaf(v.x, v.x, q0);
          q1 = fmaf(v.y, v.y, q1);
        }
      }
    }
    if (want) {
      const int col = n_blk * WS_BN + sub * 64 + 2 * lane;
      if (col < p.N) {
        red_add_f64(stats_g + col, static_cast<double>(s0));
        red_add_f64(stats_g + p.N + col, static_cast<double>(q0));
      }
      if (col + 1 < p.N) {
        red_add_f64(stats_g + col + 1, static_cast<double>(s1));
        red_add_f64(stats_g + p.N + col + 1, static_cast<double>(q1));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * WS_BN);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the weight-stationary kernel (cta_group::2): the two SMs of a TPC work on one 256-row x 256-column
// output tile.  Each CTA streams ITS 128 activation rows and keeps ITS 128 weight rows (half of the pair's 256-column
// slice) resident; the pair-MMA (M = 256, N = 256, issued by the leader CTA) reads both halves of B across the pair,
// so every activation tile fetched from L2 now feeds 256 output columns instead of 128 -- the L2 -> SM activation
// traffic that paces the single-CTA kernel (tile period = activation tile / ~57 GB/s per SM) is halved.
// Barriers: TMA loads of both CTAs count bytes on the LEADER's full / weight barriers; the leader's MMA commits
// multicast to both CTAs' empty and accumulator-full barriers; both CTAs' epilogue warps arrive on the leader's
// accumulator-empty barrier.  Per CTA: 128 rows x 256 fp32 columns x 2 buffers = all 512 TMEM columns.
// ------------------------------------------------------------------------------------------------
template <bool B_MN, bool TR = false>
__global__ void __launch_bounds__(384, 1)
gemm_ws2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_a2,
                const __grid_constant__ CUtensorMap tma_b2, const __grid_constant__ CUtensorMap tma_c2,
                const GemmWsParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* s_w = smem;                                        // [num_k_blocks][16 KB]: this CTA's 128 weight rows
  uint8_t* s_a = s_w + p.num_k_blocks * WS_WKB_BYTES;         // [stages][16 KB]: this CTA's 128 activation rows
  uint8_t* s_stg = s_a + p.stages * WS_A_BYTES;               // [2][128 rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stg + WS_STG_BYTES);
  uint64_t* empty_bar = full_bar + WS_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WS_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool is_leader = crank == 0;
  const int grp = (p.groups > 1 && static_cast<int>(blockIdx.x) >= p.group_ctas) ? 1 : 0;
  const CUtensorMap* ma = grp ? &tma_a2 : &tma_a;
  const CUtensorMap* mb = grp ? &tma_b2 : &tma_b;
  const CUtensorMap* mc = grp ? &tma_c2 : &tma_c;
  const int32_t* lengths_g = grp ? p.lengths2 : p.lengths;
  double* stats_g = grp ? p.stats2 : p.stats;
  const int pair = (static_cast<int>(blockIdx.x) - grp * p.group_ctas) >> 1;
  const int n_blk2 = pair % p.num_n_blocks;     // 256-column slice of the pair
  const int m_first = pair / p.num_n_blocks;    // 256-row tiles m_first, m_first + ctas_per_n, ...
  const int n0 = n_blk2 * 256;                  // first output column of the pair
  const int nw0 = n0 + crank * WS_BN;           // first weight row (output column) this CTA keeps

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(ma);
    tma_prefetch_desc(mb);
    tma_prefetch_desc(mc);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's: its producer's arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(&empty_bar[s], 1);  // each CTA's: the leader's multicast commit
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 8);  // leader's: 4 epilogue warps of each CTA
    mbar_init(&tmem_empty_bar[1], 8);
    mbar_init(w_bar, 1);
    mbar_fence_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before any remote arrive / byte count
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp_idx != 0) pdl_wait();

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs; bytes are counted on the leader's barriers) ==================
    const bool leader_lane = elect_one();
    if (!p.w_early) pdl_wait();
    if (leader_lane && is_leader) mbar_arrive_expect_tx(w_bar, 2u * static_cast<uint32_t>(p.num_k_blocks) * WS_WKB_BYTES);
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      uint8_t* dst = s_w + kb * WS_WKB_BYTES;
      if (leader_lane) {
        if constexpr (!B_MN) {
          tma_load_2d_2sm(dst, mb, w_bar, kb * WS_BK, nw0);
        } else {
#pragma unroll
          for (int c = 0; c < WS_BN / 64; ++c)
            tma_load_2d_2sm(dst + c * (64 * WS_BK * 2), mb, w_bar, nw0 + c * 64, kb * WS_BK);
        }
      }
    }
    __syncwarp();
    pdl_wait();
    int stage = 0;
    uint32_t phase = 0;
    for (int m2 = m_first; m2 < p.num_m_blocks; m2 += p.ctas_per_n) {
      const int row0 = m2 * 256 + crank * WS_BM;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (leader_lane) {
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * WS_A_BYTES);
          if constexpr (TR) {
            // this CTA's 128-position block of the pair tile (a block past the end: utterance index out of range, the
            // whole box is zero-filled)
            const int ub = 2 * m2 + crank;
            const int un = ub / p.tr_bpu, uj = ub - un * p.tr_bpu;
            tma_load_3d_2sm(s_a + stage * WS_A_BYTES, ma, &full_bar[stage], kb * WS_BK, uj * WS_BM - p.tr_off, un);
          } else {
            tma_load_2d_2sm(s_a + stage * WS_A_BYTES, ma, &full_bar[stage], kb * WS_BK, row0);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (is_leader) {
      const bool leader_lane = elect_one();
      // TR: operands swapped -- A = the pair's 256 weight rows (MN-major), B = its 256 positions (K-major)
      constexpr uint32_t idesc = TR ? umma_idesc_bf16(256, 256, 1, 0) : umma_idesc_bf16(256, 256, 0, B_MN ? 1 : 0);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int local_tile = 0;
      for (int m2 = m_first; m2 < p.num_m_blocks; m2 += p.ctas_per_n, ++local_tile) {
        const int acc = local_tile & 1;
        const uint32_t acc_phase = (local_tile >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(s_a + stage * WS_A_BYTES);
          const uint32_t sb = smem_u32(s_w + kb * WS_WKB_BYTES);
#pragma unroll
          for (int k = 0; k < WS_BK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, 64 * WS_BK * 2, 1024)
                                     : umma_desc_sw128(sb + k * 32, 16, 1024);
            if (leader_lane) {
              if constexpr (TR)
                umma_bf16_2sm(tmem_d, db, da, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else
                umma_bf16_2sm(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          if (leader_lane) umma_commit_2sm(&empty_bar[stage]);
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (leader_lane) umma_commit_2sm(&tmem_full_bar[acc]);
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    // ===================== epilogue: this CTA's 128 rows x 256 columns, two 128-column passes ================
    const int ew = warp_idx - 4;
    const int r = ew * 32 + lane;
    int local_tile = 0;
    for (int m2 = m_first; m2 < p.num_m_blocks; m2 += p.ctas_per_n, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t acc_phase = (local_tile >> 1) & 1;
      const int row = m2 * 256 + crank * WS_BM + r;
      bool keep = row < p.M;
      if (keep && lengths_g != nullptr) {
        const int n = row / p.T;
        keep = (row - n * p.T) < lengths_g[n];
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (threadIdx.x == 128) tma_store_wait_read<0>();
        named_bar_sync(1, 256);
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 256 + h * WS_BN;
        if constexpr (TR) {
          // lanes = this CTA's 128 channels (weight rows nw0 ..), columns [128 h, +128) = the positions of block 2 m2 + h
#pragma unroll
          for (int ch = 0; ch < WS_BM / 32; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(taddr0 + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
              u.y = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
              u.z = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
              u.w = f32x2_to_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
              const int gl = ch * 4 + q;
              const int glp = gl ^ ((gl >> 3) & 1);
              uint8_t* sub = s_stg + (glp >> 3) * (WS_BM * 128) + r * 128;
              *reinterpret_cast<uint4*>(sub + (((glp & 7) ^ (r & 7)) << 4)) = u;
            }
          }
        } else {
#pragma unroll
        for (int ch = 0; ch < WS_BN / 32; ++ch) {
          const int col0 = n0 + h * WS_BN + ch * 32;
          uint4 rr[4];
          const bool use_res = p.residual != nullptr && row < p.M && col0 + 32 <= p.N;
          if (use_res) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(row) * p.ld_res + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) rr[q] = __ldg(rp + q);
          }
          uint32_t v[32];
          tmem_ld_32x32(taddr0 + ch * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias_after_mask && !keep) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += (col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.f;
          }
          if (!p.bias_after_mask && !keep) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
          if (use_res) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 u = rr[q];
              const float2 a0 = bf16x2_to_f32x2(u.x), a1 = bf16x2_to_f32x2(u.y), a2 = bf16x2_to_f32x2(u.z),
                           a3 = bf16x2_to_f32x2(u.w);
              f[8 * q + 0] += a0.x; f[8 * q + 1] += a0.y; f[8 * q + 2] += a1.x; f[8 * q + 3] += a1.y;
              f[8 * q + 4] += a2.x; f[8 * q + 5] += a2.y; f[8 * q + 6] += a3.x; f[8 * q + 7] += a3.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint8_t* sub = s_stg + (ch >> 1) * (WS_BM * 128) + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = f32x2_to_bf16x2(f[8 * q + 0], f[8 * q + 1]);
            u.y = f32x2_to_bf16x2(f[8 * q + 2], f[8 * q + 3]);
            u.z = f32x2_to_bf16x2(f[8 * q + 4], f[8 * q + 5]);
            u.w = f32x2_to_bf16x2(f[8 * q + 6], f[8 * q + 7]);
            const int c16 = (ch & 1) * 4 + q;
            *reinterpret_cast<uint4*>(sub + ((c16 ^ (r & 7)) << 4)) = u;
          }
        }
        }
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        if (threadIdx.x == 128) {
          if constexpr (TR) {
            const int ub = 2 * m2 + h;
            if (ub < p.tr_blocks) {
              const int un = ub / p.tr_bpu, uj = ub - un * p.tr_bpu;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                tma_store_2d(mc, s_stg + j * (WS_BM * 128), un * p.tr_S + uj * WS_BM + j * 64, nw0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < WS_BN / 64; ++j)
              tma_store_2d(mc, s_stg + j * (WS_BM * 128), n0 + h * WS_BN + j * 64, m2 * 256 + crank * WS_BM);
          }
          tma_store_commit();
        }
      }
    }
    if (threadIdx.x == 128) tma_store_wait<0>();
  } else if (warp_idx >= 8) {
    // ===================== BatchNorm statistics from the staging tile (per 128-column pass) =====================
    const int sw = warp_idx - 8;
    const int sub = sw & 1;
    const int r0 = (sw >> 1) * 64;
    float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f}, q0[2] = {0.f, 0.f}, q1[2] = {0.f, 0.f};
    const bool want = stats_g != nullptr;
    for (int m2 = m_first; m2 < p.num_m_blocks; m2 += p.ctas_per_n) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        named_bar_sync(1, 256);
        named_bar_sync(2, 256);
        if (want) {
          const uint8_t* base = s_stg + sub * (WS_BM * 128);
#pragma unroll 8
          for (int rr = 0; rr < 64; ++rr) {
            const int r = r0 + rr;
            const uint32_t w =
                *reinterpret_cast<const uint32_t*>(base + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
            const float2 v = bf16x2_to_f32x2(w);
            s0[h] += v.x;
            s1[h] += v.y;
            q0[h] = fmaf(v.x, v.x, q0[h]);
            q1[h] = fmaf(v.y, v.y, q1[h]);
          }
        }
      }
    }
    if (want) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n0 + h * WS_BN + sub * 64 + 2 * lane;
        if (col < p.N) {
          red_add_f64(stats_g + col, static_cast<double>(s0[h]));
          red_add_f64(stats_g + p.N + col, static_cast<double>(q0[h]));
        }
        if (col + 1 < p.N) {
          red_add_f64(stats_g + col + 1, static_cast<double>(s1[h]));
          red_add_f64(stats_g + p.N + col + 1, static_cast<double>(q1[h]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its partner may still signal it or read its operands
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair weight gradient: dW[Cout, Cin] += dY[frames, Cout]^T X[frames, Cin] with a 256 (Cout) x 256 (Cin) pair tile.
// Each CTA streams ITS 128 Cout columns of dY and ITS 128 Cin columns of X (32 KB per 64-frame k-block instead of the
// 48 KB of the single-CTA 128 x 256 tile) and ends up with its 128 rows x 256 columns of the accumulator, reduced into
// the gradient through the same TMA reduce-add epilogue.  One accumulator buffer (the split-K tiles are one per pair).
// ------------------------------------------------------------------------------------------------
constexpr int WG2_STAGE = 2 * 64 * 128 * 2;  // A 128 x 64 + B 128 x 64 bf16 = 32 KB
constexpr int WG2_STAGES = 6;
constexpr int WG2_SMEM = WG2_STAGES * WG2_STAGE + 32768 + 1024 + 256;

__global__ void __launch_bounds__(256, 1)
gemm_wgrad2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tma_c, const GemmTcParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* s_red = smem + WG2_STAGES * WG2_STAGE;  // [4 warps][2][32 rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_red + 32768);
  uint64_t* empty_bar = full_bar + WG2_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WG2_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool is_leader = crank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < WG2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 8);
    mbar_fence_init();
  }
  if (warp_idx == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, 256);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;  // 256 x 256 pair tiles
  const int num_tiles = tiles_mn * p.k_splits;

  if (warp_idx == 0) {
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * WG2_STAGE;
        uint8_t* sb = sa + 16384;
        if (leader_lane) {
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * WG2_STAGE);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tma_load_2d_2sm(sa + c * 8192, &tma_a, &full_bar[stage], m_blk * 256 + crank * 128 + c * 64, kb * 64);
            tma_load_2d_2sm(sb + c * 8192, &tma_b, &full_bar[stage], n_blk * 256 + crank * 128 + c * 64, kb * 64);
          }
        }
        __syncwarp();
        if (++stage == WG2_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx == 1) {
    if (is_leader) {
      const bool leader_lane = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int local_tile = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local_tile) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        mbar_wait(tmem_empty_bar, (local_tile & 1) ^ 1u);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * WG2_STAGE);
          const uint32_t sb = sa + 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 2048, 64 * 64 * 2, 1024);
            const uint64_t db = umma_desc_sw128(sb + k * 2048, 64 * 64 * 2, 1024);
            if (leader_lane) umma_bf16_2sm(tmem_base, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (leader_lane) umma_commit_2sm(&empty_bar[stage]);
          __syncwarp();
          if (++stage == WG2_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (leader_lane) umma_commit_2sm(tmem_full_bar);
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4) {
    const int ew = warp_idx - 4;
    int local_tile = 0;
    int nred = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local_tile) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      mbar_wait(tmem_full_bar, local_tile & 1);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        const int col0 = n_blk * 256 + ch * 32;
        if (col0 >= p.N) break;
        uint32_t v[32];
        tmem_ld_32x32(taddr0 + ch * 32, v);
        tmem_ld_wait();
        uint8_t* buf = s_red + ew * 8192 + (nred & 1) * 4096;
        if (nred >= 2) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tma_c, buf, col0, m_blk * 256 + crank * 128 + ew * 32);
          tma_store_commit();
        }
        ++nred;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty_bar);
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_bytes,
                      uint32_t box0, uint32_t box1, bool swizzle128) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return LASR_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0) return LASR_ERR_ALIGNMENT;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LASR_OK : LASR_ERR_DRIVER;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_bytes,
                     uint32_t box0, uint32_t box1, bool swizzle128) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return LASR_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0) return LASR_ERR_ALIGNMENT;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LASR_OK : LASR_ERR_DRIVER;
}

// rank-N bf16 tensor map, dim0 contiguous; strides_bytes has rank-1 entries (dims 1..rank-1); 128B swizzle optional
int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool swizzle128) {
  PFN_encodeTiled fn = get_encode_fn();
  if (fn == nullptr) return LASR_ERR_DRIVER;
  if (rank < 1 || rank > 5) return LASR_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return LASR_ERR_ALIGNMENT;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      if (strides_bytes[i - 1] & 15) return LASR_ERR_ALIGNMENT;
      st[i - 1] = strides_bytes[i - 1];
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, st, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LASR_OK : LASR_ERR_DRIVER;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmTcParams& p, int grid,
                       cudaStream_t stream, const CUtensorMap* tc = nullptr, const CUtensorMap* ta2 = nullptr,
                       const CUtensorMap* tb2 = nullptr, const CUtensorMap* tc2 = nullptr) {
  using Cfg = GemmTcCfg<BN>;
  constexpr int kSmem = EPI == 1 ? Cfg::SMEM_BYTES_RED : Cfg::SMEM_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  LASR_CHECK_PDL(launch_pdl(1, gemm_tc_kernel<BN, A_MN, B_MN, EPI>, dim3(grid), dim3(256), kSmem, stream, ta, tb,
                            tc != nullptr ? *tc : ta, ta2 != nullptr ? *ta2 : ta, tb2 != nullptr ? *tb2 : tb,
                            tc2 != nullptr ? *tc2 : ta, p));
  return LASR_OK;
}

static unsigned long long* g_ws_trace = nullptr;
extern "C" void lasr_debug_set_gemm_trace(unsigned long long* buf) { g_ws_trace = buf; }

struct WsFused {
  const void* residual = nullptr;
  int ld_res = 0;
  int relu = 0;
  int bias_after_mask = 0;
  // second problem of a grouped launch (same M, N, K and leading dimensions): activations, weights, output and its
  // own MaskCNN lengths / BatchNorm statistics
  const void* a2 = nullptr;
  const void* b2 = nullptr;
  void* out2 = nullptr;
  const int32_t* lengths2 = nullptr;
  double* stats2 = nullptr;
  // series output (data gradients, b_mn = true): a = dy [sr_n utterances x sr_T frames, K], out(2) = channel-major
  // series tensors [N][sr_n][sr_S] with frame offset sr_off (include/lasr.h)
  int sr_n = 0, sr_T = 0, sr_S = 0, sr_off = 0;
};

// tensor maps of the series mode: activations as {K, T, utterance} (rows outside [0, T) zero-filled), output as
// {sr_n * sr_S positions, N channels}
static int series_maps(CUtensorMap* ta, CUtensorMap* tc, const void* a, void* out, const WsFused& fu, int N, int K,
                       int lda) {
  const uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(fu.sr_T), static_cast<uint64_t>(fu.sr_n)};
  const uint64_t strides[2] = {static_cast<uint64_t>(lda) * 2, static_cast<uint64_t>(lda) * 2 * fu.sr_T};
  const uint32_t box[3] = {64, 128, 1};
  int rc = make_tmap_nd_bf16(ta, a, 3, dims, strides, box, true);
  if (rc) return rc;
  const uint64_t pos = static_cast<uint64_t>(fu.sr_n) * fu.sr_S;
  return make_tmap_2d_bf16(tc, out, pos, N, pos * 2, 64, 128, true);
}

// weight-stationary launch; returns LASR_ERR_UNSUPPORTED when the shape does not qualify (caller falls back)
static int launch_ws(bool b_mn, const void* a, const void* b, void* out, const float* bias, const int32_t* lengths, int T,
                     double* stats, int M, int N, int K, int lda, int ldb, int ldc, cudaStream_t stream,
                     const WsFused& fu = WsFused()) {
  const int kbs = cdiv(K, WS_BK);
  static const bool disabled = getenv("LASR_GEMM_STREAMED") != nullptr;  // A/B switch for profiling
  if (disabled) return LASR_ERR_UNSUPPORTED;
  if (kbs * WS_WKB_BYTES > 131072) return LASR_ERR_UNSUPPORTED;
  if ((ldc % 8) || (reinterpret_cast<uintptr_t>(out) & 15)) return LASR_ERR_UNSUPPORTED;
  const int nblk = cdiv(N, WS_BN);
  // Cluster multicast of the activation tiles (each CTA fetches 1/cluster of the rows for all n-slices) is implemented
  // but OFF by default: measured on B200 it does not shorten the tile period (L2 already de-duplicates the <= 4
  // concurrent readers of a tile) and co-residency of clusters costs CTAs.  LASR_GEMM_MULTICAST=1 enables it.
  static const bool use_mc = getenv("LASR_GEMM_MULTICAST") != nullptr;
  const int cluster = (use_mc && (nblk == 2 || nblk == 4 || nblk == 8)) ? nblk : 1;
  const bool series = fu.sr_n > 0;
  if (series && (!b_mn || cluster > 1 || (N % WS_BN) || (fu.sr_S % 128))) return LASR_ERR_UNSUPPORTED;
  CUtensorMap ta, tb, tc;
  int rc = series ? series_maps(&ta, &tc, a, out, fu, N, K, lda)
                  : make_tmap_2d_bf16(&ta, a, K, M, static_cast<uint64_t>(lda) * 2, 64, 128 / cluster, true);
  if (rc) return rc;
  if (!b_mn)
    rc = make_tmap_2d_bf16(&tb, b, K, N, static_cast<uint64_t>(ldb) * 2, 64, WS_BN, true);
  else
    rc = make_tmap_2d_bf16(&tb, b, N, K, static_cast<uint64_t>(ldb) * 2, 64, 64, true);
  if (rc) return rc;
  if (!series) rc = make_tmap_2d_bf16(&tc, out, N, M, static_cast<uint64_t>(ldc) * 2, 64, 128, true);
  if (rc) return rc;
  const bool grouped = fu.a2 != nullptr;
  if (grouped && (cluster > 1 || (reinterpret_cast<uintptr_t>(fu.out2) & 15))) return LASR_ERR_UNSUPPORTED;
  CUtensorMap ta2 = ta, tb2 = tb, tc2 = tc;
  if (grouped) {
    rc = series ? series_maps(&ta2, &tc2, fu.a2, fu.out2, fu, N, K, lda)
                : make_tmap_2d_bf16(&ta2, fu.a2, K, M, static_cast<uint64_t>(lda) * 2, 64, 128, true);
    if (rc) return rc;
    if (!b_mn)
      rc = make_tmap_2d_bf16(&tb2, fu.b2, K, N, static_cast<uint64_t>(ldb) * 2, 64, WS_BN, true);
    else
      rc = make_tmap_2d_bf16(&tb2, fu.b2, N, K, static_cast<uint64_t>(ldb) * 2, 64, 64, true);
    if (rc) return rc;
    if (!series) rc = make_tmap_2d_bf16(&tc2, fu.out2, N, M, static_cast<uint64_t>(ldc) * 2, 64, 128, true);
    if (rc) return rc;
  }
  const int sms = grouped ? sm_budget() / 2 : sm_budget();  // CTAs available to ONE problem
  GemmWsParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = cdiv(M, WS_BM);
  if (series) {
    p.tr_bpu = fu.sr_S / 128;
    p.tr_off = fu.sr_off;
    p.tr_S = fu.sr_S;
    p.tr_blocks = fu.sr_n * p.tr_bpu;
    p.num_m_blocks = p.tr_blocks;
  }
  p.num_n_blocks = cdiv(N, WS_BN);
  p.num_k_blocks = kbs;
  if (p.num_n_blocks > sms) return LASR_ERR_UNSUPPORTED;
  int per_n = sms / p.num_n_blocks;
  if (per_n > p.num_m_blocks) per_n = p.num_m_blocks;
  // equalise: the same number of rounds with as few CTAs as needed keeps every CTA's tile count within one
  p.cluster = cluster;
  const int budget = 232448 - 1024 - 512 - kbs * WS_WKB_BYTES - WS_STG_BYTES;
  int stages = budget / WS_A_BYTES;
  if (stages > WS_MAX_STAGES) stages = WS_MAX_STAGES;
  if (stages < 2) return LASR_ERR_UNSUPPORTED;
  p.stages = stages;
  p.bias = bias;
  p.lengths = lengths;
  p.T = T;
  p.stats = stats;
  p.residual = static_cast<const __nv_bfloat16*>(fu.residual);
  p.ld_res = fu.ld_res;
  p.relu = fu.relu;
  p.bias_after_mask = fu.bias_after_mask;
  p.trace = g_ws_trace;
  p.w_early = early_param_loads() ? 1 : 0;
  const int smem = 1024 + 512 + kbs * WS_WKB_BYTES + stages * WS_A_BYTES + WS_STG_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_ws_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(1) ? 2 : 1;
  if (cluster > 1) {
    // clusters must be co-resident (persistent kernel): cap the grid at what the GPCs can hold at once
    static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (max_clusters[cluster] == 0) {
      cfg.gridDim = dim3(cluster * 16);
      cfg.dynamicSmemBytes = 232448;
      int n = 0;
      cudaError_t e = b_mn ? cudaOccupancyMaxActiveClusters(&n, gemm_ws_kernel<true>, &cfg)
                           : cudaOccupancyMaxActiveClusters(&n, gemm_ws_kernel<false>, &cfg);
      if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        n = (sm_budget() / cluster) * 3 / 4;
      }
      max_clusters[cluster] = n;
      cfg.dynamicSmemBytes = smem;
    }
    if (per_n > max_clusters[cluster]) per_n = max_clusters[cluster];
  }
  const int rounds = cdiv(p.num_m_blocks, per_n);
  per_n = cdiv(p.num_m_blocks, rounds);
  p.ctas_per_n = per_n;
  p.groups = grouped ? 2 : 1;
  p.group_ctas = p.num_n_blocks * per_n;
  p.lengths2 = fu.lengths2;
  p.stats2 = fu.stats2;
  cfg.gridDim = dim3(p.group_ctas * p.groups);
  cudaError_t le = series ? cudaLaunchKernelEx(&cfg, gemm_ws_kernel<true, true>, ta, tb, tc, ta2, tb2, tc2, p)
                   : b_mn ? cudaLaunchKernelEx(&cfg, gemm_ws_kernel<true>, ta, tb, tc, ta2, tb2, tc2, p)
                          : cudaLaunchKernelEx(&cfg, gemm_ws_kernel<false>, ta, tb, tc, ta2, tb2, tc2, p);
  if (le != cudaSuccess) {
    lasr_set_cuda_error(le);
    return LASR_ERR_CUDA;
  }
  return LASR_OK;
}

// CTA-pair launch (gemm_ws2_kernel); LASR_ERR_UNSUPPORTED when the shape does not qualify
static int launch_ws2(bool b_mn, const void* a, const void* b, void* out, const float* bias, const int32_t* lengths,
                      int T, double* stats, int M, int N, int K, int lda, int ldb, int ldc, cudaStream_t stream,
                      const WsFused& fu = WsFused()) {
  // Measured (tools/bench_kernels.py, M = 25 632): the pair wins where the activation tile is large, K >= 512
  // (512->512 fwd+stats 29.5 -> 25.4 us, 512->1024 48 -> 33.7 us, dgrad of 256->512 21.2 -> 19.3 us) and loses at K = 256
  // (256->256 16.4 -> 21.5 us: 101 pair tiles over 74 pairs is two full rounds).  LASR_GEMM_PAIR=0 disables it,
  // LASR_GEMM_PAIR=2 forces it for every N >= 256.
  static const int mode = getenv("LASR_GEMM_PAIR") != nullptr ? atoi(getenv("LASR_GEMM_PAIR")) : 1;
  if (mode == 0) return LASR_ERR_UNSUPPORTED;
  const int kbs = cdiv(K, WS_BK);
  if (mode == 1 && K < 512) return LASR_ERR_UNSUPPORTED;
  if (N < 256 || kbs * WS_WKB_BYTES > 131072) return LASR_ERR_UNSUPPORTED;
  if ((ldc % 8) || (reinterpret_cast<uintptr_t>(out) & 15)) return LASR_ERR_UNSUPPORTED;
  const bool series = fu.sr_n > 0;
  if (series && (!b_mn || (N % 256) || (fu.sr_S % 128))) return LASR_ERR_UNSUPPORTED;
  CUtensorMap ta, tb, tc;
  int rc = series ? series_maps(&ta, &tc, a, out, fu, N, K, lda)
                  : make_tmap_2d_bf16(&ta, a, K, M, static_cast<uint64_t>(lda) * 2, 64, 128, true);
  if (rc) return rc;
  if (!b_mn)
    rc = make_tmap_2d_bf16(&tb, b, K, N, static_cast<uint64_t>(ldb) * 2, 64, WS_BN, true);
  else
    rc = make_tmap_2d_bf16(&tb, b, N, K, static_cast<uint64_t>(ldb) * 2, 64, 64, true);
  if (rc) return rc;
  if (!series) rc = make_tmap_2d_bf16(&tc, out, N, M, static_cast<uint64_t>(ldc) * 2, 64, 128, true);
  if (rc) return rc;
  const bool grouped = fu.a2 != nullptr;
  if (grouped && (reinterpret_cast<uintptr_t>(fu.out2) & 15)) return LASR_ERR_UNSUPPORTED;
  CUtensorMap ta2 = ta, tb2 = tb, tc2 = tc;
  if (grouped) {
    rc = series ? series_maps(&ta2, &tc2, fu.a2, fu.out2, fu, N, K, lda)
                : make_tmap_2d_bf16(&ta2, fu.a2, K, M, static_cast<uint64_t>(lda) * 2, 64, 128, true);
    if (rc) return rc;
    if (!b_mn)
      rc = make_tmap_2d_bf16(&tb2, fu.b2, K, N, static_cast<uint64_t>(ldb) * 2, 64, WS_BN, true);
    else
      rc = make_tmap_2d_bf16(&tb2, fu.b2, N, K, static_cast<uint64_t>(ldb) * 2, 64, 64, true);
    if (rc) return rc;
    if (!series) rc = make_tmap_2d_bf16(&tc2, fu.out2, N, M, static_cast<uint64_t>(ldc) * 2, 64, 128, true);
    if (rc) return rc;
  }
  GemmWsParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = cdiv(M, 256);   // 256-row pair tiles
  if (series) {
    p.tr_bpu = fu.sr_S / 128;
    p.tr_off = fu.sr_off;
    p.tr_S = fu.sr_S;
    p.tr_blocks = fu.sr_n * p.tr_bpu;
    p.num_m_blocks = cdiv(p.tr_blocks, 2);
  }
  p.num_n_blocks = cdiv(N, 256);   // 256-column pair slices
  p.num_k_blocks = kbs;
  const int pairs = (sm_budget() / 2) / (grouped ? 2 : 1);  // pairs available to ONE problem
  if (p.num_n_blocks > pairs) return LASR_ERR_UNSUPPORTED;
  int per_n = pairs / p.num_n_blocks;
  if (per_n > p.num_m_blocks) per_n = p.num_m_blocks;
  const int rounds = cdiv(p.num_m_blocks, per_n);
  per_n = cdiv(p.num_m_blocks, rounds);
  p.ctas_per_n = per_n;
  p.cluster = 2;
  const int budget = 232448 - 1024 - 512 - kbs * WS_WKB_BYTES - WS_STG_BYTES;
  int stages = budget / WS_A_BYTES;
  if (stages > WS_MAX_STAGES) stages = WS_MAX_STAGES;
  if (stages < 2) return LASR_ERR_UNSUPPORTED;
  p.stages = stages;
  p.bias = bias;
  p.lengths = lengths;
  p.T = T;
  p.stats = stats;
  p.residual = static_cast<const __nv_bfloat16*>(fu.residual);
  p.ld_res = fu.ld_res;
  p.relu = fu.relu;
  p.bias_after_mask = fu.bias_after_mask;
  p.trace = nullptr;
  p.w_early = early_param_loads() ? 1 : 0;
  const int smem = 1024 + 512 + kbs * WS_WKB_BYTES + stages * WS_A_BYTES + WS_STG_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ws2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_ws2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_ws2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  p.groups = grouped ? 2 : 1;
  p.group_ctas = 2 * p.num_n_blocks * per_n;
  p.lengths2 = fu.lengths2;
  p.stats2 = fu.stats2;
  cfg.gridDim = dim3(p.group_ctas * p.groups);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(1) ? 2 : 1;
  cudaError_t le = series ? cudaLaunchKernelEx(&cfg, gemm_ws2_kernel<true, true>, ta, tb, tc, ta2, tb2, tc2, p)
                   : b_mn ? cudaLaunchKernelEx(&cfg, gemm_ws2_kernel<true>, ta, tb, tc, ta2, tb2, tc2, p)
                          : cudaLaunchKernelEx(&cfg, gemm_ws2_kernel<false>, ta, tb, tc, ta2, tb2, tc2, p);
  if (le != cudaSuccess) {
    lasr_set_cuda_error(le);
    return LASR_ERR_CUDA;
  }
  return LASR_OK;
}

static int pick_bn(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256)); }

// y[M, N] = x[M, K] w[N, K]^T, bf16 in, bf16/fp32 out
int gemm_tc_nt(const void* a, const void* b, void* out, const float* bias, const int32_t* lengths, int T, double* stats,
               int M, int N, int K, int lda, int ldb, int ldc, int out_f32, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return LASR_ERR_BAD_SHAPE;
  if ((lda % 8) || (ldb % 8)) return LASR_ERR_ALIGNMENT;
  if (!out_f32 && N > 64) {
    const int rc_ws2 = launch_ws2(false, a, b, out, bias, lengths, T, stats, M, N, K, lda, ldb, ldc, stream);
    if (rc_ws2 != LASR_ERR_UNSUPPORTED) return rc_ws2;
    const int rc_ws = launch_ws(false, a, b, out, bias, lengths, T, stats, M, N, K, lda, ldb, ldc, stream);
    if (rc_ws != LASR_ERR_UNSUPPORTED) return rc_ws;
  }
  const int BN = pick_bn(N);
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, a, K, M, static_cast<uint64_t>(lda) * 2, 64, 128, true);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, b, K, N, static_cast<uint64_t>(ldb) * 2, 64, BN, true);
  if (rc) return rc;
  GemmTcParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = cdiv(M, 128);
  p.num_n_blocks = cdiv(N, BN);
  p.num_k_blocks = cdiv(K, 64);
  p.k_splits = 1;
  p.kb_per_split = p.num_k_blocks;
  p.out = out;
  p.ldc = ldc;
  p.out_f32 = out_f32;
  const int esz = out_f32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((static_cast<size_t>(ldc) * esz) % 16 == 0);
  p.bias = bias;
  p.lengths = lengths;
  p.T = T;
  p.stats = stats;
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < sm_budget() ? tiles : sm_budget();
  switch (BN) {
    case 32: return launch_inst<32, false, false, 0>(ta, tb, p, grid, stream);
    case 64: return launch_inst<64, false, false, 0>(ta, tb, p, grid, stream);
    case 128: return launch_inst<128, false, false, 0>(ta, tb, p, grid, stream);
    default: return launch_inst<256, false, false, 0>(ta, tb, p, grid, stream);
  }
}

// Two forward (b_mn = false: y = x w^T with mask / statistics) or data-gradient (b_mn = true: dx = dy w) GEMMs of
// identical shape in ONE launch: the CTAs are split between the problems, each CTA walks twice as many tiles, and the
// launch, prologue and weight-slice load are paid once.  LASR_ERR_UNSUPPORTED: caller issues two launches.
int gemm_tc_grouped2(bool b_mn, const void* a1, const void* b1, void* out1, const int32_t* lengths1, double* stats1,
                     const void* a2, const void* b2, void* out2, const int32_t* lengths2, double* stats2, int T, int M,
                     int N, int K, int lda, int ldb, int ldc, cudaStream_t stream) {
  static const bool off = getenv("LASR_GEMM_GROUPED") != nullptr && atoi(getenv("LASR_GEMM_GROUPED")) == 0;
  if (off || M <= 0 || N <= 64 || K <= 0) return LASR_ERR_UNSUPPORTED;
  if ((lda % 8) || (ldb % 8)) return LASR_ERR_UNSUPPORTED;
  WsFused fu;
  fu.a2 = a2;
  fu.b2 = b2;
  fu.out2 = out2;
  fu.lengths2 = lengths2;
  fu.stats2 = stats2;
  const int rc2 = launch_ws2(b_mn, a1, b1, out1, nullptr, lengths1, T, stats1, M, N, K, lda, ldb, ldc, stream, fu);
  if (rc2 != LASR_ERR_UNSUPPORTED) return rc2;
  return launch_ws(b_mn, a1, b1, out1, nullptr, lengths1, T, stats1, M, N, K, lda, ldb, ldc, stream, fu);
}

// inference epilogue: y = act(mask(x w^T) + bias + residual), bf16; weight-stationary kernels only
int gemm_tc_nt_fused(const void* a, const void* b, void* out, const float* bias, const void* residual, int ld_res,
                     const int32_t* lengths, int T, int relu, int M, int N, int K, int lda, int ldb, int ldc,
                     cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return LASR_ERR_BAD_SHAPE;
  if ((lda % 8) || (ldb % 8) || (ld_res % 8) || (N % 32)) return LASR_ERR_ALIGNMENT;
  if (residual != nullptr && (reinterpret_cast<uintptr_t>(residual) & 15)) return LASR_ERR_ALIGNMENT;
  WsFused fu;
  fu.residual = residual;
  fu.ld_res = ld_res;
  fu.relu = relu;
  fu.bias_after_mask = 1;
  const int rc2 = launch_ws2(false, a, b, out, bias, lengths, T, nullptr, M, N, K, lda, ldb, ldc, stream, fu);
  if (rc2 != LASR_ERR_UNSUPPORTED) return rc2;
  return launch_ws(false, a, b, out, bias, lengths, T, nullptr, M, N, K, lda, ldb, ldc, stream, fu);
}

// y[M, N] = a[M, K] b[K, N]: the data gradient dx = dy * W with W in its native [Cout = K, Cin = N] layout, i.e. an
// MN-major B operand (no transposed weight copy).  bf16 in, bf16/fp32 out.
int gemm_tc_nn(const void* a, const void* b, void* out, int M, int N, int K, int lda, int ldb, int ldc, int out_f32,
               cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return LASR_ERR_BAD_SHAPE;
  if ((lda % 8) || (ldb % 8)) return LASR_ERR_ALIGNMENT;
  if (!out_f32 && N > 64) {
    const int rc_ws2 = launch_ws2(true, a, b, out, nullptr, nullptr, 0, nullptr, M, N, K, lda, ldb, ldc, stream);
    if (rc_ws2 != LASR_ERR_UNSUPPORTED) return rc_ws2;
    const int rc_ws = launch_ws(true, a, b, out, nullptr, nullptr, 0, nullptr, M, N, K, lda, ldb, ldc, stream);
    if (rc_ws != LASR_ERR_UNSUPPORTED) return rc_ws;
  }
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, a, K, M, static_cast<uint64_t>(lda) * 2, 64, 128, true);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, b, N, K, static_cast<uint64_t>(ldb) * 2, 64, 64, true);
  if (rc) return rc;
  GemmTcParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = cdiv(M, 128);
  p.num_n_blocks = cdiv(N, BN);
  p.num_k_blocks = cdiv(K, 64);
  p.k_splits = 1;
  p.kb_per_split = p.num_k_blocks;
  p.out = out;
  p.ldc = ldc;
  p.out_f32 = out_f32;
  const int esz = out_f32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((static_cast<size_t>(ldc) * esz) % 16 == 0);
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < sm_budget() ? tiles : sm_budget();
  switch (BN) {
    case 64: return launch_inst<64, false, true, 0>(ta, tb, p, grid, stream);
    case 128: return launch_inst<128, false, true, 0>(ta, tb, p, grid, stream);
    default: return launch_inst<256, false, true, 0>(ta, tb, p, grid, stream);
  }
}

// Data gradient(s) written as channel-major series: outT[c][n][off + t] = sum_k dy[n, t, k] w[k, c] (and the same for a
// second problem of identical shape: the block's residual conv), the operand format of the TMA-fed depthwise kernels.
// Weight-stationary kernels with swapped operands; LASR_ERR_UNSUPPORTED when the shape does not qualify.
int gemm_tc_nn_series(const void* dy1, const void* w1, void* outT1, const void* dy2, const void* w2, void* outT2,
                      int n_utt, int T, int N, int K, int lda, int ldb, int S, int off, cudaStream_t stream) {
  if (n_utt <= 0 || T <= 0 || N <= 64 || K <= 0 || (lda % 8) || (ldb % 8)) return LASR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(outT1) & 255) || (dy2 != nullptr && (reinterpret_cast<uintptr_t>(outT2) & 255)))
    return LASR_ERR_UNSUPPORTED;
  WsFused fu;
  fu.sr_n = n_utt;
  fu.sr_T = T;
  fu.sr_S = S;
  fu.sr_off = off;
  fu.a2 = dy2;
  fu.b2 = w2;
  fu.out2 = outT2;
  const int M = n_utt * T;
  const int rc2 = launch_ws2(true, dy1, w1, outT1, nullptr, nullptr, 0, nullptr, M, N, K, lda, ldb, 8, stream, fu);
  if (rc2 != LASR_ERR_UNSUPPORTED) return rc2;
  return launch_ws(true, dy1, w1, outT1, nullptr, nullptr, 0, nullptr, M, N, K, lda, ldb, 8, stream, fu);
}

// Two weight gradients of identical shape (the block's pointwise conv and its residual conv) in ONE launch: every CTA
// still owns one output tile, with a twice longer slice of the frame axis -- launch, prologue and the split-K reduction
// epilogue are paid once instead of twice.  LASR_ERR_UNSUPPORTED: caller issues two single launches.
int gemm_tc_tn_accum2(const void* dy1, const void* x1, float* dw1, const void* dy2, const void* x2, float* dw2, int R,
                      int Cout, int Cin, int lddy, int ldx, int lddw, cudaStream_t stream) {
  static const bool off = getenv("LASR_WGRAD_GROUPED") != nullptr && atoi(getenv("LASR_WGRAD_GROUPED")) == 0;
  if (off) return LASR_ERR_UNSUPPORTED;
  if (R <= 0 || Cout <= 0 || Cin < 32) return LASR_ERR_UNSUPPORTED;
  if ((lddy % 8) || (ldx % 8)) return LASR_ERR_UNSUPPORTED;
  const bool aligned = ((reinterpret_cast<uintptr_t>(dw1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dw2) & 15) == 0) &&
                       ((static_cast<size_t>(lddw) * 4) % 16 == 0);
  if (!aligned) return LASR_ERR_UNSUPPORTED;
  const int BN = Cin <= 64 ? 64 : (Cin <= 128 ? 128 : 256);
  CUtensorMap ta, tb, tc, ta2, tb2, tc2;
  int rc = make_tmap_2d_bf16(&ta, dy1, Cout, R, static_cast<uint64_t>(lddy) * 2, 64, 64, true);
  if (!rc) rc = make_tmap_2d_bf16(&tb, x1, Cin, R, static_cast<uint64_t>(ldx) * 2, 64, 64, true);
  if (!rc) rc = make_tmap_2d_f32(&tc, dw1, Cin, Cout, static_cast<uint64_t>(lddw) * 4, 32, 32, true);
  if (!rc) rc = make_tmap_2d_bf16(&ta2, dy2, Cout, R, static_cast<uint64_t>(lddy) * 2, 64, 64, true);
  if (!rc) rc = make_tmap_2d_bf16(&tb2, x2, Cin, R, static_cast<uint64_t>(ldx) * 2, 64, 64, true);
  if (!rc) rc = make_tmap_2d_f32(&tc2, dw2, Cin, Cout, static_cast<uint64_t>(lddw) * 4, 32, 32, true);
  if (rc) return rc;
  GemmTcParams p{};
  p.M = Cout;
  p.N = Cin;
  p.K = R;
  p.num_m_blocks = cdiv(Cout, 128);
  p.num_n_blocks = cdiv(Cin, BN);
  p.num_k_blocks = cdiv(R, 64);
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  int splits = sm_budget() / (2 * tiles);
  if (splits < 1) return LASR_ERR_UNSUPPORTED;
  if (splits > p.num_k_blocks) splits = p.num_k_blocks;
  p.kb_per_split = cdiv(p.num_k_blocks, splits);
  p.k_splits = cdiv(p.num_k_blocks, p.kb_per_split);
  p.out = dw1;
  p.ldc = lddw;
  p.out_f32 = 1;
  p.vec_ok = 1;
  p.groups = 2;
  const int total = 2 * tiles * p.k_splits;
  const int grid = total < sm_budget() ? total : sm_budget();
  switch (BN) {
    case 64: return launch_inst<64, true, true, 1>(ta, tb, p, grid, stream, &tc, &ta2, &tb2, &tc2);
    case 128: return launch_inst<128, true, true, 1>(ta, tb, p, grid, stream, &tc, &ta2, &tb2, &tc2);
    default: return launch_inst<256, true, true, 1>(ta, tb, p, grid, stream, &tc, &ta2, &tb2, &tc2);
  }
}

// dw[M=Cout, N=Cin] += dy[R, Cout]^T x[R, Cin]   (R = frames), fp32 RED accumulate
int gemm_tc_tn_accum(const void* dy, const void* x, float* dw, int R, int Cout, int Cin, int lddy, int ldx, int lddw,
                     cudaStream_t stream) {
  if (R <= 0 || Cout <= 0 || Cin <= 0) return LASR_ERR_BAD_SHAPE;
  if ((lddy % 8) || (ldx % 8)) return LASR_ERR_ALIGNMENT;
  {
    // CTA-pair path: 256 x 256 pair tiles, needs the TMA reduce epilogue (aligned fp32 gradient).  OPT-IN: measured
    // equal to the single-CTA kernel (512x512: 23.5 vs 23.3 us, 1024x512: 32.8 vs 33.6 us) -- the weight gradient is
    // paced by its split-K reduction traffic, not by operand loads.  LASR_WGRAD_PAIR=1: Cout + Cin >= 1024, =2: all.
    static const int pair_mode = getenv("LASR_WGRAD_PAIR") != nullptr ? atoi(getenv("LASR_WGRAD_PAIR")) : 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(dw) & 15) == 0) && ((static_cast<size_t>(lddw) * 4) % 16 == 0);
    if (pair_mode != 0 && aligned && Cout >= 256 && Cin >= 256 && (pair_mode == 2 || Cout + Cin >= 1024)) {
      CUtensorMap ta, tb, tc;
      int rc = make_tmap_2d_bf16(&ta, dy, Cout, R, static_cast<uint64_t>(lddy) * 2, 64, 64, true);
      if (rc) return rc;
      rc = make_tmap_2d_bf16(&tb, x, Cin, R, static_cast<uint64_t>(ldx) * 2, 64, 64, true);
      if (rc) return rc;
      rc = make_tmap_2d_f32(&tc, dw, Cin, Cout, static_cast<uint64_t>(lddw) * 4, 32, 32, true);
      if (rc) return rc;
      GemmTcParams p{};
      p.M = Cout;
      p.N = Cin;
      p.K = R;
      p.num_m_blocks = cdiv(Cout, 256);
      p.num_n_blocks = cdiv(Cin, 256);
      p.num_k_blocks = cdiv(R, 64);
      const int tiles = p.num_m_blocks * p.num_n_blocks;
      const int pairs = sm_budget() / 2;
      int splits = pairs / tiles;
      if (splits < 1) splits = 1;
      if (splits > p.num_k_blocks) splits = p.num_k_blocks;
      p.kb_per_split = cdiv(p.num_k_blocks, splits);
      p.k_splits = cdiv(p.num_k_blocks, p.kb_per_split);
      p.out = dw;
      p.ldc = lddw;
      p.out_f32 = 1;
      p.vec_ok = 1;
      const int total = tiles * p.k_splits;
      const int npairs = total < pairs ? total : pairs;
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM);
        if (e != cudaSuccess) {
          lasr_set_cuda_error(e);
          return LASR_ERR_CUDA;
        }
        configured = true;
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * npairs);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = WG2_SMEM;
      cfg.stream = stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_enabled(1) ? 2 : 1;
      LASR_CHECK_PDL(cudaLaunchKernelEx(&cfg, gemm_wgrad2_kernel, ta, tb, tc, p));
      return LASR_OK;
    }
  }
  const int BN = Cin <= 64 ? 64 : (Cin <= 128 ? 128 : 256);
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, dy, Cout, R, static_cast<uint64_t>(lddy) * 2, 64, 64, true);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, x, Cin, R, static_cast<uint64_t>(ldx) * 2, 64, 64, true);
  if (rc) return rc;
  GemmTcParams p{};
  p.M = Cout;
  p.N = Cin;
  p.K = R;
  p.num_m_blocks = cdiv(Cout, 128);
  p.num_n_blocks = cdiv(Cin, BN);
  p.num_k_blocks = cdiv(R, 64);
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  // one tile per CTA: every extra split adds a full Cout x Cin pass of L2 reductions
  static const int split_mult = getenv("LASR_WGRAD_SPLIT_MULT") ? atoi(getenv("LASR_WGRAD_SPLIT_MULT")) : 1;
  int splits = (split_mult * sm_budget()) / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.num_k_blocks) splits = p.num_k_blocks;
  p.kb_per_split = cdiv(p.num_k_blocks, splits);
  p.k_splits = cdiv(p.num_k_blocks, p.kb_per_split);
  p.out = dw;
  p.ldc = lddw;
  p.out_f32 = 1;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(dw) & 15) == 0) && ((static_cast<size_t>(lddw) * 4) % 16 == 0) && Cin >= 32;
  static const bool no_tma_red = getenv("LASR_WGRAD_NO_TMA_REDUCE") != nullptr;
  if (no_tma_red) p.vec_ok = 0;
  CUtensorMap tc = ta;
  if (p.vec_ok) {
    rc = make_tmap_2d_f32(&tc, dw, Cin, Cout, static_cast<uint64_t>(lddw) * 4, 32, 32, true);
    if (rc) return rc;
  }
  const int total = tiles * p.k_splits;
  const int grid = total < sm_budget() ? total : sm_budget();
  switch (BN) {
    case 64: return launch_inst<64, true, true, 1>(ta, tb, p, grid, stream, &tc);
    case 128: return launch_inst<128, true, true, 1>(ta, tb, p, grid, stream, &tc);
    default: return launch_inst<256, true, true, 1>(ta, tb, p, grid, stream, &tc);
  }
}

}  // namespace lasr
