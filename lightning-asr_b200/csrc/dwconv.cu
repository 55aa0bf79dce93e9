// Depthwise Conv1d over channels-last activations (replaces nn.Conv1d(C, C, k, groups=C) at
// models/QuartNet.py:14-21,30 and its autograd backward).
//
// With k = 33..87 taps this op is NOT memory-bound on B200: k FMAs per element against 4 bytes of traffic puts it
// on the fp32 pipe, so the kernels are built around Blackwell's packed FFMA2 (fma.rn.f32x2): one thread owns a
// PAIR of adjacent channels (one 32-bit bf16x2 word in shared memory, one 64-bit accumulator), a warp owns 32
// pairs = 64 channels, so every shared-memory access is a conflict-free 128-byte row and every global store is a
// full 128-byte line.  The time tile plus its (k-1)-frame halo is staged in shared memory by TMA (3-D map over
// [N, T, C]; out-of-range frames -- the conv's zero padding and the utterance boundary -- are zero-filled by the
// TMA unit, so there is no bounds logic in the inner loop).
//
//   fwd   : thread = (channel pair, strip of R output frames); taps are consumed in chunks of 8 with a
//           sliding register window, fully unrolled at compile time (k is a template parameter).
//   wgrad : warp = contiguous chunk of taps, lane = channel pair; accumulators stay in registers across all the
//           time tiles a CTA walks (double-buffered TMA), then one fp32 RED per (tap, channel).
// Weights and weight gradients keep the reference's layout: fp32 [C, 1, K] (nn.Conv1d(groups=C).weight).
#include "common.cuh"

namespace lasr {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn3() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 3-D map over a channels-last activation [N, T, C]; box = {64 channels, box_rows frames, 1 utterance}
static int make_tmap_ntc(CUtensorMap* out, const void* base, int N, int T, int C, int dtype, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn3();
  if (fn == nullptr) return LASR_ERR_DRIVER;
  const int es = dtype == LASR_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((static_cast<size_t>(C) * es) & 15) != 0)
    return LASR_ERR_ALIGNMENT;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * es, static_cast<cuuint64_t>(C) * es * T};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype == LASR_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LASR_OK : LASR_ERR_DRIVER;
}

template <typename T>
struct PairIO;
template <>
struct PairIO<__nv_bfloat16> {
  using word = uint32_t;  // two bf16
  static __device__ __forceinline__ float2 load(const void* row, int pair) {
    return bf16x2_to_f32x2(reinterpret_cast<const uint32_t*>(row)[pair]);
  }
  static __device__ __forceinline__ void store(void* row, int pair, float2 v) {
    reinterpret_cast<uint32_t*>(row)[pair] = f32x2_to_bf16x2(v.x, v.y);
  }
};
template <>
struct PairIO<float> {
  using word = float2;
  static __device__ __forceinline__ float2 load(const void* row, int pair) {
    return reinterpret_cast<const float2*>(row)[pair];
  }
  static __device__ __forceinline__ void store(void* row, int pair, float2 v) {
    reinterpret_cast<float2*>(row)[pair] = v;
  }
};

constexpr int DW_CC = 64;       // channels per CTA
constexpr int DW_TT = 128;      // output frames per CTA tile
constexpr int DW_J = 8;         // taps per register chunk
constexpr int DW_TMA_ROWS = 64; // frames per TMA box

template <int K, int S>
struct DwFwdCfg {
  static constexpr int R = 16 / S;                        // outputs per thread
  static constexpr int STRIPS = DW_TT / R;                // strips per tile (== warps)
  static constexpr int THREADS = STRIPS * 32;
  static constexpr int ROWS_IN = (DW_TT - 1) * S + K;     // input frames incl. halo
  static constexpr int BOXES = (ROWS_IN + DW_TMA_ROWS - 1) / DW_TMA_ROWS;
  static constexpr int ROWS_ALLOC = BOXES * DW_TMA_ROWS;
};

template <typename T, int K, int S>
__global__ void __launch_bounds__(DwFwdCfg<K, S>::THREADS)
dwconv_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ wt, T* __restrict__ y,
                  const T* __restrict__ addend, int T_out, int C, int flip) {
  using Cfg = DwFwdCfg<K, S>;
  constexpr int R = Cfg::R;
  constexpr int ROW_BYTES = DW_CC * sizeof(T);
  extern __shared__ __align__(128) uint8_t dsmem[];
  uint8_t* sx = dsmem;                                                        // [ROWS_ALLOC][64] of T
  float2* sw = reinterpret_cast<float2*>(dsmem + Cfg::ROWS_ALLOC * ROW_BYTES);  // [K][32] float2
  __shared__ uint64_t bar;

  const int c0 = blockIdx.x * DW_CC;
  const int t0 = blockIdx.y * DW_TT;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int pair = tid & 31;
  const int strip = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar, Cfg::BOXES * DW_TMA_ROWS * ROW_BYTES);
    const int tin0 = t0 * S - K / 2;
#pragma unroll
    for (int b = 0; b < Cfg::BOXES; ++b)
      tma_load_3d(sx + b * DW_TMA_ROWS * ROW_BYTES, &tmap_x, &bar, c0, tin0 + b * DW_TMA_ROWS, n);
  }
  // taps -> smem as [K][32] channel pairs from the reference's native layout weight[c, 0, j] (global reads run
  // along j, i.e. coalesced); flip reverses the tap order (data-gradient)
  for (int i = tid; i < K * 32; i += Cfg::THREADS) {
    const int p = i / K, j = i - p * K;
    const int c = c0 + 2 * p;
    const int jj = flip ? (K - 1 - j) : j;
    float2 v = make_float2(0.f, 0.f);
    if (c < C) v = make_float2(wt[static_cast<size_t>(c) * K + jj], wt[static_cast<size_t>(c + 1) * K + jj]);
    sw[j * 32 + p] = v;
  }
  __syncthreads();
  mbar_wait(&bar, 0);

  float2 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);

  const uint8_t* xrow0 = sx + static_cast<size_t>(strip * R * S) * ROW_BYTES;
  constexpr int WIN = (R - 1) * S + DW_J;  // register window of input frames
  float2 xw[WIN];
#pragma unroll
  for (int i = 0; i < WIN; ++i) xw[i] = PairIO<T>::load(xrow0 + i * ROW_BYTES, pair);

#pragma unroll
  for (int jc = 0; jc < K; jc += DW_J) {
    float2 wj[DW_J];
#pragma unroll
    for (int j = 0; j < DW_J; ++j) wj[j] = (jc + j < K) ? sw[(jc + j) * 32 + pair] : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int j = 0; j < DW_J; ++j) {
        if (jc + j < K) acc[r] = ffma2(wj[j], xw[r * S + j], acc[r]);
      }
    }
    if (jc + DW_J < K) {
#pragma unroll
      for (int i = 0; i < WIN - DW_J; ++i) xw[i] = xw[i + DW_J];
#pragma unroll
      for (int i = WIN - DW_J; i < WIN; ++i) {
        // frames past the last one any remaining tap touches are never used
        if (jc + DW_J + i < (R - 1) * S + K) xw[i] = PairIO<T>::load(xrow0 + (jc + DW_J + i) * ROW_BYTES, pair);
      }
    }
  }

  const int c = c0 + 2 * pair;
  if (c < C) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int t = t0 + strip * R + r;
      if (t < T_out) {
        const size_t off = (static_cast<size_t>(n) * T_out + t) * C + c0;
        float2 v = acc[r];
        if (addend != nullptr) {
          const float2 a = PairIO<T>::load(addend + off, pair);
          v.x += a.x;
          v.y += a.y;
        }
        PairIO<T>::store(y + off, pair, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int DWG_TT = 128;   // output frames per tile
constexpr int DWG_WARPS = 8;
constexpr int DWG_RT = 8;     // frames per register block

template <int K, int S>
struct DwGradCfg {
  static constexpr int JW = (K + DWG_WARPS - 1) / DWG_WARPS;  // taps per warp
  static constexpr int ROWS_X = (DWG_TT - 1) * S + K;
  static constexpr int BOXES_X = (ROWS_X + DW_TMA_ROWS - 1) / DW_TMA_ROWS;
  static constexpr int ROWS_X_ALLOC = BOXES_X * DW_TMA_ROWS;
  static constexpr int BOXES_DY = DWG_TT / DW_TMA_ROWS;
};

template <typename T, int K, int S>
__global__ void __launch_bounds__(DWG_WARPS * 32)
dwconv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                    float* __restrict__ dwt, int N, int T_out, int C, int tiles_per_utt, int tiles_per_cta) {
  using Cfg = DwGradCfg<K, S>;
  constexpr int JW = Cfg::JW;
  constexpr int ROW_BYTES = DW_CC * sizeof(T);
  constexpr int STAGE_BYTES = (Cfg::ROWS_X_ALLOC + DWG_TT) * ROW_BYTES;
  extern __shared__ __align__(128) uint8_t dsmem[];
  __shared__ uint64_t bar[2];

  const int c0 = blockIdx.x * DW_CC;
  const int tid = threadIdx.x;
  const int pair = tid & 31;
  const int warp = tid >> 5;
  const int j0 = warp * JW;
  const int total_tiles = N * tiles_per_utt;
  const int tile_begin = blockIdx.y * tiles_per_cta;
  const int tile_end = min(total_tiles, tile_begin + tiles_per_cta);

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int tile, int stage) {
    const int n = tile / tiles_per_utt;
    const int t0 = (tile - n * tiles_per_utt) * DWG_TT;
    uint8_t* sx = dsmem + stage * STAGE_BYTES;
    uint8_t* sdy = sx + Cfg::ROWS_X_ALLOC * ROW_BYTES;
    mbar_arrive_expect_tx(&bar[stage], (Cfg::BOXES_X + Cfg::BOXES_DY) * DW_TMA_ROWS * ROW_BYTES);
#pragma unroll
    for (int b = 0; b < Cfg::BOXES_X; ++b)
      tma_load_3d(sx + b * DW_TMA_ROWS * ROW_BYTES, &tmap_x, &bar[stage], c0, t0 * S - K / 2 + b * DW_TMA_ROWS, n);
#pragma unroll
    for (int b = 0; b < Cfg::BOXES_DY; ++b)
      tma_load_3d(sdy + b * DW_TMA_ROWS * ROW_BYTES, &tmap_dy, &bar[stage], c0, t0 + b * DW_TMA_ROWS, n);
  };

  float2 acc[JW];
#pragma unroll
  for (int j = 0; j < JW; ++j) acc[j] = make_float2(0.f, 0.f);

  if (tid == 0 && tile_begin < tile_end) issue(tile_begin, 0);

  for (int tile = tile_begin, it = 0; tile < tile_end; ++tile, ++it) {
    const int stage = it & 1;
    if (tid == 0 && tile + 1 < tile_end) issue(tile + 1, stage ^ 1);
    mbar_wait(&bar[stage], (it >> 1) & 1);
    const uint8_t* sx = dsmem + stage * STAGE_BYTES;
    const uint8_t* sdy = sx + Cfg::ROWS_X_ALLOC * ROW_BYTES;
    // dy rows past T_out inside the tile are zero-filled by TMA, so they contribute nothing
#pragma unroll 1
    for (int tb = 0; tb < DWG_TT; tb += DWG_RT) {
      float2 dyv[DWG_RT];
#pragma unroll
      for (int r = 0; r < DWG_RT; ++r) dyv[r] = PairIO<T>::load(sdy + (tb + r) * ROW_BYTES, pair);
      constexpr int WIN = (DWG_RT - 1) * S + JW;
      float2 xw[WIN];
#pragma unroll
      for (int i = 0; i < WIN; ++i) {
        const int row = tb * S + j0 + i;  // x frame (tile-relative, halo included) used by (r, jj): r*S + j0 + jj
        xw[i] = (row < Cfg::ROWS_X_ALLOC) ? PairIO<T>::load(sx + row * ROW_BYTES, pair) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < DWG_RT; ++r)
#pragma unroll
        for (int jj = 0; jj < JW; ++jj) acc[jj] = ffma2(dyv[r], xw[r * S + jj], acc[jj]);
    }
    __syncthreads();  // everyone done with this stage before it is refilled
  }

  const int c = c0 + 2 * pair;
  if (c < C) {
#pragma unroll
    for (int jj = 0; jj < JW; ++jj) {
      const int j = j0 + jj;
      if (j < K) {
        atomicAdd(dwt + static_cast<size_t>(c) * K + j, acc[jj].x);
        atomicAdd(dwt + static_cast<size_t>(c + 1) * K + j, acc[jj].y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
template <typename T, int K, int S>
static int launch_fwd(const void* x, const float* wt, void* y, const void* addend, int N, int T_in, int T_out, int C,
                      int flip, int dtype, cudaStream_t stream) {
  using Cfg = DwFwdCfg<K, S>;
  CUtensorMap tm;
  int rc = make_tmap_ntc(&tm, x, N, T_in, C, dtype, DW_TMA_ROWS);
  if (rc) return rc;
  const int smem = Cfg::ROWS_ALLOC * DW_CC * sizeof(T) + K * 32 * sizeof(float2);
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(dwconv_fwd_kernel<T, K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid(cdiv(C, DW_CC), cdiv(T_out, DW_TT), N);
  dwconv_fwd_kernel<T, K, S><<<grid, Cfg::THREADS, smem, stream>>>(tm, wt, static_cast<T*>(y),
                                                                   static_cast<const T*>(addend), T_out, C, flip);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

template <typename T, int K, int S>
static int launch_wgrad(const void* x, const void* dy, float* dwt, int N, int T_in, int T_out, int C, int dtype,
                        cudaStream_t stream) {
  using Cfg = DwGradCfg<K, S>;
  CUtensorMap tx, tdy;
  int rc = make_tmap_ntc(&tx, x, N, T_in, C, dtype, DW_TMA_ROWS);
  if (rc) return rc;
  rc = make_tmap_ntc(&tdy, dy, N, T_out, C, dtype, DW_TMA_ROWS);
  if (rc) return rc;
  const int smem = 2 * (Cfg::ROWS_X_ALLOC + DWG_TT) * DW_CC * sizeof(T);
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(dwconv_wgrad_kernel<T, K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  const int chunks = cdiv(C, DW_CC);
  const int tiles_per_utt = cdiv(T_out, DWG_TT);
  const int total_tiles = N * tiles_per_utt;
  int ctas_y = (2 * kNumSMs) / chunks;
  if (ctas_y < 1) ctas_y = 1;
  if (ctas_y > total_tiles) ctas_y = total_tiles;
  const int tiles_per_cta = cdiv(total_tiles, ctas_y);
  ctas_y = cdiv(total_tiles, tiles_per_cta);
  dim3 grid(chunks, ctas_y);
  dwconv_wgrad_kernel<T, K, S>
      <<<grid, DWG_WARPS * 32, smem, stream>>>(tx, tdy, dwt, N, T_out, C, tiles_per_utt, tiles_per_cta);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

#define LASR_DW_DISPATCH(FN, ...)                                                        \
  do {                                                                                   \
    if (stride == 1) {                                                                   \
      switch (K) {                                                                       \
        case 33: return FN<T, 33, 1>(__VA_ARGS__);                                       \
        case 39: return FN<T, 39, 1>(__VA_ARGS__);                                       \
        case 51: return FN<T, 51, 1>(__VA_ARGS__);                                       \
        case 63: return FN<T, 63, 1>(__VA_ARGS__);                                       \
        case 75: return FN<T, 75, 1>(__VA_ARGS__);                                       \
        case 87: return FN<T, 87, 1>(__VA_ARGS__);                                       \
        default: return LASR_ERR_UNSUPPORTED;                                            \
      }                                                                                  \
    } else if (stride == 2 && K == 33) {                                                 \
      return FN<T, 33, 2>(__VA_ARGS__);                                                  \
    }                                                                                    \
    return LASR_ERR_UNSUPPORTED;                                                         \
  } while (0)

template <typename T>
static int fwd_t(const void* x, const float* wt, void* y, const void* addend, int N, int T_in, int T_out, int C, int K,
                 int stride, int flip, int dtype, cudaStream_t stream) {
  LASR_DW_DISPATCH(launch_fwd, x, wt, y, addend, N, T_in, T_out, C, flip, dtype, stream);
}
template <typename T>
static int wgrad_t(const void* x, const void* dy, float* dwt, int N, int T_in, int T_out, int C, int K, int stride,
                   int dtype, cudaStream_t stream) {
  LASR_DW_DISPATCH(launch_wgrad, x, dy, dwt, N, T_in, T_out, C, dtype, stream);
}

}  // namespace lasr

namespace lasr {
int dwconv_tc_supported(int C, int K, int stride);
int dwconv_tc_fwd(const void* x, const float* w, void* y, const void* addend, int N, int T, int C, int K, int flip,
                  cudaStream_t stream);
int dwconv_tc_wgrad(const void* x, const void* dy, float* dw, int N, int T, int C, int K, cudaStream_t stream);
int dwconv_tc_bwd(const void* x, const void* dy, const float* w, const void* addend, void* dx, float* dw, int N, int T,
                  int C, int K, cudaStream_t stream);
}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_dwconv1d_fwd(const void* x, const float* wt, void* y, const void* addend, int N, int T_in, int T_out, int C,
                      int K, int stride, int flip, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T_in <= 0 || C <= 0 || (C & 1)) return LASR_ERR_BAD_SHAPE;
  if (T_out != (T_in + 2 * (K / 2) - K) / stride + 1) return LASR_ERR_BAD_SHAPE;
  if (flip && stride != 1) return LASR_ERR_UNSUPPORTED;
  if (dtype == LASR_F32) return fwd_t<float>(x, wt, y, addend, N, T_in, T_out, C, K, stride, flip, dtype, stream);
  if (dtype == LASR_BF16 && dwconv_tc_supported(C, K, stride))  // tensor-core Toeplitz path (dwconv_tc.cu)
    return dwconv_tc_fwd(x, wt, y, addend, N, T_in, C, K, flip, stream);
  if (dtype == LASR_BF16)
    return fwd_t<__nv_bfloat16>(x, wt, y, addend, N, T_in, T_out, C, K, stride, flip, dtype, stream);
  return LASR_ERR_BAD_DTYPE;
}

int lasr_dwconv1d_bwd(const void* x, const void* dy, const float* wt, const void* addend, void* dx, float* dwt, int N,
                      int T, int C, int K, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T <= 0 || C <= 0 || (C & 1)) return LASR_ERR_BAD_SHAPE;
  if (dtype == LASR_BF16 && dwconv_tc_supported(C, K, 1)) {
    const int rc = dwconv_tc_bwd(x, dy, wt, addend, dx, dwt, N, T, C, K, stream);
    if (rc != LASR_ERR_UNSUPPORTED) return rc;
  }
  const int rc1 = lasr_dwconv1d_wgrad(x, dy, dwt, N, T, T, C, K, 1, dtype, stream);
  if (rc1) return rc1;
  return lasr_dwconv1d_fwd(dy, wt, dx, addend, N, T, T, C, K, 1, 1, dtype, stream);
}

int lasr_dwconv1d_wgrad(const void* x, const void* dy, float* dwt, int N, int T_in, int T_out, int C, int K,
                        int stride, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T_in <= 0 || C <= 0 || (C & 1)) return LASR_ERR_BAD_SHAPE;
  if (T_out != (T_in + 2 * (K / 2) - K) / stride + 1) return LASR_ERR_BAD_SHAPE;
  if (dtype == LASR_F32) return wgrad_t<float>(x, dy, dwt, N, T_in, T_out, C, K, stride, dtype, stream);
  if (dtype == LASR_BF16 && dwconv_tc_supported(C, K, stride)) return dwconv_tc_wgrad(x, dy, dwt, N, T_in, C, K, stream);
  if (dtype == LASR_BF16) return wgrad_t<__nv_bfloat16>(x, dy, dwt, N, T_in, T_out, C, K, stride, dtype, stream);
  return LASR_ERR_BAD_DTYPE;
}

}  // extern "C"
