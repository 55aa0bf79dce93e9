// fp32 pointwise-convolution GEMMs on the FFMA pipe: the exact-fp32 parity mode (single-pass
// TF32/bf16 tensor-core operands cannot meet rel 1e-4, so the fp32 mode stays on fp32 FMAs).
//   nt:       y[M, N]  = x[M, K] * w[N, K]^T (+bias) with the MaskCNN row mask
//   tn_accum: dw[Cout, Cin] += dy[R, Cout]^T * x[R, Cin]   (split over R, fp32 atomics)
//   colstats: per-32-row-group column sum / sum of squares in the layout the tcgen05 epilogue produces
#include "common.cuh"

namespace lasr {

// 128 x 64 output tile, BK = 16, 256 threads, 8 x 4 outputs per thread
__global__ void __launch_bounds__(256)
gemm_simt_nt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                    const float* __restrict__ bias, const int32_t* __restrict__ lengths, int T, int M, int N, int K,
                    int lda, int ldb, int ldc) {
  constexpr int BM = 128, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;  // tx: 16 column groups of 4, ty: 16 row groups of 8
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // software pipeline: the next k-tile's 8 + 4 values per thread are fetched into registers before the current tile is
  // multiplied, and stored to shared memory after it (the unpipelined loop exposed one global-load latency per 16 k:
  // at M = 2004 rows -- 128 CTAs, less than a wave -- that latency WAS the kernel)
  float ra[8], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int gr = m0 + r, gk = k0 + kk;
      ra[i] = (gr < M && gk < K) ? A[static_cast<size_t>(gr) * lda + gk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int gr = n0 + r, gk = k0 + kk;
      rb[i] = (gr < N && gk < K) ? B[static_cast<size_t>(gr) * ldb + gk] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      As[idx & 15][idx >> 4] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      Bs[idx & 15][idx >> 4] = rb[i];
    }
  };
  fetch(0);
  stash();
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) stash();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + ty * 8 + i;
    if (r >= M) continue;
    bool keep = true;
    if (lengths != nullptr) {
      const int n = r / T;
      keep = (r - n * T) < lengths[n];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < N) {
        float v = acc[i][j] + (bias ? bias[c] : 0.f);
        C[static_cast<size_t>(r) * ldc + c] = keep ? v : 0.f;
      }
    }
  }
}

// dw[Cout, Cin] += dy[r0:r1, Cout]^T x[r0:r1, Cin]; 64 x 64 tile, 4 x 4 per thread
__global__ void __launch_bounds__(256)
gemm_simt_tn_kernel(const float* __restrict__ DY, const float* __restrict__ X, float* __restrict__ DW, int R,
                    int Cout, int Cin, int lddy, int ldx, int lddw, int rows_per_split) {
  constexpr int BT = 64, BK = 16;
  __shared__ float As[BK][BT];
  __shared__ float Bs[BK][BT];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BT, n0 = blockIdx.y * BT;
  const int r0 = blockIdx.z * rows_per_split;
  const int r1 = min(R, r0 + rows_per_split);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];  // the next k-tile, fetched while the current one is multiplied (see gemm_simt_nt_kernel)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx >> 6, c = idx & 63;
      const int gr = k0 + kk;
      ra[i] = (gr < r1 && m0 + c < Cout) ? DY[static_cast<size_t>(gr) * lddy + m0 + c] : 0.f;
      rb[i] = (gr < r1 && n0 + c < Cin) ? X[static_cast<size_t>(gr) * ldx + n0 + c] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      As[idx >> 6][idx & 63] = ra[i];
      Bs[idx >> 6][idx & 63] = rb[i];
    }
  };
  fetch(r0);
  stash();
  __syncthreads();
  for (int k0 = r0; k0 < r1; k0 += BK) {
    const bool more = k0 + BK < r1;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) stash();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Cin) atomicAdd(&DW[static_cast<size_t>(m) * lddw + n], acc[i][j]);
    }
  }
}

// column statistics of a [M, N] matrix: stats[0][c] += sum_r y[r, c], stats[1][c] += sum_r y[r, c]^2 (double, RED)
template <typename T>
__global__ void __launch_bounds__(256)
colstats_kernel(const T* __restrict__ y, double* __restrict__ stats, int M, int N, int ld, int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int rb = r0; rb < r1; rb += 32) {
      float ps = 0.f, pq = 0.f;
      const int re = min(r1, rb + 32);
      for (int r = rb; r < re; ++r) {
        const float v = to_f32<T>(y[static_cast<size_t>(r) * ld + c]);
        ps += v;
        pq = fmaf(v, v, pq);
      }
      s += ps;
      q += pq;
    }
    atomicAdd(stats + c, s);
    atomicAdd(stats + N + c, q);
  }
}

// C[M, N] = A[M, K] * B[K, N] (row-major B): fp32 data gradient dx = dy * W with W in its native [Cout, Cin] layout
__global__ void __launch_bounds__(256)
gemm_simt_nn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N,
                    int K, int lda, int ldb, int ldc) {
  constexpr int BM = 128, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[4];  // the next k-tile, fetched while the current one is multiplied (see gemm_simt_nt_kernel)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int gr = m0 + r, gk = k0 + kk;
      ra[i] = (gr < M && gk < K) ? A[static_cast<size_t>(gr) * lda + gk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx >> 6, c = idx & 63;
      const int gk = k0 + kk, gc = n0 + c;
      rb[i] = (gk < K && gc < N) ? B[static_cast<size_t>(gk) * ldb + gc] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      As[idx & 15][idx >> 4] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      Bs[idx >> 6][idx & 63] = rb[i];
    }
  };
  fetch(0);
  stash();
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) stash();
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + ty * 8 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < N) C[static_cast<size_t>(r) * ldc + c] = acc[i][j];
    }
  }
}

int gemm_simt_nn(const float* a, const float* b, float* c, int M, int N, int K, int lda, int ldb, int ldc,
                 cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return LASR_ERR_BAD_SHAPE;
  dim3 grid(cdiv(M, 128), cdiv(N, 64));
  gemm_simt_nn_kernel<<<grid, 256, 0, stream>>>(a, b, c, M, N, K, lda, ldb, ldc);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int gemm_simt_nt(const float* a, const float* b, float* c, const float* bias, const int32_t* lengths, int T, int M,
                 int N, int K, int lda, int ldb, int ldc, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return LASR_ERR_BAD_SHAPE;
  dim3 grid(cdiv(M, 128), cdiv(N, 64));
  gemm_simt_nt_kernel<<<grid, 256, 0, stream>>>(a, b, c, bias, lengths, T, M, N, K, lda, ldb, ldc);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int gemm_simt_tn_accum(const float* dy, const float* x, float* dw, int R, int Cout, int Cin, int lddy, int ldx,
                       int lddw, cudaStream_t stream) {
  if (R <= 0 || Cout <= 0 || Cin <= 0) return LASR_ERR_BAD_SHAPE;
  const int tiles = cdiv(Cout, 64) * cdiv(Cin, 64);
  int splits = (4 * kNumSMs) / tiles;
  if (splits < 1) splits = 1;
  int rows_per_split = cdiv(R, splits);
  rows_per_split = cdiv(rows_per_split, 16) * 16;
  splits = cdiv(R, rows_per_split);
  dim3 grid(cdiv(Cout, 64), cdiv(Cin, 64), splits);
  gemm_simt_tn_kernel<<<grid, 256, 0, stream>>>(dy, x, dw, R, Cout, Cin, lddy, ldx, lddw, rows_per_split);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int colstats(const void* y, double* stats, int M, int N, int ld, int dtype, cudaStream_t stream) {
  const int rows_per_block = 128;
  const int blocks = cdiv(M, rows_per_block);
  if (dtype == LASR_F32)
    colstats_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(y), stats, M, N, ld, rows_per_block);
  else
    colstats_kernel<__nv_bfloat16>
        <<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), stats, M, N, ld, rows_per_block);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // namespace lasr
