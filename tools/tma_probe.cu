// Micro-benchmark (tools/, not part of the library): how fast can ONE CTA per SM pull / push narrow column slices of a
// channels-last [M, C] bf16 matrix?  The depthwise kernels read and write 16-channel (32-byte) slices of every frame
// row; through LDG / STG that touches one 128-byte line per 32 useful bytes and the L1TEX wavefront rate becomes the
// limit.  This probe measures the alternatives: TMA tensor loads / stores of [R rows x W bytes] boxes for W = 32, 64, 128
// and the LDG / STG patterns the kernels use today.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

constexpr int STAGES = 8;
constexpr int ROWS = 256;  // rows per box

// each CTA streams `boxes` boxes of [ROWS x W bytes]: column slice (blockIdx % slices), rows advancing
__global__ void __launch_bounds__(128) tma_load_kernel(const __grid_constant__ CUtensorMap map, int wbytes, int slices,
                                                         int boxes, int row_groups, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  const int box_bytes = ROWS * wbytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int slice = blockIdx.x % slices;
  const int rg0 = (blockIdx.x / slices) * boxes;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < boxes + STAGES; ++i) {
      if (i >= STAGES || true) {
        if (i >= STAGES) mbar_wait(&full[(i - STAGES) % STAGES], ((i - STAGES) / STAGES) & 1);
      }
      if (i < boxes) {
        const int s = i % STAGES;
        mbar_expect(&full[s], box_bytes);
        tma_load_2d(smem + s * box_bytes, &map, &full[s], slice * (wbytes / 2), ((rg0 + i) % row_groups) * ROWS);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(128) tma_store_kernel(const __grid_constant__ CUtensorMap map, int wbytes, int slices,
                                                          int boxes, int row_groups, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int box_bytes = ROWS * wbytes;
  for (int i = threadIdx.x; i < box_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int slice = blockIdx.x % slices;
  const int rg0 = (blockIdx.x / slices) * boxes;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < boxes; ++i) {
      tma_store_2d(&map, smem, slice * (wbytes / 2), ((rg0 + i) % row_groups) * ROWS);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

// LDG.128: lane = (h = lane >> 4 channel octet, b = lane & 15 frame) -> a warp instruction covers 16 rows x W = 32 bytes
// (W = 32), 8 rows x 64 bytes (W = 64) or 4 rows x 128 bytes (W = 128); 8 loads in flight per lane like the producers
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ x, int C, int wbytes, int slices, int rows_per_cta,
                                                   int total_rows, unsigned long long* cycles, uint4* sink) {
  const int lanes_per_row = wbytes / 16;
  const int rows_per_warp = 32 / lanes_per_row;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_in = lane / lanes_per_row, c_in = lane % lanes_per_row;
  const int slice = blockIdx.x % slices;
  const long long row0 = static_cast<long long>(blockIdx.x / slices) * rows_per_cta;
  const int pitch16 = C * 2 / 16;
  uint4 acc = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = warp * rows_per_warp * 8; r < rows_per_cta; r += 16 * rows_per_warp * 8) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long row = (row0 + r + i * rows_per_warp + r_in) % total_rows;
      v[i] = __ldg(x + row * pitch16 + slice * lanes_per_row + c_in);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc.x ^= v[i].x; acc.y ^= v[i].y; acc.z ^= v[i].z; acc.w ^= v[i].w;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (acc.x == 0x12345678u) sink[threadIdx.x] = acc;
}

// STG 32 bytes per lane: each lane owns 16 consecutive rows (like the depthwise epilogue), a warp store touches 32 lines
__global__ void __launch_bounds__(128) stg_kernel(uint4* __restrict__ y, int C, int slices, int rows_per_cta, int total_rows,
                                                   unsigned long long* cycles) {
  const int slice = blockIdx.x % slices;
  const long long row0 = static_cast<long long>(blockIdx.x / slices) * rows_per_cta;
  const int pitch16 = C * 2 / 16;
  __syncthreads();
  const long long t0 = clock64();
  for (int base = 0; base < rows_per_cta; base += 128 * 16) {
    const int r = base + threadIdx.x * 16;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      const long long row = (row0 + r + j) % total_rows;
      uint4* p = y + row * pitch16 + slice * 2;
      asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(j) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int C = 512;
  const long long M = 32LL * 801 * 8;  // 8 layers' worth of rows so the data does not sit in L2 (210 MB)
  void* x;
  CK(cudaMalloc(&x, M * C * 2));
  CK(cudaMemset(x, 1, M * C * 2));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, 1024 * 8));
  uint4* sink;
  CK(cudaMalloc(&sink, 512 * 16));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const int ctas = 148;
  unsigned long long h[1024];
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int hot = 0; hot < 2; ++hot) {
    // hot = 1: a working set that stays in L2 (rows wrap inside 64 MB)
    const long long rows_ws = hot ? 32LL * 801 * 2 : M;
    const int row_groups = static_cast<int>(rows_ws / ROWS);
    printf("== %s working set: %.0f MB\n", hot ? "L2-resident" : "DRAM-sized", rows_ws * C * 2 / 1e6);
    for (int wbytes : {32, 64, 128}) {
      const int slices = C * 2 / wbytes;
      CUtensorMap map;
      cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(rows_ws)};
      cuuint64_t strides[1] = {static_cast<cuuint64_t>(C) * 2};
      cuuint32_t box[2] = {static_cast<cuuint32_t>(wbytes / 2), ROWS};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); return 1; }
      const int boxes = 256;  // per CTA
      const int smem = STAGES * ROWS * wbytes;
      CK(cudaFuncSetAttribute(tma_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        tma_load_kernel<<<ctas, 128, smem>>>(map, wbytes, slices, boxes, row_groups, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < ctas; ++i) avg += h[i];
      avg /= ctas;
      const double bytes = static_cast<double>(boxes) * ROWS * wbytes;
      printf("TMA load  W=%3d B: %7.1f B/clk/SM  (%6.2f TB/s chip, %.1f us)\n", wbytes, bytes / avg,
             bytes * ctas / (ms * 1e-3) / 1e12, ms * 1e3);
      CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * wbytes));
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        tma_store_kernel<<<ctas, 128, ROWS * wbytes>>>(map, wbytes, slices, boxes, row_groups, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
      avg = 0;
      for (int i = 0; i < ctas; ++i) avg += h[i];
      avg /= ctas;
      printf("TMA store W=%3d B: %7.1f B/clk/SM  (%6.2f TB/s chip, %.1f us)\n", wbytes, bytes / avg,
             bytes * ctas / (ms * 1e-3) / 1e12, ms * 1e3);
      // LDG pattern
      const int rows_per_cta = boxes * ROWS;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        ldg_kernel<<<ctas, 512>>>(static_cast<const uint4*>(x), C, wbytes, slices, rows_per_cta, static_cast<int>(rows_ws), cyc,
                                  sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
      avg = 0;
      for (int i = 0; i < ctas; ++i) avg += h[i];
      avg /= ctas;
      printf("LDG.128   W=%3d B: %7.1f B/clk/SM  (%6.2f TB/s chip, %.1f us)\n", wbytes, bytes / avg,
             bytes * ctas / (ms * 1e-3) / 1e12, ms * 1e3);
    }
    {
      const int slices = C * 2 / 32;
      const int rows_per_cta = 256 * ROWS;
      float ms;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        stg_kernel<<<ctas, 128>>>(static_cast<uint4*>(x), C, slices, rows_per_cta, static_cast<int>(rows_ws), cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < ctas; ++i) avg += h[i];
      avg /= ctas;
      const double bytes = static_cast<double>(rows_per_cta) * 32;
      printf("STG 32 B/lane, 16 rows per lane: %7.1f B/clk/SM  (%6.2f TB/s chip, %.1f us)\n", bytes / avg,
             bytes * ctas / (ms * 1e-3) / 1e12, ms * 1e3);
    }
  }
  printf("sm clock attr %d kHz\n", clock_khz);
  return 0;
}
