// Log-mel frontend on the tensor cores (replaces AudioParser.parse_audio, data_module.py:150-174, from the waveform
// tensor onward: pre-emphasis :157, MelSpectrogram(sr=16000, n_fft=512, pad=32, win_length=320, hop_length=160,
// n_mels=64) :68-70,160, AmplitudeToDB(power) :71,161, global (x - mean) / std normalisation :171-172).
//
// The STFT is a windowed-DFT GEMM:  spec[t, f] = sum_i e[160 t + i] * win[i] * exp(-2 pi j f i / 512), i < 320
// where e[] is the pre-emphasised, zero-padded (32) and reflect-padded (256) sample stream shifted by the 96 leading
// zeros of the centred window (a pure phase, irrelevant for the power).  Frames overlap (hop 160 < 320), so the A
// operand [frames, 320] is never materialised: a TMA map with a 160-element row pitch over the sample stream loads
// the overlapping frames straight into 128-byte-swizzled shared memory.
//
// Precision: the power spectrum spans > 60 dB and low-energy bins are cancellation sums, so the bf16
// tensor-core GEMM is error-compensated: samples and basis are each split into three bf16 terms (x = h + m + l,
// 24 mantissa bits in total) and the products hh, hm, mh, mm, hl, lh are accumulated into the SAME fp32 TMEM
// accumulator ("bf16 x 6"); `products` = 3 selects the cheaper hh + hm + mh.
//
//   lasr_logmel_prepare    waveform -> e[] split in 3 bf16 terms            (elementwise)
//   lasr_logmel_fwd        tcgen05 GEMM, 128 frames x 256 basis columns per accumulator (2 accumulators in TMEM, so the
//                          epilogue of one half overlaps the MMAs of the next); epilogue: re^2 + im^2, sparse
//                          triangular mel filterbank (<= 2 filters per bin), 10 log10(max(., 1e-10)), per-utterance
//                          sum / sum of squares (fp64 RED), dB features written channels-last [N, T, 64]
//   lasr_logmel_normalize  (dB - mean) / unbiased std, zero after each utterance's own length (the collate padding,
//                          data_module.py:230,243); writes the reference layout [N, 1, 64, T] fp32 and / or the
//                          encoder's channels-last [N, T, 64] input
#include "common.cuh"

namespace lasr {

constexpr int LM_NFFT = 512, LM_WIN = 320, LM_HOP = 160, LM_PAD = 32, LM_MELS = 64, LM_BINS = 257;
constexpr int LM_BM = 128, LM_BN = 256, LM_BK = 64;
constexpr int LM_KB = LM_WIN / LM_BK;  // 5 k-blocks per product
constexpr int LM_A_BYTES = LM_BM * LM_BK * 2, LM_B_BYTES = LM_BN * LM_BK * 2;
constexpr int LM_STAGE_BYTES = LM_A_BYTES + LM_B_BYTES;
constexpr int LM_STAGES = 3;
constexpr int LM_MEL_PITCH = LM_MELS + 1;
constexpr int LM_SMEM = LM_STAGES * LM_STAGE_BYTES + LM_BM * LM_MEL_PITCH * 4 + LM_BINS * 16 + 1024 + 256;

__device__ __forceinline__ int lm_frames(int num_samples) { return 1 + (num_samples + 2 * LM_PAD) / LM_HOP; }

// ------------------------------------------------------------------------------------------------
// prepare: parts[p][n][k] = bf16 term p of e_n[k], k < Lp
// ------------------------------------------------------------------------------------------------
template <typename W>
__device__ __forceinline__ float lm_sample_f32(W v);
template <>
__device__ __forceinline__ float lm_sample_f32<float>(float v) { return v; }
// 16-bit PCM -> [-1, 1): torchaudio.load(normalize=True) divides by 32768 (data_module.py:153)
template <>
__device__ __forceinline__ float lm_sample_f32<int16_t>(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }

// standard normal deviate of (seed, utterance, sample): the device-side dither `torch.randn_like(y)` of data_module.py:155
// (the reference's stream is unseeded; this one is a pure function of its key, so sample(s) and sample(s-1) agree
// between neighbouring threads)
__device__ __forceinline__ float lm_dither_normal(unsigned long long seed, int n, int i) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(n), 0x6d656cu, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float u1 = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

template <typename W>
__global__ void logmel_prepare_kernel(const W* __restrict__ wave, const float* __restrict__ dither,
                                      const unsigned long long dither_seed,
                                      const unsigned long long* __restrict__ seed_dev,
                                      const int32_t* __restrict__ starts, const int32_t* __restrict__ num_samples,
                                      __nv_bfloat16* __restrict__ parts, int N, int S_max, int Lp) {
  const int n = blockIdx.y;
  const int S = num_samples[n];
  const int L = S + 2 * LM_PAD;  // after MelSpectrogram(pad=32)
  // train-time crop (sub_secquence, data_module.py:138-148,158-159) happens AFTER dither + pre-emphasis of the whole
  // utterance: the cropped stream starts at sample `st` of the pre-emphasised signal, so its first sample still sees
  // its predecessor
  const int st = starts != nullptr ? starts[n] : 0;
  const W* w = wave + static_cast<size_t>(n) * S_max + st;
  const float* dth = dither ? dither + static_cast<size_t>(n) * S_max + st : nullptr;
  const bool gen = dither == nullptr && dither_seed != 0ull;
  const unsigned long long seed = dither_seed + (seed_dev != nullptr ? __ldg(seed_dev) : 0ull);
  const size_t plane = static_cast<size_t>(N) * Lp;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Lp; k += gridDim.x * blockDim.x) {
    // k + 96 indexes the reflect-padded stream; j indexes the zero-padded one
    int j = k + 96 - LM_NFFT / 2;
    float v = 0.f;
    if (j < L + LM_NFFT / 2) {
      if (j < 0) j = -j;
      if (j >= L) j = 2 * (L - 1) - j;
      const int s = j - LM_PAD;
      if (s >= 0 && s < S) {
        auto sample = [&](int i) {
          float x = lm_sample_f32<W>(w[i]);
          if (dth) x = __fadd_rn(x, __fmul_rn(1e-5f, dth[i]));  // y += 1e-5 * randn  (data_module.py:155)
          if (gen) x = __fadd_rn(x, __fmul_rn(1e-5f, lm_dither_normal(seed, n, st + i)));
          return x;
        };
        v = sample(s);
        if (s + st > 0) v = __fsub_rn(v, __fmul_rn(0.97f, sample(s - 1)));  // :157, two roundings like torch
      }
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    const __nv_bfloat16 l = __float2bfloat16_rn(r2);
    const size_t o = static_cast<size_t>(n) * Lp + k;
    parts[o] = h;
    parts[plane + o] = m;
    parts[2 * plane + o] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// GEMM + power + mel + dB + statistics
// ------------------------------------------------------------------------------------------------
struct LogmelParams {
  int N, T_max, t_blocks, products;
  int pa[8], pb[8];             // product p multiplies sample term pa[p] with basis term pb[p]
  const int32_t* mel_idx;       // [257][2] filter indices fed by each bin
  const float* mel_w;           // [257][2] their weights (0 when unused)
  const int32_t* num_samples;   // [N]
  float* db;                    // [N, T_max, 64]
  double* stats;                // [N, 2]
};

struct MelEntry {
  int i0, i1;
  float w0, w1;
};

__global__ void __launch_bounds__(256, 1)
logmel_fwd_kernel(const __grid_constant__ CUtensorMap tma_a0, const __grid_constant__ CUtensorMap tma_a1,
                  const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b,
                  const LogmelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* smel = reinterpret_cast<float*>(smem + LM_STAGES * LM_STAGE_BYTES);          // [128][65]
  MelEntry* stab = reinterpret_cast<MelEntry*>(smel + LM_BM * LM_MEL_PITCH);            // [257]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stab) + LM_BINS * 16);
  full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~uintptr_t(7));
  uint64_t* empty_bar = full_bar + LM_STAGES;
  uint64_t* tmem_full_bar = empty_bar + LM_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < LM_BINS; i += blockDim.x) {
    MelEntry e;
    e.i0 = p.mel_idx[2 * i];
    e.i1 = p.mel_idx[2 * i + 1];
    e.w0 = p.mel_w[2 * i];
    e.w1 = p.mel_w[2 * i + 1];
    stab[i] = e;
  }
  for (int i = threadIdx.x; i < LM_BM * LM_MEL_PITCH; i += blockDim.x) smel[i] = 0.f;
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a0);
    tma_prefetch_desc(&tma_a1);
    tma_prefetch_desc(&tma_a2);
    tma_prefetch_desc(&tma_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < LM_STAGES; ++s) {
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
    tmem_alloc(tmem_ptr_smem, 2 * LM_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // work unit u = (tile, half): tile = (utterance n, 128-frame block), half = 256 basis columns = 128 bins
  const int num_tiles = p.N * p.t_blocks;
  const int kbs = p.products * LM_KB;

  if (warp_idx == 0) {
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n = tile / p.t_blocks;
        const int t0 = (tile - n * p.t_blocks) * LM_BM;
        for (int half = 0; half < 2; ++half) {
          for (int kb = 0; kb < kbs; ++kb) {
            const int prod = kb / LM_KB, k0 = (kb - prod * LM_KB) * LM_BK;
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * LM_STAGE_BYTES;
            uint8_t* sb = sa + LM_A_BYTES;
            if (leader) {
              mbar_arrive_expect_tx(&full_bar[stage], LM_STAGE_BYTES);
              const int ap = p.pa[prod];
              const CUtensorMap* ma = ap == 0 ? &tma_a0 : (ap == 1 ? &tma_a1 : &tma_a2);
              tma_load_3d(sa, ma, &full_bar[stage], k0, t0, n);
              tma_load_2d(sb, &tma_b, &full_bar[stage], k0, p.pb[prod] * LM_NFFT + half * LM_BN);
            }
            __syncwarp();
            if (++stage == LM_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(LM_BM, LM_BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int unit = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int half = 0; half < 2; ++half, ++unit) {
          const int acc = unit & 1;
          const uint32_t acc_phase = (unit >> 1) & 1;
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * LM_BN;
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * LM_STAGE_BYTES);
            const uint32_t sb = sa + LM_A_BYTES;
#pragma unroll
            for (int k = 0; k < LM_BK / 16; ++k) {
              const uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024);
              const uint64_t db = umma_desc_sw128(sb + k * 32, 16, 1024);
              if (leader) umma_bf16(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            if (leader) umma_commit(&empty_bar[stage]);
            __syncwarp();
            if (++stage == LM_STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          if (leader) umma_commit(&tmem_full_bar[acc]);
          __syncwarp();
        }
      }
    }
  } else if (warp_idx >= 4) {
    const int ew = warp_idx - 4;
    const int row = ew * 32 + lane;
    float* myrow = smel + row * LM_MEL_PITCH;
    int unit = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n = tile / p.t_blocks;
      const int t0 = (tile - n * p.t_blocks) * LM_BM;
      for (int half = 0; half < 2; ++half, ++unit) {
        const int acc = unit & 1;
        const uint32_t acc_phase = (unit >> 1) & 1;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * LM_BN;
#pragma unroll 1
        for (int ch = 0; ch < LM_BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(taddr0 + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int b = 0; b < 16; ++b) {
            const int f = half * 128 + ch * 16 + b;
            const float re = __uint_as_float(v[2 * b]), im = __uint_as_float(v[2 * b + 1]);
            float pw;
            if (f == 0) {
              // column 1 of the basis carries the (purely real) Nyquist bin 256 instead of im(0) == 0
              const MelEntry ny = stab[LM_BINS - 1];
              const float pn = im * im;
              myrow[ny.i0] += ny.w0 * pn;
              myrow[ny.i1] += ny.w1 * pn;
              pw = re * re;
            } else {
              pw = re * re + im * im;
            }
            const MelEntry e = stab[f];
            myrow[e.i0] += e.w0 * pw;
            myrow[e.i1] += e.w1 * pw;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
      // both halves of this tile are in smel: dB, statistics, coalesced store (a warp owns its 32 rows)
      __syncwarp();
      const int T_n = lm_frames(p.num_samples[n]);
      float s1 = 0.f, s2 = 0.f;
      for (int r = 0; r < 32; ++r) {
        float* src = smel + (ew * 32 + r) * LM_MEL_PITCH;
        const int t = t0 + ew * 32 + r;
        const float m0 = src[lane], m1 = src[lane + 32];
        src[lane] = 0.f;
        src[lane + 32] = 0.f;
        // AmplitudeToDB(stype="power"): 10 * log10(clamp(x, min=1e-10))
        const float d0 = 10.f * log10f(fmaxf(m0, 1e-10f));
        const float d1 = 10.f * log10f(fmaxf(m1, 1e-10f));
        if (t < p.T_max) {
          float* dst = p.db + (static_cast<size_t>(n) * p.T_max + t) * LM_MELS;
          dst[lane] = d0;
          dst[lane + 32] = d1;
          if (t < T_n) {
            s1 += d0 + d1;
            s2 += d0 * d0 + d1 * d1;
          }
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0 && t0 + ew * 32 < T_n) {
        atomicAdd(p.stats + 2 * n, static_cast<double>(s1));
        atomicAdd(p.stats + 2 * n + 1, static_cast<double>(s2));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * LM_BN);
  }
}

// ------------------------------------------------------------------------------------------------
// normalise
// ------------------------------------------------------------------------------------------------
// SpecAugment in the dB domain, BEFORE the normalisation (spec_augment, data_module.py:97-122,163-165): one frequency
// band [f0, f0+fw) over all frames and one time band [t0, t0+tw) over all mel bins are set to 0; the utterance's
// (sum, sum of squares) statistics are corrected by what was removed, so the normalise pass sees exactly the masked
// spectrogram.  Band positions come from the host (the reference draws them from an unseeded random.Random()).
// grid (chunks, N), 256 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spec_augment_kernel(float* __restrict__ db, double* __restrict__ stats, const int32_t* __restrict__ num_samples,
                    const int32_t* __restrict__ bands, int T_max) {
  const int n = blockIdx.y;
  const int T_n = lm_frames(num_samples[n]);
  const int f0 = bands[4 * n], fw = bands[4 * n + 1], t0 = bands[4 * n + 2], tw = bands[4 * n + 3];
  float* d = db + static_cast<size_t>(n) * T_max * LM_MELS;
  double s1 = 0.0, s2 = 0.0;
  const int total = T_n * LM_MELS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i / LM_MELS, m = i - t * LM_MELS;
    const bool masked = (m >= f0 && m < f0 + fw) || (t >= t0 && t < t0 + tw);
    if (masked) {
      const float v = d[i];
      s1 += static_cast<double>(v);
      s2 += static_cast<double>(v) * static_cast<double>(v);
      d[i] = 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0 && (s1 != 0.0 || s2 != 0.0)) {
    atomicAdd(stats + 2 * n, -s1);
    atomicAdd(stats + 2 * n + 1, -s2);
  }
}

// ------------------------------------------------------------------------------------------------
// The random draws of parse_audio(mask=True) for one utterance, on the device: restates the arithmetic
// of sub_secquence(weight=0.98) (data_module.py:138-148) and spec_augment(freq_mask=27, time_mask=0.07) (:97-122) in
// the reference's draw order, random.uniform(a, b) = a + (b - a) * random() in double precision:
//   target_length = int(S * U(0.98, 1));  location = int(U(0, S - target_length));  kept = max(target_length - location, 0)
//   T = 1 + (kept + 64) / 160;  w_x = int(U(0, 27));  w_y = int(U(0, int(T * 0.07)));
//   rect_x = int(U(0, 64 - w_x));  rect_y = int(U(0, T - w_y))
// `uniforms` [N, 6] double (nullable): the six random() values per utterance supplied by the caller (parity hook: feed
// the reference's own seeded random.Random() stream); otherwise Philox keyed by (seed + *seed_dev, utterance).
// crop == 0 keeps the whole utterance, spec == 0 writes empty bands.  percents[n] = T_n / T_max (data_module.py:244).
// ------------------------------------------------------------------------------------------------
__global__ void augment_draw_kernel(const int32_t* __restrict__ num_samples, const double* __restrict__ uniforms,
                                    unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                                    int32_t* __restrict__ starts, int32_t* __restrict__ kept, int32_t* __restrict__ bands,
                                    float* __restrict__ percents, int N, int T_max, int crop, int spec) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double u[6];
  if (uniforms != nullptr) {
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = uniforms[6 * n + i];
  } else {
    const unsigned long long sd = seed + (seed_dev != nullptr ? __ldg(seed_dev) : 0ull);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(n), static_cast<uint32_t>(h), 0x617567u, 0u),
                                    make_uint2(static_cast<uint32_t>(sd), static_cast<uint32_t>(sd >> 32)));
      // 53-bit uniforms in [0, 1) like Python's random(): (a >> 5) * 2^26 + (b >> 6), / 2^53 -- three per Philox call
      // would need 6 words; use 2 x (27 + 26 bits) from (x, y), (z, w) and a third from a second lane of the counter
      u[3 * h + 0] = (static_cast<double>(r.x >> 5) * 67108864.0 + static_cast<double>(r.y >> 6)) * (1.0 / 9007199254740992.0);
      u[3 * h + 1] = (static_cast<double>(r.z >> 5) * 67108864.0 + static_cast<double>(r.w >> 6)) * (1.0 / 9007199254740992.0);
      const uint4 q = philox4x32_10(make_uint4(static_cast<uint32_t>(n), static_cast<uint32_t>(h), 0x617567u, 1u),
                                    make_uint2(static_cast<uint32_t>(sd), static_cast<uint32_t>(sd >> 32)));
      u[3 * h + 2] = (static_cast<double>(q.x >> 5) * 67108864.0 + static_cast<double>(q.y >> 6)) * (1.0 / 9007199254740992.0);
    }
  }
  const int S = num_samples[n];
  int loc = 0, k = S;
  if (crop) {
    const int target_length = static_cast<int>(static_cast<double>(S) * (0.98 + (1.0 - 0.98) * u[0]));
    loc = static_cast<int>(0.0 + (static_cast<double>(S - target_length) - 0.0) * u[1]);
    k = target_length - loc;
    if (k < 0) k = 0;
  }
  const int T = lm_frames(k);
  int w_x = 0, w_y = 0, rect_x = 0, rect_y = 0;
  if (spec) {
    w_x = static_cast<int>(27.0 * u[2]);
    w_y = static_cast<int>(static_cast<double>(static_cast<int>(static_cast<double>(T) * 0.07)) * u[3]);
    rect_x = static_cast<int>(static_cast<double>(LM_MELS - w_x) * u[4]);
    rect_y = static_cast<int>(static_cast<double>(T - w_y) * u[5]);
  }
  starts[n] = loc;
  kept[n] = k;
  bands[4 * n + 0] = rect_x;
  bands[4 * n + 1] = w_x;
  bands[4 * n + 2] = rect_y;
  bands[4 * n + 3] = w_y;
  if (percents != nullptr) percents[n] = static_cast<float>(static_cast<double>(T) / static_cast<double>(T_max));
}

// ------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void logmel_normalize_kernel(const float* __restrict__ db, const double* __restrict__ stats,
                                        const int32_t* __restrict__ num_samples, float* __restrict__ out_nct,
                                        OutT* __restrict__ out_ntc, int T_max) {
  __shared__ float tile[64][LM_MELS + 1];
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * 64;
  const int T_n = lm_frames(num_samples[n]);
  const double cnt = static_cast<double>(T_n) * LM_MELS;
  const double mean = stats[2 * n] / cnt;
  // torch.std_mean: unbiased
  const double var = (stats[2 * n + 1] - stats[2 * n] * mean) / (cnt - 1.0);
  const float fmean = static_cast<float>(mean);
  const float fstd = static_cast<float>(sqrt(var > 0.0 ? var : 0.0));
  for (int i = threadIdx.x; i < 64 * LM_MELS; i += blockDim.x) {
    const int r = i / LM_MELS, m = i - r * LM_MELS;
    const int t = t0 + r;
    float v = 0.f;
    if (t < T_n) v = (db[(static_cast<size_t>(n) * T_max + t) * LM_MELS + m] - fmean) / fstd;
    tile[r][m] = v;
    if (out_ntc != nullptr && t < T_max)
      out_ntc[(static_cast<size_t>(n) * T_max + t) * LM_MELS + m] = from_f32<OutT>(v);
  }
  if (out_nct == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * LM_MELS; i += blockDim.x) {
    const int m = i / 64, r = i - m * 64;
    const int t = t0 + r;
    if (t < T_max) out_nct[(static_cast<size_t>(n) * LM_MELS + m) * T_max + t] = tile[r][m];
  }
}

}  // namespace lasr

using namespace lasr;

extern "C" {

int lasr_logmel_padded_len(int T_max) { return LM_HOP * (T_max + 1); }

int lasr_logmel_prepare_crop(const float* wave, const float* dither, const int32_t* starts,
                             const int32_t* num_samples, void* parts, int N, int S_max, int T_max,
                             lasr_stream_t stream) {
  if (N <= 0 || S_max <= 0 || T_max <= 0) return LASR_ERR_BAD_SHAPE;
  return lasr_logmel_prepare_wave(wave, LASR_WAVE_F32, dither, 0ull, nullptr, starts, num_samples, parts, N, S_max, T_max,
                                 stream);
}

int lasr_logmel_prepare_wave(const void* wave, int wave_dtype, const float* dither, uint64_t dither_seed,
                             const uint64_t* seed_dev, const int32_t* starts, const int32_t* num_samples, void* parts,
                             int N, int S_max, int T_max, lasr_stream_t stream) {
  if (N <= 0 || S_max <= 0 || T_max <= 0 || wave == nullptr) return LASR_ERR_BAD_SHAPE;
  const int Lp = lasr_logmel_padded_len(T_max);
  dim3 grid(cdiv(Lp, 256 * 4) < 1 ? 1 : cdiv(Lp, 256 * 4), N);
  const unsigned long long* sd = reinterpret_cast<const unsigned long long*>(seed_dev);
  if (wave_dtype == LASR_WAVE_F32)
    logmel_prepare_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(wave), dither, dither_seed, sd, starts,
                                                           num_samples, static_cast<__nv_bfloat16*>(parts), N, S_max, Lp);
  else if (wave_dtype == LASR_WAVE_I16)
    logmel_prepare_kernel<int16_t><<<grid, 256, 0, stream>>>(static_cast<const int16_t*>(wave), dither, dither_seed, sd,
                                                             starts, num_samples, static_cast<__nv_bfloat16*>(parts), N,
                                                             S_max, Lp);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_augment_draw(const int32_t* num_samples, const double* uniforms, uint64_t seed, const uint64_t* seed_dev,
                      int32_t* starts, int32_t* kept, int32_t* bands, float* percents, int N, int T_max, int crop,
                      int spec, lasr_stream_t stream) {
  if (N <= 0 || T_max <= 0 || num_samples == nullptr || starts == nullptr || kept == nullptr || bands == nullptr)
    return LASR_ERR_BAD_SHAPE;
  augment_draw_kernel<<<cdiv(N, 128), 128, 0, stream>>>(num_samples, uniforms, seed,
                                                        reinterpret_cast<const unsigned long long*>(seed_dev), starts,
                                                        kept, bands, percents, N, T_max, crop, spec);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_logmel_prepare(const float* wave, const float* dither, const int32_t* num_samples, void* parts, int N,
                        int S_max, int T_max, lasr_stream_t stream) {
  return lasr_logmel_prepare_crop(wave, dither, nullptr, num_samples, parts, N, S_max, T_max, stream);
}

int lasr_spec_augment(float* db, double* stats, const int32_t* num_samples, const int32_t* bands, int N, int T_max,
                      lasr_stream_t stream) {
  if (N <= 0 || T_max <= 0 || db == nullptr || stats == nullptr || bands == nullptr) return LASR_ERR_BAD_SHAPE;
  dim3 grid(8, N);
  spec_augment_kernel<<<grid, 256, 0, stream>>>(db, stats, num_samples, bands, T_max);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_logmel_fwd(const void* parts, const void* basis, const int32_t* mel_idx, const float* mel_w,
                    const int32_t* num_samples, float* db, double* stats, int N, int T_max, int products,
                    lasr_stream_t stream) {
  if (N <= 0 || T_max <= 0) return LASR_ERR_BAD_SHAPE;
  if (products != 1 && products != 3 && products != 6) return LASR_ERR_UNSUPPORTED;
  const int Lp = lasr_logmel_padded_len(T_max);
  CUtensorMap ta[3], tb;
  const __nv_bfloat16* pbase = static_cast<const __nv_bfloat16*>(parts);
  for (int i = 0; i < 3; ++i) {
    // overlapping frames: row t starts 160 samples after row t-1 and is 320 samples long
    const uint64_t dims[3] = {LM_WIN, static_cast<uint64_t>(T_max), static_cast<uint64_t>(N)};
    const uint64_t strides[2] = {LM_HOP * 2, static_cast<uint64_t>(Lp) * 2};
    const uint32_t box[3] = {LM_BK, LM_BM, 1};
    int rc = make_tmap_nd_bf16(&ta[i], pbase + static_cast<size_t>(i) * N * Lp, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {LM_WIN, 3 * LM_NFFT};
    const uint64_t strides[1] = {LM_WIN * 2};
    const uint32_t box[2] = {LM_BK, LM_BN};
    int rc = make_tmap_nd_bf16(&tb, basis, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  LogmelParams p{};
  p.N = N;
  p.T_max = T_max;
  p.t_blocks = cdiv(T_max, LM_BM);
  p.products = products;
  // largest terms first; the small cross terms land on an accumulator that already holds the bulk
  const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
  for (int i = 0; i < 6; ++i) {
    p.pa[i] = pa[i];
    p.pb[i] = pb[i];
  }
  p.mel_idx = mel_idx;
  p.mel_w = mel_w;
  p.num_samples = num_samples;
  p.db = db;
  p.stats = stats;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(logmel_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM);
    if (e != cudaSuccess) {
      lasr_set_cuda_error(e);
      return LASR_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles = N * p.t_blocks;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  logmel_fwd_kernel<<<grid, 256, LM_SMEM, stream>>>(ta[0], ta[1], ta[2], tb, p);
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

int lasr_logmel_normalize(const float* db, const double* stats, const int32_t* num_samples, float* out_nct,
                          void* out_ntc, int N, int T_max, int dtype, lasr_stream_t stream) {
  if (N <= 0 || T_max <= 0) return LASR_ERR_BAD_SHAPE;
  dim3 grid(cdiv(T_max, 64), N);
  if (dtype == LASR_F32)
    logmel_normalize_kernel<float><<<grid, 256, 0, stream>>>(db, stats, num_samples, out_nct,
                                                             static_cast<float*>(out_ntc), T_max);
  else if (dtype == LASR_BF16)
    logmel_normalize_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(db, stats, num_samples, out_nct,
                                                                     static_cast<__nv_bfloat16*>(out_ntc), T_max);
  else
    return LASR_ERR_BAD_DTYPE;
  LASR_CHECK_LAUNCH();
  return LASR_OK;
}

}  // extern "C"
