// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[n, ho, wo, co] = sum_src sum_(r,s) sum_ci  act_src[n, ho*sy + r - pad, wo*sx + s - pad, ci]
//                                                  * wgt_src[tap(r,s)][co][ci]
//
// One kernel serves forward convs (multi-source == virtual channel concat) and dgrad
// (act = dY of every consumer, wgt = per-consumer transposed/flipped pack). Activations are
// NHWC bf16 with channels padded to 64; a tile of 128 output pixels (tn x th x tw) is the UMMA M,
// the Cout tile BN is the UMMA N, and the K loop walks (source, tap, 64-channel chunk). Each
// k-iteration is one 4-D TMA box of the shifted input window (zero fill supplies the padding) and one
// 3-D TMA box of weights, both landing in 128B-swizzled K-major shared memory.
//
// Warp roles: warp0 TMA producer, warp1 MMA issuer (+TMEM owner), warps2-5 epilogue
// (TMEM -> regs -> bias/act -> bf16 -> swizzled smem -> TMA store, plus per-channel sum / sum-of-
// squares partials for InstanceNorm). TMEM accumulators are double buffered so the epilogue of tile i
// overlaps the MMAs of tile i+1. Persistent CTAs, static round-robin tile schedule.
#pragma once
#include "tg_ptx.cuh"

namespace tg {

constexpr int kMaxSrc = 6;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * kChunkK * 2;     // 16 KiB
constexpr int kStoreBytes = kTileM * 64 * 2;      // one 64-channel output chunk
constexpr int kNumThreads = 192;

enum : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_SIGMOID = 2, ACT_RELU = 3, ACT_TANH = 4 };

struct alignas(64) IgemmSrc {
  CUtensorMap act;  // {C, W, H, N}
  CUtensorMap wgt;  // {K, rows, taps}
  int c_chunks;
  int pad_[15];
};

struct alignas(64) IgemmParams {
  IgemmSrc src[kMaxSrc];
  CUtensorMap out;  // {Cout, Wo, Ho, N}
  int num_src;
  int taps, stride;                 // number of taps; input pixel = out pixel * stride + tap offset
  int8_t tap_dy[16], tap_dx[16];    // per-tap input offset (padding / phase folded in)
  int8_t tap_w[16];                 // per-tap index into the weight tensor's tap axis
  int Ho, Wo, N;
  int th, tw, tn;
  int tiles_h, tiles_w, tiles_img;  // tiles_img = ceil(N / tn)
  int n_tiles;                      // Cout / BN
  int act;
  float slope;
  const float* bias;      // [bias_len] or nullptr
  int bias_len;
  float* stats_partial;   // [N][stats_tiles_total][Cout][2] or nullptr (requires tn == 1)
  int stats_tiles_total, stats_tile_off;
  int cout;               // padded Cout (row pitch of stats)
  int pool_out;           // epilogue stores the 2x2 sum (tile is 16x8, `out` map is the half-resolution tensor)
  int* err_flag;
  // split-K: item = (tile, split); split s runs k-iterations [k_iters*s/splits, k_iters*(s+1)/splits) and writes its fp32
  // partial tile to ws + s*ws_slice (pixel-major [N*Ho*Wo][cout]); splitk_finalize_kernel sums the slices in order
  int splits;
  float* ws;
  long long ws_slice;
};

// Second half of a split-K convolution: out = act(sum_s ws[s] + bias) as bf16 NHWC through the output view's strides.
struct SplitFinParams {
  const float* ws;
  long long ws_slice;
  int splits;
  __nv_bfloat16* out;
  long long sn, sh, sw;   // element strides of the output view
  int N, Ho, Wo, cout;
  const float* bias;
  int bias_len, act;
  float slope;
};

__global__ void __launch_bounds__(256) splitk_finalize_kernel(const SplitFinParams p) {
  griddep_sync();
  const int groups = p.cout >> 3;
  const long long total = (long long)p.N * p.Ho * p.Wo * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = int(i % groups);
    const long long pix = i / groups;
    const int wo = int(pix % p.Wo), ho = int((pix / p.Wo) % p.Ho), n = int(pix / ((long long)p.Wo * p.Ho));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* src = p.ws + pix * p.cout + g * 8;
    for (int s = 0; s < p.splits; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(src + s * p.ws_slice);
      const float4 b = *reinterpret_cast<const float4*>(src + s * p.ws_slice + 4);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0 = acc[2 * j], x1 = acc[2 * j + 1];
      const int c = g * 8 + 2 * j;
      if (p.bias) {
        if (c < p.bias_len) x0 += p.bias[c];
        if (c + 1 < p.bias_len) x1 += p.bias[c + 1];
      }
      if (p.act == ACT_LRELU) { x0 = x0 > 0.f ? x0 : x0 * p.slope; x1 = x1 > 0.f ? x1 : x1 * p.slope; }
      else if (p.act == ACT_RELU) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
      else if (p.act == ACT_SIGMOID) { x0 = 1.f / (1.f + __expf(-x0)); x1 = 1.f / (1.f + __expf(-x1)); }
      else if (p.act == ACT_TANH) { x0 = tanhf(x0); x1 = tanhf(x1); }
      w[j] = pack_bf16x2(x0, x1);
    }
    *reinterpret_cast<uint4*>(p.out + n * p.sn + ho * p.sh + wo * p.sw + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int BN>
struct IgemmCfg {
  static constexpr int kBBytes = BN * kChunkK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 5 : 6);
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemStages = kStages * kStageBytes;
  static constexpr int kSmemStore = 2 * kStoreBytes;
  static constexpr int kSmemScratch = 4 * 64 * 2 * 4;  // [4 warps][64 ch][sum, sumsq]
  static constexpr int kSmemBars = 256;
  static constexpr int kSmemBias = 256;                // 64 fp32 bias values of the current output chunk
  // The small regions (statistics scratch, barriers, bias) sit FIRST and the 1 KiB-aligned stage / store buffers
  // follow at the next 1 KiB boundary: with the dynamic window starting 1 KiB-aligned (it does: the driver's reserved
  // 1 KiB precedes it) BN = 256 fits FOUR 48 KiB stages in exactly 227 KiB. The kernel traps if the carve-up would
  // run past the window.
  static constexpr int kSmemMisc = kSmemScratch + kSmemBars + kSmemBias;
  static constexpr int kSmemTotal = ((kSmemMisc + 1023) / 1024) * 1024 + kSmemStages + kSmemStore;
};

// Bounded spin so a protocol bug reports an error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_guard(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s: a protocol bug, not a slow tile
      if (err) atomicExch(err, code);
      __trap();
    }
  }
}

// ---- epilogue helpers (one warp per scheduler runs these: keep the instruction count low) ----------
// 64 fp32 accumulator columns -> 32 packed bf16x2 words, with the optional bias / activation hoisted
// out of the common (InstanceNorm follows: no bias, no activation) path.
template <int ACT, bool BIAS>
__device__ __forceinline__ void epi_pack_as(const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                            uint32_t (&packed)[32], float slope, const float* __restrict__ sbias) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x0 = __uint_as_float(j < 16 ? v0[2 * j] : v1[2 * (j - 16)]);
    float x1 = __uint_as_float(j < 16 ? v0[2 * j + 1] : v1[2 * (j - 16) + 1]);
    if (BIAS) {   // this chunk's 64 bias values, staged in shared memory (zero past bias_len): broadcast reads
      const float2 b = *reinterpret_cast<const float2*>(sbias + 2 * j);
      x0 += b.x;
      x1 += b.y;
    }
    if (ACT == ACT_LRELU) {
      x0 = x0 > 0.f ? x0 : x0 * slope;
      x1 = x1 > 0.f ? x1 : x1 * slope;
    } else if (ACT == ACT_RELU) {
      x0 = fmaxf(x0, 0.f);
      x1 = fmaxf(x1, 0.f);
    } else if (ACT == ACT_SIGMOID) {
      x0 = 1.f / (1.f + __expf(-x0));
      x1 = 1.f / (1.f + __expf(-x1));
    } else if (ACT == ACT_TANH) {
      x0 = tanhf(x0);
      x1 = tanhf(x1);
    }
    packed[j] = pack_bf16x2(x0, x1);
  }
}

// The activation switch sits OUTSIDE the 32-element loop: every variant is a tight straight-line block (with
// the switch inside, the unrolled body is a chain of branches over several KB of code per tile row).
__device__ __forceinline__ void epi_pack(const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                         uint32_t (&packed)[32], int act, float slope,
                                         const float* __restrict__ sbias) {
  if (sbias == nullptr) {
    switch (act) {
      case ACT_NONE: epi_pack_as<ACT_NONE, false>(v0, v1, packed, slope, sbias); break;
      case ACT_LRELU: epi_pack_as<ACT_LRELU, false>(v0, v1, packed, slope, sbias); break;
      case ACT_RELU: epi_pack_as<ACT_RELU, false>(v0, v1, packed, slope, sbias); break;
      case ACT_SIGMOID: epi_pack_as<ACT_SIGMOID, false>(v0, v1, packed, slope, sbias); break;
      default: epi_pack_as<ACT_TANH, false>(v0, v1, packed, slope, sbias); break;
    }
  } else {
    switch (act) {
      case ACT_NONE: epi_pack_as<ACT_NONE, true>(v0, v1, packed, slope, sbias); break;
      case ACT_LRELU: epi_pack_as<ACT_LRELU, true>(v0, v1, packed, slope, sbias); break;
      case ACT_RELU: epi_pack_as<ACT_RELU, true>(v0, v1, packed, slope, sbias); break;
      case ACT_SIGMOID: epi_pack_as<ACT_SIGMOID, true>(v0, v1, packed, slope, sbias); break;
      default: epi_pack_as<ACT_TANH, true>(v0, v1, packed, slope, sbias); break;
    }
  }
}

// 2x2 SUM over a 16x8-pixel tile held one pixel row per thread (tile row = ty*8 + tx = TMEM lane, 32 rows per
// warp = 4 image rows): the x neighbour is lane^1, the y neighbour lane^8. Every lane takes part; returns true on
// the 8 lanes per warp that end up holding a pooled pixel, `prow` = its row in the 8x4 pooled tile.
__device__ __forceinline__ bool epi_pool2x2(uint32_t (&v0)[32], uint32_t (&v1)[32], int q, int lane, int& prow) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float a = __uint_as_float(v0[j]), b = __uint_as_float(v1[j]);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 8);
    v0[j] = __float_as_uint(a);
    v1[j] = __float_as_uint(b);
  }
  prow = (2 * q + (lane >> 4)) * 4 + ((lane >> 1) & 3);
  return (lane & 9) == 0;
}

// Stage bias[c_base .. c_base+64) (zero past bias_len) for the 128 epilogue threads; no-op when already staged.
__device__ __forceinline__ void epi_stage_bias(float* sbias, const float* __restrict__ bias, int bias_len,
                                               int c_base, int& staged_base, int et, uint32_t bar_id = 1) {
  if (bias == nullptr || staged_base == c_base) return;
  named_bar_sync(bar_id, 128);   // readers of the previously staged chunk are done
  if (et < 64) sbias[et] = c_base + et < bias_len ? __ldg(bias + c_base + et) : 0.f;
  named_bar_sync(bar_id, 128);
  staged_base = c_base;
}

// this thread's row (128 B = 64 bf16) of the staging tile, 128B-swizzled like the TMA store expects
__device__ __forceinline__ void epi_store_row(uint32_t stage, int row, const uint32_t (&packed)[32]) {
  const uint32_t srow = stage + uint32_t(row) * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t dst = srow + (uint32_t(j ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * j]),
                 "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                 : "memory");
  }
}

// per-channel (sum, sum of squares) over rows [r0, r0+32) of the staged bf16 tile for channel pair `lane`.
// `valid_mask` bit i == row r0+i lies inside the image.
__device__ __forceinline__ void epi_stats_rows(uint32_t sbuf, int r0, int lane, uint32_t valid_mask,
                                               float& s0, float& s1, float& q0, float& q1) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    uint32_t w0, w1;
    const int ra = r0 + i, rb = r0 + i + 1;
    const uint32_t aa = sbuf + uint32_t(ra) * 128u + (uint32_t((lane >> 2) ^ (ra & 7)) << 4) + uint32_t(lane & 3) * 4u;
    const uint32_t ab = sbuf + uint32_t(rb) * 128u + (uint32_t((lane >> 2) ^ (rb & 7)) << 4) + uint32_t(lane & 3) * 4u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(aa));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(ab));
    if (!((valid_mask >> i) & 1u)) w0 = 0u;
    if (!((valid_mask >> (i + 1)) & 1u)) w1 = 0u;
    const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&w0);
    const __nv_bfloat162 h1 = *reinterpret_cast<const __nv_bfloat162*>(&w1);
    const float f0 = __low2float(h0), f1 = __high2float(h0), g0 = __low2float(h1), g1 = __high2float(h1);
    a0 += f0; a1 += f1; b0 = fmaf(f0, f0, b0); b1 = fmaf(f1, f1, b1);
    c0 += g0; c1 += g1; d0 = fmaf(g0, g0, d0); d1 = fmaf(g1, g1, d1);
  }
  s0 = a0 + c0; s1 = a1 + c1; q0 = b0 + d0; q1 = b1 + d1;
}

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 2)
igemm_conv_kernel(const __grid_constant__ IgemmParams p) {
  using Cfg = IgemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t scratch_base = smem_base;
  const uint32_t bar_base = scratch_base + Cfg::kSmemScratch;
  const uint32_t stage_base = (smem_base + uint32_t(Cfg::kSmemMisc) + 1023u) & ~1023u;
  const uint32_t store_base = stage_base + Cfg::kSmemStages;
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (store_base + Cfg::kSmemStore - smem_base > dyn) {
      if (threadIdx.x == 0 && p.err_flag) atomicExch(p.err_flag, 20);
      __trap();
    }
  }
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.num_src; ++s) {
      tma_prefetch_desc(&p.src[s].act);
      tma_prefetch_desc(&p.src[s].wgt);
    }
    tma_prefetch_desc(&p.out);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < Cfg::kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tfull_bar(s), 1);
        mbar_init(tempty_bar(s), 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_sync();   // everything above is set-up; global memory is touched from here on
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int taps = p.taps;
  int k_iters = 0;
  for (int s = 0; s < p.num_src; ++s) k_iters += taps * p.src[s].c_chunks;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int m_tiles = p.tiles_img * tiles_per_img;
  const int splits = p.splits > 1 ? p.splits : 1;
  const int total_tiles = m_tiles * p.n_tiles * splits;     // work items: (tile, split of the K loop)

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / splits, sp = item % splits;
        const int k_lo = k_iters * sp / splits, k_hi = k_iters * (sp + 1) / splits;
        int kcount = 0;
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int img = m_tile / tiles_per_img;
        const int t_in = m_tile % tiles_per_img;
        const int ho0 = (t_in / p.tiles_w) * p.th;
        const int wo0 = (t_in % p.tiles_w) * p.tw;
        const int n0 = img * p.tn;
        for (int s = 0; s < p.num_src; ++s) {
          const IgemmSrc& src = p.src[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int hi0 = ho0 * p.stride + p.tap_dy[tap];
            const int wi0 = wo0 * p.stride + p.tap_dx[tap];
            const int wtap = p.tap_w[tap];
            for (int cc = 0; cc < src.c_chunks; ++cc, ++kcount) {
              if (kcount < k_lo || kcount >= k_hi) continue;
              mbar_wait_guard(empty_bar(stage), phase ^ 1, p.err_flag, 1);
              const uint32_t a_dst = stage_base + stage * Cfg::kStageBytes;
              const uint32_t b_dst = a_dst + kABytes;
              mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_4d(a_dst, &src.act, full_bar(stage), cc * kChunkK, wi0, hi0, n0);
              tma_load_3d(b_dst, &src.wgt, full_bar(stage), cc * kChunkK, n_tile * BN, wtap);
              if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int sp = item % splits;
        const int n_k = k_iters * (sp + 1) / splits - k_iters * sp / splits;
        mbar_wait_guard(tempty_bar(as), aphase ^ 1, p.err_flag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int ki = 0; ki < n_k; ++ki) {
          mbar_wait_guard(full_bar(stage), phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t a_addr = stage_base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = umma_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(a_addr + kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k) {
            // +32 bytes along K inside the 128B swizzle row == +2 in the (addr >> 4) field
            umma_f16(d_tmem, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc,
                     (ki | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(as));
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (128 threads)
    const int et = threadIdx.x - 64;        // 0..127
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;          // tile row == TMEM lane
    const int ew = et >> 5;                 // 0..3 (row group for the stats pass)
    int as = 0;
    uint32_t aphase = 0;
    uint32_t chunk_ctr = 0;
    float* scratch = reinterpret_cast<float*>(smem_gen + (scratch_base - smem_base));
    float* sbias = reinterpret_cast<float*>(smem_gen + (bar_base + Cfg::kSmemBars - smem_base));
    int staged_base = -1;
    const int e_act = p.act, e_bias_len = p.bias_len;
    const float e_slope = p.slope;
    const float* e_bias = p.bias;
    float* e_stats = p.stats_partial;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
      const int tile = item / splits, sp = item % splits;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int t_in = m_tile % tiles_per_img;
      const int ho0 = (t_in / p.tiles_w) * p.th;
      const int wo0 = (t_in % p.tiles_w) * p.tw;
      const int n0 = img * p.tn;
      if (splits > 1) {
        // split-K: this item's fp32 partial tile goes to its own workspace slice (plain stores, no atomics: the
        // finalize kernel adds the slices in a fixed order); tile row -> (image, y, x) in the box order n, h, w
        const int iw = row % p.tw, ih = (row / p.tw) % p.th, in = row / (p.tw * p.th);
        const bool inside = n0 + in < p.N && ho0 + ih < p.Ho && wo0 + iw < p.Wo;
        float* dst = p.ws + (long long)sp * p.ws_slice +
                     ((long long)((n0 + in) * p.Ho + ho0 + ih) * p.Wo + wo0 + iw) * p.cout + n_tile * BN;
        mbar_wait_guard(tfull_bar(as), aphase, p.err_flag, 4);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 64; ++chunk) {
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + chunk * 64);
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(taddr, v0);
          tmem_ld_32x32(taddr + 32, v1);
          tmem_ld_wait();
          if (chunk == BN / 64 - 1) {
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
          }
          if (inside) {
            float4* d4 = reinterpret_cast<float4*>(dst + chunk * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              d4[j] = make_float4(__uint_as_float(v0[4 * j]), __uint_as_float(v0[4 * j + 1]),
                                  __uint_as_float(v0[4 * j + 2]), __uint_as_float(v0[4 * j + 3]));
              d4[8 + j] = make_float4(__uint_as_float(v1[4 * j]), __uint_as_float(v1[4 * j + 1]),
                                      __uint_as_float(v1[4 * j + 2]), __uint_as_float(v1[4 * j + 3]));
            }
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
        continue;
      }
      // which of this warp-group's 32 statistic rows lie inside the image (tiles may overhang)
      uint32_t valid_mask = 0xffffffffu;
      if (e_stats && (ho0 + p.th > p.Ho || wo0 + p.tw > p.Wo)) {
        valid_mask = 0u;
        for (int i = 0; i < 32; ++i) {
          const int r = ew * 32 + i;
          if (ho0 + r / p.tw < p.Ho && wo0 + r % p.tw < p.Wo) valid_mask |= 1u << i;
        }
      }
      mbar_wait_guard(tfull_bar(as), aphase, p.err_flag, 4);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 64; ++chunk, ++chunk_ctr) {
        const uint32_t sb = chunk_ctr & 1;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + chunk * 64);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        if (chunk == BN / 64 - 1) {
          tc_fence_before();
          mbar_arrive(tempty_bar(as));
        }
        const int c_base = n_tile * BN + chunk * 64;
        uint32_t packed[32];
        if (p.pool_out) {
          int prow;
          const bool keep = epi_pool2x2(v0, v1, q, lane, prow);
          epi_pack(v0, v1, packed, ACT_NONE, 0.f, nullptr);
          if (et == 0) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          if (keep) epi_store_row(store_base + sb * kStoreBytes, prow, packed);
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (et == 0) {
            tma_store_4d(&p.out, store_base + sb * kStoreBytes, c_base, wo0 >> 1, ho0 >> 1, n0);
            tma_store_commit();
          }
          continue;
        }
        epi_stage_bias(sbias, e_bias, e_bias_len, c_base, staged_base, et);
        epi_pack(v0, v1, packed, e_act, e_slope, e_bias ? sbias : nullptr);
        // staging buffer `sb` must have been drained by the TMA store issued two chunks ago
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        epi_store_row(store_base + sb * kStoreBytes, row, packed);
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (et == 0) {
          tma_store_4d(&p.out, store_base + sb * kStoreBytes, c_base, wo0, ho0, n0);
          tma_store_commit();
        }
        if (e_stats) {
          // column sums over the bf16 values actually stored (what the normalise pass will read)
          float s0, s1, q0, q1;
          epi_stats_rows(store_base + sb * kStoreBytes, ew * 32, lane, valid_mask, s0, s1, q0, q1);
          float* sc = scratch + ew * 128;
          sc[(2 * lane) * 2 + 0] = s0;
          sc[(2 * lane) * 2 + 1] = q0;
          sc[(2 * lane + 1) * 2 + 0] = s1;
          sc[(2 * lane + 1) * 2 + 1] = q1;
          named_bar_sync(1, 128);
          // et -> (channel = et >> 1, stat = et & 1)
          const float tot = scratch[et] + scratch[128 + et] + scratch[256 + et] + scratch[384 + et];
          const size_t tile_lin = size_t(n0) * p.stats_tiles_total + p.stats_tile_off + t_in;
          e_stats[(tile_lin * p.cout + c_base) * 2 + et] = tot;
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace tg
