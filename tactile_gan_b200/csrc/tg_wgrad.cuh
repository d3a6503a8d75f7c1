// Weight-gradient implicit GEMM on tcgen05 (sm_100a).
//
//   dw[tap][qc][pc] += sum_pixels  P[n, y*stride + dy(tap), x*stride + dx(tap), pc] * Q[n, y, x, qc]
//
// P is the shifted/strided operand (the conv input X for Conv2d, dY for ConvTranspose2d) and may be
// a virtual concat of several tensors; Q is the fixed operand (dY for Conv2d). The reduction (UMMA K)
// runs over pixels, so both operands are consumed **MN-major** straight from the NHWC layout: a TMA
// box of 64 pixels x 64 channels lands as 64 rows of 128 B (128B swizzle), rows = K, bytes = M/N.
// UMMA M = 128 P-channels (two boxes), UMMA N = BN Q-channels (BN/64 boxes), K = 64 pixels / stage.
//
// Work item = (tap, m_tile, n_tile, k_split); partial sums are reduced into fp32 dw with red.add.
#pragma once
#include "tg_igemm.cuh"

namespace tg {

constexpr int kWgBoxPix = 64;
constexpr int kWgBoxBytes = kWgBoxPix * 128;  // 8 KiB: 64 pixel rows x 64 channels bf16

struct alignas(64) WgradSrc {
  CUtensorMap act;  // {C, W, H, N}, box {64, tw*stride, th*stride, tn}, elementStrides {1,s,s,1}
  int c_chunks;
  int pad_[15];
};

struct alignas(64) WgradParams {
  WgradSrc src[kMaxSrc];
  CUtensorMap q;  // {Cq, Wq, Hq, N}, box {64, tw, th, tn}
  int num_src;
  int taps, stride;
  int8_t tap_dy[16], tap_dx[16], tap_w[16];
  int th, tw, tn;
  int tiles_h, tiles_w, tiles_img;  // pixel blocks of 64
  int m_tiles, total_chunks;        // P side: 128-channel tiles, 64-channel chunks over all sources
  int n_tiles;                      // Cq / BN
  int splits;
  float* dw;                        // [taps_total][n_total][m_total] fp32
  int m_total, n_total;
  int* err_flag;
};

template <int BN>
struct WgradCfg {
  static constexpr int kABytes = 2 * kWgBoxBytes;
  static constexpr int kBBytes = (BN / 64) * kWgBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
#ifndef TG_WG256_STAGES
#define TG_WG256_STAGES 4
#endif
  static constexpr int kStages = BN == 256 ? TG_WG256_STAGES : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemTotal = 1024 + kStages * kStageBytes + 256;
};

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 2)
wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t bar_base = stage_base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.src[s].act);
    tma_prefetch_desc(&p.q);
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

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int k_blocks = p.tiles_img * tiles_per_img;
  const int kb_per_split = (k_blocks + p.splits - 1) / p.splits;
  const int total_items = p.taps * p.m_tiles * p.n_tiles * p.splits;

  // item -> (tap fastest, then n_tile, m_tile; pixel split slowest): CTAs running at the same time work on
  // the same pixel range, so the (shifted) P windows and Q blocks they stream are shared through L2 and the
  // operands cross HBM once even when they are larger than L2
  auto decode = [&](int item, int& tap, int& m_tile, int& n_tile, int& kb0, int& kb1) {
    int r = item;
    tap = r % p.taps; r /= p.taps;
    n_tile = r % p.n_tiles; r /= p.n_tiles;
    m_tile = r % p.m_tiles; r /= p.m_tiles;
    const int split = r;
    kb0 = split * kb_per_split;
    kb1 = min(k_blocks, kb0 + kb_per_split);
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int tap, m_tile, n_tile, kb0, kb1;
        decode(item, tap, m_tile, n_tile, kb0, kb1);
        // resolve the two P chunks of this m_tile to (source, channel offset)
        int csrc[2], coff[2];
        for (int h = 0; h < 2; ++h) {
          int j = m_tile * 2 + h;
          csrc[h] = -1; coff[h] = 0;
          if (j < p.total_chunks) {
            for (int s = 0; s < p.num_src; ++s) {
              if (j < p.src[s].c_chunks) { csrc[h] = s; coff[h] = j * 64; break; }
              j -= p.src[s].c_chunks;
            }
          }
        }
        const uint32_t tx_bytes = (csrc[1] >= 0 ? 2 : 1) * kWgBoxBytes + Cfg::kBBytes;
        const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
        for (int kb = kb0; kb < kb1; ++kb) {
          const int img = kb / tiles_per_img;
          const int t_in = kb % tiles_per_img;
          const int y0 = (t_in / p.tiles_w) * p.th;
          const int x0 = (t_in % p.tiles_w) * p.tw;
          const int n0 = img * p.tn;
          mbar_wait_guard(empty_bar(stage), phase ^ 1, p.err_flag, 11);
          const uint32_t a_dst = stage_base + stage * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + Cfg::kABytes;
          mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          for (int h = 0; h < 2; ++h)
            if (csrc[h] >= 0)
              tma_load_4d(a_dst + h * kWgBoxBytes, &p.src[csrc[h]].act, full_bar(stage), coff[h],
                          x0 * p.stride + dx, y0 * p.stride + dy, n0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(b_dst + j * kWgBoxBytes, &p.q, full_bar(stage), n_tile * BN + j * 64, x0, y0,
                        n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int tap, m_tile, n_tile, kb0, kb1;
        decode(item, tap, m_tile, n_tile, kb0, kb1);
        mbar_wait_guard(tempty_bar(as), aphase ^ 1, p.err_flag, 12);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_guard(full_bar(stage), phase, p.err_flag, 13);
          tc_fence_after();
          const uint32_t a_addr = stage_base + stage * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kWgBoxPix / 16; ++k) {
            // 16 pixels (K) = 16 rows of 128 B = 2 KiB further into every 64-channel box
            const uint64_t a_desc = umma_smem_desc_sw128(a_addr + k * 2048, kWgBoxBytes, 1024);
            const uint64_t b_desc = umma_smem_desc_sw128(b_addr + k * 2048, kWgBoxBytes, 1024);
            umma_f16(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
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
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int tap, m_tile, n_tile, kb0, kb1;
      decode(item, tap, m_tile, n_tile, kb0, kb1);
      mbar_wait_guard(tfull_bar(as), aphase, p.err_flag, 14);
      tc_fence_after();
      const int pc = m_tile * 128 + row;
      const bool row_ok = pc < p.m_total && kb1 > kb0;
      float* dst = p.dw + (size_t(p.tap_w[tap]) * p.n_total + size_t(n_tile) * BN) * p.m_total + pc;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + c0), v);
        tmem_ld_wait();
        if (c0 + 32 >= BN) {
          tc_fence_before();
          mbar_arrive(tempty_bar(as));
        }
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            atomicAdd(dst + size_t(c0 + j) * p.m_total, __uint_as_float(v[j]));
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace tg
