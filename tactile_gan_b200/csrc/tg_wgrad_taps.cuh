// Tap-tiled weight gradient for 3x3-window stride-1 convolutions, padded or valid (sm_100a, tcgen05).
//
//   dW[(dy,dx)][co][ci] = sum_{r,c} X[r+dy, c+dx][ci] * dY[r, c][co]
//                       = sum_{r,c'} X[r+dy, c'][ci] * dY[r, c'-dx][co]          (c' = c + dx)
//
// so the row shift can live on X and the column shift on dY. Per 16x8 pixel tile ONE box of X with a row
// halo (18 x 8 px x 64 ch, 18 KiB) and ONE box of dY with a column halo (16 x 10 px x 64 ch, 20 KiB) feed
//   UMMA  M = 128 = {dy_a, dy_b} x 64 ci   (A = X,  MN-major, second M group one image row = 1024 B further)
//         N = 192 = {dx=+1,0,-1} x 64 co   (B = dY, MN-major, next N group one pixel = 128 B further)
//         K = 16 pixels (two tile rows) per instruction.
// All nine taps of a 64-channel input chunk against 64 output channels take 16 instructions per tile
// (accumulator 0: dy = -1,0; accumulator 1: dy = +1, upper half idle) and 10 KiB of operand reads per 96
// tensor cycles -- under the 128 B/cycle shared-memory port that bounds the N = 64 formulation (6 KiB per
// 32 cycles). The shifted / overlapping operand groups rely on the 128B swizzle being a function of the
// absolute shared-memory address (profiles/r01_umma_offset_probe.log).
// Work item = (input chunk, 64 output channels, pixel split); fp32 partial sums land with red.add.
#pragma once
#include "tg_wgrad.cuh"

namespace tg {

constexpr int kWtTH = 16, kWtTW = 8;
constexpr int kWtXBytes = (kWtTH + 2) * kWtTW * 128;   // 18 KiB
constexpr int kWtYBytes = kWtTH * (kWtTW + 2) * 128;   // 20 KiB
constexpr int kWtStageBytes = kWtXBytes + kWtYBytes;   // 38 KiB, both parts 1 KiB aligned
#ifndef TG_WT_STAGES
#define TG_WT_STAGES 5
#endif
constexpr int kWtStages = TG_WT_STAGES;
constexpr int kWtSmem = 1024 + kWtStages * kWtStageBytes + 256;

struct alignas(64) WgradTapsParams {
  WgradSrc src[kMaxSrc];  // X sources: box {64, 8, 18, 1}
  CUtensorMap q;          // dY: box {64, 10, 16, 1}
  int num_src;
  int8_t tap_w[3][3];     // [dy+1][dx+1] -> index on dw's tap axis, or -1 if the tap is absent
  int tiles_h, tiles_w, N;
  int org_y, org_x;       // X pixel of tap (dy, dx) = (0, 0) relative to the dY pixel (0 for pad = 1, 1 for pad = 0)
  int total_chunks, n_tiles, splits;
  float* dw;              // [taps_total][n_total][m_total] fp32
  int m_total, n_total;
  int* err_flag;
};

__global__ void __launch_bounds__(kNumThreads, 2) wgrad_taps_kernel(const __grid_constant__ WgradTapsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kWtStages * kWtStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWtStages + s); };
  const uint32_t tfull = bar_base + 8u * (2 * kWtStages);
  const uint32_t tempty = bar_base + 8u * (2 * kWtStages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWtStages + 2);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.src[s].act);
    tma_prefetch_desc(&p.q);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < kWtStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      mbar_init(tfull, 1);
      mbar_init(tempty, 128);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_sync();   // everything above is set-up; global memory is touched from here on
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int k_tiles = p.N * tiles_per_img;
  const int per_split = (k_tiles + p.splits - 1) / p.splits;
  const int total_items = p.total_chunks * p.n_tiles * p.splits;
  // chunk fastest, pixel split slowest: CTAs running together stream the same pixel range (dY shared by the
  // chunks, X shared by the output-channel tiles) through L2
  auto decode = [&](int item, int& chunk, int& n_tile, int& k0, int& k1) {
    int r = item;
    chunk = r % p.total_chunks; r /= p.total_chunks;
    n_tile = r % p.n_tiles; r /= p.n_tiles;
    k0 = r * per_split;
    k1 = min(k_tiles, k0 + per_split);
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int chunk, n_tile, k0, k1;
        decode(item, chunk, n_tile, k0, k1);
        int s = 0, j = chunk;
        while (j >= p.src[s].c_chunks) { j -= p.src[s].c_chunks; ++s; }
        const CUtensorMap* xmap = &p.src[s].act;
        const int xc = j * 64;
        for (int kt = k0; kt < k1; ++kt) {
          const int img = kt / tiles_per_img, t_in = kt % tiles_per_img;
          // tile columns run over c' = c + dx in [-org_x, W + org_x): for a valid conv the shifted column can
          // fall one pixel outside the dY grid on either side while still addressing a real X column
          const int y0 = (t_in / p.tiles_w) * kWtTH, x0 = (t_in % p.tiles_w) * kWtTW - p.org_x;
          mbar_wait_guard(empty_bar(stage), phase ^ 1, p.err_flag, 41);
          const uint32_t xs = smem_base + stage * kWtStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), uint32_t(kWtStageBytes));
          tma_load_4d(xs, xmap, full_bar(stage), xc, x0 + p.org_x, y0 - 1 + p.org_y, img);
          tma_load_4d(xs + kWtXBytes, &p.q, full_bar(stage), n_tile * 64, x0 - 1, y0, img);
          if (++stage == kWtStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, 192, 1, 1);
      const uint32_t a_hi = umma_desc_hi_sw128(1024u);                  // K groups: one image row of X
      const uint32_t b_hi = umma_desc_hi_sw128((kWtTW + 2) * 128u);     // K groups: one image row of dY (10 px)
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int chunk, n_tile, k0, k1;
        decode(item, chunk, n_tile, k0, k1);
        mbar_wait_guard(tempty, tphase ^ 1, p.err_flag, 42);
        tc_fence_after();
        for (int kt = k0; kt < k1; ++kt) {
          mbar_wait_guard(full_bar(stage), phase, p.err_flag, 43);
          tc_fence_after();
          const uint32_t xs = smem_base + stage * kWtStageBytes;
          // A: rows of the X box are 1 KiB; accumulator 0 starts at box row 0 (dy = -1) with the second M group
          //    one row further (dy = 0); accumulator 1 starts at box row 2 (dy = +1), second group unused
          const uint32_t a0 = umma_desc_lo(xs, 1024), a1 = umma_desc_lo(xs + 2048, 0);
          // B: N groups dx = +1, 0, -1 start at box pixel 0, 1, 2 of the row (128 B apart)
          const uint32_t b0 = umma_desc_lo(xs + kWtXBytes, 128);
          const uint32_t acc = kt > k0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t ak = uint32_t(k) * (2048u >> 4), bk = uint32_t(k) * ((2u * (kWtTW + 2) * 128u) >> 4);
            umma_f16_split(tmem_base, a0 + ak, a_hi, b0 + bk, b_hi, idesc, k > 0 ? 1u : acc);
            umma_f16_split(tmem_base + 192u, a1 + ak, a_hi, b0 + bk, b_hi, idesc, k > 0 ? 1u : acc);
          }
          umma_commit(empty_bar(stage));
          if (++stage == kWtStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull);
        tphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int chunk, n_tile, k0, k1;
      decode(item, chunk, n_tile, k0, k1);
      mbar_wait_guard(tfull, tphase, p.err_flag, 44);
      tphase ^= 1;
      tc_fence_after();
      const int ci = chunk * 64 + (row & 63);
      if (k1 > k0) {
#pragma unroll 1
        for (int a = 0; a < 2; ++a) {
          const int dyi = a == 0 ? (row >> 6) : 2;   // index dy + 1
          if (a == 1 && row >= 64) break;
#pragma unroll 1
          for (int g = 0; g < 3; ++g) {
            const int tw = p.tap_w[dyi][2 - g];        // N group g <-> dx = 1 - g
            if (tw < 0) continue;
            float* dst = p.dw + (size_t(tw) * p.n_total + size_t(n_tile) * 64) * p.m_total + ci;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t v[32];
              tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(a * 192 + g * 64 + c0), v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(dst + size_t(c0 + j) * p.m_total, __uint_as_float(v[j]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tg
