// Halo-resident weight gradient for 3x3-window stride-1 convolutions with 64 output channels per tile
// (the Cout = 64 / 128 layers, where the per-tap wgrad kernel is L2- and issue-bound).
//
//   dW[tap][co][ci] = sum_p X[p + off(tap)][ci] * dY[p][co] = sum_{p' in tile} X[p'][ci] * dY[p' - off(tap)][co]
//
// Per pixel tile (16x8): ONE 32 KiB load of X (two 64-channel boxes, UMMA A, MN-major, M = 128 input
// channels) and ONE 23 KiB halo box of dY (UMMA B, MN-major, N = 64 output channels); the tap shift is a
// descriptor start offset into the dY halo tile (rows = pixels = K; SBO = 10 rows). A work item owns up to
// 5 taps (5 x 64 TMEM columns), one 128-channel slice of the (virtual-concat) input and a range of pixel
// tiles; partial sums are reduced into the fp32 gradient with red.add.
#pragma once
#include "tg_wgrad.cuh"
#include "tg_igemm_halo.cuh"

namespace tg {

constexpr int kWhTaps = 5;                               // taps per work item
constexpr int kWhXBytes = 2 * kTileM * 128;              // 32 KiB: 128 pixels x (2 x 64 ch)
constexpr int kWhYStage = 23 * 1024;                     // >= 18*10*128
constexpr int kWhStageBytes = kWhXBytes + kWhYStage;
constexpr int kWhStages = 4;
constexpr int kWhSmem = 1024 + kWhStages * kWhStageBytes + 256;

struct alignas(64) WgradHaloParams {
  WgradSrc src[kMaxSrc];  // X sources: box {64, 8, 16, 1}
  CUtensorMap q;          // dY: box {64, 8+ew, 16+eh, 1}
  int num_src;
  int taps;
  int8_t tap_dy[16], tap_dx[16], tap_w[16];  // halo-relative row / col of each tap's dY window
  int org_dy, org_dx, halo_w, y_bytes;
  int tiles_h, tiles_w, N;
  int m_tiles, total_chunks, n_tiles, tap_groups, splits;
  float* dw;
  int m_total, n_total;
  int* err_flag;
};

__global__ void __launch_bounds__(kNumThreads, 2) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kWhStages * kWhStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWhStages + s); };
  const uint32_t tfull = bar_base + 8u * (2 * kWhStages);
  const uint32_t tempty = bar_base + 8u * (2 * kWhStages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWhStages + 2);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.src[s].act);
    tma_prefetch_desc(&p.q);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < kWhStages; ++s) {
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
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int k_tiles = p.N * tiles_per_img;
  const int per_split = (k_tiles + p.splits - 1) / p.splits;
  const int total_items = p.tap_groups * p.m_tiles * p.n_tiles * p.splits;
  auto decode = [&](int item, int& tg0, int& ntap, int& m_tile, int& n_tile, int& k0, int& k1) {
    // tap group fastest, pixel split slowest: concurrent CTAs share the X / dY tiles through L2
    int r = item;
    tg0 = (r % p.tap_groups) * kWhTaps; r /= p.tap_groups;
    n_tile = r % p.n_tiles; r /= p.n_tiles;
    m_tile = r % p.m_tiles; r /= p.m_tiles;
    const int split = r;
    ntap = min(kWhTaps, p.taps - tg0);
    k0 = split * per_split;
    k1 = min(k_tiles, k0 + per_split);
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int tg0, ntap, m_tile, n_tile, k0, k1;
        decode(item, tg0, ntap, m_tile, n_tile, k0, k1);
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
        const uint32_t tx = (csrc[1] >= 0 ? 2u : 1u) * (kTileM * 128u) + uint32_t(p.y_bytes);
        for (int kt = k0; kt < k1; ++kt) {
          const int img = kt / tiles_per_img, t_in = kt % tiles_per_img;
          const int y0 = (t_in / p.tiles_w) * kHaloTH, x0 = (t_in % p.tiles_w) * kHaloTW;
          mbar_wait_guard(empty_bar(stage), phase ^ 1, p.err_flag, 31);
          const uint32_t xs = smem_base + stage * kWhStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), tx);
          for (int h = 0; h < 2; ++h)
            if (csrc[h] >= 0)
              tma_load_4d(xs + h * (kTileM * 128), &p.src[csrc[h]].act, full_bar(stage), coff[h], x0, y0, img);
          tma_load_4d(xs + kWhXBytes, &p.q, full_bar(stage), n_tile * 64, x0 + p.org_dx, y0 + p.org_dy, img);
          if (++stage == kWhStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, 64, 1, 1);
      const uint32_t sbo_y = uint32_t(p.halo_w) * 128u;
      // constant descriptor halves and per-tap start steps (16-byte units) hoisted out of the issue loop
      const uint32_t a_hi = umma_desc_hi_sw128(1024u);
      const uint32_t b_hi = umma_desc_hi_sw128(sbo_y);
      const uint32_t b_kstep = 2u * (sbo_y >> 4);          // 16 pixels (K) = two halo rows
      uint32_t y_step[9];
#pragma unroll
      for (int t = 0; t < 9; ++t)
        y_step[t] = t < p.taps ? uint32_t(p.tap_dy[t] * p.halo_w + p.tap_dx[t]) * 8u : 0u;
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int tg0, ntap, m_tile, n_tile, k0, k1;
        decode(item, tg0, ntap, m_tile, n_tile, k0, k1);
        // this item's taps -> registers (tg0 is 0 or kWhTaps)
        uint32_t ys[kWhTaps];
#pragma unroll
        for (int t = 0; t < kWhTaps; ++t) ys[t] = tg0 == 0 ? y_step[t] : (t + kWhTaps < 9 ? y_step[t + kWhTaps] : 0u);
        mbar_wait_guard(tempty, tphase ^ 1, p.err_flag, 32);
        tc_fence_after();
        for (int kt = k0; kt < k1; ++kt) {
          mbar_wait_guard(full_bar(stage), phase, p.err_flag, 33);
          tc_fence_after();
          const uint32_t xs = smem_base + stage * kWhStageBytes;
          // K step = 16 pixels = two image rows of the tile: X rows are dense (8 px = 1 KiB, +2 KiB per step),
          // dY rows sit in the halo tile with a pitch of halo_w pixels
          const uint32_t a_lo0 = umma_desc_lo(xs, kTileM * 128);
          const uint32_t b_lo0 = umma_desc_lo(xs + kWhXBytes, 16);
          const uint32_t first = kt > k0 ? 1u : 0u;
#pragma unroll
          for (int t = 0; t < kWhTaps; ++t) {
            if (t < ntap) {
              const uint32_t d_tmem = tmem_base + uint32_t(t * 64);
              const uint32_t b_lo = b_lo0 + ys[t];
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_f16_split(d_tmem, a_lo0 + k * 128, a_hi, b_lo + k * b_kstep, b_hi, idesc, k > 0 ? 1u : first);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == kWhStages) { stage = 0; phase ^= 1; }
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
      int tg0, ntap, m_tile, n_tile, k0, k1;
      decode(item, tg0, ntap, m_tile, n_tile, k0, k1);
      mbar_wait_guard(tfull, tphase, p.err_flag, 34);
      tphase ^= 1;
      tc_fence_after();
      const int pc = m_tile * 128 + row;
      const bool row_ok = pc < p.m_total && k1 > k0;
      for (int t = 0; t < ntap; ++t) {
        float* dst = p.dw + (size_t(p.tap_w[tg0 + t]) * p.n_total + size_t(n_tile) * 64) * p.m_total + pc;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(t * 64 + c0), v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + size_t(c0 + j) * p.m_total, __uint_as_float(v[j]));
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
