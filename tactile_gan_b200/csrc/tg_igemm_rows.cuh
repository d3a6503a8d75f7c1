// Row-resident implicit GEMM for 3x3 stride-1 same-size convolutions with a 64-wide output tile on maps whose
// width is a multiple of 128 (the 256^2 / 128^2 levels of UNet++ and BCDUNet at Cout 64 / 128).
//
// The halo kernel (tg_igemm_halo.cuh) issues one M128 x N64 x K16 UMMA per (tap, k-step): 4 KiB of pixels + 2 KiB
// of weights per 32 tensor cycles = 192 B/cycle against a 128 B/cycle shared-memory port. Here a tile is 128
// consecutive pixels of ONE image row, and an input row is multiplied against the weights of all three vertical
// taps at once: N = 192 = {dy = 2, 1, 0} x 64 co. Input row y' feeds output rows y'-1, y', y'+1, whose accumulators
// sit side by side in TMEM (64 columns each), so the three 64-column slices of the product land directly in the
// accumulators of three different output rows -- the vertical scatter costs nothing, and no epilogue shuffle is
// needed. The horizontal tap is a start-address offset of the pixel descriptor into a 130-pixel row (128B swizzle
// is a function of the absolute address, as in the halo kernel).
//
//   item = 4 output rows x 128 columns x 64 output channels; per 64-channel input chunk:
//     TMA: 6 input rows (y0-1 .. y0+4) of 130 pixels (16.25 KiB each), 3 weight groups [dx][dy=2|1|0][64 co][64 ci]
//     for dx: for input row j: 4 k-steps of ONE UMMA with N = 64/128/192/192/128/64 (rows at the strip's edge feed
//     fewer output rows): 72 instructions and 576 KiB of operand reads per item-chunk instead of 144 and 864 KiB.
//
// The first touch of an accumulator (its dy = 0 slice at chunk 0, dx 0, k 0) is split off as an N = 64 instruction
// that overwrites; everything else accumulates. Weights stay resident when the layer has a single input chunk.
#pragma once
#include "tg_igemm.cuh"

namespace tg {

constexpr int kRowsG = 4;                    // output rows (128-pixel tiles) per item
constexpr int kRowsIn = kRowsG + 2;          // input rows per item-chunk
constexpr int kRowsPix = 130;                // 128 + one halo pixel on each side
constexpr int kRowsAStage = 17 * 1024;       // >= 130 * 128 B, 1 KiB aligned
constexpr int kRowsABytes = kRowsPix * 128;
constexpr int kRowsBBlock = 64 * 128;        // one tap: 64 co x 64 ci
constexpr int kRowsBGroup = 3 * kRowsBBlock; // one dx: [dy=2 | dy=1 | dy=0]
constexpr int kRowsThreads = 64 + 2 * 128;   // TMA warp, MMA warp, two epilogue groups of four warps
constexpr int kRowsSmem = 1024 + 3 * kRowsBGroup + kRowsIn * kRowsAStage + 2 * kStoreBytes + 2 * (4 * 64 * 2 * 4) + 256 + 512;

struct alignas(64) RowsParams {
  IgemmSrc src[kMaxSrc];  // act box {64, 130, 1, 1}; wgt box {64, 64, 1}
  CUtensorMap out;        // box {64, 128, 1, 1}
  int num_src;
  int8_t tap_w[3][4];     // [dy][dx] -> index on the weight tensor's tap axis
  int org_dy, org_dx;     // input offset of tap (0, 0) relative to the output pixel
  int Ho, Wo, N;
  int segs, groups_h;     // Wo / 128, Ho / 4
  int n_tiles;            // Cout / 64
  int act;
  float slope;
  const float* bias;
  int bias_len;
  float* stats_partial;
  int stats_tiles_total, stats_tile_off;
  int cout;
  int prefetch;
  int pool_out;           // epilogue stores the 2x2 sum (`out` map = half-resolution tensor, box {64, 64, 1, 1})
  int* err_flag;
};

__global__ void __launch_bounds__(kRowsThreads, 1) igemm_rows_kernel(const __grid_constant__ RowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base;
  const uint32_t a_base = b_base + 3 * kRowsBGroup;
  const uint32_t store_base = a_base + kRowsIn * kRowsAStage;
  const uint32_t scratch_base = store_base + 2 * kStoreBytes;
  const uint32_t bar_base = scratch_base + 2 * (4 * 64 * 2 * 4);
  auto bfull = [&](int s) { return bar_base + 8u * s; };              // 3
  auto bempty = [&](int s) { return bar_base + 8u * (3 + s); };       // 3
  auto afull = [&](int s) { return bar_base + 8u * (6 + s); };        // 6
  auto aempty = [&](int s) { return bar_base + 8u * (12 + s); };      // 6
  auto tfull = [&](int s) { return bar_base + 8u * (18 + s); };       // 2
  auto tempty = [&](int s) { return bar_base + 8u * (20 + s); };      // 2
  const uint32_t tmem_slot = bar_base + 8u * 22;
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
      for (int s = 0; s < 3; ++s) {
        mbar_init(bfull(s), 1);
        mbar_init(bempty(s), 1);
      }
      for (int s = 0; s < kRowsIn; ++s) {
        mbar_init(afull(s), 1);
        mbar_init(aempty(s), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tfull(s), 1);
        mbar_init(tempty(s), 256);
      }
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

  int chunks = 0;
  for (int s = 0; s < p.num_src; ++s) chunks += p.src[s].c_chunks;
  const int groups = p.N * p.groups_h * p.segs;
  const int total_items = groups * p.n_tiles;
  // single-chunk layers keep their 72 KiB of weights for the whole kernel (the CTA's n_tile never changes)
  const bool resident = chunks == 1 && (int(gridDim.x) % p.n_tiles) == 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int n_tile = item % p.n_tiles;
        const int g = item / p.n_tiles;
        const int seg = g % p.segs, yg = (g / p.segs) % p.groups_h, img = g / (p.segs * p.groups_h);
        const int x0 = seg * kTileM + p.org_dx, y0 = yg * kRowsG + p.org_dy;
        for (int s = 0; s < p.num_src; ++s) {
          const IgemmSrc& src = p.src[s];
          for (int cc = 0; cc < src.c_chunks; ++cc, ++it) {
            const uint32_t ph = it & 1u;
            auto load_b = [&](int d) {
              if (resident && it != 0) return;
              mbar_wait_guard(bempty(d), ph ^ 1u, p.err_flag, 31);
              mbar_arrive_expect_tx(bfull(d), uint32_t(kRowsBGroup));
#pragma unroll
              for (int b = 0; b < 3; ++b)
                tma_load_3d(b_base + d * kRowsBGroup + b * kRowsBBlock, &src.wgt, bfull(d), cc * kChunkK,
                            n_tile * 64, p.tap_w[2 - b][d]);
            };
            // same order as the issuer consumes: weights of dx 0, the six rows, then the other two weight groups
            load_b(0);
            for (int j = 0; j < kRowsIn; ++j) {
              mbar_wait_guard(aempty(j), ph ^ 1u, p.err_flag, 32);
              mbar_arrive_expect_tx(afull(j), uint32_t(kRowsABytes));
              tma_load_4d(a_base + j * kRowsAStage, &src.act, afull(j), cc * kChunkK, x0, y0 + j, img);
            }
            load_b(1);
            load_b(2);
            if (p.prefetch) {
              // next chunk's rows (next source / next item at the end) into L2 while this chunk computes
              int ps = s, pcc = cc + 1, px0 = x0, py0 = y0, pimg = img;
              bool ok = true;
              if (pcc == src.c_chunks) { pcc = 0; ++ps; }
              if (ps == p.num_src) {
                ps = 0;
                const int nitem = item + int(gridDim.x);
                ok = nitem < total_items;
                const int ng = nitem / p.n_tiles;
                px0 = (ng % p.segs) * kTileM + p.org_dx;
                py0 = ((ng / p.segs) % p.groups_h) * kRowsG + p.org_dy;
                pimg = ng / (p.segs * p.groups_h);
              }
              if (ok)
                for (int j = 0; j < kRowsIn; ++j) tma_prefetch_4d(&p.src[ps].act, pcc * kChunkK, px0, py0 + j, pimg);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc64 = umma_idesc_bf16(kTileM, 64, 0, 0);
      constexpr uint32_t idesc128 = umma_idesc_bf16(kTileM, 128, 0, 0);
      constexpr uint32_t idesc192 = umma_idesc_bf16(kTileM, 192, 0, 0);
      const uint32_t hi = umma_desc_hi_sw128(1024u);
      uint32_t it = 0, tphase = 0;
      int set = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        mbar_wait_guard(tempty(set), tphase ^ 1u, p.err_flag, 33);
        tc_fence_after();
        for (int c = 0; c < chunks; ++c, ++it) {
          const uint32_t ph = it & 1u;
#pragma unroll 1
          for (int d = 0; d < 3; ++d) {
            if (!resident || it == 0) mbar_wait_guard(bfull(d), ph, p.err_flag, 34);
            tc_fence_after();
            const uint32_t b_lo0 = umma_desc_lo(b_base + d * kRowsBGroup, 16);
#pragma unroll
            for (int j = 0; j < kRowsIn; ++j) {
              if (d == 0) {
                mbar_wait_guard(afull(j), ph, p.err_flag, 35);
                tc_fence_after();
              }
              // input row j feeds output rows r = j - dy, dy in [max(0, j-3), min(2, j)]; columns ascend with r,
              // i.e. with descending dy = the block order of the weight group
              constexpr int kLo[kRowsIn] = {0, 0, 0, 1, 2, 3};      // first output row
              constexpr int kNb[kRowsIn] = {1, 2, 3, 3, 2, 1};      // number of 64-column blocks
              constexpr int kB0[kRowsIn] = {2, 1, 0, 0, 0, 0};      // first weight block (2 - dy_max)
              const int nb = kNb[j];
              const uint32_t idesc = nb == 3 ? idesc192 : (nb == 2 ? idesc128 : idesc64);
              const uint32_t a_lo = umma_desc_lo(a_base + j * kRowsAStage, 16) + uint32_t(d) * 8u;
              const uint32_t b_lo = b_lo0 + uint32_t(kB0[j]) * (kRowsBBlock >> 4);
              const uint32_t d_tmem = tmem_base + uint32_t((set * kRowsG + kLo[j]) * 64);
              const bool first = c == 0 && d == 0;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k == 0 && first && j <= 3) {
                  // the last block (dy = 0 -> output row j) is that accumulator's first contribution: overwrite
                  if (nb > 1)
                    umma_f16_split(d_tmem, a_lo, hi, b_lo, hi, nb == 3 ? idesc128 : idesc64, 1u);
                  umma_f16_split(d_tmem + uint32_t((nb - 1) * 64), a_lo, hi, b_lo + uint32_t(nb - 1) * (kRowsBBlock >> 4),
                                 hi, idesc64, 0u);
                } else {
                  umma_f16_split(d_tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, 1u);
                }
              }
              if (d == 2) umma_commit(aempty(j));
            }
            if (!resident) umma_commit(bempty(d));
          }
        }
        umma_commit(tfull(set));
        set ^= 1;
        if (set == 0) tphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: two groups of 128 threads
    // One epilogue warp per scheduler is latency-bound (~1800 cycles per tile against 1152 tensor cycles per
    // tile-chunk), which capped the one- and two-chunk layers; group e drains output rows e and e+2 of every item
    // with its own staging tile, barrier, statistics scratch and bulk-store group.
    const int e = (warp - 2) >> 2;
    const int et = threadIdx.x - 64 - e * 128;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ew = et >> 5;
    const uint32_t bar_id = 1u + uint32_t(e);
    const uint32_t stage = store_base + uint32_t(e) * kStoreBytes;
    int set = 0;
    uint32_t tphase = 0;
    float* scratch = reinterpret_cast<float*>(smem_gen + (scratch_base - smem_base)) + e * 512;
    float* sbias = reinterpret_cast<float*>(smem_gen + (bar_base + 256 - smem_base)) + e * 64;
    int staged_base = -1;
    const int e_act = p.act, e_bias_len = p.bias_len;
    const float e_slope = p.slope;
    const float* e_bias = p.bias;
    float* e_stats = p.stats_partial;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int n_tile = item % p.n_tiles;
      const int g = item / p.n_tiles;
      const int seg = g % p.segs, yg = (g / p.segs) % p.groups_h, img = g / (p.segs * p.groups_h);
      const int c_base = n_tile * 64;
      mbar_wait_guard(tfull(set), tphase, p.err_flag, 36);
      tc_fence_after();
      if (p.pool_out) {
        // input gradient through a nearest-upsampled copy: rows (2e, 2e+1) are two accumulators on the SAME lanes,
        // so the vertical pair is a register add and the horizontal pair one lane shuffle; even lanes keep the pixel
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t((set * kRowsG + 2 * e) * 64);
        uint32_t v0[32], v1[32], t[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 64, t);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v0[j] = __float_as_uint(__uint_as_float(v0[j]) + __uint_as_float(t[j]));
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_32x32(taddr + 96, t);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tempty(set));
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(v0[j]), b = __uint_as_float(v1[j]) + __uint_as_float(t[j]);
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          b += __shfl_xor_sync(0xffffffffu, b, 1);
          v0[j] = __float_as_uint(a);
          v1[j] = __float_as_uint(b);
        }
        uint32_t packed[32];
        epi_pack(v0, v1, packed, ACT_NONE, 0.f, nullptr);
        if (et == 0) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        if ((lane & 1) == 0) epi_store_row(stage, row >> 1, packed);
        fence_proxy_async();
        named_bar_sync(bar_id, 128);
        if (et == 0) {
          tma_store_4d(&p.out, stage, c_base, seg * (kTileM / 2), yg * (kRowsG / 2) + e, img);
          tma_store_commit();
        }
        set ^= 1;
        if (set == 0) tphase ^= 1u;
        continue;
      }
#pragma unroll 1
      for (int r = e; r < kRowsG; r += 2) {
        const int y = yg * kRowsG + r;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t((set * kRowsG + r) * 64);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        if (r + 2 >= kRowsG) {
          tc_fence_before();
          mbar_arrive(tempty(set));
        }
        uint32_t packed[32];
        epi_stage_bias(sbias, e_bias, e_bias_len, c_base, staged_base, et, bar_id);
        epi_pack(v0, v1, packed, e_act, e_slope, e_bias ? sbias : nullptr);
        if (et == 0) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
        epi_store_row(stage, row, packed);
        fence_proxy_async();
        named_bar_sync(bar_id, 128);
        if (et == 0) {
          tma_store_4d(&p.out, stage, c_base, seg * kTileM, y, img);
          tma_store_commit();
        }
        if (e_stats) {
          float s0, s1, q0, q1;
          epi_stats_rows(stage, ew * 32, lane, 0xffffffffu, s0, s1, q0, q1);
          float* sc = scratch + ew * 128;
          sc[(2 * lane) * 2 + 0] = s0;
          sc[(2 * lane) * 2 + 1] = q0;
          sc[(2 * lane + 1) * 2 + 0] = s1;
          sc[(2 * lane + 1) * 2 + 1] = q1;
          named_bar_sync(bar_id, 128);
          const float tot = scratch[et] + scratch[128 + et] + scratch[256 + et] + scratch[384 + et];
          const size_t tile_lin = size_t(img) * p.stats_tiles_total + p.stats_tile_off + size_t(y) * p.segs + seg;
          e_stats[(tile_lin * p.cout + c_base) * 2 + et] = tot;
        }
      }
      set ^= 1;
      if (set == 0) tphase ^= 1u;
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tg
