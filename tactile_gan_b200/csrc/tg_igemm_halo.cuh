// Halo-resident implicit GEMM for 3x3 (any <=3x3-window) stride-1 convolutions with a 64-wide output
// tile -- the Cout=64/128 layers that carry ~70 % of UNet++'s FLOPs and are L2-bound in the per-tap kernel.
//
//   for each work item = (group of R pixel tiles, 64 output channels):
//     for each 64-channel input chunk (over all concat sources):          <- weights stationary
//       TMA: all taps of the weight chunk  [taps][64][64]  (72 KiB, once per R tiles)
//       for each tile r < R:
//         TMA: ONE halo box (16+2)x(8+2) pixels x 64 ch (23 KiB)          <- instead of 9 shifted boxes
//         36 UMMAs (9 taps x 4 K-steps); the tap shift is a descriptor start offset into the halo tile
//         (rows are 128 B; 8-pixel image rows are the 8-row swizzle groups, SBO = 10 rows = 1280 B).
//
// The shifted-start / SBO=1280 descriptors rely on the 128B swizzle being a function of the absolute
// shared-memory address (verified on B200: profiles/r01_umma_offset_probe.log).
// TMEM: R=4 accumulators of 64 columns, double buffered (512 columns) so the epilogue overlaps the MMAs.
#pragma once
#include "tg_igemm.cuh"

namespace tg {

constexpr int kHaloR = 4;                 // pixel tiles per weight load
constexpr int kHaloTH = 16, kHaloTW = 8;  // tile = 16 rows x 8 cols = 128 pixels
constexpr int kHaloBN = 64;
constexpr int kHaloAStage = 23 * 1024;    // >= (16+2)*(8+2)*128 = 23040, 1 KiB aligned
constexpr int kHaloAStages = 2;
constexpr int kHaloBStage = 9 * 64 * 128;  // 72 KiB: up to 9 taps
constexpr int kHaloBStages = 2;
constexpr int kHaloSmem = 1024 + kHaloBStages * kHaloBStage + kHaloAStages * kHaloAStage + 2 * kStoreBytes +
                          4 * 64 * 2 * 4 + 256 + 256;   // ... + barriers + staged bias chunk

struct alignas(64) HaloParams {
  IgemmSrc src[kMaxSrc];  // act box {64, 8+ww, 16+hh, 1}; wgt box {64, 64, wgt_taps}
  CUtensorMap out;        // box {64, 8, 16, 1}
  int num_src;
  int taps;
  int8_t tap_dy[16], tap_dx[16], tap_w[16];  // dy/dx relative to the halo origin (>= 0)
  int org_dy, org_dx;                          // halo origin relative to the output tile origin
  int halo_w;                                  // 8 + extra columns (row pitch of the halo tile in pixels)
  int a_bytes, b_bytes;                        // TMA transaction sizes
  int Ho, Wo, N;
  int tiles_h, tiles_w;
  int n_tiles;  // Cout / 64
  int act;
  float slope;
  const float* bias;
  int bias_len;
  float* stats_partial;
  int stats_tiles_total, stats_tile_off;
  int cout;
  int prefetch;   // issue L2 prefetches one chunk ahead
  int pool_out;   // epilogue stores the 2x2 sum of the tile (`out` map = half-resolution tensor, box {64,4,8,1})
  int* err_flag;
};

__global__ void __launch_bounds__(kNumThreads, 2) igemm_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base;
  const uint32_t a_base = b_base + kHaloBStages * kHaloBStage;
  const uint32_t store_base = a_base + kHaloAStages * kHaloAStage;
  const uint32_t scratch_base = store_base + 2 * kStoreBytes;
  const uint32_t bar_base = scratch_base + 4 * 64 * 2 * 4;
  auto bfull = [&](int s) { return bar_base + 8u * s; };
  auto bempty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto afull = [&](int s) { return bar_base + 8u * (4 + s); };
  auto aempty = [&](int s) { return bar_base + 8u * (6 + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (10 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 12;
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
      for (int s = 0; s < 2; ++s) {
        mbar_init(bfull(s), 1);
        mbar_init(bempty(s), 1);
        mbar_init(afull(s), 1);
        mbar_init(aempty(s), 1);
        mbar_init(tfull(s), 1);
        mbar_init(tempty(s), 128);
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
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int m_tiles = p.N * tiles_per_img;
  const int groups = (m_tiles + kHaloR - 1) / kHaloR;
  const int total_items = groups * p.n_tiles;
  // Weight-stationary mode: when every weight chunk of this CTA's output-channel tile fits the B stages (one or
  // two 64-channel input chunks) and the CTA keeps the same n_tile for all its items (item stride is a multiple
  // of n_tiles), the weights are loaded ONCE per CTA instead of once per item -- for the single-chunk layers
  // that is 44 % of the bytes staged through shared memory.
  const bool resident = chunks <= kHaloBStages && (int(gridDim.x) % p.n_tiles) == 0;
  const bool prefetch = p.prefetch != 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int bs = 0, as = 0;
      uint32_t bphase = 0, aphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int n_tile = item % p.n_tiles;
        const int group = item / p.n_tiles;
        const int t0 = group * kHaloR;
        const int cnt = min(kHaloR, m_tiles - t0);
        for (int s = 0; s < p.num_src; ++s) {
          const IgemmSrc& src = p.src[s];
          for (int cc = 0; cc < src.c_chunks; ++cc) {
            if (!resident || item == int(blockIdx.x)) {
              mbar_wait_guard(bempty(bs), bphase ^ 1, p.err_flag, 21);
              mbar_arrive_expect_tx(bfull(bs), uint32_t(p.b_bytes));
              tma_load_3d(b_base + bs * kHaloBStage, &src.wgt, bfull(bs), cc * kChunkK, n_tile * kHaloBN, 0);
              if (++bs == kHaloBStages) { bs = 0; bphase ^= 1; }
            }
            if (prefetch) {
              // the shared-memory ring holds only two halo tiles (~1 tile = 0.6 us of MMA work ahead), less than
              // an HBM round trip: pull the NEXT chunk's tiles (next source / next item at the end) into L2 now
              int ps = s, pcc = cc + 1, pt0 = t0, pcnt = cnt;
              if (pcc == src.c_chunks) { pcc = 0; ++ps; }
              if (ps == p.num_src) {
                ps = 0;
                const int nitem = item + int(gridDim.x);
                pt0 = (nitem / p.n_tiles) * kHaloR;
                pcnt = nitem < total_items ? min(kHaloR, m_tiles - pt0) : 0;
              }
              for (int r = 0; r < pcnt; ++r) {
                const int mt = pt0 + r;
                const int img = mt / tiles_per_img, t_in = mt % tiles_per_img;
                tma_prefetch_4d(&p.src[ps].act, pcc * kChunkK, (t_in % p.tiles_w) * kHaloTW + p.org_dx,
                                (t_in / p.tiles_w) * kHaloTH + p.org_dy, img);
              }
            }
            for (int r = 0; r < cnt; ++r) {
              const int mt = t0 + r;
              const int img = mt / tiles_per_img, t_in = mt % tiles_per_img;
              const int ho0 = (t_in / p.tiles_w) * kHaloTH, wo0 = (t_in % p.tiles_w) * kHaloTW;
              mbar_wait_guard(aempty(as), aphase ^ 1, p.err_flag, 22);
              mbar_arrive_expect_tx(afull(as), uint32_t(p.a_bytes));
              tma_load_4d(a_base + as * kHaloAStage, &src.act, afull(as), cc * kChunkK, wo0 + p.org_dx,
                          ho0 + p.org_dy, img);
              if (++as == kHaloAStages) { as = 0; aphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, kHaloBN, 0, 0);
      // descriptor halves that never change, and the per-tap start-address steps (in 16-byte units), are
      // computed once: the issue loop below is then two adds + one MMA per instruction, which a single
      // thread must sustain at one MMA every 32 cycles (N = 64)
      const uint32_t a_hi = umma_desc_hi_sw128(uint32_t(p.halo_w) * 128u);
      const uint32_t b_hi = umma_desc_hi_sw128(1024u);
      uint32_t a_step[9], b_step[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        a_step[t] = t < p.taps ? uint32_t(p.tap_dy[t] * p.halo_w + p.tap_dx[t]) * 8u : 0u;
        b_step[t] = t < p.taps ? uint32_t(p.tap_w[t]) * (kHaloBN * 8u) : 0u;
      }
      const int ntaps = p.taps;
      int bs = 0, as = 0, set = 0;
      uint32_t bphase = 0, aphase = 0, tphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int group = item / p.n_tiles;
        const int cnt = min(kHaloR, m_tiles - group * kHaloR);
        mbar_wait_guard(tempty(set), tphase ^ 1, p.err_flag, 23);
        tc_fence_after();
        const bool first_item = item == int(blockIdx.x);
        for (int c = 0; c < chunks; ++c) {
          if (resident) bs = c;
          if (!resident || first_item) mbar_wait_guard(bfull(bs), bphase, p.err_flag, 24);
          const uint32_t b_lo0 = umma_desc_lo(b_base + bs * kHaloBStage, 16);
          for (int r = 0; r < cnt; ++r) {
            mbar_wait_guard(afull(as), aphase, p.err_flag, 25);
            tc_fence_after();
            const uint32_t a_lo0 = umma_desc_lo(a_base + as * kHaloAStage, 16);
            const uint32_t d_tmem = tmem_base + uint32_t((set * kHaloR + r) * kHaloBN);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (tap < ntaps) {
                const uint32_t a_lo = a_lo0 + a_step[tap], b_lo = b_lo0 + b_step[tap];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_split(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc,
                                 (tap | k) != 0 ? 1u : uint32_t(c != 0));
              }
            }
            umma_commit(aempty(as));
            if (++as == kHaloAStages) { as = 0; aphase ^= 1; }
          }
          if (!resident) {
            umma_commit(bempty(bs));
            if (++bs == kHaloBStages) { bs = 0; bphase ^= 1; }
          }
        }
        umma_commit(tfull(set));
        set ^= 1;
        if (set == 0) tphase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (128 threads)
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ew = et >> 5;
    int set = 0;
    uint32_t tphase = 0, chunk_ctr = 0;
    float* scratch = reinterpret_cast<float*>(smem_gen + (scratch_base - smem_base));
    float* sbias = reinterpret_cast<float*>(smem_gen + (bar_base + 256 - smem_base));
    int staged_base = -1;
    const int e_act = p.act, e_bias_len = p.bias_len;
    const float e_slope = p.slope;
    const float* e_bias = p.bias;
    float* e_stats = p.stats_partial;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int n_tile = item % p.n_tiles;
      const int group = item / p.n_tiles;
      const int t0 = group * kHaloR;
      const int cnt = min(kHaloR, m_tiles - t0);
      const int c_base = n_tile * kHaloBN;
      mbar_wait_guard(tfull(set), tphase, p.err_flag, 26);
      tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < cnt; ++r, ++chunk_ctr) {
        const int mt = t0 + r;
        const int img = mt / tiles_per_img, t_in = mt % tiles_per_img;
        const int ho0 = (t_in / p.tiles_w) * kHaloTH, wo0 = (t_in % p.tiles_w) * kHaloTW;
        const uint32_t sb = chunk_ctr & 1;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t((set * kHaloR + r) * kHaloBN);
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        if (r == cnt - 1) {
          tc_fence_before();
          mbar_arrive(tempty(set));
        }
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
            tma_store_4d(&p.out, store_base + sb * kStoreBytes, c_base, wo0 >> 1, ho0 >> 1, img);
            tma_store_commit();
          }
          continue;
        }
        epi_stage_bias(sbias, e_bias, e_bias_len, c_base, staged_base, et);
        epi_pack(v0, v1, packed, e_act, e_slope, e_bias ? sbias : nullptr);
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        epi_store_row(store_base + sb * kStoreBytes, row, packed);
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (et == 0) {
          tma_store_4d(&p.out, store_base + sb * kStoreBytes, c_base, wo0, ho0, img);
          tma_store_commit();
        }
        if (e_stats) {
          uint32_t valid_mask = 0xffffffffu;
          if (ho0 + kHaloTH > p.Ho || wo0 + kHaloTW > p.Wo) {
            valid_mask = 0u;
            for (int i = 0; i < 32; ++i) {
              const int rr = ew * 32 + i;
              if (ho0 + (rr >> 3) < p.Ho && wo0 + (rr & 7) < p.Wo) valid_mask |= 1u << i;
            }
          }
          float s0, s1, q0, q1;
          epi_stats_rows(store_base + sb * kStoreBytes, ew * 32, lane, valid_mask, s0, s1, q0, q1);
          float* sc = scratch + ew * 128;
          sc[(2 * lane) * 2 + 0] = s0;
          sc[(2 * lane) * 2 + 1] = q0;
          sc[(2 * lane + 1) * 2 + 0] = s1;
          sc[(2 * lane + 1) * 2 + 1] = q1;
          named_bar_sync(1, 128);
          const float tot = scratch[et] + scratch[128 + et] + scratch[256 + et] + scratch[384 + et];
          const size_t tile_lin = size_t(img) * p.stats_tiles_total + p.stats_tile_off + t_in;
          e_stats[(tile_lin * p.cout + c_base) * 2 + et] = tot;
        }
      }
      set ^= 1;
      if (set == 0) tphase ^= 1;
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
