// Streaming form of the InstanceNorm passes whose gradient routes sit at the unit's own resolution, optionally plus the
// gradient of the unit's 2x2 average-pooled copy (all 30 of UNet++'s units backward, every forward without pooled /
// upsampled copies, the discriminator's units): normalise+activation forward, backward statistics, backward apply.
//
// Why: the register-staged passes (in_act_fwd_kernel / in_bwd_reduce_kernel) keep <= 8 x 16 B loads per thread in
// flight at 2 x 256 threads per SM and drain them before every compute phase; ncu showed dram 37-49 %, sm 34-45 %
// on the 256^2 launches (profiles/r01_ncu_top_kernels.txt) -- latency bound, ~0.6 of the copy bandwidth. Here the
// bytes in flight do not depend on registers or occupancy: one producer lane per CTA keeps a ring of 8 KiB
// cp.async.bulk (1-D TMA) chunks per input tensor in shared memory (2-8 stages x up to 4 tensors, two CTAs per SM =
// 130-190 KiB in flight per SM against the ~35 KiB that 6.4 TB/s x ~800 ns needs), eight consumer warps read the
// chunks conflict-free (a thread owns one 8-channel group, so the per-channel constants live in registers) and write
// results straight to HBM with 16-byte stores. The grid is persistent: every CTA gets the same number of chunks of
// the flattened (image, pixel) space, so there is no wave quantisation (1216 CTAs on 296 slots before).
#pragma once
#include "tg_ptx.cuh"

namespace tg {

constexpr int kStreamConsumers = 256;             // consumer threads (8 warps); + 1 producer warp
constexpr int kStreamThreads = kStreamConsumers + 32;
constexpr int kStreamPPT = 2;                     // pixels per consumer thread per chunk (1 in the slim form)
constexpr int kStreamChunkBytes = kStreamPPT * kStreamConsumers * 16;   // 8 KiB per tensor per stage (upper bound)
__host__ __device__ constexpr int stream_chunk_bytes(int ppt) { return ppt * kStreamConsumers * 16; }

struct StreamArgs {
  const __nv_bfloat16* in0;   // raw (conv output)
  const __nv_bfloat16* in1;   // gradient route 1 (backward) / unused (forward)
  const __nv_bfloat16* in2;   // optional second same-resolution gradient route
  const __nv_bfloat16* pool;  // optional gradient of the 2x2 average-pooled copy, [N][H/2][W/2][C]: each pixel adds a quarter
  int W;                      // map width (pool route only: a chunk must lie inside one image row)
  __nv_bfloat16* out;         // forward: y; statistics pass: dn (optional); apply pass: dz
  const float* mr;            // [N][C][2] mean, rstd
  const float* gamma;
  const float* beta;
  float* red;                 // statistics pass: out (atomics); apply pass: in
  float* dgamma;
  float* dbeta;
  int N, HW, C, c_valid, act;
  float slope;
  int stages;
  int rev;                    // walk the chunks from the tensor's end: a pass that follows a kernel which walked them
                              // upwards starts on what that kernel left in L2 (tg_in_stream_serpentine)
  int slim;                   // small-footprint form that shares an SM with a persistent weight-gradient CTA
                              // (tg_in_stream_slim): 4 KiB chunks, one CTA per SM, statistics folded through
                              // warp shuffles + shared-memory atomics instead of the [PL][C][2] scratch
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// MODE 0: y = act(S*raw + T)
// MODE 1: dn = (g1 + g2 + pool/4) * act'(A*raw + B); red[n][c] += (sum dn, rstd * (sum dn*raw - mean * sum dn)); dn stored if out
// MODE 2: dz = P*dn + Q*raw + R with dn recomputed as in MODE 1 (P, Q, R from red)
// G2 / POOL: a second same-resolution gradient route / the gradient of the 2x2 average-pooled copy are present
// (compile-time, like the activation's branch-free form: the first version of this kernel decided both per element
// and spent ~245 SASS instructions per 8-channel group -- issue bound at 0.78 of the copy bandwidth; profiles/
// r02_stream_sass.txt has the before / after counts).
//
// One pixel group (8 channels of one pixel) of a chunk, fully inlined; OK = the group lies inside the chunk's valid part.
template <int MODE, bool G2, bool POOL>
struct StreamMath {
  float A[8], B[8];                 // pre-activation n = A*raw + B  (MODE 0: the output itself before the activation)
  float P[8], Pn[8], Q[8], R[8];    // MODE 2 (Pn = P * neg)
  float s0[8], s1[8];               // MODE 1
  float neg;                        // act'(n <= 0): 0 ReLU, slope LeakyReLU, 1 none

  __device__ __forceinline__ void group(const uint4& vr, const uint4& v1, const uint4& v2, const uint4& vp,
                                        __nv_bfloat16* out, bool store) {
    float r[8];
    unpack8(vr, r);
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float n = fmaf(r[q], A[q], B[q]);
        r[q] = fmaf(neg, fminf(n, 0.f), fmaxf(n, 0.f));
      }
      stg16(out, pack8(r));
      return;
    }
    float g[8];
    unpack8(v1, g);
    if (G2) {
      float f[8];
      unpack8(v2, f);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] += f[q];
    }
    if (POOL) {
      float f[8];
      unpack8(vp, f);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] = fmaf(0.25f, f[q], g[q]);
    }
    if (MODE == 1) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        g[q] *= fmaf(r[q], A[q], B[q]) > 0.f ? 1.f : neg;
        s0[q] += g[q];
        s1[q] = fmaf(g[q], r[q], s1[q]);
      }
      if (store) stg16(out, pack8(g));
    } else {
      // dz = P*act'(n)*g + Q*raw + R: the activation's slope is folded into P (Pn = P*neg), one select per element
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float pm = fmaf(r[q], A[q], B[q]) > 0.f ? P[q] : Pn[q];
        g[q] = fmaf(pm, g[q], fmaf(Q[q], r[q], R[q]));
      }
      stg16(out, pack8(g));
    }
  }
};

template <int MODE, int PPT, bool G2, bool POOL>
__global__ void __launch_bounds__(kStreamThreads, 2) in_stream_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // ring slots of one stage: raw | g1 | g2 | pooled gradient (absent routes take no slot)
  constexpr int slot_g2 = 2;
  constexpr int slot_pool = G2 ? 3 : 2;
  constexpr int NIN = MODE == 0 ? 1 : 2 + (G2 ? 1 : 0) + (POOL ? 1 : 0);
  const int CG = a.C >> 3;
  const int PL = kStreamConsumers / CG;                 // pixel lanes; threads with pl >= PL idle (C/8 not a divisor)
  const int CP = PPT * PL;                              // pixels per chunk
  const int CPI = (a.HW + CP - 1) / CP;                 // chunks per image
  const long long TC = (long long)a.N * CPI;
  const long long k0 = TC * blockIdx.x / gridDim.x, k1 = TC * (blockIdx.x + 1) / gridDim.x;
  const int S = a.stages;
  // carve-up: [stages][NIN][chunk] | reduction scratch (MODE 1) | barriers
  const uint32_t base = smem_u32(smem_raw);
  constexpr uint32_t chunk_stride = uint32_t(stream_chunk_bytes(PPT));
  const uint32_t ring_bytes = uint32_t(S) * NIN * chunk_stride;
  float* scratch = reinterpret_cast<float*>(smem_raw + ring_bytes);
  const uint32_t scratch_bytes = MODE == 1 ? uint32_t(a.slim ? 1 : PL) * a.C * 2 * sizeof(float) : 0;
  const uint32_t bar0 = base + ring_bytes + scratch_bytes;       // full[S], then empty[S]
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar0 + 8 * s, 1);
      mbar_init(bar0 + 8 * (S + s), kStreamConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  griddep_sync();   // launched as a programmatic dependent (tg_launch): set-up above, global memory below

  if (tid >= kStreamConsumers) {
    // ------------------------------------------------------------------ producer warp (one lane issues)
    if (tid == kStreamConsumers) {
      int s = 0;
      uint32_t ph = 1;       // the first pass over the ring finds every stage free
      for (long long k = k0; k < k1; ++k) {
        const long long kk = a.rev ? TC - 1 - k : k;
        const int n = int(kk / CPI), j = int(kk % CPI);
        const int p0 = j * CP;
        const int np = min(CP, a.HW - p0);
        const uint32_t bytes = uint32_t(np) * a.C * 2;
        mbar_wait(bar0 + 8 * (S + s), ph);
        const size_t off = (size_t(n) * a.HW + p0) * a.C;
        const uint32_t st = base + uint32_t(s) * NIN * chunk_stride;
        const uint32_t pool_bytes = POOL ? bytes >> 1 : 0;
        mbar_arrive_expect_tx(bar0 + 8 * s, bytes * (NIN - (POOL ? 1 : 0)) + pool_bytes);
        bulk_load_1d(st, a.in0 + off, bytes, bar0 + 8 * s);
        if (MODE != 0) bulk_load_1d(st + chunk_stride, a.in1 + off, bytes, bar0 + 8 * s);
        if (G2) bulk_load_1d(st + slot_g2 * chunk_stride, a.in2 + off, bytes, bar0 + 8 * s);
        if (POOL) {
          // the chunk lies inside image row y (launcher: W % CP == 0): its pooled gradients are np / 2 consecutive
          // pixels of row y / 2 of the half-resolution tensor
          const int y = p0 / a.W, x = p0 % a.W;
          const size_t poff = ((size_t(n) * (a.HW / a.W / 2) + (y >> 1)) * (a.W >> 1) + (x >> 1)) * a.C;
          bulk_load_1d(st + slot_pool * chunk_stride, a.pool + poff, pool_bytes, bar0 + 8 * s);
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int cg = tid % CG, pl = tid / CG;
  const bool active = pl < PL;
  const int c0 = cg * 8;
  if (MODE == 2 && (a.dgamma || a.dbeta) && blockIdx.x == 0) {
    // affine gradients ride along: dgamma[c] += sum_n red[n][c][1], dbeta[c] += sum_n red[n][c][0]
    for (int c = tid; c < a.c_valid; c += kStreamConsumers) {
      float g = 0.f, b = 0.f;
      for (int k = 0; k < a.N; ++k) {
        b += a.red[(size_t(k) * a.C + c) * 2];
        g += a.red[(size_t(k) * a.C + c) * 2 + 1];
      }
      if (a.dgamma) atomicAdd(a.dgamma + c, g);
      if (a.dbeta) atomicAdd(a.dbeta + c, b);
    }
  }
  if (MODE == 1 && a.slim) {
    for (int c = tid; c < 2 * a.C; c += kStreamConsumers) scratch[c] = 0.f;
    named_bar_sync(1, kStreamConsumers);
  }
  StreamMath<MODE, G2, POOL> m;
  m.neg = a.act == 3 ? 0.f : (a.act == 1 ? a.slope : 1.f);
  int cur_n = -1;
  const float inv_hw = 1.f / float(a.HW);
  // per-thread constants of the chunk geometry: byte offset of pixel group p inside a ring slot (and inside the
  // half-size pooled slot), element offset inside the chunk's slice of the output tensor
  uint32_t so[PPT], po[PPT], eo[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int lp = p * PL + pl;
    so[p] = uint32_t(lp) * a.C * 2 + c0 * 2;
    po[p] = uint32_t(lp >> 1) * a.C * 2 + c0 * 2;
    eo[p] = uint32_t(lp) * a.C + c0;
  }
  const bool store_dn = MODE != 1 || a.out != nullptr;

  auto flush = [&]() {
    // MODE 1: block-level reduction of this image's partial sums, then one atomic per (n, c) and CTA
    if (MODE != 1) return;
    if (a.slim) {
      // lanes of a warp that own the same channel group sit CG apart (CG = 8 or 16); other widths go straight
      // to the shared-memory accumulators (8 warps -> at most 8-way contention per address)
      const bool fold = CG == 8 || CG == 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v0 = m.s0[j], v1 = m.s1[j];
        if (fold)
          for (int off = 16; off >= CG; off >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, off);
            v1 += __shfl_xor_sync(0xffffffffu, v1, off);
          }
        if (active && (!fold || (tid & 31) < CG)) {
          atomicAdd(scratch + (c0 + j) * 2, v0);
          atomicAdd(scratch + (c0 + j) * 2 + 1, v1);
        }
      }
      named_bar_sync(1, kStreamConsumers);
      for (int c = tid; c < a.C; c += kStreamConsumers) {
        const float d0 = scratch[c * 2], d1 = scratch[c * 2 + 1];
        scratch[c * 2] = 0.f;
        scratch[c * 2 + 1] = 0.f;
        const float mean = a.mr[(size_t(cur_n) * a.C + c) * 2], rstd = a.mr[(size_t(cur_n) * a.C + c) * 2 + 1];
        atomicAdd(a.red + (size_t(cur_n) * a.C + c) * 2, d0);
        atomicAdd(a.red + (size_t(cur_n) * a.C + c) * 2 + 1, rstd * (d1 - mean * d0));
      }
      named_bar_sync(1, kStreamConsumers);
      return;
    }
    if (active) {
      float* shp = scratch + (size_t(pl) * a.C + c0) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) { shp[2 * j] = m.s0[j]; shp[2 * j + 1] = m.s1[j]; }
    }
    named_bar_sync(1, kStreamConsumers);
    for (int c = tid; c < a.C; c += kStreamConsumers) {
      float d0 = 0.f, d1 = 0.f;
      for (int k = 0; k < PL; ++k) {
        d0 += scratch[(size_t(k) * a.C + c) * 2];
        d1 += scratch[(size_t(k) * a.C + c) * 2 + 1];
      }
      const float mean = a.mr[(size_t(cur_n) * a.C + c) * 2], rstd = a.mr[(size_t(cur_n) * a.C + c) * 2 + 1];
      atomicAdd(a.red + (size_t(cur_n) * a.C + c) * 2, d0);
      atomicAdd(a.red + (size_t(cur_n) * a.C + c) * 2 + 1, rstd * (d1 - mean * d0));
    }
    named_bar_sync(1, kStreamConsumers);
  };

  int s = 0;
  uint32_t ph = 0;
  for (long long k = k0; k < k1; ++k) {
    const long long kk = a.rev ? TC - 1 - k : k;
    const int n = int(kk / CPI), j = int(kk % CPI);
    if (n != cur_n) {
      if (cur_n >= 0) flush();
      cur_n = n;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const size_t kq = size_t(n) * a.C + c0 + q;
        const float g = ld_aff(a.gamma, c0 + q, a.c_valid, 1.f);
        const float b = ld_aff(a.beta, c0 + q, a.c_valid, 0.f);
        const float mean = active ? a.mr[kq * 2] : 0.f, rstd = active ? a.mr[kq * 2 + 1] : 0.f;
        m.A[q] = g * rstd;
        m.B[q] = b - mean * m.A[q];
        m.s0[q] = 0.f;
        m.s1[q] = 0.f;
        if (MODE == 2) {
          const float am = (active ? a.red[kq * 2] : 0.f) * inv_hw, bm = (active ? a.red[kq * 2 + 1] : 0.f) * inv_hw;
          m.P[q] = g * rstd;
          m.Q[q] = -g * rstd * rstd * bm;
          m.R[q] = -m.P[q] * am - m.Q[q] * mean;
          m.Pn[q] = m.P[q] * m.neg;
        }
      }
    }
    const int p0 = j * CP;
    const int np = min(CP, a.HW - p0);
    mbar_wait(bar0 + 8 * s, ph);
    if (active) {
      const uint32_t st = base + uint32_t(s) * NIN * chunk_stride;
      __nv_bfloat16* outc = a.out ? a.out + (size_t(n) * a.HW + p0) * a.C : nullptr;
      // every load of the chunk first (the groups are independent), then the math
      uint4 vr[PPT], v1[PPT], v2[PPT], vp[PPT];
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        vr[p] = lds16(st + so[p]);
        if (MODE != 0) v1[p] = lds16(st + chunk_stride + so[p]);
        if (G2) v2[p] = lds16(st + slot_g2 * chunk_stride + so[p]);
        if (POOL) vp[p] = lds16(st + slot_pool * chunk_stride + po[p]);
      }
      if (np == CP) {
#pragma unroll
        for (int p = 0; p < PPT; ++p) m.group(vr[p], v1[p], v2[p], vp[p], outc + eo[p], store_dn);
      } else {
        // last chunk of an image whose pixel count is not a chunk multiple: groups past the end hold stale bytes
#pragma unroll
        for (int p = 0; p < PPT; ++p)
          if (p * PL + pl < np) m.group(vr[p], v1[p], v2[p], vp[p], outc + eo[p], store_dn);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar0 + 8 * (S + s));
    if (++s == S) { s = 0; ph ^= 1; }
  }
  if (cur_n >= 0) flush();
}

}  // namespace tg
