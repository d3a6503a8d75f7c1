// Bandwidth-bound tail of the G+D step: layout packs, InstanceNorm (+ReLU/LeakyReLU) forward /
// backward / double-backward, pooling & nearest-upsample fused into the normalise pass, the 1x1
// feature-map head, loss reductions, gradient-penalty pieces and the fused Adam + weight re-pack.
// All activations are NHWC bf16 with C % 64 == 0; every access is a 128-bit vector of 8 channels.
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include "tg_api_internal.h"
#include "tg_ptx.cuh"

namespace tg {

static_assert(sizeof(AdamTensor) == 136, "AdamTensor layout is part of the C-ABI (tg_adam_step)");

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  // bf16 -> fp32 is the upper half of the word: one shift / one mask per element
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// affine / bias vectors hold only the real channels; padded channels behave as (1, 0)
__device__ __forceinline__ float ld_aff(const float* p, int c, int c_valid, float dflt) {
  return (p && c < c_valid) ? __ldg(p + c) : dflt;
}

// Branch-free: neg = act'(x <= 0) (0 ReLU, slope LeakyReLU, 1 none) is loop invariant, so the per-element work is
// two min/max + one fma (forward) or one compare-select (gradient) instead of a switch on `act`.
__device__ __forceinline__ float act_neg(int act, float slope) { return act == 3 ? 0.f : (act == 1 ? slope : 1.f); }
__device__ __forceinline__ float act_fwd(float x, int act, float slope) {
  return fmaf(act_neg(act, slope), fminf(x, 0.f), fmaxf(x, 0.f));
}
__device__ __forceinline__ float act_grad(float x, int act, float slope) {
  return x > 0.f ? 1.f : act_neg(act, slope);
}

}  // namespace tg
#include "tg_stream.cuh"
namespace tg {

// ------------------------------------------------------------------ layout packs
// out[n,h,w, c_off + j] = wa[n]*A[n,j,h,w] + wb[n]*B[n,j,h,w]   (fp32 NCHW -> bf16 NHWC), j < cj <= 8.
// Only the 8-channel group containing c_off is written; the other groups of `out` keep their value.
__global__ void pack_nchw_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                 const float* __restrict__ wa, const float* __restrict__ wb,
                                 __nv_bfloat16* __restrict__ out, int N, int HW, int cj, int C,
                                 int c_off) {
  const size_t total = size_t(N) * HW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const int n = int(i / HW);
    const int pix = int(i % HW);
    const float fa = wa ? wa[n] : 1.f;
    const float fb = wb ? wb[n] : 0.f;
    __nv_bfloat16* o = out + i * C + c_off;
    for (int j = 0; j < cj; ++j) {
      float v = fa * A[(size_t(n) * cj + j) * HW + pix];
      if (B) v += fb * B[(size_t(n) * cj + j) * HW + pix];
      o[j] = __float2bfloat16(v);
    }
  }
}

// bf16 NHWC (C channels, first cj used, starting at c_off) -> fp32 NCHW
__global__ void unpack_nhwc_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                   int N, int HW, int C, int c_off, int cj, float scale) {
  const size_t total = size_t(N) * HW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const int n = int(i / HW);
    const int pix = int(i % HW);
    const __nv_bfloat16* s = in + i * C + c_off;
    for (int j = 0; j < cj; ++j) out[(size_t(n) * cj + j) * HW + pix] = scale * __bfloat162float(s[j]);
  }
}

// ------------------------------------------------------------------ thin-channel layers as 1x1 GEMMs
// The discriminator's first conv (6 -> 64, 3x3, stride 2) and its head (512 -> 1, 3x3) would waste >= 90 % of
// a 64-channel-padded implicit GEMM. Both are restated so the tensor-core kernel sees a dense 1-tap problem:
//   first layer: the pack writes the im2col rows directly, cols[n,oy,ox, tap*Ct + c] = x[n, c, oy*s+dy, ox*s+dx]
//                (Ct = ca + cb <= 7 -> 9*Ct <= 64 channels), so the conv is a 1x1 GEMM with K = 64;
//   head:        hc[n,y,x,tap] = sum_ci X[n,y,x,ci] * W[tap,ci] is a 1x1 GEMM with 9 output channels and
//                z[n,oy,ox] = sum_tap hc[n,oy+dy,ox+dx,tap]; the backward scatters dz to dzc[n,y,x,tap] = dz[n,y-dy,x-dx].

// cols <- im2col of cat(A, wa*B + wb*B2) (fp32 NCHW sources), k x k taps, stride s, no padding. One thread per
// output pixel writes its 128-byte row. KK / CA / CB > 0: compile-time shape (the row lives in registers);
// 0: run-time shape (generic fallback, row in local memory).
template <int KK, int CA, int CB>
__global__ void __launch_bounds__(128)
im2col_pack_kernel(const float* __restrict__ A, const float* __restrict__ B,
                   const float* __restrict__ B2, const float* __restrict__ wa,
                   const float* __restrict__ wb, __nv_bfloat16* __restrict__ cols, int N, int ca_,
                   int cb_, int H, int W, int Ho, int Wo, int k_, int stride) {
  const int k = KK ? KK : k_, ca = KK ? CA : ca_, cb = KK ? CB : cb_;
  const size_t total = size_t(N) * Ho * Wo;
  const int ct = ca + cb;
  const size_t HW = size_t(H) * W;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int ox = int(i % Wo), oy = int((i / Wo) % Ho), n = int(i / (size_t(Wo) * Ho));
    const float fa = wa ? wa[n] : 1.f, fb = wb ? wb[n] : 0.f;
    __align__(16) __nv_bfloat16 row[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) row[j] = __float2bfloat16(0.f);
    const float* An = A ? A + size_t(n) * ca * HW : nullptr;
    const float* Bn = B + size_t(n) * cb * HW;
    const float* B2n = B2 ? B2 + size_t(n) * cb * HW : nullptr;
#pragma unroll
    for (int t = 0; t < (KK ? KK * KK : 16); ++t) {
      if (t >= k * k) break;
      const size_t pix = size_t(oy * stride + t / k) * W + ox * stride + t % k;
#pragma unroll
      for (int c = 0; c < (KK ? CA : 8); ++c) {
        if (c >= ca) break;
        if (An) row[t * ct + c] = __float2bfloat16(__ldg(An + c * HW + pix));
      }
#pragma unroll
      for (int c = 0; c < (KK ? CB : 8); ++c) {
        if (c >= cb) break;
        float v = fa * __ldg(Bn + c * HW + pix);
        if (B2n) v += fb * __ldg(B2n + c * HW + pix);
        row[t * ct + ca + c] = __float2bfloat16(v);
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(cols + i * 64);
    const uint4* src = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = src[j];
  }
}

// g[n, c, y, x] = sum over taps whose window covers (y, x) of dcols[n, oy, ox, tap*ct + c_off + c]  (fp32 NCHW)
__global__ void col2im_grad_kernel(const __nv_bfloat16* __restrict__ dcols, float* __restrict__ g, int N, int ct,
                                   int c_off, int cj, int H, int W, int Ho, int Wo, int k, int stride, float scale) {
  const size_t total = size_t(N) * H * W;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H), n = int(i / (size_t(W) * H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < k * k; ++t) {
      const int yy = y - t / k, xx = x - t % k;
      if (yy < 0 || xx < 0 || yy % stride || xx % stride) continue;
      const int oy = yy / stride, ox = xx / stride;
      if (oy >= Ho || ox >= Wo) continue;
      const __nv_bfloat16* r = dcols + ((size_t(n) * Ho + oy) * Wo + ox) * 64 + t * ct + c_off;
      for (int c = 0; c < cj; ++c) acc[c] += __bfloat162float(r[c]);
    }
    for (int c = 0; c < cj; ++c) g[(size_t(n) * cj + c) * H * W + size_t(y) * W + x] = scale * acc[c];
  }
}

// out[n,oy,ox,0] = act(sum_tap hc[n, oy + t/k, ox + t%k, tap] + bias); channels 1..7 of the pixel are zeroed
__global__ void head_gather_kernel(const __nv_bfloat16* __restrict__ hc, const float* __restrict__ bias,
                                   __nv_bfloat16* __restrict__ out, int N, int Hi, int Wi, int Ho, int Wo, int k,
                                   int C, int act, float slope) {
  const size_t total = size_t(N) * Ho * Wo;
  const float b = bias ? __ldg(bias) : 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int ox = int(i % Wo), oy = int((i / Wo) % Ho), n = int(i / (size_t(Wo) * Ho));
    float v = b;
    for (int t = 0; t < k * k; ++t)
      v += __bfloat162float(hc[((size_t(n) * Hi + oy + t / k) * Wi + ox + t % k) * 64 + t]);
    if (act == 2) v = 1.f / (1.f + __expf(-v));
    else v = act_fwd(v, act, slope);
    float f[8] = {v, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    stg16(out + i * C, pack8(f));
  }
}

// dzc[n,y,x,tap] = dz[n, y - t/k, x - t%k, 0] (0 outside), taps < 16; the first 16 channels of every row are written
__global__ void head_scatter_kernel(const __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dzc, int N,
                                    int Hi, int Wi, int Ho, int Wo, int k, int C) {
  const size_t total = size_t(N) * Hi * Wi;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % Wi), y = int((i / Wi) % Hi), n = int(i / (size_t(Wi) * Hi));
    float f[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      f[t] = 0.f;
      if (t < k * k) {
        const int oy = y - t / k, ox = x - t % k;
        if (oy >= 0 && ox >= 0 && oy < Ho && ox < Wo)
          f[t] = __bfloat162float(dz[((size_t(n) * Ho + oy) * Wo + ox) * C]);
      }
    }
    float lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo[j] = f[j]; hi[j] = f[8 + j]; }
    stg16(dzc + i * 64, pack8(lo));
    stg16(dzc + i * 64 + 8, pack8(hi));
  }
}

// nsq[n] += sum over an fp32 NCHW image gradient of (g + 1e-16)^2   (gradient penalty norm, util.py:92)
__global__ void gp_normsq_img_kernel(const float* __restrict__ g, size_t per_img, float* __restrict__ nsq) {
  __shared__ float sh[32];
  const int n = blockIdx.y;
  float acc = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < per_img; i += size_t(gridDim.x) * blockDim.x) {
    const float v = g[size_t(n) * per_img + i] + 1e-16f;
    acc += v * v;
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(nsq + n, tot);
}

// ------------------------------------------------------------------ InstanceNorm forward
// partial: [N][T][C][2] (sum, sumsq) from the conv epilogue -> mr: [N][C][2] (mean, rstd)
__global__ void in_finalize_kernel(const float* __restrict__ partial, float* __restrict__ mr, int T,
                                   int C, float inv_count, float eps) {
  // block = 16 channels x 64 tile lanes (grid (C/16, N)): a (tile, 16-channel) row is one 128-byte line
  __shared__ float sh[64][16][2];
  griddep_sync();
  const int n = blockIdx.y;
  const int cl = threadIdx.x & 15;
  const int c = blockIdx.x * 16 + cl;
  const int tl = threadIdx.x >> 4;  // 0..63
  float s = 0.f, q = 0.f;
  const float2* src = reinterpret_cast<const float2*>(partial) + (size_t(n) * T) * C + c;
  for (int t = tl; t < T; t += 64) {
    const float2 v = __ldg(src + size_t(t) * C);
    s += v.x;
    q += v.y;
  }
  sh[tl][cl][0] = s;
  sh[tl][cl][1] = q;
  __syncthreads();
  if (tl == 0) {
    for (int k = 1; k < 64; ++k) {
      s += sh[k][cl][0];
      q += sh[k][cl][1];
    }
    const float mean = s * inv_count;
    const float var = fmaxf(q * inv_count - mean * mean, 0.f);
    mr[(size_t(n) * C + c) * 2 + 0] = mean;
    mr[(size_t(n) * C + c) * 2 + 1] = rsqrtf(var + eps);
  }
}

// Direct statistics for a tensor the conv epilogue could not reduce (tiles spanning images).
__global__ void in_stats_direct_kernel(const __nv_bfloat16* __restrict__ raw, float* __restrict__ mr,
                                       int HW, int C, float eps) {
  // grid (C/64, N), block 256: 8 channel groups x 32 pixel lanes
  __shared__ float sh[32][64][2];
  const int n = blockIdx.y;
  const int g = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + g * 8;
  float s[8] = {0}, q[8] = {0};
  for (int pix = pl; pix < HW; pix += 32) {
    float f[8];
    unpack8(ldg16(raw + (size_t(n) * HW + pix) * C + c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[pl][g * 8 + j][0] = s[j]; sh[pl][g * 8 + j][1] = q[j]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < 32; ++k) { ss += sh[k][threadIdx.x][0]; qq += sh[k][threadIdx.x][1]; }
    const float mean = ss / HW;
    const float var = fmaxf(qq / HW - mean * mean, 0.f);
    const int c = blockIdx.x * 64 + threadIdx.x;
    mr[(size_t(n) * C + c) * 2 + 0] = mean;
    mr[(size_t(n) * C + c) * 2 + 1] = rsqrtf(var + eps);
  }
}

// Thread layout shared by the normalise / backward passes: a block works on one image (blockIdx.y) and a
// strip of its pixels (blockIdx.x); a thread owns one 8-channel group (so per-channel constants live in
// registers for the whole strip) and walks pixels with stride PL = blockDim.x / (C/8). Consecutive threads
// touch consecutive 16-byte vectors, i.e. whole 128-byte lines.
struct StripIdx {
  int cg, pl, PL, c0, p0, p1;
};
__device__ __forceinline__ StripIdx strip_index(int C, int units, int rev = 0) {
  StripIdx s;
  const int CG = C >> 3;
  s.PL = blockDim.x / CG;
  s.cg = threadIdx.x % CG;
  s.pl = threadIdx.x / CG;
  s.c0 = s.cg * 8;
  const int strip = (units + gridDim.x - 1) / gridDim.x;
  s.p0 = (rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * strip;   // rev: serpentine order, see serp_mask()
  s.p1 = min(units, s.p0 + strip);
  return s;
}

// y = act(gamma * (raw - mean) * rstd + beta) = act(S*raw + T); optional 2x2 pooled copy and 2x
// nearest-upsampled copy written by the same pass (units = pixels, or 2x2 quads when POOL/UP).
template <int POOL, bool UP>
__global__ void __launch_bounds__(256)
in_act_fwd_kernel(const __nv_bfloat16* __restrict__ raw, const float* __restrict__ mr,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ pool,
                  __nv_bfloat16* __restrict__ up, int H, int W, int C, int c_valid, int act, float slope,
                  int rev) {
  griddep_sync();
  const int n = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  constexpr bool QUAD = POOL != 0 || UP;
  const int H2 = H >> 1, W2 = W >> 1;
  const StripIdx t = strip_index(C, QUAD ? H2 * W2 : H * W, rev);
  if (t.pl >= t.PL) return;
  float S[8], T[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float g = ld_aff(gamma, t.c0 + j, c_valid, 1.f);
    const float b = ld_aff(beta, t.c0 + j, c_valid, 0.f);
    const float mean = mr[(size_t(n) * C + t.c0 + j) * 2];
    const float rstd = mr[(size_t(n) * C + t.c0 + j) * 2 + 1];
    S[j] = g * rstd;
    T[j] = b - mean * S[j];
  }
  const size_t img = size_t(n) * H * W;
  if (!QUAD) {
    int pix = t.p0 + t.pl;
    for (; pix + 3 * t.PL < t.p1; pix += 4 * t.PL) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ldg16(raw + (img + pix + k * t.PL) * C + t.c0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fwd(fmaf(f[j], S[j], T[j]), act, slope);
        stg16(y + (img + pix + k * t.PL) * C + t.c0, pack8(f));
      }
    }
    for (; pix < t.p1; pix += t.PL) {
      float f[8];
      unpack8(ldg16(raw + (img + pix) * C + t.c0), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = act_fwd(fmaf(f[j], S[j], T[j]), act, slope);
      stg16(y + (img + pix) * C + t.c0, pack8(f));
    }
  } else {
    for (int q = t.p0 + t.pl; q < t.p1; q += t.PL) {
      const int qy = q / W2, qx = q % W2;
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = ldg16(raw + (img + size_t(2 * qy + (k >> 1)) * W + 2 * qx + (k & 1)) * C + t.c0);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = POOL == 2 ? -3.0e38f : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = 2 * qy + (k >> 1), xx = 2 * qx + (k & 1);
        float f[8];
        unpack8(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fwd(fmaf(f[j], S[j], T[j]), act, slope);
        const uint4 pv = pack8(f);
        stg16(y + (img + size_t(yy) * W + xx) * C + t.c0, pv);
        if (POOL) {
          float fr[8];
          unpack8(pv, fr);  // pool the bf16-rounded values the next layer would read
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = POOL == 2 ? fmaxf(acc[j], fr[j]) : acc[j] + fr[j];
        }
        if (UP) {
          const size_t ubase = (size_t(n) * 2 * H + 2 * yy) * (2 * W) + 2 * xx;
          stg16(up + ubase * C + t.c0, pv);
          stg16(up + (ubase + 1) * C + t.c0, pv);
          stg16(up + (ubase + 2 * W) * C + t.c0, pv);
          stg16(up + (ubase + 2 * W + 1) * C + t.c0, pv);
        }
      }
      if (POOL) {
        if (POOL == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
        }
        stg16(pool + ((size_t(n) * H2 + qy) * W2 + qx) * C + t.c0, pack8(acc));
      }
    }
  }
}

// ------------------------------------------------------------------ InstanceNorm backward
// dn = (g_same + 0.25*g_pool[h/2,w/2] + sum_{2x2} g_up[2h+a,2w+b] (+ g_max routed)) * act'(n)
// red[n][c] += (sum dn, sum dn*xhat). Block = strip of pixels of one image, all channels.
// pool_mode: 1 avg (g_pool spread evenly), 2 max (g_pool routed to the arg-max element of y).
struct InBwdArgs {
  const __nv_bfloat16* raw;   // conv output (pre-norm); nullptr => no norm (y is act(raw+bias))
  const __nv_bfloat16* y;     // post-activation output (needed for no-norm layers and max-pool routing)
  const float* mr;
  const float* gamma;
  const float* beta;
  const __nv_bfloat16* g_same;
  const __nv_bfloat16* g_pool;
  const __nv_bfloat16* g_up;
  __nv_bfloat16* dn;
  float* red;                 // [N][C][2]
  int N, H, W, C, c_valid, act, pool_mode;
  int up_pooled;              // g_up is already summed to this tensor's resolution (pool_out conv epilogue)
  float slope;
  // second pass (PASS == 1): dz = P*dn + Q*raw + R with dn recomputed from the same loads; optional affine gradients
  __nv_bfloat16* dz;
  float* dgamma;
  float* dbeta;
  int rev;                    // serpentine block order (serp_mask())
};

// Per-pixel work of the backward reduce for the general case (pooled / upsampled gradient routes).
__device__ __forceinline__ void in_bwd_gather(const InBwdArgs& a, int n, int pix, int c0, float (&g)[8]) {
  const int yy = pix / a.W, xx = pix % a.W;
  if (a.g_up && a.up_pooled) {
    float f[8];
    unpack8(ldg16(a.g_up + (size_t(n) * a.H * a.W + pix) * a.C + c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += f[j];
  } else if (a.g_up) {
    const int WU = 2 * a.W;
    const size_t ub = (size_t(n) * 2 * a.H + 2 * yy) * WU + 2 * xx;
    const uint4 u0 = ldg16(a.g_up + ub * a.C + c0), u1 = ldg16(a.g_up + (ub + 1) * a.C + c0);
    const uint4 u2 = ldg16(a.g_up + (ub + WU) * a.C + c0), u3 = ldg16(a.g_up + (ub + WU + 1) * a.C + c0);
    float f[8];
    unpack8(u0, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += f[j];
    unpack8(u1, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += f[j];
    unpack8(u2, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += f[j];
    unpack8(u3, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += f[j];
  }
  if (a.g_pool) {
    const int H2 = a.H >> 1, W2 = a.W >> 1;
    float f[8];
    unpack8(ldg16(a.g_pool + ((size_t(n) * H2 + (yy >> 1)) * W2 + (xx >> 1)) * a.C + c0), f);
    if (a.pool_mode == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += 0.25f * f[j];
    } else {
      // max-pool routing: first element (row-major in the 2x2 window) equal to the window max
      float best[8];
      int first[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -3.0e38f; first[j] = 0; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y2 = (yy & ~1) + (k >> 1), x2 = (xx & ~1) + (k & 1);
        float o[8];
        unpack8(ldg16(a.y + ((size_t(n) * a.H + y2) * a.W + x2) * a.C + c0), o);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (o[j] > best[j]) { best[j] = o[j]; first[j] = k; }
      }
      const int mine = ((yy & 1) << 1) | (xx & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (first[j] == mine) g[j] += f[j];
    }
  }
}

// PLAIN: the only gradient route is g_same (no pooled / upsampled copies) -> four pixels per iteration with
// all eight 16-byte loads in flight before the first use.
// PASS 0: statistics pass -- red[n][c] += (sum dn, sum dn*xhat); dn is stored only if a.dn is given (the GP double
//         backward keeps it) or the layer has no norm (then dn IS dz).
// PASS 1: apply pass -- the same loads again (raw + the gradient routes, 2 tensors for a plain unit), dn recomputed
//         in fp32 and dz = rstd*gamma*(dn - mean(dn) - xhat*mean(dn*xhat)) written: the backward of a unit moves
//         5 tensor-sizes through HBM instead of the 6 of "store dn, re-read dn".
template <bool PLAIN, int PASS>
__global__ void __launch_bounds__(256, 2) in_bwd_reduce_kernel(const InBwdArgs a) {
  extern __shared__ float shm[];  // [PL][C][2]
  griddep_sync();
  const int n = a.rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int HW = a.H * a.W;
  const StripIdx t = strip_index(a.C, HW, a.rev);
  const int c0 = t.c0;
  if (PASS == 1 && (a.dgamma || a.dbeta) && blockIdx.x == 0 && blockIdx.y == 0) {
    // affine gradients ride along (one block): dgamma[c] += sum_n red[n][c][1], dbeta[c] += sum_n red[n][c][0]
    for (int c = threadIdx.x; c < a.c_valid; c += blockDim.x) {
      float g = 0.f, b = 0.f;
      for (int k = 0; k < a.N; ++k) {
        b += a.red[(size_t(k) * a.C + c) * 2];
        g += a.red[(size_t(k) * a.C + c) * 2 + 1];
      }
      if (a.dgamma) atomicAdd(a.dgamma + c, g);
      if (a.dbeta) atomicAdd(a.dbeta + c, b);
    }
  }
  // n = A*raw + B (pre-activation); xhat = (raw - mean)*rstd is folded into the finalisation:
  // sum dn*xhat = rstd * (sum dn*raw - mean * sum dn)
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float g = ld_aff(a.gamma, c0 + j, a.c_valid, 1.f);
    const float b = ld_aff(a.beta, c0 + j, a.c_valid, 0.f);
    const float mean = a.mr ? a.mr[(size_t(n) * a.C + c0 + j) * 2] : 0.f;
    const float rstd = a.mr ? a.mr[(size_t(n) * a.C + c0 + j) * 2 + 1] : 1.f;
    A[j] = g * rstd;
    B[j] = b - mean * A[j];
  }
  float s0[8] = {0}, s1[8] = {0};
  float P[8], Q[8], R[8];
  if (PASS == 1) {
    const float inv = 1.f / float(HW);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t k = size_t(n) * a.C + c0 + j;
      const float g = ld_aff(a.gamma, c0 + j, a.c_valid, 1.f);
      const float mean = a.mr[k * 2], rstd = a.mr[k * 2 + 1];
      const float am = a.red[k * 2] * inv, bm = a.red[k * 2 + 1] * inv;
      P[j] = g * rstd;
      Q[j] = -g * rstd * rstd * bm;
      R[j] = -P[j] * am - Q[j] * mean;
    }
  }
  const __nv_bfloat16* src = a.raw ? a.raw : a.y;
  const bool has_raw = a.raw != nullptr;
  const int act = a.act;
  const float slope = a.slope;
  auto finish = [&](const uint4& vr, float (&g)[8], size_t lin) {
    float r[8];
    unpack8(vr, r);
    if (PASS == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dn = g[j] * act_grad(fmaf(r[j], A[j], B[j]), act, slope);
        g[j] = fmaf(P[j], dn, fmaf(Q[j], r[j], R[j]));
      }
      stg16(a.dz + lin, pack8(g));
      return;
    }
    if (has_raw) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] *= act_grad(fmaf(r[j], A[j], B[j]), act, slope);
        s0[j] += g[j];
        s1[j] = fmaf(g[j], r[j], s1[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= act_grad(r[j], act, slope);
    }
    if (a.dn) stg16(a.dn + lin, pack8(g));
  };
  if (t.pl < t.PL) {
    const size_t img = size_t(n) * HW;
    int pix = t.p0 + t.pl;
    if (PLAIN) {
      // same-resolution routes only: g_same and / or the pre-summed upsample route (either may be absent)
      const __nv_bfloat16* ga = a.g_same ? a.g_same : a.g_up;
      const __nv_bfloat16* gb = (a.g_same && a.g_up) ? a.g_up : nullptr;
      for (; pix + 3 * t.PL < t.p1; pix += 4 * t.PL) {
        uint4 vr[4], vg[4], vu[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const size_t lin = (img + pix + k * t.PL) * a.C + c0;
          vr[k] = ldg16(src + lin);
          vg[k] = ldg16(ga + lin);
          if (gb) vu[k] = ldg16(gb + lin);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float g[8];
          unpack8(vg[k], g);
          if (gb) {
            float f[8];
            unpack8(vu[k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] += f[j];
          }
          finish(vr[k], g, (img + pix + k * t.PL) * a.C + c0);
        }
      }
      for (; pix < t.p1; pix += t.PL) {
        const size_t lin = (img + pix) * a.C + c0;
        float g[8];
        const uint4 vr = ldg16(src + lin);
        unpack8(ldg16(ga + lin), g);
        if (gb) {
          float f[8];
          unpack8(ldg16(gb + lin), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += f[j];
        }
        finish(vr, g, lin);
      }
    } else {
      for (; pix + t.PL < t.p1; pix += 2 * t.PL) {
        const size_t l0 = (img + pix) * a.C + c0, l1 = (img + pix + t.PL) * a.C + c0;
        const uint4 r0 = ldg16(src + l0), r1 = ldg16(src + l1);
        float g0[8] = {0}, g1[8] = {0};
        if (a.g_same) {
          const uint4 v0 = ldg16(a.g_same + l0), v1 = ldg16(a.g_same + l1);
          unpack8(v0, g0);
          unpack8(v1, g1);
        }
        in_bwd_gather(a, n, pix, c0, g0);
        in_bwd_gather(a, n, pix + t.PL, c0, g1);
        finish(r0, g0, l0);
        finish(r1, g1, l1);
      }
      for (; pix < t.p1; pix += t.PL) {
        const size_t lin = (img + pix) * a.C + c0;
        const uint4 vr = ldg16(src + lin);
        float g[8] = {0};
        if (a.g_same) unpack8(ldg16(a.g_same + lin), g);
        in_bwd_gather(a, n, pix, c0, g);
        finish(vr, g, lin);
      }
    }
  }
  if (PASS == 1 || !a.red) return;
  float* shp = shm + (size_t(t.pl) * a.C + c0) * 2;
  if (t.pl < t.PL) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { shp[2 * j] = s0[j]; shp[2 * j + 1] = s1[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float d0 = 0.f, d1 = 0.f;
    for (int k = 0; k < t.PL; ++k) {
      d0 += shm[(size_t(k) * a.C + c) * 2];
      d1 += shm[(size_t(k) * a.C + c) * 2 + 1];
    }
    const float mean = a.mr[(size_t(n) * a.C + c) * 2], rstd = a.mr[(size_t(n) * a.C + c) * 2 + 1];
    atomicAdd(a.red + (size_t(n) * a.C + c) * 2, d0);
    atomicAdd(a.red + (size_t(n) * a.C + c) * 2 + 1, rstd * (d1 - mean * d0));
  }
}

// dz = rstd * gamma * (dn - mean(dn) - xhat * mean(dn * xhat)) = P*dn + Q*raw + R
__global__ void __launch_bounds__(256)
in_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dn, const __nv_bfloat16* __restrict__ raw,
                    const float* __restrict__ mr, const float* __restrict__ gamma,
                    const float* __restrict__ red, __nv_bfloat16* __restrict__ dz, int HW, int C,
                    int c_valid, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int n = blockIdx.y;
  // affine gradients ride along (one block): dgamma[c] += sum_n red[n][c][1], dbeta[c] += sum_n red[n][c][0]
  if ((dgamma || dbeta) && blockIdx.x == 0 && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < c_valid; c += blockDim.x) {
      float g = 0.f, b = 0.f;
      for (int k = 0; k < int(gridDim.y); ++k) {
        b += red[(size_t(k) * C + c) * 2];
        g += red[(size_t(k) * C + c) * 2 + 1];
      }
      // atomic: sub-batch engines of the same network run this concurrently on their own streams
      if (dgamma) atomicAdd(dgamma + c, g);
      if (dbeta) atomicAdd(dbeta + c, b);
    }
  }
  const StripIdx t = strip_index(C, HW);
  if (t.pl >= t.PL) return;
  const float inv = 1.f / float(HW);
  float P[8], Q[8], R[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const size_t k = size_t(n) * C + t.c0 + j;
    const float g = ld_aff(gamma, t.c0 + j, c_valid, 1.f);
    const float mean = mr[k * 2], rstd = mr[k * 2 + 1];
    const float am = red[k * 2] * inv, bm = red[k * 2 + 1] * inv;
    P[j] = g * rstd;
    Q[j] = -g * rstd * rstd * bm;
    R[j] = -P[j] * am - Q[j] * mean;
  }
  const size_t img = size_t(n) * HW;
  int pix = t.p0 + t.pl;
  for (; pix + 3 * t.PL < t.p1; pix += 4 * t.PL) {
    uint4 vd[4], vr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      vd[k] = ldg16(dn + (img + pix + k * t.PL) * C + t.c0);
      vr[k] = ldg16(raw + (img + pix + k * t.PL) * C + t.c0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float d[8], r[8];
      unpack8(vd[k], d);
      unpack8(vr[k], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = fmaf(P[j], d[j], fmaf(Q[j], r[j], R[j]));
      stg16(dz + (img + pix + k * t.PL) * C + t.c0, pack8(d));
    }
  }
  for (; pix < t.p1; pix += t.PL) {
    float d[8], r[8];
    unpack8(ldg16(dn + (img + pix) * C + t.c0), d);
    unpack8(ldg16(raw + (img + pix) * C + t.c0), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = fmaf(P[j], d[j], fmaf(Q[j], r[j], R[j]));
    stg16(dz + (img + pix) * C + t.c0, pack8(d));
  }
}

// dgamma[c] += sum_n red[n][c][1]; dbeta[c] += sum_n red[n][c][0]
__global__ void affine_grad_kernel(const float* __restrict__ red, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, int N, int C, int c_valid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_valid) return;
  float g = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    b += red[(size_t(n) * C + c) * 2];
    g += red[(size_t(n) * C + c) * 2 + 1];
  }
  if (dgamma) dgamma[c] += g;
  if (dbeta) dbeta[c] += b;
}

// db[c] += sum over rows of dz[row][c], c < c_valid.  grid (C/64, strips)
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ dz, float* __restrict__ db,
                                 size_t rows, int C, int c_valid) {
  __shared__ float sh[32][65];
  const int g = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + g * 8;
  const size_t strip = (rows + gridDim.y - 1) / gridDim.y;
  const size_t r0 = blockIdx.y * strip, r1 = min(rows, r0 + strip);
  float s[8] = {0};
  for (size_t r = r0 + pl; r < r1; r += 32) {
    float f[8];
    unpack8(ldg16(dz + r * C + c0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[pl][g * 8 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += sh[k][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < c_valid) atomicAdd(db + c, t);
  }
}

// ------------------------------------------------------------------ InstanceNorm double backward
// (gradient penalty). With xhat, r = rstd, dxh = gamma*dn (first backward), u = adj(dz):
//   adj(dn) = gamma * r * (u - mean(u) - xhat*mean(u*xhat))              [continues the upward sweep]
//   adj(z)  = r^2 * ( -b*(u - c1) - c2*(dxh - a) + xhat*(3*b*c2 + a*c1 - e) ) [injected downward]
//   adj(gamma) += sum dn * r*(u - c1 - xhat*c2)
// where a = mean(dxh), b = mean(dxh*xhat), c1 = mean(u), c2 = mean(u*xhat), e = mean(u*dxh).
// Pass 1 reduces (sum u, sum u*xhat, sum u*dn) into red2[N][C][4].
__global__ void in_bwd2_reduce_kernel(const __nv_bfloat16* __restrict__ u,
                                      const __nv_bfloat16* __restrict__ raw,
                                      const __nv_bfloat16* __restrict__ dn,
                                      const float* __restrict__ mr, float* __restrict__ red2, int HW,
                                      int C) {
  extern __shared__ float shm[];  // [PL][C][3]
  const int CG = C >> 3;
  const int PL = blockDim.x / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int n = blockIdx.y;
  const int strip = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * strip, p1 = min(HW, p0 + strip);
  const int c0 = cg * 8;
  float s0[8] = {0}, s1[8] = {0}, s2[8] = {0};
  if (pl < PL) {
    float mean[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = mr[(size_t(n) * C + c0 + j) * 2];
      rstd[j] = mr[(size_t(n) * C + c0 + j) * 2 + 1];
    }
    for (int pix = p0 + pl; pix < p1; pix += PL) {
      const size_t lin = (size_t(n) * HW + pix) * C + c0;
      float fu[8], fr[8], fd[8];
      unpack8(ldg16(u + lin), fu);
      unpack8(ldg16(raw + lin), fr);
      unpack8(ldg16(dn + lin), fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fr[j] - mean[j]) * rstd[j];
        s0[j] += fu[j];
        s1[j] += fu[j] * xh;
        s2[j] += fu[j] * fd[j];
      }
    }
    float* shp = shm + (size_t(pl) * C + c0) * 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) { shp[3 * j] = s0[j]; shp[3 * j + 1] = s1[j]; shp[3 * j + 2] = s2[j]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < PL; ++k) t += shm[size_t(k) * C * 3 + i];
    atomicAdd(red2 + (size_t(n) * C + i / 3) * 4 + (i % 3), t);
  }
}

// Pass 2: writes adj_da = adj(dn) * act'(n) (upward) and adj_z (downward injection); accumulates
// adj(gamma) partial sums into red2[..][3].
__global__ void in_bwd2_apply_kernel(const __nv_bfloat16* __restrict__ u,
                                     const __nv_bfloat16* __restrict__ raw,
                                     const __nv_bfloat16* __restrict__ dn,
                                     const float* __restrict__ mr, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ red1,
                                     float* __restrict__ red2, __nv_bfloat16* __restrict__ adj_da,
                                     __nv_bfloat16* __restrict__ adj_z, int HW, int C, int c_valid,
                                     int act, float slope) {
  extern __shared__ float shm[];  // [PL][C]
  const int CG = C >> 3;
  const int PL = blockDim.x / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int n = blockIdx.y;
  const int strip = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * strip, p1 = min(HW, p0 + strip);
  const int c0 = cg * 8;
  const float inv = 1.f / float(HW);
  float sg[8] = {0};
  if (pl < PL) {
    float mean[8], rstd[8], gm[8], bt[8], A[8], Bc[8], c1[8], c2[8], e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t k = size_t(n) * C + c0 + j;
      mean[j] = mr[k * 2];
      rstd[j] = mr[k * 2 + 1];
      gm[j] = ld_aff(gamma, c0 + j, c_valid, 1.f);
      bt[j] = ld_aff(beta, c0 + j, c_valid, 0.f);
      A[j] = gm[j] * red1[k * 2] * inv;       // mean(dxh)
      Bc[j] = gm[j] * red1[k * 2 + 1] * inv;  // mean(dxh * xhat)
      c1[j] = red2[k * 4] * inv;
      c2[j] = red2[k * 4 + 1] * inv;
      e[j] = gm[j] * red2[k * 4 + 2] * inv;   // mean(u * dxh)
    }
    for (int pix = p0 + pl; pix < p1; pix += PL) {
      const size_t lin = (size_t(n) * HW + pix) * C + c0;
      float fu[8], fr[8], fd[8], oa[8], oz[8];
      unpack8(ldg16(u + lin), fu);
      unpack8(ldg16(raw + lin), fr);
      unpack8(ldg16(dn + lin), fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fr[j] - mean[j]) * rstd[j];
        const float adxh = rstd[j] * (fu[j] - c1[j] - xh * c2[j]);
        sg[j] += fd[j] * adxh;
        oa[j] = gm[j] * adxh * act_grad(gm[j] * xh + bt[j], act, slope);
        const float dxh = gm[j] * fd[j];
        oz[j] = rstd[j] * rstd[j] *
                (-Bc[j] * (fu[j] - c1[j]) - c2[j] * (dxh - A[j]) +
                 xh * (3.f * Bc[j] * c2[j] + A[j] * c1[j] - e[j]));
      }
      stg16(adj_da + lin, pack8(oa));
      stg16(adj_z + lin, pack8(oz));
    }
    float* shp = shm + size_t(pl) * C + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) shp[j] = sg[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < PL; ++k) t += shm[size_t(k) * C + i];
    atomicAdd(red2 + (size_t(n) * C + i) * 4 + 3, t);
  }
}

// dgamma[c] += sum_n red2[n][c][3]
__global__ void gamma_grad2_kernel(const float* __restrict__ red2, float* __restrict__ dgamma, int N,
                                   int C, int c_valid) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_valid) return;
  float g = 0.f;
  for (int n = 0; n < N; ++n) g += red2[(size_t(n) * C + c) * 4 + 3];
  dgamma[c] += g;
}

// ------------------------------------------------------------------ elementwise helpers
// out = a + b (bf16, same shape); b optional
__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           __nv_bfloat16* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8;
       i += size_t(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(ldg16(a + i * 8), x);
    unpack8(ldg16(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    stg16(out + i * 8, pack8(x));
  }
}

// out = g * act'(y)  where y is the post-activation value (sign-preserving activations only)
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                               __nv_bfloat16* __restrict__ out, size_t n8, int act, float slope) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8;
       i += size_t(gridDim.x) * blockDim.x) {
    float x[8], o[8];
    unpack8(ldg16(g + i * 8), x);
    unpack8(ldg16(y + i * 8), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= act_grad(o[j], act, slope);
    stg16(out + i * 8, pack8(x));
  }
}

// ------------------------------------------------------------------ 1x1 feature-map head (64 -> co<=4)
// out[n,o,h,w] = act(sum_c x[n,h,w,c] * w[o][c] + b[o]); fp32 NCHW out.
// C == 64, HW % 4 == 0: eight lanes share a pixel (one 16-byte load each = the pixel's whole 128-byte line), four
// pixels per lane group and iteration; partial dot products are folded over the eight lanes with shuffles and lane o
// of the group stores output plane o as one float4. (The per-thread-pixel form below reads a different line per
// lane and load: 8x the L1 wavefronts for the same bytes.)
__global__ void __launch_bounds__(256)
fmap_fwd64_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                  float* __restrict__ out, int N, int HW, int co, int use_tanh, int ci) {
  __shared__ float sw[4 * 64 + 4];
  for (int i = threadIdx.x; i < 4 * 64; i += blockDim.x)
    sw[i] = (i / 64 < co && (i % 64) < ci) ? w[(i / 64) * ci + i % 64] : 0.f;
  if (threadIdx.x < 4) sw[4 * 64 + threadIdx.x] = threadIdx.x < co ? b[threadIdx.x] : 0.f;
  __syncthreads();
  const int grp = threadIdx.x & 7;
  float wreg[4][8];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) wreg[o][j] = sw[o * 64 + grp * 8 + j];
  const float bias = sw[4 * 64 + (grp & 3)];
  const size_t quads = size_t(N) * HW / 4;
  // the loop bound is per warp (four lane groups): the shuffles below need all 32 lanes
  for (size_t qb = blockIdx.x * size_t(32) + (threadIdx.x >> 5) * 4; qb < quads; qb += size_t(gridDim.x) * 32) {
    const size_t q = qb + ((threadIdx.x >> 3) & 3);
    const bool valid = q < quads;
    const size_t i0 = (valid ? q : quads - 1) * 4;
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ldg16(x + (i0 + k) * 64 + grp * 8);
    float acc[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8(v[k], f);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a = fmaf(f[j], wreg[o][j], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        acc[k][o] = a;
      }
    }
    if (valid && grp < co) {
      float r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t = acc[k][0];
        if (grp == 1) t = acc[k][1];
        if (grp == 2) t = acc[k][2];
        if (grp == 3) t = acc[k][3];
        t += bias;
        r[k] = use_tanh ? tanhf(t) : t;
      }
      const size_t n = i0 / HW, pix = i0 % HW;
      *reinterpret_cast<float4*>(out + (n * co + grp) * HW + pix) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

__global__ void fmap_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ b, float* __restrict__ out, int N, int HW,
                                int C, int co, int use_tanh, int ci) {
  __shared__ float sw[4 * 64 + 4];
  // w is the module's own [co][ci] weight (torch layout, ci <= C real input channels); padded channels read 0
  for (int i = threadIdx.x; i < co * C; i += blockDim.x) sw[i] = (i % C) < ci ? w[(i / C) * ci + i % C] : 0.f;
  if (threadIdx.x < co) sw[4 * 64 + threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const size_t total = size_t(N) * HW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const int n = int(i / HW), pix = int(i % HW);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < C; c0 += 8) {
      float f[8];
      unpack8(ldg16(x + i * C + c0), f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < co) acc[o] += f[j] * sw[o * C + c0 + j];
    }
    for (int o = 0; o < co; ++o) {
      float v = acc[o] + sw[4 * 64 + o];
      if (use_tanh) v = tanhf(v);
      out[(size_t(n) * co + o) * HW + pix] = v;
    }
  }
}

// dz[o] = (g1 + g2)[n,o,pix] * (1 - out^2) ; dx[n,pix,c] = sum_o dz[o] w[o][c];
// dw[o][c] += sum dz[o] x[c]; db[o] += sum dz[o]
// grid (strips, N): a block walks a strip of one image's pixels, thread = (pixel lane, 8-channel group), two pixels
// per iteration so ~20 independent loads are in flight per thread.
__global__ void __launch_bounds__(256)
fmap_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ out, const float* __restrict__ g1,
                const float* __restrict__ g2, __nv_bfloat16* __restrict__ dx,
                float* __restrict__ dw, float* __restrict__ db, int N, int HW, int C,
                int co, int use_tanh, int replicas, int ci) {
  __shared__ float sw[4 * 64];
  __shared__ float sacc[4 * 64 + 4];
  for (int i = threadIdx.x; i < co * C; i += blockDim.x) sw[i] = (i % C) < ci ? w[(i / C) * ci + i % C] : 0.f;
  for (int i = threadIdx.x; i < 4 * 64 + 4; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float lw[4][8];  // this lane accumulates dw for channel group (lane & 7) only
  float lb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) lw[o][j] = 0.f;
  const int grp = threadIdx.x & 7, pl = threadIdx.x >> 3;  // C == 64 -> 8 groups of 8 channels, 32 pixel lanes
  const int n = blockIdx.y;
  const int strip = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * strip, p1 = min(HW, p0 + strip);
  float wreg[4][8];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) wreg[o][j] = o < co ? sw[o * C + grp * 8 + j] : 0.f;

  // branch-free loads (a null g1 / g2 reads the other operand with weight 0), so all of an iteration's loads issue
  // back to back instead of one dependent round trip per `if`
  const float* ga = g1 ? g1 : g2;
  const float* gb = g2 ? g2 : g1;
  const float wa = g1 ? 1.f : 0.f, wb = g2 ? 1.f : 0.f, wt = use_tanh ? 1.f : 0.f;
  auto load_dz = [&](int pix, float (&dz)[4]) {
    float a[4], b[4], t[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const size_t k = (size_t(n) * co + (o < co ? o : 0)) * HW + pix;
      a[o] = __ldg(ga + k);
      b[o] = __ldg(gb + k);
      t[o] = __ldg(out + k);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) dz[o] = o < co ? (wa * a[o] + wb * b[o]) * (1.f - wt * t[o] * t[o]) : 0.f;
  };
  auto emit = [&](int pix, const uint4& vx, const float (&dz)[4]) {
    float f[8], d[8];
    unpack8(vx, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        acc = fmaf(dz[o], wreg[o][j], acc);
        lw[o][j] = fmaf(dz[o], f[j], lw[o][j]);
      }
      d[j] = acc;
    }
    if (grp == 0) {
#pragma unroll
      for (int o = 0; o < 4; ++o) lb[o] += dz[o];
    }
    if (dx) stg16(dx + (size_t(n) * HW + pix) * C + grp * 8, pack8(d));
  };
  int pix = p0 + pl;
  for (; pix + 32 < p1; pix += 64) {
    float dza[4], dzb[4];
    const uint4 xa = ldg16(x + (size_t(n) * HW + pix) * C + grp * 8);
    const uint4 xb = ldg16(x + (size_t(n) * HW + pix + 32) * C + grp * 8);
    load_dz(pix, dza);
    load_dz(pix + 32, dzb);
    emit(pix, xa, dza);
    emit(pix + 32, xb, dzb);
  }
  for (; pix < p1; pix += 32) {
    float dza[4];
    const uint4 xa = ldg16(x + (size_t(n) * HW + pix) * C + grp * 8);
    load_dz(pix, dza);
    emit(pix, xa, dza);
  }
  // lanes with equal grp are 8 apart: reduce over them with shuffles, then smem, then atomics
#pragma unroll
  for (int o = 0; o < 4; ++o) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = lw[o][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      lw[o][j] = v;
    }
    float v = lb[o];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    lb[o] = v;
  }
  if ((threadIdx.x & 31) < 8) {
    for (int o = 0; o < co; ++o) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[o * 64 + grp * 8 + j], lw[o][j]);
      if (grp == 0) atomicAdd(&sacc[256 + o], lb[o]);
    }
  }
  __syncthreads();
  // ~1200 blocks adding into the same 6 cache lines serialise in L2: spread them over `replicas` copies of dw
  float* dwr = dw + size_t((blockIdx.y * gridDim.x + blockIdx.x) % replicas) * co * 64;
  for (int i = threadIdx.x; i < co * 64; i += blockDim.x) atomicAdd(dwr + i, sacc[i]);
  if (threadIdx.x < co) atomicAdd(db + threadIdx.x, sacc[256 + threadIdx.x]);
}

// ------------------------------------------------------------------ losses
// GAN loss on channel 0 of pred (NHWC bf16, C channels). `pred` holds sigmoid(z) when has_sigmoid.
// mode: 0 ls, 1 ce, 2 w, 3 hinge. Writes dz (grad wrt the pre-activation z, channel 0; other
// channels of dz must already be zero) scaled by `scale`/numel and adds scale*mean-loss to *loss.
// For samples >= n_loss_samples (GP interpolates riding in the same batch) dz = gp_seed (d sum(pred)/dz).
__global__ void gan_loss_kernel(const __nv_bfloat16* __restrict__ pred, const float* __restrict__ label,
                                float label_const, int mode, int target_is_real, int for_disc,
                                int has_sigmoid, float scale, int n0, int n1, int HW, int C,
                                float* __restrict__ loss, __nv_bfloat16* __restrict__ dz, int write_dz) {
  __shared__ float sh[32];
  const size_t numel = size_t(n1 - n0) * HW;
  float acc = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < numel;
       i += size_t(gridDim.x) * blockDim.x) {
    const size_t k = (size_t(n0) * HW + i) * C;
    const float p = __bfloat162float(pred[k]);
    const float t = label ? label[i] : label_const;
    float l = 0.f, dp = 0.f;  // loss element, dloss/dp (p = network output incl. sigmoid if any)
    if (mode == 0) { const float d = p - t; l = d * d; dp = 2.f * d; }
    else if (mode == 1) {  // BCE with logits on p
      l = fmaxf(p, 0.f) - p * t + log1pf(__expf(-fabsf(p)));
      dp = 1.f / (1.f + __expf(-p)) - t;
    } else if (mode == 2) { l = target_is_real ? -p : p; dp = target_is_real ? -1.f : 1.f; }
    else {
      if (for_disc) {
        const float m = target_is_real ? p - 1.f : -p - 1.f;
        l = -fminf(m, 0.f);
        dp = m < 0.f ? (target_is_real ? -1.f : 1.f) : 0.f;
      } else { l = -p; dp = -1.f; }
    }
    acc += l;
    if (write_dz) {
      float g = dp * scale / float(numel);
      if (has_sigmoid) g *= p * (1.f - p);
      dz[k] = __float2bfloat16(g);
    }
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, tot * scale / float(numel));
}

// dz[k] = p(1-p) (or 1 without sigmoid) for samples [n0,n1): the seed of d sum(pred) / d input.
__global__ void gp_first_seed_kernel(const __nv_bfloat16* __restrict__ pred, int has_sigmoid, int n0,
                                     int n1, int HW, int C, __nv_bfloat16* __restrict__ dz) {
  const size_t numel = size_t(n1 - n0) * HW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < numel;
       i += size_t(gridDim.x) * blockDim.x) {
    const size_t k = (size_t(n0) * HW + i) * C;
    const float p = __bfloat162float(pred[k]);
    dz[k] = __float2bfloat16(has_sigmoid ? p * (1.f - p) : 1.f);
  }
}

// Second-order seed at the top of D: adj(z5) = w * (1-2p) * p(1-p), w = adj(dz5) (channel 0).
__global__ void gp_top_kernel(const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ pred,
                              int has_sigmoid, size_t numel, int C, __nv_bfloat16* __restrict__ out) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < numel;
       i += size_t(gridDim.x) * blockDim.x) {
    const size_t k = i * C;
    const float p = __bfloat162float(pred[k]);
    const float v = has_sigmoid ? __bfloat162float(w[k]) * (1.f - 2.f * p) * p * (1.f - p) : 0.f;
    out[k] = __float2bfloat16(v);
  }
}

// L1 mean between fp32 NCHW tensors: *loss += mean|a-b| (unscaled, as the reference logs it,
// train.py:145-146); grad_a = sign(a-b) * scale / numel (scale = lambda_a) optional.
__global__ void l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t numel,
                               float scale, float* __restrict__ loss, float* __restrict__ grad_a) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < numel;
       i += size_t(gridDim.x) * blockDim.x) {
    const float d = a[i] - b[i];
    acc += fabsf(d);
    if (grad_a) grad_a[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * scale / float(numel);
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, tot / float(numel));
}

// weighted L1 (or L2) mean between two bf16 tensors (feature maps), all channels valid.
__global__ void feat_loss_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                 size_t n8, float weight, int l2, float* __restrict__ loss) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8;
       i += size_t(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(ldg16(a + i * 8), x);
    unpack8(ldg16(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = x[j] - y[j];
      acc += l2 ? d * d : fabsf(d);
    }
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, tot * weight / float(n8 * 8));
}

// d/da of  weight * mean|a - b|  as a bf16 tensor: g = sign(a - b) * weight / numel  (version-1 perceptual term)
__global__ void feat_loss_grad_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                      size_t n8, float weight, __nv_bfloat16* __restrict__ g) {
  const float s = weight / float(n8 * 8);
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(ldg16(a + i * 8), x);
    unpack8(ldg16(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = x[j] > y[j] ? s : (x[j] < y[j] ? -s : 0.f);
    stg16(g + i * 8, pack8(x));
  }
}

// 2x2 pooling of a bf16 NHWC tensor (mode 1 average, 2 max) for layers without a normalise pass (VGG16)
__global__ void pool_fwd_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ pool, int N, int H,
                                int W, int C, int mode) {
  const int CG = C >> 3, H2 = H >> 1, W2 = W >> 1;
  const size_t total = size_t(N) * H2 * W2 * CG;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int cg = int(i % CG);
    const size_t q = i / CG;
    const int qx = int(q % W2), qy = int((q / W2) % H2), n = int(q / (size_t(W2) * H2));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = mode == 2 ? -3.0e38f : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8(ldg16(y + ((size_t(n) * H + 2 * qy + (k >> 1)) * W + 2 * qx + (k & 1)) * C + cg * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = mode == 2 ? fmaxf(acc[j], f[j]) : acc[j] + 0.25f * f[j];
    }
    stg16(pool + q * C + cg * 8, pack8(acc));
  }
}

// VGGPerceptualLoss input transform (util.py:120-129): repeat a 1-channel image to 3, (x - mean) / std, bilinear
// resize (align_corners = False, as F.interpolate) to OH x OW; fp32 NCHW -> bf16 NHWC (first 8-channel group).
struct BilinTap { int i0, i1; float l0, l1; };
__device__ __forceinline__ BilinTap bilin_tap(int o, int in_size, float scale) {
  float src = scale * (float(o) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  BilinTap t;
  t.i0 = min(int(src), in_size - 1);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = src - float(t.i0);
  t.l0 = 1.f - t.l1;
  return t;
}
__constant__ float kVggMean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float kVggStd[3] = {0.229f, 0.224f, 0.225f};

__global__ void vgg_prep_fwd_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int cs,
                                    int H, int W, int OH, int OW, int C, int resize) {
  const size_t total = size_t(N) * OH * OW;
  const float sy = resize ? float(H) / float(OH) : 1.f, sx = resize ? float(W) / float(OW) : 1.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int ox = int(i % OW), oy = int((i / OW) % OH), n = int(i / (size_t(OW) * OH));
    const BilinTap ty = bilin_tap(oy, H, sy), tx = bilin_tap(ox, W, sx);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = x + (size_t(n) * cs + (cs == 3 ? c : 0)) * H * W;
      const float m = kVggMean[c], is = 1.f / kVggStd[c];
      const float v00 = (p[size_t(ty.i0) * W + tx.i0] - m) * is, v01 = (p[size_t(ty.i0) * W + tx.i1] - m) * is;
      const float v10 = (p[size_t(ty.i1) * W + tx.i0] - m) * is, v11 = (p[size_t(ty.i1) * W + tx.i1] - m) * is;
      f[c] = ty.l0 * (tx.l0 * v00 + tx.l1 * v01) + ty.l1 * (tx.l0 * v10 + tx.l1 * v11);
    }
    stg16(out + i * C, pack8(f));
  }
}

// transpose of the transform above: grad[n, c, y, x] += scale * sum over output pixels of w * g[n, oy, ox, c] / std
__global__ void vgg_prep_bwd_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ grad, int N, int cs,
                                    int H, int W, int OH, int OW, int C, int resize, float scale) {
  const size_t total = size_t(N) * OH * OW;
  const float sy = resize ? float(H) / float(OH) : 1.f, sx = resize ? float(W) / float(OW) : 1.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int ox = int(i % OW), oy = int((i / OW) % OH), n = int(i / (size_t(OW) * OH));
    const BilinTap ty = bilin_tap(oy, H, sy), tx = bilin_tap(ox, W, sx);
    float f[8];
    unpack8(ldg16(g + i * C), f);
    float gc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gc[c] = scale * f[c] / kVggStd[c];
    if (cs != 3) { gc[0] += gc[1] + gc[2]; }
    for (int c = 0; c < (cs == 3 ? 3 : 1); ++c) {
      float* p = grad + (size_t(n) * cs + c) * H * W;
      atomicAdd(p + size_t(ty.i0) * W + tx.i0, gc[c] * ty.l0 * tx.l0);
      atomicAdd(p + size_t(ty.i0) * W + tx.i1, gc[c] * ty.l0 * tx.l1);
      atomicAdd(p + size_t(ty.i1) * W + tx.i0, gc[c] * ty.l1 * tx.l0);
      atomicAdd(p + size_t(ty.i1) * W + tx.i1, gc[c] * ty.l1 * tx.l1);
    }
  }
}

// Device-side augmentation + ToTensor / Normalize of one uint8 HWC pair (PairedDataset.py:30-44,80-92): optional
// horizontal flip, then an affine warp given as the INVERSE map in 16.16 fixed point (so the index math is exact
// integer arithmetic, identical on host and device); source image bilinear, target mask nearest, zero border.
//   q[n] = {flip, a00, a01, a02, a10, a11, a12, 0}: src_x = (a00*x + a01*y + a02) / 65536, src_y likewise.
// out_a[n,c,y,x] = (img/255 - 0.5) / 0.5 (fp32 NCHW), out_b[n,c,y,x] = mask/255.
__global__ void augment_pair_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                    const long long* __restrict__ q, float* __restrict__ out_a,
                                    float* __restrict__ out_b, int N, int H, int W, int ca, int cb) {
  const size_t total = size_t(N) * H * W;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % W), y = int((i / W) % H), n = int(i / (size_t(W) * H));
    const long long* p = q + size_t(n) * 8;
    long long sx = p[1] * x + p[2] * y + p[3];
    const long long sy = p[4] * x + p[5] * y + p[6];
    if (p[0]) sx = ((long long)(W - 1) << 16) - sx;
    // ---- image: bilinear, zero outside
    const long long x0 = sx >> 16, y0 = sy >> 16;
    const float fx = float(sx & 0xFFFF) * (1.f / 65536.f), fy = float(sy & 0xFFFF) * (1.f / 65536.f);
    const uint8_t* ib = img + size_t(n) * H * W * ca;
    for (int c = 0; c < ca; ++c) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long xx = x0 + (k & 1), yy = y0 + (k >> 1);
        v[k] = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? float(ib[(size_t(yy) * W + xx) * ca + c]) : 0.f;
      }
      const float top = v[0] + fx * (v[1] - v[0]), bot = v[2] + fx * (v[3] - v[2]);
      const float val = (top + fy * (bot - top)) / 255.f;   // ToTensor divides by 255
      out_a[(size_t(n) * ca + c) * H * W + size_t(y) * W + x] = (val - 0.5f) / 0.5f;
    }
    // ---- mask: nearest (round half up), zero outside
    const long long xn = (sx + 32768) >> 16, yn = (sy + 32768) >> 16;
    const bool in = xn >= 0 && xn < W && yn >= 0 && yn < H;
    const uint8_t* mb = mask + size_t(n) * H * W * cb;
    for (int c = 0; c < cb; ++c)
      out_b[(size_t(n) * cb + c) * H * W + size_t(y) * W + x] =
          in ? float(mb[(size_t(yn) * W + xn) * cb + c]) / 255.f : 0.f;
  }
}

// Fuzzy evaluation sums of test.py:113-124 per image: stats[n] = (sum o*r, sum o^2 + r^2, sum min(o, r), sum r)
// for the generator output o and the target r (fp32, any layout, per_img elements each).
__global__ void eval_fuzzy_kernel(const float* __restrict__ o, const float* __restrict__ r, size_t per_img,
                                  float* __restrict__ stats) {
  __shared__ float sh[32];
  const int n = blockIdx.y;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < per_img; i += size_t(gridDim.x) * blockDim.x) {
    const float x = o[size_t(n) * per_img + i], y = r[size_t(n) * per_img + i];
    a0 = fmaf(x, y, a0);
    a1 += x * x + y * y;
    a2 += fminf(x, y);
    a3 += y;
  }
  float t;
  t = block_sum(a0, sh); if (threadIdx.x == 0) atomicAdd(stats + n * 4 + 0, t);
  t = block_sum(a1, sh); if (threadIdx.x == 0) atomicAdd(stats + n * 4 + 1, t);
  t = block_sum(a2, sh); if (threadIdx.x == 0) atomicAdd(stats + n * 4 + 2, t);
  t = block_sum(a3, sh); if (threadIdx.x == 0) atomicAdd(stats + n * 4 + 3, t);
}

// ------------------------------------------------------------------ ConvLSTM (generators/BCDUNet.py:6-103)
// fp32 NCHW (image stride sn elements, so a time slice X[:, t] of a (B,T,C,H,W) sequence needs no copy) ->
// bf16 NHWC [N][HW][Cpad]; channels >= C of the row are written as zero. One 64-pixel x 64-channel tile per block
// through shared memory: coalesced 256-byte reads along pixels, 16-byte vector writes along channels.
__global__ void __launch_bounds__(256)
pack_nchw_tiled_kernel(const float* __restrict__ in, long long sn, __nv_bfloat16* __restrict__ out, int C, int HW,
                       int Cpad) {
  __shared__ float t[64][65];
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 4096; i += 256) {
    const int cl = i >> 6, pl = i & 63;
    const int c = c0 + cl, p = p0 + pl;
    t[cl][pl] = (c < C && p < HW) ? __ldg(in + size_t(n) * sn + size_t(c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < 512; v += 256) {
    const int pl = v >> 3, k = v & 7;
    const int p = p0 + pl, c = c0 + 8 * k;
    if (p < HW && c < Cpad) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = t[8 * k + j][pl];
      stg16(out + (size_t(n) * HW + p) * Cpad + c, pack8(f));
    }
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// Peephole gates of one ConvLSTM time step (BCDUNet.py:32-47). z = conv(cat[X, H_prev]) + bias arrives from the
// implicit-GEMM kernel as bf16 NHWC with the four gate groups i | f | g | o at channel offsets 0, C, 2C, 3C; the
// cell state, the peepholes W_c* [C][H][W] and the returned H live in the module's own fp32 NCHW layout. A block
// owns 32 pixels x 32 channels: the z tile is transposed through shared memory so that phase 1 reads NHWC rows with
// 16-byte vectors and phase 2 walks pixel-fastest over the NCHW tensors (128-byte segments per channel row).
//   i = s(zi + W_ci*c)   f = s(zf + W_cf*c)   c' = f*c + i*act(zg)   o = s(zo + W_co*c')   h = o*act(c')
// h is written twice: fp32 NCHW (module output / sequence slot, image stride h_sn) and bf16 NHWC (the next step's
// conv operand). c_prev == nullptr: zero state (first step). c_prev may alias c_out.
__global__ void __launch_bounds__(256)
convlstm_gates_kernel(const __nv_bfloat16* __restrict__ z, int zc, const float* __restrict__ w_ci,
                      const float* __restrict__ w_cf, const float* __restrict__ w_co, const float* c_prev,
                      long long cp_sn, float* c_out, long long co_sn, float* __restrict__ h_out, long long h_sn,
                      __nv_bfloat16* __restrict__ h_nhwc, int hc, int HW, int C, int act) {
  __shared__ float zt[4][32][33];
  __shared__ float ht[32][33];
  const int n = blockIdx.z, cb = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int v = threadIdx.x; v < 512; v += 256) {
    const int g = v >> 7, pl = (v >> 2) & 31, k = v & 3;
    const int p = p0 + pl, c = cb + 8 * k;
    if (p < HW && c < C) {
      float f[8];
      unpack8(ldg16(z + (size_t(n) * HW + p) * zc + g * C + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) zt[g][pl][8 * k + j] = f[j];
    }
  }
  __syncthreads();
  const int pl = threadIdx.x & 31, p = p0 + pl;
  for (int cl = threadIdx.x >> 5; cl < 32; cl += 8) {
    const int c = cb + cl;
    float h = 0.f;
    if (c < C && p < HW) {
      const size_t idx = size_t(c) * HW + p;
      const float cp = c_prev ? c_prev[size_t(n) * cp_sn + idx] : 0.f;
      const float gi = sigmoid_f(zt[0][pl][cl] + __ldg(w_ci + idx) * cp);
      const float gf = sigmoid_f(zt[1][pl][cl] + __ldg(w_cf + idx) * cp);
      const float zg = zt[2][pl][cl];
      const float cn = gf * cp + gi * (act == 4 ? tanhf(zg) : fmaxf(zg, 0.f));
      const float go = sigmoid_f(zt[3][pl][cl] + __ldg(w_co + idx) * cn);
      h = go * (act == 4 ? tanhf(cn) : fmaxf(cn, 0.f));
      c_out[size_t(n) * co_sn + idx] = cn;
      if (h_out) h_out[size_t(n) * h_sn + idx] = h;
    }
    ht[pl][cl] = h;
  }
  __syncthreads();
  if (h_nhwc && threadIdx.x < 128) {
    const int ql = threadIdx.x >> 2, k = threadIdx.x & 3;
    const int q = p0 + ql, c = cb + 8 * k;
    if (q < HW && c < C) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = ht[ql][8 * k + j];
      stg16(h_nhwc + (size_t(n) * HW + q) * hc + c, pack8(f));
    }
  }
}

// Backward of one ConvLSTM time step (the adjoint of convlstm_gates_kernel). The gates are recomputed from the saved
// z (bf16 NHWC), c_prev and c (fp32 NCHW); dh = dh_ext (fp32 NCHW: the gradient of this frame's output, image stride
// dhe_sn) + dh_rec (bf16 NHWC: d/dH_prev of step t+1, from the gate conv's input gradient); dc_in carries the cell
// gradient of step t+1 (NULL = 0):
//   dzo = dh*act(c)*o(1-o)          dc = dc_in + dh*o*act'(c) + dzo*W_co
//   dzi = dc*g*i(1-i)   dzf = dc*c_prev*f(1-f)   dzg = dc*i*act'(zg)   dc_prev = dc*f + dzi*W_ci + dzf*W_cf
//   dW_co += sum_n dzo*c   dW_ci += sum_n dzi*c_prev   dW_cf += sum_n dzf*c_prev
// A block owns 32 pixels x 32 channels and walks the batch itself, so the peephole gradients are plain
// read-modify-writes (deterministic, no atomics); dz leaves as bf16 NHWC [N][HW][zc] -- the operand of the gate
// conv's input-gradient and weight-gradient GEMMs.
__global__ void __launch_bounds__(256)
convlstm_gates_bwd_kernel(const __nv_bfloat16* __restrict__ z, int zc, const float* __restrict__ w_ci,
                          const float* __restrict__ w_cf, const float* __restrict__ w_co,
                          const float* __restrict__ c_prev, long long cp_sn, const float* __restrict__ c_cur,
                          long long cc_sn, const float* __restrict__ dh_ext, long long dhe_sn,
                          const __nv_bfloat16* __restrict__ dh_rec, int hc, const float* dc_in, float* dc_out,
                          __nv_bfloat16* __restrict__ dz, float* __restrict__ dw_ci, float* __restrict__ dw_cf,
                          float* __restrict__ dw_co, int N, int HW, int C, int act) {
  __shared__ float zt[4][32][33];
  __shared__ float ht[32][33];
  const int cb = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int pl = threadIdx.x & 31, p = p0 + pl;
  float aci[4] = {0.f, 0.f, 0.f, 0.f}, acf[4] = {0.f, 0.f, 0.f, 0.f}, aco[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = 0; n < N; ++n) {
    for (int v = threadIdx.x; v < 512; v += 256) {
      const int g = v >> 7, ql = (v >> 2) & 31, k = v & 3;
      const int q = p0 + ql, c = cb + 8 * k;
      if (q < HW && c < C) {
        float f[8];
        unpack8(ldg16(z + (size_t(n) * HW + q) * zc + g * C + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) zt[g][ql][8 * k + j] = f[j];
      }
    }
    if (threadIdx.x < 128) {
      const int ql = threadIdx.x >> 2, k = threadIdx.x & 3;
      const int q = p0 + ql, c = cb + 8 * k;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (dh_rec && q < HW && c < C) unpack8(ldg16(dh_rec + (size_t(n) * HW + q) * hc + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ht[ql][8 * k + j] = f[j];
    }
    __syncthreads();
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
      const int cl = (threadIdx.x >> 5) + 8 * slot;
      const int c = cb + cl;
      float dzi = 0.f, dzf = 0.f, dzg = 0.f, dzo = 0.f;
      if (c < C && p < HW) {
        const size_t idx = size_t(c) * HW + p;
        const float cp = c_prev ? c_prev[size_t(n) * cp_sn + idx] : 0.f;
        const float cn = c_cur[size_t(n) * cc_sn + idx];
        const float wci = __ldg(w_ci + idx), wcf = __ldg(w_cf + idx), wco = __ldg(w_co + idx);
        const float gi = sigmoid_f(zt[0][pl][cl] + wci * cp);
        const float gf = sigmoid_f(zt[1][pl][cl] + wcf * cp);
        const float zg = zt[2][pl][cl];
        const float gg = act == 4 ? tanhf(zg) : fmaxf(zg, 0.f);
        const float dgg = act == 4 ? 1.f - gg * gg : (zg > 0.f ? 1.f : 0.f);
        const float go = sigmoid_f(zt[3][pl][cl] + wco * cn);
        const float ac = act == 4 ? tanhf(cn) : fmaxf(cn, 0.f);
        const float dac = act == 4 ? 1.f - ac * ac : (cn > 0.f ? 1.f : 0.f);
        const float dh = (dh_ext ? dh_ext[size_t(n) * dhe_sn + idx] : 0.f) + ht[pl][cl];
        dzo = dh * ac * go * (1.f - go);
        const float dc = (dc_in ? dc_in[size_t(n) * C * HW + idx] : 0.f) + dh * go * dac + dzo * wco;
        dzi = dc * gg * gi * (1.f - gi);
        dzf = dc * cp * gf * (1.f - gf);
        dzg = dc * gi * dgg;
        dc_out[size_t(n) * C * HW + idx] = dc * gf + dzi * wci + dzf * wcf;
        aco[slot] += dzo * cn;
        aci[slot] += dzi * cp;
        acf[slot] += dzf * cp;
      }
      // the z tile is dead for this (pixel, channel): reuse it for dz
      zt[0][pl][cl] = dzi; zt[1][pl][cl] = dzf; zt[2][pl][cl] = dzg; zt[3][pl][cl] = dzo;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < 512; v += 256) {
      const int g = v >> 7, ql = (v >> 2) & 31, k = v & 3;
      const int q = p0 + ql, c = cb + 8 * k;
      if (q < HW && c < C) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = zt[g][ql][8 * k + j];
        stg16(dz + (size_t(n) * HW + q) * zc + g * C + c, pack8(f));
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int slot = 0; slot < 4; ++slot) {
    const int cl = (threadIdx.x >> 5) + 8 * slot;
    const int c = cb + cl;
    if (c < C && p < HW) {
      const size_t idx = size_t(c) * HW + p;
      dw_ci[idx] += aci[slot];
      dw_cf[idx] += acf[slot];
      dw_co[idx] += aco[slot];
    }
  }
}

// bf16 NHWC [N][HW][Cpad] -> fp32 NCHW (image stride out_sn, first C channels), transposed through shared memory;
// accumulate != 0 adds into the destination (the input gradient of a frame that feeds two cells)
__global__ void __launch_bounds__(256)
unpack_nhwc_tiled_kernel(const __nv_bfloat16* __restrict__ in, int Cpad, float* __restrict__ out, long long out_sn,
                         int C, int HW, int accumulate) {
  __shared__ float t[64][65];
  const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  for (int v = threadIdx.x; v < 512; v += 256) {
    const int pl = v >> 3, k = v & 7;
    const int p = p0 + pl, c = c0 + 8 * k;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p < HW && c < Cpad) unpack8(ldg16(in + (size_t(n) * HW + p) * Cpad + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[8 * k + j][pl] = f[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += 256) {
    const int cl = i >> 6, pl = i & 63;
    const int c = c0 + cl, p = p0 + pl;
    if (c < C && p < HW) {
      float* dst = out + size_t(n) * out_sn + size_t(c) * HW + p;
      *dst = accumulate ? *dst + t[cl][pl] : t[cl][pl];
    }
  }
}

// Gradient penalty: nsq[n] = sum_{pix, j<cj} (g[n,pix,c_off+j] + 1e-16)^2
__global__ void gp_normsq_kernel(const __nv_bfloat16* __restrict__ g, int HW, int C, int c_off, int cj,
                                 float* __restrict__ nsq) {
  __shared__ float sh[32];
  const int n = blockIdx.y;
  float acc = 0.f;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    const __nv_bfloat16* s = g + (size_t(n) * HW + pix) * C + c_off;
    for (int j = 0; j < cj; ++j) {
      const float v = __bfloat162float(s[j]) + 1e-16f;
      acc += v * v;
    }
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(nsq + n, tot);
}

// penalty = lambda * mean_n (sqrt(nsq) - constant)^2 ; coef[n] = lambda * 2/N * (norm - constant)/norm
__global__ void gp_finish_kernel(const float* __restrict__ nsq, int N, float lambda, float constant,
                                 float* __restrict__ loss, float* __restrict__ coef) {
  float acc = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float nm = sqrtf(nsq[n]);
    const float d = nm - constant;
    acc += d * d;
    coef[n] = nm > 0.f ? lambda * 2.f / float(N) * d / nm : 0.f;
  }
  __shared__ float sh[32];
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, lambda * tot / float(N));
}

// seed[n,pix,c_off+j] = coef[n] * (g[n,pix,c_off+j] + 1e-16); other channels of the 8-group zero.
__global__ void gp_seed_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ coef,
                               int N, int HW, int C, int c_off, int cj, __nv_bfloat16* __restrict__ seed) {
  const size_t total = size_t(N) * HW;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const int n = int(i / HW);
    const int g0 = (c_off / 8) * 8;
    float f[8], o[8];
    unpack8(ldg16(g + i * C + g0), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g0 + j;
      o[j] = (c >= c_off && c < c_off + cj) ? coef[n] * (f[j] + 1e-16f) : 0.f;
    }
    stg16(seed + i * C + g0, pack8(o));
  }
}

// ------------------------------------------------------------------ fused Adam + weight re-pack
// hyper != nullptr: {lr, beta1, beta2, eps, 1 - beta1^t, 1 - beta2^t, grad_scale} are read from device memory, so a
// captured CUDA graph of the whole step can be replayed with a new learning rate / step count (tg_adam_step_dev).
__global__ void adam_kernel(const AdamTensor* __restrict__ tab, int ntensors, float lr, float beta1,
                            float beta2, float eps, float bc1, float bc2, float grad_scale,
                            const float* __restrict__ hyper) {
  if (hyper) {
    lr = hyper[0]; beta1 = hyper[1]; beta2 = hyper[2]; eps = hyper[3]; bc1 = hyper[4]; bc2 = hyper[5];
    grad_scale = hyper[6];
  }
  const AdamTensor t = tab[blockIdx.y];
  const size_t numel = t.numel;
  const int taps = t.kh * t.kw;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < numel;
       i += size_t(gridDim.x) * blockDim.x) {
    // torch order: conv [O][I][kh][kw]; convT [I][O][kh][kw]; vectors: kind 0
    size_t gi = i;
    int o = 0, ic = 0, tap = 0;
    if (t.kind != 0) {
      tap = int(i % taps);
      const size_t r = i / taps;
      const int d1 = int(r % t.dim1), d0 = int(r / t.dim1);
      if (t.kind == 2) { ic = d0; o = d1; } else { o = d0; ic = d1; }
      if (t.kind == 4) ic = tap * t.dim1 + ic;   // im2col column of the thin first layer
      else
        for (int sgi = 0; sgi < t.nseg; ++sgi)
          if (ic < t.seg_end[sgi]) { ic += t.seg_shift[sgi]; break; }
      // gradient lives in the packed forward layout [tap][rows_pad][cols_pad] (kind 3: rows = taps, kind 4: one tap)
      gi = t.kind == 1   ? (size_t(tap) * t.o_pad + o) * t.i_pad + ic
           : t.kind == 2 ? (size_t(tap) * t.i_pad + ic) * t.o_pad + o
           : t.kind == 3 ? size_t(tap) * t.i_pad + ic
                         : size_t(o) * t.i_pad + ic;
    }
    float p = t.param[i];
    if (t.grad) {
      const float g = t.grad[gi] * grad_scale;
      float m = t.m[i], v = t.v[i];
      m = beta1 * m + (1.f - beta1) * g;
      v = beta2 * v + (1.f - beta2) * g * g;
      t.m[i] = m;
      t.v[i] = v;
      const float denom = sqrtf(v) / sqrtf(bc2) + eps;
      p -= (lr / bc1) * (m / denom);
      t.param[i] = p;
    }
    if (t.kind != 0) {
      const __nv_bfloat16 b = __float2bfloat16(p);
      if (t.pack_fwd) {
        // forward operand: [tap][rows][cols], cols contiguous = reduction channel of the forward op
        const size_t k = t.kind <= 2   ? (size_t(tap) * t.o_pad + o) * t.i_pad + ic
                         : t.kind == 3 ? size_t(tap) * t.i_pad + ic
                                       : size_t(o) * t.i_pad + ic;
        t.pack_fwd[k] = b;
      }
      if (t.pack_bwd) {
        // backward-data operand: transposed roles, taps mirrored for Conv2d (flip), same for convT
        const int btap = t.kind == 1 ? (taps - 1 - tap) : tap;
        const size_t k = t.kind <= 2   ? (size_t(btap) * t.i_pad + ic) * t.o_pad + o
                         : t.kind == 3 ? size_t(ic) * t.o_pad + tap
                                       : size_t(ic) * t.o_pad + o;
        t.pack_bwd[k] = b;
      }
    }
  }
}

}  // namespace tg

// ====================================================================== C-ABI launchers
using namespace tg;

static inline int grid_for(size_t work, int block, int cap = 148 * 16) {
  size_t g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > size_t(cap)) g = cap;
  return int(g);
}
#define TG_STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define TG_RET() return tg_check_launch(__func__)

// grad[o][c] += sum_r dw[r][o][c] over the replicas tg_fmap_bwd spread its atomics on; the replicas are cleared for
// the next backward pass (one launch instead of a fill, a reduction and an add)
__global__ void fmap_wgrad_fold_kernel(float* __restrict__ dw, int replicas, int co, int ci, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= co * 64) return;
  const int o = i >> 6, c = i & 63;
  float acc = 0.f;
  for (int r = 0; r < replicas; ++r) {
    acc += dw[(size_t(r) * co + o) * 64 + c];
    dw[(size_t(r) * co + o) * 64 + c] = 0.f;
  }
  if (c < ci && grad) grad[o * ci + c] += acc;
}

// dst[i] = v.f[i]: up to 16 scalars travel as kernel arguments (copied at launch time), so the host may overwrite its
// own copy immediately -- unlike an asynchronous memcpy from a reused pinned buffer
struct Floats16 { float f[16]; };
__global__ void write_floats_kernel(float* __restrict__ dst, Floats16 v, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = v.f[threadIdx.x];
}

// util.py:79-83: alpha ~ U[0,1) per sample; version 2 maps it to [0.5, 1). Writes alpha and 1 - alpha (the
// per-sample weights tg_im2col_pack mixes real_B and fake_B with).
__global__ void gp_alpha_kernel(const float* __restrict__ u, int version2, float* __restrict__ alpha,
                                float* __restrict__ one_minus, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = u[i];
  if (version2) a = (a + 1.f) / 2.f;
  alpha[i] = a;
  one_minus[i] = 1.f - a;
}

// ---------------------------------------------------------------------------------------------------------------
// One-shot all-reduce (sum) of a small fp32 buffer over NVLink peer memory -- the discriminator's 6 MB gradient arena,
// whose collective sits on the step's critical path (the G step needs the updated D). `peers[r]` is rank r's
// SYMMETRIC staging buffer (torch symmetric memory: cudaMalloc'd, IPC-mapped into every process of the node), two
// halves of `numel` floats used alternately so that a rank still reading call k is never overwritten by call k+1;
// `flags[r]` its flag pad, gridDim.x * world words. Per CTA, on its own slice: (1) copy local -> own staging half,
// (2) release-store this call's epoch into every peer's flag [cta][rank] and acquire-spin until every peer's flag for
// this CTA has arrived -- a per-slice barrier, no grid-wide or host synchronisation, (3) read the slice from ALL ranks'
// staging buffers over NVLink (16-byte volatile loads) and sum in rank order 0..world-1, so every rank computes
// bit-identical sums (the replicas' weights must stay identical), writing the result over the local buffer.
// (world-1) x numel x 4 bytes cross NVLink per rank: 44 MB at 8 GPUs, ~0.1 ms, against ~0.4 ms exposed for the
// same buffer through an NCCL all-reduce launched behind the step's own kernels.
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
constexpr int kP2PMaxWorld = 16;

__global__ void __launch_bounds__(512)
allreduce_oneshot_kernel(float* __restrict__ local, float* const* __restrict__ peers, uint32_t* const* __restrict__ flags,
                         int rank, int world, size_t numel, size_t half_stride, uint32_t epoch) {
  __shared__ float* sp[kP2PMaxWorld];
  if (threadIdx.x < world) sp[threadIdx.x] = peers[threadIdx.x] + size_t(epoch & 1u) * half_stride;
  __syncthreads();
  const size_t n4 = numel >> 2;
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t s0 = blockIdx.x * per, s1 = s0 + per < n4 ? s0 + per : n4;
  float4* mine = reinterpret_cast<float4*>(sp[rank]);
  float4* loc = reinterpret_cast<float4*>(local);
  for (size_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) mine[i] = loc[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    st_release_sys_u32(flags[threadIdx.x] + size_t(blockIdx.x) * world + rank, epoch);
    const uint32_t* f = flags[rank] + size_t(blockIdx.x) * world + threadIdx.x;
    while (int32_t(ld_acquire_sys_u32(f) - epoch) < 0) {
    }
  }
  __syncthreads();
  for (size_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = ld_volatile_f4(reinterpret_cast<const float4*>(sp[r]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    loc[i] = acc;
  }
}

// Streaming (cp.async.bulk ring) form of the same-resolution passes, tg_stream.cuh. TG_STREAM=0 routes everything
// back to the register-staged kernels (A/B runs).
static int g_stream_policy = -1;      // 0 never, 1 per-shape choice, 2 whenever the shape allows
static int stream_policy() {
  if (g_stream_policy < 0) {
    const char* e = getenv("TG_STREAM");
    g_stream_policy = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return g_stream_policy;
}
static bool stream_enabled() { return stream_policy() != 0; }
// Serpentine order (TG_SERP / tg_in_stream_serpentine, bit m = pass m walks its chunks from the tensor's end): the
// GEMMs walk images upwards, so the last tens of MB they wrote are what the 126 MB L2 still holds when the pass that
// consumes them starts; the apply pass then walks the other way than the statistics pass, for the same reason.
static int g_serp = -1;
static int serp_mask() {
  if (g_serp < 0) {
    const char* e = getenv("TG_SERP");
    g_serp = (e && e[0] >= '0' && e[0] <= '7') ? e[0] - '0' : 0;   // measured neutral in the step: off
  }
  return g_serp;
}
// Slim form (tg_in_stream_slim): the engine switches it on for the backward passes it issues while a weight-gradient
// GEMM runs on the side stream. A persistent wgrad CTA (191-194 KiB of shared memory, 25 K registers) leaves ~33 KiB
// and 40 K registers on its SM: one 288-thread CTA of this kernel with a ring of 4 KiB chunks fits beside it, so the
// bandwidth-bound pass and the tensor-bound GEMM share the SM instead of taking turns (the 108 KiB ring cannot).
static int g_slim = 0;

static int slim_budget() {
  static int kb = -1;
  if (kb < 0) {
    const char* e = getenv("TG_SLIM_KB");
    kb = e ? atoi(e) : 32;
    if (kb < 16 || kb > 108) kb = 32;
  }
  return kb * 1024;
}
template <int MODE, int PPT, bool G2, bool POOL>
static int launch_stream_cfg(tg::StreamArgs a, cudaStream_t s, bool slim) {
  using namespace tg;
  constexpr int nin = MODE == 0 ? 1 : 2 + (G2 ? 1 : 0) + (POOL ? 1 : 0);
  const int CG = a.C >> 3, PL = kStreamConsumers / CG;
  const size_t chunk = size_t(stream_chunk_bytes(PPT));
  const size_t scratch = MODE == 1 ? size_t(slim ? 1 : PL) * a.C * 2 * sizeof(float) : 0;
  // two CTAs per SM: <= ~110 KiB each; slim: what a persistent GEMM CTA leaves
  const size_t budget = slim ? size_t(slim_budget()) : size_t(108 * 1024);
  if (scratch + 256 >= budget) return tg_set_error("in_stream: shared memory budget");
  int stages = int((budget - scratch - 256) / (size_t(nin) * chunk));
  if (stages > 8) stages = 8;
  if (stages < 2) return tg_set_error("in_stream: shared memory budget");
  a.stages = stages;
  a.slim = slim ? 1 : 0;
  a.rev = (serp_mask() >> MODE) & 1;
  const size_t smem = size_t(stages) * nin * chunk + scratch + 16 * stages;
  auto kern = in_stream_kernel<MODE, PPT, G2, POOL>;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return tg_set_error("in_stream: cudaFuncSetAttribute");
    // the same (maximum) shared-memory carve-out as the GEMM kernels: a CTA can only join an SM whose L1 / shared
    // split already matches, and the slim form is meant to run beside a resident GEMM CTA
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, int(cudaSharedmemCarveoutMaxShared));
    configured = smem;
  }
  const int CP = PPT * PL;
  const long long chunks = (long long)a.N * ((a.HW + CP - 1) / CP);
  long long grid = slim ? 148 : 2 * 148;
  if (grid > chunks) grid = chunks;
  if (grid < 1) grid = 1;
  if (tg_launch(kern, dim3(unsigned(grid)), dim3(kStreamThreads), smem, s, a) != cudaSuccess)
    return tg_check_launch("in_stream_kernel") ? -1 : tg_set_error("in_stream_kernel: launch failed");
  return tg_check_launch("in_stream_kernel");
}
template <int MODE, int PPT>
static int launch_stream_ppt(tg::StreamArgs a, cudaStream_t s, bool slim) {
  if constexpr (MODE == 0) {
    return launch_stream_cfg<MODE, PPT, false, false>(a, s, slim);
  } else {
    if (!a.in1) return tg_set_error("in_stream: the backward passes need a same-resolution gradient route");
    if (a.in2 && a.pool) return launch_stream_cfg<MODE, PPT, true, true>(a, s, slim);
    if (a.in2) return launch_stream_cfg<MODE, PPT, true, false>(a, s, slim);
    if (a.pool) return launch_stream_cfg<MODE, PPT, false, true>(a, s, slim);
    return launch_stream_cfg<MODE, PPT, false, false>(a, s, slim);
  }
}
static bool stream_pool_ok(int H, int W, int C, int ppt);   // defined below
// slim only where it fits and the chunk geometry allows (backward passes; the pooled route needs even chunks)
static bool slim_applies(int mode, const tg::StreamArgs& a) {
  if (!g_slim || mode == 0) return false;
  const int nin = 1 + (a.in1 ? 1 : 0) + (a.in2 ? 1 : 0) + (a.pool ? 1 : 0);
  const size_t scratch = mode == 1 ? size_t(a.C) * 2 * sizeof(float) : 0;
  if (scratch + 256 + 2 * size_t(nin) * tg::stream_chunk_bytes(1) > size_t(slim_budget())) return false;
  if (a.pool && !stream_pool_ok(a.HW / a.W, a.W, a.C, 1)) return false;
  return true;
}
template <int MODE>
static int launch_stream(tg::StreamArgs a, cudaStream_t s) {
  if constexpr (MODE != 0) {
    if (slim_applies(MODE, a)) return launch_stream_ppt<MODE, 1>(a, s, true);
  }
  return launch_stream_ppt<MODE, tg::kStreamPPT>(a, s, false);
}
static bool stream_shape_ok(int C) { return C >= 64 && C <= 2048 && (C & 63) == 0; }
// the average-pool gradient route streams when a chunk (2 * (256 / (C/8)) pixels) lies inside one image row
static bool stream_pool_ok(int H, int W, int C, int ppt = tg::kStreamPPT) {
  const int CP = ppt * (tg::kStreamConsumers / (C >> 3));
  return CP >= 2 && (CP & 1) == 0 && W % CP == 0 && (H & 1) == 0 && (W & 1) == 0;
}
// Which form is faster, measured per shape (profiles/r02_tail_microbench_*.txt, tools/tail_bench.py): since the
// ring's per-element work became branch-free it wins every backward pass (statistics 0.95-0.97 vs 0.72 of the copy
// bandwidth on the 268 MB tensors, apply 0.85 vs 0.72-0.88). The forward pass -- one read and one write stream --
// stays register-staged on the wide maps with few channels (C <= 128 and >= 100 MB: 94.5 vs 107 us at 268 MB, 52.5
// vs 56 us at 134 MB; 1184 short-lived CTAs walk the tensor as one moving window, 296 persistent ones as 296
// separate streams) and under ~20 MB, where it starts up ~1.5 us faster.
static bool stream_wins(int mode, int N, int HW, int C) {
  if (stream_policy() == 2) return true;
  const double bytes = 2.0 * N * double(HW) * C;
  if (mode == 0 && C <= 128 && bytes >= 100e6) return false;
  if (mode == 0 && bytes <= 20e6) return false;
  return true;
}

extern "C" {

int tg_pack_nchw(const float* A, const float* B, const float* wa, const float* wb, void* out, int N,
                 int HW, int cj, int C, int c_off, void* stream) {
  if (cj > 8 - (c_off & 7)) return tg_set_error("tg_pack_nchw: channels must stay inside one 8-group");
  pack_nchw_kernel<<<grid_for(size_t(N) * HW, 256), 256, 0, TG_STREAM(stream)>>>(
      A, B, wa, wb, (__nv_bfloat16*)out, N, HW, cj, C, c_off);
  TG_RET();
}

int tg_unpack_nhwc(const void* in, float* out, int N, int HW, int C, int c_off, int cj, float scale,
                   void* stream) {
  unpack_nhwc_kernel<<<grid_for(size_t(N) * HW, 256), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)in, out, N, HW, C, c_off, cj, scale);
  TG_RET();
}

int tg_im2col_pack(const float* A, const float* B, const float* B2, const float* wa, const float* wb, void* cols,
                   int N, int ca, int cb, int H, int W, int k, int stride, void* stream) {
  if (k * k * (ca + cb) > 64) return tg_set_error("tg_im2col_pack: k*k*(ca+cb) must be <= 64");
  if (!B) return tg_set_error("tg_im2col_pack: B is required");
  const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
  if (ca > 8 || cb > 8 || k > 4) return tg_set_error("tg_im2col_pack: at most 8 + 8 channels and 4x4 taps");
  const int grid = grid_for(size_t(N) * Ho * Wo, 128, 148 * 16);
  if (k == 3 && ca == 3 && cb == 3)
    im2col_pack_kernel<3, 3, 3><<<grid, 128, 0, TG_STREAM(stream)>>>(A, B, B2, wa, wb, (__nv_bfloat16*)cols, N, ca, cb,
                                                                      H, W, Ho, Wo, k, stride);
  else
    im2col_pack_kernel<0, 0, 0><<<grid, 128, 0, TG_STREAM(stream)>>>(A, B, B2, wa, wb, (__nv_bfloat16*)cols, N, ca, cb,
                                                                      H, W, Ho, Wo, k, stride);
  TG_RET();
}

int tg_col2im_grad(const void* dcols, float* g, int N, int ct, int c_off, int cj, int H, int W, int k, int stride,
                   float scale, void* stream) {
  if (cj > 8 || k * k * ct > 64) return tg_set_error("tg_col2im_grad: at most 8 channels, k*k*ct <= 64");
  const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
  col2im_grad_kernel<<<grid_for(size_t(N) * H * W, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)dcols, g, N, ct, c_off, cj, H, W, Ho, Wo, k, stride, scale);
  TG_RET();
}

int tg_head_gather(const void* hc, const float* bias, void* out, int N, int Hi, int Wi, int k, int C, int act,
                   float slope, void* stream) {
  if (k * k > 16) return tg_set_error("tg_head_gather: at most 16 taps");
  const int Ho = Hi - k + 1, Wo = Wi - k + 1;
  head_gather_kernel<<<grid_for(size_t(N) * Ho * Wo, 256, 148 * 8), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)hc, bias, (__nv_bfloat16*)out, N, Hi, Wi, Ho, Wo, k, C, act, slope);
  TG_RET();
}

int tg_head_scatter(const void* dz, void* dzc, int N, int Hi, int Wi, int k, int C, void* stream) {
  if (k * k > 16) return tg_set_error("tg_head_scatter: at most 16 taps");
  const int Ho = Hi - k + 1, Wo = Wi - k + 1;
  head_scatter_kernel<<<grid_for(size_t(N) * Hi * Wi, 256, 148 * 8), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)dz, (__nv_bfloat16*)dzc, N, Hi, Wi, Ho, Wo, k, C);
  TG_RET();
}

int tg_gp_normsq_img(const float* g, int N, long long per_img, float* nsq, void* stream) {
  dim3 grid(grid_for(size_t(per_img), 256, 64), N);
  gp_normsq_img_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>(g, size_t(per_img), nsq);
  TG_RET();
}

int tg_in_stream_policy(int policy) {
  const int prev = stream_policy();
  if (policy >= 0 && policy <= 2) g_stream_policy = policy;
  return prev;
}

int tg_in_stream_slim(int on) {
  const int prev = g_slim;
  if (on == 0 || on == 1) g_slim = on;
  return prev;
}

int tg_in_stream_serpentine(int mask) {
  const int prev = serp_mask();
  if (mask >= 0 && mask <= 7) g_serp = mask;
  return prev;
}

int tg_in_finalize(const float* partial, float* mr, int N, int T, int C, int count, float eps,
                   void* stream) {
  dim3 grid(C / 16, N);
  tg_launch(in_finalize_kernel, grid, dim3(1024), 0, TG_STREAM(stream), partial, (float*)mr, T, C, 1.f / float(count), eps);
  TG_RET();
}

int tg_in_stats_direct(const void* raw, float* mr, int N, int HW, int C, float eps, void* stream) {
  dim3 grid(C / 64, N);
  in_stats_direct_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>((const __nv_bfloat16*)raw, mr, HW, C, eps);
  TG_RET();
}

static inline int strip_block(int C) { return (C >> 3) >= 256 ? (C >> 3) : 256; }
static inline int strip_count(int units, int C, int N, int per_thread) {
  const int PL = strip_block(C) / (C >> 3);
  int strips = (units + PL * per_thread - 1) / (PL * per_thread);
  const int cap = (148 * 8 + N - 1) / N;
  if (strips > cap) strips = cap;
  return strips < 1 ? 1 : strips;
}

int tg_in_act_fwd(const void* raw, const float* mr, const float* gamma, const float* beta, void* y,
                  void* pool, int pool_mode, void* up, int N, int H, int W, int C, int c_valid, int act,
                  float slope, void* stream) {
  const __nv_bfloat16* r = (const __nv_bfloat16*)raw;
  __nv_bfloat16 *yy = (__nv_bfloat16*)y, *pp = (__nv_bfloat16*)pool, *uu = (__nv_bfloat16*)up;
  cudaStream_t s = TG_STREAM(stream);
  const bool quad = pool || up;
  if (quad && ((H | W) & 1)) return tg_set_error("tg_in_act_fwd: pool/upsample need even H, W");
  if (C > 8192) return tg_set_error("tg_in_act_fwd: C too large");
  if (!quad && raw && mr && stream_enabled() && stream_shape_ok(C) && stream_wins(0, N, H * W, C)) {
    tg::StreamArgs a{};
    a.in0 = r; a.out = yy; a.mr = mr; a.gamma = gamma; a.beta = beta;
    a.N = N; a.HW = H * W; a.C = C; a.c_valid = c_valid; a.act = act; a.slope = slope;
    return launch_stream<0>(a, s);
  }
  const int units = quad ? (H / 2) * (W / 2) : H * W;
  dim3 grid(strip_count(units, C, N, quad ? 4 : 16), N);
  const int block = strip_block(C);
  const int pm = pool ? pool_mode : 0;
  const int rev = serp_mask() & 1;
#define LAUNCH(P, U) tg_launch(in_act_fwd_kernel<P, U>, grid, dim3(block), 0, s, r, mr, gamma, beta, yy, pp, uu, H, W, C, c_valid, act, slope, rev)
  if (pm == 0 && !up) LAUNCH(0, false);
  else if (pm == 0 && up) LAUNCH(0, true);
  else if (pm == 1 && !up) LAUNCH(1, false);
  else if (pm == 1 && up) LAUNCH(1, true);
  else if (pm == 2 && !up) LAUNCH(2, false);
  else LAUNCH(2, true);
#undef LAUNCH
  TG_RET();
}

int tg_in_bwd_reduce(const void* raw, const void* y, const float* mr, const float* gamma,
                     const float* beta, const void* g_same, const void* g_pool, int pool_mode,
                     const void* g_up, int g_up_pooled, void* dn, float* red, int N, int H, int W, int C,
                     int c_valid, int act, float slope, void* stream) {
  InBwdArgs a;
  a.rev = (serp_mask() >> 1) & 1;
  a.up_pooled = g_up_pooled;
  a.dz = nullptr; a.dgamma = nullptr; a.dbeta = nullptr;
  a.raw = (const __nv_bfloat16*)raw; a.y = (const __nv_bfloat16*)y; a.mr = mr; a.gamma = gamma;
  a.beta = beta; a.g_same = (const __nv_bfloat16*)g_same; a.g_pool = (const __nv_bfloat16*)g_pool;
  a.g_up = (const __nv_bfloat16*)g_up; a.dn = (__nv_bfloat16*)dn; a.red = red;
  a.N = N; a.H = H; a.W = W; a.C = C; a.c_valid = c_valid; a.act = act; a.pool_mode = pool_mode; a.slope = slope;
  if (red && !mr) return tg_set_error("tg_in_bwd_reduce: reductions need the (mean, rstd) table");
  const int block = strip_block(C);
  if (block > 1024) return tg_set_error("tg_in_bwd_reduce: C too large");
  const int PL = block / (C / 8);
  const size_t smem = red ? size_t(PL) * C * 2 * sizeof(float) : 0;
  dim3 grid(strip_count(H * W, C, N, 16), N);
  if (!raw && !dn) return tg_set_error("tg_in_bwd_reduce: a layer without norm needs the dn (= dz) output");
  const bool plain = !g_pool && (g_same || g_up) && (!g_up || g_up_pooled);
  // every route at the unit's own resolution, plus (optionally) the gradient of its 2x2 average-pooled copy
  const bool streamable = (plain || (g_pool && pool_mode == 1 && (g_same || g_up) && (!g_up || g_up_pooled) &&
                                     stream_pool_ok(H, W, C))) &&
                          raw && mr && stream_enabled() && stream_shape_ok(C);
  if (streamable && red && (g_pool || stream_wins(1, N, H * W, C))) {
    tg::StreamArgs sa{};
    sa.in0 = a.raw; sa.in1 = a.g_same ? a.g_same : a.g_up; sa.in2 = (a.g_same && a.g_up) ? a.g_up : nullptr;
    sa.pool = a.g_pool; sa.W = W;
    sa.out = a.dn; sa.mr = mr; sa.gamma = gamma; sa.beta = beta; sa.red = red;
    sa.N = N; sa.HW = H * W; sa.C = C; sa.c_valid = c_valid; sa.act = act; sa.slope = slope;
    return launch_stream<1>(sa, TG_STREAM(stream));
  }
  if (plain) tg_launch(in_bwd_reduce_kernel<true, 0>, grid, dim3(block), smem, TG_STREAM(stream), a);
  else tg_launch(in_bwd_reduce_kernel<false, 0>, grid, dim3(block), smem, TG_STREAM(stream), a);
  TG_RET();
}

// Second pass of the InstanceNorm backward without a stored dn: re-reads raw and the gradient routes, recomputes
// dn and writes dz (see in_bwd_reduce_kernel PASS 1). red must hold the sums of tg_in_bwd_reduce on the same inputs.
int tg_in_bwd_apply_re(const void* raw, const void* y, const float* mr, const float* gamma, const float* beta,
                       const void* g_same, const void* g_pool, int pool_mode, const void* g_up, int g_up_pooled,
                       const float* red, void* dz, int N, int H, int W, int C, int c_valid, int act, float slope,
                       float* dgamma, float* dbeta, void* stream) {
  if (!raw || !mr || !red || !dz) return tg_set_error("tg_in_bwd_apply_re: null argument");
  InBwdArgs a;
  a.rev = (serp_mask() >> 2) & 1;
  a.raw = (const __nv_bfloat16*)raw; a.y = (const __nv_bfloat16*)y; a.mr = mr; a.gamma = gamma;
  a.beta = beta; a.g_same = (const __nv_bfloat16*)g_same; a.g_pool = (const __nv_bfloat16*)g_pool;
  a.g_up = (const __nv_bfloat16*)g_up; a.dn = nullptr; a.red = const_cast<float*>(red);
  a.N = N; a.H = H; a.W = W; a.C = C; a.c_valid = c_valid; a.act = act; a.pool_mode = pool_mode; a.slope = slope;
  a.up_pooled = g_up_pooled;
  a.dz = (__nv_bfloat16*)dz; a.dgamma = dgamma; a.dbeta = dbeta;
  const int block = strip_block(C);
  if (block > 1024) return tg_set_error("tg_in_bwd_apply_re: C too large");
  dim3 grid(strip_count(H * W, C, N, 16), N);
  const bool plain = !g_pool && (g_same || g_up) && (!g_up || g_up_pooled);
  const bool streamable = (plain || (g_pool && pool_mode == 1 && (g_same || g_up) && (!g_up || g_up_pooled) &&
                                     stream_pool_ok(H, W, C))) &&
                          stream_enabled() && stream_shape_ok(C);
  if (streamable && (g_pool || stream_wins(2, N, H * W, C))) {
    tg::StreamArgs sa{};
    sa.in0 = a.raw; sa.in1 = a.g_same ? a.g_same : a.g_up; sa.in2 = (a.g_same && a.g_up) ? a.g_up : nullptr;
    sa.pool = a.g_pool; sa.W = W;
    sa.out = a.dz; sa.mr = mr; sa.gamma = gamma; sa.beta = beta; sa.red = a.red; sa.dgamma = dgamma; sa.dbeta = dbeta;
    sa.N = N; sa.HW = H * W; sa.C = C; sa.c_valid = c_valid; sa.act = act; sa.slope = slope;
    return launch_stream<2>(sa, TG_STREAM(stream));
  }
  if (plain) tg_launch(in_bwd_reduce_kernel<true, 1>, grid, dim3(block), 0, TG_STREAM(stream), a);
  else tg_launch(in_bwd_reduce_kernel<false, 1>, grid, dim3(block), 0, TG_STREAM(stream), a);
  TG_RET();
}

int tg_in_bwd_apply(const void* dn, const void* raw, const float* mr, const float* gamma,
                    const float* red, void* dz, int N, int HW, int C, int c_valid, float* dgamma, float* dbeta,
                    void* stream) {
  dim3 grid(strip_count(HW, C, N, 16), N);
  in_bwd_apply_kernel<<<grid, strip_block(C), 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)dn, (const __nv_bfloat16*)raw, mr, gamma, red, (__nv_bfloat16*)dz, HW, C, c_valid,
      dgamma, dbeta);
  TG_RET();
}

int tg_affine_grad(const float* red, float* dgamma, float* dbeta, int N, int C, int c_valid, void* stream) {
  affine_grad_kernel<<<(C + 127) / 128, 128, 0, TG_STREAM(stream)>>>(red, dgamma, dbeta, N, C, c_valid);
  TG_RET();
}

int tg_bias_grad(const void* dz, float* db, long long rows, int C, int c_valid, void* stream) {
  int strips = int((rows + 2047) / 2048);
  if (strips > 148 * 4) strips = 148 * 4;
  if (strips < 1) strips = 1;
  dim3 grid(C / 64, strips);
  bias_grad_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>((const __nv_bfloat16*)dz, db, size_t(rows), C, c_valid);
  TG_RET();
}

int tg_in_bwd2(const void* u, const void* raw, const void* dn, const float* mr, const float* gamma,
               const float* beta, const float* red1, float* red2, void* adj_da, void* adj_z, int N,
               int HW, int C, int c_valid, int act, float slope, void* stream) {
  const int CG = C / 8;
  const int block = CG >= 256 ? CG : 256;
  if (block > 1024) return tg_set_error("tg_in_bwd2: C too large");
  const int PL = block / CG;
  int strips = (HW + PL * 8 - 1) / (PL * 8);
  const int cap = (148 * 8 + N - 1) / N;
  if (strips > cap) strips = cap;
  if (strips < 1) strips = 1;
  dim3 grid(strips, N);
  cudaStream_t s = TG_STREAM(stream);
  in_bwd2_reduce_kernel<<<grid, block, size_t(PL) * C * 3 * sizeof(float), s>>>(
      (const __nv_bfloat16*)u, (const __nv_bfloat16*)raw, (const __nv_bfloat16*)dn, mr, red2, HW, C);
  in_bwd2_apply_kernel<<<grid, block, size_t(PL) * C * sizeof(float), s>>>(
      (const __nv_bfloat16*)u, (const __nv_bfloat16*)raw, (const __nv_bfloat16*)dn, mr, gamma, beta, red1,
      red2, (__nv_bfloat16*)adj_da, (__nv_bfloat16*)adj_z, HW, C, c_valid, act, slope);
  TG_RET();
}

int tg_gamma_grad2(const float* red2, float* dgamma, int N, int C, int c_valid, void* stream) {
  gamma_grad2_kernel<<<(C + 127) / 128, 128, 0, TG_STREAM(stream)>>>(red2, dgamma, N, C, c_valid);
  TG_RET();
}

int tg_add(const void* a, const void* b, void* out, long long numel, void* stream) {
  add_kernel<<<grid_for(size_t(numel) / 8, 256, 148 * 32), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, size_t(numel) / 8);
  TG_RET();
}

int tg_act_bwd(const void* g, const void* y, void* out, long long numel, int act, float slope,
               void* stream) {
  act_bwd_kernel<<<grid_for(size_t(numel) / 8, 256, 148 * 32), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)g, (const __nv_bfloat16*)y, (__nv_bfloat16*)out, size_t(numel) / 8, act, slope);
  TG_RET();
}

int tg_fmap_fwd(const void* x, const float* w, const float* b, float* out, int N, int HW, int C, int co,
                int use_tanh, int ci, void* stream) {
  if (C != 64 || co > 4 || ci > C || ci < 1) return tg_set_error("tg_fmap_fwd: expects C == 64, co <= 4, ci <= C");
  if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    fmap_fwd64_kernel<<<grid_for(size_t(N) * HW / 4, 32, 148 * 8), 256, 0, TG_STREAM(stream)>>>(
        (const __nv_bfloat16*)x, w, b, out, N, HW, co, use_tanh, ci);
  else
    fmap_fwd_kernel<<<grid_for(size_t(N) * HW, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
        (const __nv_bfloat16*)x, w, b, out, N, HW, C, co, use_tanh, ci);
  TG_RET();
}

int tg_fmap_bwd(const void* x, const float* w, const float* out, const float* g1, const float* g2,
                void* dx, float* dw, int dw_replicas, float* db, int N, int HW, int C, int co, int use_tanh, int ci,
                void* stream) {
  if (dw_replicas < 1) return tg_set_error("tg_fmap_bwd: dw_replicas >= 1");
  if (C != 64 || co > 4 || ci > C || ci < 1) return tg_set_error("tg_fmap_bwd: expects C == 64, co <= 4, ci <= C");
  int strips = (HW + 1023) / 1024;                 // >= 32 pixels per lane per block
  const int cap = (148 * 8 + N - 1) / N;
  if (strips > cap) strips = cap;
  if (strips < 1) strips = 1;
  fmap_bwd_kernel<<<dim3(strips, N), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)x, w, out, g1, g2, (__nv_bfloat16*)dx, dw, db, N, HW, C, co, use_tanh, dw_replicas, ci);
  TG_RET();
}

int tg_fmap_wgrad_fold(float* dw, int dw_replicas, int co, int ci, float* grad, void* stream) {
  if (co > 4 || ci > 64) return tg_set_error("tg_fmap_wgrad_fold: co <= 4, ci <= 64");
  fmap_wgrad_fold_kernel<<<(co * 64 + 127) / 128, 128, 0, TG_STREAM(stream)>>>(dw, dw_replicas, co, ci, grad);
  TG_RET();
}

int tg_allreduce_oneshot(float* local, const void* peer_bufs_dev, const void* peer_flags_dev, int rank, int world,
                          long long numel, long long half_stride, unsigned epoch, int ctas, void* stream) {
  if (world < 1 || world > kP2PMaxWorld || rank < 0 || rank >= world) return tg_set_error("tg_allreduce_oneshot: world");
  if ((numel & 3) || (half_stride & 3) || numel > half_stride)
    return tg_set_error("tg_allreduce_oneshot: numel, half_stride multiples of 4, numel <= half_stride");
  if (ctas < 1 || epoch == 0) return tg_set_error("tg_allreduce_oneshot: ctas >= 1, epoch >= 1");
  allreduce_oneshot_kernel<<<ctas, 512, 0, TG_STREAM(stream)>>>(local, (float* const*)peer_bufs_dev,
                                                                 (uint32_t* const*)peer_flags_dev, rank, world,
                                                                 size_t(numel), size_t(half_stride), epoch);
  TG_RET();
}

int tg_gp_alpha(const float* u, int version, float* alpha, float* one_minus_alpha, int n, void* stream) {
  gp_alpha_kernel<<<(n + 127) / 128, 128, 0, TG_STREAM(stream)>>>(u, version == 2, alpha, one_minus_alpha, n);
  TG_RET();
}

int tg_gan_loss(const void* pred, const float* label, float label_const, int mode, int target_is_real,
                int for_disc, int has_sigmoid, float scale, int n0, int n1, int HW, int C, float* loss,
                void* dz, void* stream) {
  gan_loss_kernel<<<grid_for(size_t(n1 - n0) * HW, 256, 148 * 4), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)pred, label, label_const, mode, target_is_real, for_disc, has_sigmoid, scale,
      n0, n1, HW, C, loss, (__nv_bfloat16*)dz, dz != nullptr);
  TG_RET();
}

int tg_gp_first_seed(const void* pred, int has_sigmoid, int n0, int n1, int HW, int C, void* dz,
                     void* stream) {
  gp_first_seed_kernel<<<grid_for(size_t(n1 - n0) * HW, 256, 148 * 4), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)pred, has_sigmoid, n0, n1, HW, C, (__nv_bfloat16*)dz);
  TG_RET();
}

int tg_gp_top(const void* w, const void* pred, int has_sigmoid, long long numel, int C, void* out,
              void* stream) {
  gp_top_kernel<<<grid_for(size_t(numel), 256, 148 * 4), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)w, (const __nv_bfloat16*)pred, has_sigmoid, size_t(numel), C, (__nv_bfloat16*)out);
  TG_RET();
}

int tg_l1_loss(const float* a, const float* b, long long numel, float scale, float* loss, float* grad_a,
               void* stream) {
  l1_loss_kernel<<<grid_for(size_t(numel), 256, 148 * 8), 256, 0, TG_STREAM(stream)>>>(
      a, b, size_t(numel), scale, loss, grad_a);
  TG_RET();
}

int tg_feat_loss(const void* a, const void* b, long long numel, float weight, int l2, float* loss,
                 void* stream) {
  feat_loss_kernel<<<grid_for(size_t(numel) / 8, 256, 148 * 8), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, size_t(numel) / 8, weight, l2, loss);
  TG_RET();
}

int tg_feat_loss_grad(const void* a, const void* b, long long numel, float weight, void* g, void* stream) {
  feat_loss_grad_kernel<<<grid_for(size_t(numel) / 8, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, size_t(numel) / 8, weight, (__nv_bfloat16*)g);
  TG_RET();
}

int tg_pool_fwd(const void* y, void* pool, int N, int H, int W, int C, int mode, void* stream) {
  if ((H | W) & 1) return tg_set_error("tg_pool_fwd: H and W must be even");
  if (mode != 1 && mode != 2) return tg_set_error("tg_pool_fwd: mode 1 (avg) or 2 (max)");
  pool_fwd_kernel<<<grid_for(size_t(N) * (H / 2) * (W / 2) * (C / 8), 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)y, (__nv_bfloat16*)pool, N, H, W, C, mode);
  TG_RET();
}

int tg_vgg_prep_fwd(const float* x, void* out, int N, int cs, int H, int W, int OH, int OW, int C, int resize,
                    void* stream) {
  if (cs != 1 && cs != 3) return tg_set_error("tg_vgg_prep: 1 or 3 source channels");
  vgg_prep_fwd_kernel<<<grid_for(size_t(N) * OH * OW, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      x, (__nv_bfloat16*)out, N, cs, H, W, OH, OW, C, resize);
  TG_RET();
}

int tg_vgg_prep_bwd(const void* g, float* grad, int N, int cs, int H, int W, int OH, int OW, int C, int resize,
                    float scale, void* stream) {
  if (cs != 1 && cs != 3) return tg_set_error("tg_vgg_prep: 1 or 3 source channels");
  vgg_prep_bwd_kernel<<<grid_for(size_t(N) * OH * OW, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)g, grad, N, cs, H, W, OH, OW, C, resize, scale);
  TG_RET();
}

int tg_augment_pair(const void* img_u8, const void* mask_u8, const long long* params, float* out_a, float* out_b,
                    int N, int H, int W, int ca, int cb, void* stream) {
  augment_pair_kernel<<<grid_for(size_t(N) * H * W, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      (const uint8_t*)img_u8, (const uint8_t*)mask_u8, params, out_a, out_b, N, H, W, ca, cb);
  TG_RET();
}

int tg_eval_fuzzy(const float* out, const float* real, int N, long long per_img, float* stats, void* stream) {
  dim3 grid(grid_for(size_t(per_img), 256, 32), N);
  eval_fuzzy_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>(out, real, size_t(per_img), stats);
  TG_RET();
}

int tg_pack_nchw_tiled(const float* in, long long in_stride_n, void* out, int N, int C, int HW, int Cpad,
                       void* stream) {
  if (Cpad % 8 || C > Cpad) return tg_set_error("tg_pack_nchw_tiled: Cpad must be a multiple of 8 and >= C");
  dim3 grid((HW + 63) / 64, (Cpad + 63) / 64, N);
  pack_nchw_tiled_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>(in, in_stride_n, (__nv_bfloat16*)out, C, HW, Cpad);
  TG_RET();
}

int tg_convlstm_gates(const void* z, int zc, const float* w_ci, const float* w_cf, const float* w_co,
                      const float* c_prev, long long c_prev_stride_n, float* c_out, long long c_out_stride_n,
                      float* h_out, long long h_out_stride_n, void* h_nhwc, int hc, int N, int HW, int C, int act,
                      void* stream) {
  if (C % 8 || zc < 4 * C || zc % 8) return tg_set_error("tg_convlstm_gates: C % 8 == 0 and zc >= 4*C required");
  if (h_nhwc && (hc < C || hc % 8)) return tg_set_error("tg_convlstm_gates: hc must be a multiple of 8 and >= C");
  if (act != 3 && act != 4) return tg_set_error("tg_convlstm_gates: activation must be relu (3) or tanh (4)");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  convlstm_gates_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)z, zc, w_ci, w_cf, w_co, c_prev, c_prev_stride_n, c_out, c_out_stride_n, h_out,
      h_out_stride_n, (__nv_bfloat16*)h_nhwc, hc, HW, C, act);
  TG_RET();
}

int tg_convlstm_gates_bwd(const void* z, int zc, const float* w_ci, const float* w_cf, const float* w_co,
                          const float* c_prev, long long c_prev_stride_n, const float* c_cur, long long c_cur_stride_n,
                          const float* dh_ext, long long dh_ext_stride_n, const void* dh_rec, int hc,
                          const float* dc_in, float* dc_out, void* dz, float* dw_ci, float* dw_cf, float* dw_co,
                          int N, int HW, int C, int act, void* stream) {
  if (C % 8 || zc < 4 * C || zc % 8) return tg_set_error("tg_convlstm_gates_bwd: C % 8 == 0 and zc >= 4*C required");
  if (dh_rec && (hc < C || hc % 8)) return tg_set_error("tg_convlstm_gates_bwd: hc must be a multiple of 8 and >= C");
  if (act != 3 && act != 4) return tg_set_error("tg_convlstm_gates_bwd: activation must be relu (3) or tanh (4)");
  if (!z || !c_cur || !dc_out || !dz || !dw_ci || !dw_cf || !dw_co)
    return tg_set_error("tg_convlstm_gates_bwd: null argument");
  dim3 grid((HW + 31) / 32, (C + 31) / 32);
  convlstm_gates_bwd_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)z, zc, w_ci, w_cf, w_co, c_prev, c_prev_stride_n, c_cur, c_cur_stride_n, dh_ext,
      dh_ext_stride_n, (const __nv_bfloat16*)dh_rec, hc, dc_in, dc_out, (__nv_bfloat16*)dz, dw_ci, dw_cf, dw_co, N, HW,
      C, act);
  TG_RET();
}

int tg_unpack_nhwc_tiled(const void* in, int Cpad, float* out, long long out_stride_n, int N, int C, int HW,
                         int accumulate, void* stream) {
  if (Cpad % 8 || C > Cpad) return tg_set_error("tg_unpack_nhwc_tiled: Cpad must be a multiple of 8 and >= C");
  dim3 grid((HW + 63) / 64, (C + 63) / 64, N);
  unpack_nhwc_tiled_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>((const __nv_bfloat16*)in, Cpad, out, out_stride_n, C,
                                                                HW, accumulate);
  TG_RET();
}

int tg_gp_normsq(const void* g, int N, int HW, int C, int c_off, int cj, float* nsq, void* stream) {
  dim3 grid(grid_for(size_t(HW), 256, 64), N);
  gp_normsq_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>((const __nv_bfloat16*)g, HW, C, c_off, cj, nsq);
  TG_RET();
}

int tg_gp_finish(const float* nsq, int N, float lambda, float constant, float* loss, float* coef,
                 void* stream) {
  gp_finish_kernel<<<1, 256, 0, TG_STREAM(stream)>>>(nsq, N, lambda, constant, loss, coef);
  TG_RET();
}

int tg_gp_seed(const void* g, const float* coef, int N, int HW, int C, int c_off, int cj, void* seed,
               void* stream) {
  if (cj > 8 - (c_off & 7)) return tg_set_error("tg_gp_seed: channels must stay inside one 8-group");
  gp_seed_kernel<<<grid_for(size_t(N) * HW, 256, 148 * 16), 256, 0, TG_STREAM(stream)>>>(
      (const __nv_bfloat16*)g, coef, N, HW, C, c_off, cj, (__nv_bfloat16*)seed);
  TG_RET();
}

int tg_adam_step(const void* table_dev, int ntensors, long long max_numel, float lr, float beta1,
                 float beta2, float eps, int step, float grad_scale, void* stream) {
  const float bc1 = 1.f - powf(beta1, float(step));
  const float bc2 = 1.f - powf(beta2, float(step));
  dim3 grid(grid_for(size_t(max_numel), 256, 256), ntensors);
  adam_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>((const AdamTensor*)table_dev, ntensors, lr, beta1,
                                                    beta2, eps, bc1, bc2, grad_scale, nullptr);
  TG_RET();
}

int tg_write_floats(float* dst, const float* values_host, int n, void* stream) {
  if (n < 1 || n > 16 || !dst || !values_host) return tg_set_error("tg_write_floats: 1 <= n <= 16");
  Floats16 v;
  for (int i = 0; i < 16; ++i) v.f[i] = i < n ? values_host[i] : 0.f;
  write_floats_kernel<<<1, 32, 0, TG_STREAM(stream)>>>(dst, v, n);
  TG_RET();
}

int tg_adam_step_dev(const void* table_dev, int ntensors, long long max_numel, const float* hyper_dev, void* stream) {
  if (!hyper_dev) return tg_set_error("tg_adam_step_dev: null hyper-parameter buffer");
  dim3 grid(grid_for(size_t(max_numel), 256, 256), ntensors);
  adam_kernel<<<grid, 256, 0, TG_STREAM(stream)>>>((const AdamTensor*)table_dev, ntensors, 0.f, 0.f, 0.f, 1.f, 1.f,
                                                    1.f, 1.f, hyper_dev);
  TG_RET();
}

}  // extern "C"
