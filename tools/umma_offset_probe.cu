// Probe: can a tcgen05 shared-memory descriptor (128B swizzle) start at an arbitrary 128-byte row of a
// TMA-written tile, and use a non-1024 stride between 8-row groups?  That is what a halo-resident 3x3
// window needs (one TMA box per tile, nine shifted views). Tests K-major and MN-major operands with
// base_offset = 0 and base_offset = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_offset_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../tactile_gan_b200/csrc/tg_ptx.cuh"

using namespace tg;

struct Params {
  CUtensorMap a;  // {64, 256 rows}
  CUtensorMap b;  // {64, 64 rows}
  float* out;     // [variants][128][64]
  int row_off[16], sbo[16], base_off_mode[16], mn_major[16];
  int variants;
};

__device__ __forceinline__ uint64_t desc_full(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = umma_smem_desc_sw128(saddr, lbo, sbo);
  d |= uint64_t(base_off & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a0 = base;               // 256 rows x 128 B = 32 KiB
  const uint32_t a1 = base + 32768;       // second copy (MN-major: channels 64..127 -> same data)
  const uint32_t b0 = base + 65536;       // 64 rows x 128 B = 8 KiB
  const uint32_t bar = base + 65536 + 8192;
  const uint32_t mbar = bar + 8;
  const uint32_t slot = bar + 16;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 32768 * 2 + 8192);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(a0), "l"(reinterpret_cast<uint64_t>(&p.a)), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(a1), "l"(reinterpret_cast<uint64_t>(&p.a)), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(b0), "l"(reinterpret_cast<uint64_t>(&p.b)), "r"(bar), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(bar, 0);
  __syncthreads();
  uint32_t phase = 0;
  for (int v = 0; v < p.variants; ++v) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t off = uint32_t(p.row_off[v]) * 128u;
      const uint32_t sbo = uint32_t(p.sbo[v]);
      if (!p.mn_major[v]) {
        // K-major: D[m][n] = sum_k A[row(m)][k] B[n][k], 4 x K16
        const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        for (int k = 0; k < 4; ++k) {
          const uint32_t aa = a0 + off + k * 32;
          const uint32_t bo = p.base_off_mode[v] ? ((aa >> 7) & 7) : 0;
          umma_f16(tmem, desc_full(aa, 16, sbo, bo), desc_full(b0 + k * 32, 16, 1024, 0), idesc, k > 0);
        }
      } else {
        // MN-major: D[m][n] = sum_{k<64} A[k-row][m] B[k-row][n]; A = two 64-wide boxes (LBO = 32 KiB)
        const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
        for (int k = 0; k < 4; ++k) {
          // rows advance by 16 per MMA: two 8-row groups `sbo` apart
          const uint32_t aa = a0 + off + k * 2 * sbo;
          const uint32_t bb = b0 + k * 2048;
          const uint32_t bo = p.base_off_mode[v] ? ((aa >> 7) & 7) : 0;
          umma_f16(tmem, desc_full(aa, 32768, sbo, bo), desc_full(bb, 8192, 1024, 0), idesc, k > 0);
        }
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16), r0);
    tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + 32, r1);
    tmem_ld_wait();
    float* o = p.out + (size_t(v) * 128 + warp * 32 + lane) * 64;
    for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(r0[j]); o[32 + j] = __uint_as_float(r1[j]); }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int AR = 256, BR = 64, K = 64;
  std::vector<__nv_bfloat16> ha(AR * K), hb(BR * K);
  std::vector<float> fa(AR * K), fb(BR * K);
  srand(1);
  for (int i = 0; i < AR * K; ++i) { fa[i] = float(rand() % 7 - 3); ha[i] = __float2bfloat16(fa[i]); }
  for (int i = 0; i < BR * K; ++i) { fb[i] = float(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  Params p;
  auto mk = [&](CUtensorMap* m, void* ptr, int rows) {
    cuuint64_t dims[2] = {64, cuuint64_t(rows)};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, cuuint32_t(rows)};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  if (mk(&p.a, da, AR) || mk(&p.b, db, BR)) { printf("encode failed\n"); return 1; }
  struct V { int off, sbo, bo, mn; };
  std::vector<V> vs = {{0, 1024, 0, 0}, {1, 1024, 0, 0}, {1, 1024, 1, 0}, {3, 1024, 0, 0}, {3, 1024, 1, 0},
                       {0, 1280, 0, 0}, {11, 1280, 0, 0}, {11, 1280, 1, 0},
                       {0, 1024, 0, 1}, {1, 1024, 0, 1}, {1, 1024, 1, 1}, {0, 1280, 0, 1}, {11, 1280, 0, 1},
                       {11, 1280, 1, 1}};
  p.variants = int(vs.size());
  for (int i = 0; i < p.variants; ++i) {
    p.row_off[i] = vs[i].off; p.sbo[i] = vs[i].sbo; p.base_off_mode[i] = vs[i].bo; p.mn_major[i] = vs[i].mn;
  }
  cudaMalloc(&p.out, size_t(p.variants) * 128 * 64 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  probe_kernel<<<1, 128, 80 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(size_t(p.variants) * 128 * 64);
  cudaMemcpy(out.data(), p.out, out.size() * 4, cudaMemcpyDeviceToHost);
  for (int v = 0; v < p.variants; ++v) {
    const int pitch_rows = vs[v].sbo / 128;  // rows between 8-row groups
    int bad = 0;
    for (int m = 0; m < 128 && bad < 100000; ++m)
      for (int n = 0; n < 64; ++n) {
        float ref = 0.f;
        if (!vs[v].mn) {
          const int row = vs[v].off + (m / 8) * pitch_rows + (m % 8);
          for (int k = 0; k < K; ++k) ref += fa[row * K + k] * fb[n * K + k];
        } else {
          // K index kk (0..63) -> row off + (kk/8)*pitch + kk%8 of A; B rows are dense (kk)
          for (int kk = 0; kk < 64; ++kk) {
            const int row = vs[v].off + (kk / 8) * pitch_rows + (kk % 8);
            ref += fa[row * K + (m % 64)] * fb[kk * K + n];
          }
        }
        if (out[(size_t(v) * 128 + m) * 64 + n] != ref) ++bad;
      }
    printf("variant %2d: %s row_off=%2d sbo=%4d base_offset=%s -> %s (%d mismatches)\n", v,
           vs[v].mn ? "MN-major" : "K-major ", vs[v].off, vs[v].sbo, vs[v].bo ? "(addr>>7)&7" : "0",
           bad == 0 ? "MATCH" : "differs", bad);
  }
  return 0;
}
