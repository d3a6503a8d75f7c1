// Internal glue shared by the translation units of libtactile_gan_b200.so (not part of the C-ABI).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_bf16.h>

#include <cuda_runtime.h>

int tg_set_error(const char* msg);          // records msg, returns -1
int tg_check_launch(const char* what);      // cudaGetLastError() -> 0 / -1
int tg_pdl_enabled();                       // env TG_PDL=1 (default off): programmatic dependent launch for tg_launch

// Launch with the programmatic-stream-serialization attribute (when TG_PDL enables it): the kernel may be scheduled before the previous kernel
// of the stream has finished and MUST call tg::griddep_sync() before its first global-memory access. Inside a stream
// capture (the inference forward's CUDA graph) the attribute is left off.
template <typename... KArgs, typename... Args>
static inline cudaError_t tg_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                    Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int on = tg_pdl_enabled();
  if (on == 1) {   // TG_PDL=1: eager launches only; TG_PDL=2: also inside stream capture (programmatic graph edges)
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) on = 0;
  }
  if (on) on = 1;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = on;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

namespace tg {

// One row of the fused-Adam descriptor table (device resident).
//   kind 0: plain vector (bias / affine / 1x1 head), grad has the same layout as param
//   kind 1: Conv2d weight          torch [O][I][kh][kw]
//   kind 2: ConvTranspose2d weight torch [I][O][kh][kw]
struct AdamTensor {
  float* param;
  const float* grad;        // nullptr: re-pack only (no optimiser update)
  float* m;
  float* v;
  __nv_bfloat16* pack_fwd;  // [taps][o_pad][i_pad]
  __nv_bfloat16* pack_bwd;  // [taps][i_pad][o_pad]
  long long numel;
  int kind, kh, kw, dim1, o_pad, i_pad;  // i_pad = total padded reduction width (all concat sources)
  // input channel -> padded position: the concat sources are padded to 64 separately, so channel ic of
  // segment s (ic < seg_end[s]) sits at ic + seg_shift[s]
  int nseg, seg_end[6], seg_shift[6];
  int pad_;
};

}  // namespace tg
