// C-ABI front end: tensor-map construction, tile/split heuristics, plan objects, launches.
#include <cstdio>
#include <cstring>
#include <new>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/tactile_gan_b200.h"
#include "tg_api_internal.h"
#include "tg_igemm.cuh"
#include "tg_wgrad.cuh"
#include "tg_igemm_halo.cuh"
#include "tg_igemm_rows.cuh"
#include "tg_wgrad_taps.cuh"
#include <cstdlib>

static thread_local char g_err[512] = "";

int tg_set_error(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
static int g_pdl = -1;
extern "C" int tg_pdl_policy(int policy) {
  const int prev = tg_pdl_enabled();
  if (policy >= 0 && policy <= 2) g_pdl = policy;
  return prev;
}
int tg_pdl_enabled() {
  int& on = g_pdl;
  if (on < 0) {
    // -1 % on the headline step, +3..13 % on latency-bound configurations (profiles/r02_pdl_ab.txt): off by default here,
    // switched on per configuration by the Python side (step.TrainStep, engine.GraphEngine) through tg_pdl_policy
    const char* e = getenv("TG_PDL");
    on = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  return on;
}
int tg_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -1;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int* g_err_flag = nullptr;
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// 4-D NHWC map {C, W, H, N}; box {64, bw, bh, bn}; optional element strides on W/H.
int make_act_map(CUtensorMap* m, const tg_view& v, int c_off, int c_len, int bw, int bh, int bn,
                 int estride);

// 3-D weight map with a box covering `box_taps` taps: {64, rows_box, box_taps}
int make_wgt_map_taps(CUtensorMap* m, const void* base, int k_len, int rows, int taps, int pitch_k,
                      int pitch_rows, int rows_box, int box_taps);

int make_act_map(CUtensorMap* m, const tg_view& v, int c_off, int c_len, int bw, int bh, int bn,
                 int estride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return tg_set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {cuuint64_t(c_len), cuuint64_t(v.w), cuuint64_t(v.h), cuuint64_t(v.n)};
  cuuint64_t strides[3] = {cuuint64_t(v.sw) * 2, cuuint64_t(v.sh) * 2, cuuint64_t(v.sn) * 2};
  cuuint32_t box[4] = {64, cuuint32_t(bw * estride), cuuint32_t(bh * estride), cuuint32_t(bn)};
  cuuint32_t es[4] = {1, cuuint32_t(estride), cuuint32_t(estride), 1};
  if (box[1] > 256 || box[2] > 256) return tg_set_error("TMA box too large");
  void* base = static_cast<char*>(v.ptr) + size_t(c_off) * 2;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeTiled(act) failed: %d dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) "
             "box=(%u,%u,%u,%u) es=%d ptr=%p",
             int(r), (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
             (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
             (unsigned long long)strides[2], box[0], box[1],
             box[2], box[3], estride, base);
    return -1;
  }
  return 0;
}

// 3-D weight map {K, rows, taps}; box {64, bn, 1}
int make_wgt_map(CUtensorMap* m, const void* base, int k_len, int rows, int taps, int pitch_k,
                 int pitch_rows, int bn) {
  return make_wgt_map_taps(m, base, k_len, rows, taps, pitch_k, pitch_rows, bn, 1);
}

int make_wgt_map_taps(CUtensorMap* m, const void* base, int k_len, int rows, int taps, int pitch_k,
                      int pitch_rows, int bn, int box_taps) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return tg_set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {cuuint64_t(k_len), cuuint64_t(rows), cuuint64_t(taps)};
  cuuint64_t strides[2] = {cuuint64_t(pitch_k) * 2, cuuint64_t(pitch_k) * pitch_rows * 2};
  cuuint32_t box[3] = {64, cuuint32_t(bn), cuuint32_t(box_taps)};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled(wgt) failed: %d k=%d rows=%d taps=%d pitch=%d",
             int(r), k_len, rows, taps, pitch_k);
    return -1;
  }
  return 0;
}

// choose (tn, th, tw) with tn*th*tw == pixels minimising padded work; stats need tn == 1
void choose_tile(int n, int h, int w, int pixels, bool single_image, int* th, int* tw, int* tn) {
  double best = 1e30;
  *th = 1; *tw = pixels; *tn = 1;
  for (int cw = 1; cw <= pixels; cw *= 2) {
    for (int ch = 1; ch * cw <= pixels; ch *= 2) {
      const int cn = pixels / (cw * ch);
      if (single_image && cn != 1) continue;
      if (cw > 128 || ch > 128 || cn > 256) continue;
      const double tiles = double((w + cw - 1) / cw) * ((h + ch - 1) / ch) * ((n + cn - 1) / cn);
      // tie-break: prefer 8..16 wide tiles (L2 locality of the shifted windows), then fewer images
      const double cost = tiles * (1.0 + 1e-3 * (cw < 8 ? 8 - cw : 0) + 1e-4 * (cw > 16 ? 1 : 0) + 1e-5 * cn);
      if (cost < best) { best = cost; *th = ch; *tw = cw; *tn = cn; }
    }
  }
}

}  // namespace

struct tg_plan {
  int kind;  // 0 conv, 1 wgrad, 2 halo-resident conv, 3 row-resident conv, 4 tap-tiled wgrad
  int bn;
  int grid;
  size_t smem;
  int fin_grid;              // > 0: split-K conv, splitk_finalize_kernel follows the GEMM
  tg::SplitFinParams fin;
  tg::IgemmParams conv;
  tg::WgradParams wg;
  tg::HaloParams halo;
  tg::RowsParams rows;
  tg::WgradTapsParams wgt;
};

namespace {
// 3x3-window stride-1 convolutions with a 64/128-wide output run the halo-resident kernel.
bool halo_eligible(const tg_conv_desc* d, int* dy0, int* dx0, int* eh, int* ew) {
  static const bool disabled = getenv("TG_DISABLE_HALO") != nullptr;
  if (disabled || d->stride != 1 || d->taps > 9 || d->taps < 2) return false;
  if (d->pool_out && (d->stats_partial || d->bias || d->act != 0)) return false;
  if (d->out.c != 64 && d->out.c != 128) return false;
  int ymin = 127, ymax = -127, xmin = 127, xmax = -127;
  for (int t = 0; t < d->taps; ++t) {
    ymin = d->tap_dy[t] < ymin ? d->tap_dy[t] : ymin; ymax = d->tap_dy[t] > ymax ? d->tap_dy[t] : ymax;
    xmin = d->tap_dx[t] < xmin ? d->tap_dx[t] : xmin; xmax = d->tap_dx[t] > xmax ? d->tap_dx[t] : xmax;
  }
  if (ymax - ymin > 2 || xmax - xmin > 2) return false;
  for (int s = 0; s < d->num_src; ++s)
    if (d->src[s].wgt_taps > 9 || d->src[s].wgt_taps != d->src[0].wgt_taps) return false;
  for (int t = 0; t < d->taps; ++t)
    if (d->tap_w[t] < 0 || d->tap_w[t] >= d->src[0].wgt_taps) return false;
  *dy0 = ymin; *dx0 = xmin; *eh = ymax - ymin; *ew = xmax - xmin;
  return true;
}

int create_halo_plan(const tg_conv_desc* d, tg_plan* pl, int dy0, int dx0, int eh, int ew) {
  pl->kind = 2;
  tg::HaloParams& p = pl->halo;
  memset(&p, 0, sizeof(p));
  const int cout = d->out.c;
  p.num_src = d->num_src;
  p.taps = d->taps;
  for (int t = 0; t < d->taps; ++t) {
    p.tap_dy[t] = int8_t(d->tap_dy[t] - dy0);
    p.tap_dx[t] = int8_t(d->tap_dx[t] - dx0);
    p.tap_w[t] = d->tap_w[t];
  }
  p.org_dy = dy0; p.org_dx = dx0;
  p.halo_w = tg::kHaloTW + ew;
  p.a_bytes = (tg::kHaloTW + ew) * (tg::kHaloTH + eh) * 128;
  p.b_bytes = d->src[0].wgt_taps * tg::kHaloBN * 128;
  const int up = d->pool_out ? 2 : 1;
  p.pool_out = d->pool_out;
  p.Ho = d->out.h * up; p.Wo = d->out.w * up; p.N = d->out.n;
  p.tiles_h = (p.Ho + tg::kHaloTH - 1) / tg::kHaloTH;
  p.tiles_w = (p.Wo + tg::kHaloTW - 1) / tg::kHaloTW;
  p.n_tiles = cout / tg::kHaloBN;
  p.act = d->act; p.slope = d->slope;
  p.bias = d->bias; p.bias_len = d->bias_len;
  p.stats_partial = d->stats_partial;
  p.stats_tiles_total = d->stats_tiles_total > 0 ? d->stats_tiles_total : p.tiles_h * p.tiles_w;
  p.stats_tile_off = d->stats_tile_off;
  if (p.stats_partial && p.stats_tiles_total < p.stats_tile_off + p.tiles_h * p.tiles_w)
    return tg_set_error("tg_conv_plan_create: stats buffer has too few tile slots");
  p.cout = cout;
  {
    static const int pf = getenv("TG_HALO_PREFETCH") ? atoi(getenv("TG_HALO_PREFETCH")) : 1;
    p.prefetch = pf;
  }
  p.err_flag = tg_error_flag_device_ptr();
  for (int s = 0; s < d->num_src; ++s) {
    const tg_conv_src& cs = d->src[s];
    if (cs.act.c % 64) return tg_set_error("tg_conv_plan_create: source C must be a multiple of 64");
    if (make_act_map(&p.src[s].act, cs.act, 0, cs.act.c, tg::kHaloTW + ew, tg::kHaloTH + eh, 1, 1)) return -1;
    const char* wbase = static_cast<const char*>(cs.wgt) + (size_t(cs.row_off) * cs.wgt_k + cs.k_off) * 2;
    if (make_wgt_map_taps(&p.src[s].wgt, wbase, cs.act.c, cout, cs.wgt_taps, cs.wgt_k, cs.wgt_rows, tg::kHaloBN,
                          cs.wgt_taps))
      return -1;
    p.src[s].c_chunks = cs.act.c / 64;
  }
  if (make_act_map(&p.out, d->out, 0, cout, tg::kHaloTW / up, tg::kHaloTH / up, 1, 1)) return -1;
  const int m_tiles = p.N * p.tiles_h * p.tiles_w;
  const int total = ((m_tiles + tg::kHaloR - 1) / tg::kHaloR) * p.n_tiles;
  pl->grid = total < sm_count() ? total : sm_count();
  pl->smem = tg::kHaloSmem;
  cudaError_t e = cudaFuncSetAttribute(tg::igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(halo): %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}
// Full 3x3 stride-1 same-size convolutions on maps with W % 128 == 0 and H % 4 == 0 run the row-resident kernel
// (N = 192 vertical-scatter UMMAs, tg_igemm_rows.cuh). TG_ROWS=0 keeps them on the halo kernel.
bool rows_eligible(const tg_conv_desc* d, int dy0, int dx0, int eh, int ew) {
  static const bool enabled = !(getenv("TG_ROWS") && atoi(getenv("TG_ROWS")) == 0);
  if (!enabled || d->taps != 9 || eh != 2 || ew != 2) return false;
  const int up = d->pool_out ? 2 : 1;
  const int ho = d->out.h * up, wo = d->out.w * up;
  if (wo % tg::kTileM || ho % tg::kRowsG) return false;
  bool seen[3][3] = {};
  for (int t = 0; t < 9; ++t) seen[d->tap_dy[t] - dy0][d->tap_dx[t] - dx0] = true;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      if (!seen[a][b]) return false;
  for (int s = 0; s < d->num_src; ++s)
    if (d->src[s].act.h != ho || d->src[s].act.w != wo) return false;
  return true;
}

int create_rows_plan(const tg_conv_desc* d, tg_plan* pl, int dy0, int dx0) {
  pl->kind = 3;
  tg::RowsParams& p = pl->rows;
  memset(&p, 0, sizeof(p));
  const int cout = d->out.c;
  p.num_src = d->num_src;
  for (int t = 0; t < 9; ++t) p.tap_w[d->tap_dy[t] - dy0][d->tap_dx[t] - dx0] = d->tap_w[t];
  p.org_dy = dy0; p.org_dx = dx0;
  const int up = d->pool_out ? 2 : 1;
  p.pool_out = d->pool_out;
  p.Ho = d->out.h * up; p.Wo = d->out.w * up; p.N = d->out.n;
  p.segs = p.Wo / tg::kTileM;
  p.groups_h = p.Ho / tg::kRowsG;
  p.n_tiles = cout / 64;
  p.act = d->act; p.slope = d->slope;
  p.bias = d->bias; p.bias_len = d->bias_len;
  p.stats_partial = d->stats_partial;
  p.stats_tiles_total = d->stats_tiles_total > 0 ? d->stats_tiles_total : p.Ho * p.segs;
  p.stats_tile_off = d->stats_tile_off;
  if (p.stats_partial && p.stats_tiles_total < p.stats_tile_off + p.Ho * p.segs)
    return tg_set_error("tg_conv_plan_create: stats buffer has too few tile slots");
  p.cout = cout;
  {
    static const int pf = getenv("TG_HALO_PREFETCH") ? atoi(getenv("TG_HALO_PREFETCH")) : 1;
    p.prefetch = pf;
  }
  p.err_flag = tg_error_flag_device_ptr();
  for (int s = 0; s < d->num_src; ++s) {
    const tg_conv_src& cs = d->src[s];
    if (cs.act.c % 64) return tg_set_error("tg_conv_plan_create: source C must be a multiple of 64");
    if (make_act_map(&p.src[s].act, cs.act, 0, cs.act.c, tg::kRowsPix, 1, 1, 1)) return -1;
    const char* wbase = static_cast<const char*>(cs.wgt) + (size_t(cs.row_off) * cs.wgt_k + cs.k_off) * 2;
    if (make_wgt_map_taps(&p.src[s].wgt, wbase, cs.act.c, cout, cs.wgt_taps, cs.wgt_k, cs.wgt_rows, 64, 1)) return -1;
    p.src[s].c_chunks = cs.act.c / 64;
  }
  if (make_act_map(&p.out, d->out, 0, cout, tg::kTileM / up, 1, 1, 1)) return -1;
  const int total = p.N * p.groups_h * p.segs * p.n_tiles;
  pl->grid = total < sm_count() ? total : sm_count();
  pl->smem = tg::kRowsSmem;
  cudaError_t e = cudaFuncSetAttribute(tg::igemm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(rows): %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}
}  // namespace

extern "C" {

int tg_version(void) { return 100; }
const char* tg_last_error(void) { return g_err; }
int tg_device_sm_count(void) { return sm_count(); }

int* tg_error_flag_device_ptr(void) {
  if (!g_err_flag) {
    if (cudaMalloc(&g_err_flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(g_err_flag, 0, sizeof(int));
  }
  return g_err_flag;
}

int tg_error_flag_read(void) {
  int v = 0;
  if (!g_err_flag) return 0;
  if (cudaMemcpy(&v, g_err_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

int tg_conv_query_tiles(int n, int ho, int wo, int want_stats, int* out4) {
  int th, tw, tn;
  choose_tile(n, ho, wo, tg::kTileM, want_stats != 0, &th, &tw, &tn);
  out4[0] = th; out4[1] = tw; out4[2] = tn;
  out4[3] = ((ho + th - 1) / th) * ((wo + tw - 1) / tw);
  // the halo-resident kernel tiles 16x8: report enough tile slots for either kernel
  const int halo_tiles = ((ho + tg::kHaloTH - 1) / tg::kHaloTH) * ((wo + tg::kHaloTW - 1) / tg::kHaloTW);
  if (want_stats && halo_tiles > out4[3]) out4[3] = halo_tiles;
  return 0;
}

int tg_conv_plan_create(const tg_conv_desc* d, tg_plan** out) {
  if (!d || !out) return tg_set_error("tg_conv_plan_create: null argument");
  if (d->num_src < 1 || d->num_src > TG_MAX_SRC) return tg_set_error("tg_conv_plan_create: bad num_src");
  if (d->taps < 1 || d->taps > TG_MAX_TAPS) return tg_set_error("tg_conv_plan_create: bad taps");
  if (d->out.c % 64) return tg_set_error("tg_conv_plan_create: Cout must be a multiple of 64");
  tg_plan* pl = new (std::nothrow) tg_plan();
  if (!pl) return tg_set_error("out of host memory");
  {
    int dy0, dx0, eh, ew;
    if (halo_eligible(d, &dy0, &dx0, &eh, &ew)) {
      for (int s = 0; s < d->num_src; ++s)
        if (d->src[s].act.n != d->out.n) { delete pl; return tg_set_error("tg_conv_plan_create: batch mismatch"); }
      if (rows_eligible(d, dy0, dx0, eh, ew)) {
        if (create_rows_plan(d, pl, dy0, dx0)) { delete pl; return -1; }
      } else if (create_halo_plan(d, pl, dy0, dx0, eh, ew)) { delete pl; return -1; }
      *out = pl;
      return 0;
    }
  }
  pl->kind = 0;
  tg::IgemmParams& p = pl->conv;
  memset(&p, 0, sizeof(p));
  const int cout = d->out.c;
  const int bn = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  pl->bn = bn;
  int th, tw, tn;
  const int up = d->pool_out ? 2 : 1;
  if (d->pool_out) {
    if (d->stats_partial || d->bias || d->act != 0 || d->stride != 1) {
      delete pl;
      return tg_set_error("tg_conv_plan_create: pool_out is a plain stride-1 input-gradient mode (no bias/act/stats)");
    }
    th = 16; tw = 8; tn = 1;       // the pooled epilogue pairs lanes of a 16x8 tile
  } else {
    choose_tile(d->out.n, d->out.h, d->out.w, tg::kTileM, d->stats_partial != nullptr, &th, &tw, &tn);
  }
  p.pool_out = d->pool_out;
  p.num_src = d->num_src;
  p.taps = d->taps;
  p.stride = d->stride;
  memcpy(p.tap_dy, d->tap_dy, 16);
  memcpy(p.tap_dx, d->tap_dx, 16);
  memcpy(p.tap_w, d->tap_w, 16);
  p.Ho = d->out.h * up; p.Wo = d->out.w * up; p.N = d->out.n;
  p.th = th; p.tw = tw; p.tn = tn;
  p.tiles_h = (p.Ho + th - 1) / th;
  p.tiles_w = (p.Wo + tw - 1) / tw;
  p.tiles_img = (p.N + tn - 1) / tn;
  p.n_tiles = cout / bn;
  p.act = d->act;
  p.slope = d->slope;
  p.bias = d->bias;
  p.bias_len = d->bias_len;
  p.stats_partial = d->stats_partial;
  p.stats_tiles_total = d->stats_tiles_total > 0 ? d->stats_tiles_total : p.tiles_h * p.tiles_w;
  p.stats_tile_off = d->stats_tile_off;
  p.cout = cout;
  p.err_flag = tg_error_flag_device_ptr();
  for (int s = 0; s < d->num_src; ++s) {
    const tg_conv_src& cs = d->src[s];
    if (cs.act.c % 64) { delete pl; return tg_set_error("tg_conv_plan_create: source C must be a multiple of 64"); }
    if (cs.act.n != d->out.n) { delete pl; return tg_set_error("tg_conv_plan_create: batch mismatch"); }
    if (make_act_map(&p.src[s].act, cs.act, 0, cs.act.c, tw, th, tn, d->stride)) { delete pl; return -1; }
    const char* wbase = static_cast<const char*>(cs.wgt) + (size_t(cs.row_off) * cs.wgt_k + cs.k_off) * 2;
    if (make_wgt_map(&p.src[s].wgt, wbase, cs.act.c, cout, cs.wgt_taps, cs.wgt_k, cs.wgt_rows, bn)) {
      delete pl; return -1;
    }
    p.src[s].c_chunks = cs.act.c / 64;
  }
  if (make_act_map(&p.out, d->out, 0, cout, tw / up, th / up, tn, 1)) { delete pl; return -1; }
  const int total = p.tiles_img * p.tiles_h * p.tiles_w * p.n_tiles;
  // Split-K: on maps of up to 8x8 pixels (UNet's deepest levels) a launch has a handful of tiles, each with a serial K
  // loop of up to 16 taps x 16 chunks, on a handful of CTAs while the rest of the GPU idles. Spread the K iterations
  // of every tile over `splits` CTAs (>= 4 iterations each) when the launch has at most a quarter of the SMs' worth of
  // tiles (measured: UNet batch 4 +6.8 %; at batch 32 the 8x8 level already has 32 tiles and the extra finalize pass
  // costs more than the split returns). For small batches the split COUNT is capped by the layer (min(16, k/4)), not by
  // the batch, so chunked inference reproduces the single-engine result bit for bit.
  p.splits = 1;
  pl->fin_grid = 0;
  {
    static const bool off = getenv("TG_SPLITK") != nullptr && getenv("TG_SPLITK")[0] == '0';
    int k_iters = 0;
    for (int s = 0; s < d->num_src; ++s) k_iters += d->taps * (d->src[s].act.c / 64);
    const long long slice = (long long)p.N * p.Ho * p.Wo * cout;
    if (!off && d->splitk_ws && !d->stats_partial && !d->pool_out && p.Ho * p.Wo <= 64 && k_iters >= 16 &&
        total * 4 <= sm_count()) {
      // all (tile, split) items in ONE wave; fewer than four splits do not pay for the finalize pass
      int splits = k_iters / 4 < 16 ? k_iters / 4 : 16;
      if (splits > sm_count() / total) splits = sm_count() / total;
      if (splits < 4) splits = 1;
      if ((long long)splits * slice * 4 > d->splitk_ws_bytes) splits = 1;   // enormous batch of tiny maps: enough tiles anyway
      if (splits >= 2) {
        p.splits = splits;
        p.ws = d->splitk_ws;
        p.ws_slice = slice;
        tg::SplitFinParams& f = pl->fin;
        f.ws = d->splitk_ws; f.ws_slice = slice; f.splits = splits;
        f.out = static_cast<__nv_bfloat16*>(d->out.ptr);
        f.sn = d->out.sn; f.sh = d->out.sh; f.sw = d->out.sw;
        f.N = p.N; f.Ho = p.Ho; f.Wo = p.Wo; f.cout = cout;
        f.bias = d->bias; f.bias_len = d->bias_len; f.act = d->act; f.slope = d->slope;
        const long long work = (long long)p.N * p.Ho * p.Wo * (cout / 8);
        pl->fin_grid = int(work / 256 + 1 < 148 * 8 ? work / 256 + 1 : 148 * 8);
      }
    }
  }
  const int items = total * p.splits;
  pl->grid = items < sm_count() ? items : sm_count();
  cudaError_t e;
  if (bn == 256) {
    pl->smem = tg::IgemmCfg<256>::kSmemTotal;
    e = cudaFuncSetAttribute(tg::igemm_conv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  } else if (bn == 128) {
    pl->smem = tg::IgemmCfg<128>::kSmemTotal;
    e = cudaFuncSetAttribute(tg::igemm_conv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  } else {
    pl->smem = tg::IgemmCfg<64>::kSmemTotal;
    e = cudaFuncSetAttribute(tg::igemm_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  }
  if (e != cudaSuccess) {
    delete pl;
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(conv): %s", cudaGetErrorString(e));
    return -1;
  }
  *out = pl;
  return 0;
}

// 3x3-window stride-1 same-size convolutions: tap-tiled kernel (tg_wgrad_taps.cuh)
static int create_wgrad_taps_plan(const tg_wgrad_desc* d, tg_plan* pl) {
  pl->kind = 4;
  tg::WgradTapsParams& p = pl->wgt;
  memset(&p, 0, sizeof(p));
  p.num_src = d->num_src;
  memset(p.tap_w, -1, sizeof(p.tap_w));
  int ymin = 127, xmin = 127;
  for (int t = 0; t < d->taps; ++t) {
    ymin = d->tap_dy[t] < ymin ? d->tap_dy[t] : ymin;
    xmin = d->tap_dx[t] < xmin ? d->tap_dx[t] : xmin;
  }
  p.org_y = ymin + 1; p.org_x = xmin + 1;     // window centre: taps become (dy, dx) in [-1, 1]
  for (int t = 0; t < d->taps; ++t) p.tap_w[d->tap_dy[t] - ymin][d->tap_dx[t] - xmin] = d->tap_w[t];
  const int H = d->q.h, W = d->q.w;
  p.N = d->q.n;
  p.tiles_h = (H + tg::kWtTH - 1) / tg::kWtTH;
  p.tiles_w = (W + 2 * p.org_x + tg::kWtTW - 1) / tg::kWtTW;
  int chunks = 0;
  for (int s = 0; s < d->num_src; ++s) {
    if (d->p[s].c % 64) return tg_set_error("tg_wgrad_plan_create: P channels must be a multiple of 64");
    if (d->p[s].h != d->p[0].h || d->p[s].w != d->p[0].w || d->p[s].n != d->q.n)
      return tg_set_error("tg_wgrad_plan_create: source size mismatch");
    if (make_act_map(&p.src[s].act, d->p[s], 0, d->p[s].c, tg::kWtTW, tg::kWtTH + 2, 1, 1)) return -1;
    p.src[s].c_chunks = d->p[s].c / 64;
    chunks += p.src[s].c_chunks;
  }
  if (make_act_map(&p.q, d->q, 0, d->q.c, tg::kWtTW + 2, tg::kWtTH, 1, 1)) return -1;
  p.total_chunks = chunks;
  p.n_tiles = d->q.c / 64;
  p.dw = d->dw;
  p.m_total = d->dw_cols;
  p.n_total = d->dw_rows;
  if (p.m_total != chunks * 64) return tg_set_error("tg_wgrad_plan_create: dw_cols != sum of P channels");
  if (p.n_total < d->q.c) return tg_set_error("tg_wgrad_plan_create: dw_rows < Q channels");
  p.err_flag = tg_error_flag_device_ptr();
  const int k_tiles = p.N * p.tiles_h * p.tiles_w;
  const int items0 = chunks * p.n_tiles;
  static const int waves = getenv("TG_WGRAD_WAVES") ? atoi(getenv("TG_WGRAD_WAVES")) : 2;
  // items are dealt round-robin to one persistent CTA per SM: keep items0 * splits just BELOW a whole number of
  // rounds (rounding up leaves a nearly empty extra round: 300 items on 148 SMs take 3 item-times, 294 take 2)
  int splits = waves * sm_count() / items0;
  const int max_splits = k_tiles / 4 > 1 ? k_tiles / 4 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  while (splits > 1 && ((k_tiles + splits - 1) / splits) * (splits - 1) >= k_tiles) --splits;
  p.splits = splits;
  const int total = items0 * splits;
  pl->grid = total < sm_count() ? total : sm_count();
  pl->smem = tg::kWtSmem;
  cudaError_t e = cudaFuncSetAttribute(tg::wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(wgrad taps): %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

int tg_wgrad_plan_create(const tg_wgrad_desc* d, tg_plan** out) {
  if (!d || !out) return tg_set_error("tg_wgrad_plan_create: null argument");
  if (d->num_src < 1 || d->num_src > TG_MAX_SRC) return tg_set_error("tg_wgrad_plan_create: bad num_src");
  if (d->q.c % 64) return tg_set_error("tg_wgrad_plan_create: Q channels must be a multiple of 64");
  tg_plan* pl = new (std::nothrow) tg_plan();
  if (!pl) return tg_set_error("out of host memory");
  {
    static const bool disabled = getenv("TG_DISABLE_HALO") != nullptr;
    int ymin = 127, ymax = -127, xmin = 127, xmax = -127;
    for (int t = 0; t < d->taps; ++t) {
      ymin = d->tap_dy[t] < ymin ? d->tap_dy[t] : ymin; ymax = d->tap_dy[t] > ymax ? d->tap_dy[t] : ymax;
      xmin = d->tap_dx[t] < xmin ? d->tap_dx[t] : xmin; xmax = d->tap_dx[t] > xmax ? d->tap_dx[t] : xmax;
    }
    static const int max_qc = getenv("TG_WGRAD_TAPS_MAXC") ? atoi(getenv("TG_WGRAD_TAPS_MAXC")) : 128;
    // any <= 3x3 tap window of a stride-1 conv, padded (X grid == dY grid) or valid (X grid larger)
    const bool window = d->taps <= 9 && ymax - ymin <= 2 && xmax - xmin <= 2;
    const bool grids = ymin >= -1 && xmin >= -1 && ymin <= 0 && xmin <= 0;   // centre offset 0 (padded) or 1 (valid)
    // wide outputs: the per-tap kernel re-streams both operands nine times, which only pays while they sit in L2
    double op_bytes = double(d->q.n) * d->q.h * d->q.w * d->q.c * 2.0;
    for (int s2 = 0; s2 < d->num_src; ++s2) op_bytes += double(d->p[s2].n) * d->p[s2].h * d->p[s2].w * d->p[s2].c * 2.0;
    const bool big = op_bytes > 200.0 * 1024 * 1024;
    if (!disabled && d->stride == 1 && d->taps >= 2 && window && grids && (d->q.c <= max_qc || big)) {
      if (create_wgrad_taps_plan(d, pl)) { delete pl; return -1; }
      *out = pl;
      return 0;
    }
  }
  pl->kind = 1;
  tg::WgradParams& p = pl->wg;
  memset(&p, 0, sizeof(p));
  const int qc = d->q.c;
  const int bn = qc % 256 == 0 ? 256 : (qc % 128 == 0 ? 128 : 64);
  pl->bn = bn;
  int th, tw, tn;
  choose_tile(d->q.n, d->q.h, d->q.w, tg::kWgBoxPix, false, &th, &tw, &tn);
  p.num_src = d->num_src;
  p.taps = d->taps;
  p.stride = d->stride;
  memcpy(p.tap_dy, d->tap_dy, 16);
  memcpy(p.tap_dx, d->tap_dx, 16);
  memcpy(p.tap_w, d->tap_w, 16);
  p.th = th; p.tw = tw; p.tn = tn;
  p.tiles_h = (d->q.h + th - 1) / th;
  p.tiles_w = (d->q.w + tw - 1) / tw;
  p.tiles_img = (d->q.n + tn - 1) / tn;
  int chunks = 0;
  for (int s = 0; s < d->num_src; ++s) {
    if (d->p[s].c % 64) { delete pl; return tg_set_error("tg_wgrad_plan_create: P channels must be a multiple of 64"); }
    if (make_act_map(&p.src[s].act, d->p[s], 0, d->p[s].c, tw, th, tn, d->stride)) { delete pl; return -1; }
    p.src[s].c_chunks = d->p[s].c / 64;
    chunks += p.src[s].c_chunks;
  }
  if (make_act_map(&p.q, d->q, 0, qc, tw, th, tn, 1)) { delete pl; return -1; }
  p.total_chunks = chunks;
  p.m_tiles = (chunks + 1) / 2;
  p.n_tiles = qc / bn;
  p.dw = d->dw;
  p.m_total = d->dw_cols;
  p.n_total = d->dw_rows;
  if (p.m_total != chunks * 64) { delete pl; return tg_set_error("tg_wgrad_plan_create: dw_cols != sum of P channels"); }
  if (p.n_total < qc) { delete pl; return tg_set_error("tg_wgrad_plan_create: dw_rows < Q channels"); }
  p.err_flag = tg_error_flag_device_ptr();
  const int k_blocks = p.tiles_img * p.tiles_h * p.tiles_w;
  const int items0 = p.taps * p.m_tiles * p.n_tiles;
  static const int waves = getenv("TG_WGRAD_WAVES") ? atoi(getenv("TG_WGRAD_WAVES")) : 2;
  int splits = waves * sm_count() / items0;   // whole rounds of the persistent grid, see create_wgrad_taps_plan
  {
    // the pixel ranges being streamed at any one time (one per concurrently active split) should stay
    // L2-resident, because every (tap, m, n) item of a split re-reads them
    double bytes = double(d->q.n) * d->q.h * d->q.w * d->q.c * 2.0;
    for (int s = 0; s < d->num_src; ++s) bytes += double(d->p[s].n) * d->p[s].h * d->p[s].w * d->p[s].c * 2.0;
    const double concurrent = items0 >= sm_count() ? 1.0 : double(sm_count()) / items0;
    const int need = int(bytes * concurrent / (48.0 * 1024 * 1024)) + 1;
    if (need > splits) splits = need;
  }
  const int max_splits = k_blocks / 8 > 1 ? k_blocks / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  // no empty splits: splits <= k_blocks and ceil-division ranges must all be non-empty
  while (splits > 1 && ((k_blocks + splits - 1) / splits) * (splits - 1) >= k_blocks) --splits;
  p.splits = splits;
  const int total = items0 * splits;
  pl->grid = total < sm_count() ? total : sm_count();
  cudaError_t e;
  if (bn == 256) {
    pl->smem = tg::WgradCfg<256>::kSmemTotal;
    e = cudaFuncSetAttribute(tg::wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  } else if (bn == 128) {
    pl->smem = tg::WgradCfg<128>::kSmemTotal;
    e = cudaFuncSetAttribute(tg::wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  } else {
    pl->smem = tg::WgradCfg<64>::kSmemTotal;
    e = cudaFuncSetAttribute(tg::wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl->smem));
  }
  if (e != cudaSuccess) {
    delete pl;
    snprintf(g_err, sizeof(g_err), "cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e));
    return -1;
  }
  *out = pl;
  return 0;
}

int tg_plan_run(tg_plan* pl, void* stream) {
  if (!pl) return tg_set_error("tg_plan_run: null plan");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // every GEMM kernel starts with tg::griddep_sync(): launched as programmatic dependents (tg_launch)
  const dim3 g(pl->grid), b(tg::kNumThreads);
  cudaError_t e;
  if (pl->kind == 2) {
    e = tg_launch(tg::igemm_halo_kernel, g, b, pl->smem, s, pl->halo);
  } else if (pl->kind == 3) {
    e = tg_launch(tg::igemm_rows_kernel, g, dim3(tg::kRowsThreads), pl->smem, s, pl->rows);
  } else if (pl->kind == 4) {
    e = tg_launch(tg::wgrad_taps_kernel, g, b, pl->smem, s, pl->wgt);
  } else if (pl->kind == 0) {
    if (pl->bn == 256) e = tg_launch(tg::igemm_conv_kernel<256>, g, b, pl->smem, s, pl->conv);
    else if (pl->bn == 128) e = tg_launch(tg::igemm_conv_kernel<128>, g, b, pl->smem, s, pl->conv);
    else e = tg_launch(tg::igemm_conv_kernel<64>, g, b, pl->smem, s, pl->conv);
  } else {
    if (pl->bn == 256) e = tg_launch(tg::wgrad_kernel<256>, g, b, pl->smem, s, pl->wg);
    else if (pl->bn == 128) e = tg_launch(tg::wgrad_kernel<128>, g, b, pl->smem, s, pl->wg);
    else e = tg_launch(tg::wgrad_kernel<64>, g, b, pl->smem, s, pl->wg);
  }
  if (e == cudaSuccess && pl->kind == 0 && pl->fin_grid > 0)
    e = tg_launch(tg::splitk_finalize_kernel, dim3(pl->fin_grid), dim3(256), 0, s, pl->fin);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "tg_plan_run: %s", cudaGetErrorString(e));
    return -1;
  }
  return tg_check_launch("tg_plan_run");
}

void tg_plan_destroy(tg_plan* pl) { delete pl; }

}  // extern "C"
