// liborb_b200.so - extraction path: handle life cycle, batch geometry, launch sequence, C ABI.
// Replaces ORBextractor (reference: include/ORBextractor.h:44-105, src/ORBextractor.cc).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cudaTypedefs.h>

#include <climits>
#include <cmath>

#include <mutex>

#include "orb_kernels_extract.cuh"
#include "orb_kernel_octree_passes.cuh"

static const int8_t h_pattern[1024] = {
#include "orb_pattern_31.inc"
};

int orb_set_error(orb_handle* h, int status, const std::string& msg) {
  if (h) h->last_error = msg;
  return status;
}

int orb_use_device(orb_handle* h) {
  ORB_CUDA_CHECK(h, cudaSetDevice(h->device));
  return ORB_OK;
}

int orb_peer_read_begin(orb_handle* reader, orb_handle* owner) {
  if (reader == owner) return ORB_OK;
  ORB_CUDA_CHECK(reader, cudaEventRecord(owner->ev_sync, owner->stream));
  ORB_CUDA_CHECK(reader, cudaStreamWaitEvent(reader->stream, owner->ev_sync, 0));
  return ORB_OK;
}

int orb_peer_read_end(orb_handle* reader, orb_handle* owner) {
  if (reader == owner) return ORB_OK;
  ORB_CUDA_CHECK(reader, cudaEventRecord(owner->ev_peer, reader->stream));
  ORB_CUDA_CHECK(reader, cudaStreamWaitEvent(owner->stream, owner->ev_peer, 0));
  return ORB_OK;
}

int orb_ensure(orb_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.bytes && b.p) return ORB_OK;
  if (b.p) {
    ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
  }
  bytes = (bytes + 255) / 256 * 256;
  if (bytes == 0) bytes = 256;
  ORB_CUDA_CHECK(h, cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return ORB_OK;
}

// Is `p` page-locked host memory the device can write (cudaHostAlloc / cudaHostRegister under unified addressing)? Asked on every call
// (well under a microsecond): an address can change hands between a page-locked and a pageable allocation.
bool orb_host_buffer_is_device_writable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer == p;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the FUNCTION (per device), not to a handle, and handles are used
// from concurrent host threads (src/Frame.cc:194-197): the limit is only ever raised, under one process-wide lock.
int orb_raise_dyn_smem(orb_handle* h, const void* func, size_t bytes) {
  static std::mutex mu;
  static std::vector<std::pair<std::pair<const void*, int>, size_t>> seen;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : seen)
    if (e.first.first == func && e.first.second == h->device) {
      if (bytes <= e.second) return ORB_OK;
      ORB_CUDA_CHECK(h, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      e.second = bytes;
      return ORB_OK;
    }
  ORB_CUDA_CHECK(h, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  seen.push_back({{func, h->device}, bytes});
  return ORB_OK;
}

static inline int round_half_even(float v) { return (int)lrintf(v); }   // cvRound
static inline int round_half_even(double v) { return (int)lrint(v); }
static inline int floor_i(float v) { int i = (int)v; return i - (i > v); }    // cvFloor
static inline int ceil_i(float v) { int i = (int)v; return i + (i < v); }     // cvCeil
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- constructor tables: src/ORBextractor.cc:413-463 (float/double mix kept as in the reference) ----
static void build_scale_tables(const orb_params& prm, std::vector<float>& scale, std::vector<float>& inv_scale, std::vector<float>& sigma2,
                               std::vector<float>& inv_sigma2, std::vector<int>& nfeat) {
  const int nl = prm.nlevels;
  const double scaleFactor = (double)prm.scale_factor;  // stored in a double member (include/ORBextractor.h:92)
  scale.assign(nl, 0.f); inv_scale.assign(nl, 0.f); sigma2.assign(nl, 0.f); inv_sigma2.assign(nl, 0.f);
  scale[0] = 1.0f; sigma2[0] = 1.0f;
  for (int i = 1; i < nl; ++i) {
    scale[i] = (float)(scale[i - 1] * scaleFactor);
    sigma2[i] = scale[i] * scale[i];
  }
  for (int i = 0; i < nl; ++i) { inv_scale[i] = 1.0f / scale[i]; inv_sigma2[i] = 1.0f / sigma2[i]; }
  nfeat.assign(nl, 0);
  float factor = (float)(1.0f / scaleFactor);
  float nDesired = prm.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    nfeat[l] = round_half_even(nDesired);
    sum += nfeat[l];
    nDesired *= factor;
  }
  nfeat[nl - 1] = std::max(prm.nfeatures - sum, 0);
}

static void build_tables(orb_handle* h) {
  build_scale_tables(h->params, h->scale, h->inv_scale, h->sigma2, h->inv_sigma2, h->nfeat);
  int v, v0, vmax = floor_i(ORB_HALF_PATCH * std::sqrt(2.f) / 2 + 1);
  int vmin = ceil_i(ORB_HALF_PATCH * std::sqrt(2.f) / 2);
  const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
  for (v = 0; v < 16; ++v) h->umax[v] = 0;
  for (v = 0; v <= vmax; ++v) h->umax[v] = round_half_even(std::sqrt(hp2 - v * v));
  for (v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
    while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
    h->umax[v] = v0;
    ++v0;
  }
}

// ---- geometry of a batch of w x h images: level sizes (:1091-1092), FAST cells (:747-761), tree roots (:545-547)
static int build_geometry(orb_handle* h, int w, int hgt, int batch_cap, OrbGeom* out) {
  OrbGeom g;
  std::memset(&g, 0, sizeof(g));
  const int nl = h->params.nlevels;
  g.nlevels = nl;
  g.batch_cap = batch_cap;
  size_t off = 0;
  int cells = 0, blur_tiles = 0, max_root = 0, worst = 0;
  unsigned int scratch = 0;
  for (int l = 0; l < nl; ++l) {
    g.w[l] = round_half_even((float)w * h->inv_scale[l]);
    g.h[l] = round_half_even((float)hgt * h->inv_scale[l]);
    if (g.w[l] > ORB_MAX_DIM || g.h[l] > ORB_MAX_DIM) return ORB_ERR_UNSUPPORTED_SIZE;
    g.pitch[l] = (int)align_up((size_t)g.w[l], 16);
    g.level_base[l] = off;
    g.level_fstride[l] = (size_t)g.pitch[l] * g.h[l];
    off += align_up(g.level_fstride[l] * (size_t)batch_cap, 256);
    const int maxBX = g.w[l] - ORB_EDGE + 3, maxBY = g.h[l] - ORB_EDGE + 3;
    const float width = (float)(maxBX - ORB_BORDER), height = (float)(maxBY - ORB_BORDER);
    if (width < 35.f || height < 35.f) return ORB_ERR_UNSUPPORTED_SIZE;  // reference: nCols == 0 -> division by zero
    g.ncols[l] = (int)(width / 35.f);
    g.nrows[l] = (int)(height / 35.f);
    g.wcell[l] = (int)std::ceil(width / g.ncols[l]);
    g.hcell[l] = (int)std::ceil(height / g.nrows[l]);
    if (g.wcell[l] + 6 > ORB_ROI_MAX || g.hcell[l] + 6 > ORB_ROI_MAX) return ORB_ERR_UNSUPPORTED_SIZE;
    g.cell_start[l] = cells;
    cells += g.ncols[l] * g.nrows[l];
    g.level_cap[l] = (int)std::min((long long)ORB_LEVEL_CAP, (long long)g.ncols[l] * g.nrows[l] * ORB_CELL_CAP);
    g.scratch_off[l] = scratch;
    scratch += 2u * (unsigned)g.level_cap[l];  // global fallback buffers of the quad-tree (levels with many candidates)
    const int rw = maxBX - ORB_BORDER, rh = maxBY - ORB_BORDER;
    g.nini[l] = (int)std::round((float)rw / (float)rh);
    if (g.nini[l] < 1) return ORB_ERR_UNSUPPORTED_SIZE;  // reference: hX = w / 0
    g.hx[l] = (float)rw / g.nini[l];
    g.nfeat[l] = h->nfeat[l];
    max_root = std::max(max_root, std::max(g.nfeat[l], 4 * g.nini[l]));
    // a level may overshoot N by 3, and a tiny-N level can emit up to 4 * nIni leaves
    worst += std::max(g.nfeat[l] + 3, 4 * g.nini[l]);
    g.scale[l] = h->scale[l];
    g.inv_scale[l] = h->inv_scale[l];
    g.patch_size[l] = (int)(31 * h->scale[l]);
    g.blur_tile_start[l] = blur_tiles;
    g.blur_tiles_x[l] = (g.w[l] + BLUR_TW - 1) / BLUR_TW;
    blur_tiles += g.blur_tiles_x[l] * ((g.h[l] + BLUR_TH - 1) / BLUR_TH);
  }
  g.cell_start[nl] = cells;
  g.blur_tile_start[nl] = blur_tiles;
  g.scratch_frame = scratch;
  g.lvl_kcap = max_root + 8;
  g.node_cap = max_root + 16;
  g.kcap = std::max(h->params.nfeatures + 3 * nl, worst);
  g.ini_th = h->params.ini_th_fast;
  g.min_th = h->params.min_th_fast;
  *out = g;
  return ORB_OK;
}
static size_t slab_total(const OrbGeom& g) {
  const int l = g.nlevels - 1;
  return g.level_base[l] + align_up(g.level_fstride[l] * (size_t)g.batch_cap, 256);
}

// ---- resize coefficient tables, exactly as cv::resize builds them for INTER_LINEAR 8U
static inline int sat_short(float v) { int i = round_half_even(v); return std::min(std::max(i, -32768), 32767); }
static void axis_table(int S, int D, bool horizontal, std::vector<int>& tab) {
  const double inv_scale = (double)D / S;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < D; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = floor_i(f);
    f -= s;
    if (horizontal) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= S - 1) { s = S - 1; f = 0.f; }
    }
    const int c0 = sat_short((1.f - f) * 2048.f), c1 = sat_short(f * 2048.f);
    tab.push_back(s);
    tab.push_back((c0 & 0xffff) | (c1 << 16));
  }
}

static int upload_resize_tables(orb_handle* h) {
  const OrbGeom& g = h->g;
  std::vector<int> tab;
  for (int l = 1; l < g.nlevels; ++l) {
    h->xtab_off[l] = (int)(tab.size() / 2);
    axis_table(g.w[l - 1], g.w[l], true, tab);
    h->ytab_off[l] = (int)(tab.size() / 2);
    axis_table(g.h[l - 1], g.h[l], false, tab);
    h->area2x[l] = (g.w[l - 1] == 2 * g.w[l] && g.h[l - 1] == 2 * g.h[l]) ? 1 : 0;
    // source window of the tile kernel: widest / tallest window any RS_OW x RS_OH destination tile needs
    const int* xt = tab.data() + 2 * (size_t)h->xtab_off[l];
    const int* yt = tab.data() + 2 * (size_t)h->ytab_off[l];
    const int sw = g.w[l - 1], sh = g.h[l - 1];
    int bw = 0, bh = 0;
    for (int ox0 = 0; ox0 < g.w[l]; ox0 += RS_OW) {
      const int last = std::min(ox0 + RS_OW, g.w[l]) - 1;
      bw = std::max(bw, std::min(xt[2 * last] + 1, sw - 1) - (xt[2 * ox0] & ~15) + 1);
    }
    for (int oy0 = 0; oy0 < g.h[l]; oy0 += RS_OH) {
      const int last = std::min(oy0 + RS_OH, g.h[l]) - 1;
      const int y0 = std::min(std::max(yt[2 * oy0], 0), sh - 1), y1 = std::min(std::max(yt[2 * last] + 1, 0), sh - 1);
      bh = std::max(bh, y1 - y0 + 1);
    }
    h->rs_bw[l] = (bw + 15) & ~15;
    h->rs_bh[l] = bh;
    h->rs_tiles[l] = !h->area2x[l] && h->rs_bw[l] <= 256 && bh <= 256;
    if (h->rs_tiles[l]) {   // word variant of the horizontal pass: every group of 4 destination columns within 8 source bytes
      bool words = true;
      for (int x0 = 0; x0 < g.w[l] && words; x0 += 4) {
        const int first = xt[2 * x0];
        for (int i = 0; i < 4; ++i) {
          const int c = xt[2 * std::min(x0 + i, g.w[l] - 1)];
          if (c < first || std::min(c + 1, sw - 1) - first > 7) words = false;
        }
      }
      if (words) h->rs_tiles[l] = 2;
    }
  }
  if (tab.empty()) tab.resize(2, 0);
  int st = orb_ensure(h, h->d_tab, tab.size() * sizeof(int));
  if (st) return st;
  // blur tile table: blockIdx.x -> level | tile column << 4 | tile row << 16
  std::vector<uint32_t> tiles;
  for (int l = 0; l < g.nlevels; ++l) {
    const int ntx = g.blur_tiles_x[l], nty = (g.h[l] + BLUR_TH - 1) / BLUR_TH;
    for (int ty = 0; ty < nty; ++ty)
      for (int tx = 0; tx < ntx; ++tx) tiles.push_back((uint32_t)l | ((uint32_t)tx << 4) | ((uint32_t)ty << 16));
  }
  if ((st = orb_ensure(h, h->d_blur_tiles, tiles.size() * sizeof(uint32_t)))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpy(h->d_blur_tiles.p, tiles.data(), tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy(h->d_tab.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
  return ORB_OK;
}

// Quad-tree launch sizing per level: node slots follow the level's feature budget; the shared-memory key
// buffers hold ~8 candidates per requested feature (more go through the global fallback buffers, same code).
// Small footprints matter: the kernel is latency-bound, one warp per (frame, level), so throughput is the
// number of resident warps.
static int octree_node_cap(const OrbGeom& g, int l) { return std::max(g.nfeat[l], 4 * g.nini[l]) + 16; }
static int octree_smem_keys(const OrbGeom& g, int l) {
  return std::min(ORB_TREE_SMEM_KEYS, std::min(g.level_cap[l], (8 * g.nfeat[l] + 512 + 31) & ~31));
}
static size_t octree_smem_bytes(int node_cap, int smem_keys) {
  return (size_t)node_cap * (8 + 8 + 4 * 4 + 2 + 1) + 2 * sizeof(uint32_t) * (size_t)smem_keys + 64;
}
static size_t octree_smem_max(const OrbGeom& g) {
  size_t m = 0;
  for (int l = 0; l < g.nlevels; ++l) m = std::max(m, octree_smem_bytes(octree_node_cap(g, l), octree_smem_keys(g, l)));
  return m;
}

// ---- FAST tile kernel: tile geometry and one TMA descriptor per level. The level's frames are contiguous
//      (level-major pyramid, frame stride = pitch * h), so they form ONE 2-D u8 tensor of pitch x (h * batch)
//      with row stride pitch; a tile of any frame is a box of FT_TP x bh bytes at ((X0 - 4) & ~15, frame * h + Y0 - 3).
static PFN_cuTensorMapEncodeTiled get_tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled)p;
  }
  return fn;
}

// TMA descriptor over level l of `base` (d_pyr or d_blur layout) with a box of box_w x box_h bytes
static int encode_level_map(orb_handle* h, const uint8_t* base, int l, int box_w, int box_h, CUtensorMap* out) {
  const OrbGeom& g = h->g;
  PFN_cuTensorMapEncodeTiled encode = get_tensor_map_encoder();
  if (!encode) return orb_set_error(h, ORB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)g.pitch[l], (cuuint64_t)g.h[l] * (cuuint64_t)g.batch_cap};
  const cuuint64_t gstride[1] = {(cuuint64_t)g.pitch[l]};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base) + g.level_base[l], gdim, gstride, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed for level %d (CUresult %d)", l, (int)r);
    return orb_set_error(h, ORB_ERR_CUDA, buf);
  }
  return ORB_OK;
}

static int setup_fast_tiles(orb_handle* h) {
  const OrbGeom& g = h->g;
  size_t smem_max = 0;
  int st;
  for (int l = 0; l < g.nlevels; ++l) {
    FastTileGeom& t = h->ftg[l];
    const int wc = g.wcell[l], hc = g.hcell[l];
    if (wc > FT_MAXW || hc > FT_MAXH) return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "FAST cell larger than the tile kernel's limits");
    t.nbx = std::max(1, std::min(g.ncols[l], FT_MAXW / wc));
    t.nby = std::max(1, std::min(g.nrows[l], FT_TILE_H / hc));
    t.bh = t.nby * hc + 6;
    t.sp = (t.nbx * wc + 3 + 2 + 3) & ~3;   // xt columns: up to 3 before the interior, plus the zero ring
    t.wpr = (wc + 31) / 32;
    t.list_cap = (t.nbx * wc * t.nby * hc + 31) & ~31;
    t.list1_cap = (((t.nbx * wc + 3 + 3) / 4) * t.nby * hc + 31) & ~31;
    t.mul_w = (65536u + wc - 1) / wc;
    t.mul_h = (65536u + hc - 1) / hc;
    smem_max = std::max(smem_max, fast_tile_smem(t, hc));
    if ((st = encode_level_map(h, h->d_pyr.as<uint8_t>(), l, FT_TP, t.bh, &h->tmap_fast[l]))) return st;
    if ((st = encode_level_map(h, h->d_pyr.as<uint8_t>(), l, BLUR_TP, BLUR_TR, &h->blur_maps.m[l]))) return st;
    if ((st = encode_level_map(h, h->d_blur.as<uint8_t>(), l, DESC_BOXW, 37, &h->desc_maps.m[l]))) return st;   // k_orient_describe's patch
    if ((st = encode_level_map(h, h->d_pyr.as<uint8_t>(), l, ORB_IC_BOXW, 31, &h->ic_maps.m[l]))) return st;     // its moment disc
    if (l > 0 && h->rs_tiles[l] &&
        (st = encode_level_map(h, h->d_pyr.as<uint8_t>(), l - 1, h->rs_bw[l], h->rs_bh[l], &h->tmap_resize[l])))
      return st;
  }
  if (smem_max > 227 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "FAST tile does not fit shared memory");
  if ((st = orb_raise_dyn_smem(h, (const void*)k_fast_tiles, smem_max))) return st;
  if ((st = orb_raise_dyn_smem(h, (const void*)k_resize_tiles<false>, 256 * 256 + 64))) return st;
  if ((st = orb_raise_dyn_smem(h, (const void*)k_resize_tiles<true>, 256 * 256 + 64))) return st;
  return ORB_OK;
}

// ---- FAST cell kernel (orb_kernel_fast_cells.cuh): items, per-level constants, TMA descriptors, shared-memory slice
static int setup_fast_cells(orb_handle* h) {
  const OrbGeom& g = h->g;
  FastCellGeom& f = h->fcg;
  std::memset(&f, 0, sizeof(f));
  std::vector<uint32_t> items;
  int tile_max = 0, score_max = 0, mask_max = 0, px_max = 0, st;
  for (int l = 0; l < g.nlevels; ++l) {
    const int wc = g.wcell[l], hc = g.hcell[l];
    if (hc > 120 || wc > 107 || g.nrows[l] > 255 || g.ncols[l] > 255)
      return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "FAST cell larger than the cell kernel's limits");
    const int SP = (wc + 2 + 3) & ~3, SCELL = (int)align_up((size_t)(hc + 2) * SP, 16), WPR = (wc + 31) / 32;
    auto box_w = [&](int G) { return std::max(64, (int)align_up((size_t)G * wc + 21, 16)); };
    auto foot = [&](int G) { return box_w(G) * (hc + 6) + G * SCELL + (int)align_up((size_t)G * hc * WPR * 4, 16); };
    int G = 1;
    if (g.ncols[l] >= 2 && box_w(2) <= 128 && foot(2) <= FC_SLICE_BUDGET) G = 2;
    const int bw = box_w(G);
    if (bw > 128) return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "FAST cell wider than a 128-byte tile row");
    f.G[l] = G; f.PW[l] = bw / 4; f.SP[l] = SP; f.SCELL[l] = SCELL; f.WPR[l] = WPR;
    f.MPW[l] = ((1u << 20) + f.PW[l] - 1) / f.PW[l];
    tile_max = std::max(tile_max, bw * (hc + 6));
    score_max = std::max(score_max, G * SCELL);   // also holds the word list of passes A / B
    mask_max = std::max(mask_max, (int)align_up((size_t)G * hc * WPR * 4, 16));
    px_max = std::max(px_max, G * wc * hc);
    if ((st = encode_level_map(h, h->d_pyr.as<uint8_t>(), l, bw, hc + 6, &h->fast_maps.m[l]))) return st;
    h->fc_lvl_first[l] = (int)items.size();
    h->fc_lvl_items[l] = g.nrows[l] * ((g.ncols[l] + G - 1) / G);
    for (int i = 0; i < g.nrows[l]; ++i)
      for (int j = 0; j < g.ncols[l]; j += G) items.push_back(fc_item_code(l, i, j, std::min(G, g.ncols[l] - j)));
  }
  f.items_per_frame = (int)items.size();
  f.score_off = 128 + (int)align_up((size_t)tile_max, 16);
  f.mask_off = f.score_off + score_max;
  f.list_off = f.mask_off + mask_max;
  f.bar_off = f.list_off + FC_L2S * 2 + 2 * 32 * 4;   // corner list, then the two cells' lists of local maxima
  f.warp_stride = (int)align_up((size_t)f.bar_off + 32, 128);
  f.spill_cap = std::max(px_max - FC_L2S, 0) + 8;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  const int per_cta = std::min(optin, (228 * 1024) / FC_MINB - 1024);   // FC_MINB CTAs per SM, 1 KB reserved per CTA
  h->fc_wpc_max = std::min(FC_MAX_WARPS, per_cta / f.warp_stride);
  if (h->fc_wpc_max < 1) return orb_set_error(h, ORB_ERR_CAPACITY, "FAST cell slice does not fit shared memory");
  if ((st = orb_ensure(h, h->d_fast_items, items.size() * sizeof(uint32_t)))) return st;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy(h->d_fast_items.p, items.data(), items.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  const size_t grid_warps = (size_t)FC_MINB * h->sm_count * h->fc_wpc_max;
  if ((st = orb_ensure(h, h->d_fast_spill, grid_warps * f.spill_cap * sizeof(uint16_t)))) return st;
  if ((st = orb_ensure(h, h->d_fast_work, 256))) return st;
  const size_t smem = (size_t)h->fc_wpc_max * f.warp_stride;
  if ((st = orb_raise_dyn_smem(h, (const void*)k_fast_cells<false>, smem))) return st;
  if ((st = orb_raise_dyn_smem(h, (const void*)k_fast_cells<true>, smem))) return st;
  return ORB_OK;
}

static int ensure_buffers(orb_handle* h, const OrbGeom& g, int batch) {
  int st;
  const size_t B = (size_t)batch;
  const size_t cells = (size_t)g.cell_start[g.nlevels];
  // + 256: the stereo matcher stages patches with aligned word loads that may run a few bytes past the last row
  if ((st = orb_ensure(h, h->d_pyr, slab_total(g) + 256))) return st;
  if ((st = orb_ensure(h, h->d_blur, slab_total(g)))) return st;
  if ((st = orb_ensure(h, h->d_cell_count, B * cells * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_cell_keys, B * cells * ORB_CELL_CAP * sizeof(uint32_t)))) return st;
  if ((st = orb_ensure(h, h->d_lvl_count, B * g.nlevels * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_tree_scratch, B * (size_t)g.scratch_frame * sizeof(uint32_t)))) return st;
  if ((st = orb_ensure(h, h->d_sel_count, B * g.nlevels * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_sel_keys, B * g.nlevels * g.lvl_kcap * sizeof(uint32_t)))) return st;
  if ((st = orb_ensure(h, h->d_ord_src, B * g.kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_ord_dst, B * g.kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_kps, B * g.kcap * sizeof(orb_keypoint)))) return st;
  if ((st = orb_ensure(h, h->d_desc, B * g.kcap * 32))) return st;
  if ((st = orb_ensure(h, h->d_n, B * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_mono, B * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_status, B * sizeof(int)))) return st;
  // zero once here: every extraction's k_assemble hands the status words to the host and clears them for the next one
  ORB_CUDA_CHECK(h, cudaMemsetAsync(h->d_status.p, 0, B * sizeof(int), h->stream));
  if (h->h_cap < batch) {
    if (h->h_n) { cudaFreeHost(h->h_n); cudaFreeHost(h->h_mono); cudaFreeHost(h->h_status); }
    ORB_CUDA_CHECK(h, cudaMallocHost((void**)&h->h_n, B * sizeof(int)));
    ORB_CUDA_CHECK(h, cudaMallocHost((void**)&h->h_mono, B * sizeof(int)));
    ORB_CUDA_CHECK(h, cudaMallocHost((void**)&h->h_status, B * sizeof(int)));
    h->h_cap = batch;
  }
  return ORB_OK;
}

// (re)configure the handle for w x h x batch; keeps geometry when nothing changed
static int configure(orb_handle* h, int w, int hgt, int batch) {
  int st;
  const int batch_cap = std::max(batch, h->max_batch);
  if (h->cur_w != w || h->cur_h != hgt || h->g.batch_cap != batch_cap) {
    OrbGeom g;
    if ((st = build_geometry(h, w, hgt, batch_cap, &g)))
      return orb_set_error(h, st, "image size unsupported: every pyramid level needs at least one 35-px FAST cell "
                                  "inside its 16-px border, width/height ratio >= 0.5, and sides <= 4095");
    h->g = g;
    h->geom_gen++;   // buffers / tensor maps / geometry change below: a captured pipeline graph is stale from here on
    h->have_batch = false;
    h->have_stereo = false;
    h->have_fe = false;
    h->have_fe_tri = false;
    // the new size is committed only when every step below succeeded: after a failure (out of memory, tensor-map encoding,
    // capacity) the next call with the same size must configure again instead of launching on stale buffers
    h->cur_w = h->cur_h = 0;
    if ((st = ensure_buffers(h, g, batch_cap))) return st;
    if ((st = upload_resize_tables(h))) return st;
    if ((st = setup_fast_tiles(h))) return st;
    if ((st = setup_fast_cells(h))) return st;
    const size_t smem = octree_smem_max(g);
    if (smem > 227 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "nfeatures too large for the quad-tree kernel");
    if ((st = orb_raise_dyn_smem(h, (const void*)k_octree, smem))) return st;
    h->cur_w = w; h->cur_h = hgt;
    h->max_batch = batch_cap;   // a larger batch grew the buffers: keep them for the smaller ones that follow
  }
  return ORB_OK;
}

#define ORB_GRAPH_MAX_BATCH 8   // batches up to this size are latency chains: per-level branches, pipeline replayed as one CUDA graph
static void stage_mark(orb_handle* h, int i) {
  if (h->stage_timing) cudaEventRecord(h->ev_stage[i], h->stream);
}

// enqueue the whole extraction for `batch` frames whose level 0 is already in d_pyr
// Enqueue the whole extraction for `batch` frames whose level 0 is already in d_pyr.
// Stream plan (one main stream per handle plus auxiliary streams, fork/join with events):
//   main : resize levels 1..L-1 -> FAST per level -> compaction -> quad-trees -> assemble -> [join blur] -> orient+describe
//   aux 0: blur of all levels (needs only the pyramid; overlaps FAST and the latency-bound quad-trees)
// With stage timing enabled the blur stays on the main stream so that every stage is bracketed by events there.
static int launch_pipeline(orb_handle* h, int batch, int lap0, int lap1) {
  const OrbGeom& g = h->g;
  cudaStream_t s = h->stream;
  uint8_t* pyr = h->d_pyr.as<uint8_t>();
  uint8_t* blur = h->d_blur.as<uint8_t>();
  const int cells = g.cell_start[g.nlevels];
  const bool fork_blur = !h->stage_timing;
  stage_mark(h, 0);   // (d_status is zero here: cleared at allocation and by the previous extraction's k_assemble)
  // Small batches are latency chains (one pair is how Tracking calls the path): there FAST and the quad-tree of a level start on the
  // level's own stream as soon as the level exists - level 0 with the upload - instead of after the whole pyramid. The longest link,
  // the level-0 quad-tree, then runs underneath the pyramid and the other levels (critical path per extraction at batch 1:
  // pyramid + FAST + quad-tree 0.14 ms -> FAST(0) + quad-tree(0) 0.09 ms). Large batches keep one launch per stage.
  bool per_level = h->level_pipe && !h->stage_timing && batch <= ORB_GRAPH_MAX_BATCH && h->fast_mode == 1 && h->octree_passes;
  if (per_level) {   // every branch's FAST launch needs its own slice of the per-warp spill buffer
    size_t warps = 0;
    for (int l = 0; l < g.nlevels; ++l) warps += (size_t)batch * h->fc_lvl_items[l] + FC_MAX_WARPS;
    if (warps > (size_t)FC_MINB * h->sm_count * h->fc_wpc_max) per_level = false;
  }
  size_t spill_warp0 = 0;
  auto level_branch = [&](int l) -> int {
    cudaStream_t sl = h->aux[1 + l];
    ORB_CUDA_CHECK(h, cudaEventRecord(h->ev_fork[1 + l], s));
    ORB_CUDA_CHECK(h, cudaStreamWaitEvent(sl, h->ev_fork[1 + l], 0));
    FastCellGeom f = h->fcg;
    f.items_per_frame = h->fc_lvl_items[l];
    const int total = batch * f.items_per_frame, ctas_max = FC_MINB * h->sm_count;
    const int wpc = std::min(h->fc_wpc_max, std::max(1, (total + ctas_max - 1) / ctas_max));
    const int grid = std::min(ctas_max, (total + wpc - 1) / wpc);
    const size_t smem = (size_t)wpc * f.warp_stride;
    const uint32_t* items = h->d_fast_items.as<uint32_t>() + h->fc_lvl_first[l];
    uint16_t* spill = h->d_fast_spill.as<uint16_t>() + spill_warp0 * f.spill_cap;
    spill_warp0 += (size_t)grid * wpc;
    if (g.ini_th < 128)
      k_fast_cells<false><<<grid, wpc * 32, smem, sl>>>(h->fast_maps, g, f, items, total, h->d_cell_count.as<int>(), h->d_cell_keys.as<uint32_t>(),
                                                     cells, spill, h->d_status.as<int>(), nullptr);
    else
      k_fast_cells<true><<<grid, wpc * 32, smem, sl>>>(h->fast_maps, g, f, items, total, h->d_cell_count.as<int>(), h->d_cell_keys.as<uint32_t>(),
                                                    cells, spill, h->d_status.as<int>(), nullptr);
    h->launches++;
    const int nc = octree_node_cap(g, l), sk = octree_smem_keys(g, l);
    k_octree_passes<<<dim3(batch, 1), OP_THREADS, octree_passes_smem_bytes(nc, sk), sl>>>(
        g, h->d_cell_count.as<int>(), h->d_cell_keys.as<uint32_t>(), cells, h->d_tree_scratch.as<uint32_t>(), h->d_lvl_count.as<int>(),
        h->d_sel_count.as<int>(), h->d_sel_keys.as<uint32_t>(), h->d_status.as<int>(), l, nc, sk, nullptr, 0);
    h->launches++;
    ORB_CUDA_CHECK(h, cudaEventRecord(h->ev_join[1 + l], sl));
    return ORB_OK;
  };
  if (per_level) {
    int nc = 0, sk = 0;
    for (int l = 0; l < g.nlevels; ++l) { nc = std::max(nc, octree_node_cap(g, l)); sk = std::max(sk, octree_smem_keys(g, l)); }
    int st2;
    if ((st2 = orb_raise_dyn_smem(h, (const void*)k_octree_passes, octree_passes_smem_bytes(nc, sk)))) return st2;
    if ((st2 = level_branch(0))) return st2;
  }
  for (int l = 1; l < g.nlevels; ++l) {
    const int2* xtab = h->d_tab.as<int2>() + h->xtab_off[l];
    const int2* ytab = h->d_tab.as<int2>() + h->ytab_off[l];
    if (h->rs_tiles[l]) {
      const dim3 grd((g.w[l] + RS_OW - 1) / RS_OW, (g.h[l] + RS_OH - 1) / RS_OH, batch);
      static const bool byte_loads = [] { const char* e = getenv("ORB_B200_RESIZE"); return e && !strcmp(e, "bytes"); }();   // measurement switch
      if (h->rs_tiles[l] == 2 && !byte_loads)
        k_resize_tiles<true><<<grd, RS_WARPS * 32, (size_t)h->rs_bw[l] * h->rs_bh[l] + 16, s>>>(h->tmap_resize[l], g, pyr, l, xtab, ytab,
                                                                                              h->rs_bw[l], h->rs_bh[l]);
      else
        k_resize_tiles<false><<<grd, RS_WARPS * 32, (size_t)h->rs_bw[l] * h->rs_bh[l] + 16, s>>>(h->tmap_resize[l], g, pyr, l, xtab, ytab,
                                                                                               h->rs_bw[l], h->rs_bh[l]);
    } else {  // exact 2x levels (OpenCV's box filter) and ratios whose source window exceeds a TMA box
      dim3 blk(32, 8), grd((g.w[l] + 127) / 128, (g.h[l] + 7) / 8, batch);
      k_resize_level<<<grd, blk, 0, s>>>(g, pyr, l, xtab, ytab, h->area2x[l]);
    }
    h->launches++;
    if (per_level) { const int st2 = level_branch(l); if (st2) return st2; }
  }
  stage_mark(h, 1);
  cudaStream_t sb = fork_blur ? h->aux[0] : s;
  if (fork_blur) {
    ORB_CUDA_CHECK(h, cudaEventRecord(h->ev_fork[0], s));
    ORB_CUDA_CHECK(h, cudaStreamWaitEvent(sb, h->ev_fork[0], 0));
  }
  k_blur7<<<dim3(g.blur_tile_start[g.nlevels], batch), 256, BLUR_SMEM, sb>>>(h->blur_maps, g, h->d_blur_tiles.as<uint32_t>(), blur);
  h->launches++;
  if (fork_blur) ORB_CUDA_CHECK(h, cudaEventRecord(h->ev_join[0], sb));
  stage_mark(h, 2);
  if (per_level) {
    for (int l = 0; l < g.nlevels; ++l) ORB_CUDA_CHECK(h, cudaStreamWaitEvent(s, h->ev_join[1 + l], 0));
  } else if (h->fast_mode == 1) {
    // one launch for all levels and frames: a warp per item (one or two cells), persistent over the item list
    const FastCellGeom& f = h->fcg;
    const int total = batch * f.items_per_frame, ctas_max = FC_MINB * h->sm_count;
    const int wpc = std::min(h->fc_wpc_max, std::max(1, (total + ctas_max - 1) / ctas_max));
    const int grid = std::min(ctas_max, (total + wpc - 1) / wpc);
    const size_t smem = (size_t)wpc * f.warp_stride;
    // more than a few items per warp: dynamic hand-out behind an atomic counter (d_fast_work, zeroed here)
    int* work = nullptr;
    if (h->fast_dynamic && total > 4 * grid * wpc) {
      work = h->d_fast_work.as<int>();
      ORB_CUDA_CHECK(h, cudaMemsetAsync(work, 0, sizeof(int), s));
    }
    if (g.ini_th < 128)
      k_fast_cells<false><<<grid, wpc * 32, smem, s>>>(h->fast_maps, g, f, h->d_fast_items.as<uint32_t>(), total, h->d_cell_count.as<int>(),
                                                  h->d_cell_keys.as<uint32_t>(), cells, h->d_fast_spill.as<uint16_t>(), h->d_status.as<int>(), work);
    else
      k_fast_cells<true><<<grid, wpc * 32, smem, s>>>(h->fast_maps, g, f, h->d_fast_items.as<uint32_t>(), total, h->d_cell_count.as<int>(),
                                                  h->d_cell_keys.as<uint32_t>(), cells, h->d_fast_spill.as<uint16_t>(), h->d_status.as<int>(), work);
    h->launches++;
  } else {
    for (int l = 0; l < g.nlevels; ++l) {
      const FastTileGeom& t = h->ftg[l];
      const dim3 grd((g.ncols[l] + t.nbx - 1) / t.nbx, (g.nrows[l] + t.nby - 1) / t.nby, batch);
      k_fast_tiles<<<grd, FT_THREADS, fast_tile_smem(t, g.hcell[l]), s>>>(h->tmap_fast[l], g, l, t, h->d_cell_count.as<int>(),
                                                                         h->d_cell_keys.as<uint32_t>(), cells, h->d_status.as<int>());
      h->launches++;
    }
  }
  stage_mark(h, 3);
  if (!per_level) {
    // one launch for all (frame, level) quad-trees: one warp each, shared memory sized for the largest level
    int nc = 0, sk = 0;
    for (int l = 0; l < g.nlevels; ++l) { nc = std::max(nc, octree_node_cap(g, l)); sk = std::max(sk, octree_smem_keys(g, l)); }
    if (h->octree_passes) {   // block-parallel pass form (default), see orb_kernel_octree_passes.cuh; ORB_B200_OCTREE=warp selects k_octree
      const size_t sm = octree_passes_smem_bytes(nc, sk);
      { int st2; if ((st2 = orb_raise_dyn_smem(h, (const void*)k_octree_passes, sm))) return st2; }
      k_octree_passes<<<dim3(batch, g.nlevels), OP_THREADS, sm, s>>>(
          g, h->d_cell_count.as<int>(), h->d_cell_keys.as<uint32_t>(), cells, h->d_tree_scratch.as<uint32_t>(),
          h->d_lvl_count.as<int>(), h->d_sel_count.as<int>(), h->d_sel_keys.as<uint32_t>(), h->d_status.as<int>(), -1, nc, sk,
          nullptr, 0);
    } else
    k_octree<<<dim3(batch, g.nlevels), 32, octree_smem_bytes(nc, sk), s>>>(
        g, h->d_cell_count.as<int>(), h->d_cell_keys.as<uint32_t>(), cells, h->d_tree_scratch.as<uint32_t>(),
        h->d_lvl_count.as<int>(), h->d_sel_count.as<int>(), h->d_sel_keys.as<uint32_t>(), h->d_status.as<int>(), -1, nc, sk,
        nullptr, 0);
    h->launches++;
  }
  stage_mark(h, 4);
  k_assemble<<<batch, 256, 0, s>>>(g, h->d_sel_count.as<int>(), h->d_sel_keys.as<uint32_t>(), lap0, lap1,
                                   h->d_ord_src.as<int>(), h->d_ord_dst.as<int>(), h->d_n.as<int>(), h->d_mono.as<int>(),
                                   h->d_status.as<int>(), h->h_n, h->h_mono, h->h_status);
  h->launches++;
  stage_mark(h, 5);
  if (fork_blur) ORB_CUDA_CHECK(h, cudaStreamWaitEvent(s, h->ev_join[0], 0));
  k_orient_describe<<<dim3((g.kcap + DESC_WARPS - 1) / DESC_WARPS, batch), DESC_WARPS * 32, 0, s>>>(
      h->desc_maps, h->ic_maps, g, h->d_n.as<int>(), h->d_ord_src.as<uint32_t>(), h->d_ord_dst.as<int>(), h->d_pattern_f.as<float4>(), h->d_ic_tab.as<uint2>(),
      h->d_kps.as<orb_keypoint>(), h->d_desc.as<uint8_t>(), h->zc_kps, h->zc_desc, h->zc_cap);
  h->launches++;
  stage_mark(h, 6);
  ORB_CUDA_CHECK(h, cudaGetLastError());
  return ORB_OK;
}

// Small batches are launch-bound (21 kernels of a few microseconds each per extraction; a single EuRoC pair spends more host time
// enqueueing than the GPU spends computing), so the pipeline of a batch of at most ORB_GRAPH_MAX_BATCH frames is captured once into a
// CUDA graph - both streams, the fork / join events become graph edges - and replayed with one launch per extraction. The graph
// holds the kernel arguments of launch_pipeline, which only depend on (geometry, batch, lapping area): it is rebuilt when those change.
static void drop_pipeline_graph(orb_handle* h) {
  if (h->pipe_exec) cudaGraphExecDestroy(h->pipe_exec);
  h->pipe_exec = nullptr;
}

static int run_pipeline(orb_handle* h, int batch, int lap0, int lap1) {
  if (h->stage_timing || batch > ORB_GRAPH_MAX_BATCH || h->graph_disabled) return launch_pipeline(h, batch, lap0, lap1);
  if (h->pipe_exec && (h->pipe_batch != batch || h->pipe_lap0 != lap0 || h->pipe_lap1 != lap1 || h->pipe_gen != h->geom_gen ||
                       h->pipe_zc_kps != h->zc_kps || h->pipe_zc_desc != h->zc_desc || h->pipe_zc_cap != h->zc_cap))
    drop_pipeline_graph(h);   // the graph holds the kernel arguments, the caller's result buffers included
  if (!h->pipe_exec) {
    // the first extraction of a configuration runs as plain launches (it also loads the kernels' modules, which must not happen
    // inside a capture); the second one is captured
    if (h->seen_batch != batch || h->seen_lap0 != lap0 || h->seen_lap1 != lap1 || h->seen_gen != h->geom_gen) {
      h->seen_batch = batch; h->seen_lap0 = lap0; h->seen_lap1 = lap1; h->seen_gen = h->geom_gen;
      return launch_pipeline(h, batch, lap0, lap1);
    }
    const int64_t l0 = h->launches;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      h->graph_disabled = true;
      return launch_pipeline(h, batch, lap0, lap1);
    }
    const int st = launch_pipeline(h, batch, lap0, lap1);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    const int64_t n_launch = h->launches - l0;
    h->launches = l0;
    if (st == ORB_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&h->pipe_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (st != ORB_OK || e != cudaSuccess || !h->pipe_exec) {   // capture refused (e.g. another capture in this thread): plain launches from now on
      cudaGetLastError();
      h->pipe_exec = nullptr;
      h->graph_disabled = true;
      return launch_pipeline(h, batch, lap0, lap1);
    }
    h->pipe_batch = batch; h->pipe_lap0 = lap0; h->pipe_lap1 = lap1; h->pipe_gen = h->geom_gen; h->pipe_launches = n_launch;
    h->pipe_zc_kps = h->zc_kps; h->pipe_zc_desc = h->zc_desc; h->pipe_zc_cap = h->zc_cap;
  }
  ORB_CUDA_CHECK(h, cudaGraphLaunch(h->pipe_exec, h->stream));
  h->launches += h->pipe_launches;
  return ORB_OK;
}

static int finish_batch(orb_handle* h) {
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  if (h->stage_timing) {
    for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&h->stage_ms[i], h->ev_stage[i], h->ev_stage[i + 1]);
  }
  int rc = ORB_OK;
  if (h->pending) {
    h->pending = false;
    for (int i = 0; i < h->pending_batch; ++i) {
      int n = h->h_n[i], mono = h->h_mono[i];
      if (h->h_status[i] != 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "frame %d exceeded an internal capacity (status bits 0x%x: 1 cell, 2 level, 4 tree nodes, 8 output)",
                 i, h->h_status[i]);
        rc = orb_set_error(h, ORB_ERR_CAPACITY, buf);
        n = -1; mono = -1;
      }
      if (h->pending_n_out) h->pending_n_out[i] = n;
      if (h->pending_mono_out) h->pending_mono_out[i] = mono;
    }
  }
  return rc;
}

extern "C" {

const char* orb_status_string(int s) {
  switch (s) {
    case ORB_OK: return "ok";
    case ORB_ERR_EMPTY_IMAGE: return "empty image";
    case ORB_ERR_INVALID_ARG: return "invalid argument";
    case ORB_ERR_CUDA: return "CUDA error / no device";
    case ORB_ERR_UNSUPPORTED_SIZE: return "unsupported image size";
    case ORB_ERR_CAPACITY: return "capacity exceeded";
    case ORB_ERR_STATE: return "invalid call order";
  }
  return "unknown";
}

const char* orb_last_error(const orb_handle* h) { return h ? h->last_error.c_str() : "null handle"; }

int orb_create(const orb_params* p, int max_width, int max_height, int max_batch, int device, orb_handle** out) {
  if (!p || !out) return ORB_ERR_INVALID_ARG;
  *out = nullptr;
  if (p->nlevels < 1 || p->nlevels > ORB_MAX_LEVELS || p->nfeatures < 1 || p->min_th_fast < 1 ||
      p->min_th_fast > 127 || p->ini_th_fast < p->min_th_fast || p->ini_th_fast > 254 || !(p->scale_factor > 1.0f) || max_width < 1 ||
      max_height < 1 || max_batch < 1)
    return ORB_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ORB_ERR_CUDA;
  orb_handle* h = new orb_handle();
  h->device = device;
  h->params = *p;
  h->max_w = max_width; h->max_h = max_height; h->max_batch = max_batch;
  { const char* e = getenv("ORB_B200_OCTREE"); h->octree_passes = !(e && !strcmp(e, "warp")); }   // measurement switch: the one-warp list kernel of round 1
  { const char* e = getenv("ORB_B200_FAST"); h->fast_mode = (e && !strcmp(e, "tiles")) ? 0 : 1; }   // measurement switch: round-1 tile kernel
  { const char* e = getenv("ORB_B200_FAST_STATIC"); h->fast_dynamic = !(e && e[0] == '1'); }   // measurement switch: fixed-stride item assignment at every batch size
  { const char* e = getenv("ORB_B200_LEVEL_PIPE"); h->level_pipe = !(e && e[0] == '0'); }   // measurement switch: one FAST / quad-tree launch for all levels at small batches too
  { const char* e = getenv("ORB_B200_NO_GRAPH"); h->graph_disabled = (e && e[0] == '1'); }   // measurement switch: plain launches for small batches too
  auto fail = [&](int st) { orb_destroy(h); return st; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(ORB_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(ORB_ERR_CUDA);
  cudaEventCreate(&h->ev_start); cudaEventCreate(&h->ev_stop);
  cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_peer, cudaEventDisableTiming);
  for (int i = 0; i < ORB_MAX_LEVELS + 1; ++i) {
    if (cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking) != cudaSuccess) return fail(ORB_ERR_CUDA);
    cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 10; ++i) cudaEventCreate(&h->ev_stage[i]);
  build_tables(h);
  if (cudaMemcpyToSymbol(c_pattern, h_pattern, sizeof(h_pattern)) != cudaSuccess) return fail(ORB_ERR_CUDA);
  if (cudaMemcpyToSymbol(c_umax, h->umax, sizeof(int) * 16) != cudaSuccess) return fail(ORB_ERR_CUDA);
  if (orb_ensure(h, h->d_pattern, sizeof(h_pattern)) != ORB_OK) return fail(ORB_ERR_CUDA);
  if (cudaMemcpy(h->d_pattern.p, h_pattern, sizeof(h_pattern), cudaMemcpyHostToDevice) != cudaSuccess) return fail(ORB_ERR_CUDA);
  {
    // float copy of the pattern for k_orient_describe, transposed so that lane i reads comparison 8 * i + j at j * 32 + i
    std::vector<float> pf(1024);
    for (int lane = 0; lane < 32; ++lane)
      for (int j = 0; j < 8; ++j)
        for (int c = 0; c < 4; ++c) pf[(size_t)(j * 32 + lane) * 4 + c] = (float)h_pattern[(8 * lane + j) * 4 + c];
    if (orb_ensure(h, h->d_pattern_f, pf.size() * sizeof(float)) != ORB_OK) return fail(ORB_ERR_CUDA);
    if (cudaMemcpy(h->d_pattern_f.p, pf.data(), pf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return fail(ORB_ERR_CUDA);
    // IC_Angle (src/ORBextractor.cc:75-99) as byte dot products: item = (row v + 15, word j) of the 31 x 48-byte box that starts at the
    // 16-byte boundary at or before cx - 15; per byte offset o = (cx - 15) & 15 the weight of byte b of word j is u = 4 j + b - o - 15
    // for m10 and v for m01 when |u| <= u_max[|v|] (the centre row takes all of -15 .. 15, :82-83), else 0 (padding items: 0)
    std::vector<uint32_t> ic((size_t)16 * ORB_IC_ITEMS * 2, 0u);
    for (int o = 0; o < 16; ++o)
      for (int i = 0; i < 31 * 12; ++i) {
        const int row = i / 12, j = i % 12, v = row - ORB_HALF_PATCH, av = v < 0 ? -v : v;
        const int d = av == 0 ? ORB_HALF_PATCH : h->umax[av];
        uint32_t wu = 0, wv = 0;
        for (int b = 0; b < 4; ++b) {
          const int u = 4 * j + b - o - ORB_HALF_PATCH;
          if (u >= -d && u <= d) {
            wu |= (uint32_t)(uint8_t)(int8_t)u << (8 * b);
            wv |= (uint32_t)(uint8_t)(int8_t)v << (8 * b);
          }
        }
        uint32_t* e = &ic[((size_t)o * ORB_IC_ITEMS + i) * 2];
        e[0] = wu; e[1] = wv;
      }
    if (orb_ensure(h, h->d_ic_tab, ic.size() * sizeof(uint32_t)) != ORB_OK) return fail(ORB_ERR_CUDA);
    if (cudaMemcpy(h->d_ic_tab.p, ic.data(), ic.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) return fail(ORB_ERR_CUDA);
  }
  int st = configure(h, max_width, max_height, max_batch);
  if (st) { fprintf(stderr, "orb_create: %s\n", h->last_error.c_str()); return fail(st); }
  *out = h;
  return ORB_OK;
}

int orb_destroy(orb_handle* h) {
  if (!h) return ORB_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  DevBuf* bufs[] = {&h->d_pyr, &h->d_blur, &h->d_tab, &h->d_blur_tiles, &h->d_pattern, &h->d_pattern_f, &h->d_ic_tab, &h->d_cell_count, &h->d_cell_keys, &h->d_lvl_count, &h->d_tree_scratch,
                    &h->d_sel_count, &h->d_sel_keys, &h->d_ord_src, &h->d_ord_dst, &h->d_kps, &h->d_desc, &h->d_n, &h->d_mono,
                    &h->d_status, &h->d_uright, &h->d_depth, &h->d_sad, &h->d_best_idx, &h->d_best_dist, &h->d_rband, &h->d_row_items,
                    &h->d_scratch, &h->d_scratch2, &h->d_grid_off, &h->d_grid_idx, &h->d_grid_cell, &h->d_sp_cand, &h->d_sp_cnt,
                    &h->d_sp_match, &h->d_sp_nm, &h->d_sp2, &h->d_bow2, &h->d_bow_fword, &h->d_bow_fnode, &h->d_bow_fw, &h->d_bow_n, &h->d_bow_word, &h->d_bow_val,
                    &h->d_fv_node, &h->d_fv_off, &h->d_fv_feat, &h->d_fast_items, &h->d_fast_spill, &h->d_raw, &h->d_mapx, &h->d_mapy, &h->d_map_tiles, &h->d_in_tab, &h->d_kps_un, &h->d_fe_idx, &h->d_fe_dist, &h->d_fe_pass,
                    &h->d_fe_l2r, &h->d_fe_r2l, &h->d_fe_depth, &h->d_fe_p3d, &h->d_fe_code, &h->d_fe_part, &h->d_fast_work};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (h->h_n) { cudaFreeHost(h->h_n); cudaFreeHost(h->h_mono); cudaFreeHost(h->h_status); }
  drop_pipeline_graph(h);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->ev_stop) cudaEventDestroy(h->ev_stop);
  if (h->ev_sync) cudaEventDestroy(h->ev_sync);
  if (h->ev_peer) cudaEventDestroy(h->ev_peer);
  for (int i = 0; i < ORB_MAX_LEVELS + 1; ++i) {
    if (h->aux[i]) { cudaStreamSynchronize(h->aux[i]); cudaStreamDestroy(h->aux[i]); }
    if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  for (int i = 0; i < 10; ++i)
    if (h->ev_stage[i]) cudaEventDestroy(h->ev_stage[i]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return ORB_OK;
}

int orb_keypoint_capacity(const orb_handle* h) { return h ? h->g.kcap : ORB_ERR_INVALID_ARG; }

int orb_get_tables(const orb_handle* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* nfeat) {
  if (!h) return ORB_ERR_INVALID_ARG;
  for (int i = 0; i < h->params.nlevels; ++i) {
    if (scale) scale[i] = h->scale[i];
    if (inv_scale) inv_scale[i] = h->inv_scale[i];
    if (sigma2) sigma2[i] = h->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
    if (nfeat) nfeat[i] = h->nfeat[i];
  }
  return ORB_OK;
}

int orb_compute_tables(const orb_params* p, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* nfeat) {
  if (!p || p->nlevels < 1 || p->nlevels > ORB_MAX_LEVELS || p->nfeatures < 1 || !(p->scale_factor > 1.0f)) return ORB_ERR_INVALID_ARG;
  std::vector<float> sc, inv, s2, is2;
  std::vector<int> nf;
  build_scale_tables(*p, sc, inv, s2, is2, nf);
  for (int i = 0; i < p->nlevels; ++i) {
    if (scale) scale[i] = sc[i];
    if (inv_scale) inv_scale[i] = inv[i];
    if (sigma2) sigma2[i] = s2[i];
    if (inv_sigma2) inv_sigma2[i] = is2[i];
    if (nfeat) nfeat[i] = nf[i];
  }
  return ORB_OK;
}

int orb_extract_batch(orb_handle* h, const uint8_t* images, int batch, int width, int height, size_t stride,
                      size_t image_stride, int lap0, int lap1, orb_keypoint* kps_out, uint8_t* desc_out, int cap,
                      int* n_out, int* mono_out, int flags) {
  if (!h) return ORB_ERR_INVALID_ARG;
  if (!images || width <= 0 || height <= 0) return orb_set_error(h, ORB_ERR_EMPTY_IMAGE, "empty image");
  if (batch < 1 || stride < (size_t)width) return orb_set_error(h, ORB_ERR_INVALID_ARG, "bad batch/stride");
  const bool remap = (flags & ORB_INPUT_REMAP) != 0, resize_in = (flags & ORB_INPUT_RESIZE) != 0;
  if (remap && resize_in) return orb_set_error(h, ORB_ERR_INVALID_ARG, "ORB_INPUT_REMAP and ORB_INPUT_RESIZE exclude each other (src/System.cc:254-268)");
  if (remap && !h->map_w) return orb_set_error(h, ORB_ERR_STATE, "ORB_INPUT_REMAP without orb_set_rectify_maps");
  if (resize_in && !h->in_w) return orb_set_error(h, ORB_ERR_STATE, "ORB_INPUT_RESIZE without orb_set_input_size");
  const int raw_w = width, raw_h = height;
  if (remap) { width = h->map_w; height = h->map_h; }
  if (resize_in) { width = h->in_w; height = h->in_h; }
  if (width > h->max_w || height > h->max_h)
    return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "image larger than the handle's max_width x max_height");
  int st;
  if ((st = orb_use_device(h))) return st;
  if (h->pending) { if ((st = finish_batch(h))) return st; }
  if ((st = configure(h, width, height, batch))) return st;
  const OrbGeom& g = h->g;
  if (!(flags & ORB_NO_OUTPUT) && (kps_out || desc_out) && cap < 1) return orb_set_error(h, ORB_ERR_INVALID_ARG, "cap < 1");
  // level 0 = the input image (the reference's copyMakeBorder at :1108-1109 is only a copy + margin)
  uint8_t* l0 = h->d_pyr.as<uint8_t>() + g.level_base[0];
  if (remap || resize_in) {
    // raw frames -> d_raw (tight rows) -> k_remap / k_resize_input -> level 0 (System::TrackStereo, src/System.cc:254-264)
    const size_t fbytes = (size_t)raw_w * raw_h;
    if ((st = orb_ensure(h, h->d_raw, fbytes * batch + 64))) return st;   // the staging loads of k_remap read up to 15 bytes past a box row
    uint8_t* d_raw = h->d_raw.as<uint8_t>();
    if (image_stride == stride * (size_t)raw_h && stride == (size_t)raw_w) {
      ORB_CUDA_CHECK(h, cudaMemcpyAsync(d_raw, images, fbytes * batch, cudaMemcpyDefault, h->stream));
    } else {
      for (int f = 0; f < batch; ++f)
        ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(d_raw + (size_t)f * fbytes, raw_w, images + (size_t)f * image_stride, stride, raw_w, raw_h,
                                            cudaMemcpyDefault, h->stream));
    }
    if (resize_in) {
      if (h->in_src_w != raw_w || h->in_src_h != raw_h) {
        std::vector<int> tab;
        axis_table(raw_w, width, true, tab);
        axis_table(raw_h, height, false, tab);
        if ((st = orb_ensure(h, h->d_in_tab, tab.size() * sizeof(int)))) return st;
        ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));   // an earlier batch may still read the old tables
        ORB_CUDA_CHECK(h, cudaMemcpy(h->d_in_tab.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
        h->in_src_w = raw_w; h->in_src_h = raw_h;
        h->in_area2x = (raw_w == 2 * width && raw_h == 2 * height) ? 1 : 0;
      }
      const dim3 blk(32, 8), grd((((width + 3) / 4) + 31) / 32, (height + 7) / 8, batch);
      k_resize_input<<<grd, blk, 0, h->stream>>>(d_raw, raw_w, raw_h, fbytes, l0, width, height, g.pitch[0], (size_t)g.level_fstride[0],
                                                  h->d_in_tab.as<int2>(), h->d_in_tab.as<int2>() + width, h->in_area2x);
    } else {
    const int tiles_x = (width + RM_TW - 1) / RM_TW, tiles_y = (height + RM_TH - 1) / RM_TH;
    k_remap<<<dim3(tiles_x * tiles_y, (batch + RM_FRAMES - 1) / RM_FRAMES), 256, 0, h->stream>>>(
        d_raw, raw_w, raw_h, (size_t)raw_w, fbytes, h->d_mapx.as<float>(), h->d_mapy.as<float>(), width, height, h->d_map_tiles.as<int4>(),
        tiles_x, l0, g.pitch[0], (size_t)g.level_fstride[0], batch);
    }
    h->launches++;
    ORB_CUDA_CHECK(h, cudaGetLastError());
  } else if (image_stride == stride * (size_t)height && stride == (size_t)width && g.pitch[0] == width) {
    // contiguous frames whose rows need no re-pitching: one linear copy (2-D copies of short rows are slow DMA)
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(l0, images, (size_t)width * height * batch, cudaMemcpyDefault, h->stream));
  } else if (image_stride == stride * (size_t)height && stride == (size_t)width && !(flags & ORB_SRC_DEVICE)) {
    // tight host frames whose width is not the level-0 pitch (e.g. KITTI, 1241 px): a 2-D host copy of short rows runs at a
    // fraction of the PCIe rate, so the frames cross the bus as one linear copy and are re-pitched device to device
    const size_t bytes = (size_t)width * height * batch;
    if ((st = orb_ensure(h, h->d_raw, bytes + 64))) return st;
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(h->d_raw.p, images, bytes, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(l0, g.pitch[0], h->d_raw.p, width, width, (size_t)height * batch, cudaMemcpyDeviceToDevice, h->stream));
  } else if (image_stride == stride * (size_t)height) {
    ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(l0, g.pitch[0], images, stride, width, (size_t)height * batch, cudaMemcpyDefault, h->stream));
  } else {
    for (int f = 0; f < batch; ++f)
      ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(l0 + (size_t)f * g.level_fstride[0], g.pitch[0], images + (size_t)f * image_stride,
                                          stride, width, height, cudaMemcpyDefault, h->stream));
  }
  // small batches (latency path) whose result buffers are page-locked: the descriptor kernel writes the records into them itself
  h->zc_kps = nullptr; h->zc_desc = nullptr; h->zc_cap = 0;
  const bool zero_copy = batch <= ORB_GRAPH_MAX_BATCH && !(flags & (ORB_DST_DEVICE | ORB_NO_OUTPUT)) && kps_out && desc_out && cap > 0 &&
                         !h->stage_timing && orb_host_buffer_is_device_writable(kps_out) && orb_host_buffer_is_device_writable(desc_out);
  if (zero_copy) { h->zc_kps = kps_out; h->zc_desc = desc_out; h->zc_cap = cap; }
  if ((st = run_pipeline(h, batch, lap0, lap1))) return st;
  h->cur_batch = batch;
  h->have_batch = true;
  h->frames_loaded = false;
  h->have_stereo = false;
  h->have_fe = false;
  h->have_fe_tri = false;
  h->have_grid = false;
  h->have_undist = false;
  h->have_bow = false;
  h->have_bow2 = false;
  h->lap0 = lap0; h->lap1 = lap1;
  // h_n / h_mono / h_status: k_assemble wrote them into the pinned words itself (no copies)
  if (!(flags & ORB_NO_OUTPUT) && !zero_copy) {
    const int rows = std::min(cap, g.kcap);
    if (cap == g.kcap) {
      if (kps_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(kps_out, h->d_kps.p, (size_t)batch * cap * sizeof(orb_keypoint), cudaMemcpyDefault, h->stream));
      if (desc_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(desc_out, h->d_desc.p, (size_t)batch * cap * 32, cudaMemcpyDefault, h->stream));
    } else {
      if (kps_out)
        ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(kps_out, (size_t)cap * sizeof(orb_keypoint), h->d_kps.p, (size_t)g.kcap * sizeof(orb_keypoint),
                                            (size_t)rows * sizeof(orb_keypoint), batch, cudaMemcpyDefault, h->stream));
      if (desc_out)
        ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(desc_out, (size_t)cap * 32, h->d_desc.p, (size_t)g.kcap * 32, (size_t)rows * 32, batch,
                                            cudaMemcpyDefault, h->stream));
    }
  }
  h->pending = true;
  h->pending_batch = batch;
  h->pending_n_out = (flags & ORB_DST_DEVICE) ? nullptr : n_out;
  h->pending_mono_out = (flags & ORB_DST_DEVICE) ? nullptr : mono_out;
  if (flags & ORB_DST_DEVICE) {
    if (n_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(n_out, h->d_n.p, batch * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    if (mono_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(mono_out, h->d_mono.p, batch * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  st = finish_batch(h);
  if (st) return st;
  if (!(flags & ORB_NO_OUTPUT) && (kps_out || desc_out)) {
    for (int i = 0; i < batch; ++i)
      if (h->h_n[i] > cap) return orb_set_error(h, ORB_ERR_CAPACITY, "caller capacity smaller than the number of keypoints");
  }
  return ORB_OK;
}

int orb_set_input_size(orb_handle* h, int new_w, int new_h) {
  if (!h) return ORB_ERR_INVALID_ARG;
  if (new_w <= 0 || new_h <= 0) { h->in_w = h->in_h = 0; return ORB_OK; }
  if (new_w > h->max_w || new_h > h->max_h)
    return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "resized image larger than the handle's max_width x max_height");
  if (new_w != h->in_w || new_h != h->in_h) h->in_src_w = h->in_src_h = 0;
  h->in_w = new_w; h->in_h = new_h;
  return ORB_OK;
}

int orb_set_rectify_maps(orb_handle* h, const float* map_x, const float* map_y, int map_w, int map_h) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  if (h->pending) { if ((st = finish_batch(h))) return st; }
  if (!map_x || !map_y) { h->map_w = h->map_h = 0; return ORB_OK; }
  if (map_w < 1 || map_h < 1) return orb_set_error(h, ORB_ERR_INVALID_ARG, "bad map size");
  if (map_w > h->max_w || map_h > h->max_h)
    return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "rectified image larger than the handle's max_width x max_height");
  const size_t bytes = (size_t)map_w * map_h * sizeof(float);
  if ((st = orb_ensure(h, h->d_mapx, bytes)) || (st = orb_ensure(h, h->d_mapy, bytes))) return st;
  // integer tap bounds of every 64 x 16 destination tile (unclamped, the kernel clamps them to the raw image it is given)
  const int tiles_x = (map_w + RM_TW - 1) / RM_TW, tiles_y = (map_h + RM_TH - 1) / RM_TH;
  std::vector<int> tb((size_t)tiles_x * tiles_y * 4);
  for (int ty = 0; ty < tiles_y; ++ty)
    for (int tx = 0; tx < tiles_x; ++tx) {
      int lo_x = INT_MAX, hi_x = INT_MIN, lo_y = INT_MAX, hi_y = INT_MIN;
      for (int y = ty * RM_TH; y < std::min((ty + 1) * RM_TH, map_h); ++y)
        for (int x = tx * RM_TW; x < std::min((tx + 1) * RM_TW, map_w); ++x) {
          const float vx = map_x[(size_t)y * map_w + x] * 32.f, vy = map_y[(size_t)y * map_w + x] * 32.f;
          const int fsx = (vx >= -2147483648.f && vx < 2147483648.f) ? (int)lrintf(vx) : INT_MIN;   // cvtss2si
          const int fsy = (vy >= -2147483648.f && vy < 2147483648.f) ? (int)lrintf(vy) : INT_MIN;
          const int sx = rm_sat_short(fsx >> 5), sy = rm_sat_short(fsy >> 5);
          lo_x = std::min(lo_x, sx); hi_x = std::max(hi_x, sx); lo_y = std::min(lo_y, sy); hi_y = std::max(hi_y, sy);
        }
      int* o = &tb[((size_t)ty * tiles_x + tx) * 4];
      o[0] = lo_x; o[1] = hi_x; o[2] = lo_y; o[3] = hi_y;
    }
  if ((st = orb_ensure(h, h->d_map_tiles, tb.size() * sizeof(int)))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(h->d_map_tiles.p, tb.data(), tb.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(h->d_mapx.p, map_x, bytes, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(h->d_mapy.p, map_y, bytes, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  h->map_w = map_w; h->map_h = map_h;
  return ORB_OK;
}

int orb_sync(orb_handle* h) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  return finish_batch(h);
}

int orb_extract(orb_handle* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1,
                orb_keypoint* kps_out, uint8_t* desc_out, int cap, int* n_out, int* mono_out) {
  if (!h) return ORB_ERR_INVALID_ARG;
  if (n_out) *n_out = 0;
  if (mono_out) *mono_out = -1;
  if (!image || width <= 0 || height <= 0) return orb_set_error(h, ORB_ERR_EMPTY_IMAGE, "empty image");
  int n = 0, mono = 0;
  int st = orb_extract_batch(h, image, 1, width, height, stride, stride * (size_t)height, lap0, lap1, kps_out, desc_out, cap,
                             &n, &mono, 0);
  if (st) return st;
  if (n_out) *n_out = n;
  if (mono_out) *mono_out = mono;
  return ORB_OK;
}

int orb_pyramid_level_size(const orb_handle* h, int level, int* width, int* height) {
  if (!h || level < 0 || level >= h->params.nlevels) return ORB_ERR_INVALID_ARG;
  if (width) *width = h->g.w[level];
  if (height) *height = h->g.h[level];
  return ORB_OK;
}

static int copy_level(orb_handle* h, const DevBuf& buf, int frame, int level, uint8_t* dst, size_t dst_stride) {
  if (!h || !dst) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  if (h->frames_loaded) return orb_set_error(h, ORB_ERR_STATE, "the resident frames were loaded with orb_load_frames: no pyramid");
  if (level < 0 || level >= h->g.nlevels || frame < 0 || frame >= h->cur_batch) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const OrbGeom& g = h->g;
  if (dst_stride == 0) dst_stride = g.w[level];
  const uint8_t* src = buf.as<uint8_t>() + g.level_base[level] + (size_t)frame * g.level_fstride[level];
  ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(dst, dst_stride, src, g.pitch[level], g.w[level], g.h[level], cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_pyramid_level(orb_handle* h, int frame, int level, uint8_t* dst, size_t dst_stride) {
  return copy_level(h, h->d_pyr, frame, level, dst, dst_stride);
}
int orb_debug_get_blurred(orb_handle* h, int frame, int level, uint8_t* dst, size_t dst_stride) {
  return copy_level(h, h->d_blur, frame, level, dst, dst_stride);
}

int orb_debug_get_candidates(orb_handle* h, int frame, int level, int32_t* xys, int cap, int* n_out) {
  if (!h || !xys || !n_out) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  const OrbGeom& g = h->g;
  if (level < 0 || level >= g.nlevels || frame < 0 || frame >= h->cur_batch) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const int cells = g.cell_start[g.nlevels];
  const int c0 = g.cell_start[level], c1 = g.cell_start[level + 1];
  std::vector<int> cnt(c1 - c0);
  std::vector<uint32_t> keys((size_t)(c1 - c0) * ORB_CELL_CAP);
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy(cnt.data(), h->d_cell_count.as<int>() + (size_t)frame * cells + c0, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
  ORB_CUDA_CHECK(h, cudaMemcpy(keys.data(), h->d_cell_keys.as<uint32_t>() + ((size_t)frame * cells + c0) * ORB_CELL_CAP,
                               keys.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  int n = 0;
  for (int c = 0; c < c1 - c0; ++c)
    for (int i = 0; i < cnt[c]; ++i) {
      if (n >= cap) return orb_set_error(h, ORB_ERR_CAPACITY, "candidate buffer too small");
      const uint32_t k = keys[(size_t)c * ORB_CELL_CAP + i];
      xys[3 * n] = orb_px(k); xys[3 * n + 1] = orb_py(k); xys[3 * n + 2] = orb_ps(k);
      ++n;
    }
  *n_out = n;
  return ORB_OK;
}

int orb_debug_get_level_counts(orb_handle* h, int32_t* counts, int cap) {
  if (!h || !counts) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  const int n = h->cur_batch * h->g.nlevels;
  if (cap < n) return orb_set_error(h, ORB_ERR_CAPACITY, "counts buffer too small");
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy(counts, h->d_lvl_count.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return ORB_OK;
}

int orb_debug_get_selected(orb_handle* h, int frame, int level, int32_t* xys, int cap, int* n_out) {
  if (!h || !xys || !n_out) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  const OrbGeom& g = h->g;
  if (level < 0 || level >= g.nlevels || frame < 0 || frame >= h->cur_batch) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  int n = 0;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy(&n, h->d_sel_count.as<int>() + (size_t)frame * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  if (n > cap) return orb_set_error(h, ORB_ERR_CAPACITY, "selected buffer too small");
  std::vector<uint32_t> keys(std::max(n, 1));
  ORB_CUDA_CHECK(h, cudaMemcpy(keys.data(), h->d_sel_keys.as<uint32_t>() + ((size_t)frame * g.nlevels + level) * g.lvl_kcap,
                               (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) { xys[3 * i] = orb_px(keys[i]); xys[3 * i + 1] = orb_py(keys[i]); xys[3 * i + 2] = orb_ps(keys[i]); }
  *n_out = n;
  return ORB_OK;
}

int orb_debug_std_sort(orb_handle* h, const uint32_t* keys, int n, uint32_t* keys_out, uint32_t* payload_out) {
  if (!h || !keys || !keys_out || !payload_out || n < 0) return ORB_ERR_INVALID_ARG;
  if (n > 8192) return orb_set_error(h, ORB_ERR_CAPACITY, "at most 8192 records");
  if (n == 0) return ORB_OK;
  int st;
  if ((st = orb_use_device(h))) return st;
  if (h->pending) { if ((st = finish_batch(h))) return st; }
  if ((st = orb_ensure(h, h->d_scratch, (size_t)n * 12))) return st;
  uint32_t* d = h->d_scratch.as<uint32_t>();
  const size_t smem = (size_t)n * 24 + 16;
  if ((st = orb_raise_dyn_smem(h, (const void*)k_debug_std_sort, smem))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(d, keys, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  k_debug_std_sort<<<1, OP_THREADS, smem, h->stream>>>(d, n, d + n, d + 2 * n);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(keys_out, d + n, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(payload_out, d + 2 * n, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_debug_distribute(orb_handle* h, const int32_t* cands, int n, int region_w, int region_h, int N, int32_t* out, int cap,
                         int* n_out) {
  if (!h || !cands || !out || !n_out || n < 0 || N < 1 || region_w < 1 || region_h < 1) return ORB_ERR_INVALID_ARG;
  if (n > ORB_LEVEL_CAP || region_w > ORB_MAX_DIM || region_h > ORB_MAX_DIM) return orb_set_error(h, ORB_ERR_CAPACITY, "too many candidates");
  int st;
  if ((st = orb_use_device(h))) return st;
  if (h->pending) { if ((st = finish_batch(h))) return st; }
  OrbGeom g;
  std::memset(&g, 0, sizeof(g));
  g.nlevels = 1;
  g.w[0] = region_w + 2 * ORB_BORDER; g.h[0] = region_h + 2 * ORB_BORDER;
  g.nfeat[0] = N;
  g.nini[0] = (int)std::round((float)region_w / (float)region_h);
  if (g.nini[0] < 1) return orb_set_error(h, ORB_ERR_UNSUPPORTED_SIZE, "region narrower than half its height");
  g.hx[0] = (float)region_w / g.nini[0];
  const int max_root = std::max(N, 4 * g.nini[0]);
  g.lvl_kcap = max_root + 8;
  g.node_cap = max_root + 16;
  g.level_cap[0] = ORB_LEVEL_CAP;
  g.scratch_off[0] = 0;
  g.scratch_frame = 2 * ORB_LEVEL_CAP;
  const int dbg_keys_cap = std::min(ORB_TREE_SMEM_KEYS, (8 * N + 512 + 31) & ~31);
  const size_t smem = octree_smem_bytes(g.node_cap, dbg_keys_cap);
  if (smem > 227 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "N too large");
  std::vector<uint32_t> keys(std::max(n, 1));
  for (int i = 0; i < n; ++i) keys[i] = orb_pack(cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]);
  // scratch: [keys n][tree scratch 2*LEVEL_CAP][sel lvl_kcap][sel_count 1][status 1][lvl_count 1]
  const size_t words = (size_t)ORB_LEVEL_CAP + 2 * ORB_LEVEL_CAP + g.lvl_kcap + 8;
  if ((st = orb_ensure(h, h->d_scratch, words * 4))) return st;
  uint32_t* d = h->d_scratch.as<uint32_t>();
  uint32_t* d_keys = d; uint32_t* d_tree = d + ORB_LEVEL_CAP; uint32_t* d_sel = d_tree + 2 * ORB_LEVEL_CAP;
  int* d_cnt = (int*)(d_sel + g.lvl_kcap); int* d_stat = d_cnt + 1; int* d_lvl = d_cnt + 2;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(d_keys, keys.data(), (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaMemsetAsync(d_cnt, 0, 16, h->stream));
  if ((st = orb_raise_dyn_smem(h, (const void*)k_octree, std::max(smem, octree_smem_max(h->g))))) return st;
  if (h->octree_passes) {
    const size_t sm = octree_passes_smem_bytes(g.node_cap, dbg_keys_cap);
    if ((st = orb_raise_dyn_smem(h, (const void*)k_octree_passes, sm))) return st;
    k_octree_passes<<<1, OP_THREADS, sm, h->stream>>>(g, nullptr, nullptr, 0, d_tree, d_lvl, d_cnt, d_sel, d_stat, 0, g.node_cap, dbg_keys_cap, d_keys, n);
  } else
  k_octree<<<1, 32, smem, h->stream>>>(g, nullptr, nullptr, 0, d_tree, d_lvl, d_cnt, d_sel, d_stat, 0, g.node_cap, dbg_keys_cap, d_keys, n);
  h->launches++;
  int res[2] = {0, 0};
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(res, d_cnt, 8, cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  if (res[1]) return orb_set_error(h, ORB_ERR_CAPACITY, "quad-tree capacity exceeded");
  if (res[0] > cap) return orb_set_error(h, ORB_ERR_CAPACITY, "output buffer too small");
  std::vector<uint32_t> sel(std::max(res[0], 1));
  ORB_CUDA_CHECK(h, cudaMemcpy(sel.data(), d_sel, (size_t)res[0] * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < res[0]; ++i) { out[3 * i] = orb_px(sel[i]); out[3 * i + 1] = orb_py(sel[i]); out[3 * i + 2] = orb_ps(sel[i]); }
  *n_out = res[0];
  return ORB_OK;
}

int orb_hamming_distance(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    d += __builtin_popcount(x ^ y);
  }
  return d;
}

int orb_host_alloc(void** p, size_t bytes) {
  if (!p) return ORB_ERR_INVALID_ARG;
  return cudaMallocHost(p, bytes) == cudaSuccess ? ORB_OK : ORB_ERR_CUDA;
}
int orb_host_alloc_ex(void** p, size_t bytes, int write_combined) {
  if (!p) return ORB_ERR_INVALID_ARG;
  return cudaHostAlloc(p, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess ? ORB_OK : ORB_ERR_CUDA;
}
int orb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? ORB_OK : ORB_ERR_CUDA; }
int orb_device_alloc(orb_handle* h, void** p, size_t bytes) {
  if (!h || !p) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaMalloc(p, bytes));
  return ORB_OK;
}
int orb_device_free(orb_handle* h, void* p) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaFree(p));
  return ORB_OK;
}
int orb_memcpy_h2d(orb_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}
int orb_memcpy_d2h(orb_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_memcpy_h2d_async(orb_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return ORB_OK;
}
int orb_memcpy_d2h_async(orb_handle* h, void* dst, const void* src, size_t bytes) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  return ORB_OK;
}

int orb_timer_start(orb_handle* h) {
  if (!h) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaEventRecord(h->ev_start, h->stream));
  return ORB_OK;
}
int orb_timer_stop(orb_handle* h, float* ms_out) {
  if (!h || !ms_out) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaEventRecord(h->ev_stop, h->stream));
  ORB_CUDA_CHECK(h, cudaEventSynchronize(h->ev_stop));
  ORB_CUDA_CHECK(h, cudaEventElapsedTime(ms_out, h->ev_start, h->ev_stop));
  return ORB_OK;
}
int64_t orb_launch_count(const orb_handle* h) { return h ? h->launches : -1; }
int orb_set_stage_timing(orb_handle* h, int enabled) {
  if (!h) return ORB_ERR_INVALID_ARG;
  h->stage_timing = enabled != 0;
  return ORB_OK;
}
int orb_get_stage_times(orb_handle* h, float* ms8) {
  if (!h || !ms8) return ORB_ERR_INVALID_ARG;
  for (int i = 0; i < 8; ++i) ms8[i] = h->stage_ms[i];
  return ORB_OK;
}

}  // extern "C"
