// Internal definitions shared by the CUDA translation units of liborb_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "orb_b200.h"

#define ORB_BORDER 16        // EDGE_THRESHOLD - 3, FAST working border (src/ORBextractor.cc:747)
#define ORB_EDGE 19          // EDGE_THRESHOLD (src/ORBextractor.cc:73)
#define ORB_HALF_PATCH 15    // HALF_PATCH_SIZE (:72)
#define ORB_CELL_CAP 512     // FAST candidates kept per 35-px cell (overflow -> ORB_ERR_CAPACITY)
#define ORB_LEVEL_CAP 65535  // FAST candidates per (frame, level) the quad-tree accepts
#define ORB_TREE_SMEM_KEYS 3072  // candidates per level that fit the shared-memory fast path
#define ORB_MAX_DIM 4095     // 12-bit packed coordinates
#define ORB_ROI_MAX 80       // largest FAST cell ROI side (cell + 6) the tile kernel stages

// per-frame status bits (device side)
#define ORB_ST_CELL_OVERFLOW 1
#define ORB_ST_LEVEL_OVERFLOW 2
#define ORB_ST_NODE_OVERFLOW 4
#define ORB_ST_OUT_OVERFLOW 8

// packed FAST candidate / keypoint: x (12 bits) | y (12 bits) << 12 | score (8 bits) << 24,
// x,y relative to the 16-px border of the level
__host__ __device__ inline uint32_t orb_pack(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__host__ __device__ inline int orb_px(uint32_t k) { return (int)(k & 0xfffu); }
__host__ __device__ inline int orb_py(uint32_t k) { return (int)((k >> 12) & 0xfffu); }
__host__ __device__ inline int orb_ps(uint32_t k) { return (int)(k >> 24); }

// Geometry of one batch (same for every frame of the batch); passed to kernels by value.
struct OrbGeom {
  int nlevels;
  int w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS], pitch[ORB_MAX_LEVELS];
  // pyramid storage is level-major: level l of frame f starts at level_base[l] + f * level_fstride[l]
  // (all frames of one level are contiguous, so a whole batch uploads into level 0 with one copy)
  unsigned long long level_base[ORB_MAX_LEVELS];
  unsigned long long level_fstride[ORB_MAX_LEVELS];  // pitch[l] * h[l]
  int batch_cap;                                      // frames the level regions are laid out for
  // FAST cell grid (src/ORBextractor.cc:755-761)
  int ncols[ORB_MAX_LEVELS], nrows[ORB_MAX_LEVELS], wcell[ORB_MAX_LEVELS], hcell[ORB_MAX_LEVELS];
  int cell_start[ORB_MAX_LEVELS + 1];            // first cell index of each level inside a frame
  // quad-tree
  int nfeat[ORB_MAX_LEVELS];                     // mnFeaturesPerLevel
  int nini[ORB_MAX_LEVELS];                      // round(w/h) of the FAST region (:545)
  float hx[ORB_MAX_LEVELS];                      // (float)w / nIni (:547)
  int level_cap[ORB_MAX_LEVELS];                 // most candidates level l can hold: min(ORB_LEVEL_CAP, cells * ORB_CELL_CAP)
  unsigned int scratch_off[ORB_MAX_LEVELS];      // key offset of level l's global ping-pong buffers inside a frame's scratch
  unsigned int scratch_frame;                    // keys of tree scratch per frame (both buffers)
  int lvl_kcap;                                  // selected keypoints kept per (frame, level)
  int node_cap;                                  // quad-tree node slots
  int kcap;                                      // keypoints per frame (nfeatures + 3 * nlevels)
  // per-level constants
  float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS];
  int patch_size[ORB_MAX_LEVELS];                // (int)(31 * scale) (:826)
  int ini_th, min_th;
  // tile tables for the per-pixel kernels: tiles of all levels flattened into one grid dimension
  int blur_tile_start[ORB_MAX_LEVELS + 1], blur_tiles_x[ORB_MAX_LEVELS];
  int pyr_tile_start[ORB_MAX_LEVELS + 1], pyr_tiles_x[ORB_MAX_LEVELS];
};

// Host-computed launch geometry of one level of the FAST tile kernel (orb_kernel_fast.cuh).
struct FastTileGeom {
  int nbx, nby;           // cells per tile
  int bh;                 // TMA box height = nby * hcell + 6
  int sp;                 // score map pitch (bytes, multiple of 4)
  int wpr;                // mask words per cell row
  int list_cap;           // entries of the corner list (pixels of the tile)
  int list1_cap;          // entries of the word list (4-pixel words of the tile)
  unsigned mul_w, mul_h;  // ceil(65536 / wcell), ceil(65536 / hcell): x / wcell == (x * mul_w) >> 16 for x < 885
};

struct BlurMaps {
  CUtensorMap m[ORB_MAX_LEVELS];
};

// FAST cell kernel (orb_kernel_fast_cells.cuh): one TMA descriptor per level, host-computed constants
struct FastMaps {
  CUtensorMap m[ORB_MAX_LEVELS];
};
struct FastCellGeom {
  int G[ORB_MAX_LEVELS];     // cells per item of level l (1 or 2)
  int PW[ORB_MAX_LEVELS];    // tile pitch in 32-bit words = TMA box width / 4 (16..32)
  int SP[ORB_MAX_LEVELS];    // score map pitch in bytes (multiple of 4, >= wcell + 2)
  int SCELL[ORB_MAX_LEVELS]; // bytes of one cell's score map (multiple of 16)
  int WPR[ORB_MAX_LEVELS];   // mask words per cell row
  unsigned MPW[ORB_MAX_LEVELS];  // ceil(2^20 / PW): idx / PW == (idx * MPW) >> 20 for idx < 2^15
  int items_per_frame;
  // per-warp shared-memory slice (sizes = maximum over the levels): valid-byte masks (128 B), tile (at +128), score maps
  // (= word list of passes A / B), row masks, corner list, mbarrier
  int score_off, mask_off, list_off, bar_off, warp_stride;
  int spill_cap;             // u16 corner-list entries per warp in global memory
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  template <typename T> T* as() const { return (T*)p; }
};

struct orb_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux[ORB_MAX_LEVELS + 1] = {nullptr};      // 0: blur, 1 + l: quad-tree of level l
  cudaEvent_t ev_fork[ORB_MAX_LEVELS + 1] = {nullptr}, ev_join[ORB_MAX_LEVELS + 1] = {nullptr};
  orb_params params{};
  int max_w = 0, max_h = 0, max_batch = 0;
  std::string last_error;
  int64_t launches = 0;

  // host tables (reference ctor)
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int> nfeat;
  int umax[16];

  // geometry of the current batch
  OrbGeom g{};
  int cur_w = 0, cur_h = 0, cur_batch = 0;
  int tab_w = 0, tab_h = 0;       // image size the resize tables were built for
  bool have_batch = false;
  bool frames_loaded = false;   // the resident batch came from orb_load_frames: keypoints / descriptors only, the pyramids are stale
  bool have_stereo = false;
  int lap0 = 0, lap1 = 0;
  int xtab_off[ORB_MAX_LEVELS], ytab_off[ORB_MAX_LEVELS];  // offsets (in int2) into d_tab
  int area2x[ORB_MAX_LEVELS];

  // FAST tile kernel: one TMA descriptor per level over that level's frames (re-encoded when d_pyr moves)
  CUtensorMap tmap_fast[ORB_MAX_LEVELS];
  BlurMaps blur_maps;                        // source of k_blur7: level l of d_pyr, box BLUR_TP x BLUR_TR
  BlurMaps desc_maps;                        // blurred patch of k_orient_describe: level l of d_blur, box 64 x 37
  BlurMaps ic_maps;                          // its moment disc: level l of d_pyr, box 48 x 31
  CUtensorMap tmap_resize[ORB_MAX_LEVELS];   // source window of k_resize_tiles: level l - 1, box rs_bw x rs_bh
  int rs_bw[ORB_MAX_LEVELS], rs_bh[ORB_MAX_LEVELS], rs_tiles[ORB_MAX_LEVELS];
  FastTileGeom ftg[ORB_MAX_LEVELS];
  // FAST cell kernel (default): descriptors, constants, item table, spill buffer of the corner lists
  FastMaps fast_maps;
  FastCellGeom fcg{};
  DevBuf d_fast_items;   // uint32 [items_per_frame] item codes (level | cell row << 4 | first column << 12 | cells << 20)
  DevBuf d_fast_spill;   // uint16 [grid warps][spill_cap]
  int fast_mode = 1;     // 1: k_fast_cells (warp per item), 0: k_fast_tiles (round-1 CTA per tile; ORB_B200_FAST=tiles)
  int fc_wpc_max = 1;    // most warps per CTA that leave two CTAs per SM
  int fc_lvl_first[ORB_MAX_LEVELS] = {0}, fc_lvl_items[ORB_MAX_LEVELS] = {0};   // item list of k_fast_cells: first item / items of level l
  bool fast_dynamic = true;  // large batches: k_fast_cells hands its items out behind an atomic counter (ORB_B200_FAST_STATIC=1: fixed stride)
  DevBuf d_fast_work;        // that counter
  bool level_pipe = true;   // small batches: FAST + quad-tree of a level start as soon as the level exists (ORB_B200_LEVEL_PIPE=0: off)
  int sm_count = 148;

  // device buffers (grown on demand, sized in orb_create for max_width x max_height x max_batch)
  DevBuf d_pyr;        // un-blurred pyramids, one slab per frame
  DevBuf d_blur;       // blurred pyramids
  DevBuf d_pattern;    // rBRIEF pattern, 1024 int8
  DevBuf d_pattern_f;  // the same as float4 (x0 y0 x1 y1) per comparison, transposed: entry j * 32 + lane = comparison 8 * lane + j
  DevBuf d_ic_tab;     // uint2 [4][ORB_IC_ITEMS]: IC_Angle u weights and in-disc masks per word alignment (orb_kernel_describe.cuh)
  DevBuf d_blur_tiles; // blur tile table: blockIdx.x -> level | tile column << 4 | tile row << 16
  DevBuf d_tab;        // resize tables: int2 (offset, c0 | c1 << 16) per destination column / row and level
  DevBuf d_cell_count; // int [batch][cells]
  DevBuf d_cell_keys;  // uint32 [batch][cells][ORB_CELL_CAP]
  DevBuf d_lvl_count;  // int [batch][levels] FAST candidates per level
  DevBuf d_tree_scratch;  // uint32 [batch][scratch_frame] global fallback key buffers of the quad-tree
  DevBuf d_sel_count;  // int [batch][levels]
  DevBuf d_sel_keys;   // uint32 [batch][levels][lvl_kcap]
  DevBuf d_ord_src;    // uint32 [batch][kcap] packed keypoint (x, y, score) per output ordinal
  DevBuf d_ord_dst;    // int [batch][kcap] (destination slot << 4 | level) per ordinal
  DevBuf d_kps;        // orb_keypoint [batch][kcap]
  DevBuf d_desc;       // uint8 [batch][kcap][32]
  DevBuf d_n, d_mono, d_status;  // int [batch]
  // stereo
  DevBuf d_uright, d_depth;      // float [batch][kcap]
  DevBuf d_sad, d_best_idx, d_best_dist;  // int [batch][kcap]
  DevBuf d_rband;      // int [batch][H + 1] row table offsets of the right keypoints
  DevBuf d_row_items;  // uint2 [batch][items_cap] right keypoints grouped by image row: index | octave << 16, x
  // fisheye stereo (orb_knn.cu: k_fisheye_knn2): int [batch][kcap][2] train index / distance, uint8 [batch][kcap] ratio test
  DevBuf d_fe_idx, d_fe_dist, d_fe_pass, d_fe_part;   // d_fe_part: per-chunk top-2 keys of the small-batch fisheye kNN
  // fisheye triangulation (orb_fisheye.cu): mvLeftToRightMatch / mvRightToLeftMatch / mvDepth / mvStereo3Dpoints / reject code
  DevBuf d_fe_l2r, d_fe_r2l, d_fe_depth, d_fe_p3d, d_fe_code;
  bool have_fe = false;   // orb_stereo_fisheye_match_batch ran on the current batch
  bool have_fe_tri = false;   // orb_stereo_fisheye_triangulate_batch ran on it: d_fe_l2r / d_fe_r2l hold mvLeftToRightMatch / mvRightToLeftMatch
  // windowed matcher (orb_match.cu)
  DevBuf d_grid_off;   // int [batch][3073] CSR offsets of the 64 x 48 grid, cell = ix * 48 + iy
  DevBuf d_grid_idx;   // uint16 [batch][kcap] keypoint indices grouped by cell, ascending inside a cell
  DevBuf d_grid_cell;  // uint16 [batch][kcap] cell of every keypoint (0xffff = outside the grid)
  DevBuf d_sp_cand;    // uint32 [batch][qcap][4] best candidates per query (distance << 16 | keypoint)
  DevBuf d_sp_cnt;     // uint8 [batch][qcap] candidates with distance <= TH_HIGH (255 = list overflow)
  DevBuf d_sp_match, d_sp_nm;  // int [batch][kcap], int [batch]
  DevBuf d_sp2;        // two-camera searches: right camera's candidates / counts / matches, split queries, area flags
  // bag of words (orb_bow.cu)
  DevBuf d_bow_fword, d_bow_fnode, d_bow_fw;   // int / int / double [batch][kcap]: word, node, weight of every feature
  DevBuf d_bow_n;                             // int [2][batch]: BowVector sizes, FeatureVector sizes
  DevBuf d_bow_word, d_bow_val;               // uint32 / double [batch][kcap] in ascending word order
  DevBuf d_fv_node, d_fv_off, d_fv_feat;      // uint32 [batch][kcap], int [batch][kcap + 1], uint32 [batch][kcap]
  DevBuf d_kps_un;     // orb_keypoint [batch][kcap] Frame::mvKeysUn when orb_undistort_keypoints ran on the batch
  bool have_undist = false;
  bool have_bow = false;   // orb_compute_bow ran on the current batch (d_fv_* hold the frames' FeatureVectors)
  // two-camera frames: ComputeBoW over the left descriptors followed by the right ones (orb_compute_bow_stereo, on the left handle)
  DevBuf d_bow2;
  uint8_t* bow2_r[16] = {nullptr};   // regions of d_bow2: see orb_bow.cu
  int bow2_cap = 0;                  // entries per frame of the combined vectors: kcap of hL + kcap of hR
  bool have_bow2 = false;
  orb_grid_params grid_params{};
  bool have_grid = false;
  // generic scratch (kNN, debug uploads)
  // input rectification (orb_kernel_remap.cuh)
  DevBuf d_raw;        // uint8 [batch][raw_h][raw_w] raw camera frames
  DevBuf d_mapx, d_mapy;  // float [map_h][map_w]
  DevBuf d_map_tiles;     // int4 per 64 x 16 destination tile: min / max integer tap column and row
  int map_w = 0, map_h = 0;
  // input resize (k_resize_input): target size, the raw size the tables were built for, int2 tables (x then y)
  int in_w = 0, in_h = 0, in_src_w = 0, in_src_h = 0, in_area2x = 0;
  DevBuf d_in_tab;
  DevBuf d_scratch, d_scratch2;
  // pinned host mirrors
  int* h_n = nullptr;
  int* h_mono = nullptr;
  int* h_status = nullptr;
  int h_cap = 0;
  // captured pipeline of small batches (orb_extract.cu: run_pipeline)
  cudaGraphExec_t pipe_exec = nullptr;
  int pipe_batch = -1, pipe_lap0 = 0, pipe_lap1 = 0;
  // small batches with page-locked result buffers: the descriptor kernel writes keypoints / descriptors straight into them (no copies)
  orb_keypoint* zc_kps = nullptr; uint8_t* zc_desc = nullptr; int zc_cap = 0;          // of the extraction being enqueued (null: copies)
  orb_keypoint* pipe_zc_kps = nullptr; uint8_t* pipe_zc_desc = nullptr; int pipe_zc_cap = 0;   // what the captured graph holds
  int seen_batch = -1, seen_lap0 = 0, seen_lap1 = 0;
  unsigned long long seen_gen = 0;
  unsigned long long pipe_gen = 0, geom_gen = 1;
  int64_t pipe_launches = 0;
  bool graph_disabled = false;
  bool octree_passes = true;    // block-parallel quad-tree k_octree_passes (default); ORB_B200_OCTREE=warp: one-warp list kernel k_octree
  // pending async completion
  int* pending_n_out = nullptr;
  int* pending_mono_out = nullptr;
  int pending_batch = 0;
  bool pending = false;
  // timing
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_sync = nullptr, ev_peer = nullptr;
  bool stage_timing = false;
  cudaEvent_t ev_stage[10] = {nullptr};
  float stage_ms[8] = {0};
};

int orb_ensure(orb_handle* h, DevBuf& b, size_t bytes);
int orb_use_device(orb_handle* h);
// Kernels on `reader`'s stream are about to read `owner`'s device-resident results (stereo / fisheye matchers): begin orders
// reader's stream after everything queued on owner's stream; end (after the last such kernel) orders owner's stream after the
// reader's kernels, so that the owner's next extraction cannot overwrite buffers that are still being read (ORB_ASYNC, R-then-L
// call order, the reference's two extraction threads). No-ops when both are the same handle.
int orb_peer_read_begin(orb_handle* reader, orb_handle* owner);
int orb_peer_read_end(orb_handle* reader, orb_handle* owner);
// raise (never lower) a kernel's dynamic shared-memory limit under a process-wide lock (orb_extract.cu)
int orb_raise_dyn_smem(orb_handle* h, const void* func, size_t bytes);
bool orb_host_buffer_is_device_writable(const void* p);   // page-locked host memory the device can write at the same address
#define ORB_SMALL_BATCH 8   // batches up to this size are latency paths: per-level branches, graph replay, results written straight to page-locked buffers

// error helpers -------------------------------------------------------------------------------
int orb_set_error(orb_handle* h, int status, const std::string& msg);
#define ORB_CUDA_CHECK(h, call)                                                                 \
  do {                                                                                          \
    cudaError_t _e = (call);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return orb_set_error((h), ORB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

