// liborb_b200.so - bag of words on the device-resident descriptors of the extractor:
//   Frame::ComputeBoW                                  reference src/Frame.cc:822-827
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
//                                                      reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1200
//   TemplatedVocabulary::transform(feature, word, weight, nid, levelsup)   (the tree descent)        :1217-1260
//   FORB::distance                                     reference Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
//   BowVector::addWeight / addIfNotExist / normalize   reference Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-98
//   FeatureVector::addFeature                          reference Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:34-48
//   TemplatedVocabulary::loadFromTextFile              reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1426 (host)
//
// Layout: the tree is stored by CHILD SLOT - the children of a node occupy consecutive slots in the order
// loadFromTextFile appends them (ascending node id), each slot holds the child's node id and its 32-byte descriptor, so
// the lanes of a group read the k descriptors of one level as one contiguous run.
//   k_bow_descend   a group of G lanes (G = 4 .. 32, the smallest power of two >= k) per feature: lane j takes child j,
//                   256-bit Hamming with __popc, redux.min of (distance << 8 | j) = first minimum of the reference's
//                   strict "<" scan; L dependent levels. Writes word, node at level L - levelsup and the word weight.
//   k_bow_assemble  one CTA per frame: the std::map insertions become two bitonic sorts in shared memory, by
//                   (word, feature) and by (node, feature); a word's value is its weight added once per feature in
//                   feature order (addWeight), the norm is summed in ascending word order by ONE thread because the
//                   order of the double additions is part of the result (BowVector::normalize iterates the map).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "orb_internal.h"

struct orb_vocab {
  int device = 0;
  int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, n_words = 0, max_children = 0;
  int* d_child_start = nullptr;      // [n_nodes + 1] first child slot of every node
  int* d_child_id = nullptr;         // [n_nodes - 1] node id per slot
  uint4* d_child_desc = nullptr;     // [n_nodes - 1][2] descriptor per slot
  unsigned int* d_word = nullptr;    // [n_nodes] word id (0 unless the file flags the node as a leaf, like Node::word_id)
  double* d_weight = nullptr;        // [n_nodes]
};

// ---- tree descent ----------------------------------------------------------------------------------------------
#ifndef BOW_MINB
#define BOW_MINB 6   // six dependent levels per feature: 40 registers for 48 resident warps per SM (0.33 -> 0.25 ms per 256 frames; 8: 0.31)
#endif
template <int G>
__global__ void __launch_bounds__(256, BOW_MINB) k_bow_descend(const uint8_t* __restrict__ desc, const int* __restrict__ n_arr, int cap,
                                                     const int* __restrict__ child_start, const int* __restrict__ child_id,
                                                     const uint4* __restrict__ child_desc, const unsigned int* __restrict__ word,
                                                     const double* __restrict__ weight, int nid_level, int* __restrict__ feat_word,
                                                     int* __restrict__ feat_node, double* __restrict__ feat_w) {
  const int frame = blockIdx.y;
  const int f = (blockIdx.x * 256 + threadIdx.x) / G, j = threadIdx.x & (G - 1);
  const int n = min(n_arr[frame], cap);
  if (f >= n) return;                       // whole groups leave together (256 % G == 0)
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const size_t fo = (size_t)frame * cap + f;
  const uint4* d = reinterpret_cast<const uint4*>(desc + fo * 32);
  const uint4 a0 = d[0], a1 = d[1];
  int node = 0, level = 0, nid = 0;         // nid_level <= 0 -> root (:1228)
  int b = child_start[0], e = child_start[1];
  while (e > b) {                           // do { } while (!isLeaf()) (:1233-1256); the root of a non-empty vocabulary has children
    ++level;
    unsigned int best = 0xffffffffu;
    for (int c = b + j; c < e; c += G) {
      const uint4 b0 = child_desc[2 * (size_t)c], b1 = child_desc[2 * (size_t)c + 1];
      const unsigned int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
                                __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
      best = min(best, (dist << 8) | (unsigned int)(c - b));   // first minimum in child order (strict "<", :1244)
    }
    best = __reduce_min_sync(gmask, best);
    node = child_id[b + (int)(best & 0xffu)];
    if (level == nid_level) nid = node;
    b = child_start[node]; e = child_start[node + 1];
  }
  // a leaf above nid_level leaves *nid unset in the reference (uninitialised local, :1149): defined here as that leaf
  if (nid_level > 0 && level < nid_level) nid = node;
  if (j == 0) {
    const double w = weight[node];
    feat_w[fo] = w;
    feat_word[fo] = w > 0 ? (int)word[node] : -1;             // w > 0: not stopped (:1157)
    feat_node[fo] = w > 0 ? nid : -1;
  }
}

// ---- BowVector / FeatureVector assembly --------------------------------------------------------------------------
static __device__ void bitonic_sort(unsigned long long* a, int npad, int tid, int nthreads) {
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += nthreads) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long x = a[i], y = a[p];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
}

// exclusive block scan of one flag per element, elements strided over the threads in chunks of 256
static __device__ int block_scan_flags(const unsigned long long* keys, int n, int tid, int* s_warp, int* s_carry, int* out_index) {
  // out_index[i] = number of heads before element i (heads: first element, or upper 32 bits differ from the previous one)
  const int lane = tid & 31, wid = tid >> 5;
  if (tid == 0) *s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + tid;
    const int head = (i < n && (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32))) ? 1 : 0;
    int incl = head;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = *s_carry, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (k < wid) before += s_warp[k]; total += s_warp[k]; }
    if (i < n) out_index[i] = head ? before + incl - 1 : -1;   // index of the head's group, -1 for the others
    __syncthreads();
    if (tid == 0) *s_carry += total;
    __syncthreads();
  }
  return *s_carry;
}

// dynamic shared memory: keys u64[npad] | vals f64[npad] | index i32[npad]
__global__ void __launch_bounds__(256) k_bow_assemble(const int* __restrict__ n_arr, int cap, int npad, const int* __restrict__ feat_word,
                                                      const int* __restrict__ feat_node, const double* __restrict__ feat_w, int weighting,
                                                      int norm_kind /* 0 none (divide by size for TF / TF_IDF), 1 L1, 2 L2 */,
                                                      int* __restrict__ bow_n, unsigned int* __restrict__ bow_word, double* __restrict__ bow_val,
                                                      int* __restrict__ fv_n, unsigned int* __restrict__ fv_node, int* __restrict__ fv_off,
                                                      unsigned int* __restrict__ fv_feat) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(s_raw);
  double* vals = reinterpret_cast<double*>(keys + npad);
  int* index = reinterpret_cast<int*>(vals + npad);
  __shared__ int s_warp[8];
  __shared__ int s_carry, s_valid;
  __shared__ double s_norm;
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int n = min(n_arr[frame], cap);
  const size_t fo = (size_t)frame * cap;
  if (tid == 0) s_valid = 0;
  __syncthreads();
  // ---- BowVector: sort (word, feature)
  int mine = 0;
  for (int i = tid; i < npad; i += 256) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const int w = feat_word[fo + i];
      if (w >= 0) { key = ((unsigned long long)(unsigned int)w << 32) | (unsigned int)i; ++mine; }
    }
    keys[i] = key;
  }
  atomicAdd(&s_valid, mine);
  __syncthreads();
  const int nv = s_valid;                         // features that are not stopped
  bitonic_sort(keys, npad, tid, 256);
  const int nb = block_scan_flags(keys, nv, tid, s_warp, &s_carry, index);
  for (int i = tid; i < nv; i += 256) {
    const int g = index[i];
    if (g < 0) continue;
    // addWeight (BowVector.cpp:34-46): the word's weight once per feature, in feature order; addIfNotExist keeps the first
    const unsigned int word = (unsigned int)(keys[i] >> 32);
    const double w = feat_w[fo + (unsigned int)keys[i]];
    double v = w;
    if (weighting == 0 || weighting == 1)
      for (int t = i + 1; t < nv && (unsigned int)(keys[t] >> 32) == word; ++t) v = __dadd_rn(v, w);
    vals[g] = v;
    bow_word[fo + g] = word;
  }
  __syncthreads();
  if (tid == 0) {
    double norm = 0.0;
    if (norm_kind == 1) for (int i = 0; i < nb; ++i) norm = __dadd_rn(norm, fabs(vals[i]));                 // :66-70
    else if (norm_kind == 2) { for (int i = 0; i < nb; ++i) norm = __dadd_rn(norm, __dmul_rn(vals[i], vals[i])); norm = sqrt(norm); }  // :72-76
    else norm = (weighting == 0 || weighting == 1) ? (double)nb : 0.0;   // TemplatedVocabulary.h:1163-1169
    s_norm = norm;
    bow_n[frame] = nb;
  }
  __syncthreads();
  {
    const double norm = s_norm;
    for (int i = tid; i < nb; i += 256) bow_val[fo + i] = norm > 0.0 ? __ddiv_rn(vals[i], norm) : vals[i];
  }
  __syncthreads();
  // ---- FeatureVector: sort (node, feature)
  for (int i = tid; i < npad; i += 256) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const int nd = feat_node[fo + i];
      if (feat_word[fo + i] >= 0) key = ((unsigned long long)(unsigned int)nd << 32) | (unsigned int)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic_sort(keys, npad, tid, 256);
  const int nn = block_scan_flags(keys, nv, tid, s_warp, &s_carry, index);
  int* off = fv_off + (size_t)frame * (cap + 1);
  for (int i = tid; i < nv; i += 256) {
    fv_feat[fo + i] = (unsigned int)keys[i];
    const int g = index[i];
    if (g >= 0) { fv_node[fo + g] = (unsigned int)(keys[i] >> 32); off[g] = i; }
  }
  if (tid == 0) { off[nn] = nv; fv_n[frame] = nn; }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int vocab_upload(orb_vocab* v, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
  // children in the order loadFromTextFile appends them (:1388): ascending node id per parent
  std::vector<int> start(n_nodes + 1, 0);
  for (int i = 1; i < n_nodes; ++i) {
    if (parent[i] < 0 || parent[i] >= i) return ORB_ERR_INVALID_ARG;     // the reference indexes m_nodes[pid] of a node it has not read yet
    start[parent[i] + 1]++;
  }
  int maxc = 0;
  for (int i = 0; i < n_nodes; ++i) { maxc = std::max(maxc, start[i + 1]); start[i + 1] += start[i]; }
  std::vector<int> cur(start.begin(), start.end() - 1), cid(std::max(n_nodes - 1, 1));
  std::vector<uint8_t> cdesc((size_t)std::max(n_nodes - 1, 1) * 32);
  std::vector<unsigned int> word(n_nodes, 0);
  int nwords = 0;
  for (int i = 1; i < n_nodes; ++i) {
    const int s = cur[parent[i]]++;
    cid[s] = i;
    std::memcpy(&cdesc[(size_t)s * 32], desc + (size_t)i * 32, 32);
    if (is_leaf[i]) word[i] = nwords++;                                   // :1408-1415
  }
  if (maxc > 255) return ORB_ERR_CAPACITY;
  v->n_nodes = n_nodes; v->n_words = nwords; v->max_children = maxc;
  if (cudaSetDevice(v->device) != cudaSuccess) return ORB_ERR_CUDA;
  const size_t ns = (size_t)std::max(n_nodes - 1, 1);
  if (cudaMalloc(&v->d_child_start, (size_t)(n_nodes + 1) * 4) != cudaSuccess || cudaMalloc(&v->d_child_id, ns * 4) != cudaSuccess ||
      cudaMalloc(&v->d_child_desc, ns * 32) != cudaSuccess || cudaMalloc(&v->d_word, (size_t)n_nodes * 4) != cudaSuccess ||
      cudaMalloc(&v->d_weight, (size_t)n_nodes * 8) != cudaSuccess)
    return ORB_ERR_CUDA;
  if (cudaMemcpy(v->d_child_start, start.data(), (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(v->d_child_id, cid.data(), ns * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(v->d_child_desc, cdesc.data(), ns * 32, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(v->d_word, word.data(), (size_t)n_nodes * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(v->d_weight, weight, (size_t)n_nodes * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    return ORB_ERR_CUDA;
  return ORB_OK;
}

static bool vocab_header_ok(int k, int L, int scoring, int weighting) {
  // the acceptance test of loadFromTextFile (:1359)
  return !(k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3);
}

// two-camera frame: features of the left camera followed by those of the right one (mDescriptors = vconcat(left, right))
__global__ void k_bow_concat(const int* __restrict__ nL_arr, int kL, const int* __restrict__ nR_arr, int kR, const int* __restrict__ wL,
                             const int* __restrict__ ndL, const double* __restrict__ fwL, const int* __restrict__ wR,
                             const int* __restrict__ ndR, const double* __restrict__ fwR, int* __restrict__ n2, int* __restrict__ w2,
                             int* __restrict__ nd2, double* __restrict__ fw2) {
  const int frame = blockIdx.y, cap2 = kL + kR;
  const int nL = min(nL_arr[frame], kL), nR = min(nR_arr[frame], kR);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) n2[frame] = nL + nR;
  if (i >= nL + nR) return;
  const size_t o = (size_t)frame * cap2 + i;
  if (i < nL) {
    const size_t s = (size_t)frame * kL + i;
    w2[o] = wL[s]; nd2[o] = ndL[s]; fw2[o] = fwL[s];
  } else {
    const size_t s = (size_t)frame * kR + (i - nL);
    w2[o] = wR[s]; nd2[o] = ndR[s]; fw2[o] = fwR[s];
  }
}

extern "C" {

int orb_vocab_destroy(orb_vocab* v) {
  if (!v) return ORB_OK;
  cudaSetDevice(v->device);
  cudaFree(v->d_child_start); cudaFree(v->d_child_id); cudaFree(v->d_child_desc); cudaFree(v->d_word); cudaFree(v->d_weight);
  delete v;
  return ORB_OK;
}

int orb_vocab_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                     const uint8_t* desc, const double* weight, orb_vocab** out) {
  if (!out || !parent || !is_leaf || !desc || !weight || n_nodes < 1) return ORB_ERR_INVALID_ARG;
  if (!vocab_header_ok(k, L, scoring, weighting)) return ORB_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ORB_ERR_CUDA;   // no CPU fallback
  orb_vocab* v = new orb_vocab;
  v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
  const int st = vocab_upload(v, n_nodes, parent, is_leaf, desc, weight);
  if (st != ORB_OK) { orb_vocab_destroy(v); return st; }
  *out = v;
  return ORB_OK;
}

// Text vocabulary (TemplatedVocabulary::loadFromTextFile, :1338-1426). Like the reference the file is parsed LINE BY LINE: a token
// never comes from the next line. A node line with fewer than 35 tokens is rejected (the reference would read indeterminate values
// there). Never throws: allocation failures come back as ORB_ERR_INVALID_ARG.
static int vocab_load_text_impl(int device, const char* path, orb_vocab** out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return ORB_ERR_INVALID_ARG;
  long size = -1;
  if (std::fseek(f, 0, SEEK_END) == 0) size = std::ftell(f);
  if (size < 0 || std::fseek(f, 0, SEEK_SET) != 0) { std::fclose(f); return ORB_ERR_INVALID_ARG; }   // not a seekable file
  std::vector<char> buf((size_t)size + 1);
  const size_t got = std::fread(buf.data(), 1, (size_t)size, f);
  std::fclose(f);
  buf[got] = 0;
  char* p = buf.data();
  char* const file_end = buf.data() + got;
  // next line as a NUL-terminated string; returns false at the end of the file
  auto next_line = [&](char*& line) -> bool {
    if (p >= file_end) return false;
    line = p;
    char* e = p;
    while (e < file_end && *e != '\n') ++e;
    p = e < file_end ? e + 1 : e;
    *e = 0;
    return true;
  };
  char* line;
  char* end;
  // header: k L scoring weighting (:1349-1357)
  if (!next_line(line)) return ORB_ERR_INVALID_ARG;
  char* q = line;
  const int k = (int)std::strtol(q, &end, 10); if (end == q) return ORB_ERR_INVALID_ARG; q = end;
  const int L = (int)std::strtol(q, &end, 10); if (end == q) return ORB_ERR_INVALID_ARG; q = end;
  const int scoring = (int)std::strtol(q, &end, 10); if (end == q) return ORB_ERR_INVALID_ARG; q = end;
  const int weighting = (int)std::strtol(q, &end, 10); if (end == q) return ORB_ERR_INVALID_ARG;
  if (!vocab_header_ok(k, L, scoring, weighting)) return ORB_ERR_INVALID_ARG;   // "This is not a correct text file!"
  std::vector<int32_t> parent(1, 0);
  std::vector<uint8_t> leaf(1, 0), desc(32, 0);
  std::vector<double> weight(1, 0.0);
  // one node per line: parent isLeaf d0 .. d31 weight (:1374-1421). Blank lines are skipped: the reference turns an empty last
  // line into one more child of the root with an uninitialised descriptor (its while(!f.eof()) loop), which cannot be reproduced.
  while (next_line(line)) {
    q = line;
    while (*q == '\r' || *q == ' ' || *q == '\t') ++q;
    if (!*q) continue;
    const long pid = std::strtol(q, &end, 10);
    if (end == q) return ORB_ERR_INVALID_ARG;
    q = end;
    const long isleaf = std::strtol(q, &end, 10);
    if (end == q) return ORB_ERR_INVALID_ARG;
    q = end;
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) { d[i] = (uint8_t)std::strtol(q, &end, 10); if (end == q) return ORB_ERR_INVALID_ARG; q = end; }
    const double w = std::strtod(q, &end);
    if (end == q) return ORB_ERR_INVALID_ARG;
    parent.push_back((int32_t)pid); leaf.push_back(isleaf > 0 ? 1 : 0); weight.push_back(w);
    desc.insert(desc.end(), d, d + 32);
  }
  return orb_vocab_create(device, k, L, scoring, weighting, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), out);
}

int orb_vocab_load_text(int device, const char* path, orb_vocab** out) {
  if (!path || !out) return ORB_ERR_INVALID_ARG;
  try {
    return vocab_load_text_impl(device, path, out);
  } catch (...) {   // std::bad_alloc / length_error of the buffers: the C ABI never throws
    return ORB_ERR_INVALID_ARG;
  }
}

int orb_vocab_info(const orb_vocab* v, int32_t* info6) {
  if (!v || !info6) return ORB_ERR_INVALID_ARG;
  info6[0] = v->k; info6[1] = v->L; info6[2] = v->scoring; info6[3] = v->weighting; info6[4] = v->n_nodes; info6[5] = v->n_words;
  return ORB_OK;
}

int orb_compute_bow(orb_handle* h, const orb_vocab* v, int levelsup, const orb_bow_out* out, int flags) {
  if (!h || !v) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  if (v->device != h->device) return orb_set_error(h, ORB_ERR_INVALID_ARG, "vocabulary lives on another device");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  int npad = 32;
  while (npad < kcap) npad <<= 1;
  const size_t smem = (size_t)npad * 20;
  if (smem > 200 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many keypoints per frame for the bag-of-words assembly");
  const size_t nk = (size_t)batch * kcap;
  if ((st = orb_ensure(h, h->d_bow_fword, nk * 4)) || (st = orb_ensure(h, h->d_bow_fnode, nk * 4)) || (st = orb_ensure(h, h->d_bow_fw, nk * 8)) ||
      (st = orb_ensure(h, h->d_bow_n, (size_t)batch * 8)) || (st = orb_ensure(h, h->d_bow_word, nk * 4)) || (st = orb_ensure(h, h->d_bow_val, nk * 8)) ||
      (st = orb_ensure(h, h->d_fv_node, nk * 4)) || (st = orb_ensure(h, h->d_fv_off, (size_t)batch * (kcap + 1) * 4)) ||
      (st = orb_ensure(h, h->d_fv_feat, nk * 4)))
    return st;
  int* d_bow_n = h->d_bow_n.as<int>();
  int* d_fv_n = d_bow_n + batch;
  if (v->n_words == 0) {                                         // empty() (:1133): both vectors stay empty
    ORB_CUDA_CHECK(h, cudaMemsetAsync(d_bow_n, 0, (size_t)batch * 8, h->stream));
    ORB_CUDA_CHECK(h, cudaMemsetAsync(h->d_fv_off.p, 0, (size_t)batch * (kcap + 1) * 4, h->stream));
    ORB_CUDA_CHECK(h, cudaMemsetAsync(h->d_bow_fword.p, 0xff, nk * 4, h->stream));
    ORB_CUDA_CHECK(h, cudaMemsetAsync(h->d_bow_fnode.p, 0xff, nk * 4, h->stream));
  } else {
    const int nid_level = v->L - levelsup;
    int G = 4;
    while (G < 32 && G < v->max_children) G <<= 1;
    const dim3 grid((unsigned)(((size_t)kcap * G + 255) / 256), batch);
#define BOW_DESCEND(GG)                                                                                                                  \
  k_bow_descend<GG><<<grid, 256, 0, h->stream>>>(h->d_desc.as<uint8_t>(), h->d_n.as<int>(), kcap, v->d_child_start, v->d_child_id,         \
                                                 v->d_child_desc, v->d_word, v->d_weight, nid_level, h->d_bow_fword.as<int>(),             \
                                                 h->d_bow_fnode.as<int>(), h->d_bow_fw.as<double>())
    if (G == 4) BOW_DESCEND(4); else if (G == 8) BOW_DESCEND(8); else if (G == 16) BOW_DESCEND(16); else BOW_DESCEND(32);
#undef BOW_DESCEND
    h->launches++;
    // every scoring but DOT_PRODUCT normalises: L2_NORM with L2, the others with L1 (ScoringObject.h:74-89)
    const int norm_kind = v->scoring == 5 ? 0 : (v->scoring == 1 ? 2 : 1);
    { const int st_a = orb_raise_dyn_smem(h, (const void*)k_bow_assemble, smem); if (st_a) return st_a; }
    k_bow_assemble<<<batch, 256, smem, h->stream>>>(h->d_n.as<int>(), kcap, npad, h->d_bow_fword.as<int>(), h->d_bow_fnode.as<int>(),
                                                    h->d_bow_fw.as<double>(), v->weighting, norm_kind, d_bow_n, h->d_bow_word.as<unsigned int>(),
                                                    h->d_bow_val.as<double>(), d_fv_n, h->d_fv_node.as<unsigned int>(), h->d_fv_off.as<int>(),
                                                    h->d_fv_feat.as<unsigned int>());
    h->launches++;
    ORB_CUDA_CHECK(h, cudaGetLastError());
  }
  h->have_bow = true;
  if (out && !(flags & ORB_NO_OUTPUT)) {
#define BOW_COPY(dst, src, bytes) \
  if (dst) ORB_CUDA_CHECK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream))
    BOW_COPY(out->bow_n, d_bow_n, (size_t)batch * 4);
    BOW_COPY(out->bow_word, h->d_bow_word.p, nk * 4);
    BOW_COPY(out->bow_val, h->d_bow_val.p, nk * 8);
    BOW_COPY(out->fv_n, d_fv_n, (size_t)batch * 4);
    BOW_COPY(out->fv_node, h->d_fv_node.p, nk * 4);
    BOW_COPY(out->fv_off, h->d_fv_off.p, (size_t)batch * (kcap + 1) * 4);
    BOW_COPY(out->fv_feat, h->d_fv_feat.p, nk * 4);
    BOW_COPY(out->feat_word, h->d_bow_fword.p, nk * 4);
    BOW_COPY(out->feat_node, h->d_bow_fnode.p, nk * 4);
#undef BOW_COPY
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_compute_bow_stereo(orb_handle* hL, orb_handle* hR, const orb_vocab* v, int levelsup, const orb_bow_out* out, int flags) {
  if (!hL || !hR || !v || hL == hR) return ORB_ERR_INVALID_ARG;
  if (!hL->have_batch || !hR->have_batch) return orb_set_error(hL, ORB_ERR_STATE, "no extraction has run on both handles");
  if (v->device != hL->device || hL->device != hR->device || hL->cur_batch != hR->cur_batch)
    return orb_set_error(hL, ORB_ERR_INVALID_ARG, "vocabulary and both handles must share the device, and the batches must be equal");
  int st;
  if ((st = orb_use_device(hL))) return st;
  const int batch = hL->cur_batch, kL = hL->g.kcap, kR = hR->g.kcap, cap2 = kL + kR;
  int npad = 32;
  while (npad < cap2) npad <<= 1;
  const size_t smem = (size_t)npad * 20;
  if (smem > 200 * 1024) return orb_set_error(hL, ORB_ERR_CAPACITY, "too many keypoints per frame for the bag-of-words assembly");
  const size_t B = (size_t)batch, nl = B * kL, nr = B * kR, n2 = B * cap2;
  // regions: per-camera word / node / weight (0-2 left, 3-5 right), combined (6-8), n2 (9), bow_n + fv_n (10), bow_word (11),
  //          bow_val (12), fv_node (13), fv_off (14), fv_feat (15)
  const size_t bytes[16] = {nl * 4, nl * 4, nl * 8, nr * 4, nr * 4, nr * 8, n2 * 4, n2 * 4, n2 * 8, B * 4, B * 8, n2 * 4, n2 * 8, n2 * 4,
                            B * (cap2 + 1) * 4, n2 * 4};
  size_t off = 0, o[16];
  for (int i = 0; i < 16; ++i) { o[i] = off; off += (bytes[i] + 255) & ~(size_t)255; }
  if ((st = orb_ensure(hL, hL->d_bow2, off))) return st;
  uint8_t** r = hL->bow2_r;
  for (int i = 0; i < 16; ++i) r[i] = hL->d_bow2.as<uint8_t>() + o[i];
  hL->bow2_cap = cap2;
  int* d_n2 = (int*)r[9];
  int* d_bow_n = (int*)r[10];
  int* d_fv_n = d_bow_n + batch;
  if ((st = orb_peer_read_begin(hL, hR))) return st;
  if (v->n_words == 0) {
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(d_bow_n, 0, B * 8, hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(r[14], 0, bytes[14], hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(r[6], 0xff, bytes[6], hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(r[7], 0xff, bytes[7], hL->stream));
  } else {
    const int nid_level = v->L - levelsup;
    int G = 4;
    while (G < 32 && G < v->max_children) G <<= 1;
    for (int side = 0; side < 2; ++side) {
      orb_handle* h = side ? hR : hL;
      const int kcap = side ? kR : kL;
      const dim3 grid((unsigned)(((size_t)kcap * G + 255) / 256), batch);
      int* fw_ = (int*)r[3 * side]; int* fn_ = (int*)r[3 * side + 1]; double* fd_ = (double*)r[3 * side + 2];
#define BOW_DESCEND(GG)                                                                                                            \
  k_bow_descend<GG><<<grid, 256, 0, hL->stream>>>(h->d_desc.as<uint8_t>(), h->d_n.as<int>(), kcap, v->d_child_start, v->d_child_id, \
                                                  v->d_child_desc, v->d_word, v->d_weight, nid_level, fw_, fn_, fd_)
      if (G == 4) BOW_DESCEND(4); else if (G == 8) BOW_DESCEND(8); else if (G == 16) BOW_DESCEND(16); else BOW_DESCEND(32);
#undef BOW_DESCEND
    }
    k_bow_concat<<<dim3((cap2 + 255) / 256, batch), 256, 0, hL->stream>>>(hL->d_n.as<int>(), kL, hR->d_n.as<int>(), kR, (int*)r[0], (int*)r[1],
                                                                         (double*)r[2], (int*)r[3], (int*)r[4], (double*)r[5], d_n2, (int*)r[6],
                                                                         (int*)r[7], (double*)r[8]);
    const int norm_kind = v->scoring == 5 ? 0 : (v->scoring == 1 ? 2 : 1);
    if ((st = orb_raise_dyn_smem(hL, (const void*)k_bow_assemble, smem))) return st;
    k_bow_assemble<<<batch, 256, smem, hL->stream>>>(d_n2, cap2, npad, (int*)r[6], (int*)r[7], (double*)r[8], v->weighting, norm_kind, d_bow_n,
                                                     (unsigned int*)r[11], (double*)r[12], d_fv_n, (unsigned int*)r[13], (int*)r[14],
                                                     (unsigned int*)r[15]);
    hL->launches += 4;
    ORB_CUDA_CHECK(hL, cudaGetLastError());
  }
  if ((st = orb_peer_read_end(hL, hR))) return st;
  hL->have_bow2 = true;
  if (out && !(flags & ORB_NO_OUTPUT)) {
#define BOW_COPY(dst, src, nbytes) \
  if (dst) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, hL->stream))
    BOW_COPY(out->bow_n, d_bow_n, B * 4);
    BOW_COPY(out->bow_word, r[11], bytes[11]);
    BOW_COPY(out->bow_val, r[12], bytes[12]);
    BOW_COPY(out->fv_n, d_fv_n, B * 4);
    BOW_COPY(out->fv_node, r[13], bytes[13]);
    BOW_COPY(out->fv_off, r[14], bytes[14]);
    BOW_COPY(out->fv_feat, r[15], bytes[15]);
    BOW_COPY(out->feat_word, r[6], bytes[6]);
    BOW_COPY(out->feat_node, r[7], bytes[7]);
#undef BOW_COPY
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

}  // extern "C"
