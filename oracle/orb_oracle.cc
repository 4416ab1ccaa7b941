// TEST INFRASTRUCTURE ONLY (oracle). CPU restatement of the reference's ORB front-end, written in
// the data-parallel formulation the CUDA kernels use (global FAST score map + per-cell NMS,
// array-based quad-tree with an explicit libstdc++ introsort emulation, order-free stereo predicate),
// so that the formulation itself is checked against the unmodified reference (oracle/_ref) on the
// CPU before any kernel is trusted. Each function cites the reference lines it follows.
//
// Parity status: pinned. tests/test_oracle_*.py check (i) every OpenCV primitive below against the
// cv2 4.13.0 wheel, (ii) this restatement against oracle/_ref (the reference's own
// src/ORBextractor.cc, src/Frame.cc:889-1047, src/ORBmatcher.cc:1880-1894 compiled unmodified) on
// seeded frames of all BASELINE.json configs, (iii) both against the committed tests/golden vectors.
#include "orb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "opencv2/cvshim.hpp"
#include "sincosf_restate.h"

namespace {

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;  // src/ORBextractor.cc:71-73
const int kBorder = kEdge - 3;                        // FAST working border, :747

const int8_t kPattern[1024] = {
#include "orb_pattern_31.inc"
};

struct Cand { int x, y, score; };

// ---------------------------------------------------------------------------------------------
// libstdc++ std::sort emulation (bits/stl_algo.h __sort: introsort loop, threshold 16, median of
// three to first, unguarded partition, heap-sort fallback, final insertion sort). Elements are
// (key, payload); only key takes part in comparisons, exactly like compareNodes
// (src/ORBextractor.cc:525-538) with key = (nKeys << 16) | UL.x.
// ---------------------------------------------------------------------------------------------
struct SortEl { uint32_t key, val; };
inline bool lessEl(const SortEl& a, const SortEl& b) { return a.key < b.key; }

void adjust_heap(SortEl* a, int hole, int len, SortEl value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (lessEl(a[child], a[child - 1])) child--;
    a[hole] = a[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[hole] = a[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && lessEl(a[parent], value)) {
    a[hole] = a[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[hole] = value;
}

void heap_sort(SortEl* a, int n) {  // __partial_sort(first, last, last)
  if (n >= 2) {
    int parent = (n - 2) / 2;
    while (true) {
      SortEl v = a[parent];
      adjust_heap(a, parent, n, v);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = n;
  while (last > 1) {
    --last;
    SortEl v = a[last];
    a[last] = a[0];
    adjust_heap(a, 0, last, v);
  }
}

void introsort_loop(SortEl* a, int first, int last, int depth) {
  while (last - first > 16) {
    if (depth == 0) { heap_sort(a + first, last - first); return; }
    --depth;
    int mid = first + (last - first) / 2;
    {  // __move_median_to_first(first, first+1, mid, last-1)
      int ia = first + 1, ib = mid, ic = last - 1, pick;
      if (lessEl(a[ia], a[ib])) {
        if (lessEl(a[ib], a[ic])) pick = ib;
        else if (lessEl(a[ia], a[ic])) pick = ic;
        else pick = ia;
      } else if (lessEl(a[ia], a[ic])) pick = ia;
      else if (lessEl(a[ib], a[ic])) pick = ic;
      else pick = ib;
      std::swap(a[first], a[pick]);
    }
    int lo = first + 1, hi = last;  // __unguarded_partition(first+1, last, pivot=first)
    while (true) {
      while (lessEl(a[lo], a[first])) ++lo;
      --hi;
      while (lessEl(a[first], a[hi])) --hi;
      if (!(lo < hi)) break;
      std::swap(a[lo], a[hi]);
      ++lo;
    }
    introsort_loop(a, lo, last, depth);
    last = lo;
  }
}

void std_sort_emulated(SortEl* a, int n) {
  if (n <= 0) return;
  int lg = 31 - __builtin_clz((unsigned)n);
  introsort_loop(a, 0, n, 2 * lg);
  // __final_insertion_sort
  int guarded = n > 16 ? 16 : n;
  for (int i = 1; i < guarded; ++i) {
    SortEl v = a[i];
    if (lessEl(v, a[0])) {
      for (int j = i; j > 0; --j) a[j] = a[j - 1];
      a[0] = v;
    } else {
      int j = i;
      while (lessEl(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
      a[j] = v;
    }
  }
  for (int i = guarded; i < n; ++i) {  // unguarded linear insert
    SortEl v = a[i];
    int j = i;
    while (lessEl(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
    a[j] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree on arrays (src/ORBextractor.cc:540-738; DivideNode :475-523).
// Keys live in one array; every node owns a contiguous segment of it; DivideNode is a stable
// 4-way partition of that segment. The std::list is a doubly linked list of node slots.
// ---------------------------------------------------------------------------------------------
struct QNode {
  int begin, count;
  int ulx, uly, urx, bry;
  int prev, next;
  bool noMore;
};

struct QuadTree {
  std::vector<Cand> keys, tmp;
  std::vector<QNode> nodes;
  int head = -1, tail = -1, size = 0;

  int newNode() { nodes.push_back(QNode()); return (int)nodes.size() - 1; }
  void pushFront(int n) {
    nodes[n].prev = -1; nodes[n].next = head;
    if (head >= 0) nodes[head].prev = n; else tail = n;
    head = n; ++size;
  }
  void pushBack(int n) {
    nodes[n].next = -1; nodes[n].prev = tail;
    if (tail >= 0) nodes[tail].next = n; else head = n;
    tail = n; ++size;
  }
  int erase(int n) {  // returns the following node
    int p = nodes[n].prev, q = nodes[n].next;
    if (p >= 0) nodes[p].next = q; else head = q;
    if (q >= 0) nodes[q].prev = p; else tail = p;
    --size;
    return q;
  }
  // DivideNode: children in order n1 (UL), n2 (UR), n3 (BL), n4 (BR); returns their slots (or -1 if empty)
  void divide(int n, int child[4]) {
    const QNode P = nodes[n];
    const int halfX = (int)std::ceil((float)(P.urx - P.ulx) / 2);
    const int halfY = (int)std::ceil((float)(P.bry - P.uly) / 2);
    const int midX = P.ulx + halfX, midY = P.uly + halfY;
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.count; ++i) {
      const Cand& k = keys[P.begin + i];
      int q = ((float)k.x < (float)midX) ? (((float)k.y < (float)midY) ? 0 : 2) : (((float)k.y < (float)midY) ? 1 : 3);
      cnt[q]++;
    }
    int off[4] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
    int w[4] = {off[0], off[1], off[2], off[3]};
    tmp.resize(keys.size());
    for (int i = 0; i < P.count; ++i) {
      const Cand& k = keys[P.begin + i];
      int q = (k.x < midX) ? ((k.y < midY) ? 0 : 2) : ((k.y < midY) ? 1 : 3);
      tmp[P.begin + w[q]++] = k;
    }
    for (int i = 0; i < P.count; ++i) keys[P.begin + i] = tmp[P.begin + i];
    const int rx[4][2] = {{P.ulx, midX}, {midX, P.urx}, {P.ulx, midX}, {midX, P.urx}};
    const int ry[4][2] = {{P.uly, midY}, {P.uly, midY}, {midY, P.bry}, {midY, P.bry}};
    for (int q = 0; q < 4; ++q) {
      if (cnt[q] == 0) { child[q] = -1; continue; }
      int c = newNode();
      QNode& C = nodes[c];
      C.begin = P.begin + off[q]; C.count = cnt[q];
      C.ulx = rx[q][0]; C.urx = rx[q][1]; C.uly = ry[q][0]; C.bry = ry[q][1];
      C.noMore = (cnt[q] == 1);
      C.prev = C.next = -1;
      child[q] = c;
    }
  }
};

std::vector<Cand> distribute_octree(const std::vector<Cand>& in, int w, int h, int N) {
  std::vector<Cand> out;
  if (in.empty()) return out;
  QuadTree T;
  const int nIni = (int)std::round((float)w / (float)h);  // :545 (std::round, half away from zero)
  if (nIni < 1) return out;                                // the reference divides by zero here; callers reject such sizes
  const float hX = (float)w / nIni;                        // :547
  // root membership by truncating pt.x / hX (:567-570), order preserved inside each root
  std::vector<int> rootOf(in.size());
  std::vector<int> rcount(nIni, 0);
  for (size_t i = 0; i < in.size(); ++i) {
    int r = (int)((float)in[i].x / hX);
    if (r >= nIni) r = nIni - 1;  // cannot happen for x < w; keeps the array access defined
    rootOf[i] = r;
    rcount[r]++;
  }
  std::vector<int> rbegin(nIni, 0);
  for (int r = 1; r < nIni; ++r) rbegin[r] = rbegin[r - 1] + rcount[r - 1];
  T.keys.resize(in.size());
  {
    std::vector<int> wpos = rbegin;
    for (size_t i = 0; i < in.size(); ++i) T.keys[wpos[rootOf[i]]++] = in[i];
  }
  for (int r = 0; r < nIni; ++r) {  // :554-565, then the empty / single-key pass of :574-582
    if (rcount[r] == 0) continue;
    int n = T.newNode();
    QNode& R = T.nodes[n];
    R.begin = rbegin[r]; R.count = rcount[r];
    R.ulx = (int)(hX * (float)r); R.urx = (int)(hX * (float)(r + 1));
    R.uly = 0; R.bry = h;
    R.noMore = (rcount[r] == 1);
    T.pushBack(n);
  }

  std::vector<SortEl> rec, prevRec;  // vSizeAndPointerToNode
  bool finish = false;
  while (!finish) {  // :591
    int prevSize = T.size;
    int nToExpand = 0;
    rec.clear();
    int it = T.head;
    while (it >= 0) {
      if (T.nodes[it].noMore) { it = T.nodes[it].next; continue; }
      int ch[4];
      T.divide(it, ch);
      for (int q = 0; q < 4; ++q) {
        if (ch[q] < 0) continue;
        T.pushFront(ch[q]);
        if (T.nodes[ch[q]].count > 1) {
          nToExpand++;
          rec.push_back(SortEl{((uint32_t)T.nodes[ch[q]].count << 16) | (uint32_t)T.nodes[ch[q]].ulx, (uint32_t)ch[q]});
        }
      }
      it = T.erase(it);
    }
    if (T.size >= N || T.size == prevSize) {
      finish = true;
    } else if (T.size + nToExpand * 3 > N) {  // :659
      while (!finish) {
        prevSize = T.size;
        prevRec = rec;
        rec.clear();
        std_sort_emulated(prevRec.data(), (int)prevRec.size());
        for (int j = (int)prevRec.size() - 1; j >= 0; --j) {
          int n = (int)prevRec[j].val;
          int ch[4];
          T.divide(n, ch);
          for (int q = 0; q < 4; ++q) {
            if (ch[q] < 0) continue;
            T.pushFront(ch[q]);
            if (T.nodes[ch[q]].count > 1)
              rec.push_back(SortEl{((uint32_t)T.nodes[ch[q]].count << 16) | (uint32_t)T.nodes[ch[q]].ulx, (uint32_t)ch[q]});
          }
          T.erase(n);
          if (T.size >= N) break;
        }
        if (T.size >= N || T.size == prevSize) finish = true;
      }
    }
  }
  // :718-735 best response per leaf, first maximum wins, list order
  for (int it = T.head; it >= 0; it = T.nodes[it].next) {
    const QNode& nd = T.nodes[it];
    int best = nd.begin;
    for (int k = 1; k < nd.count; ++k)
      if (T.keys[nd.begin + k].score > T.keys[best].score) best = nd.begin + k;
    out.push_back(T.keys[best]);
  }
  return out;
}

// DistributeOctTree in PASS form - the formulation a block-parallel kernel can take. The reference's loop
// (src/ORBextractor.cc:591-712) visits, in one pass, exactly the nodes that were in the list when the pass started: children
// are push_front()ed, i.e. land before the iterator. Consequences used here:
//   * the divisions of one pass are independent of each other (each works on its own key segment);
//   * the list after the pass = [children of the LAST divided node as n4 n3 n2 n1, ..., children of the FIRST divided node]
//     followed by the undivided nodes in their old order;
//   * vSizeAndPointerToNode = the children with more than one key, in visiting order (n1 .. n4 per node);
//   * the final phase divides the nodes of the previous record list from the back of the emulated std::sort order and stops as
//     soon as the list holds N nodes: the number of divisions is the first prefix of (children - 1) that reaches N.
// Every "for each node of the pass" loop below has no loop-carried state except the placement offsets (prefix sums).
// tests/test_oracle_vs_ref.py checks it against distribute_octree above and the reference's own code.
std::vector<Cand> distribute_octree_passes(const std::vector<Cand>& in, int w, int h, int N) {
  std::vector<Cand> out;
  if (in.empty()) return out;
  QuadTree T;   // only its key array, node slots and divide() are used; the list is the vector `order`
  const int nIni = (int)std::round((float)w / (float)h);
  if (nIni < 1) return out;
  const float hX = (float)w / nIni;
  std::vector<int> rootOf(in.size()), rcount(nIni, 0);
  for (size_t i = 0; i < in.size(); ++i) {
    int r = (int)((float)in[i].x / hX);
    if (r >= nIni) r = nIni - 1;
    rootOf[i] = r;
    rcount[r]++;
  }
  std::vector<int> rbegin(nIni, 0);
  for (int r = 1; r < nIni; ++r) rbegin[r] = rbegin[r - 1] + rcount[r - 1];
  T.keys.resize(in.size());
  {
    std::vector<int> wpos = rbegin;
    for (size_t i = 0; i < in.size(); ++i) T.keys[wpos[rootOf[i]]++] = in[i];
  }
  std::vector<int> order;   // the std::list, front first
  for (int r = 0; r < nIni; ++r) {
    if (rcount[r] == 0) continue;
    int n = T.newNode();
    QNode& R = T.nodes[n];
    R.begin = rbegin[r]; R.count = rcount[r];
    R.ulx = (int)(hX * (float)r); R.urx = (int)(hX * (float)(r + 1));
    R.uly = 0; R.bry = h;
    R.noMore = (rcount[r] == 1);
    order.push_back(n);
  }
  struct Div { int node; int ch[4]; int nch; };
  auto rec_of = [&](int c) { return SortEl{((uint32_t)T.nodes[c].count << 16) | (uint32_t)T.nodes[c].ulx, (uint32_t)c}; };
  std::vector<SortEl> rec, prevRec;
  bool finish = false;
  while (!finish) {
    const int prevSize = (int)order.size();
    // ---- one pass: divide every divisible node (independent), then place
    std::vector<Div> divs;
    for (int n : order) if (!T.nodes[n].noMore) divs.push_back(Div{n, {-1, -1, -1, -1}, 0});
    for (Div& d : divs) {                       // parallel over nodes
      T.divide(d.node, d.ch);
      for (int q = 0; q < 4; ++q) d.nch += d.ch[q] >= 0;
    }
    std::vector<int> next;
    next.reserve(order.size() + 3 * divs.size());
    for (int i = (int)divs.size() - 1; i >= 0; --i)
      for (int q = 3; q >= 0; --q) if (divs[i].ch[q] >= 0) next.push_back(divs[i].ch[q]);
    for (int n : order) if (T.nodes[n].noMore) next.push_back(n);
    rec.clear();
    int nToExpand = 0;
    for (const Div& d : divs)
      for (int q = 0; q < 4; ++q)
        if (d.ch[q] >= 0 && T.nodes[d.ch[q]].count > 1) { nToExpand++; rec.push_back(rec_of(d.ch[q])); }
    order.swap(next);
    const int size = (int)order.size();
    if (size >= N || size == prevSize) {
      finish = true;
    } else if (size + nToExpand * 3 > N) {
      while (!finish) {
        const int prev = (int)order.size();
        prevRec = rec;
        rec.clear();
        std_sort_emulated(prevRec.data(), (int)prevRec.size());
        // divisions from the back of the sorted order; all of them are independent, the stop index is a prefix sum
        const int m = (int)prevRec.size();
        std::vector<Div> dv(m);
        for (int t = 0; t < m; ++t) dv[t] = Div{(int)prevRec[m - 1 - t].val, {-1, -1, -1, -1}, 0};   // t = 0 is divided first
        // step 1 (parallel over the m nodes): non-empty quadrants of every node - a histogram, no key moves
        std::vector<int> nch(m, 0);
        for (int t = 0; t < m; ++t) {
          const QNode& P = T.nodes[dv[t].node];
          const int midX = P.ulx + (int)std::ceil((float)(P.urx - P.ulx) / 2), midY = P.uly + (int)std::ceil((float)(P.bry - P.uly) / 2);
          bool has[4] = {false, false, false, false};
          for (int i = 0; i < P.count; ++i) {
            const Cand& k = T.keys[P.begin + i];
            has[(k.x < midX) ? ((k.y < midY) ? 0 : 2) : ((k.y < midY) ? 1 : 3)] = true;
          }
          nch[t] = has[0] + has[1] + has[2] + has[3];
        }
        // step 2 (scan): the loop stops after the first division that brings the list to N nodes
        int used = m, sz = prev;
        for (int t = 0; t < m; ++t) {
          sz += nch[t] - 1;
          if (sz >= N) { used = t + 1; break; }
        }
        // step 3 (parallel over the first `used` nodes): the partitions themselves
        for (int t = 0; t < used; ++t) {
          T.divide(dv[t].node, dv[t].ch);
          for (int q = 0; q < 4; ++q) dv[t].nch += dv[t].ch[q] >= 0;
        }
        std::vector<char> erased(T.nodes.size(), 0);
        for (int t = 0; t < used; ++t) erased[dv[t].node] = 1;
        std::vector<int> nx;
        nx.reserve(sz);
        for (int t = used - 1; t >= 0; --t)
          for (int q = 3; q >= 0; --q) if (dv[t].ch[q] >= 0) nx.push_back(dv[t].ch[q]);
        for (int n : order) if (!erased[n]) nx.push_back(n);
        for (int t = 0; t < used; ++t)
          for (int q = 0; q < 4; ++q)
            if (dv[t].ch[q] >= 0 && T.nodes[dv[t].ch[q]].count > 1) rec.push_back(rec_of(dv[t].ch[q]));
        order.swap(nx);
        if ((int)order.size() >= N || (int)order.size() == prev) finish = true;
      }
    }
  }
  for (int n : order) {
    const QNode& nd = T.nodes[n];
    int best = nd.begin;
    for (int k = 1; k < nd.count; ++k)
      if (T.keys[nd.begin + k].score > T.keys[best].score) best = nd.begin + k;
    out.push_back(T.keys[best]);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
class Oracle {
 public:
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;  // the reference stores the float argument in a double member (include/ORBextractor.h:92)
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nFeat, umax;
  // stage outputs of the last extract()
  std::vector<int> lw, lh;
  std::vector<std::vector<uint8_t>> level, blurred;
  std::vector<std::vector<Cand>> cands;
  std::vector<std::vector<cv::KeyPoint>> levelKps;

  Oracle(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
    // src/ORBextractor.cc:413-427
    scale.resize(nl); invScale.resize(nl); sigma2.resize(nl); invSigma2.resize(nl);
    scale[0] = 1.0f; sigma2[0] = 1.0f;
    for (int i = 1; i < nl; ++i) {
      scale[i] = (float)(scale[i - 1] * scaleFactor);
      sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < nl; ++i) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
    // :431-443
    nFeat.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
      nFeat[l] = cvRound(nDesired);
      sum += nFeat[l];
      nDesired *= factor;
    }
    nFeat[nl - 1] = std::max(nfeatures - sum, 0);
    // :451-463
    umax.resize(kHalfPatch + 1);
    int v, v0, vmax = cvFloor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = cvCeil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) umax[v] = cvRound(std::sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }

  // ComputePyramid (:1088-1112) without the 19-px reflected margin (never read downstream)
  void pyramid(const uint8_t* img, int w, int h, int stride) {
    lw.assign(nlevels, 0); lh.assign(nlevels, 0);
    level.assign(nlevels, {});
    for (int l = 0; l < nlevels; ++l) {
      lw[l] = cvRound((float)w * invScale[l]);
      lh[l] = cvRound((float)h * invScale[l]);
      level[l].resize((size_t)lw[l] * lh[l]);
      if (l == 0) {
        for (int y = 0; y < h; ++y) std::memcpy(&level[0][(size_t)y * w], img + (size_t)y * stride, w);
      } else {
        shim_resize(level[l - 1].data(), lw[l - 1], lh[l - 1], lw[l - 1], level[l].data(), lw[l], lh[l]);
      }
    }
  }

  // FAST corner score (OpenCV cornerScore<16>) at an arbitrary pixel
  static int score_at(const uint8_t* p, int stride) { return cv::shim_detail::fast_score(p, (size_t)stride); }

  // Cell/FAST part of ComputeKeyPointsOctTree (:744-820) in the global formulation:
  // one score map per level; per cell: local maxima inside the cell interior, threshold fallback.
  bool fast_level(int l, std::vector<Cand>& out) {
    out.clear();
    const int W = lw[l], H = lh[l];
    const uint8_t* im = level[l].data();
    const int minBX = kBorder, minBY = kBorder, maxBX = W - kEdge + 3, maxBY = H - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / 35.f), nRows = (int)(height / 35.f);
    if (nCols < 1 || nRows < 1) return false;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    // score map, clamped below at 0 (non-corners never matter: they lose against any score >= minTh)
    std::vector<uint8_t> S((size_t)W * H, 0);
    for (int y = minBY + 3; y < maxBY - 3; ++y)
      for (int x = minBX + 3; x < maxBX - 3; ++x) {
        int s = score_at(im + (size_t)y * W + x, W);
        S[(size_t)y * W + x] = (uint8_t)(s < minTh ? 0 : s);
      }
    for (int i = 0; i < nRows; ++i) {
      const int iniY = minBY + i * hCell;
      int maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = maxBY;
      for (int j = 0; j < nCols; ++j) {
        const int iniX = minBX + j * wCell;
        int maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = maxBX;
        const int x0 = iniX + 3, x1 = maxX - 3, y0 = iniY + 3, y1 = maxY - 3;  // cell interior
        std::vector<Cand> cell;
        bool anyIni = false;
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) {
            int s = S[(size_t)y * W + x];
            if (s == 0) continue;
            bool lm = true;
            for (int dy = -1; dy <= 1 && lm; ++dy)
              for (int dx = -1; dx <= 1; ++dx) {
                if (!dx && !dy) continue;
                int xx = x + dx, yy = y + dy;
                int q = (xx >= x0 && xx < x1 && yy >= y0 && yy < y1) ? S[(size_t)yy * W + xx] : 0;
                if (!(s > q)) { lm = false; break; }
              }
            if (!lm) continue;
            if (s >= iniTh) anyIni = true;
            cell.push_back(Cand{x - minBX, y - minBY, s});
          }
        const int th = anyIni ? iniTh : minTh;
        for (const Cand& c : cell)
          if (c.score >= th) out.push_back(c);
      }
    }
    return true;
  }

  // IC_Angle (:75-99)
  float ic_angle(int l, int x, int y) const {
    const int W = lw[l];
    const uint8_t* c = level[l].data() + (size_t)y * W + x;
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
      int vsum = 0, d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int vp = c[u + v * W], vm = c[u - v * W];
        vsum += (vp - vm);
        m10 += u * (vp + vm);
      }
      m01 += v * vsum;
    }
    return cv::fastAtan2((float)m01, (float)m10);
  }

  // computeOrbDescriptor (:102-145) on the blurred level
  void descriptor(int l, const cv::KeyPoint& kp, uint8_t* desc) const {
    const float factorPI = (float)(CV_PI / 180.f);
    float angle = (float)kp.angle * factorPI;
    float a = cosf(angle), b = sinf(angle);  // glibc float routines, as the reference resolves them
    const int W = lw[l];
    const uint8_t* c = blurred[l].data() + (size_t)cvRound(kp.pt.y) * W + cvRound(kp.pt.x);
    for (int i = 0; i < 32; ++i) {
      int val = 0;
      for (int k = 0; k < 8; ++k) {
        const int8_t* p = &kPattern[(i * 16 + 2 * k) * 2];
        int t0 = c[cvRound(p[0] * b + p[1] * a) * W + cvRound(p[0] * a - p[1] * b)];
        int t1 = c[cvRound(p[2] * b + p[3] * a) * W + cvRound(p[2] * a - p[3] * b)];
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // operator() (:1006-1086)
  int extract(const uint8_t* img, int w, int h, int stride, int lap0, int lap1, std::vector<cv::KeyPoint>& kps,
              std::vector<uint8_t>& desc) {
    kps.clear(); desc.clear();
    if (!img || w <= 0 || h <= 0) return -1;
    pyramid(img, w, h, stride);
    cands.assign(nlevels, {});
    levelKps.assign(nlevels, {});
    blurred.assign(nlevels, {});
    for (int l = 0; l < nlevels; ++l) {
      if (!fast_level(l, cands[l])) return -3;
      const int rw = lw[l] - 2 * kBorder, rh = lh[l] - 2 * kBorder;  // maxBorder - minBorder
      if ((int)std::round((float)rw / (float)rh) < 1) return -3;
      std::vector<Cand> sel = distribute_octree(cands[l], rw, rh, nFeat[l]);
      const int scaledPatch = (int)(kPatch * scale[l]);  // :826
      for (const Cand& c : sel) {
        cv::KeyPoint kp((float)(c.x + kBorder), (float)(c.y + kBorder), (float)scaledPatch, -1.f, (float)c.score, l, -1);
        levelKps[l].push_back(kp);
      }
    }
    for (int l = 0; l < nlevels; ++l)
      for (cv::KeyPoint& kp : levelKps[l]) kp.angle = ic_angle(l, cvRound(kp.pt.x), cvRound(kp.pt.y));
    int n = 0;
    for (int l = 0; l < nlevels; ++l) n += (int)levelKps[l].size();
    kps.assign(n, cv::KeyPoint());
    desc.assign((size_t)n * 32, 0);
    int mono = 0, stereo = n - 1;
    for (int l = 0; l < nlevels; ++l) {
      blurred[l].assign((size_t)lw[l] * lh[l], 0);
      if (levelKps[l].empty()) continue;
      shim_gauss7(level[l].data(), lw[l], lh[l], lw[l], blurred[l].data());
      for (const cv::KeyPoint& k0 : levelKps[l]) {
        uint8_t d[32];
        descriptor(l, k0, d);
        cv::KeyPoint kp = k0;
        if (l != 0) { kp.pt.x = kp.pt.x * scale[l]; kp.pt.y = kp.pt.y * scale[l]; }
        int slot;
        if (kp.pt.x >= lap0 && kp.pt.x <= lap1) slot = stereo--; else slot = mono++;
        kps[slot] = kp;
        std::memcpy(&desc[(size_t)slot * 32], d, 32);
      }
    }
    return mono;
  }
};

int hamming256(const uint8_t* a, const uint8_t* b) {
  // ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1880-1894): 8 x 32-bit xor + popcount
  int d = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    d += __builtin_popcount(x ^ y);
  }
  return d;
}

}  // namespace

extern "C" {

void* oro_create(int nf, float sf, int nl, int ini, int mn) { return new Oracle(nf, sf, nl, ini, mn); }
void oro_destroy(void* h) { delete (Oracle*)h; }
void oro_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* nfeat, int* umax) {
  Oracle* o = (Oracle*)h;
  for (int i = 0; i < o->nlevels; ++i) {
    scale[i] = o->scale[i]; inv_scale[i] = o->invScale[i]; sigma2[i] = o->sigma2[i]; inv_sigma2[i] = o->invSigma2[i];
    nfeat[i] = o->nFeat[i];
  }
  for (int i = 0; i < 16; ++i) umax[i] = o->umax[i];
}
int oro_extract(void* h, const uint8_t* img, int w, int hgt, int stride, int lap0, int lap1, void* kps_out,
                uint8_t* desc_out, int cap, int* n_out) {
  Oracle* o = (Oracle*)h;
  std::vector<cv::KeyPoint> kps;
  std::vector<uint8_t> desc;
  int mono = o->extract(img, w, hgt, stride, lap0, lap1, kps, desc);
  *n_out = (int)kps.size();
  if (mono < 0) return mono;
  if ((int)kps.size() > cap) return -2;
  if (!kps.empty()) {
    std::memcpy(kps_out, kps.data(), kps.size() * sizeof(cv::KeyPoint));
    std::memcpy(desc_out, desc.data(), desc.size());
  }
  return mono;
}
int oro_level_size(void* h, int l, int* w, int* hgt) { Oracle* o = (Oracle*)h; *w = o->lw[l]; *hgt = o->lh[l]; return 0; }
int oro_get_level(void* h, int l, uint8_t* dst) { Oracle* o = (Oracle*)h; std::memcpy(dst, o->level[l].data(), o->level[l].size()); return 0; }
int oro_get_blurred(void* h, int l, uint8_t* dst) { Oracle* o = (Oracle*)h; std::memcpy(dst, o->blurred[l].data(), o->blurred[l].size()); return 0; }
int oro_get_candidates(void* h, int l, int32_t* xys, int cap) {
  Oracle* o = (Oracle*)h;
  int n = (int)o->cands[l].size();
  if (n > cap) return -2;
  for (int i = 0; i < n; ++i) { xys[3 * i] = o->cands[l][i].x; xys[3 * i + 1] = o->cands[l][i].y; xys[3 * i + 2] = o->cands[l][i].score; }
  return n;
}
int oro_get_level_keypoints(void* h, int l, void* kps_out, int cap) {
  Oracle* o = (Oracle*)h;
  int n = (int)o->levelKps[l].size();
  if (n > cap) return -2;
  if (n) std::memcpy(kps_out, o->levelKps[l].data(), n * sizeof(cv::KeyPoint));
  return n;
}

int oro_distribute_passes(const int32_t* cands, int n, int w, int hgt, int N, int32_t* out, int cap) {
  std::vector<Cand> in(n);
  for (int i = 0; i < n; ++i) in[i] = Cand{cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]};
  std::vector<Cand> r = distribute_octree_passes(in, w, hgt, N);
  if ((int)r.size() > cap) return -2;
  for (size_t i = 0; i < r.size(); ++i) { out[3 * i] = r[i].x; out[3 * i + 1] = r[i].y; out[3 * i + 2] = r[i].score; }
  return (int)r.size();
}

int oro_distribute(const int32_t* cands, int n, int w, int hgt, int N, int32_t* out, int cap) {
  std::vector<Cand> in(n);
  for (int i = 0; i < n; ++i) in[i] = Cand{cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]};
  std::vector<Cand> r = distribute_octree(in, w, hgt, N);
  if ((int)r.size() > cap) return -2;
  for (size_t i = 0; i < r.size(); ++i) { out[3 * i] = r[i].x; out[3 * i + 1] = r[i].y; out[3 * i + 2] = r[i].score; }
  return (int)r.size();
}

// Frame::ComputeStereoMatches (src/Frame.cc:889-1047) in the order-free form of SURVEY.md a12':
// a right keypoint iR is a candidate of iL iff (int)vL lies in [floor(yR - r), ceil(yR + r)],
// |octR - octL| <= 1 and uL - maxD <= uR <= uL; winner = min (hamming, iR) with hamming < TH_HIGH.
int oro_stereo(void* hl, void* hr, const void* kpsL_, const uint8_t* descL, int nL, const void* kpsR_,
               const uint8_t* descR, int nR, float mbf, float maxD, float* uRight, float* depth,
               int32_t* best_idx, int32_t* best_dist) {
  Oracle* L = (Oracle*)hl;
  Oracle* R = (Oracle*)hr;
  const cv::KeyPoint* kL = (const cv::KeyPoint*)kpsL_;
  const cv::KeyPoint* kR = (const cv::KeyPoint*)kpsR_;
  const int TH_HIGH = 100, TH_LOW = 50;          // src/ORBmatcher.cc:34-35
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;  // :893
  const float minD = 0;
  std::vector<int> minr(nR), maxr(nR);
  for (int i = 0; i < nR; ++i) {  // :904-912
    const float r = 2.0f * L->scale[kR[i].octave];
    maxr[i] = (int)std::ceil(kR[i].pt.y + r);
    minr[i] = (int)std::floor(kR[i].pt.y - r);
  }
  std::vector<std::pair<int, int>> distIdx;
  for (int iL = 0; iL < nL; ++iL) {
    uRight[iL] = -1.f; depth[iL] = -1.f;
    if (best_idx) best_idx[iL] = -1;
    if (best_dist) best_dist[iL] = -1;
    const cv::KeyPoint& kp = kL[iL];
    const int levelL = kp.octave;
    const float vL = kp.pt.y, uL = kp.pt.x;
    const int row = (int)vL;  // vRowIndices[vL], :929
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH, bestR = -1;
    for (int iR = 0; iR < nR; ++iR) {
      if (row < minr[iR] || row > maxr[iR]) continue;
      if (kR[iR].octave < levelL - 1 || kR[iR].octave > levelL + 1) continue;
      const float uR = kR[iR].pt.x;
      if (uR >= minU && uR <= maxU) {
        int d = hamming256(descL + 32 * (size_t)iL, descR + 32 * (size_t)iR);
        if (d < bestDist) { bestDist = d; bestR = iR; }
      }
    }
    if (best_idx) best_idx[iL] = bestR;
    if (best_dist) best_dist[iL] = bestR >= 0 ? bestDist : -1;
    if (!(bestDist < thOrbDist) || bestR < 0) continue;
    // :966-970 coordinates at the left keypoint's pyramid level (std::round: half away from zero)
    const float uR0 = kR[bestR].pt.x;
    const float sf = L->invScale[kp.octave];
    const float scaleduL = std::round(kp.pt.x * sf);
    const float scaledvL = std::round(kp.pt.y * sf);
    const float scaleduR0 = std::round(uR0 * sf);
    const int w = 5, Ls = 5;
    const int lv = kp.octave;
    const int WL = L->lw[lv], WR = R->lw[lv];
    const float iniu = scaleduR0 + Ls - w;
    const float endu = scaleduR0 + Ls + w + 1;
    if (iniu < 0 || endu >= WR) continue;  // :984-988 (verbatim)
    const uint8_t* IL = L->level[lv].data();
    const uint8_t* IR = R->level[lv].data();
    const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
    int bestSad = INT_MAX, bestInc = 0;
    float dists[11];
    for (int inc = -Ls; inc <= Ls; ++inc) {
      int sad = 0;
      for (int dy = -w; dy <= w; ++dy)
        for (int dx = -w; dx <= w; ++dx)
          sad += std::abs((int)IL[(size_t)(cy + dy) * WL + cxl + dx] - (int)IR[(size_t)(cy + dy) * WR + cxr + inc + dx]);
      float dist = (float)sad;
      if (dist < (float)bestSad) { bestSad = (int)dist; bestInc = inc; }
      dists[Ls + inc] = dist;
    }
    if (bestInc == -Ls || bestInc == Ls) continue;
    const float d1 = dists[Ls + bestInc - 1], d2 = dists[Ls + bestInc], d3 = dists[Ls + bestInc + 1];
    const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = L->scale[kp.octave] * ((float)scaleduR0 + (float)bestInc + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
      depth[iL] = mbf / disparity;
      uRight[iL] = bestuR;
      distIdx.push_back(std::make_pair(bestSad, iL));
    }
  }
  if (distIdx.empty()) return 0;  // the reference indexes an empty vector here
  std::sort(distIdx.begin(), distIdx.end());
  const float median = distIdx[distIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = 0;
  for (auto& p : distIdx) {
    if ((float)p.first < thDist) ++kept;
    else { uRight[p.second] = -1; depth[p.second] = -1; }
  }
  return kept;
}

int oro_descriptor_distance(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

int oro_knn2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int32_t* idx, int32_t* dist, int threads) {
  if (threads < 1) threads = 1;
  auto work = [&](int t) {
    for (int i = t; i < nq; i += threads) {
      int d0 = INT_MAX, d1 = INT_MAX, i0 = -1, i1 = -1;
      for (int64_t j = 0; j < ndb; ++j) {
        int d = hamming256(q + 32 * (size_t)i, db + 32 * (size_t)j);
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = (int)j; }
        else if (d < d1) { d1 = d; i1 = (int)j; }
      }
      idx[2 * i] = i0; idx[2 * i + 1] = i1;
      dist[2 * i] = i0 >= 0 ? d0 : -1; dist[2 * i + 1] = i1 >= 0 ? d1 : -1;
    }
  };
  if (threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  return 0;
}

int oro_ratio_test(const int32_t* dist, int nq, uint8_t* pass) {
  for (int i = 0; i < nq; ++i) {
    if (dist[2 * i] < 0 || dist[2 * i + 1] < 0) { pass[i] = 0; continue; }
    float d0 = (float)dist[2 * i], d1 = (float)dist[2 * i + 1];
    pass[i] = (d0 < d1 * 0.7) ? 1 : 0;  // float < float * double -> evaluated in double (src/Frame.cc:1250)
  }
  return 0;
}

void shim_resize(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh) {
  cv::Mat s(sh, sw, CV_8UC1, (void*)src, (size_t)sstride), d(dh, dw, CV_8UC1, (void*)dst, (size_t)dw);
  cv::resize(s, d, cv::Size(dw, dh), 0, 0, cv::INTER_LINEAR);
}
void shim_undistort_points(const float* pts, int n, const float* K, const float* dist, int ndist, const float* P, float* out) {
  cv::undistort_points_pinhole(pts, n, K, dist, ndist, P, out);
}

void shim_remap(const uint8_t* src, int sw, int sh, int sstride, const float* mapx, const float* mapy, int dw, int dh, uint8_t* dst) {
  cv::remap_linear_8u(src, sw, sh, sstride, mapx, mapy, dw, dh, dst, dw);
}

void shim_gauss7(const uint8_t* src, int w, int hgt, int stride, uint8_t* dst) {
  cv::Mat s(hgt, w, CV_8UC1, (void*)src, (size_t)stride), d(hgt, w, CV_8UC1, (void*)dst, (size_t)w);
  cv::GaussianBlur(s, d, cv::Size(7, 7), 2, 2, cv::BORDER_REFLECT_101);
}
int shim_fast(const uint8_t* img, int w, int hgt, int stride, int threshold, int32_t* xys, int cap) {
  cv::Mat s(hgt, w, CV_8UC1, (void*)img, (size_t)stride);
  std::vector<cv::KeyPoint> k;
  cv::FAST(s, k, threshold, true);
  if ((int)k.size() > cap) return -2;
  for (size_t i = 0; i < k.size(); ++i) { xys[3 * i] = (int)k[i].pt.x; xys[3 * i + 1] = (int)k[i].pt.y; xys[3 * i + 2] = (int)k[i].response; }
  return (int)k.size();
}
float shim_fastatan2(float y, float x) { return cv::fastAtan2(y, x); }
void shim_border101(const uint8_t* src, int w, int hgt, uint8_t* dst, int b) {
  cv::Mat s(hgt, w, CV_8UC1, (void*)src, (size_t)w), d(hgt + 2 * b, w + 2 * b, CV_8UC1, (void*)dst, (size_t)(w + 2 * b));
  cv::copyMakeBorder(s, d, b, b, b, b, cv::BORDER_REFLECT_101);
}
float restated_sinf(float x) { return sincosf_restate::sinf_r(x); }
float restated_cosf(float x) { return sincosf_restate::cosf_r(x); }
float libm_sinf(float x) { return sinf(x); }
float libm_cosf(float x) { return cosf(x); }
void oro_introsort(uint32_t* keys, uint32_t* payload, int n) {
  std::vector<SortEl> a(n);
  for (int i = 0; i < n; ++i) a[i] = SortEl{keys[i], payload[i]};
  std_sort_emulated(a.data(), n);
  for (int i = 0; i < n; ++i) { keys[i] = a[i].key; payload[i] = a[i].val; }
}

}  // extern "C"
