// TEST INFRASTRUCTURE ONLY (oracle). C interface of the CPU restatement of the reference's ORB
// front-end. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product path (morb_slam_b200/, include/) never does.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// restatement of ORBextractor (src/ORBextractor.cc:406-464 ctor, :1006-1086 operator())
void* oro_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void oro_destroy(void* h);
void oro_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* nfeat, int* umax);
// returns monoIndex (-1 on empty image, -2 capacity, -3 unsupported size); kps_out is 28-byte cv::KeyPoint records
int oro_extract(void* h, const uint8_t* img, int w, int hgt, int stride, int lap0, int lap1, void* kps_out,
                uint8_t* desc_out, int cap, int* n_out);
// stage outputs of the last oro_extract on this handle
int oro_level_size(void* h, int level, int* w, int* hgt);
int oro_get_level(void* h, int level, uint8_t* dst);        // un-blurred pyramid level, w*h contiguous
int oro_get_blurred(void* h, int level, uint8_t* dst);      // 7x7 sigma-2 blurred level (all zeros if level had no keypoint)
int oro_get_candidates(void* h, int level, int32_t* xys, int cap);  // FAST candidates (x,y,score) rel. to the 16-px border
int oro_get_level_keypoints(void* h, int level, void* kps_out, int cap);  // after octree + orientation, level coords

// DistributeOctTree restated on arrays with an explicit libstdc++-introsort emulation
// (src/ORBextractor.cc:540-738). cands/out are (x,y,score) int triples, region = [0,w) x [0,h).
int oro_distribute(const int32_t* cands, int n, int w, int hgt, int N, int32_t* out, int cap);

// Frame::ComputeStereoMatches restated (src/Frame.cc:889-1047), pyramids taken from the two handles
int oro_stereo(void* hl, void* hr, const void* kpsL, const uint8_t* descL, int nL, const void* kpsR,
               const uint8_t* descR, int nR, float mbf, float maxD, float* uRight, float* depth,
               int32_t* best_idx, int32_t* best_dist);

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1880-1894)
int oro_descriptor_distance(const uint8_t* a, const uint8_t* b);
// cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) + tie rule (src/Frame.cc:1242); idx/dist are nq x 2, -1 when absent
int oro_knn2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int32_t* idx, int32_t* dist, int threads);
// Lowe ratio gate of src/Frame.cc:1250 evaluated as the reference does (float < float * double)
int oro_ratio_test(const int32_t* dist, int nq, uint8_t* pass);

// Windowed matcher (orb_oracle_match.cc). gp = {mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv}.
// Frame::AssignFeaturesToGrid (src/Frame.cc:501-528) as CSR in the reference's cell order (ix * 48 + iy); returns entries
int oro_assign_grid(const void* kps, int n, const float* gp, int* cell_off /*[3073]*/, int* idx /*[n]*/);
// Frame::GetFeaturesInArea (src/Frame.cc:742-807)
int oro_features_in_area(const void* kps, int n, const float* gp, float x, float y, float r, int minLevel, int maxLevel, int* out,
                         int cap);
// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1521-1733), Nleft == -1, from the
// projected point on; q = {u, v, z, angle, octave, flags} per last-frame keypoint; returns nmatches
int oro_search_by_projection(const void* kpsC, const uint8_t* descC, const float* uRightC, int nC, const float* scale, int nlevels,
                             const float* gp, float mb, float mbf, const void* q, const uint8_t* qdesc, int nq, float th, int bMono,
                             float tlc_z, int check_orientation, int* match_out);

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (src/ORBmatcher.cc:42-209), Nleft == -1;
// q = {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mnTrackScaleLevel, flags}; locked0[i2]: keypoint already holds a
// map point with observations; returns nmatches
int oro_search_local_points(const void* kpsC, const uint8_t* descC, const float* uRightC, const uint8_t* locked0, int nC,
                            const float* scale, int nlevels, const float* gp, const void* q, const uint8_t* qdesc, int nq, float th,
                            float nnratio, int* match_out);

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:218-395), single camera; feature vectors as CSR in map
// order; kf_flags[i] != 0: keyframe keypoint i holds a good map point; match_out[iF] = keyframe keypoint or -1; returns nmatches
int oro_search_by_bow(const uint8_t* descKF, const float* angleKF, const uint8_t* kf_flags, int nKF, const uint32_t* kf_node, const int* kf_off,
                      const uint32_t* kf_feat, int kf_nn, const uint8_t* descF, const float* angleF, int nF, const uint32_t* f_node,
                      const int* f_off, const uint32_t* f_feat, int f_nn, float nnratio, int check_orientation, int* match_out);

// Frame::ComputeBoW = DBoW2 TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1260). Vocabulary as arrays in file order (node 0 = root);
// outputs in std::map order: BowVector (word, value), FeatureVector as CSR (node, offsets, feature indices);
// feat_word / feat_node (optional): word and node of every feature, -1 for stopped words
void* oro_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                       const uint8_t* desc, const double* weight);
void oro_vocab_free(void* v);
int oro_bow_transform(void* v, const uint8_t* desc, int n, int levelsup, int cap, uint32_t* bow_word, double* bow_val, int* bow_n,
                      uint32_t* fv_node, int* fv_off, uint32_t* fv_feat, int* fv_n, int32_t* feat_word, int32_t* feat_node);

// OpenCV-primitive restatements (the shim), exported so tests can pin them against cv2
void shim_resize(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh);
void shim_gauss7(const uint8_t* src, int w, int hgt, int stride, uint8_t* dst);
// cv::undistortPoints(pts, pts, K, dist, noArray(), P) of Frame::UndistortKeyPoints (src/Frame.cc:845-846); K, P 3 x 3 float
void shim_undistort_points(const float* pts, int n, const float* K, const float* dist, int ndist, const float* P, float* out);
// cv::remap(src, dst, mapx CV_32FC1, mapy CV_32FC1, INTER_LINEAR) with BORDER_CONSTANT 0 (System::TrackStereo, src/System.cc:260-261)
void shim_remap(const uint8_t* src, int sw, int sh, int sstride, const float* mapx, const float* mapy, int dw, int dh, uint8_t* dst);
int shim_fast(const uint8_t* img, int w, int hgt, int stride, int threshold, int32_t* xys, int cap);
float shim_fastatan2(float y, float x);
void shim_border101(const uint8_t* src, int w, int hgt, uint8_t* dst, int b);
float restated_sinf(float x);
float restated_cosf(float x);
float libm_sinf(float x);
float libm_cosf(float x);
// emulation of libstdc++ std::sort on (key, payload) pairs compared by key only (for tests)
void oro_introsort(uint32_t* keys, uint32_t* payload, int n);

#ifdef __cplusplus
}
#endif
