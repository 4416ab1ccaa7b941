/*
 * orb_b200.h - C ABI of the B200-native ORB front-end (liborb_b200.so).
 *
 * This is the drop-in boundary for the hot path of Soldann/MORB_SLAM (an ORB-SLAM3 fork). The
 * reference has no FFI layer: its boundary is the C++ class ORB_SLAM3::ORBextractor
 * (include/ORBextractor.h:44-105), Frame::ComputeStereoMatches (include/Frame.h:116,
 * src/Frame.cc:889-1047), the brute-force kNN inside Frame::ComputeStereoFishEyeMatches
 * (src/Frame.cc:1222-1274) and ORBmatcher::DescriptorDistance (include/ORBmatcher.h:41,
 * src/ORBmatcher.cc:1880-1894). The C++ shim classes in morb_slam_b200/cpp/ keep those C++
 * signatures and call the functions below; INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes only; every function returns an int status (0 = ok,
 * < 0 = error, see ORB_ERR_*), never throws. One orb_handle per camera (like one ORBextractor
 * instance per camera, src/Tracking.cc:615-624); a handle owns one CUDA stream and must not be
 * used from two threads at once; different handles may be used concurrently. All results are
 * bit-identical to the reference on the same input (angles/disparities included).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * ORB_ERR_CUDA.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_MAX_LEVELS 12

enum {
  ORB_OK = 0,
  ORB_ERR_EMPTY_IMAGE = -1,      /* ORBextractor::operator() returns -1 (src/ORBextractor.cc:1011) */
  ORB_ERR_INVALID_ARG = -2,
  ORB_ERR_CUDA = -3,             /* CUDA runtime error or no device; see orb_last_error() */
  ORB_ERR_UNSUPPORTED_SIZE = -4, /* image larger than the handle was created for, or a pyramid level
                                    too small for one 35-px FAST cell (the reference divides by zero there) */
  ORB_ERR_CAPACITY = -5,         /* an internal or caller-supplied capacity was exceeded (never silent) */
  ORB_ERR_STATE = -6             /* call order violated (e.g. stereo match before extraction) */
};

/* flags for the batch entry points */
enum {
  ORB_SRC_DEVICE = 1,  /* image pointer(s) are device memory on the handle's device */
  ORB_DST_DEVICE = 2,  /* output pointers are device memory */
  ORB_ASYNC = 4,       /* enqueue only; results are valid after orb_sync() */
  ORB_NO_OUTPUT = 8,   /* keep results device-resident only (stereo match / debug getters read them) */
  ORB_INPUT_REMAP = 16, /* orb_extract_batch: the images are raw camera frames, rectify them first (orb_set_rectify_maps) */
  ORB_INPUT_RESIZE = 32 /* orb_extract_batch: resize the images to the size set with orb_set_input_size first */
};

/* The five constructor arguments of ORBextractor (include/ORBextractor.h:48-49). */
typedef struct orb_params {
  int nfeatures;
  float scale_factor;
  int nlevels;
  int ini_th_fast;
  int min_th_fast;
} orb_params;

/* Same 28-byte layout as cv::KeyPoint: pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct orb_keypoint {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orb_keypoint;

typedef struct orb_handle orb_handle;

/* ---- extractor life cycle: replaces ORBextractor::ORBextractor (src/ORBextractor.cc:406-464) ---- */
int orb_create(const orb_params* params, int max_width, int max_height, int max_batch, int device,
               orb_handle** out);
int orb_destroy(orb_handle* h);
const char* orb_last_error(const orb_handle* h); /* text of the last failure on this handle */
const char* orb_status_string(int status);
/* maximum number of keypoints one image can yield: nfeatures + 3 * nlevels */
int orb_keypoint_capacity(const orb_handle* h);
/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:60-74) and mnFeaturesPerLevel; arrays of nlevels entries, NULL to skip */
int orb_get_tables(const orb_handle* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                   int* features_per_level);
/* The same tables from the constructor arguments alone, without a handle or a device (pure host arithmetic of
 * src/ORBextractor.cc:413-443): the reference fills them in its constructor and every Frame constructor reads the getters
 * BEFORE the first extraction (src/Frame.cc:181-187), so the drop-in class fills its members with this call. */
int orb_compute_tables(const orb_params* params, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                       int* features_per_level);

/* ---- ORBextractor::operator() (src/ORBextractor.cc:1006-1086), one host image, synchronous ----
 * image: 8-bit single channel, `stride` bytes per row. lap0/lap1 = vLappingArea. Outputs: kps_out
 * (cap records), desc_out (cap x 32 bytes), *n_out = number of keypoints, *mono_out = the value
 * operator() returns (monoIndex). Returns ORB_ERR_EMPTY_IMAGE for a null/empty image. */
int orb_extract(orb_handle* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1,
                orb_keypoint* kps_out, uint8_t* desc_out, int cap, int* n_out, int* mono_out);

/* ---- the same for a batch of equally sized images (independent frames, one launch sequence) ----
 * images: base pointer of frame 0; frame i starts at images + i * image_stride. kps_out / desc_out
 * hold `cap` records per frame (frame i at i * cap); n_out / mono_out hold `batch` ints. With
 * ORB_ASYNC the call only enqueues work on the handle's stream (host buffers should be pinned,
 * see orb_host_alloc) and orb_sync() completes it. A frame whose internal capacities overflow is
 * reported by orb_sync()/the call returning ORB_ERR_CAPACITY and n_out[i] = -1. */
int orb_extract_batch(orb_handle* h, const uint8_t* images, int batch, int width, int height, size_t stride,
                      size_t image_stride, int lap0, int lap1, orb_keypoint* kps_out, uint8_t* desc_out, int cap,
                      int* n_out, int* mono_out, int flags);
int orb_sync(orb_handle* h);

/* ---- the step before the extractor: System::TrackStereo's cv::remap(imLeft, imLeftToFeed, M1l, M2l, cv::INTER_LINEAR)
 * (src/System.cc:254-261) for settings with needToRectify() (EuRoC "PinHole" stereo). map_x / map_y are the CV_32FC1
 * maps of cv::initUndistortRectifyMap (src/Settings.cc:540-545), map_w x map_h = the rectified image size (host
 * pointers, uploaded once per handle = per camera; NULL clears them). With ORB_INPUT_REMAP, orb_extract_batch takes
 * the RAW frames (width x height = the camera's size), remaps them on the device with OpenCV's fixed-point bilinear
 * arithmetic (5-bit fractions, 15-bit weights, BORDER_CONSTANT 0; bit-identical to cv::remap) into level 0 of the
 * pyramid and extracts from that; orb_pyramid_level(h, frame, 0, ...) returns the rectified image (imLeftToFeed). ---- */
int orb_set_rectify_maps(orb_handle* h, const float* map_x, const float* map_y, int map_w, int map_h);
/* The other input stage of System::TrackStereo / TrackMonocular / TrackRGBD: cv::resize(im, imToFeed, settings_->newImSize())
 * for settings with needToResize() (src/System.cc:262-264, INTER_LINEAR). With ORB_INPUT_RESIZE, orb_extract_batch resizes the
 * frames it is given to new_w x new_h on the device (cv::resize's 8-bit fixed-point arithmetic, bit-identical; an exact 2x shrink
 * is OpenCV's 2x2 box filter) into level 0 and extracts from that. new_w = 0 clears. */
int orb_set_input_size(orb_handle* h, int new_w, int new_h);

/* ---- ORBextractor::mvImagePyramid (include/ORBextractor.h:76): un-blurred level `level` of frame
 * `frame` of the last call, copied to host memory (dst_stride bytes per row) ---- */
int orb_pyramid_level_size(const orb_handle* h, int level, int* width, int* height);
int orb_pyramid_level(orb_handle* h, int frame, int level, uint8_t* dst, size_t dst_stride);

/* ---- Frame::ComputeStereoMatches (src/Frame.cc:889-1047) ----
 * Batch form: matches frame i of hL's last batch against frame i of hR's last batch using the
 * device-resident keypoints, descriptors and pyramids. uright_out / depth_out: `cap` floats per
 * frame (-1 = no match), mvuRight / mvDepth of the reference. max_d = mbf / mb (the reference reads
 * mb before assigning it, src/Frame.cc:915 vs :253, so the caller passes it). */
int orb_stereo_match_batch(orb_handle* hL, orb_handle* hR, float mbf, float max_d, float* uright_out,
                           float* depth_out, int cap, int flags);
/* Single-frame form with host keypoints/descriptors (those of mvKeys/mvKeysRight, mDescriptors/
 * mDescriptorsRight); the pyramids are those of frame 0 of the two handles' last extraction. */
int orb_stereo_match(orb_handle* hL, orb_handle* hR, const orb_keypoint* kpsL, const uint8_t* descL, int nL,
                     const orb_keypoint* kpsR, const uint8_t* descR, int nR, float mbf, float max_d,
                     float* uright_out, float* depth_out);

/* ---- Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1222-1250) up to the triangulation (next entry point), for every frame of the two
 * handles' last batches (extracted with a lapping area: `mono` = monoLeft / monoRight): per frame
 * BFmatcher.knnMatch(mDescriptors.rowRange(monoLeft, N), mDescriptorsRight.rowRange(monoRight, Nright), matches, 2) and Lowe's
 * ratio (*it)[0].distance < (*it)[1].distance * 0.7 on the device-resident descriptors. For query i (left keypoint
 * monoLeft + i): idx_out[(frame * cap + i) * 2 + k] = trainIdx of neighbour k (right keypoint monoRight + trainIdx, -1 when
 * the right side has fewer than k + 1 keypoints), dist_out likewise, pass_out[frame * cap + i] = the ratio test. ---- */
int orb_stereo_fisheye_match_batch(orb_handle* hL, orb_handle* hR, int32_t* idx_out, int32_t* dist_out, uint8_t* pass_out,
                                   int cap, int flags);

/* ---- the rest of Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1252-1277): for every left keypoint whose best match passed
 * the ratio test, KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:323-395: unproject both keypoints
 * :116-147, parallax gate cos > 0.9998, KannalaBrandt8::Triangulate :415-428 = smallest right singular vector of the 4 x 4 DLT
 * matrix, positive depth in both cameras, reprojection error against 5.991 * mvLevelSigma2[octave] in both images via
 * project :68-94) and the acceptance depth > 0.0001f. Runs on the results orb_stereo_fisheye_match_batch left on the device and
 * the resident keypoints of the two handles.
 * The rig is what the Frame holds: mpCamera / mpCamera2 parameters (fx fy cx cy k0 k1 k2 k3 = KannalaBrandt8::mvParameters),
 * KannalaBrandt8::precision (1e-6 in Settings), mRlr row-major and mtlr (src/Frame.cc:1260-1262).
 * Floating point: the reference computes this in float with glibc's tanf / atan2f / cos / sin and Eigen's float JacobiSVD; here the
 * float expressions are kept as written (no contraction), the four libm routines are glibc 2.39's restated for the device (pinned
 * exhaustively against the image's libm on the host) and the null vector comes from Eigen's two-sided float Jacobi SVD restated from
 * its published algorithm. Return codes, depths and 3-D points equal the oracle's bit for bit (tests/test_gpu_fisheye.py); what
 * cannot be checked in this image is the restated SVD / reduction order against a build of Eigen itself.
 * Outputs (`cap` entries per frame, host or device like the other batch calls, any may be NULL):
 *   left_to_right[frame * cap + i]  = mvLeftToRightMatch[i]  (right keypoint index incl. monoRight, -1 = none), i < Nleft
 *   right_to_left[frame * cap + j]  = mvRightToLeftMatch[j]  (the LAST accepted left keypoint in query order, like the loop)
 *   depth[frame * cap + i]          = mvDepth[i] (-1 = none);   p3d[(frame * cap + i) * 3 ..] = mvStereo3Dpoints[i] (0 0 0 = none)
 *   code[frame * cap + i]           = 0: no ratio-test match, 1: accepted, -1 .. -5: the negative return value of
 *                                     TriangulateMatches (parallax, z1, z2, reprojection 1, reprojection 2), -6: 0 < depth <= 0.0001 ---- */
typedef struct orb_kb8_rig {
  float cam1[8], cam2[8];        /* KannalaBrandt8::mvParameters of mpCamera / mpCamera2 */
  float precision1, precision2;  /* KannalaBrandt8::precision of the two cameras */
  float R12[9], t12[3];          /* mRlr (row-major), mtlr */
} orb_kb8_rig;
int orb_stereo_fisheye_triangulate_batch(orb_handle* hL, orb_handle* hR, const orb_kb8_rig* rig, int32_t* left_to_right,
                                         int32_t* right_to_left, float* depth, float* p3d, int8_t* code, int cap, int flags);
/* KannalaBrandt8::TriangulateMatches on n explicit keypoint pairs (host arrays; xy1 / xy2: n x 2 floats, sigma1 / sigma2: n):
 * ret[i] = the function's return value (depth, or -1 .. -5), p3d[3 i ..] = the point when ret > 0. */
int orb_kb8_triangulate_matches(orb_handle* h, const orb_kb8_rig* rig, const float* xy1, const float* xy2, const float* sigma1,
                                const float* sigma2, int n, float* ret, float* p3d);

/* ---- brute-force top-2 Hamming kNN: cv::BFMatcher(NORM_HAMMING).knnMatch(q, db, k=2) as used at
 * src/Frame.cc:1242, ties resolved towards the lower database index ----
 * q: nq x 32 bytes, db: ndb x 32 bytes. idx_out/dist_out: nq x 2 int32 (idx = index_base + row,
 * -1/-1 when fewer than 2 rows). Pointers are host memory unless the ORB_*_DEVICE flags are set.
 * `h` supplies the device and stream (any handle created on that device). */
int orb_hamming_knn2(orb_handle* h, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int32_t index_base,
                     int32_t* idx_out, int32_t* dist_out, int flags);
/* Merge `nparts` partial top-2 lists (as produced on disjoint, index-contiguous database shards and
 * gathered rank-major: part p at p * nq * 2) into the global top-2 by (distance, index). */
int orb_knn2_merge(orb_handle* h, const int32_t* idx_parts, const int32_t* dist_parts, int nparts, int nq,
                   int32_t* idx_out, int32_t* dist_out, int flags);
/* ---- the sharded search (BASELINE.json configs[4]: database rows sharded contiguously over the GPUs of one box) with the exchange
 * fused into the kernels over peer memory instead of a collective call: every rank scans its shard, the kernel that merges the
 * scan's partial lists stores the rank's top-2 straight into every rank's exchange buffer (NVLink peer stores) and publishes an
 * epoch flag; a second kernel waits for all ranks' flags and merges the `world` lists by (distance, global index). The result is
 * identical on every rank and equal to one brute-force scan (shards are index-contiguous, ties keep the lowest index).
 * Set-up, one process per GPU: orb_knn_exchange_create on every rank -> exchange the ORB_IPC_HANDLE_BYTES-byte handles by any
 * transport (e.g. one torch.distributed all_gather) -> orb_knn_exchange_connect with all handles in rank order. Ranks that live in
 * ONE process (several handles, tests) connect with orb_knn_exchange_connect_local instead. The search is a collective: every rank
 * calls it the same number of times. q / db_local / outputs are device pointers (ORB_SRC_DEVICE | ORB_DST_DEVICE required);
 * a peer that does not arrive within ORB_KNN_PEER_TIMEOUT_S seconds turns into ORB_ERR_STATE, never a hang. ---- */
#define ORB_KNN_MAX_RANKS 16
#define ORB_IPC_HANDLE_BYTES 64
#define ORB_KNN_PEER_TIMEOUT_S 30   /* default; ORB_B200_KNN_TIMEOUT_S in the environment of orb_knn_exchange_create overrides it (1 .. 3600) */
typedef struct orb_knn_exchange orb_knn_exchange;
int orb_knn_exchange_create(orb_handle* h, int rank, int world, int max_nq, orb_knn_exchange** out, uint8_t* ipc_handle_out);
int orb_knn_exchange_connect(orb_knn_exchange* x, const uint8_t* all_handles);
int orb_knn_exchange_connect_local(orb_knn_exchange* x, orb_knn_exchange* const* all);
/* synchronises the handle's stream and returns ORB_ERR_STATE if a search enqueued with ORB_ASYNC ran into the peer time-out */
int orb_knn_exchange_check(orb_knn_exchange* x);
int orb_knn_exchange_destroy(orb_knn_exchange* x);
int orb_hamming_knn2_sharded(orb_handle* h, orb_knn_exchange* x, const uint8_t* q, int nq, const uint8_t* db_local, int64_t ndb_local,
                             int32_t index_base, int32_t* idx_out, int32_t* dist_out, int flags);
/* Lowe ratio gate of src/Frame.cc:1250: pass[i] = has two neighbours && (double)d0 < (double)d1 * 0.7 */
int orb_ratio_test(orb_handle* h, const int32_t* dist, int nq, uint8_t* pass_out, int flags);

/* ---- Frame::UndistortKeyPoints (src/Frame.cc:829-857) on the device-resident keypoints of the handle's last batch:
 * cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) - OpenCV's five fixed iterations in double, results bit-identical
 * floats. K = static_cast<Pinhole*>(mpCamera)->toK(), P = mK, both 3 x 3 row-major float; dist = mDistCoef (k1 k2 p1 p2 [k3 ...],
 * at most 12 coefficients). dist[0] == 0 (or ndist == 0) means mvKeysUn = mvKeys like the reference. The result becomes the
 * frame's mvKeysUn: orb_assign_features_to_grid and the two searches read it instead of mvKeys until the next extraction.
 * kps_un_out (optional, `cap` records per frame) receives mvKeysUn. ---- */
int orb_undistort_keypoints(orb_handle* h, const float* K, const float* dist, int ndist, const float* P, orb_keypoint* kps_un_out,
                            int cap, int flags);

/* ---- windowed matcher on the device-resident results of the last extraction ----
 * Frame::AssignFeaturesToGrid (src/Frame.cc:501-528, PosInGrid :809-820) for frames without a second camera
 * (Nleft == -1: monocular / rectified stereo, mvKeysUn == the extractor's keypoints): builds the 64 x 48 cell lists
 * (include/Frame.h:44-45,252) of every frame of the handle's last batch on the device. The parameters are the
 * static members the reference computes once in the Frame constructor (src/Frame.cc:236-241). */
typedef struct orb_grid_params {
  float min_x, min_y, max_x, max_y; /* mnMinX, mnMinY, mnMaxX, mnMaxY */
  float grid_w_inv, grid_h_inv;     /* mfGridElementWidthInv, mfGridElementHeightInv */
} orb_grid_params;
int orb_assign_features_to_grid(orb_handle* h, const orb_grid_params* gp, int flags);
/* one frame's grid as CSR in the reference's cell order (cell = ix * 48 + iy): cell_off[3073], idx[cap] */
int orb_debug_get_grid(orb_handle* h, int frame, int32_t* cell_off, int32_t* idx, int cap, int* n_out);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
 * (src/ORBmatcher.cc:1521-1733, Nleft == -1) for every frame of the handle's last batch = the CurrentFrames.
 * The caller projects the map points of the LastFrame with its own pose / camera model (:1545-1553: x3Dc = Tcw * x3Dw,
 * uv = project(x3Dc)) and passes one query per last-frame keypoint i: (u, v) = uv, z = x3Dc(2), the angle and octave
 * of LastFrame.mvKeys[i] and flags (bit 0: mvpMapPoints[i] != NULL && !mvbOutlier[i]; bit 1: Observations() > 0),
 * plus the map point's descriptor (qdesc, 32 bytes per query). tlc_z[frame] = tlc(2) (:1534), mb / mbf those of the
 * CurrentFrame, mono = bMono. Everything after the projection is done here, in the reference's order: image bounds,
 * GetFeaturesInArea window (src/Frame.cc:742-807) with the forward / backward level gates, mvuRight gate, best
 * descriptor with strict "<", greedy assignment that skips keypoints already holding a map point with observations,
 * rotation histogram and ComputeThreeMaxima (:1844-1876). mvuRight comes from the handle's last stereo match
 * (-1 everywhere when none ran). match_out[frame * kcap + i2] = index i of the query whose map point
 * CurrentFrame.mvpMapPoints[i2] holds at the end, or -1; nmatches_out[frame] = the function's return value.
 * queries / qdesc hold qcap records per frame, nq[frame] of them used. ORB_SRC_DEVICE / ORB_DST_DEVICE / ORB_ASYNC /
 * ORB_NO_OUTPUT as for the other batch calls. */
typedef struct orb_proj_query {
  float u, v, z; /* projection of the map point and its depth in the current camera */
  float angle;   /* LastFrame keypoint angle (degrees) */
  int32_t octave;
  int32_t flags;
} orb_proj_query;
int orb_search_by_projection(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                             float th, int mono, const float* tlc_z, float mb, float mbf, int check_orientation,
                             int32_t* match_out, int32_t* nmatches_out, int flags);

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:42-209, Nleft == -1; RadiusByViewingCos :211-216) - the local-map search of
 * Tracking::SearchLocalPoints (src/Tracking.cc) - for every frame of the handle's last batch = F.
 * The caller runs Frame::isInFrustum (pose, camera model, viewing angle, predicted scale: host glue) and passes one
 * query per local map point: mTrackProjX / mTrackProjY / mTrackProjXR / mTrackViewCos / mnTrackScaleLevel and
 * flags (bit 0: mbTrackInView && !isBad() && !(bFarPoints && mTrackDepth > thFarPoints); bit 1: Observations() > 0),
 * plus GetDescriptor() (qdesc, 32 bytes per query). locked0[frame * kcap + i2] != 0 when F.mvpMapPoints[i2] already
 * holds a map point with Observations() > 0 when the call starts (NULL: none). Everything after that is done here in
 * the reference's order: window radius from the viewing angle, GetFeaturesInArea at levels [level - 1, level], lock and
 * mvuRight gates, best / second-best descriptor with strict "<", ratio test (nnratio = mfNNratio) when both share the
 * level, TH_HIGH, greedy assignment in map-point order. match_out[frame * kcap + i2] = index of the query the call
 * assigned to keypoint i2, or -1 (keypoints that kept their previous map point report -1); nmatches_out[frame] = the
 * function's return value. Flags as for orb_search_by_projection. */
typedef struct orb_track_query {
  float proj_x, proj_y, proj_xr; /* MapPoint::mTrackProjX / Y / XR (include/MapPoint.h:171-179) */
  float view_cos;                /* mTrackViewCos */
  int32_t level;                 /* mnTrackScaleLevel */
  int32_t flags;
} orb_track_query;
int orb_search_local_points(orb_handle* h, const orb_track_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                            const uint8_t* locked0, float th, float nnratio, int32_t* match_out, int32_t* nmatches_out, int flags);

/* ---- Frame::ComputeBoW (src/Frame.cc:822-827): mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) on the
 * device-resident descriptors of the handle's last batch. The vocabulary is DBoW2's k-ary tree
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h); orb_vocab_load_text reads the ORBvoc.txt text format and replaces
 * ORBVocabulary::loadFromTextFile (:1338-1426, called at src/System.cc:132), orb_vocab_create takes the same content as
 * arrays in file order: node 0 is the root, node i > 0 has parent[i] < i, is_leaf[i] (the file's flag: leaves are
 * numbered as words in file order), desc[i] (32 bytes) and weight[i]. scoring / weighting are DBoW2's ScoringType /
 * WeightingType values (ORBvoc.txt: 0 = L1_NORM, 0 = TF_IDF). A vocabulary belongs to one device and may be shared by
 * the handles of that device. Errors: ORB_ERR_INVALID_ARG for a header the reference rejects (:1359) or a malformed
 * file, ORB_ERR_CUDA without a device. info6 = k, L, scoring, weighting, nodes, words. ---- */
typedef struct orb_vocab orb_vocab;
int orb_vocab_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                     const uint8_t* is_leaf, const uint8_t* desc, const double* weight, orb_vocab** out);
int orb_vocab_load_text(int device, const char* path, orb_vocab** out);
int orb_vocab_info(const orb_vocab* v, int32_t* info6);
int orb_vocab_destroy(orb_vocab* v);
/* Outputs of orb_compute_bow, every pointer optional (NULL = not wanted), host or device memory, per frame of the batch
 * with kcap = orb_keypoint_capacity() entries per frame:
 *   bow_n[frame], bow_word / bow_val[frame * kcap + i]: the BowVector in std::map order (ascending word id), values
 *     after the normalisation the scoring asks for (BowVector.cpp:62-86), bit-identical doubles;
 *   fv_n[frame], fv_node[frame * kcap + j], fv_off[frame * (kcap + 1) + j], fv_feat[frame * kcap + o]: the
 *     FeatureVector in std::map order as CSR: node j holds the feature indices fv_feat[fv_off[j] .. fv_off[j + 1]);
 *   feat_word / feat_node[frame * kcap + f]: word and node (levelsup levels above the leaf) of feature f, -1 when the
 *     word is stopped (weight 0, :1157). */
typedef struct orb_bow_out {
  int32_t* bow_n;
  uint32_t* bow_word;
  double* bow_val;
  int32_t* fv_n;
  uint32_t* fv_node;
  int32_t* fv_off;
  uint32_t* fv_feat;
  int32_t* feat_word;
  int32_t* feat_node;
} orb_bow_out;
int orb_compute_bow(orb_handle* h, const orb_vocab* v, int levelsup, const orb_bow_out* out, int flags);

/* ---- ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) (src/ORBmatcher.cc:218-395, single
 * camera; called by Tracking::TrackReferenceKeyFrame and Tracking::Relocalization): every frame of the handle's last batch (whose
 * FeatureVector orb_compute_bow left on the device) against one keyframe each. Per frame the caller passes the keyframe's
 * mDescriptors (desc, 32 bytes per keypoint), the angles of its mvKeysUn, flags (!= 0: GetMapPointMatches()[i] is a map point that
 * is not bad), its keypoint count n and its mFeatVec in the CSR form orb_compute_bow produces (fv_node / fv_off / fv_feat / fv_n),
 * `cap` records per frame (fv_off: cap + 1). Keypoints of shared vocabulary nodes are compared all against all in the
 * reference's order: best frame keypoint not taken yet, distance <= TH_LOW, ratio nnratio = mfNNratio against the second best,
 * rotation histogram + ComputeThreeMaxima when check_orientation. match_out[frame * kcap + iF] = keyframe keypoint whose map point
 * vpMapPointMatches[iF] receives, or -1; nmatches_out[frame] = the return value. ---- */
typedef struct orb_bow_keyframes {
  const uint8_t* desc;
  const float* angle;
  const uint8_t* flags;
  const int32_t* n;
  const uint32_t* fv_node;
  const int32_t* fv_off;
  const uint32_t* fv_feat;
  const int32_t* fv_n;
  int32_t cap;
} orb_bow_keyframes;
int orb_search_by_bow(orb_handle* h, const orb_bow_keyframes* kf, float nnratio, int check_orientation, int32_t* match_out,
                      int32_t* nmatches_out, int flags);

/* ---- two-camera frames (Frame::Nleft != -1: the fisheye rig of TUM-VI) --------------------------------------
 * A two-camera Frame keeps the left keypoints (mvKeys), the right ones (mvKeysRight), the second grid mGridRight
 * (src/Frame.cc:510-526) and mvpMapPoints over the combined index space [0, Nleft) + [Nleft, N). Here the two cameras are the
 * two handles hL / hR (same device, same batch): orb_assign_features_to_grid on hR IS mGridRight, and a window search on hR's
 * grid is GetFeaturesInArea(..., bRight = true) (src/Frame.cc:783-790). Results come back per camera: match_left_out
 * [batch][kcap of hL], match_right_out [batch][kcap of hR] (right keypoint j = combined index Nleft + j).
 *
 * orb_search_by_projection_stereo: ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) with a two-camera
 * CurrentFrame (src/ORBmatcher.cc:1521-1733; right-camera half :1638-1707). Queries as for orb_search_by_projection plus
 * (ur, vr) = mpCamera->project(GetRelativePoseTrl() * x3Dc) (:1639-1640, host glue like the left projection). A query whose
 * left window is empty or invalid skips the right camera too (the `continue`s of :1552-1581); there is no mvuRight gate
 * (Nleft != -1, :1595); the rotation histogram is shared by both cameras.
 *
 * orb_search_local_points_stereo: ORBmatcher::SearchByProjection(F, vpMapPoints, th, ...) with a two-camera F
 * (src/ORBmatcher.cc:42-209; right-camera half :127-205). Flags: bit 0 mbTrackInView, bit 1 Observations() > 0, bit 2
 * mbTrackInViewR (bits 0 / 2 after the isBad / far-point filter of :52-56). A left match also assigns the stereo partner
 * mvLeftToRightMatch[best] in the right camera and vice versa (:127-133, :192-198) - an unconditional overwrite that can
 * also release a lock; a left ratio-test failure skips the right camera (:123). The right window radius has no `th` factor
 * (:131). left_to_right [batch][kcap of hL] / right_to_left [batch][kcap of hR]: NULL = the device-resident result of
 * orb_stereo_fisheye_triangulate_batch on (hL, hR). locked0_left / locked0_right as locked0 of orb_search_local_points. */
/* orb_compute_bow_stereo: Frame::ComputeBoW of a two-camera frame - mDescriptors holds the left rows followed by the right rows
 * (src/Frame.cc:1157 cv::vconcat), so feature i < Nleft is left keypoint i and feature Nleft + j right keypoint j. Outputs as
 * orb_compute_bow with cap2 = kcap(hL) + kcap(hR) entries per frame; the vectors stay on the device (left handle) for
 * orb_search_by_bow_stereo: ORBmatcher::SearchByBoW(pKF, F, ...) with a two-camera F (src/ORBmatcher.cc:218-395; :264-360 keep a
 * best / second best per camera, the right-camera match needs the LEFT best below TH_LOW and skips its own ratio test, `|| true`). */
int orb_compute_bow_stereo(orb_handle* hL, orb_handle* hR, const orb_vocab* v, int levelsup, const orb_bow_out* out, int flags);
int orb_search_by_bow_stereo(orb_handle* hL, orb_handle* hR, const orb_bow_keyframes* kf, float nnratio, int check_orientation,
                             int32_t* match_left_out, int32_t* match_right_out, int32_t* nmatches_out, int flags);
typedef struct orb_proj_query2 {
  float u, v, z; /* left projection of the map point and its depth */
  float angle;   /* LastFrame keypoint angle (degrees) */
  int32_t octave;
  int32_t flags;
  float ur, vr;  /* projection into the right camera */
} orb_proj_query2;
int orb_search_by_projection_stereo(orb_handle* hL, orb_handle* hR, const orb_proj_query2* queries, const uint8_t* qdesc, const int32_t* nq,
                                    int qcap, float th, int mono, const float* tlc_z, float mb, int check_orientation,
                                    int32_t* match_left_out, int32_t* match_right_out, int32_t* nmatches_out, int flags);
typedef struct orb_track_query2 {
  float proj_x, proj_y, view_cos;     /* mTrackProjX / Y, mTrackViewCos */
  int32_t level;                      /* mnTrackScaleLevel */
  float proj_xr, proj_yr, view_cos_r; /* mTrackProjXR / YR, mTrackViewCosR */
  int32_t level_r;                    /* mnTrackScaleLevelR (-1: none) */
  int32_t flags, pad;
} orb_track_query2;
int orb_search_local_points_stereo(orb_handle* hL, orb_handle* hR, const orb_track_query2* queries, const uint8_t* qdesc, const int32_t* nq,
                                   int qcap, const uint8_t* locked0_left, const uint8_t* locked0_right, const int32_t* left_to_right,
                                   const int32_t* right_to_left, float th, float nnratio, int32_t* match_left_out,
                                   int32_t* match_right_out, int32_t* nmatches_out, int flags);

/* ---- widening beyond SURVEY.md 8: the LocalMapping / Relocalization consumers of the same window and
 * Hamming primitives. KeyFrames are frames the host's map already holds: orb_load_frames makes up to max_batch of them the handle's
 * resident batch (as if they had just been extracted), after which orb_assign_features_to_grid and the searches below run on them.
 * kps = mvKeysUn (the keypoints the grid is built from, src/KeyFrame.cc:112-113 copies Frame's grid), desc = mDescriptors,
 * uright = mvuRight (NULL: monocular, every entry -1), n[frame] keypoints, `cap` records per frame (n[frame] <= cap and
 * <= orb_keypoint_capacity()). ORB_SRC_DEVICE / ORB_ASYNC as elsewhere. The pyramids of the handle are NOT those of the loaded
 * frames: orb_stereo_match_* / orb_pyramid_level return ORB_ERR_STATE until the next extraction. ---- */
int orb_load_frames(orb_handle* h, const orb_keypoint* kps, const uint8_t* desc, const float* uright, const int32_t* n, int batch, int cap,
                    int flags);

/* ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th, bRight) (src/ORBmatcher.cc:1044-1215; called by
 * LocalMapping::SearchInNeighbors) and ORBmatcher::Fuse(KeyFrame *pKF, Sophus::Sim3f &Scw, vpPoints, th, vpReplacePoint) (:1217-1322,
 * loop closing), the SEARCH of both: for every frame of the resident batch (= pKF, loaded with orb_load_frames, grid built) and
 * every map point the window search and the choice of the best keypoint. The caller does what precedes the window with its own
 * pose / camera model and map-point state (:1076-1128: isBad, IsInKeyFrame, p3Dc = Tcw * p3Dw, depth sign, uv = project(p3Dc),
 * IsInImage, distance range, viewing angle, nPredictedLevel = PredictScale(dist3D, pKF)) and passes per map point (u, v) = uv,
 * ur = uv(0) - bf * invz, level = nPredictedLevel, flags bit 0 = "reached the window" and GetDescriptor() (qdesc). Done here in the
 * reference's order: radius = th * mvScaleFactors[level], KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:729-774), the level gate
 * [level - 1, level], mode 0 only: the reprojection gates e2 * mvInvLevelSigma2[kpLevel] > 7.8 (mvuRight[idx] >= 0) / > 5.99
 * (:1149-1171), best descriptor with strict "<" in GetFeaturesInArea order.
 * best_idx_out[frame * qcap + i] = bestIdx (keypoint of the frame) or -1, best_dist_out = bestDist (256 when none; the Sim3 overload's
 * INT_MAX start value cannot be told apart from 256 by its `bestDist <= TH_LOW` test). The caller replays the side effects
 * (Replace / AddObservation / AddMapPoint, :1195-1210) in map-point order: they only ever turn later map points into "skip"
 * (isBad after Replace), never change a later search, so the replay re-checks isBad() / IsInKeyFrame() and is exact.
 * mode 0: first overload (with the reprojection gates), 1: Sim3 overload (none). ---- */
typedef struct orb_fuse_query {
  float u, v;    /* uv */
  float ur;      /* uv(0) - bf * invz (mode 0) */
  int32_t level; /* nPredictedLevel */
  int32_t flags; /* bit 0: the map point reaches the window search */
} orb_fuse_query;
int orb_fuse_search(orb_handle* h, const orb_fuse_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap, float th, int mode,
                    int32_t* best_idx_out, int32_t* best_dist_out, int flags);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 * (src/ORBmatcher.cc:1735-1842; Tracking::Relocalization) for every frame of the resident batch = CurrentFrame. One query per
 * keyframe map point i (vpMPs[i]): (u, v) = project(Tcw * x3Dw), angle = pKF->mvKeysUn[i].angle, octave = nPredictedLevel =
 * PredictScale(dist3D, &CurrentFrame), flags bit 0 = pMP && !isBad() && !sAlreadyFound.count(pMP) && distance in range (z is not
 * read). locked0[frame * kcap + i2] != 0: CurrentFrame.mvpMapPoints[i2] != NULL when the call starts (NULL: none). Done here: image
 * bounds, window th * mvScaleFactors[octave] at levels [octave - 1, octave + 1], best descriptor among the keypoints without a map
 * point (every assignment of this call locks its keypoint), bestDist <= orb_dist, rotation histogram + ComputeThreeMaxima.
 * Outputs as for orb_search_by_projection. ---- */
int orb_search_by_projection_kf(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                const uint8_t* locked0, float th, int orb_dist, int check_orientation, int32_t* match_out,
                                int32_t* nmatches_out, int flags);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (src/ORBmatcher.cc:821-1042;
 * LocalMapping::CreateNewMapPoints) for `npairs` keyframe pairs, single-camera keyframes (mpCamera2 == NULL; Pinhole
 * epipolarConstrain, src/CameraModels/Pinhole.cpp). The keyframes are rows of a set: keypoints (mvKeysUn), descriptors, mvuRight
 * (NULL: all -1), has_mp[i] != 0 when GetMapPoint(i) != NULL, n, and mFeatVec in the CSR form orb_compute_bow produces, `cap`
 * records per keyframe (fv_off: cap + 1). pair p = (kf1[p], kf2[p]) indexes the set; F12[p * 9 ..] is the fundamental matrix the
 * reference builds inside every epipolarConstrain call (K1^-T [t12]x R12 K2^-1, row-major; the caller computes it once per pair
 * with its own pose types), ep[p * 2 ..] the epipole of :835. sigma2 = mvLevelSigma2 and scale = mvScaleFactors of the handle.
 * Every keypoint idx1 of pKF1 without a map point scans the keypoints of the same vocabulary node of pKF2 in the reference's
 * order: distance <= TH_LOW and <= bestDist (ties move to the later keypoint, :926), epipole distance (:943-950), epipolar line
 * distance dsqr < 3.84 * unc unless coarse; then the rotation histogram. match12_out[p * cap + idx1] = vMatches12[idx1] (-1 = none;
 * vMatchedPairs is its list of (i, match) in ascending i), nmatches_out[p] = the return value. The reference never sets
 * vbMatched2 (:868, :916), so the keypoints of pKF1 are independent. ---- */
typedef struct orb_kf_set {
  const orb_keypoint* kps;
  const uint8_t* desc;
  const float* uright;
  const uint8_t* has_mp;
  const int32_t* n;
  const uint32_t* fv_node;
  const int32_t* fv_off;
  const uint32_t* fv_feat;
  const int32_t* fv_n;
  int32_t count; /* keyframes in the set */
  int32_t cap;
} orb_kf_set;
int orb_search_for_triangulation(orb_handle* h, const orb_kf_set* kfs, const int32_t* kf1, const int32_t* kf2, const float* F12,
                                 const float* ep, int npairs, int only_stereo, int coarse, int check_orientation, int32_t* match12_out,
                                 int32_t* nmatches_out, int flags);

/* The same between TWO-CAMERA keyframes (mpCamera2 != NULL on both: the fisheye rig; src/ORBmatcher.cc:891-903, :935-981). A keyframe's
 * row of the set holds the left keypoints (mvKeys) followed by the right ones (mvKeysRight), nleft[k] = NLeft of keyframe k, and
 * descriptors / has_mp / mFeatVec cover that combined index space (uright is not read: bStereo1 / bStereo2 are false with a second
 * camera, so only_stereo yields no match, like the reference). The epipolar constraint is KannalaBrandt8::epipolarConstrain
 * (src/CameraModels/KannalaBrandt8.cpp:229-236) = TriangulateMatches(...) > 0.0001f with the cameras and the relative pose of the
 * (camera of kp1, camera of kp2) combination: rigs[p * 4 + 0 .. 3] = left-left, left-right, right-left, right-right, each with
 * cam1 = the camera of kp1 (pKF1->mpCamera or mpCamera2), cam2 = the camera of kp2, R12 / t12 = rotation and translation of
 * Tll / Tlr / Trl / Trr (:846-855; the caller multiplies the poses with its own pose types). Outputs as above. */
int orb_search_for_triangulation_fisheye(orb_handle* h, const orb_kf_set* kfs, const int32_t* nleft, const int32_t* kf1, const int32_t* kf2,
                                         const orb_kb8_rig* rigs, int npairs, int only_stereo, int coarse, int check_orientation,
                                         int32_t* match12_out, int32_t* nmatches_out, int flags);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (src/ORBmatcher.cc:702-819; LoopClosing's
 * candidate matching) for `npairs` pairs of a keyframe set (single-camera keyframes). has_mp[i] != 0 here means "GetMapPointMatches()[i] is
 * a map point that is not bad". Inside every shared vocabulary node the keypoints of pKF1 are visited in order, each takes the best
 * keypoint of pKF2 that is not matched yet: bestDist1 < TH_LOW (strict), ratio nnratio = mfNNratio, then the rotation histogram.
 * match12_out[p * cap + idx1] = keypoint of pKF2 whose map point vpMatches12[idx1] receives, or -1; nmatches_out[p] = the return value. */
int orb_search_by_bow_kf(orb_handle* h, const orb_kf_set* kfs, const int32_t* kf1, const int32_t* kf2, int npairs, float nnratio,
                         int check_orientation, int32_t* match12_out, int32_t* nmatches_out, int flags);

/* ORBmatcher::SearchByProjection(KeyFrame *pKF, Sophus::Sim3f &Scw, vpPoints, vpMatched, th, ratioHamming) (src/ORBmatcher.cc:397-494)
 * and its overload with vpPointsKFs / vpMatchedKF (:496-601) - LoopClosing - for every frame of the resident batch = pKF (loaded with
 * orb_load_frames, grid built). One query per candidate map point: (u, v) = project(Tcw * p3Dw), octave = nPredictedLevel, flags bit 0 =
 * everything before the window holds (:420-447: not bad, not in spAlreadyFound, positive depth, IsInImage, distance range, viewing
 * angle); angle and z are not read. matched0[frame * kcap + idx] != 0: vpMatched[idx] != NULL when the call starts. Done here in
 * map-point order: window th * mvScaleFactors[level], levels [level - 1, level], best keypoint without a match (every match of the call
 * locks its keypoint), bestDist <= TH_LOW * ratio_hamming. match_out[frame * kcap + idx] = query whose map point vpMatched[idx]
 * receives (-1: unchanged), nmatches_out[frame] = the return value.
 * ORBmatcher::SearchBySim3 (:1323-1519) needs no entry point of its own: its two loops are orb_fuse_search(mode 1) on pKF2 with the map
 * points of pKF1 and on pKF1 with those of pKF2 (vnMatch = best_idx where best_dist <= TH_HIGH), followed by the agreement test on the
 * host (tests/test_gpu_map.py: test_search_by_sim3). */
int orb_search_by_projection_sim3(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                  const uint8_t* matched0, float th, float ratio_hamming, int32_t* match_out, int32_t* nmatches_out,
                                  int flags);

/* ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:603-700;
 * Tracking::MonocularInitialization, windowSize = 100) for every frame of the resident batch = F2 (grid built). One query per keypoint
 * i1 of the initial frame F1: (x, y) = vbPrevMatched[i1], angle and octave of F1.mvKeysUn[i1], qdesc = F1.mDescriptors. Done here in the
 * reference's order: only level-0 keypoints, GetFeaturesInArea(x, y, windowSize, 0, 0) on F2, candidates an earlier i1 holds at a
 * distance <= this one are skipped (vMatchedDistance, :638), best <= TH_LOW, bestDist < bestDist2 * nnratio in float, a better i1 takes
 * the keypoint over (:653-656), rotation histogram + ComputeThreeMaxima when check_orientation.
 * matches12_out[frame * qcap + i1] = vnMatches12[i1]; prev_matched_out[(frame * qcap + i1) * 2 ..] = vbPrevMatched[i1] after the update
 * of :694-697 (the matched keypoint's position, else unchanged); nmatches_out[frame] = the return value. */
typedef struct orb_init_query {
  float x, y;     /* vbPrevMatched[i1] */
  float angle;    /* F1.mvKeysUn[i1].angle */
  int32_t octave; /* F1.mvKeysUn[i1].octave */
} orb_init_query;
int orb_search_for_initialization(orb_handle* h, const orb_init_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                  int window_size, float nnratio, int check_orientation, int32_t* matches12_out, float* prev_matched_out,
                                  int32_t* nmatches_out, int flags);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:367-431; LocalMapping after every fusion / new observation) for
 * `npoints` map points at once: desc holds the observed descriptors of all map points back to back (the vDescriptors of :385-399
 * in observation order), off[p] .. off[p + 1] those of map point p. Per map point: all pairwise distances, the median of every row
 * = sorted row [0.5 * (N - 1)] (:421-423, the row includes the 0 of the diagonal), the first row with the smallest median wins.
 * best_out[p] = BestIdx relative to off[p] (-1 for a map point without descriptors, which the reference leaves untouched),
 * median_out[p] (optional) = BestMedian. ---- */
int orb_distinctive_descriptors(orb_handle* h, const uint8_t* desc, const int32_t* off, int npoints, int32_t* best_out, int32_t* median_out,
                                int flags);

/* ---- host-side formats: the fragments KeyFrame::serialize (include/KeyFrame.h:116-124) writes for the
 * front-end's results into the binary Atlas file (.osa; boost::archive::binary_oarchive, src/System.cc:1434, stores primitives and
 * make_array() blocks as their native bytes):
 *   serializeVectorKeyPoints(ar, mvKeys / mvKeysUn) (include/SerializationUtils.h:115-152):
 *       int32 NumEl, then per keypoint float angle, response, size, pt.x, pt.y, int32 class_id, octave          = 4 + 28 N bytes
 *   serializeMatrix(ar, mDescriptors) (:74-113): int32 cols, rows, type, bool continuous, rows * cols bytes    = 13 + 32 N bytes
 * orb_serialize_frame produces one fragment from the device-resident results of frame `frame` of the handle's last batch (the
 * keypoint records are permuted on the device, the fragment leaves the GPU as one copy); `out` is host memory of `cap` bytes,
 * `written` receives the fragment size. The loaders are the loading branch of the same templates (host only). ---- */
enum { ORB_SER_KEYS = 0, ORB_SER_KEYS_UN = 1, ORB_SER_DESCRIPTORS = 2 };
size_t orb_serialized_keypoints_size(int n);
size_t orb_serialized_descriptors_size(int n);
int orb_serialize_frame(orb_handle* h, int frame, int what, uint8_t* out, size_t cap, size_t* written);
int orb_deserialize_keypoints(const uint8_t* in, size_t len, orb_keypoint* kps, int cap, int* n_out);
int orb_deserialize_descriptors(const uint8_t* in, size_t len, uint8_t* desc, int cap_rows, int* rows_out);

/* ---- ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1880-1894): scalar host helper for the
 * 18 scalar call sites (no device work) ---- */
int orb_hamming_distance(const uint8_t* a, const uint8_t* b);

/* ---- pinned host / device memory helpers for callers that batch ---- */
int orb_host_alloc(void** p, size_t bytes);
/* write_combined != 0: cudaHostAllocWriteCombined - for INPUT buffers the CPU only writes (images on their way to the device):
 * the DMA reads bypass the CPU caches; reading such memory from the CPU is slow */
int orb_host_alloc_ex(void** p, size_t bytes, int write_combined);
int orb_host_free(void* p);
int orb_device_alloc(orb_handle* h, void** p, size_t bytes);
int orb_device_free(orb_handle* h, void* p);
int orb_memcpy_h2d(orb_handle* h, void* dst, const void* src, size_t bytes);
int orb_memcpy_d2h(orb_handle* h, void* dst, const void* src, size_t bytes);
/* the same on the handle's stream without the synchronisation (orb_sync completes them) */
int orb_memcpy_h2d_async(orb_handle* h, void* dst, const void* src, size_t bytes);
int orb_memcpy_d2h_async(orb_handle* h, void* dst, const void* src, size_t bytes);

/* ---- measurement support: CUDA-event timing on the handle's stream, kernel launch counter ---- */
int orb_timer_start(orb_handle* h);
int orb_timer_stop(orb_handle* h, float* ms_out); /* synchronises the stream */
int64_t orb_launch_count(const orb_handle* h);   /* kernels launched by this handle so far */
/* per-stage device time (ms) of the last orb_extract_batch when stage timing is enabled:
 * 0 pyramid, 1 blur, 2 fast, 3 octree, 4 assemble, 5 orient+describe, 6 stereo match, 7 stereo gate */
int orb_set_stage_timing(orb_handle* h, int enabled);
int orb_get_stage_times(orb_handle* h, float* ms8);

/* ---- stage outputs for parity tests (device -> host copies of intermediate results) ---- */
int orb_debug_get_blurred(orb_handle* h, int frame, int level, uint8_t* dst, size_t dst_stride);
/* FAST candidates of one level in reference order, (x, y, score) triples relative to the 16-px border */
int orb_debug_get_candidates(orb_handle* h, int frame, int level, int32_t* xys, int cap, int* n_out);
/* number of FAST candidates per (frame, level) of the last batch: counts[frame * nlevels + level] */
int orb_debug_get_level_counts(orb_handle* h, int32_t* counts, int cap);
/* keypoints of one level after the quad-tree, (x, y, score) triples relative to the border, list order */
int orb_debug_get_selected(orb_handle* h, int frame, int level, int32_t* xys, int cap, int* n_out);
/* the quad-tree kernel's emulation of libstdc++ std::sort(first, last, compareNodes) (src/ORBextractor.cc:665-667) on `n` records
 * compared by keys[i] only, payload = the record's input index: keys_out / payload_out receive the sorted order, which for equal
 * keys is the order the reference's (unstable) std::sort leaves them in. n <= 8192. */
int orb_debug_std_sort(orb_handle* h, const uint32_t* keys, int n, uint32_t* keys_out, uint32_t* payload_out);
/* run only the quad-tree stage on caller-supplied candidates (x,y,score; region w x h, target N) */
int orb_debug_distribute(orb_handle* h, const int32_t* cands, int n, int region_w, int region_h, int N,
                         int32_t* out, int cap, int* n_out);
/* best right index / Hamming distance per left keypoint of the last stereo match (before SAD) */
int orb_debug_get_stereo_best(orb_handle* hL, int frame, int32_t* best_idx, int32_t* best_dist, int cap);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H */
