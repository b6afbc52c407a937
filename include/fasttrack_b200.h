/* fasttrack_b200.h -- C ABI of the B200-native stereo tracking front-end.
 *
 * Drop-in boundary for the hot path of sfu-rsl/FastTrack (ORB-SLAM3 + CUDA): ORB
 * extraction (L+R), stereo matching and SearchByProjection of local MapPoints.
 * Plain C: opaque handle, POD structs, raw pointers and sizes. Every entry point
 * returns an ft_status (0 = ok) and never exits/aborts; ft_last_error() gives the text.
 * One context per (GPU, camera rig, sequence); calls on one context are single-threaded,
 * contexts on different GPUs are independent (no NCCL, no peer access).
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   ft_context_create / destroy   <- KernelController::initializeKernels / shutdownKernels
 *                                    (include/Kernels/KernelController.h:27-29),
 *                                    CudaUtils::loadSetting (src/Kernels/CudaUtils.cu:24-40),
 *                                    ORBextractor ctor/dtor device setup (src/ORBextractor.cc:416-442,1546-1564)
 *   ft_extract_stereo             <- ORBextractor::operator() x2 as driven by Frame::ExtractORB on two
 *                                    threads (include/ORBextractor.h:113-115, src/Frame.cc:127-130,442-449)
 *   ft_stereo_match               <- Frame::ComputeStereoMatches / KernelController::launchStereoMatchKernel
 *                                    (src/Frame.cc:835-1005, KernelController.h:33-39)
 *   ft_stereo_match_fisheye       <- Frame::ComputeStereoFishEyeMatches / launchFisheyeStereoMatchKernel
 *                                    (src/Frame.cc:1231-1271, KernelController.h:41)
 *   ft_frame_download             <- the host vectors those operators fill (mvKeys, mDescriptors, mvuRight,
 *                                    mvDepth, mvLeftToRightMatch, mvRightToLeftMatch, mvStereo3Dpoints)
 *   ft_set_pose                   <- Frame::SetPose / UpdatePoseMatrices (src/Frame.cc:455-506)
 *   ft_search_local_points        <- Frame::isInFrustum loop of Tracking::SearchLocalPoints
 *                                    (src/Tracking.cc:3504-3522) + ORBmatcher::SearchByProjection #1
 *                                    (src/ORBmatcher.cc:49-312) / launchSearchLocalPointsKernel
 *                                    (KernelController.h:43-45)
 */
#ifndef FASTTRACK_B200_H
#define FASTTRACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FT_MAX_LEVELS 16
#define FT_GRID_COLS 64 /* FRAME_GRID_COLS, include/Frame.h:47 */
#define FT_GRID_ROWS 48 /* FRAME_GRID_ROWS, include/Frame.h:46 */

typedef enum {
  FT_OK = 0,
  FT_ERR_INVALID = 1,   /* bad argument / unsupported configuration */
  FT_ERR_CUDA = 2,      /* CUDA runtime error (text in ft_last_error) */
  FT_ERR_CAPACITY = 3,  /* a device-side buffer bound was exceeded; nothing was truncated silently */
  FT_ERR_STATE = 4      /* call out of order (e.g. stereo match before extract) */
} ft_status;

typedef enum { FT_CAM_PINHOLE = 0, FT_CAM_KB8 = 1 } ft_camera_type;
typedef enum { FT_SENSOR_STEREO = 0, FT_SENSOR_MONOCULAR = 1, FT_SENSOR_RGBD = 2 } ft_sensor;   /* System::eSensor subset */

/* Settings the reference reads from YAML (src/Settings.cc) plus the rig geometry.
 * Limits checked by ft_context_create (FT_ERR_INVALID / FT_ERR_CAPACITY with the numbers in ft_last_error): image at most
 * 4000 x 4000; every pyramid level at least one 35-px FAST cell; aspect ratio at least 1:2; about 6500 features at scale
 * factor 1.2 (the per-level octree keeps its node list in shared memory). */
typedef struct {
  int device_id;
  int width, height;        /* Camera.width / Camera.height */
  int nfeatures;            /* ORBextractor.nFeatures */
  int nlevels;              /* ORBextractor.nLevels (<= FT_MAX_LEVELS) */
  float scale_factor;       /* ORBextractor.scaleFactor */
  int ini_th_fast;          /* ORBextractor.iniThFAST */
  int min_th_fast;          /* ORBextractor.minThFAST */
  int camera_type;          /* ft_camera_type */
  float cam1[8];            /* fx fy cx cy [k1 k2 k3 k4] (KB8) */
  float cam2[8];
  int lap_left[2];          /* KannalaBrandt8::mvLappingArea of camera 1 ({0,0} for pinhole) */
  int lap_right[2];
  float bf;                 /* mbf = baseline * fx */
  float Tlr[12];            /* 3x4 row-major T_c1_c2 (fisheye rigs); ignored for pinhole */
  int max_map_points;       /* capacity of one ft_search_local_points call (reference: 25000) */
} ft_config;

/* cv::KeyPoint without class_id (always -1 in the reference): 24 bytes. */
typedef struct {
  float x, y;      /* pt, level-0 coordinates */
  float size;      /* float(int(31*scale[octave])) */
  float angle;     /* degrees [0,360) */
  float response;  /* FAST score */
  int octave;
} ft_keypoint;

typedef struct ft_context ft_context;

const char* ft_last_error(void);
const char* ft_version(void);

ft_status ft_context_create(const ft_config* cfg, ft_context** out);
ft_status ft_context_destroy(ft_context* ctx);

/* Scale tables exactly as ORBextractor exposes them (GetScaleFactors etc., ORBextractor.h:117-146). */
ft_status ft_get_scale_tables(ft_context* ctx, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                              int* features_per_level);

/* Extract both eyes. imgL/imgR: HOST pointers to 8-bit grayscale, row pitch stepL/stepR bytes.
 * Uploads, builds both pyramids, FAST + octree + orientation + rBRIEF; results stay on the device.
 * Asynchronous w.r.t. the host: returns once the work is enqueued. */
ft_status ft_extract_stereo(ft_context* ctx, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR);

/* Same, for images already resident in device memory (tight pitch == width not required). */
ft_status ft_extract_stereo_device(ft_context* ctx, const uint8_t* d_imgL, int stepL, const uint8_t* d_imgR, int stepR);

/* Frame::ComputeStereoMatches on the device-resident frame (pinhole / rectified rigs). */
ft_status ft_stereo_match(ft_context* ctx);
/* Frame::ComputeStereoFishEyeMatches (KannalaBrandt8 rigs). */
ft_status ft_stereo_match_fisheye(ft_context* ctx);

/* Number of keypoints per eye and monoIndex (operator()'s return value). Synchronises. */
ft_status ft_frame_counts(ft_context* ctx, int* n_left, int* n_right, int* mono_left, int* mono_right);

/* Copy one eye's result to host buffers (any pointer may be NULL). kps/desc capacity = cap entries.
 * u_right/depth (left eye only, length n_left), l2r (n_left) / r2l (n_right) / p3d (n_left*3) for fisheye. */
ft_status ft_frame_download(ft_context* ctx, int eye, int cap, ft_keypoint* kps, uint8_t* desc, int* n, int* mono_index,
                            float* u_right, float* depth, int* l2r, int* r2l, float* p3d);

/* Stereo rectification in front of the extractor: cv::remap(im, M1, M2, INTER_LINEAR) of System::TrackStereo
 * (reference src/System.cc:273-281) fused into the level-0 load. M1x/M2x: the CV_32F x / y maps of
 * cv::initUndistortRectifyMap (src/Settings.cc:506-509), width*height floats each (the context's camera size); raw
 * images handed to ft_extract_stereo / ft_frame_construct are then raw_width x raw_height. NULL maps switch it off. */
ft_status ft_set_rectification(ft_context* ctx, int raw_width, int raw_height, const float* M1l, const float* M2l,
                               const float* M1r, const float* M2r);

/* Monocular / RGB-D sensors (RGB-D half). The reference's monocular and RGB-D Frame constructors
 * (src/Frame.cc:226-325, 328-419) run one ORBextractor::operator(), UndistortKeyPoints, then either
 * ComputeStereoFromRGBD (src/Frame.cc:1065-1086: depth = imDepth.at<float>(kp.pt.y, kp.pt.x), mvuRight = kpU.pt.x -
 * mbf / depth where depth > 0) or leave mvuRight / mvDepth at -1. ft_set_sensor switches a pinhole context between
 * the three modes (default FT_SENSOR_STEREO); ft_extract_mono extracts eye 0 only; ft_depth_from_rgbd takes the CV_32F
 * depth image (step in bytes; NULL on a monocular context) and also builds the frame grid, after which
 * ft_frame_download(eye 0), ft_frame_keypoints_undistorted and every projection search work as for a stereo frame. */
ft_status ft_set_sensor(ft_context* ctx, int sensor);
ft_status ft_extract_mono(ft_context* ctx, const uint8_t* img, int step);
ft_status ft_depth_from_rgbd(ft_context* ctx, const float* depth, int step_bytes);

/* cv::resize(im, imToFeed, newImSize) in front of the extractor (reference src/System.cc:282-285, taken when
 * Settings::needToResize): raw images are raw_width x raw_height, level 0 of the pyramid is their INTER_LINEAR resize
 * to the context's camera size. raw_width = 0 switches it off; FT_ERR_STATE while rectification is active (the
 * reference does one or the other). */
ft_status ft_set_input_resize(ft_context* ctx, int raw_width, int raw_height);

/* Frame::UndistortKeyPoints + Frame::ComputeImageBounds (reference src/Frame.cc:771-835) for a pinhole camera with
 * distortion: dist_coef = mDistCoef (k1 k2 p1 p2 [k3]; n = 4 or 5, n = 0 or k1 == 0 switches it off as Frame.cc:773
 * does). cv::undistortPoints(pt, K, mDistCoef, Mat(), K) runs per left keypoint inside the frame-grid kernel; the
 * grid (AssignFeaturesToGrid), the image bounds of isInFrustum and both projection searches then use mvKeysUn, as
 * the reference does; ComputeStereoMatches keeps mvKeys. ft_image_bounds returns {mnMinX, mnMaxX, mnMinY, mnMaxY};
 * ft_frame_keypoints_undistorted returns mvKeysUn[i].pt (x, y interleaved, 2*cap floats). */
ft_status ft_set_distortion(ft_context* ctx, const float* dist_coef, int n);
ft_status ft_image_bounds(ft_context* ctx, float* out4);
ft_status ft_frame_keypoints_undistorted(ft_context* ctx, int cap, float* xy, int* n);

/* Capacity (entries) of the per-eye keypoint arrays: nfeatures + a few per level (the octree may exceed a level
 * quota by up to 3, reference sizes its buffers nfeatures+20, include/Kernels/CudaUtils.h:16). */
int ft_max_keypoints(ft_context* ctx);

/* The Frame constructor's front-end work in one call and one synchronisation (reference src/Frame.cc:102-223,
 * 1115-1229): upload, extract both eyes, stereo-match (pinhole or fisheye according to the context), download.
 * Every output array must hold ft_max_keypoints() entries (p3d: 3x); counts4 = {n_left, mono_left, n_right,
 * mono_right}. Unused outputs may be NULL. */
ft_status ft_frame_construct(ft_context* ctx, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR,
                             ft_keypoint* kpsL, uint8_t* descL, ft_keypoint* kpsR, uint8_t* descR, int* counts4,
                             float* u_right, float* depth, int* l2r, int* r2l, float* p3d);

/* ft_frame_construct in two halves, so that a caller keeps two frames in flight on two contexts: ORB extraction and
 * stereo matching of frame t+1 do not depend on the SLAM state and overlap the tracking of frame t (the reference
 * runs them strictly in sequence inside the Frame constructor, src/Frame.cc:102-223). ft_frame_submit enqueues
 * upload + extract + stereo + the download of the result slab and returns at once; the images must stay valid until
 * ft_frame_collect, which waits for the frame and fills the host vectors exactly as ft_frame_construct does.
 * FT_ERR_STATE when no frame is pending. */
ft_status ft_frame_submit(ft_context* ctx, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR);
ft_status ft_frame_collect(ft_context* ctx, ft_keypoint* kpsL, uint8_t* descL, ft_keypoint* kpsR, uint8_t* descR,
                           int* counts4, float* u_right, float* depth, int* l2r, int* r2l, float* p3d);

/* ft_frame_construct without the downloads, for images already in device memory; asynchronous. */
ft_status ft_frame_enqueue_device(ft_context* ctx, const uint8_t* d_imgL, int stepL, const uint8_t* d_imgR, int stepR);

/* Pose of the current frame: Rcw (row-major 3x3), tcw; Rwc/Ow may be NULL (then Rwc = Rcw^T, Ow = -Rwc*tcw). */
ft_status ft_set_pose(ft_context* ctx, const float* Rcw, const float* tcw, const float* Rwc, const float* Ow);

/* Tracking::SearchLocalPoints, loop 2 + SearchByProjection. All pointers are HOST memory.
 *   pos/normal [M][3], minmax [M][2] = raw {mfMinDistance, mfMaxDistance}, desc [M][32],
 *   flags [M]: bit0 = skip (isBad() or mnLastFrameSeen == frame id), bit1 = Observations() > 0.
 *   holder [N] in/out  : F.mvpMapPoints as indices: -1 none, -2 a map point that is not in this call,
 *                        >= 0 index into this call's arrays.  N = n_left (pinhole) or n_left+n_right (fisheye).
 *   holder_obs [N] in/out: 1 when the holder has Observations() > 0.
 *   best_idx [M][2] out (may be NULL): keypoint chosen by the left / right search of each map point, -1 none.
 *   nmatches out: the function's return value in the reference. */
ft_status ft_search_local_points(ft_context* ctx, int M, const float* pos, const float* normal, const float* minmax,
                                 const uint8_t* desc, const int* flags, float th, int b_far_points,
                                 float th_far_points, float nnratio, int* holder, uint8_t* holder_obs, int* best_idx,
                                 int* nmatches);

/* The same search in three separable steps, so that the map-point snapshot can stay resident on the device
 * between frames (the reference re-marshals all map points on every call, SearchLocalPointsKernel.cu:368-409):
 *   ft_upload_map_points  H2D of the snapshot (asynchronous),
 *   ft_upload_holders     H2D of F.mvpMapPoints as indices (NULL pointers = all empty, as after the Frame ctor),
 *   ft_search_resident    frustum + candidate gathering + claim resolution, asynchronous, results stay on the device,
 *   ft_search_download    D2H of the results; synchronises. */
ft_status ft_upload_map_points(ft_context* ctx, int M, const float* pos, const float* normal, const float* minmax,
                               const uint8_t* desc, const int* flags);
ft_status ft_upload_holders(ft_context* ctx, int N, const int* holder, const uint8_t* holder_obs);
/* like ft_upload_map_points, but the SoA arrays already live in DEVICE memory and are used in place (no copy) */
ft_status ft_bind_map_points_device(ft_context* ctx, int M, const float* d_pos, const float* d_normal,
                                    const float* d_minmax, const uint8_t* d_desc, const int* d_flags);
ft_status ft_search_resident(ft_context* ctx, float th, int b_far_points, float th_far_points, float nnratio);
ft_status ft_search_download(ft_context* ctx, int* holder, uint8_t* holder_obs, int* best_idx, int* nmatches);

/* Zero-copy variant for integrations that marshal MapPoints themselves (the reference's CudaMapPoint loop,
 * src/Kernels/CudaWrappers/CudaMapPoint.cc:15-34): ft_map_point_staging returns HOST pointers into the context's
 * pinned staging buffer, laid out for M map points (pos[M][3], normal[M][3], minmax[M][2], desc[M][32], flags[M],
 * holder[N], holder_obs[N]); the caller fills them and calls ft_search_staged, which issues ONE H2D, the kernels and
 * ONE D2H. The returned result pointers point into pinned host memory owned by the context and stay valid until
 * the next search. */
ft_status ft_map_point_staging(ft_context* ctx, int M, float** pos, float** normal, float** minmax, uint8_t** desc,
                               int** flags, int** holder, uint8_t** holder_obs);
ft_status ft_search_staged(ft_context* ctx, int M, float th, int b_far_points, float th_far_points, float nnratio,
                           const int** holder_out, const uint8_t** holder_obs_out, const int** best_idx_out,
                           int* nmatches);

/* ---- Persistent device-side MapPoint store ----
 * The reference re-marshals every local MapPoint into a CudaMapPoint on every frame (an `omp parallel for` over
 * mutex-guarded getters, src/Kernels/CudaWrappers/CudaMapPoint.cc:15-34, src/Kernels/SearchLocalPointsKernel.cu:368-409,
 * called from Tracking::SearchLocalPoints, src/Tracking.cc:3595-3632) and uploads all of them. Here a MapPoint owns a
 * row of the store for its lifetime; the mapping side upserts the rows it created or changed (MapPoint::SetWorldPos,
 * UpdateNormalAndDepth, ComputeDistinctiveDescriptors: src/MapPoint.cc:83-96,383-473,475-529), and a frame's search
 * names its local map as a list of rows (the order of mvpLocalMapPoints, which the claim resolution observes) plus
 * the per-call flags. Results are identical to ft_search_local_points on the gathered arrays.
 *
 * ft_map_store_create: `capacity` rows owned by this context. ft_map_store_attach: another context of the SAME
 * sequence and device (e.g. the second context of a two-frames-in-flight pipeline) searches the same rows; updates
 * and searches are ordered across the contexts' streams inside the library. The store is freed with its last context.
 * ft_map_store_update: upsert n rows (slots[i] in [0, capacity)); arrays as in ft_search_local_points.
 * ft_search_store: as ft_search_local_points with (slots[M], flags[M]) in place of the five arrays.
 * ft_search_store_submit / ft_search_collect: the two halves of ft_search_store. The reference's launch is synchronous
 * (SearchLocalPointsKernel::launch copies in, runs the kernel and copies out before it returns,
 * src/Kernels/SearchLocalPointsKernel.cu:346-478); here the submit enqueues the H2D of holders + rows + flags, the two
 * kernels and the D2H of the result and returns, and the collect waits for the result and fills holder / holder_obs /
 * best_idx (best_idx only when want_best_idx was set) and nmatches. Between the two the tracking thread is free -- e.g.
 * to hand the next camera frame to ft_frame_submit on another context. holder / holder_obs of the submit are read before
 * it returns. One search per context may be outstanding: every other search call on the context returns FT_ERR_STATE
 * until it has been collected, and the staging arrays of ft_map_point_staging must not be written in between (the submit's
 * upload reads the same pinned buffer). With M == 0 or a frame without keypoints nothing is enqueued and the collect returns
 * FT_OK with nmatches = 0 and the output arrays untouched (the holders handed to the submit are the result).
 * Errors: FT_ERR_STATE without a store / frame, FT_ERR_CAPACITY for a row outside the store or M > max_map_points. */
ft_status ft_map_store_create(ft_context* ctx, int capacity);
ft_status ft_map_store_attach(ft_context* ctx, ft_context* owner);
ft_status ft_map_store_update(ft_context* ctx, int n, const int* slots, const float* pos, const float* normal,
                              const float* minmax, const uint8_t* desc);
ft_status ft_search_store(ft_context* ctx, int M, const int* slots, const int* flags, float th, int bFarPoints,
                          float thFarPoints, float nnratio, int* holder, uint8_t* holder_obs, int* best_idx,
                          int* nmatches);
ft_status ft_search_store_submit(ft_context* ctx, int M, const int* slots, const int* flags, float th, int bFarPoints,
                                 float thFarPoints, float nnratio, const int* holder, const uint8_t* holder_obs,
                                 int want_best_idx);
ft_status ft_search_collect(ft_context* ctx, int* holder, uint8_t* holder_obs, int* best_idx, int* nmatches);

/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) -- the frame-to-last-frame search of
 * Tracking::TrackWithMotionModel (reference src/ORBmatcher.cc:1775-2085, src/Tracking.cc:2911-2990), incl. the
 * rotation-histogram consistency check (ComputeThreeMaxima, :2210-2254) when check_orientation != 0.
 * One entry per last-frame keypoint that holds a map point: pos[n][3] = GetWorldPos(), desc[n][32] = GetDescriptor(),
 * octave[n] / angle[n] of that last-frame keypoint, flags[n]: bit0 = skip (no map point / mvbOutlier), bit1 =
 * Observations() > 0. Rlw/tlw = LastFrame pose (row-major 3x3, 3); the current pose comes from ft_set_pose.
 * holder / holder_obs / best_idx / nmatches as in ft_search_local_points (the caller clears F.mvpMapPoints first, as
 * TrackWithMotionModel does). */
ft_status ft_search_last_frame(ft_context* ctx, int n, const float* pos, const uint8_t* desc, const int* octave,
                               const float* angle, const int* flags, const float* Rlw, const float* tlw, float th,
                               int b_mono, int check_orientation, int* holder, uint8_t* holder_obs, int* best_idx,
                               int* nmatches);

/* ---- Bag of words: Frame::ComputeBoW and ORBmatcher::SearchByBoW(KeyFrame*, Frame&) ----
 * ft_vocabulary is ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (reference include/ORBVocabulary.h:28-29,
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) resident on one device, shared read-only by the contexts of that device.
 *   ft_vocabulary_load_text  <- ORBVocabulary::loadFromTextFile (TemplatedVocabulary.h:1338-1423; System.cc:118-127 loads
 *                               ORBvoc.txt with it), incl. the extra stopped node its eof loop makes after a trailing newline
 *   ft_vocabulary_create     the same tree from arrays: node i+1 (node 0 is the root) has parent[i] (must precede it),
 *                               is_leaf[i] (the file's flag: leaves get word ids in order), desc[i][32], weight[i]
 *   ft_vocabulary_transform  <- ORBVocabulary::transform(features, BowVector, FeatureVector, levelsup) (:1127-1194) for HOST
 *                               descriptors, e.g. KeyFrame::ComputeBoW (src/KeyFrame.cc:98-108). node_id[i] = the FeatureVector
 *                               node of feature i, -1 when its word is stopped (weight 0); (bow_ids ascending, bow_vals) =
 *                               the BowVector after the weighting / normalisation the vocabulary's types prescribe.
 *                               Thread-safe per vocabulary. At most 16384 features per call (FT_ERR_CAPACITY).
 *   ft_compute_bow           <- Frame::ComputeBoW (src/Frame.cc:762-769) on the context's device-resident descriptors
 *                               (left eye; left then right for fisheye rigs, the vconcat of Frame.cc:1218); asynchronous.
 *   ft_bow_download          mBowVec / mFeatVec of that frame: arrays as ft_vocabulary_transform (cap entries each).
 *   ft_search_by_bow         <- ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (src/ORBmatcher.cc:322-523,
 *                               called by Tracking::TrackReferenceKeyFrame / Relocalization, src/Tracking.cc:2787,3846) after
 *                               ft_compute_bow. KeyFrame side, HOST arrays over its (left, right) keypoints: kf_desc[n][32],
 *                               kf_angle[n] (mvKeysUn / mvKeys / mvKeysRight angle), kf_node[n] (its FeatureVector node per
 *                               feature, -1 none: node_id of ft_vocabulary_transform / ft_bow_download), kf_has_mp[n] =
 *                               vpMapPointsKF[i] && !isBad(). match[N] out: the KeyFrame feature whose MapPoint frame keypoint
 *                               i receives (vpMapPointMatches[i] = vpMapPointsKF[match[i]]), -1 none; N = n_left (+ n_right).
 *                               nnratio = mfNNratio, check_orientation = mbCheckOrientation. *nmatches = the return value. */
typedef struct ft_vocabulary ft_vocabulary;
ft_status ft_vocabulary_load_text(int device_id, const char* path, ft_vocabulary** out);
ft_status ft_vocabulary_create(int device_id, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                               const uint8_t* is_leaf, const uint8_t* desc, const double* weight, ft_vocabulary** out);
ft_status ft_vocabulary_destroy(ft_vocabulary* voc);
ft_status ft_vocabulary_info(ft_vocabulary* voc, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words);
ft_status ft_vocabulary_transform(ft_vocabulary* voc, const uint8_t* desc, int n, int levelsup, int* word_id, int* node_id,
                                  uint32_t* bow_ids, double* bow_vals, int bow_cap, int* n_bow);
ft_status ft_compute_bow(ft_context* ctx, ft_vocabulary* voc, int levelsup);
ft_status ft_bow_download(ft_context* ctx, int cap, int* word_id, int* node_id, uint32_t* bow_ids, double* bow_vals, int* n_bow,
                          int* n);
ft_status ft_search_by_bow(ft_context* ctx, int n_kf, const uint8_t* kf_desc, const float* kf_angle, const int* kf_node,
                           const uint8_t* kf_has_mp, float nnratio, int check_orientation, int* match, int* nmatches);

/* Block until everything enqueued on this context has finished. */
ft_status ft_synchronize(ft_context* ctx);

/* Host image buffers. The extractor takes any host pointer; pageable memory (a plain cv::Mat) is first copied row by row
 * into the context's pinned staging buffer, which costs one extra pass over both images per frame. A caller that owns its
 * frame buffers avoids that copy by allocating them here (page-locked, the DMA reads them directly) or by registering the
 * buffers it already has once (camera ring buffers are reused frame after frame). The reference has no counterpart: its
 * ORBextractor::ComputePyramidGPU issues cudaMemcpyAsync on the cv::Mat as it comes (src/ORBextractor.cc:1524). */
ft_status ft_host_alloc(size_t bytes, void** out);
ft_status ft_host_free(void* p);
ft_status ft_host_register(void* p, size_t bytes);
ft_status ft_host_unregister(void* p);

/* Test hook, host arithmetic only: sinf / cosf as the descriptor kernel evaluates them (glibc's polynomial in double;
 * the reference calls cos/sin on a float, src/ORBextractor.cc:74). tests/test_abi.py checks it against the host libm. */
void ft_debug_sincosf(int n, const float* angle, float* sin_out, float* cos_out);

/* ---- diagnostics used by the parity tests (device -> host copies of intermediate stages) ---- */
ft_status ft_debug_level_dims(ft_context* ctx, int level, int* w, int* h);
ft_status ft_debug_level_image(ft_context* ctx, int eye, int level, int blurred, uint8_t* out /* w*h tight */);
/* pre-octree FAST candidates of one level in canonical order: xyr[n][3], returns count in *n */
ft_status ft_debug_level_candidates(ft_context* ctx, int eye, int level, int cap, float* xyr, int* n);
/* the device introsort used by the octree, on its own: sorts n <= 4096 words by their high 32 bits exactly as
 * libstdc++ std::sort would arrange them (test hook) */
int ft_debug_sort(unsigned long long* keys_inout, int n);
/* per-level FAST candidate and kept-keypoint counts of one eye (arrays of nlevels) */
ft_status ft_debug_level_counts(ft_context* ctx, int eye, int* cand, int* kp);
/* frustum scratch of the last ft_search_local_points: track_i[M][4] = inView,inViewR,level,levelR;
 * track_f[M][9] = projX,projY,projXR,depth,viewCos,projXR_r,projYR_r,depthR,viewCosR */
ft_status ft_debug_track(ft_context* ctx, int M, int* track_i, float* track_f);
/* frame grid: counts[64*48] (ix*48+iy) and the concatenated keypoint indices */
ft_status ft_debug_grid(ft_context* ctx, int right, int* counts, int* indices, int* n);
/* measured counts of the last frame for the roofline arithmetic (bench.py):
 * stats[0..] = C_left, C_right (FAST candidates), K_left, K_right, stereo candidates tested, coarse matches
 * refined, projection-search candidates, resolve rounds */
ft_status ft_debug_stats(ft_context* ctx, long long* stats, int n);

/* Per-stage device timing with CUDA events recorded around every kernel on the stream it is launched on.
 * Enabling it switches the context to direct launches (events cannot be read out of a replayed graph).
 * enable = 1: the launch topology of the product path (branches on several streams: stage times include the contention
 * between concurrent kernels); enable = 2: every launch on one stream, so each stage is timed in isolation. */
#define FT_STAGE_COPY0 0
#define FT_STAGE_RESIZE 1
#define FT_STAGE_BLUR 2
#define FT_STAGE_FAST 3
#define FT_STAGE_OCTREE 4
#define FT_STAGE_ORIENT 5
#define FT_STAGE_GRID 6
#define FT_STAGE_STEREO 7
#define FT_STAGE_OUTLIER 8
#define FT_STAGE_FRUSTUM 9
#define FT_STAGE_GATHER 10
#define FT_STAGE_RESOLVE 11
#define FT_STAGE_FAST_L0 12   /* level 0 has its own branch of the launch graph */
#define FT_STAGE_OCTREE_L0 13
#define FT_STAGE_BLUR_L0 14
#define FT_STAGE_COUNT 15
ft_status ft_set_stage_timing(ft_context* ctx, int enable);
ft_status ft_get_stage_times(ft_context* ctx, float* ms /* [FT_STAGE_COUNT], -1 = not run */, int n);

/* The CUDA stream the context enqueues on (cudaStream_t as void*), for event timing by the caller. */
void* ft_context_stream(ft_context* ctx);
/* Enable/disable replay of the captured CUDA graph for the per-frame chain (default on). */
ft_status ft_set_use_graph(ft_context* ctx, int enable);
/* Number of kernel launches the last ft_extract_stereo + ft_stereo_match + ft_search_local_points issued. */
ft_status ft_launch_counts(ft_context* ctx, int* extract, int* stereo, int* search);

#ifdef __cplusplus
}
#endif
#endif /* FASTTRACK_B200_H */
