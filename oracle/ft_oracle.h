// ft_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
//
// Plain C++17 restatement of the reference's CPU tracking front-end
// (sfu-rsl/FastTrack, the code executed when KernelController::*RunStatus == 0).
// It exists to CHECK the CUDA path. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it. The product
// (fasttrack_b200/) never links, imports or calls anything in oracle/.
//
// Parity pin: the reference ships no tests or golden vectors for this path and does not build here as a whole (it
// needs OpenCV C++, Eigen, Sophus, Pangolin). Its code is run anyway (oracle/Makefile, `make ref`, outputs in oracle/_ref/):
//   * src/ORBextractor.cc (CPU branch of operator()) compiled as a whole against an OpenCV stand-in whose image primitives
//     are this file's, each pinned bit-exactly against cv2 4.13 (tests/test_oracle_cv2_live.py, tests/golden/cv2_*.npz);
//   * the hot-path functions of src/Frame.cc, MapPoint.cc, ORBmatcher.cc and the camera models, compiled from their own
//     text (oracle/ref_extract_fns.py) against stand-in class definitions (oracle/ref_stubs/ref_frame_shim.h);
//   * the vendored DBoW2.
// This restatement equals that code bit for bit on every input of tests/test_oracle_ref_extractor.py,
// tests/test_oracle_ref_frame.py and tests/test_oracle_bow.py (golden fixtures tests/golden/ref_*.npz, dbow2_ref.npz).
// Eigen::JacobiSVD (KannalaBrandt8::Triangulate) has no independent pin: the compiled reference function and this file
// share one Jacobi routine. Still restated only: UndistortKeyPoints, remap / input resize (each pinned against cv2).
// See DESIGN.md section 2.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace fto {

struct KeyPoint {
  float x, y, size, angle, response;
  int octave;
};

struct Img {
  int w = 0, h = 0;
  std::vector<uint8_t> d;  // tight, stride == w
  const uint8_t* row(int y) const { return d.data() + (size_t)y * w; }
  uint8_t* row(int y) { return d.data() + (size_t)y * w; }
};

struct Candidate {  // pre-octree FAST keypoint, coordinates relative to (minBorderX, minBorderY)
  float x, y, response;
};

// ---- primitives (pinned against cv2) ----
int cv_round(float v);
void resize_linear_u8(const Img& src, Img& dst, int dw, int dh);    // cv::resize INTER_LINEAR 8UC1
void gaussian_blur_7x7_s2(const Img& src, Img& dst);                // cv::GaussianBlur 7x7 sigma 2 REFLECT_101
int fast_score_9_16(const uint8_t* p, int stride);                  // cv::FAST cornerScore<16> with threshold folded out
void fast_detect(const uint8_t* roi, int stride, int w, int h, int th, std::vector<Candidate>& out);  // cv::FAST(..., nonmax=true)
float fast_atan2(float y, float x);                                 // cv::fastAtan2
void knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx2, int* dist2);  // BFMatcher knnMatch k=2
int descriptor_distance(const uint8_t* a, const uint8_t* b);        // ORBmatcher.cc:2256-2273
// cv::remap(src, dst, map1 (CV_32FC1 x), map2 (CV_32FC1 y), INTER_LINEAR, BORDER_CONSTANT 0) on 8UC1 (System.cc:279-280)
void remap_linear_u8(const Img& src, const float* mapx, const float* mapy, int dw, int dh, Img& dst);

// cv::undistortPoints(src, dst, K, dist, Mat(), K) of Frame::UndistortKeyPoints (Frame.cc:771-804); K = fx fy cx cy
void undistort_points(const float* xy, int n, const float K[4], const float* dist, int ndist, float* out);
void image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float out[4]);  // Frame.cc:806-833

// Frame::ComputeStereoFromRGBD (Frame.cc:1065-1086); depth == nullptr = monocular frame
void stereo_from_rgbd(const float* keysXY, const float* keysUnX, int n, const float* depth, int w, int h, float mbf,
                      float* uRight, float* outDepth);

// ---- ORBextractor (ORBextractor.cc CPU branches) ----
class Extractor {
 public:
  Extractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  // operator() -- returns monoIndex, -1 on empty image (ORBextractor.cc:1356-1493)
  int extract(const uint8_t* img, int w, int h, int step, int lap0, int lap1,
              std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc);

  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> featuresPerLevel;
  std::vector<int> umax;
  // state after extract()
  std::vector<Img> pyramid;                          // mvImagePyramid (inner ROI)
  std::vector<Img> blurred;                          // per-level GaussianBlur (only levels with keypoints)
  std::vector<std::vector<Candidate>> candidates;    // vToDistributeKeys per level
  std::vector<std::vector<KeyPoint>> levelKeys;      // after octree + orientation, level coordinates
  std::vector<std::vector<uint8_t>> levelDesc;
  long descBorderline = 0;                           // samples within 1e-4 of a .5 rounding boundary

  void computePyramid(const uint8_t* img, int w, int h, int step);
  void computeKeyPointsOctTree();
  std::vector<Candidate> distributeOctTree(const std::vector<Candidate>& in, int minX, int maxX,
                                           int minY, int maxY, int N) const;
};

// test entry points: IC_Angle (ORBextractor.cc:39-66) / computeOrbDescriptor (:68-108) of given keypoints
void ic_angles(const Extractor& ex, const uint8_t* img, int w, int h, const float* xy, int n, float* out);
void orb_descriptors(const uint8_t* blurred, int w, int h, const float* xy, const float* angle, int n, uint8_t* out);

// ---- pinhole stereo (Frame.cc:835-1005) ----
struct StereoResult {
  std::vector<float> uRight, depth;
  std::vector<int> bestIdxR;   // coarse match index (-1 none) -- diagnostic
  std::vector<int> sad;        // best SAD per accepted match (-1 none)
};
void compute_stereo_matches(const Extractor& exL, const Extractor& exR,
                            const std::vector<KeyPoint>& kL, const std::vector<uint8_t>& dL,
                            const std::vector<KeyPoint>& kR, const std::vector<uint8_t>& dR,
                            float mbf, float mb, StereoResult& out);

// ---- camera models ----
struct Camera {
  int type;        // 0 pinhole, 1 KannalaBrandt8
  float p[8];      // fx fy cx cy k1..k4
};
void cam_project(const Camera& c, const float P[3], float uv[2]);          // Pinhole.cpp:43-49 / KannalaBrandt8.cpp:67-84
void kb8_unproject(const Camera& c, float u, float v, float ray[3]);       // KannalaBrandt8.cpp:116-143

// ---- fisheye stereo (Frame.cc:1231-1271 + KannalaBrandt8::TriangulateMatches) ----
struct FisheyeResult {
  std::vector<int> l2r, r2l;
  std::vector<float> depth;
  std::vector<float> p3d;       // Nleft x 3
  std::vector<int> code;        // per left kp: 0 no ratio match, 1 accepted, <0 TriangulateMatches reject code
  std::vector<int> knnIdx;      // per left-subset query: best train idx (diagnostic)
};
void compute_stereo_fisheye(const Camera& c1, const Camera& c2, const float Rlr[9], const float tlr[3],
                            const std::vector<float>& sigma2,
                            const std::vector<KeyPoint>& kL, const std::vector<uint8_t>& dL, int monoLeft,
                            const std::vector<KeyPoint>& kR, const std::vector<uint8_t>& dR, int monoRight,
                            FisheyeResult& out);

// ---- frame grid + frustum + SearchByProjection ----
struct FrameModel {
  int Nleft = -1, Nright = -1;   // -1: pinhole/rectified (ORB-SLAM3 convention)
  int N = 0;
  std::vector<KeyPoint> keys;        // N (left then right for fisheye)
  std::vector<uint8_t> desc;         // N x 32
  std::vector<float> uRight;         // N (pinhole) / Nleft
  std::vector<int> l2r, r2l;         // fisheye tables
  float minX, maxX, minY, maxY, gridWInv, gridHInv;
  std::vector<float> scale;
  float logScale; int nlevels;
  Camera cam1, cam2;
  float mbf;
  float Rcw[9], tcw[3], Ow[3], Rwc[9];
  float Rrl[9], trl[3], tlr[3];      // fisheye extrinsics
  std::vector<int> grid[64][48], gridR[64][48];
  void assignFeaturesToGrid();                                      // Frame.cc:409-440,749-759
  void featuresInArea(float x, float y, float r, int minLevel, int maxLevel, bool right,
                      std::vector<int>& out) const;                  // Frame.cc:681-747
};

struct MapPointIn {
  float pos[3], normal[3];
  float minDist, maxDist;      // raw mfMinDistance / mfMaxDistance
  uint8_t desc[32];
  int flags;                   // bit0 skip (bad / already matched in this frame), bit1 Observations()>0
};
struct MapPointTrack {         // the mTrack* scratch of MapPoint.h:170-181
  int inView = 0, inViewR = 0;
  float projX = -1, projY = -1, projXR = 0, depth = 0, viewCos = 0;
  float projXR_r = 0, projYR_r = 0, depthR = 0, viewCosR = 0;
  int level = -1, levelR = -1;
  int borderline = 0;          // PredictScale or a frustum compare within 1e-5 of its decision boundary
};
void is_in_frustum(const FrameModel& F, const MapPointIn& mp, float viewCosLimit, MapPointTrack& t);  // Frame.cc:536-598,1308-1382
// SearchByProjection #1 (ORBmatcher.cc:49-225). frameMP: per keypoint -1 none, else 2*id+obs>0 encoding is NOT used:
// holder[i] = map point index (>=0) or -1, and holderObs[i] = 1 when holder has Observations()>0.
// Pre-existing holders are passed with index -2 (foreign map point) and their obs flag.
int search_by_projection(FrameModel& F, const std::vector<MapPointIn>& mps, const std::vector<MapPointTrack>& tr,
                         float th, bool bFar, float thFar, float nnratio,
                         std::vector<int>& holder, std::vector<uint8_t>& holderObs);


// ---- frame-to-last-frame SearchByProjection (ORBmatcher.cc:1775-2085) + ComputeThreeMaxima (:2210-2254) ----
struct LastFramePoint {      // one keypoint of the last frame that holds a (non-outlier) MapPoint
  float pos[3];              // pMP->GetWorldPos()
  uint8_t desc[32];          // pMP->GetDescriptor()
  int octave;                // last-frame keypoint octave
  float angle;               // last-frame keypoint angle (degrees)
  int flags;                 // bit0 skip (no MapPoint / outlier), bit1 Observations()>0
};
// direction: +1 bForward, -1 bBackward, 0 neither. holder/holderObs as in search_by_projection. Returns nmatches.
int search_by_projection_last_frame(FrameModel& F, const std::vector<LastFramePoint>& pts, float th, int direction,
                                    bool checkOrientation, float mb, std::vector<int>& holder,
                                    std::vector<uint8_t>& holderObs, std::vector<int>& borderline);

// ---- bag of words: Frame::ComputeBoW (Frame.cc:762-769) -> DBoW2 transform, ORBmatcher::SearchByBoW(KF, F) ----
// The vocabulary is the reference's vendored DBoW2 (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, FORB.cpp,
// BowVector.cpp, FeatureVector.cpp). This restatement is PINNED against that code itself: oracle/Makefile compiles the
// reference's DBoW2 sources where they lie into oracle/_ref/libft_ref_dbow2.so (tests/test_oracle_bow.py, fixtures in
// tests/golden/dbow2_ref.npz made by tools/make_dbow2_golden.py).
struct Vocabulary {
  int k = 0, L = 0, scoring = 0, weighting = 0;   // ScoringType / WeightingType enums of BowVector.h:39-56
  std::vector<int> parent;                        // [n] node 0 = root
  std::vector<std::vector<int>> children;         // in order of appearance (m_nodes[pid].children.push_back)
  std::vector<uint8_t> desc;                      // [n][32]
  std::vector<double> weight;                     // [n]
  std::vector<int> wordId;                        // [n] (0 for inner nodes, as Node() initialises it)
  int nWords = 0;
  // loadFromTextFile (TemplatedVocabulary.h:1338-1423), including the node its `while(!f.eof())` loop makes out of
  // the empty line after a trailing newline: weight 0, parent and leaf flag = the previous line's (uninitialised
  // locals in the reference), descriptor bytes unspecified in the reference (cv::Mat::create does not initialise) --
  // zero here, as in the stand-in Mat the _ref build uses.
  bool loadText(const char* path);
  // the same tree from arrays: node i+1 has parent[i], is_leaf[i], desc[i], weight[i] (one text line each)
  void fromArrays(int k_, int L_, int scoring_, int weighting_, int n, const int* parent_, const uint8_t* isLeaf,
                  const uint8_t* desc_, const double* weight_);
  // transform(feature, word_id, weight, nid, levelsup) (:1218-1260)
  void transformOne(const uint8_t* f, int levelsup, unsigned& word, double& w, unsigned& nid) const;
};
// transform(features, BowVector, FeatureVector, levelsup) (:1127-1194). featNode[i] = node of feature i in the
// FeatureVector, -1 when its word is stopped (weight 0); bow = (ids ascending, values) after normalisation.
void voc_transform(const Vocabulary& v, const uint8_t* desc, int n, int levelsup, std::vector<int>& featNode,
                   std::vector<int>& featWord, std::vector<unsigned>& bowIds, std::vector<double>& bowVals);
// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:322-523). Indices are into the
// concatenated (left, right) keypoint arrays; nodes are FeatureVector nodes per feature (-1 = not in the FeatureVector).
// kfHasMp[i] = vpMapPointsKF[i] && !isBad(). match[iF] = KeyFrame feature whose MapPoint the frame keypoint received
// (-1 none). Returns nmatches.
int search_by_bow(int nKF, const uint8_t* kfDesc, const float* kfAngle, const int* kfNode, const uint8_t* kfHasMp,
                  int nF, const uint8_t* fDesc, const float* fAngle, const int* fNode, int fNleft, float nnratio,
                  bool checkOrientation, std::vector<int>& match);

}  // namespace fto
