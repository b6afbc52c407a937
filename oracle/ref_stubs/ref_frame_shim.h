// ref_frame_shim.h -- stand-in class definitions for compiling FUNCTIONS of the reference's Frame.cc / ORBmatcher.cc /
// MapPoint.cc / camera models, TEST INFRASTRUCTURE ONLY.
//
// Those files as a whole need Eigen, Sophus, g2o and OpenCV C++, none of which is installed. The functions on the hot
// path use very little of them, so oracle/ref_extract_fns.py copies the TEXT of each function, verbatim and at build
// time, from the reference file where it lies into oracle/_ref/gen_frame_fns.inc (git-ignored; never stored in the
// repository), and ref_frame_capi.cpp compiles that text against the classes below, which declare only the members
// those functions touch, with the reference's names and types (include/Frame.h, MapPoint.h, KeyFrame.h, ORBmatcher.h).
// Eigen / Sophus are replaced by the few fixed-size operations used (unfused, evaluated left to right).
#pragma once
#include <climits>
#include <cmath>
#include <list>
#include <mutex>
#include <set>
#include <vector>

#include <opencv2/opencv.hpp>          // the stand-in of this directory
#include "DBoW2/BowVector.h"           // the reference's own (Thirdparty/DBoW2)
#include "DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48             // reference include/Frame.h:46-47
#define FRAME_GRID_COLS 64

namespace cv {
enum { NORM_L1 = 2 };
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {   // 8U: integer sum of absolute differences
  long s = 0;
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) s += std::abs((int)a.ptr(y)[x] - (int)b.ptr(y)[x]);
  return (double)s;
}
}  // namespace cv

namespace Eigen {
template <typename T, int R, int C> struct Matrix;
template <> struct Matrix<float, 3, 1> {
  float v[3];
  Matrix() : v{0, 0, 0} {}
  Matrix(float a, float b, float c) : v{a, b, c} {}
  float& operator()(int i) { return v[i]; }
  const float& operator()(int i) const { return v[i]; }
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
  float dot(const Matrix& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
  float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
  Matrix operator+(const Matrix& o) const { return Matrix(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Matrix operator-(const Matrix& o) const { return Matrix(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
};
template <> struct Matrix<float, 2, 1> {
  float v[2];
  Matrix() : v{0, 0} {}
  Matrix(float a, float b) : v{a, b} {}
  float& operator()(int i) { return v[i]; }
  const float& operator()(int i) const { return v[i]; }
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
};
template <> struct Matrix<float, 3, 3> {
  float m[9];   // row-major
  Matrix() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  Matrix<float, 3, 1> operator*(const Matrix<float, 3, 1>& x) const {
    return Matrix<float, 3, 1>(m[0] * x.v[0] + m[1] * x.v[1] + m[2] * x.v[2], m[3] * x.v[0] + m[4] * x.v[1] + m[5] * x.v[2],
                               m[6] * x.v[0] + m[7] * x.v[1] + m[8] * x.v[2]);
  }
  Matrix operator*(const Matrix& o) const {
    Matrix r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = m[3 * i] * o.m[j] + m[3 * i + 1] * o.m[3 + j] + m[3 * i + 2] * o.m[6 + j];
    return r;
  }
  Matrix transpose() const { Matrix r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = m[3 * j + i]; return r; }
};
template <> struct Matrix<float, 4, 4> { float m[16]; };
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
}  // namespace Eigen

namespace Sophus {
template <typename T> struct SE3 {
  Eigen::Matrix3f R;
  Eigen::Vector3f t;
  SE3() {}
  SE3(const Eigen::Matrix3f& R_, const Eigen::Vector3f& t_) : R(R_), t(t_) {}
  Eigen::Matrix3f rotationMatrix() const { return R; }
  Eigen::Vector3f translation() const { return t; }
  SE3 inverse() const { Eigen::Matrix3f Rt = R.transpose(); Eigen::Vector3f x = Rt * t; return SE3(Rt, Eigen::Vector3f(-x(0), -x(1), -x(2))); }
  Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return R * p + t; }
  Eigen::Matrix4f matrix() const { return Eigen::Matrix4f(); }
};
typedef SE3<float> SE3f;
}  // namespace Sophus

using namespace std;   // the reference's .cc files say so at file scope

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;

class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) = 0;
  std::vector<float> mvParameters;
};
class Pinhole : public GeometricCamera { public: Eigen::Vector2f project(const Eigen::Vector3f& v3D); };
class KannalaBrandt8 : public GeometricCamera { public: Eigen::Vector2f project(const Eigen::Vector3f& v3D); };

class ORBextractor { public: std::vector<cv::Mat> mvImagePyramid; };

class MapPoint {   // include/MapPoint.h: the tracking scratch (:170-181) and the getters the hot path calls
 public:
  Eigen::Vector3f GetWorldPos() { return mWorldPos; }
  Eigen::Vector3f GetNormal() { return mNormalVector; }
  float GetMinDistanceInvariance();
  float GetMaxDistanceInvariance();
  int PredictScale(const float& currentDist, Frame* pF);
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  float mTrackProjX = -1, mTrackProjY = -1, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
  bool mbTrackInView = false, mbTrackInViewR = false;
  int mnTrackScaleLevel = -1, mnTrackScaleLevelR = -1;
  float mTrackViewCos = 0, mTrackViewCosR = 0;
  long unsigned int mnLastFrameSeen = 0;
  // state behind the getters
  Eigen::Vector3f mWorldPos, mNormalVector;
  cv::Mat mDescriptor;
  float mfMinDistance = 0, mfMaxDistance = 0;
  int nObs = 0;
  bool mbBad = false;
  std::mutex mMutexPos;
};

class Frame {   // include/Frame.h
 public:
  void AssignFeaturesToGrid();
  bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
  bool isInFrustumChecks(MapPoint* pMP, float viewingCosLimit, bool bRight = false);
  vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                   const bool bRight = false) const;
  bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
  void ComputeStereoMatches();
  void ComputeStereoFromRGBD(const cv::Mat& imDepth);
  Sophus::SE3<float> GetPose() const { return mTcw; }
  Sophus::SE3f GetRelativePoseTrl() const { return mTrl; }

  ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
  GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
  float mbf = 0, mb = 0;
  int N = 0, Nleft = -1, Nright = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mDescriptors, mDescriptorsRight;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  long unsigned int mnId = 0;
  int mnScaleLevels = 0;
  float mfScaleFactor = 0, mfLogScaleFactor = 0;
  vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
  Sophus::SE3<float> mTcw, mTlr, mTrl;
  Eigen::Matrix<float, 3, 3> mRwc, mRcw;
  Eigen::Matrix<float, 3, 1> mOw, mtcw;
};

class KeyFrame {   // include/KeyFrame.h: what SearchByBoW reads
 public:
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  std::vector<MapPoint*> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mDescriptors;
  GeometricCamera* mpCamera2 = nullptr;
  int NLeft = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
};

class ORBmatcher {   // include/ORBmatcher.h
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                         const float thFarPoints = 50.0f);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;
 protected:
  float RadiusByViewingCos(const float& viewCos);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3

// include/Kernels/KernelController.h: the run-mode flags stay 0 so that the CPU branches of the extracted functions run
class KernelController {
 public:
  static bool searchLocalPointsKernelRunStatus, poseEstimationKernelRunStatus;
  static void launchSearchLocalPointsKernel(ORB_SLAM3::Frame&, const vector<ORB_SLAM3::MapPoint*>&, const float, const bool, const float,
                                            int*, int*, int*, int*, int*, int*, int*, int*, int*, int*);
  static void launchPoseEstimationKernel(ORB_SLAM3::Frame&, const ORB_SLAM3::Frame&, const float, const bool, const bool,
                                         Eigen::Matrix4f, int*, int*, int*, int*);
};
