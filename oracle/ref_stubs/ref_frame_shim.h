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
#include "ft_oracle.h"                 // knn2_hamming for the BFMatcher stand-in
#include "DBoW2/BowVector.h"           // the reference's own (Thirdparty/DBoW2)
#include "DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48             // reference include/Frame.h:46-47
#define FRAME_GRID_COLS 64

namespace cv {
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
enum { NORM_L1 = 2, NORM_HAMMING = 6 };
// cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2): the oracle's knn2_hamming, whose ordering (ascending
// distance, lowest trainIdx first on ties) is pinned against cv2 (tests/test_oracle_cv2_live.py)
class BFMatcher {
 public:
  explicit BFMatcher(int = NORM_HAMMING) {}
  void knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch> >& matches, int k) const {
    if (k != 2) throw std::runtime_error("stand-in BFMatcher: k = 2 only");
    matches.assign(q.rows, std::vector<DMatch>());
    if (q.rows == 0 || t.rows == 0) return;
    std::vector<unsigned char> qd((size_t)q.rows * 32), td((size_t)t.rows * 32);
    for (int i = 0; i < q.rows; i++) memcpy(&qd[(size_t)i * 32], q.ptr(i), 32);
    for (int i = 0; i < t.rows; i++) memcpy(&td[(size_t)i * 32], t.ptr(i), 32);
    std::vector<int> idx((size_t)q.rows * 2), dist((size_t)q.rows * 2);
    fto::knn2_hamming(qd.data(), q.rows, td.data(), t.rows, idx.data(), dist.data());
    for (int i = 0; i < q.rows; i++)
      for (int j = 0; j < 2; j++)
        if (idx[2 * i + j] >= 0) matches[i].push_back(DMatch{i, idx[2 * i + j], 0, (float)dist[2 * i + j]});
  }
};
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {   // 8U: integer sum of absolute differences
  long s = 0;
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) s += std::abs((int)a.ptr(y)[x] - (int)b.ptr(y)[x]);
  return (double)s;
}
// cv::undistortPoints(src, dst, K, distCoeffs, R = Mat(), P) for N x 2 float points with P == K: the oracle's restatement,
// pinned bit-exactly against cv2.undistortPoints on 80k points (tests/golden/cv2_undistort.npz + live test)
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& /*R*/, const Mat& P) {
  const float k4[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
  const float p4[4] = {P.at<float>(0, 0), P.at<float>(1, 1), P.at<float>(0, 2), P.at<float>(1, 2)};
  for (int i = 0; i < 4; i++) if (k4[i] != p4[i]) throw std::runtime_error("stand-in undistortPoints: P must equal K");
  const int n = src.rows, nd = dist.rows * dist.cols;
  std::vector<float> in((size_t)2 * n), out((size_t)2 * n), d(nd);
  for (int i = 0; i < n; i++) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
  for (int i = 0; i < nd; i++) d[i] = dist.ptr<float>()[i];
  fto::undistort_points(in.data(), n, k4, d.data(), nd, out.data());
  dst.create(n, 2, CV_32F);
  for (int i = 0; i < n; i++) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
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
  static Matrix Zero() { return Matrix(); }
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
  static Matrix Identity() { return Matrix(); }
  Matrix<float, 3, 1> row(int i) const { return Matrix<float, 3, 1>(m[3 * i], m[3 * i + 1], m[3 * i + 2]); }
};
struct Row4 {   // a row of a 3x4 / 4x4 matrix as a value
  float v[4];
  Row4 operator-(const Row4& o) const { return Row4{{v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2], v[3] - o.v[3]}}; }
};
inline Row4 operator*(float s, const Row4& r) { return Row4{{s * r.v[0], s * r.v[1], s * r.v[2], s * r.v[3]}}; }
template <> struct Matrix<float, 4, 1> {
  float v[4];
  float operator()(int i) const { return v[i]; }
  Matrix<float, 3, 1> head(int) const { return Matrix<float, 3, 1>(v[0], v[1], v[2]); }
};
template <> struct Matrix<float, 4, 4> {
  float m[16];
  struct RowRef { float* p; void operator=(const Row4& r) { for (int i = 0; i < 4; i++) p[i] = r.v[i]; } };
  RowRef row(int i) { return RowRef{m + 4 * i}; }
  Matrix<float, 4, 1> col(int j) const { Matrix<float, 4, 1> c; for (int i = 0; i < 4; i++) c.v[i] = m[4 * i + j]; return c; }
};
template <> struct Matrix<float, 3, 4> {
  float m[12];
  Row4 row(int i) const { return Row4{{m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]}}; }
  float operator()(int i, int j) const { return m[4 * i + j]; }
  Matrix<float, 3, 1> col(int j) const { return Matrix<float, 3, 1>(m[j], m[4 + j], m[8 + j]); }
  // `M << R, t;` (comma initialiser with a 3x3 block and a column)
  struct Comma {
    Matrix* M;
    Comma operator,(const Matrix<float, 3, 1>& t) { for (int i = 0; i < 3; i++) M->m[4 * i + 3] = t.v[i]; return *this; }
  };
  Comma operator<<(const Matrix<float, 3, 3>& R) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[4 * i + j] = R.m[3 * i + j]; return Comma{this}; }
};
inline Matrix<float, 3, 1> operator/(const Matrix<float, 3, 1>& a, float s) { return Matrix<float, 3, 1>(a.v[0] / s, a.v[1] / s, a.v[2] / s); }
inline Matrix<float, 3, 3> operator-(const Matrix<float, 3, 3>& a) { Matrix<float, 3, 3> r; for (int i = 0; i < 9; i++) r.m[i] = -a.m[i]; return r; }
enum { ComputeFullV = 1 };
// Eigen::JacobiSVD<Matrix4f>(A, ComputeFullV).matrixV().col(3): Eigen is not available, so the smallest right-singular
// vector comes from the same cyclic Jacobi on A^T A in double that the oracle uses (a TOLERANCE stand-in):
// this pins the control flow of TriangulateMatches, not the last ulp of the triangulated point.
template <typename M> class JacobiSVD {
 public:
  JacobiSVD(const Matrix<float, 4, 4>& A, int) {
    double Md[4][4], V[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += (double)A.m[4 * k + i] * (double)A.m[4 * k + j];
      Md[i][j] = s; V[i][j] = (i == j);
    }
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0;
      for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) off += Md[i][j] * Md[i][j];
      if (off < 1e-300) break;
      for (int p = 0; p < 3; p++) for (int q = p + 1; q < 4; q++) {
        if (std::fabs(Md[p][q]) < 1e-300) continue;
        const double theta = (Md[q][q] - Md[p][p]) / (2 * Md[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 4; k++) { const double a = Md[k][p], b = Md[k][q]; Md[k][p] = c * a - s * b; Md[k][q] = s * a + c * b; }
        for (int k = 0; k < 4; k++) { const double a = Md[p][k], b = Md[q][k]; Md[p][k] = c * a - s * b; Md[q][k] = s * a + c * b; }
        for (int k = 0; k < 4; k++) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
      }
    }
    int best = 0;
    for (int i = 1; i < 4; i++) if (Md[i][i] < Md[best][best]) best = i;
    // singular values descend in Eigen: the smallest one's vector is column 3 of V
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) v_.m[4 * i + j] = 0.f;
    for (int k = 0; k < 4; k++) v_.m[4 * k + 3] = (float)V[k][best];
  }
  const Matrix<float, 4, 4>& matrixV() const { return v_; }
 private:
  Matrix<float, 4, 4> v_;
};
typedef Matrix<float, 4, 1> Vector4f;
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
  virtual Eigen::Vector3f unprojectEig(const cv::Point2f& p2D) = 0;
  std::vector<float> mvParameters;
};
class Pinhole : public GeometricCamera {
 public:
  Eigen::Vector2f project(const Eigen::Vector3f& v3D);
  cv::Mat toK() {   // Pinhole::toK (src/CameraModels/Pinhole.cpp): fx 0 cx; 0 fy cy; 0 0 1
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32F);
    K.at<float>(0, 0) = mvParameters[0]; K.at<float>(1, 1) = mvParameters[1];
    K.at<float>(0, 2) = mvParameters[2]; K.at<float>(1, 2) = mvParameters[3]; K.at<float>(2, 2) = 1.f;
    return K;
  }
  Eigen::Vector3f unprojectEig(const cv::Point2f&) { abort(); }
};
class KannalaBrandt8 : public GeometricCamera {   // include/CameraModels/KannalaBrandt8.h
 public:
  KannalaBrandt8() : precision(1e-6) {}
  Eigen::Vector2f project(const Eigen::Vector3f& v3D);
  Eigen::Vector3f unprojectEig(const cv::Point2f& p2D);
  cv::Point3f unproject(const cv::Point2f& p2D);
  float TriangulateMatches(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                           const Eigen::Vector3f& t12, const float sigmaLevel, const float unc, Eigen::Vector3f& p3D);
 private:
  const float precision;
  void Triangulate(const cv::Point2f& p1, const cv::Point2f& p2, const Eigen::Matrix<float, 3, 4>& Tcw1,
                   const Eigen::Matrix<float, 3, 4>& Tcw2, Eigen::Vector3f& x3D);
};

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
  void UndistortKeyPoints();
  void ComputeImageBounds(const cv::Mat& imLeft);
  cv::Mat mK, mDistCoef;
  void ComputeStereoMatches();
  void ComputeStereoFishEyeMatches();
  void ComputeStereoFromRGBD(const cv::Mat& imDepth);
  int monoLeft = 0, monoRight = 0, mnCloseMPs = 0;
  std::vector<Eigen::Vector3f> mvStereo3Dpoints;
  cv::BFMatcher BFmatcher;
  Eigen::Matrix<float, 3, 3> mRlr;
  Eigen::Vector3f mtlr;
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
