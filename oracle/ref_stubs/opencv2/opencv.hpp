// Stand-in for the OpenCV headers, TEST INFRASTRUCTURE ONLY.
//
// It lets two pieces of the REFERENCE compile here, from where they lie under /root/reference, without OpenCV C++:
//   * Thirdparty/DBoW2 (cv::Mat as a 1x32 byte row)                      -> oracle/_ref/libft_ref_dbow2.so
//   * src/ORBextractor.cc, CPU branch (KernelController::orbExtractionKernelRunStatus == 0)
//                                                                         -> oracle/_ref/libft_ref_orbextractor.so
// Only the slice of the API those files touch exists. The image primitives (cv::resize INTER_LINEAR 8U, copyMakeBorder
// REFLECT_101, GaussianBlur 7x7 sigma 2, FAST 9/16 + NMS, fastAtan2, cvRound) are implemented in ft_cv_standin.cpp by
// the oracle's primitives, each of which is pinned bit-exactly against cv2 4.13 (tests/test_oracle_cv2_live.py,
// tests/golden/cv2_primitives.npz). cv::Mat::create zero-fills (real OpenCV leaves the bytes uninitialised).
#pragma once
// real opencv.hpp drags the standard headers in; the reference relies on it
#include <math.h>
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <list>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

namespace cv {

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_& operator*=(U s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

class KeyPoint {
 public:
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct MatStep {
  size_t v;
  MatStep() : v(0) {}
  operator size_t() const { return v; }
  size_t operator[](int) const { return v; }
  MatStep& operator=(size_t s) { v = s; return *this; }
};

class Mat {
 public:
  int rows, cols;
  uchar* data;
  MatStep step;
  Mat() : rows(0), cols(0), data(nullptr), type_(CV_8U) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), type_(CV_8U) { create(r, c, type); }
  Mat(Size sz, int type) : rows(0), cols(0), data(nullptr), type_(CV_8U) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* d, size_t s = 0) : rows(r), cols(c), data((uchar*)d), type_(type) { step = s ? s : (size_t)c * esz(); }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;   // OpenCV keeps a buffer of the right shape
    rows = r; cols = c; type_ = type;
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * esz() + 1, 0);
    data = buf_->data(); step = (size_t)c * esz();
  }
  static Mat zeros(int r, int c, int type) { Mat m; m.create(r, c, type); if (m.data) memset(m.data, 0, (size_t)r * m.step); return m; }
  Mat clone() const {
    Mat m;
    if (!data) return m;
    m.create(rows, cols, type_);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * esz());
    return m;
  }
  void copyTo(const Mat& dst) const {   // into an existing view of the same shape (descriptors.row(i))
    if (dst.rows != rows || dst.cols != cols || !dst.data) throw std::runtime_error("stand-in Mat::copyTo: shape mismatch");
    for (int y = 0; y < rows; y++) memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * esz());
  }
  void release() { rows = cols = 0; data = nullptr; buf_.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t step1() const { return step.v / esz(); }
  Size size() const { return Size(cols, rows); }
  Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + (size_t)r.x * esz(); m.rows = r.height; m.cols = r.width; return m; }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat reshape(int /*channels*/) const { return *this; }   // N x 2 floats <-> N two-channel points: same bytes
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
  template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data + (size_t)y * step)[x]; }
  template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data + (size_t)y * step)[x]; }
  template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }   // single-index access of a vector
  template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
 private:
  int type_;
  size_t esz() const { return type_ == CV_32F ? 4 : 1; }
  std::shared_ptr<std::vector<uchar>> buf_;
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_->empty(); }
  Mat getMat() const { return *m_; }
 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  _OutputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}   // a temporary view that is written in place
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
  Mat& ref() const { return *m_; }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

// implemented in ft_cv_standin.cpp on the oracle's cv2-pinned primitives
int cvRoundImpl(double v);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);

// YAML storage is not available in this stand-in: every use throws (DBoW2's save/load, never called on this path)
class FileNode {
 public:
  FileNode operator[](const char*) const { fail(); return FileNode(); }
  FileNode operator[](const std::string&) const { fail(); return FileNode(); }
  FileNode operator[](int) const { fail(); return FileNode(); }
  size_t size() const { fail(); return 0; }
  operator int() const { fail(); return 0; }
  operator double() const { fail(); return 0; }
  operator std::string() const { fail(); return std::string(); }
 private:
  static void fail() { throw std::runtime_error("cv::FileStorage is not available in the oracle's OpenCV stand-in"); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const char*, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { return FileNode()[0]; }
  void release() {}
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) {
  throw std::runtime_error("cv::FileStorage is not available in the oracle's OpenCV stand-in");
  return fs;
}

}  // namespace cv

// cvRound / cvFloor / cvCeil live in the global namespace in OpenCV (core/fast_math.hpp)
inline int cvRound(double v) { return cv::cvRoundImpl(v); }
inline int cvRound(float v) { return cv::cvRoundImpl((double)v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
