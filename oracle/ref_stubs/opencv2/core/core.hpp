// Minimal stand-in for <opencv2/core/core.hpp>, TEST INFRASTRUCTURE ONLY.
// It exists so that the reference's vendored DBoW2 (Thirdparty/DBoW2, compiled where it lies under /root/reference by
// oracle/Makefile -> oracle/_ref/libft_ref_dbow2.so) builds without OpenCV: DBoW2 uses cv::Mat only as a 1x32 byte
// row (FORB.cpp) and cv::FileStorage only in its YAML save/load, which the ORB-SLAM3 path never calls (it uses
// loadFromTextFile). Mat::create zero-fills (real OpenCV leaves the bytes uninitialised).
#pragma once
// real core.hpp drags these in; DBoW2 relies on it
#include <math.h>
#include <sstream>
#include <iostream>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    data_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * (type == CV_32F ? 4 : 1), 0);
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  Mat clone() const {
    Mat m; m.rows = rows; m.cols = cols; m.type_ = type_;
    if (data_) m.data_ = std::make_shared<std::vector<unsigned char>>(*data_);
    return m;
  }
  void release() { rows = cols = 0; data_.reset(); }
  bool empty() const { return !data_ || data_->empty(); }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data_->data() + (size_t)r * cols * esz()); }
  template <typename T> const T* ptr(int r = 0) const {
    return reinterpret_cast<const T*>(data_->data() + (size_t)r * cols * esz());
  }
 private:
  int type_ = CV_8U;
  size_t esz() const { return type_ == CV_32F ? 4 : 1; }
  std::shared_ptr<std::vector<unsigned char>> data_;
};

// YAML storage is not available in this stand-in: every use throws.
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
