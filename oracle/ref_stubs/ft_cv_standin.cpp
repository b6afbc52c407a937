// ft_cv_standin.cpp -- the image primitives of the OpenCV stand-in (opencv2/opencv.hpp in this directory), TEST
// INFRASTRUCTURE ONLY. Each forwards to the oracle primitive that is pinned bit-exactly against cv2 4.13
// (oracle/ft_oracle.cpp; tests/test_oracle_cv2_live.py), so that the reference's src/ORBextractor.cc, compiled against
// the stand-in, computes what it computes against the real library.
#include <cmath>

#include "../ft_oracle.h"
#include "opencv2/opencv.hpp"

namespace cv {

int cvRoundImpl(double v) { return (int)lrint(v); }   // round half to even, as cvRound's SSE2 conversion

static fto::Img to_img(const Mat& m) {
  fto::Img im; im.w = m.cols; im.h = m.rows; im.d.resize((size_t)m.cols * m.rows);
  for (int y = 0; y < m.rows; y++) memcpy(im.row(y), m.ptr(y), (size_t)m.cols);
  return im;
}
static void from_img(const fto::Img& im, Mat& m) {
  m.create(im.h, im.w, CV_8U);   // keeps a destination view of the right shape (the pyramid level inside its border)
  for (int y = 0; y < im.h; y++) memcpy(m.ptr(y), im.row(y), (size_t)im.w);
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int interpolation) {
  if (interpolation != INTER_LINEAR) throw std::runtime_error("stand-in cv::resize: INTER_LINEAR only");
  Mat src = src_.getMat();
  fto::Img out;
  fto::resize_linear_u8(to_img(src), out, dsize.width, dsize.height);
  from_img(out, dst_.ref());
}

static int reflect101(int p, int len) {   // cv::borderInterpolate(BORDER_REFLECT_101)
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) throw std::runtime_error("stand-in copyMakeBorder: REFLECT_101 only");
  Mat src = src_.getMat();
  Mat& dst = dst_.ref();
  dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
  const bool inplace = src.data == dst.data + (size_t)top * dst.step + left;
  // Without BORDER_ISOLATED OpenCV would read real pixels around a source ROI; the reference passes a whole image there.
  for (int y = 0; y < dst.rows; y++) {
    const int sy = reflect101(y - top, src.rows);
    for (int x = 0; x < dst.cols; x++) {
      const bool interior = y >= top && y < top + src.rows && x >= left && x < left + src.cols;
      if (interior && inplace) continue;
      dst.ptr(y)[x] = src.ptr(sy)[reflect101(x - left, src.cols)];
    }
  }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
  if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 || borderType != BORDER_REFLECT_101)
    throw std::runtime_error("stand-in GaussianBlur: 7x7, sigma 2, REFLECT_101 only");
  fto::Img out;
  fto::gaussian_blur_7x7_s2(to_img(src_.getMat()), out);
  from_img(out, dst_.ref());
}

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  if (!nonmaxSuppression) throw std::runtime_error("stand-in cv::FAST: nonmaxSuppression only");
  Mat m = image.getMat();
  std::vector<fto::Candidate> c;
  fto::fast_detect(m.data, (int)(size_t)m.step, m.cols, m.rows, threshold, c);
  keypoints.clear();
  for (const fto::Candidate& k : c) keypoints.push_back(KeyPoint(k.x, k.y, 7.f, -1.f, k.response));
}

float fastAtan2(float y, float x) { return fto::fast_atan2(y, x); }

}  // namespace cv
