// ref_orbextractor_capi.cpp -- C entry points over the REFERENCE's own ORBextractor (TEST INFRASTRUCTURE ONLY).
//
// oracle/Makefile compiles the reference's src/ORBextractor.cc from where it lies under /root/reference, together with
// this file, against the OpenCV stand-in of oracle/ref_stubs (image primitives = the oracle's cv2-pinned ones) and the
// real CUDA runtime headers, into oracle/_ref/libft_ref_orbextractor.so. KernelController::orbExtractionKernelRunStatus
// stays 0, so operator() runs the reference's CPU branch: ComputePyramid, ComputeKeyPointsOctTree (per-cell FAST,
// DistributeOctTree / DivideNode / the std::sort of the careful phase), IC_Angle, computeOrbDescriptor and the
// lapping-area ordering -- the operator-level logic the oracle restates (oracle/ft_oracle.cpp). Nothing of the reference
// is copied into this repository.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBextractor.h"
#include "Kernels/KernelController.h"
#include "Kernels/CudaUtils.h"

bool KernelController::orbExtractionKernelRunStatus = false;

void checkCudaError(cudaError_t, const char*) {}

// launchers of the reference's own CUDA kernels (src/*.cu, not built here): only reachable in GPU run mode
static void gpu_only(const char* what) { fprintf(stderr, "ref_orbextractor: %s belongs to the reference's GPU branch\n", what); abort(); }
void fast_extract(uchar*, uchar*, uint8_t, uint8_t, uint8_t*, uint8_t*, int*, int, ORB_SLAM3::GpuPoint*, uint*, int, int, int, float*,
                  int, cudaStream_t, cudaEvent_t, cv::Mat) { gpu_only("fast_extract"); }
void compute_orientation(uchar*, uchar*, ORB_SLAM3::GpuPoint*, uint*, int, int*, int, int, int, int, float*, cudaStream_t) { gpu_only("compute_orientation"); }
void resize(uint, uint, float*, uchar*, uchar*, uint, uint, cudaStream_t) { gpu_only("resize"); }
void gaussian_blur(uchar*, uchar*, uchar*, uchar*, float*, int, int, int, float*, int, cudaStream_t) { gpu_only("gaussian_blur"); }
void compute_descriptor(uchar*, uchar*, ORB_SLAM3::GpuPoint*, uint*, int, cv::Point*, int, int, int, int, float*, cudaStream_t) { gpu_only("compute_descriptor"); }

extern "C" {

void* ftref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int w, int h) {
  return new ORB_SLAM3::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh, w, h);
}
void ftref_extractor_destroy(void* ex) { delete static_cast<ORB_SLAM3::ORBextractor*>(ex); }

// ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea): returns monoIndex; kps6[n][6] =
// x, y, size, angle, response, octave; desc[n][32]; *n = keypoints.size()
int ftref_extract(void* ex_, const unsigned char* img, int w, int h, int step, int lap0, int lap1, float* kps6,
                  unsigned char* desc, int cap, int* n) {
  ORB_SLAM3::ORBextractor* ex = static_cast<ORB_SLAM3::ORBextractor*>(ex_);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)step), mask, descriptors;
  std::vector<cv::KeyPoint> kps;
  std::vector<int> lap = {lap0, lap1};
  const int mono = (*ex)(image, mask, kps, descriptors, lap);
  *n = (int)kps.size();
  for (int i = 0; i < *n && i < cap; i++) {
    const cv::KeyPoint& k = kps[i];
    float* o = kps6 + 6 * (size_t)i;
    o[0] = k.pt.x; o[1] = k.pt.y; o[2] = k.size; o[3] = k.angle; o[4] = k.response; o[5] = (float)k.octave;
    memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
  }
  return mono;
}

// mvImagePyramid[level] (the inner ROI), tight; returns 0 when the level does not exist
int ftref_level_image(void* ex_, int level, unsigned char* out, int* w, int* h) {
  ORB_SLAM3::ORBextractor* ex = static_cast<ORB_SLAM3::ORBextractor*>(ex_);
  if (level < 0 || level >= (int)ex->mvImagePyramid.size()) return 0;
  const cv::Mat& m = ex->mvImagePyramid[level];
  *w = m.cols; *h = m.rows;
  if (out) for (int y = 0; y < m.rows; y++) memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
  return 1;
}
void ftref_scale_tables(void* ex_, float* scale, float* inv, float* sigma2, float* invSigma2) {
  ORB_SLAM3::ORBextractor* ex = static_cast<ORB_SLAM3::ORBextractor*>(ex_);
  std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(),
                     d = ex->GetInverseScaleSigmaSquares();
  for (size_t i = 0; i < a.size(); i++) { scale[i] = a[i]; inv[i] = b[i]; sigma2[i] = c[i]; invSigma2[i] = d[i]; }
}

}  // extern "C"
