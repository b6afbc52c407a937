// ref_gpu_capi.cpp -- the REFERENCE's own CUDA extractor kernels, timed on this box (BENCH INFRASTRUCTURE ONLY).
//
// FastTrack's extractor offload is five CUDA files (/root/reference/src/{resize,gaussian_blur,fast,orientation,
// descriptor}.cu, launchers declared in include/{resize,gaussian_blur,fast,orientation,descriptor}.h). oracle/Makefile
// compiles them with nvcc from where they lie (nothing is copied) and links them, the reference's src/ORBextractor.cc and
// this file into oracle/_ref/libft_ref_orbextractor_gpu.so. Two things are exposed:
//   * ftrefgpu_extract      ORBextractor::operator() with KernelController::orbExtractionKernelRunStatus = 1, i.e. the
//                            reference's GPU branch end to end (H2D, its kernels, D2H of every corner, DistributeOctTreeGPU
//                            on the host, ORBextractor.cc:1374-1380, 1228-1286, 1522-1544);
//   * ftrefgpu_stereo_time   StereoMatchKernel::launch (src/Kernels/StereoMatchKernel.cu:351-490, compiled with CudaUtils.cu) as
//                            Frame::ComputeStereoMatchesGPU drives it (src/Frame.cc:1007-1063): row table, the pyramid left on the
//                            GPU by the reference's own resize kernel, the (distance, index) sort and the median filter on the host;
//   * ftrefgpu_stage_times  the five launchers called in the order ComputePyramidGPU / ComputeKeyPointsOctTreeGPU call them
//                            (ORBextractor.cc:1524-1531, 1233-1236), on buffers sized as allocMemory / allocInputMemory size
//                            them (:1320-1340), with CUDA events around each launcher -> the reference's per-stage GPU time.
// The reference's GPU results are NOT bit-compatible with its CPU branch; this file only measures.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBextractor.h"
#include "Kernels/KernelController.h"
#include "fast.h"
#include "resize.h"
#include "gaussian_blur.h"
#include "orientation.h"
#include "descriptor.h"
#include "Kernels/StereoMatchKernel.h"
#include <chrono>

bool KernelController::orbExtractionKernelRunStatus = true;
namespace ORB_SLAM3 { void generateGaussian(float K[]); }   // src/ORBextractor.cc:122-130 (no declaration in the header)

// checkCudaError comes from the reference's src/Kernels/CudaUtils.cu (linked in for StereoMatchKernel.cu)

extern "C" {

void* ftrefgpu_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int w, int h) {
  return new ORB_SLAM3::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh, w, h);
}
void ftrefgpu_extractor_destroy(void* ex) { delete static_cast<ORB_SLAM3::ORBextractor*>(ex); }

// the reference's GPU branch of operator(); returns the number of keypoints
int ftrefgpu_extract(void* ex_, const unsigned char* img, int w, int h, int step) {
  ORB_SLAM3::ORBextractor* ex = static_cast<ORB_SLAM3::ORBextractor*>(ex_);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)step), mask, descriptors;
  std::vector<cv::KeyPoint> kps;
  std::vector<int> lap = {0, 0};
  (*ex)(image, mask, kps, descriptors, lap);
  return (int)kps.size();
}

// Per-launcher device times in ms, averaged over `reps` images: out[0..4] = resize, gaussian_blur, fast_extract,
// compute_orientation, compute_descriptor; out[5] = corners found on the last repetition (all levels).
int ftrefgpu_stage_times(const unsigned char* img, int w, int h, int step, int nlevels, float scaleFactor, int iniTh, int minTh,
                         int reps, float* out) {
  std::vector<float> sf(nlevels, 1.f);
  for (int i = 1; i < nlevels; i++) sf[i] = sf[i - 1] * scaleFactor;
  // umax and the sampling pattern as ORBextractor's constructor builds them (:478-499, 468-474)
  const int HALF = 15;
  std::vector<int> umax(HALF + 1);
  {
    int v, v0, vmax = (int)floor(HALF * sqrt(2.f) / 2 + 1), vmin = (int)ceil(HALF * sqrt(2.f) / 2);
    const double hp2 = HALF * HALF;
    for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = HALF, v0 = 0; v >= vmin; --v) { while (umax[v0] == umax[v0 + 1]) ++v0; umax[v] = v0; ++v0; }
  }
  int points[32] = {0, 3, 1, 3, 2, 2, 3, 1, 3, 0, 3, -1, 2, -2, 1, -3, 0, -3, -1, -3, -2, -2, -3, -1, -3, 0, -3, 1, -2, 2, -1, 3};
  float k[7 * 7];
  ORB_SLAM3::generateGaussian(k);
  // the rBRIEF pattern: the extractor object owns it (protected); an instance built in GPU mode uploads it, so borrow one
  struct Peek : ORB_SLAM3::ORBextractor {
    Peek(int nf, float s, int nl, int a, int b, int w_, int h_) : ORB_SLAM3::ORBextractor(nf, s, nl, a, b, w_, h_) {}
    using ORB_SLAM3::ORBextractor::pattern;
  };
  std::vector<cv::Point> pattern;
  { Peek ex(1200, scaleFactor, nlevels, iniTh, minTh, w, h); pattern = ex.pattern; }

  cudaStream_t st;
  cudaEvent_t ev[6], inter;
  if (cudaStreamCreate(&st) != cudaSuccess) return -1;
  for (int i = 0; i < 6; i++) cudaEventCreate(&ev[i]);
  cudaEventCreateWithFlags(&inter, cudaEventDisableTiming);
  const size_t px = (size_t)w * h;
  uint8_t *d_R, *d_R_low; uchar *d_images, *d_imagesBlured, *d_in, *d_inBlured;
  ORB_SLAM3::GpuPoint* d_corner; uint* d_sizes; int *d_points, *d_umax; float *d_sf, *d_kernel; cv::Point* d_pattern;
  cudaMalloc(&d_R, px * nlevels); cudaMalloc(&d_R_low, px * nlevels);
  cudaMalloc(&d_corner, sizeof(ORB_SLAM3::GpuPoint) * px * nlevels);
  cudaMalloc(&d_images, px * nlevels); cudaMalloc(&d_imagesBlured, px * nlevels);
  cudaMalloc(&d_in, (size_t)h * step); cudaMalloc(&d_inBlured, (size_t)h * step);
  cudaMalloc(&d_sizes, sizeof(uint) * nlevels); cudaMalloc(&d_points, sizeof(points)); cudaMalloc(&d_umax, sizeof(int) * umax.size());
  cudaMalloc(&d_sf, sizeof(float) * nlevels); cudaMalloc(&d_kernel, sizeof(k)); cudaMalloc(&d_pattern, sizeof(cv::Point) * pattern.size());
  cudaMemcpy(d_points, points, sizeof(points), cudaMemcpyHostToDevice);
  cudaMemcpy(d_umax, umax.data(), sizeof(int) * umax.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_sf, sf.data(), sizeof(float) * nlevels, cudaMemcpyHostToDevice);
  cudaMemcpy(d_kernel, k, sizeof(k), cudaMemcpyHostToDevice);
  cudaMemcpy(d_pattern, pattern.data(), sizeof(cv::Point) * pattern.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_in, img, (size_t)h * step, cudaMemcpyHostToDevice);
  cv::Mat level0(h, w, CV_8UC1, (void*)img, (size_t)step);
  double acc[5] = {0, 0, 0, 0, 0};
  std::vector<uint> sizes(nlevels);
  for (int r = 0; r < reps + 2; r++) {       // two warm-up rounds
    cudaEventRecord(ev[0], st);
    resize(h, w, d_sf, d_in, d_images, nlevels, step, st);
    cudaEventRecord(ev[1], st);
    gaussian_blur(d_images, d_in, d_imagesBlured, d_inBlured, d_kernel, w, h, step, d_sf, nlevels, st);
    cudaEventRecord(ev[2], st);
    fast_extract(d_images, d_in, (uint8_t)iniTh, (uint8_t)minTh, d_R, d_R_low, d_points, 12, d_corner, d_sizes, w, h, step, d_sf, nlevels, st,
                 inter, level0);
    cudaEventRecord(ev[3], st);
    compute_orientation(d_images, d_in, d_corner, d_sizes, w * h, d_umax, step, nlevels, w, h, d_sf, st);
    cudaEventRecord(ev[4], st);
    compute_descriptor(d_imagesBlured, d_inBlured, d_corner, d_sizes, w * h, d_pattern, step, nlevels, w, h, d_sf, st);
    cudaEventRecord(ev[5], st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { fprintf(stderr, "ftrefgpu_stage_times: %s\n", cudaGetErrorString(cudaGetLastError())); return -2; }
    if (r >= 2)
      for (int i = 0; i < 5; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); acc[i] += ms; }
  }
  cudaMemcpy(sizes.data(), d_sizes, sizeof(uint) * nlevels, cudaMemcpyDeviceToHost);
  uint total = 0;
  for (int l = 0; l < nlevels; l++) total += sizes[l];
  for (int i = 0; i < 5; i++) out[i] = (float)(acc[i] / reps);
  out[5] = (float)total;
  cudaFree(d_R); cudaFree(d_R_low); cudaFree(d_corner); cudaFree(d_images); cudaFree(d_imagesBlured); cudaFree(d_in); cudaFree(d_inBlured);
  cudaFree(d_sizes); cudaFree(d_points); cudaFree(d_umax); cudaFree(d_sf); cudaFree(d_kernel); cudaFree(d_pattern);
  for (int i = 0; i < 6; i++) cudaEventDestroy(ev[i]);
  cudaEventDestroy(inter);
  cudaStreamDestroy(st);
  return 0;
}


// Wall-clock ms per call of the reference's GPU stereo matching (Frame::ComputeStereoMatchesGPU, src/Frame.cc:1007-1063) on one
// extracted stereo pair: keypoints as (x, y, octave) triples, 32-byte descriptors. out[0] = ms per call, out[1] = matches kept.
int ftrefgpu_stereo_time(const unsigned char* imgL, const unsigned char* imgR, int w, int h, int step, int nlevels, float scaleFactor,
                         const float* kL, int nL, const unsigned char* dL, const float* kR, int nR, const unsigned char* dR,
                         float mbf, float mb, int reps, float* out) {
  CudaUtils::loadSetting(nL > nR ? nL : nR, nlevels, false, scaleFactor, w, h, false);
  std::vector<float> sf(nlevels, 1.f);
  for (int i = 1; i < nlevels; i++) sf[i] = sf[i - 1] * scaleFactor;
  const size_t px = (size_t)w * h;
  uchar *d_in, *d_pyrL, *d_pyrR; float* d_sf;
  cudaMalloc(&d_in, (size_t)h * step); cudaMalloc(&d_pyrL, px * nlevels); cudaMalloc(&d_pyrR, px * nlevels);
  cudaMalloc(&d_sf, sizeof(float) * nlevels);
  cudaMemcpy(d_sf, sf.data(), sizeof(float) * nlevels, cudaMemcpyHostToDevice);
  cudaMemcpy(d_in, imgL, (size_t)h * step, cudaMemcpyHostToDevice);
  resize(h, w, d_sf, d_in, d_pyrL, nlevels, step, 0);       // ORBextractor::GetGPUPyramid() = d_images after ComputePyramidGPU
  cudaDeviceSynchronize();
  cudaMemcpy(d_in, imgR, (size_t)h * step, cudaMemcpyHostToDevice);
  resize(h, w, d_sf, d_in, d_pyrR, nlevels, step, 0);
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::vector<cv::KeyPoint> keysL(nL), keysR(nR);
  for (int i = 0; i < nL; i++) { keysL[i].pt.x = kL[3 * i]; keysL[i].pt.y = kL[3 * i + 1]; keysL[i].octave = (int)kL[3 * i + 2]; }
  for (int i = 0; i < nR; i++) { keysR[i].pt.x = kR[3 * i]; keysR[i].pt.y = kR[3 * i + 1]; keysR[i].octave = (int)kR[3 * i + 2]; }
  cv::Mat descL(nL, 32, CV_8UC1, (void*)dL, 32), descR(nR, 32, CV_8UC1, (void*)dR, 32);
  std::vector<cv::Mat> pyrL(1, cv::Mat(h, w, CV_8UC1, (void*)imgL, (size_t)step)), pyrR(1, cv::Mat(h, w, CV_8UC1, (void*)imgR, (size_t)step));
  StereoMatchKernel kernel;
  kernel.initialize();
  double total = 0;
  int kept = 0;
  for (int r = 0; r < reps + 1; r++) {
    const auto t0 = std::chrono::steady_clock::now();
    // Frame::ComputeStereoMatchesGPU (src/Frame.cc:1007-1063)
    const int thOrbDist = (100 + 50) / 2;
    std::vector<std::vector<int> > vRowIndices(h, std::vector<int>());
    for (int i = 0; i < h; i++) vRowIndices[i].reserve(MAX_FEATURES_IN_ROW_SLIDING_WINDOW);
    for (int iR = 0; iR < nR; iR++) {
      const float kpY = keysR[iR].pt.y;
      const float rr = 2.0f * sf[keysR[iR].octave];
      const int maxr = (int)ceil(kpY + rr), minr = (int)floor(kpY - rr);
      for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < h) vRowIndices[yi].push_back(iR);
    }
    const float minD = 0, maxD = mbf / mb;
    std::vector<std::pair<int, int> > vDistIdx;
    std::vector<float> mvuRight, mvDepth;
    kernel.launch(vRowIndices, d_pyrL, d_pyrR, pyrL, pyrR, keysL, keysR, descL, descR, minD, maxD, thOrbDist, mbf, true, vDistIdx, mvuRight, mvDepth);
    std::sort(vDistIdx.begin(), vDistIdx.end());
    if (!vDistIdx.empty()) {
      const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
      const float thDist = 1.5f * 1.4f * median;
      for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        if (vDistIdx[i].second >= 0) { mvuRight[vDistIdx[i].second] = -1; mvDepth[vDistIdx[i].second] = -1; }
      }
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (r > 0) total += std::chrono::duration<double, std::milli>(t1 - t0).count();
    kept = 0;
    for (float d : mvDepth) kept += d > 0;
  }
  kernel.shutdown();
  cudaFree(d_in); cudaFree(d_pyrL); cudaFree(d_pyrR); cudaFree(d_sf);
  out[0] = (float)(total / reps); out[1] = (float)kept;
  return 0;
}

}  // extern "C"
