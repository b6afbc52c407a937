// Stand-in for the reference's include/Kernels/KernelController.h, TEST INFRASTRUCTURE ONLY. The real header pulls in
// Frame.h / MapPoint.h (Eigen, Sophus, g2o); src/ORBextractor.cc only reads the run-mode flag, which stays 0 here so
// that the reference's CPU branch is the code that runs.
#pragma once
class KernelController {
 public:
  static bool orbExtractionKernelRunStatus;
};
