// ft_camera.cuh -- camera projection shared by the stereo and projection-search kernels.
// Every float operation is individually rounded (__f*_rn) in the order the reference writes it,
// so results do not depend on FMA contraction.
#pragma once
#include "ft_device.cuh"

__device__ __forceinline__ void ft_cam_project(const FtCamera& c, const float P[3], float uv[2]) {
  if (c.type == FT_CAM_PINHOLE) {   // Pinhole::project (Pinhole.cpp:43-49)
    uv[0] = __fadd_rn(__fdiv_rn(__fmul_rn(c.p[0], P[0]), P[2]), c.p[2]);
    uv[1] = __fadd_rn(__fdiv_rn(__fmul_rn(c.p[1], P[1]), P[2]), c.p[3]);
  } else {                           // KannalaBrandt8::project (KannalaBrandt8.cpp:67-84)
    const float x2y2 = __fadd_rn(__fmul_rn(P[0], P[0]), __fmul_rn(P[1], P[1]));
    const float theta = atan2f(sqrtf(x2y2), P[2]);
    const float psi = atan2f(P[1], P[0]);
    const float t2 = __fmul_rn(theta, theta), t3 = __fmul_rn(theta, t2), t5 = __fmul_rn(t3, t2), t7 = __fmul_rn(t5, t2),
                t9 = __fmul_rn(t7, t2);
    const float r = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(theta, __fmul_rn(c.p[4], t3)), __fmul_rn(c.p[5], t5)),
                                        __fmul_rn(c.p[6], t7)), __fmul_rn(c.p[7], t9));
    uv[0] = __fadd_rn(__fmul_rn(__fmul_rn(c.p[0], r), cosf(psi)), c.p[2]);
    uv[1] = __fadd_rn(__fmul_rn(__fmul_rn(c.p[1], r), sinf(psi)), c.p[3]);
  }
}

__device__ __forceinline__ void ft_mat3_vec(const float R[9], const float v[3], float o[3]) {
  for (int i = 0; i < 3; i++)
    o[i] = __fadd_rn(__fadd_rn(__fmul_rn(R[3 * i], v[0]), __fmul_rn(R[3 * i + 1], v[1])), __fmul_rn(R[3 * i + 2], v[2]));
}


// cv::undistortPoints(pt, K, distCoef, Mat(), K) as Frame::UndistortKeyPoints calls it (reference src/Frame.cc:771-804):
// OpenCV's iteration in double, five fixed-point steps (default TermCriteria(MAX_ITER, 5, 0.01)), the zero-valued
// rational / thin-prism terms kept so that every rounding (and the sign of a zero) matches. Host + device: the context
// uses it for Frame::ComputeImageBounds. Compiled with -fmad=false; the host compiler does not contract either.
__host__ __device__ inline void ft_undistort_point(const FtUndistort& u, float xf, float yf, float& xo, float& yo) {
  const double k0 = u.k[0], k1 = u.k[1], k2 = u.k[2], k3 = u.k[3], k4 = u.k[4];
  const double z = 0.0;
  double x = (double)xf, y = (double)yf;
  const double uu = x, vv = y;
  x = (x - u.cx) * u.ifx;
  y = (y - u.cy) * u.ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((z * r2 + z) * r2 + z) * r2) / (1 + ((k4 * r2 + k1) * r2 + k0) * r2);
    if (icdist < 0) {
      x = (uu - u.cx) * u.ifx;
      y = (vv - u.cy) * u.ify;
      break;
    }
    const double deltaX = 2 * k2 * x * y + k3 * (r2 + 2 * x * x) + z * r2 + z * r2 * r2;
    const double deltaY = k2 * (r2 + 2 * y * y) + 2 * k3 * x * y + z * r2 + z * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  const double xx = u.fx * x + z * y + u.cx;
  const double yy = z * x + u.fy * y + u.cy;
  const double ww = 1. / (z * x + z * y + 1.);
  xo = (float)(xx * ww);
  yo = (float)(yy * ww);
}
