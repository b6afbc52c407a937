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

// cosf / sinf as glibc computes them (the reference calls cos(float)/sin(float), ORBextractor.cc:74). glibc's
// sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM optimized-routines algorithm) evaluate a fixed degree-7/8
// minimax polynomial in double after a multiple-of-pi/2 reduction and round once to float; the result is within
// 0.56 ulp but NOT always the correctly rounded float, and a rotated rBRIEF sample can sit exactly on a .5 rounding
// boundary (seen: sinf(0.73145765f), sample (-9,-2) -> fy = -7.5). Evaluating the same polynomial in double
// reproduces libm's float except when the double result lies within ~1e-16 of a float rounding boundary (p ~ 1e-9);
// checked against libm on 2e7 angles in [0, 2pi); tests/test_abi.py pins it against the host libm through ft_debug_sincosf.
// Valid for |y| < 120 (angles here are in [0, 2pi]).
__host__ __device__ __forceinline__ double ft_sincosf_poly(double x, double x2, int n, bool neg) {
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
               c4 = 0x1.99343027bf8c3p-16;
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double t1 = s2 + x2 * s3;
    const double x7 = x3 * x2;
    const double s = x + x3 * s1;
    return s + x7 * t1;
  }
  const double sg = neg ? -1.0 : 1.0;   // the second table of glibc holds the negated cosine coefficients
  const double x4 = x2 * x2;
  const double t2 = sg * c3 + x2 * (sg * c4);
  const double t1 = sg * c0 + x2 * (sg * c1);
  const double x6 = x4 * x2;
  const double c = t1 + x4 * (sg * c2);
  return c + x6 * t2;
}
__host__ __device__ __forceinline__ void ft_glibc_sincosf(float y, float& sn, float& cs) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  double x = (double)y;
  const float ay = fabsf(y);
  if (ay < 0.75f) {   // glibc compares the top 12 bits with those of pi/4 (0x3f4): the no-reduction path is |y| < 0.75
    const double x2 = x * x;
    if (ay < 0x1p-12f) { sn = y; cs = 1.0f; return; }
    sn = (float)ft_sincosf_poly(x, x2, 0, false);
    cs = (float)ft_sincosf_poly(x, x2, 1, false);
    return;
  }
  const double r = x * hpi_inv;
  const int n = ((int)r + 0x800000) >> 24;
  x = x - (double)n * hpi;
  const double x2 = x * x;
  // sine: sign[n & 3] = {1,-1,-1,1}, negated-cosine table when n & 2
  { const int q = n & 3; const double sg = (q == 1 || q == 2) ? -1.0 : 1.0; sn = (float)ft_sincosf_poly(x * sg, x2, n, (n & 2) != 0); }
  // cosine: sign[(n + 1) & 3], table by (n + 1) & 2, polynomial parity n ^ 1
  { const int q = (n + 1) & 3; const double sg = (q == 1 || q == 2) ? -1.0 : 1.0; cs = (float)ft_sincosf_poly(x * sg, x2, n ^ 1, ((n + 1) & 2) != 0); }
}

