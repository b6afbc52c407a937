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

