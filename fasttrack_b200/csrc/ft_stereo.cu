// ft_stereo.cu -- stereo matching on the device-resident frame (sm_100a).
//
//   k_stereo_match     Frame::ComputeStereoMatches, per-keypoint part (reference src/Frame.cc:835-989):
//                      one warp per left keypoint; row-band + octave + disparity filter over the right
//                      keypoints, Hamming argmin with __popc over uint4 descriptor words, then 11x11 SAD
//                      over 11 shifts on the left keypoint's pyramid level and the parabola fit.
//                      The (SAD, iL) median filter at the end of the same function (:991-1004) runs in the last CTA.
//   k_fisheye_match    Frame::ComputeStereoFishEyeMatches (:1231-1271): brute-force 2-NN Hamming with Lowe
//                      ratio, then KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:306-406).
#include "ft_device.cuh"
#include "ft_internal.h"
#include "ft_camera.cuh"

#define ST_WARPS 8
static_assert(ST_WARPS * 32 == 256, "the outlier filter of k_stereo_match keeps one histogram bin per thread");

// per-warp body of k_stereo_match: one left keypoint
struct FtRightKp {   // right keypoint as the row-band scan needs it (staged in shared memory once per CTA)
  float x;
  short minr, maxr;
  int octave;
};

__device__ __forceinline__ void ft_stereo_one(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& s, float mbf,
                                              float mb, int iL, int lane, const FtRightKp* sR) {
  const FtEye& EL = b.eye[0];
  const FtEye& ER = b.eye[1];
  const int nR = ER.counts[0];
  if (lane == 0) { s.uRight[iL] = -1.0f; s.depth[iL] = -1.0f; s.bestIdxR[iL] = -1; s.sad[iL] = -1; }
  const ft_keypoint kpL = EL.kps[iL];
  const int levelL = kpL.octave;
  const float vL = kpL.y, uL = kpL.x;
  const int rowi = (int)vL;
  const int nRows = p.lv[0].h;
  if (rowi < 0 || rowi >= nRows) return;
  const float maxD = __fdiv_rn(mbf, mb);           // mbf / minZ, minZ = mb (:866-868)
  const float minU = __fsub_rn(uL, maxD), maxU = uL;   // minD = 0
  if (maxU < 0) return;
  const uint4* dl = reinterpret_cast<const uint4*>(EL.desc + (size_t)iL * 32);
  const uint4 dl0 = dl[0], dl1 = dl[1];
  // the row table lists candidates in ascending iR and the scan keeps the first minimum (:897-917)
  unsigned best = (100u << 16) | 0xFFFFu;   // TH_HIGH = 100, strict <
  int tested = 0;
  for (int iR = lane; iR < nR; iR += 32) {
    const FtRightKp kpR = sR[iR];
    if (rowi < kpR.minr || rowi > kpR.maxr) continue;
    if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
    if (kpR.x >= minU && kpR.x <= maxU) {
      const uint4* dr = reinterpret_cast<const uint4*>(ER.desc + (size_t)iR * 32);
      const int dist = ft_hamming256(dl0, dl1, dr[0], dr[1]);
      tested++;
      const unsigned key = ((unsigned)dist << 16) | (unsigned)iR;
      best = min(best, key);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    tested += __shfl_xor_sync(0xFFFFFFFFu, tested, o);
  }
  if (lane == 0 && tested) atomicAdd(&s.stats[0], (unsigned long long)tested);
  const int bestDist = (int)(best >> 16);
  if (bestDist >= 75) return;                      // thOrbDist = (TH_HIGH+TH_LOW)/2 (:843)
  const int bestIdxR = (int)(best & 0xFFFFu);
  if (lane == 0) s.bestIdxR[iL] = bestIdxR;
  // sub-pixel refinement by correlation (:921-989)
  const float uR0 = ER.kps[bestIdxR].x;
  const float sf = p.invScale[levelL];
  const float scaleduL = roundf(__fmul_rn(kpL.x, sf)), scaledvL = roundf(__fmul_rn(kpL.y, sf));
  const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
  const int w = 5, Lw = 5;
  const FtLevel& LV = p.lv[levelL];
  const float iniu = scaleduR0 + Lw - w, endu = scaleduR0 + Lw + w + 1;
  if (iniu < 0 || endu >= (float)LV.w) return;
  const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
  if (cy - w < 0 || cy + w >= LV.h || cxl - w < 0 || cxl + w >= LV.w || cxr - Lw - w < 0) return;
  if (lane == 0) atomicAdd(&s.stats[1], 1ull);
  const uint8_t* IL = EL.pyr + LV.offset;
  const uint8_t* IR = ER.pyr + LV.offset;
  int sum[11];
#pragma unroll
  for (int k = 0; k < 11; k++) sum[k] = 0;
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int i = lane + 32 * it;
    if (i < 121) {
      const int yy = i / 11 - w, xx = i % 11 - w;
      const int a = IL[(size_t)(cy + yy) * LV.pitch + cxl + xx];
      const uint8_t* rr = IR + (size_t)(cy + yy) * LV.pitch + cxr + xx - Lw;
#pragma unroll
      for (int k = 0; k < 11; k++) sum[k] += abs(a - (int)rr[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 11; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum[k] += __shfl_xor_sync(0xFFFFFFFFu, sum[k], o);
  }
  if (lane != 0) return;
  int bestD = 0x7FFFFFFF, bestinc = 0;
#pragma unroll
  for (int k = 0; k < 11; k++) {
    if (sum[k] < bestD) { bestD = sum[k]; bestinc = k - Lw; }
  }
  if (bestinc == -Lw || bestinc == Lw) return;
  float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
  for (int k = 1; k < 10; k++) {
    if (k == Lw + bestinc) { d1 = (float)sum[k - 1]; d2 = (float)sum[k]; d3 = (float)sum[k + 1]; }
  }
  const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
  if (deltaR < -1 || deltaR > 1) return;
  float bestuR = __fmul_rn(p.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
  float disparity = __fsub_rn(uL, bestuR);
  if (disparity >= 0 && disparity < maxD) {
    if (disparity <= 0) {
      disparity = 0.01f;                             // float(0.01)
      bestuR = (float)((double)uL - 0.01);           // uL - 0.01 evaluated in double (:982)
    }
    s.depth[iL] = __fdiv_rn(mbf, disparity);
    s.uRight[iL] = bestuR;
    s.sad[iL] = bestD;
  }
}

// Stereo matching with the outlier filter fused in: the last CTA to finish (ticket counter) runs the
// (SAD, iL) median filter of Frame.cc:991-1004. The reference sorts the pairs and reads element size/2; only
// that element's SAD is used, so a two-pass 8+7-bit radix select over shared-memory histograms gives the same
// value (SAD <= 121*255 < 2^15) without sorting.
__global__ void __launch_bounds__(ST_WARPS * 32) k_stereo_match(const __grid_constant__ FtParams p,
                                                                const __grid_constant__ FtBuffers b,
                                                                const __grid_constant__ FtStereoBuffers s, float mbf,
                                                                float mb) {
  extern __shared__ __align__(16) uint8_t sDyn[];
  FtRightKp* sR = reinterpret_cast<FtRightKp*>(sDyn);
  __shared__ int sHist[256];
  __shared__ int sLast, sCount, sBin, sBefore, sMedian;
  const int tid = threadIdx.x, lane = tid & 31;
  FT_PDL_WAIT();        // launched as a programmatic dependent of k_orient_desc
  const int nL = b.eye[0].counts[0], nR = b.eye[1].counts[0];
  const int iL = blockIdx.x * ST_WARPS + (tid >> 5);
  if (blockIdx.x * ST_WARPS < nL) {
    // row band of every right keypoint: rows floor(y-r) .. ceil(y+r), r = 2*scale[octave] (Frame.cc:852-862)
    for (int i = tid; i < nR; i += ST_WARPS * 32) {
      const ft_keypoint k = b.eye[1].kps[i];
      const float r = __fmul_rn(2.0f, p.scale[k.octave]);
      FtRightKp o;
      o.x = k.x; o.octave = k.octave;
      o.maxr = (short)(int)ceilf(__fadd_rn(k.y, r)); o.minr = (short)(int)floorf(__fsub_rn(k.y, r));
      sR[i] = o;
    }
  }
  __syncthreads();
  if (iL < nL) ft_stereo_one(p, b, s, mbf, mb, iL, lane, sR);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(&s.stats[7]), 1u);
    sLast = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!sLast) return;
  __threadfence();
  // ---- outlier filter, 256 threads ----
  // The SADs are fetched from L2 once and kept in registers across the three passes (up to ST_KEEP per thread; frames with
  // more keypoints re-read the tail).
  constexpr int ST_KEEP = 8;
  const int NT = ST_WARPS * 32;
  sHist[tid] = 0;
  if (tid == 0) { sCount = 0; reinterpret_cast<unsigned*>(&s.stats[7])[0] = 0; }
  int keep[ST_KEEP];
#pragma unroll
  for (int k = 0; k < ST_KEEP; k++) { const int i = tid + k * NT; keep[k] = i < nL ? __ldcg(&s.sad[i]) : -1; }
  __syncthreads();
  int local = 0;
#pragma unroll
  for (int k = 0; k < ST_KEEP; k++) if (keep[k] >= 0) { local++; atomicAdd(&sHist[keep[k] >> 7], 1); }
  for (int i = tid + ST_KEEP * NT; i < nL; i += NT) {
    const int v = __ldcg(&s.sad[i]);
    if (v >= 0) { local++; atomicAdd(&sHist[v >> 7], 1); }
  }
  if (local) atomicAdd(&sCount, local);
  __syncthreads();
  const int cnt = sCount;
  if (cnt == 0) return;
  const int target = cnt / 2;   // vDistIdx[vDistIdx.size()/2]
  if (tid < 32) {
    // 8 bins per lane, find the bin holding rank `target`
    int c[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k] = sHist[tid * 8 + k]; sum += c[k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (tid >= o) incl += t; }
    int before = incl - sum;
    if (target >= before && target < incl) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (target >= before && target < before + c[k]) { sBin = tid * 8 + k; sBefore = before; }
        before += c[k];
      }
    }
  }
  __syncthreads();
  const int bin = sBin, rem = target - sBefore;
  if (tid < 128) sHist[tid] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ST_KEEP; k++) if (keep[k] >= 0 && (keep[k] >> 7) == bin) atomicAdd(&sHist[keep[k] & 127], 1);
  for (int i = tid + ST_KEEP * NT; i < nL; i += NT) {
    const int v = __ldcg(&s.sad[i]);
    if (v >= 0 && (v >> 7) == bin) atomicAdd(&sHist[v & 127], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0, lo = 0;
    for (int k = 0; k < 128; k++) { if (rem >= acc && rem < acc + sHist[k]) { lo = k; break; } acc += sHist[k]; }
    sMedian = (bin << 7) | lo;
  }
  __syncthreads();
  const float thDist = __fmul_rn(1.5f * 1.4f, (float)sMedian);
#pragma unroll
  for (int k = 0; k < ST_KEEP; k++) {
    const int i = tid + k * NT;
    if (keep[k] >= 0 && !((float)keep[k] < thDist)) { s.uRight[i] = -1.0f; s.depth[i] = -1.0f; }
  }
  for (int i = tid + ST_KEEP * NT; i < nL; i += NT) {
    const int v = __ldcg(&s.sad[i]);
    if (v >= 0 && !((float)v < thDist)) { s.uRight[i] = -1.0f; s.depth[i] = -1.0f; }
  }
}

// ------------------------------------------------------------------------------------
// Fisheye (KannalaBrandt8) stereo
// ------------------------------------------------------------------------------------
__device__ void ft_kb8_unproject(const FtCamera& c, float u, float v, float ray[3]) {
  // KannalaBrandt8::unproject (KannalaBrandt8.cpp:116-143), precision 1e-6
  const float pwx = __fdiv_rn(__fsub_rn(u, c.p[2]), c.p[0]), pwy = __fdiv_rn(__fsub_rn(v, c.p[3]), c.p[1]);
  float scale = 1.f;
  float theta_d = sqrtf(__fadd_rn(__fmul_rn(pwx, pwx), __fmul_rn(pwy, pwy)));
  const float hp = (float)(3.14159265358979323846 / 2.0);
  theta_d = fminf(fmaxf(-hp, theta_d), hp);
  if (theta_d > 1e-8) {
    float theta = theta_d;
    for (int j = 0; j < 10; j++) {
      const float t2 = __fmul_rn(theta, theta), t4 = __fmul_rn(t2, t2), t6 = __fmul_rn(t4, t2), t8 = __fmul_rn(t4, t4);
      const float k0 = __fmul_rn(c.p[4], t2), k1 = __fmul_rn(c.p[5], t4), k2 = __fmul_rn(c.p[6], t6), k3 = __fmul_rn(c.p[7], t8);
      const float num = __fsub_rn(__fmul_rn(theta, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.f, k0), k1), k2), k3)), theta_d);
      const float den = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.f, __fmul_rn(3.f, k0)), __fmul_rn(5.f, k1)),
                                            __fmul_rn(7.f, k2)), __fmul_rn(9.f, k3));
      const float fix = __fdiv_rn(num, den);
      theta = __fsub_rn(theta, fix);
      if (fabsf(fix) < 1e-6f) break;
    }
    scale = __fdiv_rn(tanf(theta), theta_d);
  }
  ray[0] = __fmul_rn(pwx, scale); ray[1] = __fmul_rn(pwy, scale); ray[2] = 1.f;
}

// smallest right-singular vector of a 4x4 matrix: cyclic Jacobi on A^T A in double
// (stands in for Eigen::JacobiSVD<Matrix4f>, KannalaBrandt8.cpp:403-405).
__device__ void ft_null_vec4(const double A[4][4], double v[4]) {
  double M[4][4], V[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double acc = 0;
      for (int k = 0; k < 4; k++) acc += A[k][i] * A[k][j];
      M[i][j] = acc;
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int i = 0; i < 4; i++)
      for (int j = i + 1; j < 4; j++) off += M[i][j] * M[i][j];
    if (off < 1e-300) break;
    for (int pp = 0; pp < 3; pp++)
      for (int q = pp + 1; q < 4; q++) {
        if (fabs(M[pp][q]) < 1e-300) continue;
        const double theta = (M[q][q] - M[pp][pp]) / (2 * M[pp][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), sn = t * c;
        for (int k = 0; k < 4; k++) {
          const double a = M[k][pp], bq = M[k][q];
          M[k][pp] = c * a - sn * bq; M[k][q] = sn * a + c * bq;
        }
        for (int k = 0; k < 4; k++) {
          const double a = M[pp][k], bq = M[q][k];
          M[pp][k] = c * a - sn * bq; M[q][k] = sn * a + c * bq;
        }
        for (int k = 0; k < 4; k++) {
          const double a = V[k][pp], bq = V[k][q];
          V[k][pp] = c * a - sn * bq; V[k][q] = sn * a + c * bq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; i++) if (M[i][i] < M[best][best]) best = i;
  for (int k = 0; k < 4; k++) v[k] = V[k][best];
}

__device__ float ft_triangulate_matches(const FtCamera& c1, const FtCamera& c2, const ft_keypoint& kp1,
                                        const ft_keypoint& kp2, const float R12[9], const float t12[3], float sigmaLevel,
                                        float unc, float p3D[3]) {
  float r1[3], r2[3], r21[3];
  ft_kb8_unproject(c1, kp1.x, kp1.y, r1);
  ft_kb8_unproject(c2, kp2.x, kp2.y, r2);
  ft_mat3_vec(R12, r2, r21);
  const float n1 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r1[0], r1[0]), __fmul_rn(r1[1], r1[1])), __fmul_rn(r1[2], r1[2])));
  const float n21 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(r21[0], r21[0]), __fmul_rn(r21[1], r21[1])), __fmul_rn(r21[2], r21[2])));
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(r1[0], r21[0]), __fmul_rn(r1[1], r21[1])), __fmul_rn(r1[2], r21[2]));
  const float cosPar = __fdiv_rn(dot, __fmul_rn(n1, n21));
  if ((double)cosPar > 0.9998) return -1;
  float R21[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R21[3 * i + j] = R12[3 * j + i];
  float Rt[3];
  ft_mat3_vec(R21, t12, Rt);
  float T2[3][4];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T2[i][j] = R21[3 * i + j]; T2[i][3] = -Rt[i]; }
  const float T1[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
  double A[4][4];
  for (int j = 0; j < 4; j++) {
    A[0][j] = (double)__fsub_rn(__fmul_rn(r1[0], T1[2][j]), T1[0][j]);
    A[1][j] = (double)__fsub_rn(__fmul_rn(r1[1], T1[2][j]), T1[1][j]);
    A[2][j] = (double)__fsub_rn(__fmul_rn(r2[0], T2[2][j]), T2[0][j]);
    A[3][j] = (double)__fsub_rn(__fmul_rn(r2[1], T2[2][j]), T2[1][j]);
  }
  double vh[4];
  ft_null_vec4(A, vh);
  float x3D[3] = {(float)(vh[0] / vh[3]), (float)(vh[1] / vh[3]), (float)(vh[2] / vh[3])};
  const float z1 = x3D[2];
  if (z1 <= 0) return -2;
  const float z2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R21[6], x3D[0]), __fmul_rn(R21[7], x3D[1])), __fmul_rn(R21[8], x3D[2])), T2[2][3]);
  if (z2 <= 0) return -3;
  float uv1[2];
  ft_cam_project(c1, x3D, uv1);
  const float ex1 = __fsub_rn(uv1[0], kp1.x), ey1 = __fsub_rn(uv1[1], kp1.y);
  if ((double)__fadd_rn(__fmul_rn(ex1, ex1), __fmul_rn(ey1, ey1)) > 5.991 * (double)sigmaLevel) return -4;
  float x3D2[3];
  ft_mat3_vec(R21, x3D, x3D2);
  for (int i = 0; i < 3; i++) x3D2[i] = __fadd_rn(x3D2[i], T2[i][3]);
  float uv2[2];
  ft_cam_project(c2, x3D2, uv2);
  const float ex2 = __fsub_rn(uv2[0], kp2.x), ey2 = __fsub_rn(uv2[1], kp2.y);
  if ((double)__fadd_rn(__fmul_rn(ex2, ex2), __fmul_rn(ey2, ey2)) > 5.991 * (double)unc) return -5;
  p3D[0] = x3D[0]; p3D[1] = x3D[1]; p3D[2] = x3D[2];
  return z1;
}

__global__ void __launch_bounds__(64) k_fisheye_init(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                     const __grid_constant__ FtStereoBuffers s) {
  const int i = blockIdx.x * 64 + threadIdx.x;
  FT_PDL_WAIT();        // launched as a programmatic dependent of k_orient_desc
  if (i < p.maxKp) {
    s.l2r[i] = -1; s.r2l[i] = -1; s.depth[i] = -1.0f; s.uRight[i] = -1.0f; s.code[i] = 0;
    s.p3d[3 * i] = 0; s.p3d[3 * i + 1] = 0; s.p3d[3 * i + 2] = 0;
  }
}

__global__ void __launch_bounds__(ST_WARPS * 32) k_fisheye_match(const __grid_constant__ FtParams p,
                                                                 const __grid_constant__ FtBuffers b,
                                                                 const __grid_constant__ FtStereoBuffers s,
                                                                 const __grid_constant__ FtCamera c1,
                                                                 const __grid_constant__ FtCamera c2,
                                                                 const __grid_constant__ FtPose pose) {
  const FtEye& EL = b.eye[0];
  const FtEye& ER = b.eye[1];
  const int lane = threadIdx.x & 31;
  const int nL = EL.counts[0], nR = ER.counts[0];
  const int monoL = EL.counts[1], monoR = ER.counts[1];
  const int iL = monoL + blockIdx.x * ST_WARPS + (threadIdx.x >> 5);
  if (iL >= nL) return;
  const uint4* dl = reinterpret_cast<const uint4*>(EL.desc + (size_t)iL * 32);
  const uint4 dl0 = dl[0], dl1 = dl[1];
  // knnMatch k=2 over the right lapping subset: ascending distance, ties -> lowest train index
  unsigned b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu;
  for (int iR = monoR + lane; iR < nR; iR += 32) {
    const uint4* dr = reinterpret_cast<const uint4*>(ER.desc + (size_t)iR * 32);
    const unsigned key = ((unsigned)ft_hamming256(dl0, dl1, dr[0], dr[1]) << 16) | (unsigned)(iR - monoR);
    if (key < b0) { b1 = b0; b0 = key; }
    else if (key < b1) b1 = key;
  }
  unsigned g0 = b0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g0 = min(g0, __shfl_xor_sync(0xFFFFFFFFu, g0, o));
  unsigned g1 = (b0 == g0) ? b1 : b0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g1 = min(g1, __shfl_xor_sync(0xFFFFFFFFu, g1, o));
  if (lane != 0) return;
  if (nR - monoR > 0) atomicAdd(&s.stats[0], (unsigned long long)(nR - monoR));
  if (g1 == 0xFFFFFFFFu) return;   // fewer than two neighbours
  const float d0 = (float)(g0 >> 16), d1 = (float)(g1 >> 16);
  if (!((double)d0 < (double)d1 * 0.7)) return;
  const int iR = (int)(g0 & 0xFFFFu) + monoR;
  const ft_keypoint k1 = EL.kps[iL], k2 = ER.kps[iR];
  float p3D[3] = {0, 0, 0};
  atomicAdd(&s.stats[1], 1ull);
  // R12 = Rlr, t12 = tlr (Frame.cc:1260): rays of camera 2 expressed in camera 1
  const float depth = ft_triangulate_matches(c1, c2, k1, k2, pose.Rlr, pose.tlr, p.sigma2[k1.octave], p.sigma2[k2.octave], p3D);
  if (depth > 0.0001f) {
    s.l2r[iL] = iR;
    atomicMax(&s.r2l[iR], iL);   // sequential loop: the last (highest) left index wins
    s.p3d[3 * iL] = p3D[0]; s.p3d[3 * iL + 1] = p3D[1]; s.p3d[3 * iL + 2] = p3D[2];
    s.depth[iL] = depth;
    s.code[iL] = 1;
  } else {
    s.code[iL] = (int)depth == 0 ? -6 : (int)depth;
  }
}

// the right-keypoint table staged by every CTA is 12 bytes per keypoint: above ~4000 features it exceeds the 48 KB a kernel
// gets without opting in
cudaError_t ft_launch_stereo_setup(const FtParams& p) {
  (void)p;
  return ft_set_max_dynamic_smem((const void*)k_stereo_match);   // per function and per device: never lowered by a smaller context
}

void ft_launch_stereo_match(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& s, float mbf, float mb,
                            cudaStream_t st) {
  ft_launch_pdl(k_stereo_match, dim3((p.maxKp + ST_WARPS - 1) / ST_WARPS), dim3(ST_WARPS * 32), sizeof(FtRightKp) * p.maxKp, st,
                p, b, s, mbf, mb);
}
void ft_launch_fisheye(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& s, const FtCamera& c1,
                       const FtCamera& c2, const FtPose& pose, cudaStream_t st) {
  ft_launch_pdl(k_fisheye_init, dim3((p.maxKp + 63) / 64), dim3(64), 0, st, p, b, s);
  k_fisheye_match<<<(p.maxKp + ST_WARPS - 1) / ST_WARPS, ST_WARPS * 32, 0, st>>>(p, b, s, c1, c2, pose);
}

// ---- RGB-D / monocular frames ---------------------------------------------------------------------
// Frame::ComputeStereoFromRGBD (reference src/Frame.cc:1065-1086): d = imDepth.at<float>(kp.pt.y, kp.pt.x) (float
// coordinates truncated to int); d > 0 -> mvDepth = d, mvuRight = kpU.pt.x - mbf / d, else both stay -1. depth ==
// nullptr is the monocular Frame constructor (src/Frame.cc:330-331): everything -1. Runs after the frame grid, which
// writes the undistorted keypoints.
__global__ void k_rgbd_depth(const __grid_constant__ FtBuffers b, const __grid_constant__ FtStereoBuffers st, const float2* kpUn,
                             const float* depth, int pitchFloats, int width, int height, float mbf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.eye[0].counts[0]) return;
  float ur = -1.f, dp = -1.f;
  if (depth) {
    const ft_keypoint kp = b.eye[0].kps[i];
    const int u = (int)kp.x, v = (int)kp.y;
    if (u >= 0 && u < width && v >= 0 && v < height) {
      const float d = depth[(size_t)v * pitchFloats + u];
      if (d > 0) { dp = d; ur = __fsub_rn(kpUn[i].x, __fdiv_rn(mbf, d)); }
    }
  }
  st.uRight[i] = ur; st.depth[i] = dp;
}
void ft_launch_rgbd_depth(const FtParams& p, const FtBuffers& b, const FtStereoBuffers& st, const float2* kpUn, const float* depth,
                          int pitchFloats, float mbf, cudaStream_t s) {
  k_rgbd_depth<<<(p.maxKp + 255) / 256, 256, 0, s>>>(b, st, kpUn, depth, pitchFloats, p.width, p.height, mbf);
}
