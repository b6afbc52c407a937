// ref_frame_capi.cpp -- C entry points over FUNCTIONS OF THE REFERENCE compiled from their own text (TEST INFRASTRUCTURE
// ONLY). oracle/_ref/gen_frame_fns.inc is written at build time by oracle/ref_extract_fns.py: the verbatim text of
//   Frame::{AssignFeaturesToGrid, PosInGrid, GetFeaturesInArea, isInFrustum, isInFrustumChecks, ComputeStereoMatches,
//          ComputeStereoFishEyeMatches, ComputeStereoFromRGBD, UndistortKeyPoints, ComputeImageBounds}, MapPoint::{Get*DistanceInvariance, PredictScale},
//   Pinhole::project, KannalaBrandt8::{project, unproject, unprojectEig, TriangulateMatches, Triangulate},
//   ORBmatcher::{SearchByProjection (local map), SearchByProjection (last frame), SearchByBoW, RadiusByViewingCos,
//               ComputeThreeMaxima, DescriptorDistance}
// taken from the reference files where they lie; it is compiled against ref_stubs/ref_frame_shim.h. The entry points
// below fill the stand-in Frame / MapPoint / KeyFrame objects from flat arrays (the oracle's conventions) and call them.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ref_frame_shim.h"

bool KernelController::searchLocalPointsKernelRunStatus = false;
bool KernelController::poseEstimationKernelRunStatus = false;
void KernelController::launchSearchLocalPointsKernel(ORB_SLAM3::Frame&, const vector<ORB_SLAM3::MapPoint*>&, const float, const bool,
                                                     const float, int*, int*, int*, int*, int*, int*, int*, int*, int*, int*) { abort(); }
void KernelController::launchPoseEstimationKernel(ORB_SLAM3::Frame&, const ORB_SLAM3::Frame&, const float, const bool, const bool,
                                                  Eigen::Matrix4f, int*, int*, int*, int*) { abort(); }

namespace ORB_SLAM3 {
const int ORBmatcher::TH_HIGH = 100;      // src/ORBmatcher.cc:41-43
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
}  // namespace ORB_SLAM3

#include "_ref/gen_frame_fns.inc"

using namespace ORB_SLAM3;

namespace {

struct RefFrame {
  Frame F;
  ORBextractor exL, exR;
  Pinhole pin1, pin2;
  KannalaBrandt8 kb1, kb2;
  std::vector<std::vector<unsigned char>> pyrL, pyrR;
};

std::vector<cv::KeyPoint> keys_from(const float* k6, int n) {
  std::vector<cv::KeyPoint> v(n);
  for (int i = 0; i < n; i++) {
    const float* k = k6 + 6 * (size_t)i;
    v[i] = cv::KeyPoint(k[0], k[1], k[2], k[3], k[4], (int)k[5], -1);
  }
  return v;
}
cv::Mat desc_from(const unsigned char* d, int n) {
  cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
  if (n > 0) memcpy(m.data, d, (size_t)n * 32);
  m.rows = n;
  return m;
}
void set3(Eigen::Vector3f& v, const float* p) { v(0) = p[0]; v(1) = p[1]; v(2) = p[2]; }
void set9(Eigen::Matrix3f& m, const float* p) { for (int i = 0; i < 9; i++) m.m[i] = p[i]; }

}  // namespace

extern "C" {

// the same description of a frame as the oracle's FrameDesc (oracle/ft_oracle_capi.cpp)
struct FtRefFrameDesc {
  int Nleft, Nright, N;
  const float* keys6; const unsigned char* desc; const float* uRight; const int* l2r; const int* r2l;
  float minX, maxX, minY, maxY;
  int nlevels; const float* scale; float logScale;
  int camType; float cam1[8], cam2[8];
  float mbf;
  float Rcw[9], tcw[3], Rwc[9], Ow[3];
  float Rrl[9], trl[3], tlr[3];
  const float* keysUn6;   // mvKeysUn (pinhole), may be NULL = keys6
};

void* ftref_frame_create(const FtRefFrameDesc* d) {
  RefFrame* r = new RefFrame();
  Frame& F = r->F;
  F.N = d->N; F.Nleft = d->Nleft; F.Nright = d->Nright;
  const bool fisheye = d->Nleft != -1;
  const int nl = fisheye ? d->Nleft : d->N;
  F.mvKeys = keys_from(d->keys6, nl);
  if (fisheye) F.mvKeysRight = keys_from(d->keys6 + 6 * (size_t)nl, d->Nright);
  F.mvKeysUn = d->keysUn6 ? keys_from(d->keysUn6, nl) : F.mvKeys;
  F.mDescriptors = desc_from(d->desc, d->N);
  F.mvuRight.assign(nl, -1.f);
  if (d->uRight) F.mvuRight.assign(d->uRight, d->uRight + nl);
  F.mvDepth.assign(nl, -1.f);
  if (fisheye) { F.mvLeftToRightMatch.assign(d->l2r, d->l2r + nl); F.mvRightToLeftMatch.assign(d->r2l, d->r2l + d->Nright); }
  F.mvpMapPoints.assign(d->N, nullptr);
  F.mvbOutlier.assign(d->N, false);
  F.mnMinX = d->minX; F.mnMaxX = d->maxX; F.mnMinY = d->minY; F.mnMaxY = d->maxY;
  F.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(d->maxX - d->minX);   // Frame.cc:190-191
  F.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(d->maxY - d->minY);
  F.mnScaleLevels = d->nlevels; F.mfLogScaleFactor = d->logScale;
  F.mvScaleFactors.assign(d->scale, d->scale + d->nlevels);
  F.mvInvScaleFactors.resize(d->nlevels);
  for (int i = 0; i < d->nlevels; i++) F.mvInvScaleFactors[i] = 1.0f / F.mvScaleFactors[i];
  F.mbf = d->mbf;
  GeometricCamera *c1 = d->camType ? (GeometricCamera*)&r->kb1 : (GeometricCamera*)&r->pin1;
  GeometricCamera *c2 = d->camType ? (GeometricCamera*)&r->kb2 : (GeometricCamera*)&r->pin2;
  c1->mvParameters.assign(d->cam1, d->cam1 + 8); c2->mvParameters.assign(d->cam2, d->cam2 + 8);
  F.mpCamera = c1; F.mpCamera2 = fisheye ? c2 : nullptr;
  set9(F.mRcw, d->Rcw); set3(F.mtcw, d->tcw); set9(F.mRwc, d->Rwc); set3(F.mOw, d->Ow);
  F.mTcw = Sophus::SE3f(F.mRcw, F.mtcw);
  Eigen::Matrix3f Rrl; set9(Rrl, d->Rrl);
  Eigen::Vector3f trl, tlr; set3(trl, d->trl); set3(tlr, d->tlr);
  F.mTrl = Sophus::SE3f(Rrl, trl);
  F.mTlr = Sophus::SE3f(Rrl.transpose(), tlr);
  F.AssignFeaturesToGrid();
  return r;
}
void ftref_frame_destroy(void* r) { delete static_cast<RefFrame*>(r); }

// mGrid / mGridRight: counts[64*48] (ix*48+iy) and the concatenated indices
int ftref_frame_grid(void* r_, int right, int* counts, int* indices) {
  Frame& F = static_cast<RefFrame*>(r_)->F;
  int k = 0;
  for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) {
    const std::vector<std::size_t>& c = right ? F.mGridRight[i][j] : F.mGrid[i][j];
    counts[i * FRAME_GRID_ROWS + j] = (int)c.size();
    for (std::size_t v : c) indices[k++] = (int)v;
  }
  return k;
}

int ftref_features_in_area(void* r_, float x, float y, float rad, int minLevel, int maxLevel, int right, int* out, int cap) {
  Frame& F = static_cast<RefFrame*>(r_)->F;
  const vector<size_t> v = F.GetFeaturesInArea(x, y, rad, minLevel, maxLevel, right != 0);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
  return (int)v.size();
}

static void make_map_points(int M, const float* pos, const float* normal, const float* minmax, const unsigned char* desc,
                            const int* flags, std::vector<MapPoint>& mps) {
  mps = std::vector<MapPoint>(M);
  for (int i = 0; i < M; i++) {
    MapPoint& p = mps[i];
    set3(p.mWorldPos, pos + 3 * (size_t)i);
    if (normal) set3(p.mNormalVector, normal + 3 * (size_t)i);
    if (minmax) { p.mfMinDistance = minmax[2 * (size_t)i]; p.mfMaxDistance = minmax[2 * (size_t)i + 1]; }
    p.mDescriptor = desc_from(desc + 32 * (size_t)i, 1);
    p.mbBad = false; p.nObs = (flags[i] & 2) ? 3 : 0;
  }
}

// Tracking::SearchLocalPoints: the isInFrustum loop (src/Tracking.cc:3504-3522) + SearchByProjection(F, vpMapPoints, th,
// bFarPoints, thFarPoints). flags bit0 = skipped by the loop (bad / already seen in this frame), bit1 = Observations() > 0.
// holder / holderObs as in the oracle. track_i[M][4], track_f[M][9] = the mTrack* scratch.
int ftref_search_local_points(void* r_, int M, const float* pos, const float* normal, const float* minmax, const unsigned char* desc,
                              const int* flags, float th, int bFar, float thFar, float nnratio, int* holder, unsigned char* holderObs,
                              int* track_i, float* track_f) {
  Frame& F = static_cast<RefFrame*>(r_)->F;
  std::vector<MapPoint> mps;
  make_map_points(M, pos, normal, minmax, desc, flags, mps);
  MapPoint foreignObs, foreignNoObs;
  foreignObs.nObs = 3; foreignNoObs.nObs = 0;
  for (int i = 0; i < F.N; i++)
    F.mvpMapPoints[i] = holder[i] == -1 ? nullptr : (holder[i] >= 0 ? &mps[holder[i]] : (holderObs[i] ? &foreignObs : &foreignNoObs));
  std::vector<MapPoint*> vp;
  for (int i = 0; i < M; i++) {
    MapPoint* p = &mps[i];
    if (flags[i] & 1) { p->mbTrackInView = false; p->mbTrackInViewR = false; vp.push_back(p); continue; }   // `continue` of the loop
    F.isInFrustum(p, 0.5);
    vp.push_back(p);
  }
  if (track_i && track_f)
    for (int i = 0; i < M; i++) {
      const MapPoint& p = mps[i];
      int* ti = track_i + 4 * (size_t)i; float* tf = track_f + 9 * (size_t)i;
      ti[0] = p.mbTrackInView; ti[1] = p.mbTrackInViewR; ti[2] = p.mnTrackScaleLevel; ti[3] = p.mnTrackScaleLevelR;
      tf[0] = p.mTrackProjX; tf[1] = p.mTrackProjY; tf[2] = p.mTrackProjXR; tf[3] = p.mTrackDepth; tf[4] = p.mTrackViewCos;
      tf[5] = p.mTrackProjXR; tf[6] = p.mTrackProjYR; tf[7] = p.mTrackDepthR; tf[8] = p.mTrackViewCosR;
    }
  ORBmatcher matcher(nnratio);
  const int nm = matcher.SearchByProjection(F, vp, th, bFar != 0, thFar);
  for (int i = 0; i < F.N; i++) {
    MapPoint* q = F.mvpMapPoints[i];
    if (!q) { holder[i] = -1; holderObs[i] = 0; }
    else if (q == &foreignObs || q == &foreignNoObs) { holder[i] = -2; holderObs[i] = q == &foreignObs; }
    else { holder[i] = (int)(q - mps.data()); holderObs[i] = q->nObs > 0; }
  }
  return nm;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono): one last-frame keypoint per entry
int ftref_search_last_frame(void* r_, int n, const float* pos, const unsigned char* desc, const int* octave, const float* angle,
                            const int* flags, const float* Rlw, const float* tlw, float th, int bMono, int checkOri, float mb,
                            int* holder, unsigned char* holderObs) {
  Frame& F = static_cast<RefFrame*>(r_)->F;
  F.mb = mb;
  std::vector<MapPoint> mps;
  make_map_points(n, pos, nullptr, nullptr, desc, flags, mps);
  Frame Last;
  Last.N = n; Last.Nleft = -1;
  Last.mvKeys.resize(n); Last.mvKeysUn.resize(n);
  Last.mvpMapPoints.assign(n, nullptr); Last.mvbOutlier.assign(n, false);
  for (int i = 0; i < n; i++) {
    Last.mvKeys[i].octave = octave[i]; Last.mvKeys[i].angle = angle[i]; Last.mvKeysUn[i] = Last.mvKeys[i];
    if (!(flags[i] & 1)) Last.mvpMapPoints[i] = &mps[i];
  }
  Eigen::Matrix3f R; set9(R, Rlw); Eigen::Vector3f t; set3(t, tlw);
  Last.mTcw = Sophus::SE3f(R, t);
  MapPoint foreignObs, foreignNoObs;
  foreignObs.nObs = 3; foreignNoObs.nObs = 0;
  for (int i = 0; i < F.N; i++)
    F.mvpMapPoints[i] = holder[i] == -1 ? nullptr : (holder[i] >= 0 ? &mps[holder[i]] : (holderObs[i] ? &foreignObs : &foreignNoObs));
  ORBmatcher matcher(0.9f, checkOri != 0);
  const int nm = matcher.SearchByProjection(F, Last, th, bMono != 0);
  for (int i = 0; i < F.N; i++) {
    MapPoint* q = F.mvpMapPoints[i];
    if (!q) { holder[i] = -1; holderObs[i] = 0; }
    else if (q == &foreignObs || q == &foreignNoObs) { holder[i] = -2; holderObs[i] = q == &foreignObs; }
    else { holder[i] = (int)(q - mps.data()); holderObs[i] = q->nObs > 0; }
  }
  return nm;
}

// Frame::ComputeStereoMatches on two pyramids: levels given tight (w[l] x h[l]) back to back
void ftref_stereo_matches(int nlevels, const int* lw, const int* lh, const unsigned char* pyrL, const unsigned char* pyrR,
                          const float* scale, const float* keysL6, const unsigned char* descL, int nL, const float* keysR6,
                          const unsigned char* descR, int nR, float mbf, float mb, float* uRight, float* depth) {
  RefFrame r;
  Frame& F = r.F;
  size_t off = 0;
  for (int l = 0; l < nlevels; l++) {
    r.exL.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, (void*)(pyrL + off), (size_t)lw[l]));
    r.exR.mvImagePyramid.push_back(cv::Mat(lh[l], lw[l], CV_8U, (void*)(pyrR + off), (size_t)lw[l]));
    off += (size_t)lw[l] * lh[l];
  }
  F.mpORBextractorLeft = &r.exL; F.mpORBextractorRight = &r.exR;
  F.N = nL;
  F.mvKeys = keys_from(keysL6, nL); F.mvKeysRight = keys_from(keysR6, nR);
  F.mDescriptors = desc_from(descL, nL); F.mDescriptorsRight = desc_from(descR, nR);
  F.mvScaleFactors.assign(scale, scale + nlevels);
  F.mvInvScaleFactors.resize(nlevels);
  for (int i = 0; i < nlevels; i++) F.mvInvScaleFactors[i] = 1.0f / F.mvScaleFactors[i];
  F.mbf = mbf; F.mb = mb;
  F.ComputeStereoMatches();
  memcpy(uRight, F.mvuRight.data(), sizeof(float) * nL);
  memcpy(depth, F.mvDepth.data(), sizeof(float) * nL);
}

// Frame::ComputeStereoFishEyeMatches (+ KannalaBrandt8::TriangulateMatches / Triangulate / unproject / project)
void ftref_stereo_fisheye(const float* cam1, const float* cam2, const float* Rlr, const float* tlr, const float* sigma2, int nlevels,
                          const float* keysL6, const unsigned char* descL, int nL, int monoLeft, const float* keysR6,
                          const unsigned char* descR, int nR, int monoRight, int* l2r, int* r2l, float* depth, float* p3d) {
  RefFrame r;
  Frame& F = r.F;
  r.kb1.mvParameters.assign(cam1, cam1 + 8); r.kb2.mvParameters.assign(cam2, cam2 + 8);
  F.mpCamera = &r.kb1; F.mpCamera2 = &r.kb2;
  F.Nleft = nL; F.Nright = nR; F.N = nL + nR;
  F.monoLeft = monoLeft; F.monoRight = monoRight;
  F.mvKeys = keys_from(keysL6, nL); F.mvKeysRight = keys_from(keysR6, nR);
  F.mDescriptors = desc_from(descL, nL); F.mDescriptorsRight = desc_from(descR, nR);
  F.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
  set9(F.mRlr, Rlr); set3(F.mtlr, tlr);
  F.ComputeStereoFishEyeMatches();
  for (int i = 0; i < nL; i++) {
    l2r[i] = F.mvLeftToRightMatch[i]; depth[i] = F.mvDepth[i];
    for (int k = 0; k < 3; k++) p3d[3 * i + k] = F.mvStereo3Dpoints[i](k);
  }
  for (int i = 0; i < nR; i++) r2l[i] = F.mvRightToLeftMatch[i];
}

// Frame::UndistortKeyPoints (mvKeysUn) and Frame::ComputeImageBounds (mnMinX, mnMaxX, mnMinY, mnMaxY)
void ftref_undistort(const float* keysXY, int n, int cols, int rows, const float* K4, const float* dist, int ndist, float* outXY,
                     float* bounds4) {
  RefFrame r;
  Frame& F = r.F;
  r.pin1.mvParameters.assign(K4, K4 + 4);
  F.mpCamera = &r.pin1;
  F.mK = r.pin1.toK();
  F.mDistCoef = cv::Mat(ndist > 0 ? ndist : 1, 1, CV_32F);
  for (int i = 0; i < ndist; i++) F.mDistCoef.at<float>(i, 0) = dist[i];
  F.N = n;
  F.mvKeys.resize(n);
  for (int i = 0; i < n; i++) { F.mvKeys[i].pt.x = keysXY[2 * i]; F.mvKeys[i].pt.y = keysXY[2 * i + 1]; }
  F.UndistortKeyPoints();
  for (int i = 0; i < n; i++) { outXY[2 * i] = F.mvKeysUn[i].pt.x; outXY[2 * i + 1] = F.mvKeysUn[i].pt.y; }
  cv::Mat im(rows, cols, CV_8U);
  F.ComputeImageBounds(im);
  bounds4[0] = F.mnMinX; bounds4[1] = F.mnMaxX; bounds4[2] = F.mnMinY; bounds4[3] = F.mnMaxY;
}

// Frame::ComputeStereoFromRGBD
void ftref_stereo_from_rgbd(const float* keysXY, const float* keysUnX, int n, const float* depthImg, int w, int h, float mbf,
                            float* uRight, float* depth) {
  Frame F;
  F.N = n; F.mbf = mbf;
  F.mvKeys.resize(n); F.mvKeysUn.resize(n);
  for (int i = 0; i < n; i++) { F.mvKeys[i].pt.x = keysXY[2 * i]; F.mvKeys[i].pt.y = keysXY[2 * i + 1]; F.mvKeysUn[i].pt.x = keysUnX[i]; }
  cv::Mat im(h, w, CV_32F, (void*)depthImg, (size_t)w * 4);
  F.ComputeStereoFromRGBD(im);
  memcpy(uRight, F.mvuRight.data(), sizeof(float) * n); memcpy(depth, F.mvDepth.data(), sizeof(float) * n);
}

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) with FeatureVectors given as one node per feature (-1 = none)
int ftref_search_by_bow(int nKF, const unsigned char* kfDesc, const float* kfAngle, const int* kfNode, const unsigned char* kfHasMp,
                        int nF, const unsigned char* fDesc, const float* fAngle, const int* fNode, int fNleft, float nnratio,
                        int checkOri, int* match) {
  Frame F;
  KeyFrame KF;
  Pinhole cam2;
  F.N = nF; F.Nleft = fNleft;
  const int nl = fNleft == -1 ? nF : fNleft;
  F.mvKeys.resize(nl); F.mvKeysRight.resize(nF - nl);
  for (int i = 0; i < nF; i++) (i < nl ? F.mvKeys[i] : F.mvKeysRight[i - nl]).angle = fAngle[i];
  F.mDescriptors = desc_from(fDesc, nF);
  for (int i = 0; i < nF; i++) if (fNode[i] >= 0) F.mFeatVec.addFeature((unsigned)fNode[i], (unsigned)i);
  if (fNleft != -1) { F.mpCamera2 = &cam2; KF.mpCamera2 = &cam2; KF.NLeft = nKF; }
  std::vector<MapPoint> mps(nKF);
  KF.mvpMapPoints.assign(nKF, nullptr);
  KF.mvKeys.resize(nKF); KF.mvKeysUn.resize(nKF);
  for (int i = 0; i < nKF; i++) {
    KF.mvKeys[i].angle = kfAngle[i]; KF.mvKeysUn[i].angle = kfAngle[i];
    if (kfHasMp[i]) KF.mvpMapPoints[i] = &mps[i];
    if (kfNode[i] >= 0) KF.mFeatVec.addFeature((unsigned)kfNode[i], (unsigned)i);
  }
  KF.mDescriptors = desc_from(kfDesc, nKF);
  ORBmatcher matcher(nnratio, checkOri != 0);
  std::vector<MapPoint*> vp;
  const int nm = matcher.SearchByBoW(&KF, F, vp);
  for (int i = 0; i < nF; i++) match[i] = vp[i] ? (int)(vp[i] - mps.data()) : -1;
  return nm;
}

}  // extern "C"
