// ft_oracle_capi.cpp -- flat C entry points over the CPU ORACLE for ctypes
// (TEST INFRASTRUCTURE ONLY; see ft_oracle.h). Keypoints cross this boundary as
// float[n][6] = {x, y, size, angle, response, octave}.
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "ft_oracle.h"

using namespace fto;

static void kps_to_flat(const std::vector<KeyPoint>& k, float* out) {
  for (size_t i = 0; i < k.size(); i++) {
    out[6 * i] = k[i].x; out[6 * i + 1] = k[i].y; out[6 * i + 2] = k[i].size;
    out[6 * i + 3] = k[i].angle; out[6 * i + 4] = k[i].response; out[6 * i + 5] = (float)k[i].octave;
  }
}
static std::vector<KeyPoint> kps_from_flat(const float* in, int n) {
  std::vector<KeyPoint> k(n);
  for (int i = 0; i < n; i++)
    k[i] = {in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], in[6 * i + 4], (int)in[6 * i + 5]};
  return k;
}

extern "C" {

void* fto_extractor_create(int nfeatures, float scale, int nlevels, int iniTh, int minTh) {
  return new Extractor(nfeatures, scale, nlevels, iniTh, minTh);
}
void fto_extractor_destroy(void* ex) { delete (Extractor*)ex; }

void fto_extractor_tables(void* ex_, float* scale, float* invScale, float* sigma2, float* invSigma2,
                          int* featuresPerLevel, int* umax16) {
  Extractor* ex = (Extractor*)ex_;
  for (int i = 0; i < ex->nlevels; i++) {
    scale[i] = ex->scale[i]; invScale[i] = ex->invScale[i];
    sigma2[i] = ex->sigma2[i]; invSigma2[i] = ex->invSigma2[i];
    featuresPerLevel[i] = ex->featuresPerLevel[i];
  }
  for (int i = 0; i < 16; i++) umax16[i] = ex->umax[i];
}

// returns monoIndex (or -1); *n_out = number of keypoints (<= cap or nothing is written)
int fto_extract(void* ex_, const uint8_t* img, int w, int h, int step, int lap0, int lap1, int cap, float* kps6,
                uint8_t* desc, int* n_out) {
  Extractor* ex = (Extractor*)ex_;
  std::vector<KeyPoint> k;
  std::vector<uint8_t> d;
  const int mono = ex->extract(img, w, h, step, lap0, lap1, k, d);
  *n_out = (int)k.size();
  if ((int)k.size() <= cap) {
    kps_to_flat(k, kps6);
    if (!d.empty()) memcpy(desc, d.data(), d.size());
  }
  return mono;
}

void fto_level_dims(void* ex_, int level, int* w, int* h) {
  Extractor* ex = (Extractor*)ex_;
  *w = ex->pyramid[level].w; *h = ex->pyramid[level].h;
}
// returns 1 if the image exists (blurred levels exist only where keypoints were kept)
int fto_level_image(void* ex_, int level, int blurred, uint8_t* out) {
  Extractor* ex = (Extractor*)ex_;
  const Img& im = blurred ? ex->blurred[level] : ex->pyramid[level];
  if (im.d.empty()) return 0;
  memcpy(out, im.d.data(), im.d.size());
  return 1;
}
int fto_level_candidates(void* ex_, int level, int cap, float* xyr) {
  Extractor* ex = (Extractor*)ex_;
  const auto& c = ex->candidates[level];
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xyr[3 * i] = c[i].x; xyr[3 * i + 1] = c[i].y; xyr[3 * i + 2] = c[i].response; }
  return (int)c.size();
}
int fto_level_keys(void* ex_, int level, int cap, float* kps6, uint8_t* desc) {
  Extractor* ex = (Extractor*)ex_;
  const auto& k = ex->levelKeys[level];
  if ((int)k.size() <= cap) {
    kps_to_flat(k, kps6);
    if (desc && !ex->levelDesc[level].empty()) memcpy(desc, ex->levelDesc[level].data(), ex->levelDesc[level].size());
  }
  return (int)k.size();
}
long fto_desc_borderline(void* ex_) { return ((Extractor*)ex_)->descBorderline; }

// ---- primitives ----
void fto_resize(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  Img s, d;
  s.w = sw; s.h = sh; s.d.assign(src, src + (size_t)sw * sh);
  resize_linear_u8(s, d, dw, dh);
  memcpy(dst, d.d.data(), d.d.size());
}
void fto_blur(const uint8_t* src, int w, int h, uint8_t* dst) {
  Img s, d;
  s.w = w; s.h = h; s.d.assign(src, src + (size_t)w * h);
  gaussian_blur_7x7_s2(s, d);
  memcpy(dst, d.d.data(), d.d.size());
}
// the host libm's cosf / sinf, vectorised (what computeOrbDescriptor calls, ORBextractor.cc:74)
void fto_libm_sincosf(int n, const float* a, float* s, float* c) {
  for (int i = 0; i < n; i++) { s[i] = sinf(a[i]); c[i] = cosf(a[i]); }
}
void fto_stereo_from_rgbd(const float* xy, const float* unx, int n, const float* depth, int w, int h, float mbf, float* ur, float* dp) {
  stereo_from_rgbd(xy, unx, n, depth, w, h, mbf, ur, dp);
}
void fto_ic_angles(void* ex, const uint8_t* img, int w, int h, const float* xy, int n, float* out) {
  ic_angles(*(Extractor*)ex, img, w, h, xy, n, out);
}
void fto_orb_descriptors(const uint8_t* blurred, int w, int h, const float* xy, const float* angle, int n, uint8_t* out) {
  orb_descriptors(blurred, w, h, xy, angle, n, out);
}
void fto_undistort_points(const float* xy, int n, const float* K, const float* dist, int ndist, float* out) {
  undistort_points(xy, n, K, dist, ndist, out);
}
void fto_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* out) {
  image_bounds(cols, rows, K, dist, ndist, out);
}
void fto_remap(const uint8_t* src, int sw, int sh, const float* mapx, const float* mapy, int dw, int dh, uint8_t* dst) {
  Img s, d;
  s.w = sw; s.h = sh; s.d.assign(src, src + (size_t)sw * sh);
  remap_linear_u8(s, mapx, mapy, dw, dh, d);
  memcpy(dst, d.d.data(), d.d.size());
}
int fto_fast(const uint8_t* img, int stride, int w, int h, int th, int cap, float* xyr) {
  std::vector<Candidate> c;
  fast_detect(img, stride, w, h, th, c);
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xyr[3 * i] = c[i].x; xyr[3 * i + 1] = c[i].y; xyr[3 * i + 2] = c[i].response; }
  return (int)c.size();
}
int fto_fast_score(const uint8_t* p, int stride) { return fast_score_9_16(p, stride); }
float fto_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int fto_cv_round(float v) { return cv_round(v); }
void fto_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx2, int* dist2) {
  knn2_hamming(q, nq, t, nt, idx2, dist2);
}
int fto_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

int fto_octree(void* ex_, const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, float* out_xyr) {
  Extractor* ex = (Extractor*)ex_;
  std::vector<Candidate> in(n);
  for (int i = 0; i < n; i++) in[i] = {xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2]};
  std::vector<Candidate> out = ex->distributeOctTree(in, minX, maxX, minY, maxY, N);
  for (size_t i = 0; i < out.size(); i++) { out_xyr[3 * i] = out[i].x; out_xyr[3 * i + 1] = out[i].y; out_xyr[3 * i + 2] = out[i].response; }
  return (int)out.size();
}

// ---- pinhole stereo ----
void fto_stereo(void* exL, void* exR, const float* kL, int nL, const uint8_t* dL, const float* kR, int nR,
                const uint8_t* dR, float mbf, float mb, float* uRight, float* depth, int* bestIdxR, int* sad) {
  StereoResult r;
  std::vector<uint8_t> vdL(dL, dL + (size_t)nL * 32), vdR(dR, dR + (size_t)nR * 32);
  compute_stereo_matches(*(Extractor*)exL, *(Extractor*)exR, kps_from_flat(kL, nL), vdL, kps_from_flat(kR, nR), vdR,
                         mbf, mb, r);
  for (int i = 0; i < nL; i++) { uRight[i] = r.uRight[i]; depth[i] = r.depth[i]; bestIdxR[i] = r.bestIdxR[i]; sad[i] = r.sad[i]; }
}

// ---- fisheye stereo ----
void fto_fisheye(const float* cam1, const float* cam2, const float* Rlr, const float* tlr, const float* sigma2,
                 int nlevels, const float* kL, int nL, const uint8_t* dL, int monoLeft, const float* kR, int nR,
                 const uint8_t* dR, int monoRight, int* l2r, int* r2l, float* depth, float* p3d, int* code) {
  Camera c1, c2;
  c1.type = c2.type = 1;
  memcpy(c1.p, cam1, 32); memcpy(c2.p, cam2, 32);
  std::vector<float> s2(sigma2, sigma2 + nlevels);
  std::vector<uint8_t> vdL(dL, dL + (size_t)nL * 32), vdR(dR, dR + (size_t)nR * 32);
  FisheyeResult r;
  compute_stereo_fisheye(c1, c2, Rlr, tlr, s2, kps_from_flat(kL, nL), vdL, monoLeft, kps_from_flat(kR, nR), vdR,
                         monoRight, r);
  for (int i = 0; i < nL; i++) { l2r[i] = r.l2r[i]; depth[i] = r.depth[i]; code[i] = r.code[i]; }
  memcpy(p3d, r.p3d.data(), sizeof(float) * 3 * nL);
  for (int i = 0; i < nR; i++) r2l[i] = r.r2l[i];
}

void fto_cam_project(int type, const float* p8, const float* P, float* uv) {
  Camera c; c.type = type; memcpy(c.p, p8, 32);
  cam_project(c, P, uv);
}
void fto_kb8_unproject(const float* p8, float u, float v, float* ray) {
  Camera c; c.type = 1; memcpy(c.p, p8, 32);
  kb8_unproject(c, u, v, ray);
}

// ---- frame model + projection search ----
struct fto_frame_desc {
  int Nleft, Nright, N;
  const float* keys6;
  const uint8_t* desc;
  const float* uRight;     // N (pinhole) or null
  const int* l2r;          // Nleft or null
  const int* r2l;          // Nright or null
  float minX, maxX, minY, maxY;
  int nlevels;
  const float* scale;
  float logScale;
  int camType;
  float cam1[8], cam2[8];
  float mbf;
  float Rcw[9], tcw[3], Rwc[9], Ow[3];
  float Rrl[9], trl[3], tlr[3];
};

void* fto_frame_create(const fto_frame_desc* d) {
  FrameModel* F = new FrameModel();
  F->Nleft = d->Nleft; F->Nright = d->Nright; F->N = d->N;
  F->keys = kps_from_flat(d->keys6, d->N);
  F->desc.assign(d->desc, d->desc + (size_t)d->N * 32);
  if (d->uRight) F->uRight.assign(d->uRight, d->uRight + (d->Nleft == -1 ? d->N : d->Nleft));
  if (d->l2r) F->l2r.assign(d->l2r, d->l2r + d->Nleft);
  if (d->r2l) F->r2l.assign(d->r2l, d->r2l + d->Nright);
  F->minX = d->minX; F->maxX = d->maxX; F->minY = d->minY; F->maxY = d->maxY;
  F->gridWInv = 64.0f / (d->maxX - d->minX);
  F->gridHInv = 48.0f / (d->maxY - d->minY);
  F->nlevels = d->nlevels;
  F->scale.assign(d->scale, d->scale + d->nlevels);
  F->logScale = d->logScale;
  F->cam1.type = F->cam2.type = d->camType;
  memcpy(F->cam1.p, d->cam1, 32); memcpy(F->cam2.p, d->cam2, 32);
  F->mbf = d->mbf;
  memcpy(F->Rcw, d->Rcw, 36); memcpy(F->tcw, d->tcw, 12); memcpy(F->Rwc, d->Rwc, 36); memcpy(F->Ow, d->Ow, 12);
  memcpy(F->Rrl, d->Rrl, 36); memcpy(F->trl, d->trl, 12); memcpy(F->tlr, d->tlr, 12);
  F->assignFeaturesToGrid();
  return F;
}
void fto_frame_destroy(void* F) { delete (FrameModel*)F; }

// grid dump: counts[64*48] (index ix*48+iy) and concatenated indices in (ix, iy, insertion) order
int fto_frame_grid(void* F_, int right, int* counts, int* indices) {
  FrameModel* F = (FrameModel*)F_;
  int n = 0;
  for (int ix = 0; ix < 64; ix++)
    for (int iy = 0; iy < 48; iy++) {
      const auto& c = right ? F->gridR[ix][iy] : F->grid[ix][iy];
      counts[ix * 48 + iy] = (int)c.size();
      for (int v : c) indices[n++] = v;
    }
  return n;
}

static std::vector<MapPointIn> mps_from_soa(int M, const float* pos, const float* normal, const float* minmax,
                                            const uint8_t* desc, const int* flags) {
  std::vector<MapPointIn> mps(M);
  for (int i = 0; i < M; i++) {
    memcpy(mps[i].pos, pos + 3 * (size_t)i, 12);
    memcpy(mps[i].normal, normal + 3 * (size_t)i, 12);
    mps[i].minDist = minmax[2 * (size_t)i]; mps[i].maxDist = minmax[2 * (size_t)i + 1];
    memcpy(mps[i].desc, desc + 32 * (size_t)i, 32);
    mps[i].flags = flags[i];
  }
  return mps;
}

// track_i[M][5] = inView, inViewR, level, levelR, borderline
// track_f[M][9] = projX, projY, projXR, depth, viewCos, projXR_r, projYR_r, depthR, viewCosR
static void track_out(const std::vector<MapPointTrack>& tr, int* ti, float* tf) {
  for (size_t i = 0; i < tr.size(); i++) {
    const MapPointTrack& t = tr[i];
    ti[5 * i] = t.inView; ti[5 * i + 1] = t.inViewR; ti[5 * i + 2] = t.level; ti[5 * i + 3] = t.levelR; ti[5 * i + 4] = t.borderline;
    float* f = tf + 9 * i;
    f[0] = t.projX; f[1] = t.projY; f[2] = t.projXR; f[3] = t.depth; f[4] = t.viewCos;
    f[5] = t.projXR_r; f[6] = t.projYR_r; f[7] = t.depthR; f[8] = t.viewCosR;
  }
}

void fto_frustum(void* F_, int M, const float* pos, const float* normal, const float* minmax, const uint8_t* desc,
                 const int* flags, float viewCosLimit, int* track_i, float* track_f) {
  FrameModel* F = (FrameModel*)F_;
  std::vector<MapPointIn> mps = mps_from_soa(M, pos, normal, minmax, desc, flags);
  std::vector<MapPointTrack> tr(M);
  for (int i = 0; i < M; i++) is_in_frustum(*F, mps[i], viewCosLimit, tr[i]);
  track_out(tr, track_i, track_f);
}

// Tracking::SearchLocalPoints loop 2 (isInFrustum, Tracking.cc:3504-3522) followed by
// SearchByProjection (Tracking.cc:3555). holder[N] in/out: -1 empty, -2 pre-existing foreign
// map point, >=0 index into this call's map points. holderObs[N] in/out.
int fto_search_local_points(void* F_, int M, const float* pos, const float* normal, const float* minmax,
                            const uint8_t* desc, const int* flags, float th, int bFar, float thFar, float nnratio,
                            int* holder, uint8_t* holderObs, int* track_i, float* track_f) {
  FrameModel* F = (FrameModel*)F_;
  std::vector<MapPointIn> mps = mps_from_soa(M, pos, normal, minmax, desc, flags);
  std::vector<MapPointTrack> tr(M);
  for (int i = 0; i < M; i++) is_in_frustum(*F, mps[i], 0.5f, tr[i]);
  if (track_i && track_f) track_out(tr, track_i, track_f);
  std::vector<int> h(holder, holder + F->N);
  std::vector<uint8_t> ho(holderObs, holderObs + F->N);
  const int n = search_by_projection(*F, mps, tr, th, bFar != 0, thFar, nnratio, h, ho);
  memcpy(holder, h.data(), sizeof(int) * F->N);
  memcpy(holderObs, ho.data(), F->N);
  return n;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono): the caller passes the last frame's map points
// as SoA arrays (one entry per last-frame keypoint that holds a map point)
int fto_search_last_frame(void* F_, int n, const float* pos, const uint8_t* desc, const int* octave, const float* angle,
                          const int* flags, float th, int direction, int checkOri, int* holder, uint8_t* holderObs,
                          int* borderline) {
  FrameModel* F = (FrameModel*)F_;
  std::vector<LastFramePoint> pts(n);
  for (int i = 0; i < n; i++) {
    memcpy(pts[i].pos, pos + 3 * (size_t)i, 12); memcpy(pts[i].desc, desc + 32 * (size_t)i, 32);
    pts[i].octave = octave[i]; pts[i].angle = angle[i]; pts[i].flags = flags[i];
  }
  std::vector<int> h(holder, holder + F->N), bl;
  std::vector<uint8_t> ho(holderObs, holderObs + F->N);
  const int nm = search_by_projection_last_frame(*F, pts, th, direction, checkOri != 0, 0.f, h, ho, bl);
  memcpy(holder, h.data(), sizeof(int) * F->N); memcpy(holderObs, ho.data(), F->N);
  if (borderline) memcpy(borderline, bl.data(), sizeof(int) * n);
  return nm;
}

// ---- timing helper for bench.py's cpu_baseline / --impl reference legs ----
// One stereo frame as the reference threads it: L/R extraction on two std::threads
// (Frame.cc:127-130), then stereo matching on the caller's thread. Returns milliseconds.
double fto_time_stereo_frame(void* exL_, void* exR_, const uint8_t* imgL, const uint8_t* imgR, int w, int h, int step,
                             float mbf, float mb, int two_threads, int* nL_out, int* nR_out, int* nStereo_out) {
  Extractor* exL = (Extractor*)exL_;
  Extractor* exR = (Extractor*)exR_;
  std::vector<KeyPoint> kL, kR;
  std::vector<uint8_t> dL, dR;
  auto t0 = std::chrono::steady_clock::now();
  if (two_threads) {
    std::thread tl([&] { exL->extract(imgL, w, h, step, 0, 0, kL, dL); });
    std::thread trr([&] { exR->extract(imgR, w, h, step, 0, 0, kR, dR); });
    tl.join(); trr.join();
  } else {
    exL->extract(imgL, w, h, step, 0, 0, kL, dL);
    exR->extract(imgR, w, h, step, 0, 0, kR, dR);
  }
  StereoResult r;
  compute_stereo_matches(*exL, *exR, kL, dL, kR, dR, mbf, mb, r);
  auto t1 = std::chrono::steady_clock::now();
  int ns = 0;
  for (float d : r.depth) ns += d > 0;
  if (nL_out) *nL_out = (int)kL.size();
  if (nR_out) *nR_out = (int)kR.size();
  if (nStereo_out) *nStereo_out = ns;
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

}  // extern "C"
