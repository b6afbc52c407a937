// ft_context.cu -- C ABI of the front-end (include/fasttrack_b200.h): context, device memory,
// coefficient tables, CUDA-graph capture of the per-frame launch chain, host<->device transfers.
//
// Host-side constants are computed exactly the way ORBextractor's constructor does
// (reference src/ORBextractor.cc:393-499): running float32 scale products, cvRound'ed level
// sizes and per-level quotas, umax.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ft_internal.h"
#include "ft_camera.cuh"

static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
void ft_internal_set_err(const char* msg) { g_err = msg ? msg : ""; }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
      return FT_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Persistent device-side MapPoint store: rows of pos/normal/minmax/descriptor that outlive a
// frame, shared by the contexts of one sequence. Updates and searches are ordered across the contexts' streams by events.
struct ft_map_store {
  int cap = 0, device = 0;
  float *pos = nullptr, *normal = nullptr, *minmax = nullptr;
  uint8_t* desc = nullptr;
  std::vector<ft_context*> users;     // attached contexts: a search waits for the last update of each (ft_context::storeUpdated),
                                      // an update for the last search (storeSearched) and the last update of each
  std::mutex mu;                      // the mapping side updates while a tracking context searches: guards `users` and the
                                      // event bookkeeping of the attached contexts
};

struct ft_context {
  ft_config cfg;
  FtParams P;
  FtBuffers B;
  FtStereoBuffers S;
  FtGridBuffers G;
  FtSbpBuffers Q;
  FtCamera cam1, cam2;
  FtPose pose;
  FtUndistort und = {};   // pinhole distortion (Frame::UndistortKeyPoints); off by default
  float mbf, mb;
  float minX, maxX, minY, maxY, gridWInv, gridHInv, logScale;
  int fisheye;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> quota;
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;
  cudaStream_t lvStream[FT_MAX_LEVELS] = {};   // one branch per pyramid level in the captured graph
  cudaEvent_t lvReady[FT_MAX_LEVELS] = {}, lvDone[FT_MAX_LEVELS] = {};
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evFork2 = nullptr, evJoin2 = nullptr, evPyr = nullptr, evJoin3 = nullptr;
  uint8_t* hIn[2] = {nullptr, nullptr};     // pinned host staging
  int* hCounts = nullptr;                   // pinned: nL, monoL, nR, monoR, status, sbp cursor[4]
  // frame result slab: header | kpsL | kpsR | descL | descR | uRight | depth | l2r | r2l | p3d, one D2H brings it back
  uint8_t* dFrame = nullptr; uint8_t* hFrame = nullptr;
  size_t frameBytesPinhole = 0, frameBytesAll = 0;
  size_t offKps[2] = {0, 0}, offDesc[2] = {0, 0}, offUR = 0, offDepth = 0, offL2R = 0, offR2L = 0, offP3D = 0;
  // map-point staging: holderInit | holderObsInit | pos | normal | minmax | desc | flags (tight for the call's M), one H2D
  uint8_t* dMp = nullptr; uint8_t* hMp = nullptr;
  size_t mpHolderBytes = 0;
  // search result slab: header(cursor[8], status) | holder | holderObs | sel, one D2H
  uint8_t* dOut = nullptr; uint8_t* hOut = nullptr;
  size_t offOutHolder = 0, offOutObs = 0, offOutSel = 0;
  cudaGraphExec_t gExtract = nullptr, gStereo = nullptr, gFrame = nullptr;
  FtSearchGraph gSearch;   // gather -> resolve, parameters refreshed per call
  int useGraph = 1;
  int searchGraph = 1;     // FT_SEARCH_GRAPH=0: direct launches of gather / resolve (A/B runs)
  bool extracted = false, stereoDone = false, countsValid = false, framePending = false;
  int frameStatus = 0;   // device status bits seen for the current frame (latched until the next extraction)
  int lastM = 0;
  int nLaunchExtract = 0, nLaunchStereo = 0, nLaunchSearch = 0;
  long long pyrBytes = 0;
  int octDenseTotal = 0;   // dense octree cells per eye (all levels)
  // stereo rectification (optional): raw images are uploaded to dRaw and remapped into level 0 by k_remap
  int rectify = 0, rawW = 0, rawH = 0;
  size_t rawCap = 0;   // pixels the raw-input buffers hold
  int inResize = 0;   // cv::resize of the raw input into level 0 (Settings::needToResize)
  int grouped = 0;    // launch topology: 0 = one graph branch per pyramid level (lowest latency), 1 = levels 1.. grouped (fewest launches)
  // monocular / RGB-D sensors: only eye 0 is extracted; the depth image replaces stereo matching
  int sensor = FT_SENSOR_STEREO;
  float* dDepth = nullptr; float* hDepth = nullptr;
  uint8_t* dRaw[2] = {nullptr, nullptr};
  int2* dRemapTab = nullptr;
  // resident map-point snapshot + initial holders for ft_search_resident
  int residentM = 0;
  int* holderInit = nullptr;
  uint8_t* holderObsInit = nullptr;
  // persistent map store (optional)
  ft_map_store* store = nullptr;
  cudaEvent_t storeSearched = nullptr, updStaged = nullptr, storeUpdated = nullptr;
  bool storeSearchedValid = false, updStagedValid = false, storeUpdatedValid = false;
  uint8_t *hUpd = nullptr, *dUpd = nullptr;
  int updCap = 0;
  // asynchronous halves of ft_search_store: mpStaged = the H2D that reads hMp has finished (the staging buffer may be
  // rewritten), searchDone = the D2H of the result slab has landed in hOut
  cudaEvent_t mpStaged = nullptr, searchDone = nullptr;
  bool mpStagedValid = false, searchPending = false, searchWantBest = false, searchEmpty = false;
  int searchN = 0;
  // bag of words (ft_bow.cu): BowVector / FeatureVector of the current frame, SearchByBoW scratch; allocated on first use
  FtBowFrame W = {};
  FtBowSearch WQ = {};
  int bowCap = 0;
  bool bowValid = false;
  int nLaunchBow = 0;
  uint8_t* hBow = nullptr;   // pinned staging of ft_search_by_bow
  size_t hBowBytes = 0;
  // per-stage CUDA-event timing (direct-launch mode only)
  int timing = 0;
  cudaEvent_t evA[FT_STAGE_COUNT] = {}, evB[FT_STAGE_COUNT] = {};
  bool stageUsed[FT_STAGE_COUNT] = {};
};

struct StageScope {
  ft_context* c; int id; cudaStream_t s;
  StageScope(ft_context* c_, int id_, cudaStream_t s_) : c(c_), id(id_), s(s_) {
    if (c->timing) { cudaEventRecord(c->evA[id], s); c->stageUsed[id] = true; }
  }
  ~StageScope() { if (c->timing) cudaEventRecord(c->evB[id], s); }
};

extern "C" const char* ft_last_error(void) { return g_err.c_str(); }
extern "C" const char* ft_version(void) { return "fasttrack_b200 0.1 (sm_100a)"; }

template <typename T>
static cudaError_t dalloc(ft_context* c, T** p, size_t n) {
  void* v = nullptr;
  cudaError_t e = cudaMalloc(&v, n * sizeof(T) + 256);
  if (e == cudaSuccess) {
    c->allocs.push_back(v); *p = (T*)v;
    // the memset runs on the legacy default stream, which the context's non-blocking streams do not wait for: finish it
    // here so that no later (lazy) allocation can be zeroed after a kernel has already written into it
    e = cudaMemset(v, 0, n * sizeof(T) + 256);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
  }
  return e;
}

static void dfree(ft_context* c, void* p) {
  if (!p) return;
  for (size_t i = 0; i < c->allocs.size(); i++)
    if (c->allocs[i] == p) { c->allocs.erase(c->allocs.begin() + i); break; }
  cudaFree(p);
}

static ft_status build_params(ft_context* c) {
  const ft_config& cfg = c->cfg;
  FtParams& P = c->P;
  memset(&P, 0, sizeof(P));
  const int nl = cfg.nlevels;
  if (nl < 1 || nl > FT_MAX_LEVELS || cfg.width < 64 || cfg.height < 64 || cfg.width > 4000 || cfg.height > 4000 ||
      cfg.nfeatures < 1 || cfg.nfeatures > 10000 || cfg.scale_factor <= 1.0f || cfg.min_th_fast < 1 ||
      cfg.ini_th_fast < cfg.min_th_fast || cfg.ini_th_fast > 254) {
    set_err("ft_context_create: unsupported configuration");
    return FT_ERR_INVALID;
  }
  P.nlevels = nl; P.width = cfg.width; P.height = cfg.height; P.nfeatures = cfg.nfeatures;
  P.iniTh = cfg.ini_th_fast; P.minTh = cfg.min_th_fast; P.camType = cfg.camera_type;
  P.nEyes = 2;
  P.lap[0][0] = cfg.lap_left[0]; P.lap[0][1] = cfg.lap_left[1];
  P.lap[1][0] = cfg.lap_right[0]; P.lap[1][1] = cfg.lap_right[1];
  // scale tables (ORBextractor.cc:398-411,444-450); scaleFactor is held as double there
  const double sfD = (double)cfg.scale_factor;
  c->scale.assign(nl, 1.f); c->sigma2.assign(nl, 1.f); c->invScale.assign(nl, 1.f); c->invSigma2.assign(nl, 1.f);
  for (int i = 1; i < nl; i++) {
    c->scale[i] = (float)(c->scale[i - 1] * sfD);
    c->sigma2[i] = c->scale[i] * c->scale[i];
  }
  for (int i = 0; i < nl; i++) { c->invScale[i] = 1.0f / c->scale[i]; c->invSigma2[i] = 1.0f / c->sigma2[i]; }
  for (int i = 0; i < nl; i++) { P.scale[i] = c->scale[i]; P.invScale[i] = c->invScale[i]; P.sigma2[i] = c->sigma2[i]; }
  // quotas (:454-465)
  c->quota.assign(nl, 0);
  {
    const float factor = (float)(1.0f / sfD);
    float nDesired = cfg.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      c->quota[l] = cv_round_f(nDesired);
      sum += c->quota[l];
      nDesired *= factor;
    }
    c->quota[nl - 1] = std::max(cfg.nfeatures - sum, 0);
  }
  // umax (:478-493)
  {
    int umax[FT_HALF_PATCH + 2] = {0};
    const int vmax = (int)std::floor(FT_HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(FT_HALF_PATCH * std::sqrt(2.f) / 2);
    const double hp2 = FT_HALF_PATCH * FT_HALF_PATCH;
    int v, v0;
    for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(std::sqrt(hp2 - v * v));
    for (v = FT_HALF_PATCH, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    for (int i = 0; i < 16; i++) P.umax[i] = umax[i];
  }
  int off = 0, cellBase = 0, cellKp = 0, cand = 0, lvlKp = 0, xt = 0, yt = 0, tiles = 0, octX = 0, octY = 0;
  for (int l = 0; l < nl; l++) {
    FtLevel& L = P.lv[l];
    L.w = cv_round_f((float)cfg.width * c->invScale[l]);     // (:1500)
    L.h = cv_round_f((float)cfg.height * c->invScale[l]);
    // level 0 is only ever read: a tight pitch lets the upload be one contiguous DMA straight into the slab
    L.pitch = (l == 0 && L.w % 4 == 0) ? L.w : align_up(L.w, 64);
    L.offset = off;
    off += align_up(L.pitch * L.h, 256);
    L.maxBorderX = L.w - FT_EDGE_THRESHOLD + 3;
    L.maxBorderY = L.h - FT_EDGE_THRESHOLD + 3;
    const float width = (float)(L.maxBorderX - FT_MIN_BORDER), height = (float)(L.maxBorderY - FT_MIN_BORDER);
    L.nCols = (int)(width / 35.f);
    L.nRows = (int)(height / 35.f);
    if (L.nCols < 1 || L.nRows < 1) {
      set_err("ft_context_create: image too small for the requested number of pyramid levels");
      return FT_ERR_INVALID;
    }
    L.wCell = (int)std::ceil(width / L.nCols);
    L.hCell = (int)std::ceil(height / L.nRows);
    L.cellBase = cellBase;
    cellBase += L.nCols * L.nRows;
    L.cellCap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
    L.cellKpBase = cellKp;
    cellKp += L.nCols * L.nRows * L.cellCap;
    L.candBase = cand;
    L.candCap = std::min(L.nCols * L.nRows * L.cellCap, 0xFFFFF);
    cand += L.candCap;
    L.quota = c->quota[l];
    L.nIni = (int)std::round((float)(L.maxBorderX - FT_MIN_BORDER) / (L.maxBorderY - FT_MIN_BORDER));   // (:664)
    if (L.nIni < 1) {
      set_err("ft_context_create: aspect ratio below 1:2 is not supported (the reference indexes an empty vector)");
      return FT_ERR_INVALID;
    }
    L.hX = (float)(L.maxBorderX - FT_MIN_BORDER) / L.nIni;
    if (L.nIni > 8) { set_err("ft_context_create: aspect ratio above 8:1 is not supported"); return FT_ERR_INVALID; }
    L.nodeCap = std::max(L.quota + 4 * L.nIni + 8, 32);
    L.lvlKpBase = lvlKp;
    L.lvlKpCap = L.nodeCap;
    lvlKp += L.lvlKpCap;
    L.xTab = xt; L.yTab = yt;
    xt += L.w; yt += L.h;
    L.octTabX = octX; L.octTabY = octY;
    octX += L.maxBorderX - FT_MIN_BORDER + 1; octY += L.maxBorderY - FT_MIN_BORDER + 1;
    L.blurTileBase = tiles;
    L.blurTilesX = (L.w + 63) / 64;
    tiles += L.blurTilesX * ((L.h + 31) / 32);
  }
  P.totalCells = cellBase;
  P.totalBlurTiles = tiles;
  P.maxKp = align_up(lvlKp, 8);
  if (P.maxKp > 65535) { set_err("ft_context_create: too many keypoints"); return FT_ERR_INVALID; }
  c->pyrBytes = off;
  return FT_OK;
}

// cv::resize INTER_LINEAR coefficient tables (OpenCV resize.cpp) for one (source, destination) size pair
static void resize_tables(int sw, int sh, int dw, int dh, int2* xTab, int2* yTab) {
  const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    const short a0 = (short)cv_round_f((1.f - fx) * 2048.f), a1 = (short)cv_round_f(fx * 2048.f);
    xTab[dx] = make_int2(sx, (int)((unsigned short)a0 | ((unsigned)(unsigned short)a1 << 16)));
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    const short b0 = (short)cv_round_f((1.f - fy) * 2048.f), b1 = (short)cv_round_f(fy * 2048.f);
    const int y0 = std::min(std::max(sy, 0), sh - 1), y1 = std::min(std::max(sy + 1, 0), sh - 1);
    yTab[dy] = make_int2(y0 | (y1 << 16), (int)((unsigned short)b0 | ((unsigned)(unsigned short)b1 << 16)));
  }
}

static ft_status build_tables(ft_context* c, int cellKpTotal, int candTotal, int lvlKpTotal) {
  const FtParams& P = c->P;
  (void)cellKpTotal; (void)candTotal; (void)lvlKpTotal;
  int xt = 0, yt = 0;
  for (int l = 0; l < P.nlevels; l++) { xt += P.lv[l].w; yt += P.lv[l].h; }
  std::vector<int2> xTab(xt), yTab(yt);
  for (int l = 1; l < P.nlevels; l++) {
    const FtLevel& L = P.lv[l];
    const FtLevel& S = P.lv[l - 1];
    resize_tables(S.w, S.h, L.w, L.h, &xTab[L.xTab], &yTab[L.yTab]);
  }
  int2 *dx = nullptr, *dy = nullptr;
  CK(dalloc(c, &dx, xTab.size()));
  CK(dalloc(c, &dy, yTab.size()));
  CK(cudaMemcpy(dx, xTab.data(), xTab.size() * sizeof(int2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, yTab.data(), yTab.size() * sizeof(int2), cudaMemcpyHostToDevice));
  c->B.xTab = dx; c->B.yTab = dy;
  // quadtree path tables of the octree kernel (they depend on the level geometry only)
  int ox = 0, oy = 0;
  for (int l = 0; l < P.nlevels; l++) { ox += P.lv[l].maxBorderX - FT_MIN_BORDER + 1; oy += P.lv[l].maxBorderY - FT_MIN_BORDER + 1; }
  std::vector<uint32_t> octX(ox), octY(oy);
  for (int l = 0; l < P.nlevels; l++) ft_octree_tables(P.lv[l], &octX[P.lv[l].octTabX], &octY[P.lv[l].octTabY]);
  uint32_t *dox = nullptr, *doy = nullptr;
  CK(dalloc(c, &dox, octX.size()));
  CK(dalloc(c, &doy, octY.size()));
  CK(cudaMemcpy(dox, octX.data(), octX.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(doy, octY.data(), octY.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  c->B.octTabX = dox; c->B.octTabY = doy;
  return FT_OK;
}

extern "C" ft_status ft_context_create(const ft_config* cfg, ft_context** out) {
  if (!cfg || !out) { set_err("ft_context_create: null argument"); return FT_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("ft_context_create: no CUDA device available (this library has no CPU fallback)");
    return FT_ERR_CUDA;
  }
  if (cfg->device_id < 0 || cfg->device_id >= ndev) { set_err("ft_context_create: bad device_id"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(cfg->device_id));
  ft_context* c = new ft_context();
  c->cfg = *cfg;
  ft_status st = build_params(c);
  if (st != FT_OK) { delete c; return st; }
  FtParams& P = c->P;
  {
    // the per-level octree keeps its node list (and, when they fit, the level's candidates) in shared memory
    const size_t budget = ft_octree_smem_budget();
    const char* denseEnv = getenv("FT_OCT_DENSE");      // "0": general (sort-based) octree path only (tests, A/B runs)
    int denseTotal = 0;
    for (int l = 0; l < P.nlevels; l++) {
      const bool ok = ft_octree_plan(P.lv[l], budget);
      if (ok) {
        if (denseEnv && denseEnv[0] == '0') P.lv[l].octDenseDepth = 0;
        P.lv[l].octDenseBase = denseTotal;
        if (P.lv[l].octDenseDepth > 0) denseTotal += P.lv[l].nIni << (2 * P.lv[l].octDenseDepth);
        c->octDenseTotal = denseTotal;
        continue;
      }
      {
        char buf[200];
        snprintf(buf, sizeof(buf), "ft_context_create: %d features put %d octree nodes on level %d (60 B each, %zu KB of shared memory available); "
                 "reduce nfeatures", cfg->nfeatures, P.lv[l].nodeCap, l, budget >> 10);
        set_err(buf);
        delete c;
        return FT_ERR_CAPACITY;
      }
    }
  }
  c->fisheye = cfg->camera_type == FT_CAM_KB8;
  c->cam1.type = c->cam2.type = cfg->camera_type;
  memcpy(c->cam1.p, cfg->cam1, 32); memcpy(c->cam2.p, cfg->cam2, 32);
  c->mbf = cfg->bf;
  c->mb = cfg->bf / cfg->cam1[0];            // mb = mbf/fx (Frame.cc:199)
  c->minX = 0.f; c->maxX = (float)cfg->width; c->minY = 0.f; c->maxY = (float)cfg->height;   // ComputeImageBounds, no distortion
  c->gridWInv = (float)FT_GRID_COLS / (c->maxX - c->minX);
  c->gridHInv = (float)FT_GRID_ROWS / (c->maxY - c->minY);
  c->logScale = (float)std::log((double)cfg->scale_factor);   // mfLogScaleFactor = log(mfScaleFactor) (Frame.cc:112-113)
  // pose defaults to identity; rig extrinsics from Tlr
  memset(&c->pose, 0, sizeof(c->pose));
  for (int i = 0; i < 3; i++) { c->pose.Rcw[4 * i] = 1; c->pose.Rwc[4 * i] = 1; c->pose.Rlr[4 * i] = 1; c->pose.Rrl[4 * i] = 1; }
  if (c->fisheye) {
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) c->pose.Rlr[3 * i + j] = cfg->Tlr[4 * i + j];
      c->pose.tlr[i] = cfg->Tlr[4 * i + 3];
    }
    // Trl = Tlr^-1
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c->pose.Rrl[3 * i + j] = c->pose.Rlr[3 * j + i];
    for (int i = 0; i < 3; i++)
      c->pose.trl[i] = -(c->pose.Rrl[3 * i] * c->pose.tlr[0] + c->pose.Rrl[3 * i + 1] * c->pose.tlr[1] + c->pose.Rrl[3 * i + 2] * c->pose.tlr[2]);
  }

  int cellKpTotal = 0, candTotal = 0, lvlKpTotal = 0;
  for (int l = 0; l < P.nlevels; l++) {
    cellKpTotal += P.lv[l].nCols * P.lv[l].nRows * P.lv[l].cellCap;
    candTotal += P.lv[l].candCap;
    lvlKpTotal += P.lv[l].lvlKpCap;
  }
  auto fail = [&](ft_status s) { ft_context_destroy(c); return s; };
#define CKF(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
      return fail(FT_ERR_CUDA);                                                                    \
    }                                                                                              \
  } while (0)
  {
    // frame result slab
    const size_t K = (size_t)P.maxKp;
    size_t o = 64;
    c->offKps[0] = o; o += K * sizeof(ft_keypoint);
    c->offKps[1] = o; o += K * sizeof(ft_keypoint);
    c->offDesc[0] = o; o += K * 32;
    c->offDesc[1] = o; o += K * 32;
    c->offUR = o; o += K * 4;
    c->offDepth = o; o += K * 4;
    c->frameBytesPinhole = o;
    c->offL2R = o; o += K * 4;
    c->offR2L = o; o += K * 4;
    c->offP3D = o; o += K * 12;
    c->frameBytesAll = o;
    CKF(dalloc(c, &c->dFrame, o));
    CKF(cudaMallocHost((void**)&c->hFrame, o));
    memset(c->hFrame, 0, o);
    for (int e = 0; e < 2; e++) {
      c->B.eye[e].counts = reinterpret_cast<int*>(c->dFrame) + 2 * e;
      c->B.eye[e].kps = reinterpret_cast<ft_keypoint*>(c->dFrame + c->offKps[e]);
      c->B.eye[e].desc = c->dFrame + c->offDesc[e];
    }
    c->B.status = reinterpret_cast<int*>(c->dFrame) + 4;
    c->S.uRight = reinterpret_cast<float*>(c->dFrame + c->offUR);
    c->S.depth = reinterpret_cast<float*>(c->dFrame + c->offDepth);
    c->S.l2r = reinterpret_cast<int*>(c->dFrame + c->offL2R);
    c->S.r2l = reinterpret_cast<int*>(c->dFrame + c->offR2L);
    c->S.p3d = reinterpret_cast<float*>(c->dFrame + c->offP3D);
  }
  for (int e = 0; e < 2; e++) {
    FtEye& E = c->B.eye[e];
    CKF(dalloc(c, &E.pyr, (size_t)c->pyrBytes));
    CKF(dalloc(c, &E.blur, (size_t)c->pyrBytes));
    CKF(dalloc(c, &E.cellKp, (size_t)cellKpTotal));
    CKF(dalloc(c, &E.cellCount, (size_t)P.totalCells));
    CKF(dalloc(c, &E.cand, (size_t)candTotal));
    CKF(dalloc(c, &E.octScratch, (size_t)candTotal * 20));
    CKF(dalloc(c, &E.octCnt, (size_t)c->octDenseTotal + 4));
    CKF(dalloc(c, &E.octBest, (size_t)c->octDenseTotal + 4));
    CKF(dalloc(c, &E.lvlCandCount, (size_t)FT_MAX_LEVELS));
    CKF(dalloc(c, &E.lvlKp, (size_t)lvlKpTotal));
    CKF(dalloc(c, &E.lvlKpCount, (size_t)FT_MAX_LEVELS));
    CKF(dalloc(c, &E.octClock, (size_t)2 * FT_MAX_LEVELS * 64));
    CKF(cudaMallocHost((void**)&c->hIn[e], (size_t)cfg->width * cfg->height));
  }
  if (build_tables(c, cellKpTotal, candTotal, lvlKpTotal) != FT_OK) return fail(FT_ERR_CUDA);
  // stereo
  CKF(dalloc(c, &c->S.bestIdxR, (size_t)P.maxKp));
  CKF(dalloc(c, &c->S.sad, (size_t)P.maxKp));
  CKF(dalloc(c, &c->S.code, (size_t)P.maxKp));
  CKF(dalloc(c, &c->S.stats, (size_t)8));
  c->B.stereoStats = c->S.stats;
  // grid
  CKF(dalloc(c, &c->G.cellStart, (size_t)2 * FT_GRID_STRIDE));
  CKF(dalloc(c, &c->G.cellIdx, (size_t)2 * P.maxKp));
  CKF(dalloc(c, &c->G.rec, (size_t)2 * P.maxKp));
  CKF(dalloc(c, &c->G.kpUn, (size_t)P.maxKp));
  // projection search
  const int MM = cfg->max_map_points > 0 ? cfg->max_map_points : 25000;
  c->cfg.max_map_points = MM;
  FtSbpBuffers& Q = c->Q;
  {
    const size_t S2 = (size_t)2 * P.maxKp;
    c->mpHolderBytes = S2 * 4 + ((S2 + 15) / 16) * 16;
    const size_t mpBytes = c->mpHolderBytes + (size_t)MM * 68;
    CKF(dalloc(c, &c->dMp, mpBytes));
    CKF(cudaMallocHost((void**)&c->hMp, mpBytes));
    memset(c->hMp, 0, mpBytes);
    Q.holderInit = reinterpret_cast<int*>(c->dMp);
    Q.holderObsInit = c->dMp + S2 * 4;
    c->holderInit = Q.holderInit; c->holderObsInit = Q.holderObsInit;
    c->offOutHolder = 64; c->offOutObs = 64 + S2 * 4; c->offOutSel = c->offOutObs + ((S2 + 15) / 16) * 16;
    const size_t outBytes = c->offOutSel + (size_t)MM * 8;
    CKF(dalloc(c, &c->dOut, outBytes));
    CKF(cudaMallocHost((void**)&c->hOut, outBytes));
    memset(c->hOut, 0, outBytes);
    Q.cursor = reinterpret_cast<int*>(c->dOut);
    Q.holder = reinterpret_cast<int*>(c->dOut + c->offOutHolder);
    Q.holderObs = c->dOut + c->offOutObs;
    Q.sel = reinterpret_cast<int*>(c->dOut + c->offOutSel);
    Q.pos = Q.normal = Q.minmax = nullptr; Q.desc = nullptr; Q.flags = nullptr; Q.slot = nullptr;   // set per snapshot (tight layout for its M)
  }
  CKF(dalloc(c, &Q.trI, (size_t)MM * 4));
  CKF(dalloc(c, &Q.trF, (size_t)MM * 9));
  CKF(dalloc(c, &Q.listOff, (size_t)MM * 2));
  CKF(dalloc(c, &Q.listLen, (size_t)MM * 2));
  Q.poolCap = MM * 96;
  CKF(dalloc(c, &Q.pool, (size_t)Q.poolCap));
  CKF(dalloc(c, &Q.active, (size_t)MM));
  CKF(cudaMallocHost((void**)&c->hCounts, 64 * sizeof(int)));
  memset(c->hCounts, 0, 64 * sizeof(int));
  // Launch priorities (captured into the graph's kernel nodes): the blur is only read by the orientation + descriptor kernel
  // that closes the extraction, so it has the whole FAST -> octree phase to hide in; at t = 0 it would otherwise compete
  // with FAST of level 0 and the first resizes, which are on the critical path. FT_PRIORITIES=0: all streams equal.
  int prLeast = 0, prGreatest = 0;
  CKF(cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
  { const char* e = getenv("FT_PRIORITIES"); if (e && e[0] == '0') prLeast = prGreatest = 0; }
  CKF(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prGreatest));
  CKF(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prLeast));      // blur
  CKF(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prGreatest));   // FAST -> octree of level 0
  { const char* e = getenv("FT_TOPOLOGY"); if (e && !strcmp(e, "grouped")) c->grouped = 1; }
  { const char* e = getenv("FT_SEARCH_GRAPH"); if (e && e[0] == '0') c->searchGraph = 0; }
  for (int l = 0; l < P.nlevels; l++) {
    CKF(cudaStreamCreateWithPriority(&c->lvStream[l], cudaStreamNonBlocking, prGreatest));
    CKF(cudaEventCreateWithFlags(&c->lvReady[l], cudaEventDisableTiming));
    CKF(cudaEventCreateWithFlags(&c->lvDone[l], cudaEventDisableTiming));
  }
  CKF(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
  CKF(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
  CKF(cudaEventCreateWithFlags(&c->evFork2, cudaEventDisableTiming));
  CKF(cudaEventCreateWithFlags(&c->evJoin2, cudaEventDisableTiming));
  CKF(cudaEventCreateWithFlags(&c->evPyr, cudaEventDisableTiming));
  CKF(cudaEventCreateWithFlags(&c->evJoin3, cudaEventDisableTiming));
  CKF(ft_launch_extract_setup(P));
  CKF(ft_launch_sbp_setup(P));
  CKF(ft_launch_stereo_setup(P));
  *out = c;
  return FT_OK;
}

static void store_detach(ft_context* c) {
  ft_map_store* st = c->store;
  if (!st) return;
  c->store = nullptr;
  {
    std::lock_guard<std::mutex> lk(st->mu);
    for (size_t i = 0; i < st->users.size(); i++)
      if (st->users[i] == c) { st->users.erase(st->users.begin() + i); break; }
    if (!st->users.empty()) return;
  }
  cudaFree(st->pos); cudaFree(st->normal); cudaFree(st->minmax); cudaFree(st->desc);
  delete st;
}

extern "C" ft_status ft_context_destroy(ft_context* c) {
  if (!c) return FT_OK;
  cudaSetDevice(c->cfg.device_id);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->gExtract) cudaGraphExecDestroy(c->gExtract);
  if (c->gStereo) cudaGraphExecDestroy(c->gStereo);
  if (c->gFrame) cudaGraphExecDestroy(c->gFrame);
  ft_search_graph_destroy(&c->gSearch);
  store_detach(c);
  if (c->hUpd) cudaFreeHost(c->hUpd);
  if (c->dUpd) cudaFree(c->dUpd);
  if (c->storeSearched) cudaEventDestroy(c->storeSearched);
  if (c->storeUpdated) cudaEventDestroy(c->storeUpdated);
  if (c->updStaged) cudaEventDestroy(c->updStaged);
  if (c->mpStaged) cudaEventDestroy(c->mpStaged);
  if (c->searchDone) cudaEventDestroy(c->searchDone);
  for (void* p : c->allocs) cudaFree(p);
  for (int e = 0; e < 2; e++) if (c->hIn[e]) cudaFreeHost(c->hIn[e]);
  if (c->hCounts) cudaFreeHost(c->hCounts);
  if (c->hFrame) cudaFreeHost(c->hFrame);
  if (c->hMp) cudaFreeHost(c->hMp);
  if (c->hOut) cudaFreeHost(c->hOut);
  if (c->hDepth) cudaFreeHost(c->hDepth);
  if (c->hBow) cudaFreeHost(c->hBow);
  for (int i = 0; i < FT_STAGE_COUNT; i++) { if (c->evA[i]) cudaEventDestroy(c->evA[i]); if (c->evB[i]) cudaEventDestroy(c->evB[i]); }
  if (c->evFork) cudaEventDestroy(c->evFork);
  if (c->evJoin) cudaEventDestroy(c->evJoin);
  if (c->evFork2) cudaEventDestroy(c->evFork2);
  if (c->evJoin2) cudaEventDestroy(c->evJoin2);
  if (c->evPyr) cudaEventDestroy(c->evPyr);
  if (c->evJoin3) cudaEventDestroy(c->evJoin3);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  for (int l = 0; l < FT_MAX_LEVELS; l++) {
    if (c->lvStream[l]) cudaStreamDestroy(c->lvStream[l]);
    if (c->lvReady[l]) cudaEventDestroy(c->lvReady[l]);
    if (c->lvDone[l]) cudaEventDestroy(c->lvDone[l]);
  }
  delete c;
  return FT_OK;
}

extern "C" ft_status ft_get_scale_tables(ft_context* c, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                                         int* features_per_level) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  for (int i = 0; i < c->P.nlevels; i++) {
    if (scale) scale[i] = c->scale[i];
    if (inv_scale) inv_scale[i] = c->invScale[i];
    if (sigma2) sigma2[i] = c->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = c->invSigma2[i];
    if (features_per_level) features_per_level[i] = c->quota[i];
  }
  return FT_OK;
}

// The per-frame extraction chain. Enqueued on c->stream with the blur forked onto c->stream2;
// identical whether it is being captured into a graph or launched directly.
static int enqueue_extract(ft_context* c) {
  // Launch topology. The input images have already been copied straight into the level-0 slabs.
  // Graph mode (the product path): every pyramid level is its own branch, started the moment the level exists,
  //   s    : resize 1 -> resize 2 -> ... -> resize n-1                      (the only inherent chain)
  //   lv[l]: (after resize l; l = 0 immediately) FAST l -> octree l          (the octrees are latency-bound and
  //                                                                           independent: they overlap on different SMs)
  //   s2   : blur 0, (after the last resize) blur 1..n-1
  //   join -> orientation + descriptors
  // Direct-launch mode (per-kernel CUDA-event timing) groups levels 1..n-1 into one FAST and one octree launch so
  // that each stage is a single timed interval.
  const FtParams& P = c->P;
  const int nl = P.nlevels;
  int n = 0;
  cudaStream_t s = c->stream, s2 = c->stream2, s3 = c->stream3;
  if (c->timing == 2) { s2 = s; s3 = s; }     // isolated stage timing: every launch on one stream, nothing overlaps
  const bool perLevel = !c->timing && !c->grouped;
  if (c->rectify) { ft_launch_remap(P, c->B, c->dRaw[0], c->dRaw[1], c->dRemapTab, c->rawW, c->rawH, s); n++; }
  else if (c->inResize) { ft_launch_resize_input(P, c->B, c->dRaw[0], c->dRaw[1], c->rawW, s); n++; }
  cudaEventRecord(c->evFork, s);
  cudaStreamWaitEvent(s3, c->evFork, 0);
  cudaStreamWaitEvent(s2, c->evFork, 0);
  { StageScope t(c, FT_STAGE_FAST_L0, s3); ft_launch_fast(P, c->B, 0, 1, s3); n++; }
  { StageScope t(c, FT_STAGE_OCTREE_L0, s3); ft_launch_octree(P, c->B, 0, 1, s3); n++; }
  cudaEventRecord(c->evJoin3, s3);
  { StageScope t(c, FT_STAGE_BLUR_L0, s2); ft_launch_blur(P, c->B, 0, 1, s2); n++; }
  if (nl > 1) {
    if (perLevel) {
      for (int l = 1; l < nl; l++) {
        ft_launch_resize(P, c->B, l, s); n++;
        cudaEventRecord(c->lvReady[l], s);
        cudaStreamWaitEvent(c->lvStream[l], c->lvReady[l], 0);
        ft_launch_fast(P, c->B, l, l + 1, c->lvStream[l]); n++;
        ft_launch_octree(P, c->B, l, l + 1, c->lvStream[l]); n++;
        cudaEventRecord(c->lvDone[l], c->lvStream[l]);
      }
      cudaEventRecord(c->evPyr, s);
      cudaStreamWaitEvent(s2, c->evPyr, 0);
      ft_launch_blur(P, c->B, 1, nl, s2); n++;
      for (int l = 1; l < nl; l++) cudaStreamWaitEvent(s, c->lvDone[l], 0);
    } else {
      { StageScope t(c, FT_STAGE_RESIZE, s); for (int l = 1; l < nl; l++) { ft_launch_resize(P, c->B, l, s); n++; } }
      cudaEventRecord(c->evPyr, s);
      cudaStreamWaitEvent(s2, c->evPyr, 0);
      { StageScope t(c, FT_STAGE_BLUR, s2); ft_launch_blur(P, c->B, 1, nl, s2); n++; }
      { StageScope t(c, FT_STAGE_FAST, s); ft_launch_fast(P, c->B, 1, nl, s); n++; }
      { StageScope t(c, FT_STAGE_OCTREE, s); ft_launch_octree(P, c->B, 1, nl, s); n++; }
    }
  }
  cudaEventRecord(c->evJoin, s2);
  cudaStreamWaitEvent(s, c->evJoin, 0);
  cudaStreamWaitEvent(s, c->evJoin3, 0);
  { StageScope t(c, FT_STAGE_ORIENT, s); ft_launch_orient_desc(P, c->B, s); n++; }
  return n;
}

// Stereo matching with the frame grid (only read by the projection search) built on a parallel branch.
static int enqueue_stereo(ft_context* c) {
  cudaStream_t s = c->stream, s2 = c->stream2;
  if (c->timing == 2) s2 = s;
  cudaEventRecord(c->evFork2, s);
  cudaStreamWaitEvent(s2, c->evFork2, 0);
  { StageScope t(c, FT_STAGE_GRID, s2); ft_launch_grid(c->P, c->B, c->G, c->fisheye, c->minX, c->minY, c->gridWInv, c->gridHInv, c->und, s2); }
  cudaEventRecord(c->evJoin2, s2);
  if (c->fisheye) { StageScope t(c, FT_STAGE_STEREO, s); ft_launch_fisheye(c->P, c->B, c->S, c->cam1, c->cam2, c->pose, s); }
  else { StageScope t(c, FT_STAGE_STEREO, s); ft_launch_stereo_match(c->P, c->B, c->S, c->mbf, c->mb, s); }
  cudaStreamWaitEvent(s, c->evJoin2, 0);
  return c->fisheye ? 3 : 2;
}

static ft_status run_extract(ft_context* c) {
  // inputs are already in the level-0 slabs
  if (c->useGraph && !c->timing) {
    if (!c->gExtract) {
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      c->nLaunchExtract = enqueue_extract(c);
      CK(cudaStreamEndCapture(c->stream, &g));
      CK(cudaGraphInstantiate(&c->gExtract, g, 0));
      cudaGraphDestroy(g);
    }
    CK(cudaGraphLaunch(c->gExtract, c->stream));
  } else {
    c->nLaunchExtract = enqueue_extract(c);
    CK(cudaGetLastError());
  }
  c->extracted = true; c->stereoDone = false; c->countsValid = false; c->bowValid = false; c->frameStatus = 0;
  return FT_OK;
}

static ft_status upload_images(ft_context* c, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR) {
  if (!c || !imgL || (!imgR && c->P.nEyes == 2)) { set_err("null image (the reference returns -1 on an empty image)"); return FT_ERR_INVALID; }
  // with rectification the raw image goes to a staging buffer (k_remap writes level 0); otherwise straight into level 0
  const bool raw = c->rectify || c->inResize;
  const int w = raw ? c->rawW : c->cfg.width, h = raw ? c->rawH : c->cfg.height;
  const int dpitch = raw ? w : c->P.lv[0].pitch;
  if (stepL < w || (c->P.nEyes == 2 && stepR < w)) { set_err("image step smaller than width"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  const uint8_t* src[2] = {imgL, imgR};
  const int step[2] = {stepL, stepR};
  for (int e = 0; e < c->P.nEyes; e++) {
    cudaPointerAttributes at;
    bool pinned = cudaPointerGetAttributes(&at, src[e]) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    uint8_t* dst = raw ? c->dRaw[e] : c->B.eye[e].pyr + c->P.lv[0].offset;
    if (pinned && step[e] == w && dpitch == w) {
      CK(cudaMemcpyAsync(dst, src[e], (size_t)w * h, cudaMemcpyHostToDevice, c->stream));
    } else if (pinned) {
      CK(cudaMemcpy2DAsync(dst, dpitch, src[e], step[e], w, h, cudaMemcpyHostToDevice, c->stream));
    } else {
      // pageable memory: stage through the context's pinned buffer so the copy stays asynchronous
      CK(cudaStreamSynchronize(c->stream));   // previous frame may still be reading the staging buffer
      for (int y = 0; y < h; y++) memcpy(c->hIn[e] + (size_t)y * w, src[e] + (size_t)y * step[e], w);
      CK(cudaMemcpy2DAsync(dst, dpitch, c->hIn[e], w, w, h, cudaMemcpyHostToDevice, c->stream));
    }
  }
  return FT_OK;
}

extern "C" ft_status ft_extract_stereo(ft_context* c, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR) {
  ft_status st = upload_images(c, imgL, stepL, imgR, stepR);
  if (st != FT_OK) return st;
  return run_extract(c);
}

static ft_status copy_device_images(ft_context* c, const uint8_t* dL, int stepL, const uint8_t* dR, int stepR) {
  const bool raw = c->rectify || c->inResize;
  const int w = raw ? c->rawW : c->cfg.width, h = raw ? c->rawH : c->cfg.height;
  const int dpitch = raw ? w : c->P.lv[0].pitch;
  const uint8_t* src[2] = {dL, dR};
  const int step[2] = {stepL, stepR};
  for (int e = 0; e < c->P.nEyes; e++) {
    uint8_t* dst = raw ? c->dRaw[e] : c->B.eye[e].pyr + c->P.lv[0].offset;
    if (step[e] == w && dpitch == w) CK(cudaMemcpyAsync(dst, src[e], (size_t)w * h, cudaMemcpyDeviceToDevice, c->stream));
    else CK(cudaMemcpy2DAsync(dst, dpitch, src[e], step[e], w, h, cudaMemcpyDeviceToDevice, c->stream));
  }
  return FT_OK;
}

extern "C" ft_status ft_extract_stereo_device(ft_context* c, const uint8_t* dL, int stepL, const uint8_t* dR, int stepR) {
  if (!c || !dL || !dR) { set_err("ft_extract_stereo_device: null argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = copy_device_images(c, dL, stepL, dR, stepR);
  if (st != FT_OK) return st;
  return run_extract(c);
}

static ft_status run_stereo(ft_context* c) {
  CK(cudaSetDevice(c->cfg.device_id));
  if (c->useGraph && !c->timing) {
    if (!c->gStereo) {
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      c->nLaunchStereo = enqueue_stereo(c);
      CK(cudaStreamEndCapture(c->stream, &g));
      CK(cudaGraphInstantiate(&c->gStereo, g, 0));
      cudaGraphDestroy(g);
    }
    CK(cudaGraphLaunch(c->gStereo, c->stream));
  } else {
    c->nLaunchStereo = enqueue_stereo(c);
    CK(cudaGetLastError());
  }
  c->stereoDone = true;
  return FT_OK;
}

extern "C" ft_status ft_stereo_match(ft_context* c) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (!c->extracted) { set_err("ft_stereo_match: no extracted frame"); return FT_ERR_STATE; }
  if (c->fisheye) { set_err("ft_stereo_match: context is a KannalaBrandt8 rig; call ft_stereo_match_fisheye"); return FT_ERR_INVALID; }
  if (c->sensor != FT_SENSOR_STEREO) { set_err("ft_stereo_match: the context is monocular / RGB-D (ft_set_sensor); call ft_depth_from_rgbd"); return FT_ERR_STATE; }
  return run_stereo(c);
}

extern "C" ft_status ft_stereo_match_fisheye(ft_context* c) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (!c->extracted) { set_err("ft_stereo_match_fisheye: no extracted frame"); return FT_ERR_STATE; }
  if (!c->fisheye) { set_err("ft_stereo_match_fisheye: context is a pinhole rig"); return FT_ERR_INVALID; }
  return run_stereo(c);
}

static ft_status check_device_status(ft_context* c, int status) {
  if (!status) return FT_OK;
  char buf[256];
  snprintf(buf, sizeof(buf), "device buffer bound exceeded (status 0x%x:%s%s%s%s%s%s)", status,
           status & FT_ST_CELL_OVERFLOW ? " fast-cell" : "", status & FT_ST_NODE_OVERFLOW ? " octree-node" : "",
           status & FT_ST_KP_OVERFLOW ? " keypoints" : "", status & FT_ST_CAND_OVERFLOW ? " candidates" : "",
           status & FT_ST_SBP_POOL_OVERFLOW ? " search-pool" : "", status & FT_ST_RESOLVE_NOCONV ? " resolve" : "");
  set_err(buf);
  cudaMemsetAsync(c->B.status, 0, sizeof(int), c->stream);
  return FT_ERR_CAPACITY;
}

// A capacity overflow stays latched for the frame it happened in: every later call on that frame (downloads, searches)
// reports it again instead of handing out clamped counts; the next extraction starts clean.
static ft_status fetch_counts(ft_context* c) {
  if (!c->countsValid) {
    CK(cudaMemcpyAsync(c->hCounts, c->dFrame, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));   // counts + status header
    CK(cudaStreamSynchronize(c->stream));
    c->countsValid = true;
    c->frameStatus |= c->hCounts[4];
  }
  return check_device_status(c, c->frameStatus);
}

extern "C" ft_status ft_frame_counts(ft_context* c, int* nL, int* nR, int* monoL, int* monoR) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (!c->extracted) { set_err("ft_frame_counts: no extracted frame"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);
  if (st != FT_OK) return st;
  if (nL) *nL = c->hCounts[0];
  if (monoL) *monoL = c->hCounts[1];
  if (nR) *nR = c->hCounts[2];
  if (monoR) *monoR = c->hCounts[3];
  return FT_OK;
}

extern "C" ft_status ft_frame_download(ft_context* c, int eye, int cap, ft_keypoint* kps, uint8_t* desc, int* n,
                                       int* mono_index, float* u_right, float* depth, int* l2r, int* r2l, float* p3d) {
  if (!c || eye < 0 || eye > 1) { set_err("ft_frame_download: bad argument"); return FT_ERR_INVALID; }
  if (!c->extracted) { set_err("ft_frame_download: no extracted frame"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);
  if (st != FT_OK) return st;
  const int cnt = c->hCounts[2 * eye], nLeft = c->hCounts[0], nRight = c->hCounts[2];
  if (n) *n = cnt;
  if (mono_index) *mono_index = c->hCounts[2 * eye + 1];
  if ((kps || desc) && cap < cnt) { set_err("ft_frame_download: capacity smaller than keypoint count"); return FT_ERR_INVALID; }
  const FtEye& E = c->B.eye[eye];
  if (kps && cnt) CK(cudaMemcpyAsync(kps, E.kps, sizeof(ft_keypoint) * cnt, cudaMemcpyDeviceToHost, c->stream));
  if (desc && cnt) CK(cudaMemcpyAsync(desc, E.desc, 32 * (size_t)cnt, cudaMemcpyDeviceToHost, c->stream));
  if ((u_right || depth || l2r || r2l || p3d) && !c->stereoDone) { set_err("ft_frame_download: stereo results requested before stereo matching"); return FT_ERR_STATE; }
  if (u_right && nLeft) CK(cudaMemcpyAsync(u_right, c->S.uRight, sizeof(float) * nLeft, cudaMemcpyDeviceToHost, c->stream));
  if (depth && nLeft) CK(cudaMemcpyAsync(depth, c->S.depth, sizeof(float) * nLeft, cudaMemcpyDeviceToHost, c->stream));
  if (l2r && nLeft) CK(cudaMemcpyAsync(l2r, c->S.l2r, sizeof(int) * nLeft, cudaMemcpyDeviceToHost, c->stream));
  if (r2l && nRight) CK(cudaMemcpyAsync(r2l, c->S.r2l, sizeof(int) * nRight, cudaMemcpyDeviceToHost, c->stream));
  if (p3d && nLeft) CK(cudaMemcpyAsync(p3d, c->S.p3d, sizeof(float) * 3 * nLeft, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FT_OK;
}

// extract + stereo as ONE captured graph (the Frame constructor does both back to back)
static ft_status run_frame(ft_context* c) {
  if (c->sensor != FT_SENSOR_STEREO) { set_err("the one-call Frame constructor is stereo only; use ft_extract_mono + ft_depth_from_rgbd"); return FT_ERR_STATE; }
  if (c->useGraph && !c->timing) {
    if (!c->gFrame) {
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      c->nLaunchExtract = enqueue_extract(c);
      c->nLaunchStereo = enqueue_stereo(c);
      CK(cudaStreamEndCapture(c->stream, &g));
      CK(cudaGraphInstantiate(&c->gFrame, g, 0));
      cudaGraphDestroy(g);
    }
    CK(cudaGraphLaunch(c->gFrame, c->stream));
    c->extracted = true; c->stereoDone = true; c->countsValid = false; c->bowValid = false; c->frameStatus = 0;
    return FT_OK;
  }
  ft_status st = run_extract(c);
  if (st != FT_OK) return st;
  return run_stereo(c);
}

// Device-resident inputs: copy into the level-0 slabs, then extract + stereo-match as one graph; asynchronous.
extern "C" ft_status ft_frame_enqueue_device(ft_context* c, const uint8_t* dL, int stepL, const uint8_t* dR, int stepR) {
  if (!c || !dL || !dR) { set_err("ft_frame_enqueue_device: null argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = copy_device_images(c, dL, stepL, dR, stepR);
  if (st != FT_OK) return st;
  return run_frame(c);
}

// Device + pinned staging buffers for raw (pre-rectification / pre-resize) input images; grown on demand, the
// previous buffers are released (the caller has synchronised the stream).
static ft_status ensure_raw_buffers(ft_context* c, int raw_width, int raw_height) {
  const size_t need = (size_t)raw_width * raw_height;
  if (c->dRaw[0] && c->rawCap >= need) return FT_OK;
  const size_t hostNeed = std::max(need, (size_t)c->cfg.width * c->cfg.height);
  for (int e = 0; e < 2; e++) {
    dfree(c, c->dRaw[e]); c->dRaw[e] = nullptr;
    CK(dalloc(c, &c->dRaw[e], need));
    if (c->hIn[e]) { cudaFreeHost(c->hIn[e]); c->hIn[e] = nullptr; }
    CK(cudaMallocHost((void**)&c->hIn[e], hostNeed));
  }
  c->rawCap = need;
  return FT_OK;
}

// Stereo rectification in front of the extractor (reference System::TrackStereo, src/System.cc:273-281:
// cv::remap(im, M1, M2, INTER_LINEAR) with the CV_32F maps of Settings.cc:506-509). Maps are width x height floats
// (x map, y map) per eye; raw images are raw_width x raw_height. Passing NULL maps switches rectification off.
extern "C" ft_status ft_set_rectification(ft_context* c, int raw_width, int raw_height, const float* M1l, const float* M2l,
                                          const float* M1r, const float* M2r) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  // the launch graphs bake the topology in: drop them so that the next frame re-captures
  if (c->gExtract) { cudaGraphExecDestroy(c->gExtract); c->gExtract = nullptr; }
  if (c->gFrame) { cudaGraphExecDestroy(c->gFrame); c->gFrame = nullptr; }
  if (!M1l || !M2l || !M1r || !M2r) { c->rectify = 0; return FT_OK; }
  if (c->inResize) { set_err("ft_set_rectification: input resize is active (System.cc:273-285 does one or the other)"); return FT_ERR_STATE; }
  if (raw_width < 8 || raw_height < 8 || raw_width > 8192 || raw_height > 8192) { set_err("ft_set_rectification: bad raw size"); return FT_ERR_INVALID; }
  const int w = c->cfg.width, h = c->cfg.height;
  std::vector<int2> tab((size_t)2 * w * h);
  const float* mx[2] = {M1l, M1r};
  const float* my[2] = {M2l, M2r};
  for (int e = 0; e < 2; e++)
    for (size_t i = 0; i < (size_t)w * h; i++) {
      // OpenCV remap, CV_32FC1 pair -> fixed point: sx = cvRound(x * INTER_TAB_SIZE), integer part saturated to short
      const int sx = cv_round_f(mx[e][i] * 32.f), sy = cv_round_f(my[e][i] * 32.f);
      const int ix = std::min(std::max(sx >> 5, -32768), 32767), iy = std::min(std::max(sy >> 5, -32768), 32767);
      tab[(size_t)e * w * h + i] = make_int2((ix & 0xFFFF) | (iy << 16), (sx & 31) | ((sy & 31) << 8));
    }
  if (!c->dRemapTab) CK(dalloc(c, &c->dRemapTab, tab.size()));
  CK(cudaMemcpy(c->dRemapTab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice));
  { ft_status rs = ensure_raw_buffers(c, raw_width, raw_height); if (rs != FT_OK) return rs; }
  c->rawW = raw_width; c->rawH = raw_height; c->rectify = 1;
  return FT_OK;
}

// ---- monocular / RGB-D sensors ("next" row 4 of SURVEY.md 8f, the RGB-D half) ----
// The monocular and RGB-D Frame constructors (reference src/Frame.cc:226-325, 328-419) run ONE extractor, undistort
// the keypoints and take depth from the registered depth image (ComputeStereoFromRGBD, :1065-1086) or leave
// mvuRight / mvDepth at -1 (monocular, :330-331).
extern "C" ft_status ft_set_sensor(ft_context* c, int sensor) {
  if (!c || sensor < FT_SENSOR_STEREO || sensor > FT_SENSOR_RGBD) { set_err("ft_set_sensor: bad argument"); return FT_ERR_INVALID; }
  if (c->fisheye && sensor != FT_SENSOR_STEREO) { set_err("ft_set_sensor: monocular / RGB-D operation is implemented for pinhole cameras"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  if (c->gExtract) { cudaGraphExecDestroy(c->gExtract); c->gExtract = nullptr; }
  if (c->gStereo) { cudaGraphExecDestroy(c->gStereo); c->gStereo = nullptr; }
  if (c->gFrame) { cudaGraphExecDestroy(c->gFrame); c->gFrame = nullptr; }
  c->sensor = sensor;
  c->P.nEyes = sensor == FT_SENSOR_STEREO ? 2 : 1;
  CK(cudaMemsetAsync(c->B.eye[1].counts, 0, 2 * sizeof(int), c->stream));   // the right eye stays empty
  c->extracted = false; c->stereoDone = false; c->countsValid = false; c->bowValid = false; c->frameStatus = 0;
  return FT_OK;
}

extern "C" ft_status ft_extract_mono(ft_context* c, const uint8_t* img, int step) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (c->sensor == FT_SENSOR_STEREO) { set_err("ft_extract_mono: the context is a stereo rig (ft_set_sensor first)"); return FT_ERR_STATE; }
  ft_status st = upload_images(c, img, step, nullptr, 0);
  if (st != FT_OK) return st;
  return run_extract(c);
}

// depth: CV_32F depth image registered to the gray image (after Tracking::GrabImageRGBD's convertTo with
// mDepthMapFactor, src/Tracking.cc:1552-1553), step in bytes; NULL = monocular frame (mvuRight = mvDepth = -1).
extern "C" ft_status ft_depth_from_rgbd(ft_context* c, const float* depth, int step_bytes) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (c->sensor == FT_SENSOR_STEREO) { set_err("ft_depth_from_rgbd: the context is a stereo rig (ft_set_sensor first)"); return FT_ERR_STATE; }
  if (!c->extracted) { set_err("ft_depth_from_rgbd: no extracted frame"); return FT_ERR_STATE; }
  if (c->sensor == FT_SENSOR_RGBD && !depth) { set_err("ft_depth_from_rgbd: RGB-D context needs a depth image"); return FT_ERR_INVALID; }
  if (c->sensor == FT_SENSOR_MONOCULAR && depth) { set_err("ft_depth_from_rgbd: monocular context takes no depth image"); return FT_ERR_INVALID; }
  const int w = c->cfg.width, h = c->cfg.height;
  if (depth && (step_bytes < w * 4 || step_bytes % 4)) { set_err("ft_depth_from_rgbd: bad step"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  cudaStream_t s = c->stream;
  if (depth) {
    if (!c->dDepth) {
      CK(dalloc(c, &c->dDepth, (size_t)w * h));
      CK(cudaMallocHost((void**)&c->hDepth, sizeof(float) * (size_t)w * h));
    }
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, depth) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
      CK(cudaMemcpy2DAsync(c->dDepth, sizeof(float) * w, depth, step_bytes, sizeof(float) * w, h, cudaMemcpyHostToDevice, s));
    } else {
      CK(cudaStreamSynchronize(s));   // the previous frame may still be reading the staging buffer
      for (int y = 0; y < h; y++) memcpy(c->hDepth + (size_t)y * w, reinterpret_cast<const uint8_t*>(depth) + (size_t)y * step_bytes, sizeof(float) * w);
      CK(cudaMemcpyAsync(c->dDepth, c->hDepth, sizeof(float) * (size_t)w * h, cudaMemcpyHostToDevice, s));
    }
  }
  // frame grid first (it writes mvKeysUn), then the depth lookup
  { StageScope t(c, FT_STAGE_GRID, s); ft_launch_grid(c->P, c->B, c->G, 0, c->minX, c->minY, c->gridWInv, c->gridHInv, c->und, s); }
  { StageScope t(c, FT_STAGE_STEREO, s); ft_launch_rgbd_depth(c->P, c->B, c->S, c->G.kpUn, depth ? c->dDepth : nullptr, w, c->mbf, s); }
  CK(cudaGetLastError());
  c->nLaunchStereo = 2;
  c->stereoDone = true; c->countsValid = false;
  return FT_OK;
}

// cv::resize(im, imToFeed, settings_->newImSize()) in front of the extractor (reference src/System.cc:282-285,
// Settings::needToResize): raw images are raw_width x raw_height and level 0 is their INTER_LINEAR resize.
// raw_width == 0 switches it off. Mutually exclusive with rectification, as in the reference (if / else if).
extern "C" ft_status ft_set_input_resize(ft_context* c, int raw_width, int raw_height) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  if (c->gExtract) { cudaGraphExecDestroy(c->gExtract); c->gExtract = nullptr; }
  if (c->gFrame) { cudaGraphExecDestroy(c->gFrame); c->gFrame = nullptr; }
  if (raw_width == 0 || raw_height == 0) { c->inResize = 0; return FT_OK; }
  if (c->rectify) { set_err("ft_set_input_resize: rectification is active (the reference resizes only when it does not rectify)"); return FT_ERR_STATE; }
  if (raw_width < 8 || raw_height < 8 || raw_width > 16384 || raw_height > 16384) { set_err("ft_set_input_resize: bad raw size"); return FT_ERR_INVALID; }
  const int w = c->cfg.width, h = c->cfg.height;
  std::vector<int2> xt(w), yt(h);
  resize_tables(raw_width, raw_height, w, h, xt.data(), yt.data());
  CK(cudaMemcpy(const_cast<int2*>(c->B.xTab) + c->P.lv[0].xTab, xt.data(), sizeof(int2) * w, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(const_cast<int2*>(c->B.yTab) + c->P.lv[0].yTab, yt.data(), sizeof(int2) * h, cudaMemcpyHostToDevice));
  { ft_status rs = ensure_raw_buffers(c, raw_width, raw_height); if (rs != FT_OK) return rs; }
  c->rawW = raw_width; c->rawH = raw_height; c->inResize = 1;
  return FT_OK;
}

// Frame::UndistortKeyPoints + Frame::ComputeImageBounds (reference src/Frame.cc:771-835) for a pinhole camera with
// distortion coefficients (mDistCoef = k1 k2 p1 p2 [k3]). With k1 == 0 the reference copies mvKeys (Frame.cc:773-777).
extern "C" ft_status ft_set_distortion(ft_context* c, const float* dist_coef, int n) {
  if (!c || n < 0 || n > 5 || (n > 0 && !dist_coef)) { set_err("ft_set_distortion: bad argument (0, 4 or 5 coefficients)"); return FT_ERR_INVALID; }
  if (c->fisheye && n > 0 && dist_coef[0] != 0.0f) { set_err("ft_set_distortion: KannalaBrandt8 rigs keep mvKeys (Frame.cc:1157: no undistortion)"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  // the launch graphs bake the kernel arguments in: drop them so that the next frame re-captures
  if (c->gExtract) { cudaGraphExecDestroy(c->gExtract); c->gExtract = nullptr; }
  if (c->gStereo) { cudaGraphExecDestroy(c->gStereo); c->gStereo = nullptr; }
  if (c->gFrame) { cudaGraphExecDestroy(c->gFrame); c->gFrame = nullptr; }
  FtUndistort u = {};
  u.fx = c->cam1.p[0]; u.fy = c->cam1.p[1]; u.cx = c->cam1.p[2]; u.cy = c->cam1.p[3];
  u.ifx = 1. / u.fx; u.ify = 1. / u.fy;
  for (int i = 0; i < n; i++) u.k[i] = (double)dist_coef[i];
  u.on = (n > 0 && dist_coef[0] != 0.0f) ? 1 : 0;
  c->und = u;
  const float cols = (float)c->cfg.width, rows = (float)c->cfg.height;
  if (u.on) {   // ComputeImageBounds: the four undistorted corners
    float x[4], y[4];
    const float cx[4] = {0.f, cols, 0.f, cols}, cy[4] = {0.f, 0.f, rows, rows};
    for (int i = 0; i < 4; i++) ft_undistort_point(u, cx[i], cy[i], x[i], y[i]);
    c->minX = std::min(x[0], x[2]); c->maxX = std::max(x[1], x[3]);
    c->minY = std::min(y[0], y[1]); c->maxY = std::max(y[2], y[3]);
  } else {
    c->minX = 0.f; c->maxX = cols; c->minY = 0.f; c->maxY = rows;
  }
  c->gridWInv = (float)FT_GRID_COLS / (c->maxX - c->minX);
  c->gridHInv = (float)FT_GRID_ROWS / (c->maxY - c->minY);
  return FT_OK;
}

// mnMinX, mnMaxX, mnMinY, mnMaxY (Frame::ComputeImageBounds)
extern "C" ft_status ft_image_bounds(ft_context* c, float* out4) {
  if (!c || !out4) { set_err("ft_image_bounds: null argument"); return FT_ERR_INVALID; }
  out4[0] = c->minX; out4[1] = c->maxX; out4[2] = c->minY; out4[3] = c->maxY;
  return FT_OK;
}

// mvKeysUn[i].pt of the left eye (every other KeyPoint field equals mvKeys[i]); xy holds 2*cap floats. Needs the frame
// grid, i.e. a frame processed by ft_stereo_match / ft_frame_construct / ft_frame_enqueue_device.
extern "C" ft_status ft_frame_keypoints_undistorted(ft_context* c, int cap, float* xy, int* n) {
  if (!c || !xy) { set_err("ft_frame_keypoints_undistorted: null argument"); return FT_ERR_INVALID; }
  if (!c->extracted || !c->stereoDone) { set_err("ft_frame_keypoints_undistorted: the frame grid has not been built (stereo stage)"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);
  if (st != FT_OK) return st;
  const int nl = c->hCounts[0];
  if (n) *n = nl;
  if (cap < nl) { set_err("ft_frame_keypoints_undistorted: capacity smaller than keypoint count"); return FT_ERR_INVALID; }
  if (nl) CK(cudaMemcpyAsync(xy, c->G.kpUn, sizeof(float) * 2 * nl, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FT_OK;
}

// Test hook (host arithmetic only, no GPU needed): the cosf/sinf restatement used by the descriptor kernel
extern "C" void ft_debug_sincosf(int n, const float* angle, float* sin_out, float* cos_out) {
  for (int i = 0; i < n; i++) ft_glibc_sincosf(angle[i], sin_out[i], cos_out[i]);
}

extern "C" int ft_max_keypoints(ft_context* c) { return c ? c->P.maxKp : 0; }

// Frame constructor in one call (reference src/Frame.cc:102-223 for pinhole rigs, :1115-1229 for fisheye):
// upload both images, extract, stereo-match, and bring every host vector the constructor fills back with a
// single synchronisation. Output arrays must hold ft_max_keypoints() entries.
// Asynchronous half of the Frame constructor: upload + extract + stereo + the D2H of the result slab are enqueued and
// the call returns; ft_frame_collect waits and scatters. With two contexts a caller overlaps the extraction of frame
// t+1 with the tracking of frame t (extraction does not depend on the SLAM state).
extern "C" ft_status ft_frame_submit(ft_context* c, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR) {
  if (!c) { set_err("ft_frame_submit: null context"); return FT_ERR_INVALID; }
  ft_status st = upload_images(c, imgL, stepL, imgR, stepR);
  if (st != FT_OK) return st;
  st = run_frame(c);
  if (st != FT_OK) return st;
  // one D2H of the whole result slab (header, both eyes' keypoints + descriptors, stereo outputs)
  const size_t bytes = c->fisheye ? c->frameBytesAll : c->frameBytesPinhole;
  CK(cudaMemcpyAsync(c->hFrame, c->dFrame, bytes, cudaMemcpyDeviceToHost, c->stream));
  c->framePending = true;
  return FT_OK;
}

extern "C" ft_status ft_frame_collect(ft_context* c, ft_keypoint* kpsL, uint8_t* descL, ft_keypoint* kpsR, uint8_t* descR,
                                      int* counts4 /* nL, monoL, nR, monoR */, float* u_right, float* depth,
                                      int* l2r, int* r2l, float* p3d) {
  if (!c || !counts4) { set_err("ft_frame_collect: null argument"); return FT_ERR_INVALID; }
  if (!c->framePending) { set_err("ft_frame_collect: no submitted frame"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  c->framePending = false;
  memcpy(c->hCounts, c->hFrame, 8 * sizeof(int));
  c->countsValid = true;
  const int nl = c->hCounts[0], nr = c->hCounts[2];
  if (kpsL) memcpy(kpsL, c->hFrame + c->offKps[0], sizeof(ft_keypoint) * nl);
  if (kpsR) memcpy(kpsR, c->hFrame + c->offKps[1], sizeof(ft_keypoint) * nr);
  if (descL) memcpy(descL, c->hFrame + c->offDesc[0], (size_t)32 * nl);
  if (descR) memcpy(descR, c->hFrame + c->offDesc[1], (size_t)32 * nr);
  if (u_right) memcpy(u_right, c->hFrame + c->offUR, sizeof(float) * nl);
  if (depth) memcpy(depth, c->hFrame + c->offDepth, sizeof(float) * nl);
  if (c->fisheye) {
    if (l2r) memcpy(l2r, c->hFrame + c->offL2R, sizeof(int) * nl);
    if (r2l) memcpy(r2l, c->hFrame + c->offR2L, sizeof(int) * nr);
    if (p3d) memcpy(p3d, c->hFrame + c->offP3D, sizeof(float) * 3 * nl);
  }
  for (int i = 0; i < 4; i++) counts4[i] = c->hCounts[i];
  c->frameStatus |= c->hCounts[4];
  return check_device_status(c, c->frameStatus);
}

extern "C" ft_status ft_frame_construct(ft_context* c, const uint8_t* imgL, int stepL, const uint8_t* imgR, int stepR,
                                        ft_keypoint* kpsL, uint8_t* descL, ft_keypoint* kpsR, uint8_t* descR,
                                        int* counts4 /* nL, monoL, nR, monoR */, float* u_right, float* depth,
                                        int* l2r, int* r2l, float* p3d) {
  if (!c || !counts4) { set_err("ft_frame_construct: null argument"); return FT_ERR_INVALID; }
  ft_status st = ft_frame_submit(c, imgL, stepL, imgR, stepR);
  if (st != FT_OK) return st;
  return ft_frame_collect(c, kpsL, descL, kpsR, descR, counts4, u_right, depth, l2r, r2l, p3d);
}

extern "C" ft_status ft_set_pose(ft_context* c, const float* Rcw, const float* tcw, const float* Rwc, const float* Ow) {
  if (!c || !Rcw || !tcw) { set_err("ft_set_pose: null argument"); return FT_ERR_INVALID; }
  memcpy(c->pose.Rcw, Rcw, 36); memcpy(c->pose.tcw, tcw, 12);
  if (Rwc) memcpy(c->pose.Rwc, Rwc, 36);
  else for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c->pose.Rwc[3 * i + j] = Rcw[3 * j + i];
  if (Ow) memcpy(c->pose.Ow, Ow, 12);
  else for (int i = 0; i < 3; i++)
    c->pose.Ow[i] = -(c->pose.Rwc[3 * i] * tcw[0] + c->pose.Rwc[3 * i + 1] * tcw[1] + c->pose.Rwc[3 * i + 2] * tcw[2]);
  return FT_OK;
}

// Tight layout of one map-point snapshot behind the fixed holder region of the staging buffers.
static void mp_layout(ft_context* c, int M, uint8_t* base, float** pos, float** normal, float** minmax, uint8_t** desc,
                      int** flags) {
  uint8_t* p = base + c->mpHolderBytes;
  *pos = reinterpret_cast<float*>(p);
  *normal = reinterpret_cast<float*>(p + (size_t)12 * M);
  *minmax = reinterpret_cast<float*>(p + (size_t)24 * M);
  *desc = p + (size_t)32 * M;
  *flags = reinterpret_cast<int*>(p + (size_t)64 * M);
}
static void mp_bind_device(ft_context* c, int M) {
  FtSbpBuffers& Q = c->Q;
  mp_layout(c, M, c->dMp, &Q.pos, &Q.normal, &Q.minmax, &Q.desc, &Q.flags);
  Q.slot = nullptr;
  c->residentM = M;
}

extern "C" ft_status ft_map_point_staging(ft_context* c, int M, float** pos, float** normal, float** minmax,
                                          uint8_t** desc, int** flags, int** holder, uint8_t** holder_obs) {
  if (!c || M < 0) { set_err("ft_map_point_staging: bad argument"); return FT_ERR_INVALID; }
  if (M > c->cfg.max_map_points) {
    set_err("more map points than ft_config.max_map_points (the reference raises SIGSEGV beyond 25000)");
    return FT_ERR_CAPACITY;
  }
  float *p, *n, *mm; uint8_t* d; int* f;
  mp_layout(c, M, c->hMp, &p, &n, &mm, &d, &f);
  if (pos) *pos = p;
  if (normal) *normal = n;
  if (minmax) *minmax = mm;
  if (desc) *desc = d;
  if (flags) *flags = f;
  if (holder) *holder = reinterpret_cast<int*>(c->hMp);
  if (holder_obs) *holder_obs = c->hMp + (size_t)2 * c->P.maxKp * 4;
  return FT_OK;
}

extern "C" ft_status ft_upload_map_points(ft_context* c, int M, const float* pos, const float* normal,
                                          const float* minmax, const uint8_t* desc, const int* flags) {
  if (!c || M < 0 || (M > 0 && (!pos || !normal || !minmax || !desc || !flags))) { set_err("ft_upload_map_points: null argument"); return FT_ERR_INVALID; }
  float *p, *n, *mm; uint8_t* d; int* f;
  ft_status st = ft_map_point_staging(c, M, &p, &n, &mm, &d, &f, nullptr, nullptr);
  if (st != FT_OK) return st;
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));   // the staging buffer may still be in flight
  if (M) {
    if (pos != p) memcpy(p, pos, sizeof(float) * 3 * M);
    if (normal != n) memcpy(n, normal, sizeof(float) * 3 * M);
    if (minmax != mm) memcpy(mm, minmax, sizeof(float) * 2 * M);
    if (desc != d) memcpy(d, desc, (size_t)32 * M);
    if (flags != f) memcpy(f, flags, sizeof(int) * M);
    CK(cudaMemcpyAsync(c->dMp + c->mpHolderBytes, c->hMp + c->mpHolderBytes, (size_t)68 * M, cudaMemcpyHostToDevice, c->stream));
    c->mpStagedValid = false;   // a copy out of the staging buffer that no event covers: the next user synchronises the stream
  }
  mp_bind_device(c, M);
  return FT_OK;
}

// Bind a map-point snapshot that already lives in device memory (e.g. a persistent device-side map store kept up to
// date by the mapping thread): no copy, the search kernels read the caller's arrays directly.
extern "C" ft_status ft_bind_map_points_device(ft_context* c, int M, const float* d_pos, const float* d_normal,
                                               const float* d_minmax, const uint8_t* d_desc, const int* d_flags) {
  if (!c || M < 0 || (M > 0 && (!d_pos || !d_normal || !d_minmax || !d_desc || !d_flags))) { set_err("ft_bind_map_points_device: null argument"); return FT_ERR_INVALID; }
  if (M > c->cfg.max_map_points) { set_err("more map points than ft_config.max_map_points"); return FT_ERR_CAPACITY; }
  if ((reinterpret_cast<uintptr_t>(d_desc) & 15) != 0) { set_err("ft_bind_map_points_device: descriptors must be 16-byte aligned"); return FT_ERR_INVALID; }
  FtSbpBuffers& Q = c->Q;
  Q.pos = const_cast<float*>(d_pos); Q.normal = const_cast<float*>(d_normal); Q.minmax = const_cast<float*>(d_minmax);
  Q.desc = const_cast<uint8_t*>(d_desc); Q.flags = const_cast<int*>(d_flags);
  Q.slot = nullptr;
  c->residentM = M;
  return FT_OK;
}

extern "C" ft_status ft_upload_holders(ft_context* c, int N, const int* holder, const uint8_t* holderObs) {
  if (!c || N < 0 || N > 2 * c->P.maxKp) { set_err("ft_upload_holders: bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  cudaStream_t s = c->stream;
  if (!holder || !holderObs) {   // F.mvpMapPoints all NULL, as right after the Frame constructor (Frame.cc:173)
    CK(cudaMemsetAsync(c->holderInit, 0xFF, sizeof(int) * 2 * c->P.maxKp, s));
    CK(cudaMemsetAsync(c->holderObsInit, 0, (size_t)2 * c->P.maxKp, s));
    return FT_OK;
  }
  CK(cudaMemcpyAsync(c->holderInit, holder, sizeof(int) * N, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(c->holderObsInit, holderObs, (size_t)N, cudaMemcpyHostToDevice, s));
  return FT_OK;
}

static ft_status search_run(ft_context* c, float th, int bFar, float thFar, float nnratio, int mode, int direction, int checkOri);

extern "C" ft_status ft_search_resident(ft_context* c, float th, int bFar, float thFar, float nnratio) {
  return search_run(c, th, bFar, thFar, nnratio, 0, 0, 0);
}

static ft_status search_run(ft_context* c, float th, int bFar, float thFar, float nnratio, int mode, int direction, int checkOri) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (!c->extracted) { set_err("ft_search_resident: no extracted frame"); return FT_ERR_STATE; }
  if (!c->stereoDone) { set_err("ft_search_resident: stereo matching has not run (mvuRight / match tables are read)"); return FT_ERR_STATE; }
  if (c->searchPending) { set_err("a search submitted with ft_search_store_submit has not been collected (ft_search_collect)"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  FtSbpBuffers& Q = c->Q;
  cudaStream_t s = c->stream;
  const int M = c->residentM;
  c->lastM = M;
  c->nLaunchSearch = 0;
  if (M == 0) {   // nothing to project: the holders come back unchanged
    CK(cudaMemcpyAsync(Q.holder, c->holderInit, sizeof(int) * 2 * c->P.maxKp, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(Q.holderObs, c->holderObsInit, (size_t)2 * c->P.maxKp, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemsetAsync(Q.cursor, 0, 16 * sizeof(int), s));
    return FT_OK;
  }
  FtFrustumArgs fa;
  fa.cam1 = c->cam1; fa.cam2 = c->cam2; fa.pose = c->pose;
  fa.minX = c->minX; fa.maxX = c->maxX; fa.minY = c->minY; fa.maxY = c->maxY;
  fa.mbf = c->mbf; fa.logScale = c->logScale; fa.nlevels = c->P.nlevels; fa.fisheye = c->fisheye;
  fa.viewCosLimit = 0.5f;   // Tracking.cc:3512
  FtGatherArgs ga;
  ga.minX = c->minX; ga.minY = c->minY; ga.gridWInv = c->gridWInv; ga.gridHInv = c->gridHInv;
  ga.th = th; ga.bFactor = (th != 1.0f); ga.bFar = bFar; ga.thFar = thFar; ga.fisheye = c->fisheye;
  ga.mode = mode; ga.direction = direction;
  FtResolveArgs ra;
  ra.M = M; ra.nLeft = 0; ra.nSlots = 2 * c->P.maxKp; ra.fisheye = c->fisheye; ra.nnratio = nnratio;   // nSlots: smem sizing bound
  ra.mode = mode; ra.checkOri = checkOri;
  if (c->useGraph && !c->timing && c->searchGraph) {
    CK(ft_search_graph_run(&c->gSearch, c->P, c->B, c->G, c->S, Q, fa, ga, ra, M, s));
  } else {
    { StageScope t(c, FT_STAGE_GATHER, s); ft_launch_gather(c->P, c->B, c->G, c->S, Q, fa, ga, M, s); }
    { StageScope t(c, FT_STAGE_RESOLVE, s); ft_launch_resolve(c->B, Q, c->S, ra, s); }
  }
  c->nLaunchSearch = 2 + (c->fisheye ? 1 : 0);
  CK(cudaGetLastError());
  return FT_OK;
}

// one D2H of the search result slab (cursors + status, holders, per-point selections), then host scatter
static ft_status search_fetch(ft_context* c, int N, int* holder, uint8_t* holderObs, int* best_idx, int* nmatches) {
  cudaStream_t s = c->stream;
  const size_t bytes = c->offOutSel + (best_idx ? (size_t)8 * c->lastM : 0);
  CK(cudaMemcpyAsync(c->hOut, c->dOut, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int* hdr = reinterpret_cast<const int*>(c->hOut);
  memcpy(c->hCounts + 8, hdr, 8 * sizeof(int));
  if (holder && N) memcpy(holder, c->hOut + c->offOutHolder, sizeof(int) * N);
  if (holderObs && N) memcpy(holderObs, c->hOut + c->offOutObs, (size_t)N);
  if (best_idx && c->lastM) memcpy(best_idx, c->hOut + c->offOutSel, sizeof(int) * 2 * c->lastM);
  if (nmatches) *nmatches = hdr[1];
  return check_device_status(c, hdr[8]);
}

extern "C" ft_status ft_search_download(ft_context* c, int* holder, uint8_t* holderObs, int* best_idx, int* nmatches) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (c->searchPending) { set_err("a search submitted with ft_search_store_submit has not been collected (ft_search_collect)"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);
  if (st != FT_OK) return st;
  const int N = c->fisheye ? c->hCounts[0] + c->hCounts[2] : c->hCounts[0];
  return search_fetch(c, N, holder, holderObs, best_idx, nmatches);
}

// Search over a snapshot the caller wrote into the pinned staging buffers returned by ft_map_point_staging:
// one H2D (holders + snapshot), two kernels, one D2H.
extern "C" ft_status ft_search_staged(ft_context* c, int M, float th, int bFar, float thFar, float nnratio,
                                      const int** holder_out, const uint8_t** holder_obs_out, const int** best_idx_out,
                                      int* nmatches) {
  if (!c || M < 0 || M > c->cfg.max_map_points) { set_err("ft_search_staged: bad argument"); return FT_ERR_INVALID; }
  if (!c->extracted || !c->stereoDone) { set_err("ft_search_staged: no stereo-matched frame"); return FT_ERR_STATE; }
  if (c->searchPending) { set_err("a search submitted with ft_search_store_submit has not been collected (ft_search_collect)"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  if (nmatches) *nmatches = 0;
  CK(cudaMemcpyAsync(c->dMp, c->hMp, c->mpHolderBytes + (size_t)68 * M, cudaMemcpyHostToDevice, c->stream));
  mp_bind_device(c, M);
  ft_status st = ft_search_resident(c, th, bFar, thFar, nnratio);
  if (st != FT_OK) return st;
  {
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->hOut, c->dOut, c->offOutSel + (size_t)8 * M, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int* hdr = reinterpret_cast<const int*>(c->hOut);
    memcpy(c->hCounts + 8, hdr, 8 * sizeof(int));
    if (holder_out) *holder_out = reinterpret_cast<const int*>(c->hOut + c->offOutHolder);
    if (holder_obs_out) *holder_obs_out = c->hOut + c->offOutObs;
    if (best_idx_out) *best_idx_out = reinterpret_cast<const int*>(c->hOut + c->offOutSel);
    if (nmatches) *nmatches = hdr[1];
    return check_device_status(c, hdr[8]);
  }
}

extern "C" ft_status ft_search_local_points(ft_context* c, int M, const float* pos, const float* normal,
                                            const float* minmax, const uint8_t* desc, const int* flags, float th,
                                            int bFar, float thFar, float nnratio, int* holder, uint8_t* holderObs,
                                            int* best_idx, int* nmatches) {
  if (!c || !holder || !holderObs || M < 0 || (M > 0 && (!pos || !normal || !minmax || !desc || !flags))) {
    set_err("ft_search_local_points: null argument"); return FT_ERR_INVALID;
  }
  if (!c->extracted) { set_err("ft_search_local_points: no extracted frame"); return FT_ERR_STATE; }
  if (!c->stereoDone) { set_err("ft_search_local_points: stereo matching has not run (mvuRight / match tables are read)"); return FT_ERR_STATE; }
  if (c->searchPending) { set_err("a search submitted with ft_search_store_submit has not been collected (ft_search_collect)"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);   // synchronises: the staging buffers are free afterwards
  if (st != FT_OK) return st;
  const int N = c->fisheye ? c->hCounts[0] + c->hCounts[2] : c->hCounts[0];
  if (nmatches) *nmatches = 0;
  float *p, *n, *mm; uint8_t* d; int *f, *h; uint8_t* ho;
  st = ft_map_point_staging(c, M, &p, &n, &mm, &d, &f, &h, &ho);
  if (st != FT_OK) return st;
  if (M == 0 || N == 0) { c->lastM = 0; return FT_OK; }
  CK(cudaStreamSynchronize(c->stream));
  memcpy(p, pos, sizeof(float) * 3 * M); memcpy(n, normal, sizeof(float) * 3 * M); memcpy(mm, minmax, sizeof(float) * 2 * M);
  memcpy(d, desc, (size_t)32 * M); memcpy(f, flags, sizeof(int) * M);
  memcpy(h, holder, sizeof(int) * N); memcpy(ho, holderObs, (size_t)N);
  CK(cudaMemcpyAsync(c->dMp, c->hMp, c->mpHolderBytes + (size_t)68 * M, cudaMemcpyHostToDevice, c->stream));
  mp_bind_device(c, M);
  st = ft_search_resident(c, th, bFar, thFar, nnratio);
  if (st != FT_OK) return st;
  return search_fetch(c, N, holder, holderObs, best_idx, nmatches);
}

// ---- persistent map store (replaces the per-frame CudaMapPoint marshalling of the reference,
// src/Kernels/CudaWrappers/CudaMapPoint.cc:15-34 + src/Kernels/SearchLocalPointsKernel.cu:368-409) ----
extern "C" ft_status ft_map_store_create(ft_context* c, int capacity) {
  if (!c || capacity <= 0) { set_err("ft_map_store_create: bad argument"); return FT_ERR_INVALID; }
  if (c->store) { set_err("ft_map_store_create: the context already has a map store"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_map_store* st = new ft_map_store();
  st->cap = capacity; st->device = c->cfg.device_id;
  cudaError_t e = cudaMalloc((void**)&st->pos, (size_t)12 * capacity);
  if (e == cudaSuccess) e = cudaMalloc((void**)&st->normal, (size_t)12 * capacity);
  if (e == cudaSuccess) e = cudaMalloc((void**)&st->minmax, (size_t)8 * capacity);
  if (e == cudaSuccess) e = cudaMalloc((void**)&st->desc, (size_t)32 * capacity);
  if (e == cudaSuccess && !c->storeUpdated) e = cudaEventCreateWithFlags(&c->storeUpdated, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    cudaFree(st->pos); cudaFree(st->normal); cudaFree(st->minmax); cudaFree(st->desc);
    delete st;
    set_err(std::string("ft_map_store_create: ") + cudaGetErrorString(e));
    return FT_ERR_CUDA;
  }
  cudaMemsetAsync(st->pos, 0, (size_t)12 * capacity, c->stream); cudaMemsetAsync(st->normal, 0, (size_t)12 * capacity, c->stream);
  cudaMemsetAsync(st->minmax, 0, (size_t)8 * capacity, c->stream); cudaMemsetAsync(st->desc, 0, (size_t)32 * capacity, c->stream);
  CK(cudaEventRecord(c->storeUpdated, c->stream));
  c->storeUpdatedValid = true;
  st->users.push_back(c);
  c->store = st;
  return FT_OK;
}

extern "C" ft_status ft_map_store_attach(ft_context* c, ft_context* owner) {
  if (!c || !owner || c == owner) { set_err("ft_map_store_attach: bad argument"); return FT_ERR_INVALID; }
  if (!owner->store) { set_err("ft_map_store_attach: the other context has no map store"); return FT_ERR_STATE; }
  if (c->store) { set_err("ft_map_store_attach: the context already has a map store"); return FT_ERR_STATE; }
  if (owner->store->device != c->cfg.device_id) { set_err("ft_map_store_attach: contexts live on different devices"); return FT_ERR_INVALID; }
  c->store = owner->store;
  {
    std::lock_guard<std::mutex> lk(c->store->mu);
    c->store->users.push_back(c);
  }
  return FT_OK;
}

extern "C" ft_status ft_map_store_update(ft_context* c, int n, const int* slots, const float* pos, const float* normal,
                                         const float* minmax, const uint8_t* desc) {
  if (!c || n < 0 || (n > 0 && (!slots || !pos || !normal || !minmax || !desc))) { set_err("ft_map_store_update: null argument"); return FT_ERR_INVALID; }
  ft_map_store* st = c->store;
  if (!st) { set_err("ft_map_store_update: no map store (ft_map_store_create / ft_map_store_attach first)"); return FT_ERR_STATE; }
  if (n == 0) return FT_OK;
  for (int i = 0; i < n; i++)
    if (slots[i] < 0 || slots[i] >= st->cap) { set_err("ft_map_store_update: slot outside the store's capacity"); return FT_ERR_CAPACITY; }
  CK(cudaSetDevice(c->cfg.device_id));
  if (c->updStagedValid) CK(cudaEventSynchronize(c->updStaged));   // the previous update may still be reading the staging buffer
  if (n > c->updCap) {
    const int cap = std::max(n, std::max(1024, 2 * c->updCap));
    if (c->hUpd) cudaFreeHost(c->hUpd);
    if (c->dUpd) { CK(cudaStreamSynchronize(c->stream)); cudaFree(c->dUpd); }
    c->hUpd = nullptr; c->dUpd = nullptr; c->updCap = 0;
    CK(cudaMallocHost((void**)&c->hUpd, (size_t)72 * cap));
    CK(cudaMalloc((void**)&c->dUpd, (size_t)72 * cap));
    c->updCap = cap;
  }
  if (!c->updStaged) CK(cudaEventCreateWithFlags(&c->updStaged, cudaEventDisableTiming));
  uint8_t* h = c->hUpd;
  memcpy(h, slots, (size_t)4 * n); memcpy(h + (size_t)4 * n, pos, (size_t)12 * n); memcpy(h + (size_t)16 * n, normal, (size_t)12 * n);
  memcpy(h + (size_t)28 * n, minmax, (size_t)8 * n); memcpy(h + (size_t)36 * n, desc, (size_t)32 * n);
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(c->dUpd, h, (size_t)68 * n + (size_t)4 * n, cudaMemcpyHostToDevice, s));
  CK(cudaEventRecord(c->updStaged, s));
  c->updStagedValid = true;
  {
    std::lock_guard<std::mutex> lk(st->mu);
    if (!c->storeUpdated) CK(cudaEventCreateWithFlags(&c->storeUpdated, cudaEventDisableTiming));
    // rows may be read by a search in flight on another context of the sequence, or written by an update still in flight
    // on another context's stream: wait for the last search and the last update of each
    for (ft_context* u : st->users) {
      if (u == c) continue;
      if (u->storeSearchedValid) CK(cudaStreamWaitEvent(s, u->storeSearched, 0));
      if (u->storeUpdatedValid) CK(cudaStreamWaitEvent(s, u->storeUpdated, 0));
    }
    ft_launch_store_scatter(n, c->dUpd, st->pos, st->normal, st->minmax, st->desc, s);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->storeUpdated, s));
    c->storeUpdatedValid = true;
  }
  return FT_OK;
}

// Tracking::SearchLocalPoints against the store: the local map is a list of rows (mvpLocalMapPoints order) plus the
// per-call flags; 8 bytes per map point cross PCIe instead of 68.
// ft_search_store_submit enqueues everything (H2D of holders + rows + flags, gather -> resolve, D2H of the result slab) and
// returns; ft_search_collect waits for the slab and scatters it into the caller's arrays. Between the two the tracking
// thread is free, e.g. to submit the next camera frame on another context (ft_sequence_driver.cpp does that).
extern "C" ft_status ft_search_store_submit(ft_context* c, int M, const int* slots, const int* flags, float th, int bFar,
                                            float thFar, float nnratio, const int* holder, const uint8_t* holderObs,
                                            int want_best_idx) {
  if (!c || !holder || !holderObs || M < 0 || (M > 0 && (!slots || !flags))) { set_err("ft_search_store: null argument"); return FT_ERR_INVALID; }
  ft_map_store* st = c->store;
  if (!st) { set_err("ft_search_store: no map store (ft_map_store_create / ft_map_store_attach first)"); return FT_ERR_STATE; }
  if (M > c->cfg.max_map_points) { set_err("more map points than ft_config.max_map_points (the reference raises SIGSEGV beyond 25000)"); return FT_ERR_CAPACITY; }
  if (!c->extracted || !c->stereoDone) { set_err("ft_search_store: no stereo-matched frame"); return FT_ERR_STATE; }
  if (c->searchPending) { set_err("ft_search_store_submit: the previous submitted search has not been collected"); return FT_ERR_STATE; }
  for (int i = 0; i < M; i++)
    if (slots[i] < 0 || slots[i] >= st->cap) { set_err("ft_search_store: slot outside the store's capacity"); return FT_ERR_CAPACITY; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status rs = fetch_counts(c);
  if (rs != FT_OK) return rs;
  const int N = c->fisheye ? c->hCounts[0] + c->hCounts[2] : c->hCounts[0];
  c->searchN = N; c->searchWantBest = want_best_idx != 0;
  c->searchEmpty = false;
  if (M == 0 || N == 0) { c->lastM = 0; c->searchEmpty = true; return FT_OK; }   // nothing enqueued: the caller's holders are the result
  cudaStream_t s = c->stream;
  // the staging buffer is free once the last H2D that read it has finished: wait for exactly that (an upsert or a frame
  // enqueued on this stream since then is not waited for); copies enqueued by the other search calls are followed by a
  // stream synchronisation inside those calls, so the event of the last split search is the only one that can be open
  if (c->mpStagedValid) CK(cudaEventSynchronize(c->mpStaged));
  else CK(cudaStreamSynchronize(s));
  memcpy(c->hMp, holder, sizeof(int) * N);
  memcpy(c->hMp + (size_t)2 * c->P.maxKp * 4, holderObs, (size_t)N);
  uint8_t* hs = c->hMp + c->mpHolderBytes;
  memcpy(hs, slots, (size_t)4 * M); memcpy(hs + (size_t)4 * M, flags, (size_t)4 * M);
  CK(cudaMemcpyAsync(c->dMp, c->hMp, c->mpHolderBytes + (size_t)8 * M, cudaMemcpyHostToDevice, s));
  if (!c->mpStaged) CK(cudaEventCreateWithFlags(&c->mpStaged, cudaEventDisableTiming));
  CK(cudaEventRecord(c->mpStaged, s));
  c->mpStagedValid = true;
  FtSbpBuffers& Q = c->Q;
  Q.pos = st->pos; Q.normal = st->normal; Q.minmax = st->minmax; Q.desc = st->desc;
  Q.slot = reinterpret_cast<const int*>(c->dMp + c->mpHolderBytes);
  Q.flags = reinterpret_cast<int*>(c->dMp + c->mpHolderBytes + (size_t)4 * M);
  c->residentM = M;
  {
    std::lock_guard<std::mutex> lk(st->mu);
    // every update outstanding on any attached context's stream lands before the search reads the rows
    for (ft_context* u : st->users)
      if (u != c && u->storeUpdatedValid) CK(cudaStreamWaitEvent(s, u->storeUpdated, 0));
    rs = ft_search_resident(c, th, bFar, thFar, nnratio);
    if (rs != FT_OK) return rs;
    if (!c->storeSearched) CK(cudaEventCreateWithFlags(&c->storeSearched, cudaEventDisableTiming));
    CK(cudaEventRecord(c->storeSearched, s));
    c->storeSearchedValid = true;
  }
  // one D2H of the search result slab (cursors + status, holders, per-point selections)
  const size_t bytes = c->offOutSel + (c->searchWantBest ? (size_t)8 * c->lastM : 0);
  CK(cudaMemcpyAsync(c->hOut, c->dOut, bytes, cudaMemcpyDeviceToHost, s));
  if (!c->searchDone) CK(cudaEventCreateWithFlags(&c->searchDone, cudaEventDisableTiming));
  CK(cudaEventRecord(c->searchDone, s));
  c->searchPending = true;
  return FT_OK;
}

extern "C" ft_status ft_search_collect(ft_context* c, int* holder, uint8_t* holderObs, int* best_idx, int* nmatches) {
  if (!c) { set_err("ft_search_collect: null context"); return FT_ERR_INVALID; }
  if (nmatches) *nmatches = 0;
  if (!c->searchPending) {
    // the submit found nothing to search (no map points or no keypoints): the holders handed to it are the result, the
    // output arrays are left as they are
    if (c->searchEmpty) { c->searchEmpty = false; return FT_OK; }
    set_err("ft_search_collect: no submitted search"); return FT_ERR_STATE;
  }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaEventSynchronize(c->searchDone));
  c->searchPending = false;
  const int N = c->searchN;
  const int* hdr = reinterpret_cast<const int*>(c->hOut);
  memcpy(c->hCounts + 8, hdr, 8 * sizeof(int));
  if (holder && N) memcpy(holder, c->hOut + c->offOutHolder, sizeof(int) * N);
  if (holderObs && N) memcpy(holderObs, c->hOut + c->offOutObs, (size_t)N);
  if (best_idx && c->searchWantBest && c->lastM) memcpy(best_idx, c->hOut + c->offOutSel, sizeof(int) * 2 * c->lastM);
  if (nmatches) *nmatches = hdr[1];
  return check_device_status(c, hdr[8]);
}

extern "C" ft_status ft_search_store(ft_context* c, int M, const int* slots, const int* flags, float th, int bFar,
                                     float thFar, float nnratio, int* holder, uint8_t* holderObs, int* best_idx,
                                     int* nmatches) {
  if (nmatches) *nmatches = 0;
  ft_status rs = ft_search_store_submit(c, M, slots, flags, th, bFar, thFar, nnratio, holder, holderObs, best_idx != nullptr);
  if (rs != FT_OK) return rs;
  if (!c->searchPending) { c->searchEmpty = false; return FT_OK; }   // M == 0 or no keypoints
  return ft_search_collect(c, holder, holderObs, best_idx, nmatches);
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (reference src/ORBmatcher.cc:1775-2085), the
// frame-to-last-frame search of Tracking::TrackWithMotionModel (src/Tracking.cc:2911-2990). One entry per last-frame
// keypoint that holds a map point: world position + descriptor of the map point, octave + angle of that keypoint.
extern "C" ft_status ft_search_last_frame(ft_context* c, int n, const float* pos, const uint8_t* desc, const int* octave,
                                          const float* angle, const int* flags, const float* Rlw, const float* tlw,
                                          float th, int bMono, int checkOrientation, int* holder, uint8_t* holderObs,
                                          int* best_idx, int* nmatches) {
  if (!c || n < 0 || !holder || !holderObs || (n > 0 && (!pos || !desc || !octave || !angle || !flags)) || !Rlw || !tlw) {
    set_err("ft_search_last_frame: null argument"); return FT_ERR_INVALID;
  }
  if (!c->extracted || !c->stereoDone) { set_err("ft_search_last_frame: no stereo-matched frame"); return FT_ERR_STATE; }
  if (c->searchPending) { set_err("a search submitted with ft_search_store_submit has not been collected (ft_search_collect)"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);
  if (st != FT_OK) return st;
  const int N = c->fisheye ? c->hCounts[0] + c->hCounts[2] : c->hCounts[0];
  if (nmatches) *nmatches = 0;
  for (int i = 0; i < n; i++)
    if (octave[i] < 0 || octave[i] >= c->P.nlevels) { set_err("ft_search_last_frame: octave out of range"); return FT_ERR_INVALID; }
  float *p, *nn, *mm; uint8_t* d; int *f, *h; uint8_t* ho;
  st = ft_map_point_staging(c, n, &p, &nn, &mm, &d, &f, &h, &ho);
  if (st != FT_OK) return st;
  if (n == 0 || N == 0) { c->lastM = 0; return FT_OK; }
  // bForward / bBackward (ORBmatcher.cc:1789-1794): tlc = Tlw * twc, twc = camera centre of the current frame
  float tlcz = Rlw[6] * c->pose.Ow[0] + Rlw[7] * c->pose.Ow[1] + Rlw[8] * c->pose.Ow[2] + tlw[2];
  const int direction = bMono ? 0 : (tlcz > c->mb ? 1 : (-tlcz > c->mb ? -1 : 0));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(p, pos, sizeof(float) * 3 * n);
  memset(nn, 0, sizeof(float) * 3 * n);
  for (int i = 0; i < n; i++) { mm[2 * i] = angle[i]; mm[2 * i + 1] = (float)octave[i]; }
  memcpy(d, desc, (size_t)32 * n); memcpy(f, flags, sizeof(int) * n);
  memcpy(h, holder, sizeof(int) * N); memcpy(ho, holderObs, (size_t)N);
  CK(cudaMemcpyAsync(c->dMp, c->hMp, c->mpHolderBytes + (size_t)68 * n, cudaMemcpyHostToDevice, c->stream));
  mp_bind_device(c, n);
  st = search_run(c, th, 0, 0.f, 1.0f, 1, direction, checkOrientation ? 1 : 0);
  if (st != FT_OK) return st;
  return search_fetch(c, N, holder, holderObs, best_idx, nmatches);
}

// ---- bag of words: Frame::ComputeBoW / ORBmatcher::SearchByBoW on the device-resident frame (kernels in ft_bow.cu) ----
static FtBowSource bow_source(ft_context* c) {
  FtBowSource S = {};
  S.desc0 = c->B.eye[0].desc; S.cnt0 = c->B.eye[0].counts; S.kps0 = c->B.eye[0].kps;
  if (c->fisheye) { S.desc1 = c->B.eye[1].desc; S.cnt1 = c->B.eye[1].counts; S.kps1 = c->B.eye[1].kps; }
  return S;
}

extern "C" ft_status ft_compute_bow(ft_context* c, ft_vocabulary* voc, int levelsup) {
  if (!c || !voc || levelsup < 0) { set_err("ft_compute_bow: bad argument"); return FT_ERR_INVALID; }
  if (!c->extracted) { set_err("ft_compute_bow: no extracted frame"); return FT_ERR_STATE; }
  if (ft_vocabulary_device(voc) != c->cfg.device_id) { set_err("ft_compute_bow: vocabulary lives on another device"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  const int cap = c->fisheye ? 2 * c->P.maxKp : c->P.maxKp;
  if (cap > 16384) { set_err("ft_compute_bow: more than 16384 features per frame"); return FT_ERR_CAPACITY; }
  if (!c->bowCap) {
    CK(ft_bow_frame_alloc(&c->W, cap, c->allocs));
    c->bowCap = cap;
  }
  c->nLaunchBow = ft_launch_bow_transform(voc, bow_source(c), c->W, cap, levelsup, c->stream);
  CK(cudaGetLastError());
  c->bowValid = true;
  return FT_OK;
}

extern "C" ft_status ft_bow_download(ft_context* c, int cap, int* word_id, int* node_id, uint32_t* bow_ids, double* bow_vals,
                                     int* n_bow, int* n) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (!c->bowValid) { set_err("ft_bow_download: ft_compute_bow has not run on this frame"); return FT_ERR_STATE; }
  CK(cudaSetDevice(c->cfg.device_id));
  int meta[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(meta, c->W.meta, sizeof(meta), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int N = meta[0], nd = meta[2];
  if (n) *n = N;
  if (n_bow) *n_bow = nd;
  if (((word_id || node_id) && cap < N) || ((bow_ids || bow_vals) && cap < nd)) { set_err("ft_bow_download: capacity too small"); return FT_ERR_INVALID; }
  if (word_id && N) CK(cudaMemcpyAsync(word_id, c->W.word, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
  if (node_id && N) CK(cudaMemcpyAsync(node_id, c->W.node, sizeof(int) * N, cudaMemcpyDeviceToHost, c->stream));
  if (bow_ids && nd) CK(cudaMemcpyAsync(bow_ids, c->W.bowIds, sizeof(uint32_t) * nd, cudaMemcpyDeviceToHost, c->stream));
  if (bow_vals && nd) CK(cudaMemcpyAsync(bow_vals, c->W.bowVals, sizeof(double) * nd, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FT_OK;
}

extern "C" ft_status ft_search_by_bow(ft_context* c, int n_kf, const uint8_t* kf_desc, const float* kf_angle, const int* kf_node,
                                      const uint8_t* kf_has_mp, float nnratio, int check_orientation, int* match, int* nmatches) {
  if (!c || n_kf < 0 || !match || (n_kf > 0 && (!kf_desc || !kf_angle || !kf_node || !kf_has_mp))) {
    set_err("ft_search_by_bow: null argument"); return FT_ERR_INVALID;
  }
  if (!c->extracted || !c->bowValid) { set_err("ft_search_by_bow: ft_compute_bow has not run on this frame"); return FT_ERR_STATE; }
  if (n_kf > 16384) { set_err("ft_search_by_bow: more than 16384 KeyFrame features"); return FT_ERR_CAPACITY; }
  CK(cudaSetDevice(c->cfg.device_id));
  ft_status st = fetch_counts(c);
  if (st != FT_OK) return st;
  const int N = c->fisheye ? c->hCounts[0] + c->hCounts[2] : c->hCounts[0];
  if (nmatches) *nmatches = 0;
  for (int i = 0; i < N; i++) match[i] = -1;
  if (n_kf == 0 || N == 0) return FT_OK;
  int capKF = 1024;
  while (capKF < n_kf) capKF <<= 1;
  CK(ft_bow_search_alloc(&c->WQ, c->bowCap, capKF, c->allocs));
  // pinned staging: the KeyFrame side goes up in one H2D, (nmatches, match[N]) comes back in one D2H
  const size_t need = std::max((size_t)41 * capKF + 64, sizeof(int) * ((size_t)c->bowCap + 4));
  if (need > c->hBowBytes) {
    if (c->hBow) cudaFreeHost(c->hBow);
    c->hBow = nullptr; c->hBowBytes = 0;
    CK(cudaMallocHost((void**)&c->hBow, need));
    c->hBowBytes = need;
  }
  const size_t blob = ft_bow_search_bind(&c->WQ, n_kf);
  CK(cudaStreamSynchronize(c->stream));   // the staging buffer may still feed an earlier search
  memcpy(c->hBow, kf_desc, (size_t)32 * n_kf);
  memcpy(c->hBow + (size_t)32 * n_kf, kf_angle, sizeof(float) * n_kf);
  memcpy(c->hBow + (size_t)36 * n_kf, kf_node, sizeof(int) * n_kf);
  memcpy(c->hBow + (size_t)40 * n_kf, kf_has_mp, (size_t)n_kf);
  CK(cudaMemcpyAsync(c->WQ.kfBlob, c->hBow, blob, cudaMemcpyHostToDevice, c->stream));
  c->nLaunchBow = ft_launch_bow_search(bow_source(c), c->W, c->WQ, n_kf, c->bowCap, nnratio, check_orientation ? 1 : 0, c->stream);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c->hBow, c->WQ.result, sizeof(int) * ((size_t)N + 4), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int* out = reinterpret_cast<const int*>(c->hBow);
  const int nm = out[0];
  memcpy(match, out + 4, sizeof(int) * N);
  if (nmatches) *nmatches = nm;
  return FT_OK;
}

extern "C" ft_status ft_set_stage_timing(ft_context* c, int enable) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  if (enable && !c->evA[0]) {
    for (int i = 0; i < FT_STAGE_COUNT; i++) { CK(cudaEventCreate(&c->evA[i])); CK(cudaEventCreate(&c->evB[i])); }
  }
  c->timing = enable;
  for (int i = 0; i < FT_STAGE_COUNT; i++) c->stageUsed[i] = false;
  return FT_OK;
}

extern "C" ft_status ft_get_stage_times(ft_context* c, float* ms, int n) {
  if (!c || !ms || n < FT_STAGE_COUNT) { set_err("ft_get_stage_times: bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->stream2));
  for (int i = 0; i < FT_STAGE_COUNT; i++) {
    ms[i] = -1.f;
    if (c->timing && c->stageUsed[i]) CK(cudaEventElapsedTime(&ms[i], c->evA[i], c->evB[i]));
  }
  return FT_OK;
}

extern "C" ft_status ft_host_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) { set_err("ft_host_alloc: bad argument"); return FT_ERR_INVALID; }
  CK(cudaMallocHost(out, bytes));
  return FT_OK;
}
extern "C" ft_status ft_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return FT_OK;
}
extern "C" ft_status ft_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) { set_err("ft_host_register: bad argument"); return FT_ERR_INVALID; }
  CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return FT_OK;
}
extern "C" ft_status ft_host_unregister(void* p) {
  if (p) CK(cudaHostUnregister(p));
  return FT_OK;
}

extern "C" ft_status ft_synchronize(ft_context* c) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  return FT_OK;
}

extern "C" void* ft_context_stream(ft_context* c) { return c ? (void*)c->stream : nullptr; }

extern "C" ft_status ft_set_use_graph(ft_context* c, int enable) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  c->useGraph = enable;
  return FT_OK;
}

extern "C" ft_status ft_launch_counts(ft_context* c, int* extract, int* stereo, int* search) {
  if (!c) { set_err("null context"); return FT_ERR_INVALID; }
  if (extract) *extract = c->nLaunchExtract;
  if (stereo) *stereo = c->nLaunchStereo;
  if (search) *search = c->nLaunchSearch;
  return FT_OK;
}

// ---- diagnostics --------------------------------------------------------------------------
extern "C" ft_status ft_debug_level_dims(ft_context* c, int level, int* w, int* h) {
  if (!c || level < 0 || level >= c->P.nlevels) { set_err("bad level"); return FT_ERR_INVALID; }
  *w = c->P.lv[level].w; *h = c->P.lv[level].h;
  return FT_OK;
}

extern "C" ft_status ft_debug_level_image(ft_context* c, int eye, int level, int blurred, uint8_t* out) {
  if (!c || level < 0 || level >= c->P.nlevels || eye < 0 || eye > 1 || !out) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  const FtLevel& L = c->P.lv[level];
  const uint8_t* src = (blurred ? c->B.eye[eye].blur : c->B.eye[eye].pyr) + L.offset;
  CK(cudaMemcpy2DAsync(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FT_OK;
}

extern "C" ft_status ft_debug_level_candidates(ft_context* c, int eye, int level, int cap, float* xyr, int* n) {
  if (!c || level < 0 || level >= c->P.nlevels || eye < 0 || eye > 1 || !n) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  const FtLevel& L = c->P.lv[level];
  int cnt = 0;
  CK(cudaMemcpyAsync(&cnt, c->B.eye[eye].lvlCandCount + level, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *n = cnt;
  if (!xyr || cap < cnt || cnt == 0) return FT_OK;
  // canonical order = cell row-major, then the order inside the cell's slab (ORBextractor.cc:1131-1203)
  const int nCells = L.nCols * L.nRows;
  std::vector<int> cc(nCells);
  std::vector<uint32_t> slab((size_t)nCells * L.cellCap);
  CK(cudaMemcpy(cc.data(), c->B.eye[eye].cellCount + L.cellBase, sizeof(int) * nCells, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(slab.data(), c->B.eye[eye].cellKp + L.cellKpBase, sizeof(uint32_t) * slab.size(), cudaMemcpyDeviceToHost));
  int i = 0;
  for (int cell = 0; cell < nCells; cell++)
    for (int k = 0; k < cc[cell] && i < cnt; k++, i++) {
      const uint32_t w = slab[(size_t)cell * L.cellCap + k];
      xyr[3 * i] = (float)ft_px(w); xyr[3 * i + 1] = (float)ft_py(w); xyr[3 * i + 2] = (float)ft_ps(w);
    }
  return FT_OK;
}

extern "C" ft_status ft_debug_track(ft_context* c, int M, int* track_i, float* track_f) {
  if (!c || M > c->lastM) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  if (track_i) CK(cudaMemcpy(track_i, c->Q.trI, sizeof(int) * 4 * M, cudaMemcpyDeviceToHost));
  if (track_f) CK(cudaMemcpy(track_f, c->Q.trF, sizeof(float) * 9 * M, cudaMemcpyDeviceToHost));
  return FT_OK;
}

extern "C" ft_status ft_debug_grid(ft_context* c, int right, int* counts, int* indices, int* n) {
  if (!c || !counts || !indices || !n) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  const int cells = FT_GRID_COLS * FT_GRID_ROWS;
  std::vector<int> start(cells + 1);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(start.data(), c->G.cellStart + (right ? FT_GRID_STRIDE : 0), sizeof(int) * (cells + 1), cudaMemcpyDeviceToHost));
  for (int i = 0; i < cells; i++) counts[i] = start[i + 1] - start[i];
  *n = start[cells];
  if (*n) CK(cudaMemcpy(indices, c->G.cellIdx + (right ? c->P.maxKp : 0), sizeof(int) * (*n), cudaMemcpyDeviceToHost));
  return FT_OK;
}

extern "C" ft_status ft_debug_stats(ft_context* c, long long* stats, int n) {
  if (!c || !stats || n < 8) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  int cand[2][FT_MAX_LEVELS], kp[2][FT_MAX_LEVELS];
  for (int e = 0; e < 2; e++) {
    CK(cudaMemcpy(cand[e], c->B.eye[e].lvlCandCount, sizeof(int) * FT_MAX_LEVELS, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(kp[e], c->B.eye[e].lvlKpCount, sizeof(int) * FT_MAX_LEVELS, cudaMemcpyDeviceToHost));
  }
  unsigned long long st[8];
  CK(cudaMemcpy(st, c->S.stats, sizeof(st), cudaMemcpyDeviceToHost));
  int cur[8];
  CK(cudaMemcpy(cur, c->Q.cursor, sizeof(cur), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; i++) stats[i] = 0;
  for (int l = 0; l < c->P.nlevels; l++) { stats[0] += cand[0][l]; stats[1] += cand[1][l]; stats[2] += kp[0][l]; stats[3] += kp[1][l]; }
  stats[4] = (long long)st[0]; stats[5] = (long long)st[1]; stats[6] = cur[7]; stats[7] = cur[2];
  return FT_OK;
}

extern "C" ft_status ft_debug_level_counts(ft_context* c, int eye, int* cand, int* kp) {
  if (!c || eye < 0 || eye > 1 || !cand || !kp) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(cand, c->B.eye[eye].lvlCandCount, sizeof(int) * c->P.nlevels, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(kp, c->B.eye[eye].lvlKpCount, sizeof(int) * c->P.nlevels, cudaMemcpyDeviceToHost));
  return FT_OK;
}

#ifdef FT_RS_CLOCK
extern "C" ft_status ft_debug_rs_clock(ft_context* c, long long* out /* [16][64] */) {
  if (!c || !out) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, c->B.eye[1].octClock, sizeof(long long) * 16 * 64, cudaMemcpyDeviceToHost));
  return FT_OK;
}
#endif

extern "C" ft_status ft_debug_oct_clock(ft_context* c, long long* out /* [2][FT_MAX_LEVELS][64] */) {
  if (!c || !out) { set_err("bad argument"); return FT_ERR_INVALID; }
  CK(cudaSetDevice(c->cfg.device_id));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out, c->B.eye[0].octClock, sizeof(long long) * 2 * FT_MAX_LEVELS * 64, cudaMemcpyDeviceToHost));
  return FT_OK;
}
