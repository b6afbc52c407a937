// ft_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see ft_oracle.h header).
// Every function cites the reference file:line it restates (paths relative to
// /root/reference). Built with -ffp-contract=off so float expressions are evaluated
// unfused, left to right, exactly as written.
#include "ft_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <list>
#include <utility>

namespace fto {

static const int8_t kPattern[1024] = {
#include "../include/ft_orb_pattern.inc"
};

static const int HALF_PATCH_SIZE = 15;   // include/ORBextractor.h:30
static const int PATCH_SIZE = 31;        // :29
static const int EDGE_THRESHOLD = 19;    // :31

// cvRound(float): round-half-to-even under the default rounding mode.
int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

// ---------------------------------------------------------------------------
// P1: cv::resize(8UC1, INTER_LINEAR) -- OpenCV imgproc resize.cpp, classic fixed-point
// path (INTER_RESIZE_COEF_BITS = 11). Call site: ORBextractor.cc:1508.
// ---------------------------------------------------------------------------
void resize_linear_u8(const Img& src, Img& dst, int dw, int dh) {
  const int sw = src.w, sh = src.h;
  dst.w = dw; dst.h = dh; dst.d.assign((size_t)dw * dh, 0);
  const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
  const double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
  std::vector<int> xofs(dw), xofs1(dw);
  std::vector<short> a0(dw), a1(dw);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    xofs1[dx] = std::min(sx + 1, sw - 1);
    a0[dx] = (short)cv_round((1.f - fx) * 2048.f);
    a1[dx] = (short)cv_round(fx * 2048.f);
  }
  std::vector<int> rowbuf0(dw), rowbuf1(dw);
  int cached0 = -1, cached1 = -1;
  auto hresize = [&](int sy, std::vector<int>& buf) {
    const uint8_t* S = src.row(sy);
    for (int dx = 0; dx < dw; dx++) buf[dx] = S[xofs[dx]] * a0[dx] + S[xofs1[dx]] * a1[dx];
  };
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    const short b0 = (short)cv_round((1.f - fy) * 2048.f);
    const short b1 = (short)cv_round(fy * 2048.f);
    const int y0 = std::min(std::max(sy, 0), sh - 1);
    const int y1 = std::min(std::max(sy + 1, 0), sh - 1);
    if (y0 == cached1) { std::swap(rowbuf0, rowbuf1); std::swap(cached0, cached1); }
    if (y0 != cached0) { hresize(y0, rowbuf0); cached0 = y0; }
    if (y1 == y0) { rowbuf1 = rowbuf0; cached1 = y1; }
    else if (y1 != cached1) { hresize(y1, rowbuf1); cached1 = y1; }
    uint8_t* D = dst.row(dy);
    for (int dx = 0; dx < dw; dx++)
      D[dx] = (uint8_t)((((b0 * (rowbuf0[dx] >> 4)) >> 16) + ((b1 * (rowbuf1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

// ---------------------------------------------------------------------------
// P2: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on 8U -- OpenCV's fixed-point
// path: 8.8 kernel {18,34,48,56,48,34,18}/256, horizontal in 8.8, vertical in 16.16,
// round-half-up. Call site: ORBextractor.cc:1456-1457.
// ---------------------------------------------------------------------------
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

void gaussian_blur_7x7_s2(const Img& src, Img& dst) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  const int w = src.w, h = src.h;
  dst.w = w; dst.h = h; dst.d.assign((size_t)w * h, 0);
  std::vector<uint16_t> H((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src.row(y);
    uint16_t* Hr = H.data() + (size_t)y * w;
    for (int x = 0; x < w; x++) {
      unsigned acc = 0;
      if (x >= 3 && x + 3 < w) {
        for (int k = 0; k < 7; k++) acc += K[k] * S[x + k - 3];
      } else {
        for (int k = 0; k < 7; k++) acc += K[k] * S[reflect101(x + k - 3, w)];
      }
      Hr[x] = (uint16_t)acc;
    }
  }
  for (int y = 0; y < h; y++) {
    const uint16_t* r[7];
    for (int k = 0; k < 7; k++) r[k] = H.data() + (size_t)reflect101(y + k - 3, h) * w;
    uint8_t* D = dst.row(y);
    for (int x = 0; x < w; x++) {
      uint32_t v = 0;
      for (int k = 0; k < 7; k++) v += (uint32_t)K[k] * r[k][x];
      D[x] = (uint8_t)((v + 32768u) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------
// P3: cv::FAST(TYPE_9_16, nonmaxSuppression=true) -- OpenCV features2d fast.cpp /
// fast_score.cpp. Call sites: ORBextractor.cc:1157,1176.
// Ring order (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)
// (-3,-1)(-3,0)(-3,1)(-2,2)(-1,3).
// ---------------------------------------------------------------------------
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Largest t such that p is a FAST-9 corner for threshold t (i.e. some 9-arc has all
// |diff| > t with one sign); -1 when no arc has a consistent sign. corner(th) <=> score >= th.
int fast_score_9_16(const uint8_t* p, int stride) {
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * stride + kRingDx[k]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = 0;  // max over arcs of min over arc of (+d) and of (-d)
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) {
      mn = std::min(mn, d[k + j]);
      mx = std::max(mx, d[k + j]);
    }
    best = std::max(best, mn);
    best = std::max(best, -mx);
  }
  return best - 1;
}

void fast_detect(const uint8_t* roi, int stride, int w, int h, int th, std::vector<Candidate>& out) {
  if (w < 7 || h < 7) return;
  std::vector<int> sc((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++) {
    const uint8_t* r = roi + (size_t)y * stride;
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* p = r + x;
      const int v = p[0];
      // quick reject: any 9-arc contains one pixel of each opposite pair
      const int lo = v - th, hi = v + th;
      const int t0 = p[3 * stride], t8 = p[-3 * stride], t4 = p[3], t12 = p[-3];
      const bool darkOK = (t0 < lo || t8 < lo) && (t4 < lo || t12 < lo);
      const bool brightOK = (t0 > hi || t8 > hi) && (t4 > hi || t12 > hi);
      if (!darkOK && !brightOK) continue;
      const int s = fast_score_9_16(p, stride);
      if (s >= th) sc[(size_t)y * w + x] = s;
    }
  }
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const int s = sc[(size_t)y * w + x];
      if (!s) continue;
      const int* c = &sc[(size_t)y * w + x];
      if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] && s > c[w] &&
          s > c[w + 1])
        out.push_back({(float)x, (float)y, (float)s});
    }
}

// ---------------------------------------------------------------------------
// P4: cv::fastAtan2 (scalar tail of hal::fastAtan32f, degrees). Call site ORBextractor.cc:65.
// ---------------------------------------------------------------------------
float fast_atan2(float y, float x) {
  static const float scale = (float)(180.0 / M_PI);
  static const float p1 = 0.9997878412794807f * scale;
  static const float p3 = -0.3258083974640975f * scale;
  static const float p5 = 0.1555786518463281f * scale;
  static const float p7 = -0.04432655554792128f * scale;
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------------------
// cv::remap, 8UC1, CV_32FC1 map pair, INTER_LINEAR, BORDER_CONSTANT(0) -- OpenCV imgproc imgwarp.cpp fixed-point
// path: coordinates rounded to 1/32 px (INTER_BITS = 5), 2x2 weights from BilinearTab_i scaled by 2^15 with the
// table's sum correction (only the entry fx = fy = 0 saturates: {32767, 0, 0, 1}), result (sum + 2^14) >> 15.
// Call site: System::TrackStereo (System.cc:279-280), "next" row 2 of SURVEY.md 8f.
// ---------------------------------------------------------------------------
void remap_linear_u8(const Img& src, const float* mapx, const float* mapy, int dw, int dh, Img& dst) {
  dst.w = dw; dst.h = dh; dst.d.assign((size_t)dw * dh, 0);
  const int sw = src.w, sh = src.h;
  auto px = [&](int y, int x) -> int { return (x >= 0 && x < sw && y >= 0 && y < sh) ? src.d[(size_t)y * sw + x] : 0; };
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) {
      int sx = cv_round(mapx[(size_t)y * dw + x] * 32.f), sy = cv_round(mapy[(size_t)y * dw + x] * 32.f);
      const int fx = sx & 31, fy = sy & 31;
      int ix = sx >> 5, iy = sy >> 5;
      ix = std::min(std::max(ix, -32768), 32767); iy = std::min(std::max(iy, -32768), 32767);   // saturate_cast<short>
      int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
      if (fx == 0 && fy == 0) { w00 = 32767; w11 = 1; }
      const int v = w00 * px(iy, ix) + w01 * px(iy, ix + 1) + w10 * px(iy + 1, ix) + w11 * px(iy + 1, ix + 1);
      dst.d[(size_t)y * dw + x] = (uint8_t)std::min(std::max((v + (1 << 14)) >> 15, 0), 255);
    }
}

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:2256-2273): popcount of 256-bit XOR.
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t x, y;
    memcpy(&x, a + 8 * i, 8);
    memcpy(&y, b + 8 * i, 8);
    dist += __builtin_popcountll(x ^ y);
  }
  return dist;
}

// P5: BFMatcher(NORM_HAMMING).knnMatch(k=2): ascending distance, ties -> lowest trainIdx.
void knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int* idx2, int* dist2) {
  for (int i = 0; i < nq; i++) {
    int b0 = INT_MAX, b1 = INT_MAX, i0 = -1, i1 = -1;
    for (int j = 0; j < nt; j++) {
      const int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
      if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
      else if (d < b1) { b1 = d; i1 = j; }
    }
    idx2[2 * i] = i0; idx2[2 * i + 1] = i1;
    dist2[2 * i] = i0 >= 0 ? b0 : -1; dist2[2 * i + 1] = i1 >= 0 ? b1 : -1;
  }
}

// ---------------------------------------------------------------------------
// ORBextractor::ORBextractor (ORBextractor.cc:393-499, non-CUDA parts)
// ---------------------------------------------------------------------------
Extractor::Extractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
  scale.resize(nlevels); sigma2.resize(nlevels);
  scale[0] = 1.0f; sigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    float s = scale[i - 1] * scaleFactor;   // float * double -> double -> float (:405)
    scale[i] = s;
    sigma2[i] = scale[i] * scale[i];
  }
  invScale.resize(nlevels); invSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
  featuresPerLevel.resize(nlevels);
  float factor = 1.0f / scaleFactor;   // (:455) float = 1.0f/double
  float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    featuresPerLevel[l] = cv_round(nDesired);
    sum += featuresPerLevel[l];
    nDesired *= factor;
  }
  featuresPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
  // umax (:478-493)
  umax.assign(HALF_PATCH_SIZE + 1, 0);
  int v, v0;
  const int vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
  const int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
  const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
  for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
  for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

// ORBextractor::ComputePyramid (ORBextractor.cc:1495-1520). Only the inner ROI is kept:
// nothing on this path reads the 19-px reflect border.
void Extractor::computePyramid(const uint8_t* img, int w, int h, int step) {
  pyramid.assign(nlevels, Img());
  for (int l = 0; l < nlevels; l++) {
    const float s = invScale[l];
    const int lw = cv_round((float)w * s), lh = cv_round((float)h * s);
    if (l == 0) {
      pyramid[0].w = w; pyramid[0].h = h; pyramid[0].d.resize((size_t)w * h);
      for (int y = 0; y < h; y++) memcpy(pyramid[0].row(y), img + (size_t)y * step, w);
    } else {
      resize_linear_u8(pyramid[l - 1], pyramid[l], lw, lh);
    }
  }
}

// ExtractorNode + DivideNode (ORBextractor.cc:510-566)
namespace {
struct Node {
  int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
  std::vector<Candidate> keys;
  bool noMore = false;
  std::list<Node>::iterator lit;
};
void divide(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
  const int halfX = (int)std::ceil(static_cast<float>(n.URx - n.ULx) / 2);
  const int halfY = (int)std::ceil(static_cast<float>(n.BRy - n.ULy) / 2);
  n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
  n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
  n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
  n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
  n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
  n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
  n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
  n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
  for (const Candidate& kp : n.keys) {
    if (kp.x < n1.URx) {
      if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
    } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
    else n4.keys.push_back(kp);
  }
  if (n1.keys.size() == 1) n1.noMore = true;
  if (n2.keys.size() == 1) n2.noMore = true;
  if (n3.keys.size() == 1) n3.noMore = true;
  if (n4.keys.size() == 1) n4.noMore = true;
}
typedef std::pair<int, Node*> SizeNode;
// compareNodes (ORBextractor.cc:626-641): (size, UL.x) ascending; other ties are left to
// std::sort, which is why this oracle uses libstdc++'s std::sort as the reference does.
bool compare_nodes(SizeNode& e1, SizeNode& e2) {
  if (e1.first < e2.first) return true;
  if (e1.first > e2.first) return false;
  return e1.second->ULx < e2.second->ULx;
}
}  // namespace

// ORBextractor::DistributeOctTree (ORBextractor.cc:660-884)
std::vector<Candidate> Extractor::distributeOctTree(const std::vector<Candidate>& in, int minX, int maxX,
                                                    int minY, int maxY, int N) const {
  const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
  std::vector<Candidate> result;
  if (nIni < 1) return result;  // reference indexes an empty vector here (UB); unsupported aspect ratio
  const float hX = static_cast<float>(maxX - minX) / nIni;
  std::list<Node> nodes;
  std::vector<Node*> ini(nIni);
  for (int i = 0; i < nIni; i++) {
    Node ni;
    ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
    ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
    ni.BLx = ni.ULx; ni.BLy = maxY - minY;
    ni.BRx = ni.URx; ni.BRy = maxY - minY;
    nodes.push_back(ni);
    ini[i] = &nodes.back();
  }
  for (const Candidate& kp : in) {
    size_t r = (size_t)(kp.x / hX);
    if (r >= (size_t)nIni) r = nIni - 1;  // unreachable for in-range keypoints; guards UB
    ini[r]->keys.push_back(kp);
  }
  auto lit = nodes.begin();
  while (lit != nodes.end()) {
    if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
    else if (lit->keys.empty()) lit = nodes.erase(lit);
    else lit++;
  }
  bool finish = false;
  std::vector<SizeNode> sizeAndNode;
  sizeAndNode.reserve(nodes.size() * 4);
  auto pushChild = [&](Node& c, int* nToExpand) {
    if (c.keys.size() > 0) {
      nodes.push_front(c);
      if (c.keys.size() > 1) {
        if (nToExpand) (*nToExpand)++;
        sizeAndNode.push_back(std::make_pair((int)c.keys.size(), &nodes.front()));
        nodes.front().lit = nodes.begin();
      }
    }
  };
  while (!finish) {
    int prevSize = (int)nodes.size();
    lit = nodes.begin();
    int nToExpand = 0;
    sizeAndNode.clear();
    while (lit != nodes.end()) {
      if (lit->noMore) { lit++; continue; }
      Node n1, n2, n3, n4;
      divide(*lit, n1, n2, n3, n4);
      pushChild(n1, &nToExpand); pushChild(n2, &nToExpand); pushChild(n3, &nToExpand); pushChild(n4, &nToExpand);
      lit = nodes.erase(lit);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
      finish = true;
    } else if (((int)nodes.size() + nToExpand * 3) > N) {
      while (!finish) {
        prevSize = (int)nodes.size();
        std::vector<SizeNode> prev = sizeAndNode;
        sizeAndNode.clear();
        std::sort(prev.begin(), prev.end(), compare_nodes);
        for (int j = (int)prev.size() - 1; j >= 0; j--) {
          Node n1, n2, n3, n4;
          divide(*prev[j].second, n1, n2, n3, n4);
          pushChild(n1, nullptr); pushChild(n2, nullptr); pushChild(n3, nullptr); pushChild(n4, nullptr);
          nodes.erase(prev[j].second->lit);
          if ((int)nodes.size() >= N) break;
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
      }
    }
  }
  result.reserve(nodes.size());
  for (auto& n : nodes) {
    const Candidate* best = &n.keys[0];
    float maxResponse = best->response;
    for (size_t k = 1; k < n.keys.size(); k++)
      if (n.keys[k].response > maxResponse) { best = &n.keys[k]; maxResponse = n.keys[k].response; }
    result.push_back(*best);
  }
  return result;
}

// IC_Angle (ORBextractor.cc:39-66)
static float ic_angle(const Img& im, float px, float py, const std::vector<int>& umax) {
  int m_01 = 0, m_10 = 0;
  const int step = im.w;
  const uint8_t* center = im.row(cv_round(py)) + cv_round(px);
  for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
  for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
    int v_sum = 0;
    const int d = umax[v];
    for (int u = -d; u <= d; ++u) {
      const int val_plus = center[u + v * step], val_minus = center[u - v * step];
      v_sum += (val_plus - val_minus);
      m_10 += u * (val_plus + val_minus);
    }
    m_01 += v * v_sum;
  }
  return fast_atan2((float)m_01, (float)m_10);
}

// ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:1112-1227)
void Extractor::computeKeyPointsOctTree() {
  candidates.assign(nlevels, {});
  levelKeys.assign(nlevels, {});
  const float W = 35;
  for (int level = 0; level < nlevels; ++level) {
    const Img& im = pyramid[level];
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
    std::vector<Candidate>& toDistribute = candidates[level];
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols < 1 || nRows < 1) continue;  // reference divides by zero here; image too small
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<Candidate> cell;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBorderY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBorderY - 3) continue;
      if (maxY > maxBorderY) maxY = (float)maxBorderY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBorderX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBorderX - 6) continue;
        if (maxX > maxBorderX) maxX = (float)maxBorderX;
        const int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
        cell.clear();
        fast_detect(im.row(y0) + x0, im.w, x1 - x0, y1 - y0, iniTh, cell);
        if (cell.empty()) fast_detect(im.row(y0) + x0, im.w, x1 - x0, y1 - y0, minTh, cell);
        for (Candidate& c : cell) {
          c.x += j * wCell;
          c.y += i * hCell;
          toDistribute.push_back(c);
        }
      }
    }
    std::vector<Candidate> kept =
        distributeOctTree(toDistribute, minBorderX, maxBorderX, minBorderY, maxBorderY, featuresPerLevel[level]);
    const int scaledPatchSize = (int)(PATCH_SIZE * scale[level]);
    std::vector<KeyPoint>& out = levelKeys[level];
    out.resize(kept.size());
    for (size_t i = 0; i < kept.size(); i++) {
      out[i].x = kept[i].x + minBorderX;
      out[i].y = kept[i].y + minBorderY;
      out[i].octave = level;
      out[i].size = (float)scaledPatchSize;
      out[i].response = kept[i].response;
      out[i].angle = -1;
    }
  }
  for (int level = 0; level < nlevels; ++level)
    for (KeyPoint& kp : levelKeys[level]) kp.angle = ic_angle(pyramid[level], kp.x, kp.y, umax);
}

// computeOrbDescriptor (ORBextractor.cc:68-108)
static void orb_descriptor(const KeyPoint& kpt, const Img& img, uint8_t* desc, long* borderline) {
  const float factorPI = (float)(M_PI / 180.f);
  const float angle = (float)kpt.angle * factorPI;
  const float a = cosf(angle), b = sinf(angle);
  const uint8_t* center = img.row(cv_round(kpt.y)) + cv_round(kpt.x);
  const int step = img.w;
  const int8_t* pat = kPattern;
  auto get = [&](int idx) -> int {
    const int px = pat[2 * idx], py = pat[2 * idx + 1];
    const float fy = px * b + py * a;
    const float fx = px * a - py * b;
    if (borderline) {
      const float ry = std::fabs(fy - std::floor(fy) - 0.5f), rx = std::fabs(fx - std::floor(fx) - 0.5f);
      if (ry < 1e-4f || rx < 1e-4f) (*borderline)++;
    }
    return center[cv_round(fy) * step + cv_round(fx)];
  };
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int k = 0; k < 8; k++) {
      const int t0 = get(2 * k), t1 = get(2 * k + 1);
      val |= (t0 < t1) << k;
    }
    desc[i] = (uint8_t)val;
  }
}

// ORBextractor::operator() CPU branch (ORBextractor.cc:1356-1493)
int Extractor::extract(const uint8_t* img, int w, int h, int step, int lap0, int lap1,
                       std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc) {
  if (!img || w <= 0 || h <= 0) return -1;
  descBorderline = 0;
  computePyramid(img, w, h, step);
  computeKeyPointsOctTree();
  int nkeypoints = 0;
  for (int l = 0; l < nlevels; ++l) nkeypoints += (int)levelKeys[l].size();
  kps.assign(nkeypoints, KeyPoint());
  desc.assign((size_t)nkeypoints * 32, 0);
  blurred.assign(nlevels, Img());
  levelDesc.assign(nlevels, {});
  int monoIndex = 0, stereoIndex = nkeypoints - 1;
  for (int level = 0; level < nlevels; ++level) {
    std::vector<KeyPoint>& keys = levelKeys[level];
    const int n = (int)keys.size();
    if (n == 0) continue;
    gaussian_blur_7x7_s2(pyramid[level], blurred[level]);
    std::vector<uint8_t>& ld = levelDesc[level];
    ld.assign((size_t)n * 32, 0);
    for (int i = 0; i < n; i++) orb_descriptor(keys[i], blurred[level], &ld[(size_t)i * 32], &descBorderline);
    const float sc = scale[level];
    for (int i = 0; i < n; i++) {
      KeyPoint kp = keys[i];
      if (level != 0) { kp.x *= sc; kp.y *= sc; }
      int dst;
      if (kp.x >= lap0 && kp.x <= lap1) dst = stereoIndex--;
      else dst = monoIndex++;
      kps[dst] = kp;
      memcpy(&desc[(size_t)dst * 32], &ld[(size_t)i * 32], 32);
    }
  }
  return monoIndex;
}

// ---------------------------------------------------------------------------
// Frame::ComputeStereoMatches (Frame.cc:835-1005)
// ---------------------------------------------------------------------------
void compute_stereo_matches(const Extractor& exL, const Extractor& exR, const std::vector<KeyPoint>& kL,
                            const std::vector<uint8_t>& dL, const std::vector<KeyPoint>& kR,
                            const std::vector<uint8_t>& dR, float mbf, float mb, StereoResult& out) {
  const int N = (int)kL.size(), Nr = (int)kR.size();
  out.uRight.assign(N, -1.0f); out.depth.assign(N, -1.0f);
  out.bestIdxR.assign(N, -1); out.sad.assign(N, -1);
  const int thOrbDist = (100 + 50) / 2;
  const int nRows = exL.pyramid[0].h;
  std::vector<std::vector<int>> rowIdx(nRows);
  for (int iR = 0; iR < Nr; iR++) {
    const float kpY = kR[iR].y;
    const float r = 2.0f * exL.scale[kR[iR].octave];
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);   // reference: unchecked (UB outside)
  }
  const float minZ = mb, minD = 0, maxD = mbf / minZ;
  std::vector<std::pair<int, int>> distIdx;
  for (int iL = 0; iL < N; iL++) {
    const KeyPoint& kpL = kL[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int rowi = (int)vL;
    if (rowi < 0 || rowi >= nRows) continue;
    const std::vector<int>& cand = rowIdx[rowi];
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = 100;
    int bestIdxR = 0;
    const uint8_t* dl = &dL[(size_t)iL * 32];
    for (int iR : cand) {
      const KeyPoint& kpR = kR[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = descriptor_distance(dl, &dR[(size_t)iR * 32]);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDist < thOrbDist) {
      out.bestIdxR[iL] = bestIdxR;
      const float uR0 = kR[bestIdxR].x;
      const float sf = exL.invScale[kpL.octave];
      const float scaleduL = std::round(kpL.x * sf), scaledvL = std::round(kpL.y * sf);
      const float scaleduR0 = std::round(uR0 * sf);
      const int w = 5, L = 5;
      const Img& IL = exL.pyramid[kpL.octave];
      const Img& IR = exR.pyramid[kpL.octave];
      int bestD = INT_MAX, bestinc = 0;
      float dists[2 * L + 1];
      const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
      if (iniu < 0 || endu >= IR.w) continue;
      const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
      bool oob = (cy - w < 0 || cy + w >= IL.h || cxl - w < 0 || cxl + w >= IL.w || cxr - L - w < 0);
      if (oob) continue;  // reference: cv::Mat range assert would abort; unreachable for extractor output
      for (int inc = -L; inc <= +L; inc++) {
        int s = 0;
        for (int yy = -w; yy <= w; yy++) {
          const uint8_t* a = IL.row(cy + yy) + cxl - w;
          const uint8_t* b = IR.row(cy + yy) + cxr + inc - w;
          for (int xx = 0; xx < 2 * w + 1; xx++) s += std::abs((int)a[xx] - (int)b[xx]);
        }
        const float dist = (float)s;
        if (dist < bestD) { bestD = (int)dist; bestinc = inc; }
        dists[L + inc] = dist;
      }
      if (bestinc == -L || bestinc == L) continue;
      const float d1 = dists[L + bestinc - 1], d2 = dists[L + bestinc], d3 = dists[L + bestinc + 1];
      const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
      if (deltaR < -1 || deltaR > 1) continue;
      float bestuR = exL.scale[kpL.octave] * ((float)scaleduR0 + (float)bestinc + deltaR);
      float disparity = (uL - bestuR);
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
          disparity = 0.01;
          bestuR = uL - 0.01;
        }
        out.depth[iL] = mbf / disparity;
        out.uRight[iL] = bestuR;
        out.sad[iL] = bestD;
        distIdx.push_back(std::make_pair(bestD, iL));
      }
    }
  }
  if (distIdx.empty()) return;  // reference indexes an empty vector (UB); guarded
  std::sort(distIdx.begin(), distIdx.end());
  const float median = (float)distIdx[distIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  for (int i = (int)distIdx.size() - 1; i >= 0; i--) {
    if (distIdx[i].first < thDist) break;
    out.uRight[distIdx[i].second] = -1;
    out.depth[distIdx[i].second] = -1;
  }
}

// ---------------------------------------------------------------------------
// Camera models
// ---------------------------------------------------------------------------
void cam_project(const Camera& c, const float P[3], float uv[2]) {
  if (c.type == 0) {  // Pinhole::project (Pinhole.cpp:43-49)
    uv[0] = c.p[0] * P[0] / P[2] + c.p[2];
    uv[1] = c.p[1] * P[1] / P[2] + c.p[3];
  } else {            // KannalaBrandt8::project (KannalaBrandt8.cpp:67-84)
    const float x2_plus_y2 = P[0] * P[0] + P[1] * P[1];
    const float theta = atan2f(sqrtf(x2_plus_y2), P[2]);
    const float psi = atan2f(P[1], P[0]);
    const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2;
    const float theta7 = theta5 * theta2, theta9 = theta7 * theta2;
    const float r = theta + c.p[4] * theta3 + c.p[5] * theta5 + c.p[6] * theta7 + c.p[7] * theta9;
    uv[0] = c.p[0] * r * cosf(psi) + c.p[2];
    uv[1] = c.p[1] * r * sinf(psi) + c.p[3];
  }
}

// KannalaBrandt8::unproject (KannalaBrandt8.cpp:116-143); precision = 1e-6 (KannalaBrandt8.h)
void kb8_unproject(const Camera& c, float u, float v, float ray[3]) {
  const float pwx = (u - c.p[2]) / c.p[0], pwy = (v - c.p[3]) / c.p[1];
  float scale = 1.f;
  float theta_d = sqrtf(pwx * pwx + pwy * pwy);
  theta_d = fminf(fmaxf((float)(-M_PI / 2.f), theta_d), (float)(M_PI / 2.f));
  if (theta_d > 1e-8) {
    float theta = theta_d;
    for (int j = 0; j < 10; j++) {
      const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
      const float k0_theta2 = c.p[4] * theta2, k1_theta4 = c.p[5] * theta4;
      const float k2_theta6 = c.p[6] * theta6, k3_theta8 = c.p[7] * theta8;
      const float theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                              (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
      theta = theta - theta_fix;
      if (fabsf(theta_fix) < 1e-6f) break;
    }
    scale = tanf(theta) / theta_d;
  }
  ray[0] = pwx * scale; ray[1] = pwy * scale; ray[2] = 1.f;
}

// Smallest right-singular vector of a 4x4 matrix (stands in for Eigen::JacobiSVD<Matrix4f>,
// KannalaBrandt8.cpp:403-405). Eigen is not vendored in the reference and absent here, so this
// is a TOLERANCE oracle: cyclic Jacobi on A^T A in double. Sign/scale cancel in x/w.
static void smallest_right_singular_vec4(const double A[4][4], double v[4]) {
  double M[4][4], V[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[k][i] * A[k][j];
      M[i][j] = s;
      V[i][j] = (i == j);
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int i = 0; i < 4; i++)
      for (int j = i + 1; j < 4; j++) off += M[i][j] * M[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 4; q++) {
        if (std::fabs(M[p][q]) < 1e-300) continue;
        const double theta = (M[q][q] - M[p][p]) / (2 * M[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 4; k++) {
          const double mkp = M[k][p], mkq = M[k][q];
          M[k][p] = c * mkp - s * mkq; M[k][q] = s * mkp + c * mkq;
        }
        for (int k = 0; k < 4; k++) {
          const double mpk = M[p][k], mqk = M[q][k];
          M[p][k] = c * mpk - s * mqk; M[q][k] = s * mpk + c * mqk;
        }
        for (int k = 0; k < 4; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; i++) if (M[i][i] < M[best][best]) best = i;
  for (int k = 0; k < 4; k++) v[k] = V[k][best];
}

static inline void mat3_mul_vec(const float R[9], const float v[3], float o[3]) {
  for (int i = 0; i < 3; i++) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}

// KannalaBrandt8::TriangulateMatches (KannalaBrandt8.cpp:306-377) + Triangulate (:394-406)
static float kb8_triangulate_matches(const Camera& c1, const Camera& c2, const KeyPoint& kp1, const KeyPoint& kp2,
                                     const float R12[9], const float t12[3], float sigmaLevel, float unc,
                                     float p3D[3]) {
  float r1[3], r2[3], r21[3];
  kb8_unproject(c1, kp1.x, kp1.y, r1);
  kb8_unproject(c2, kp2.x, kp2.y, r2);
  mat3_mul_vec(R12, r2, r21);
  const float n1 = sqrtf(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
  const float n21 = sqrtf(r21[0] * r21[0] + r21[1] * r21[1] + r21[2] * r21[2]);
  const float cosPar = (r1[0] * r21[0] + r1[1] * r21[1] + r1[2] * r21[2]) / (n1 * n21);
  if (cosPar > 0.9998) return -1;
  float R21[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R21[3 * i + j] = R12[3 * j + i];
  float Rt[3];
  mat3_mul_vec(R21, t12, Rt);
  float T2[3][4];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T2[i][j] = R21[3 * i + j]; T2[i][3] = -Rt[i]; }
  const float T1[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
  double A[4][4];
  for (int j = 0; j < 4; j++) {
    A[0][j] = (double)(r1[0] * T1[2][j] - T1[0][j]);
    A[1][j] = (double)(r1[1] * T1[2][j] - T1[1][j]);
    A[2][j] = (double)(r2[0] * T2[2][j] - T2[0][j]);
    A[3][j] = (double)(r2[1] * T2[2][j] - T2[1][j]);
  }
  double vh[4];
  smallest_right_singular_vec4(A, vh);
  float x3D[3] = {(float)(vh[0] / vh[3]), (float)(vh[1] / vh[3]), (float)(vh[2] / vh[3])};
  const float z1 = x3D[2];
  if (z1 <= 0) return -2;
  const float z2 = R21[6] * x3D[0] + R21[7] * x3D[1] + R21[8] * x3D[2] + T2[2][3];
  if (z2 <= 0) return -3;
  float uv1[2];
  cam_project(c1, x3D, uv1);
  const float ex1 = uv1[0] - kp1.x, ey1 = uv1[1] - kp1.y;
  if ((ex1 * ex1 + ey1 * ey1) > 5.991 * sigmaLevel) return -4;
  float x3D2[3];
  mat3_mul_vec(R21, x3D, x3D2);
  for (int i = 0; i < 3; i++) x3D2[i] += T2[i][3];
  float uv2[2];
  cam_project(c2, x3D2, uv2);
  const float ex2 = uv2[0] - kp2.x, ey2 = uv2[1] - kp2.y;
  if ((ex2 * ex2 + ey2 * ey2) > 5.991 * unc) return -5;
  p3D[0] = x3D[0]; p3D[1] = x3D[1]; p3D[2] = x3D[2];
  return z1;
}

// Frame::ComputeStereoFishEyeMatches (Frame.cc:1231-1271)
void compute_stereo_fisheye(const Camera& c1, const Camera& c2, const float Rlr[9], const float tlr[3],
                            const std::vector<float>& sigma2, const std::vector<KeyPoint>& kL,
                            const std::vector<uint8_t>& dL, int monoLeft, const std::vector<KeyPoint>& kR,
                            const std::vector<uint8_t>& dR, int monoRight, FisheyeResult& out) {
  const int Nleft = (int)kL.size(), Nright = (int)kR.size();
  out.l2r.assign(Nleft, -1); out.r2l.assign(Nright, -1);
  out.depth.assign(Nleft, -1.0f); out.p3d.assign((size_t)Nleft * 3, 0.f);
  out.code.assign(Nleft, 0);
  const int nq = Nleft - monoLeft, nt = Nright - monoRight;
  out.knnIdx.assign(std::max(nq, 0), -1);
  if (nq <= 0 || nt <= 0) return;
  std::vector<int> idx2((size_t)nq * 2), dist2((size_t)nq * 2);
  knn2_hamming(&dL[(size_t)monoLeft * 32], nq, &dR[(size_t)monoRight * 32], nt, idx2.data(), dist2.data());
  for (int q = 0; q < nq; q++) {
    out.knnIdx[q] = idx2[2 * q];
    if (idx2[2 * q + 1] < 0) continue;  // fewer than 2 neighbours
    const float d0 = (float)dist2[2 * q], d1 = (float)dist2[2 * q + 1];
    if (d0 < d1 * 0.7) {
      const int iL = q + monoLeft, iR = idx2[2 * q] + monoRight;
      const float s1 = sigma2[kL[iL].octave], s2 = sigma2[kR[iR].octave];
      float p3D[3] = {0, 0, 0};
      const float depth = kb8_triangulate_matches(c1, c2, kL[iL], kR[iR], Rlr, tlr, s1, s2, p3D);
      if (depth > 0.0001f) {
        out.l2r[iL] = iR; out.r2l[iR] = iL;
        out.p3d[3 * iL] = p3D[0]; out.p3d[3 * iL + 1] = p3D[1]; out.p3d[3 * iL + 2] = p3D[2];
        out.depth[iL] = depth;
        out.code[iL] = 1;
      } else {
        out.code[iL] = (int)depth == 0 ? -6 : (int)depth;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Frame grid (Frame.cc:409-440, 749-759) and GetFeaturesInArea (Frame.cc:681-747)
// ---------------------------------------------------------------------------
void FrameModel::assignFeaturesToGrid() {
  for (int i = 0; i < 64; i++) for (int j = 0; j < 48; j++) { grid[i][j].clear(); gridR[i][j].clear(); }
  for (int i = 0; i < N; i++) {
    const KeyPoint& kp = keys[i];
    const int px = (int)std::round((kp.x - minX) * gridWInv);
    const int py = (int)std::round((kp.y - minY) * gridHInv);
    if (px < 0 || px >= 64 || py < 0 || py >= 48) continue;
    if (Nleft == -1 || i < Nleft) grid[px][py].push_back(i);
    else gridR[px][py].push_back(i - Nleft);
  }
}

void FrameModel::featuresInArea(float x, float y, float r, int minLevel, int maxLevel, bool right,
                                std::vector<int>& out) const {
  out.clear();
  const float factorX = r, factorY = r;
  const int nMinCellX = std::max(0, (int)std::floor((x - minX - factorX) * gridWInv));
  if (nMinCellX >= 64) return;
  const int nMaxCellX = std::min(63, (int)std::ceil((x - minX + factorX) * gridWInv));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)std::floor((y - minY - factorY) * gridHInv));
  if (nMinCellY >= 48) return;
  const int nMaxCellY = std::min(47, (int)std::ceil((y - minY + factorY) * gridHInv));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const std::vector<int>& cell = right ? gridR[ix][iy] : grid[ix][iy];
      for (int idx : cell) {
        const KeyPoint& kp = (Nleft == -1 || !right) ? keys[idx] : keys[idx + Nleft];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float distx = kp.x - x, disty = kp.y - y;
        if (std::fabs(distx) < factorX && std::fabs(disty) < factorY) out.push_back(idx);
      }
    }
}

// ---------------------------------------------------------------------------
// Frame::isInFrustum (Frame.cc:536-598), isInFrustumChecks (:1308-1382),
// MapPoint::PredictScale (MapPoint.cc:531-546), Get{Min,Max}DistanceInvariance (:502-512)
// ---------------------------------------------------------------------------
static inline bool near_rel(double a, double b, double eps) { return std::fabs(a - b) <= eps * std::max(1.0, std::fabs(b)); }

static bool frustum_checks(const FrameModel& F, const MapPointIn& mp, float viewingCosLimit, bool right,
                           bool pinholeMode, float& u, float& v, float& xr, float& depth, float& viewCosOut,
                           int& level, int& borderline, bool& projWritten) {
  float mR[9], mt[3], twc[3];
  if (right) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        mR[3 * i + j] = F.Rrl[3 * i] * F.Rcw[j] + F.Rrl[3 * i + 1] * F.Rcw[3 + j] + F.Rrl[3 * i + 2] * F.Rcw[6 + j];
    float Rt[3];
    mat3_mul_vec(F.Rrl, F.tcw, Rt);
    for (int i = 0; i < 3; i++) mt[i] = Rt[i] + F.trl[i];
    float Rw[3];
    mat3_mul_vec(F.Rwc, F.tlr, Rw);
    for (int i = 0; i < 3; i++) twc[i] = Rw[i] + F.Ow[i];
  } else {
    memcpy(mR, F.Rcw, sizeof(mR)); memcpy(mt, F.tcw, sizeof(mt)); memcpy(twc, F.Ow, sizeof(twc));
  }
  const float* P = mp.pos;
  float Pc[3];
  mat3_mul_vec(mR, P, Pc);
  for (int i = 0; i < 3; i++) Pc[i] += mt[i];
  const float Pc_dist = sqrtf(Pc[0] * Pc[0] + Pc[1] * Pc[1] + Pc[2] * Pc[2]);
  const float PcZ = Pc[2];
  if (std::fabs((double)PcZ) < 1e-6) borderline = 1;
  if (PcZ < 0.0f) return false;
  float uv[2];
  cam_project(right ? F.cam2 : F.cam1, Pc, uv);
  if (near_rel(uv[0], F.minX, 1e-5) || near_rel(uv[0], F.maxX, 1e-5) || near_rel(uv[1], F.minY, 1e-5) ||
      near_rel(uv[1], F.maxY, 1e-5)) borderline = 1;
  if (uv[0] < F.minX || uv[0] > F.maxX) return false;
  if (uv[1] < F.minY || uv[1] > F.maxY) return false;
  if (pinholeMode) { u = uv[0]; v = uv[1]; projWritten = true; }
  const float maxDistance = 1.2f * mp.maxDist, minDistance = 0.8f * mp.minDist;
  const float PO[3] = {P[0] - twc[0], P[1] - twc[1], P[2] - twc[2]};
  const float dist = sqrtf(PO[0] * PO[0] + PO[1] * PO[1] + PO[2] * PO[2]);
  if (near_rel(dist, minDistance, 1e-5) || near_rel(dist, maxDistance, 1e-5)) borderline = 1;
  if (dist < minDistance || dist > maxDistance) return false;
  const float viewCos = (PO[0] * mp.normal[0] + PO[1] * mp.normal[1] + PO[2] * mp.normal[2]) / dist;
  if (near_rel(viewCos, viewingCosLimit, 1e-5) || near_rel(viewCos, 0.998, 1e-5)) borderline = 1;
  if (viewCos < viewingCosLimit) return false;
  const float ratio = mp.maxDist / dist;
  const float q = logf(ratio) / F.logScale;
  {
    const double qd = std::log((double)ratio) / (double)F.logScale;
    if (std::fabs(qd - std::nearbyint(qd)) < 1e-5) borderline = 1;
  }
  int nScale = (int)std::ceil(q);
  if (nScale < 0) nScale = 0;
  else if (nScale >= F.nlevels) nScale = F.nlevels - 1;
  u = uv[0]; v = uv[1];
  xr = pinholeMode ? uv[0] - F.mbf * (1.0f / PcZ) : 0.f;
  depth = Pc_dist; viewCosOut = viewCos; level = nScale;
  return true;
}

void is_in_frustum(const FrameModel& F, const MapPointIn& mp, float viewCosLimit, MapPointTrack& t) {
  t = MapPointTrack();
  if (mp.flags & 1) return;
  bool pw = false;
  if (F.Nleft == -1) {
    float u = -1, v = -1, xr = 0, depth = 0, vc = 0; int level = -1;
    const bool ok = frustum_checks(F, mp, viewCosLimit, false, true, u, v, xr, depth, vc, level, t.borderline, pw);
    t.projX = u; t.projY = v;
    if (ok) { t.inView = 1; t.projXR = xr; t.depth = depth; t.level = level; t.viewCos = vc; }
  } else {
    float u = 0, v = 0, xr = 0, depth = 0, vc = 0; int level = -1;
    if (frustum_checks(F, mp, viewCosLimit, false, false, u, v, xr, depth, vc, level, t.borderline, pw)) {
      t.inView = 1; t.projX = u; t.projY = v; t.depth = depth; t.level = level; t.viewCos = vc;
    }
    if (frustum_checks(F, mp, viewCosLimit, true, false, u, v, xr, depth, vc, level, t.borderline, pw)) {
      t.inViewR = 1; t.projXR_r = u; t.projYR_r = v; t.depthR = depth; t.levelR = level; t.viewCosR = vc;
    }
  }
}

// ---------------------------------------------------------------------------
// ORBmatcher::SearchByProjection #1 (ORBmatcher.cc:49-225), RadiusByViewingCos (:314-320)
// ---------------------------------------------------------------------------
int search_by_projection(FrameModel& F, const std::vector<MapPointIn>& mps, const std::vector<MapPointTrack>& tr,
                         float th, bool bFar, float thFar, float nnratio, std::vector<int>& holder,
                         std::vector<uint8_t>& holderObs) {
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  const int TH_HIGH = 100;
  std::vector<int> vIdx;
  auto blocked = [&](int slot) { return holder[slot] != -1 && holderObs[slot]; };
  auto put = [&](int slot, int iMP) { holder[slot] = iMP; holderObs[slot] = (mps[iMP].flags >> 1) & 1; };
  for (size_t iMP = 0; iMP < mps.size(); iMP++) {
    const MapPointTrack& t = tr[iMP];
    if (!t.inView && !t.inViewR) continue;
    if (bFar && t.depth > thFar) continue;
    if (mps[iMP].flags & 1) continue;
    const uint8_t* md = mps[iMP].desc;
    if (t.inView) {
      const int lvl = t.level;
      float r = (t.viewCos > 0.998) ? 2.5f : 4.0f;
      if (bFactor) r *= th;
      F.featuresInArea(t.projX, t.projY, r * F.scale[lvl], lvl - 1, lvl, false, vIdx);
      if (!vIdx.empty()) {
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIdx) {
          if (blocked(idx)) continue;
          if (F.Nleft == -1 && F.uRight[idx] > 0) {
            const float er = std::fabs(t.projXR - F.uRight[idx]);
            if (er > r * F.scale[lvl]) continue;
          }
          const int dist = descriptor_distance(md, &F.desc[(size_t)idx * 32]);
          if (dist < bestDist) {
            bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
            bestLevel = F.keys[idx].octave; bestIdx = idx;
          } else if (dist < bestDist2) {
            bestLevel2 = F.keys[idx].octave; bestDist2 = dist;
          }
        }
        if (bestDist <= TH_HIGH) {
          if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
          if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
            put(bestIdx, (int)iMP);
            if (F.Nleft != -1 && F.l2r[bestIdx] != -1) {
              put(F.l2r[bestIdx] + F.Nleft, (int)iMP);
              nmatches++;
            }
            nmatches++;
          }
        }
      }
    }
    if (F.Nleft != -1 && t.inViewR) {
      const int lvl = t.levelR;
      if (lvl != -1) {
        const float r = (t.viewCosR > 0.998) ? 2.5f : 4.0f;
        F.featuresInArea(t.projXR_r, t.projYR_r, r * F.scale[lvl], lvl - 1, lvl, true, vIdx);
        if (vIdx.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIdx) {
          if (blocked(idx + F.Nleft)) continue;
          const int dist = descriptor_distance(md, &F.desc[(size_t)(idx + F.Nleft) * 32]);
          if (dist < bestDist) {
            bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
            bestLevel = F.keys[idx + F.Nleft].octave; bestIdx = idx;
          } else if (dist < bestDist2) {
            bestLevel2 = F.keys[idx + F.Nleft].octave; bestDist2 = dist;
          }
        }
        if (bestDist <= TH_HIGH) {
          if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
          if (F.r2l[bestIdx] != -1) {
            put(F.r2l[bestIdx], (int)iMP);
            nmatches++;
          }
          put(bestIdx + F.Nleft, (int)iMP);
          nmatches++;
        }
      }
    }
  }
  return nmatches;
}


// ---------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (ORBmatcher.cc:1775-2085),
// ORBmatcher::ComputeThreeMaxima (:2210-2254). Used by Tracking::TrackWithMotionModel (Tracking.cc:2911-2990).
// ---------------------------------------------------------------------------
static void compute_three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

int search_by_projection_last_frame(FrameModel& F, const std::vector<LastFramePoint>& pts, float th, int direction,
                                    bool checkOrientation, float mb, std::vector<int>& holder,
                                    std::vector<uint8_t>& holderObs, std::vector<int>& borderline) {
  (void)mb;
  const int HISTO_LENGTH = 30, TH_HIGH = 100;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  const bool bForward = direction > 0, bBackward = direction < 0;
  std::vector<int> vIdx;
  borderline.assign(pts.size(), 0);
  auto blocked = [&](int slot) { return holder[slot] != -1 && holderObs[slot]; };
  for (size_t i = 0; i < pts.size(); i++) {
    const LastFramePoint& lp = pts[i];
    if (lp.flags & 1) continue;
    float x3Dc[3];
    mat3_mul_vec(F.Rcw, lp.pos, x3Dc);
    for (int k = 0; k < 3; k++) x3Dc[k] += F.tcw[k];
    const float invzc = (float)(1.0 / x3Dc[2]);
    if (std::fabs((double)x3Dc[2]) < 1e-6) borderline[i] = 1;
    if (invzc < 0) continue;
    float uv[2];
    cam_project(F.cam1, x3Dc, uv);
    if (near_rel(uv[0], F.minX, 1e-5) || near_rel(uv[0], F.maxX, 1e-5) || near_rel(uv[1], F.minY, 1e-5) ||
        near_rel(uv[1], F.maxY, 1e-5)) borderline[i] = 1;
    if (uv[0] < F.minX || uv[0] > F.maxX) continue;
    if (uv[1] < F.minY || uv[1] > F.maxY) continue;
    const int nLastOctave = lp.octave;
    const float radius = th * F.scale[nLastOctave];
    if (bForward) F.featuresInArea(uv[0], uv[1], radius, nLastOctave, -1, false, vIdx);
    else if (bBackward) F.featuresInArea(uv[0], uv[1], radius, 0, nLastOctave, false, vIdx);
    else F.featuresInArea(uv[0], uv[1], radius, nLastOctave - 1, nLastOctave + 1, false, vIdx);
    if (vIdx.empty()) continue;
    const uint8_t obs = (uint8_t)((lp.flags >> 1) & 1);
    {
      int bestDist = 256, bestIdx2 = -1;
      for (int i2 : vIdx) {
        if (blocked(i2)) continue;
        if (F.Nleft == -1 && F.uRight[i2] > 0) {
          const float ur = uv[0] - F.mbf * invzc;
          const float er = std::fabs(ur - F.uRight[i2]);
          if (er > radius) continue;
        }
        const int dist = descriptor_distance(lp.desc, &F.desc[(size_t)i2 * 32]);
        if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
      }
      if (bestDist <= TH_HIGH) {
        holder[bestIdx2] = (int)i; holderObs[bestIdx2] = obs;
        nmatches++;
        if (checkOrientation) {
          float rot = lp.angle - F.keys[bestIdx2].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(bestIdx2);
        }
      }
    }
    if (F.Nleft != -1) {
      float x3Dr[3];
      mat3_mul_vec(F.Rrl, x3Dc, x3Dr);
      for (int k = 0; k < 3; k++) x3Dr[k] += F.trl[k];
      float uvr[2];
      cam_project(F.cam1, x3Dr, uvr);   // the reference projects with mpCamera (the LEFT model) here (:1921)
      if (bForward) F.featuresInArea(uvr[0], uvr[1], radius, nLastOctave, -1, true, vIdx);
      else if (bBackward) F.featuresInArea(uvr[0], uvr[1], radius, 0, nLastOctave, true, vIdx);
      else F.featuresInArea(uvr[0], uvr[1], radius, nLastOctave - 1, nLastOctave + 1, true, vIdx);
      int bestDist = 256, bestIdx2 = -1;
      for (int i2 : vIdx) {
        if (blocked(i2 + F.Nleft)) continue;
        const int dist = descriptor_distance(lp.desc, &F.desc[(size_t)(i2 + F.Nleft) * 32]);
        if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
      }
      if (bestDist <= TH_HIGH) {
        holder[bestIdx2 + F.Nleft] = (int)i; holderObs[bestIdx2 + F.Nleft] = obs;
        nmatches++;
        if (checkOrientation) {
          float rot = lp.angle - F.keys[bestIdx2 + F.Nleft].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(bestIdx2 + F.Nleft);
        }
      }
    }
  }
  if (checkOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i != ind1 && i != ind2 && i != ind3) {
        for (int slot : rotHist[i]) { holder[slot] = -1; holderObs[slot] = 0; nmatches--; }
      }
    }
  }
  return nmatches;
}

// ---------------------------------------------------------------------------
// cv::undistortPoints(src, dst, K, distCoef, Mat(), P=K) as Frame::UndistortKeyPoints / ComputeImageBounds call it
// (Frame.cc:771-835; "next" row 2 of SURVEY.md 8f). OpenCV (calib3d undistort, pinned against cv2 4.13): all in
// double; normalise with 1/fx, 1/fy; FIVE fixed-point iterations (default TermCriteria(MAX_ITER, 5, 0.01): only the
// count is active) of x = (x0 - deltaX) * icdist with the rational radial factor and the tangential terms; bail out
// to the plain normalised point when icdist < 0; re-project with P (xx*ww, ww = 1/1); store as float.
// dist = k1 k2 p1 p2 [k3] (ORB-SLAM3's mDistCoef has 4 or 5 entries, Settings / Tracking::ParseCamParamFile).
// ---------------------------------------------------------------------------
void undistort_points(const float* xy, int n, const float K[4], const float* dist, int ndist, float* out) {
  double k[14] = {0};
  for (int i = 0; i < ndist && i < 14; i++) k[i] = (double)dist[i];
  const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  const double ifx = 1. / fx, ify = 1. / fy;
  for (int i = 0; i < n; i++) {
    double x = xy[2 * i], y = xy[2 * i + 1];
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) {
        x = (u - cx) * ifx;
        y = (v - cy) * ify;
        break;
      }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0. * y + cx;
    const double yy = 0. * x + fy * y + cy;
    const double ww = 1. / (0. * x + 0. * y + 1.);
    out[2 * i] = (float)(xx * ww);
    out[2 * i + 1] = (float)(yy * ww);
  }
}

// Frame::ComputeImageBounds (Frame.cc:806-833): {mnMinX, mnMaxX, mnMinY, mnMaxY}
void image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float out[4]) {
  if (ndist == 0 || dist[0] == 0.0f) { out[0] = 0.f; out[1] = (float)cols; out[2] = 0.f; out[3] = (float)rows; return; }
  const float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
  float u[8];
  undistort_points(c, 4, K, dist, ndist, u);
  out[0] = std::min(u[0], u[4]); out[1] = std::max(u[2], u[6]);
  out[2] = std::min(u[1], u[3]); out[3] = std::max(u[5], u[7]);
}

// Frame::ComputeStereoFromRGBD (Frame.cc:1065-1086): depth image lookup at the RAW keypoint (float coordinates
// truncated to int by cv::Mat::at<float>(v, u)), mvuRight from the UNDISTORTED x. depth == nullptr: monocular frame
// (Frame.cc:330-331), everything stays -1.
void stereo_from_rgbd(const float* keysXY, const float* keysUnX, int n, const float* depth, int w, int h, float mbf,
                      float* uRight, float* outDepth) {
  for (int i = 0; i < n; i++) {
    uRight[i] = -1.f; outDepth[i] = -1.f;
    if (!depth) continue;
    const float v = keysXY[2 * i + 1], u = keysXY[2 * i];
    const int iu = (int)u, iv = (int)v;
    if (iu < 0 || iu >= w || iv < 0 || iv >= h) continue;
    const float d = depth[(size_t)iv * w + iu];
    if (d > 0) { outDepth[i] = d; uRight[i] = keysUnX[i] - mbf / d; }
  }
}

// Test entry points: orientation and descriptor of given keypoints on given images, so that IC_Angle and
// computeOrbDescriptor can be pinned against OpenCV's own ORB (cv::ORB::detect / compute use the same two routines;
// ORB-SLAM's copies derive from them).
void ic_angles(const Extractor& ex, const uint8_t* img, int w, int h, const float* xy, int n, float* out) {
  Img im; im.w = w; im.h = h; im.d.assign(img, img + (size_t)w * h);
  for (int i = 0; i < n; i++) out[i] = ic_angle(im, xy[2 * i], xy[2 * i + 1], ex.umax);
}
void orb_descriptors(const uint8_t* blurred, int w, int h, const float* xy, const float* angle, int n, uint8_t* out) {
  Img im; im.w = w; im.h = h; im.d.assign(blurred, blurred + (size_t)w * h);
  for (int i = 0; i < n; i++) {
    KeyPoint kp{xy[2 * i], xy[2 * i + 1], 31.f, angle[i], 0.f, 0};
    orb_descriptor(kp, im, out + (size_t)32 * i, nullptr);
  }
}

}  // namespace fto
