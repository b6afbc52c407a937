// ft_shim.h -- host-side mirror of the reference's operator surface over the C ABI (header-only, C++17).
//
// Same names, argument meaning and return values as the reference classes on this path, so host code written
// against ORB-SLAM3 / FastTrack keeps compiling:
//   ORB_SLAM3::ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)
//                                              (reference include/ORBextractor.h:113-115, src/ORBextractor.cc:1356-1493)
//   ORB_SLAM3::Frame::ComputeStereoMatches / ComputeStereoFishEyeMatches / AssignFeaturesToGrid
//                                              (src/Frame.cc:835-1005, 1231-1271, 409-440)
//   ORB_SLAM3::ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
//                                              (include/ORBmatcher.h:47, src/ORBmatcher.cc:49-312), including the
//                                              isInFrustum pass of Tracking::SearchLocalPoints (src/Tracking.cc:3504-3522)
//   ORB_SLAM3::ORBVocabulary::loadFromTextFile / transform, Frame::ComputeBoW, KeyFrame::ComputeBoW,
//   ORB_SLAM3::ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)
//                                              (include/ORBVocabulary.h, src/Frame.cc:762-769, src/KeyFrame.cc:98-108,
//                                              src/ORBmatcher.cc:322-523)
// Everything is computed on the GPU by libfasttrack_b200; there is no CPU branch here. When the library is
// compiled into a tree that has OpenCV, define FT_SHIM_USE_OPENCV and the cv:: types are used directly;
// otherwise the minimal ftcv:: stand-ins below carry the same fields.
#pragma once
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fasttrack_b200.h"

#ifdef FT_SHIM_USE_OPENCV
#include <opencv2/core/core.hpp>
namespace ftcv = cv;
#else
namespace ftcv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint {           // cv::KeyPoint
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};
class Mat {                 // 8-bit single-channel subset of cv::Mat
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, unsigned char* d, size_t s) : rows(r), cols(c), step(s), data(d) {}
  void create(int r, int c) { own_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c); rows = r; cols = c; step = c; data = own_->data(); }
  void release() { own_.reset(); rows = cols = 0; step = 0; data = nullptr; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  unsigned char* ptr(int r) { return data + (size_t)r * step; }
  const unsigned char* ptr(int r) const { return data + (size_t)r * step; }
 private:
  std::shared_ptr<std::vector<unsigned char>> own_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace ftcv
#endif

// The two DBoW2 containers the tracker keeps per frame (reference Thirdparty/DBoW2/DBoW2/BowVector.h:58-60,
// FeatureVector.h:23-25): plain std::maps, filled from the device results.
namespace DBoW2 {
typedef unsigned int WordId;
typedef unsigned int NodeId;
typedef double WordValue;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2

namespace ORB_SLAM3 {

class FtError : public std::runtime_error {
 public:
  FtError(int st, const std::string& m) : std::runtime_error("fasttrack_b200 status " + std::to_string(st) + ": " + m), status(st) {}
  int status;
};
inline void ft_check(ft_status st) { if (st != FT_OK) throw FtError((int)st, ft_last_error()); }

// ORBVocabulary (reference include/ORBVocabulary.h:28-29 = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>), resident
// on one device. Descriptors are passed as rows of 32 bytes.
class ORBVocabulary {
 public:
  explicit ORBVocabulary(int device_id = 0) : device_(device_id) {}
  ~ORBVocabulary() { if (voc_) ft_vocabulary_destroy(voc_); }
  ORBVocabulary(const ORBVocabulary&) = delete;
  bool loadFromTextFile(const std::string& filename) {          // TemplatedVocabulary.h:1338-1423
    if (voc_) { ft_vocabulary_destroy(voc_); voc_ = nullptr; }
    return ft_vocabulary_load_text(device_, filename.c_str(), &voc_) == FT_OK;
  }
  bool empty() const { return voc_ == nullptr; }
  ft_vocabulary* get() const { return voc_; }
  // transform(features, v, fv, levelsup) (:1127-1194) for n host descriptors of 32 bytes
  void transform(const unsigned char* desc, int n, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
    std::vector<int> node(n > 0 ? n : 1);
    std::vector<uint32_t> ids(n > 0 ? n : 1);
    std::vector<double> vals(n > 0 ? n : 1);
    int nb = 0;
    ft_check(ft_vocabulary_transform(voc_, desc, n, levelsup, nullptr, node.data(), ids.data(), vals.data(), (int)ids.size(), &nb));
    fill(node.data(), n, ids.data(), vals.data(), nb, v, fv);
  }
  static void fill(const int* node, int n, const uint32_t* ids, const double* vals, int nb, DBoW2::BowVector& v,
                   DBoW2::FeatureVector& fv) {
    v.clear(); fv.clear();
    for (int i = 0; i < nb; i++) v.insert(v.end(), std::make_pair(ids[i], vals[i]));
    for (int i = 0; i < n; i++) if (node[i] >= 0) fv[(unsigned)node[i]].push_back((unsigned)i);
  }

 private:
  int device_;
  ft_vocabulary* voc_ = nullptr;
};

// One GPU context shared by the two extractors of a stereo rig and by the Frame / ORBmatcher mirrors.
// Replaces KernelController's static singletons (reference include/Kernels/KernelController.h).
class FrontEndContext {
 public:
  explicit FrontEndContext(const ft_config& cfg) : cfg_(cfg) { ft_check(ft_context_create(&cfg_, &ctx_)); }
  ~FrontEndContext() { ft_context_destroy(ctx_); }
  FrontEndContext(const FrontEndContext&) = delete;
  ft_context* get() const { return ctx_; }
  const ft_config& config() const { return cfg_; }
  // Calls on one ft_context are single-threaded (include/fasttrack_b200.h): the two extractor threads of a stereo rig
  // take this lock around every C-ABI call they make after the rendezvous in submit().
  std::mutex& api_mutex() { return api_mu_; }

  // Frame::ExtractORB runs the two extractors on two std::threads (src/Frame.cc:127-130). The first eye to
  // arrive parks its image; the second one launches the stereo extraction for both and wakes the first.
  void submit(int eye, const unsigned char* img, int step) {
    std::unique_lock<std::mutex> lk(mu_);
    img_[eye] = img; step_[eye] = step;
    const unsigned gen = gen_;
    if (++arrived_ == 2) {
      std::lock_guard<std::mutex> api(api_mu_);
      ft_status st = ft_extract_stereo(ctx_, img_[0], step_[0], img_[1], step_[1]);
      if (st == FT_OK) st = ft_synchronize(ctx_);
      status_ = st; err_ = st == FT_OK ? "" : ft_last_error();
      arrived_ = 0; ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen_ != gen; });
    }
    if (status_ != FT_OK) throw FtError((int)status_, err_);
  }

 private:
  ft_config cfg_;
  ft_context* ctx_ = nullptr;
  std::mutex mu_, api_mu_;
  std::condition_variable cv_;
  const unsigned char* img_[2] = {nullptr, nullptr};
  int step_[2] = {0, 0};
  int arrived_ = 0;
  unsigned gen_ = 0;
  ft_status status_ = FT_OK;
  std::string err_;
};

class ORBextractor {
 public:
  // same parameter list as the reference constructor plus the shared context and which eye this instance serves
  ORBextractor(std::shared_ptr<FrontEndContext> fe, int eye) : fe_(fe), eye_(eye) {
    const ft_config& c = fe_->config();
    nfeatures = c.nfeatures; nlevels = c.nlevels; scaleFactor = c.scale_factor;
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    ft_check(ft_get_scale_tables(fe_->get(), mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                 mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()));
  }

  // Compute the ORB features and descriptors on an image; returns monoIndex, -1 on an empty image.
  int operator()(ftcv::InputArray image, ftcv::InputArray /*mask: ignored, as in the reference*/,
                 std::vector<ftcv::KeyPoint>& keypoints, ftcv::OutputArray descriptors, std::vector<int>& vLappingArea) {
    if (image.empty()) return -1;
    const ft_config& c = fe_->config();
    if (image.cols != c.width || image.rows != c.height) throw FtError(FT_ERR_INVALID, "image size differs from the configured camera");
    const int* lap = eye_ == 0 ? c.lap_left : c.lap_right;
    if (vLappingArea.size() != 2 || vLappingArea[0] != lap[0] || vLappingArea[1] != lap[1])
      throw FtError(FT_ERR_INVALID, "vLappingArea differs from the rig configuration the context was created with");
    fe_->submit(eye_, image.data, (int)image.step);
    const int cap = c.nfeatures + 64 * c.nlevels;
    std::vector<ft_keypoint> k(cap);
    std::vector<unsigned char> d((size_t)cap * 32);
    int n = 0, mono = 0;
    {
      std::lock_guard<std::mutex> api(fe_->api_mutex());     // the other eye's thread downloads from the same context
      ft_check(ft_frame_download(fe_->get(), eye_, cap, k.data(), d.data(), &n, &mono, nullptr, nullptr, nullptr, nullptr, nullptr));
    }
    keypoints.assign(n, ftcv::KeyPoint());
    for (int i = 0; i < n; i++) {
      ftcv::KeyPoint& o = keypoints[i];
      o.pt.x = k[i].x; o.pt.y = k[i].y; o.size = k[i].size; o.angle = k[i].angle; o.response = k[i].response;
      o.octave = k[i].octave; o.class_id = -1;
    }
    if (n == 0) descriptors.release();
    else {
      descriptors.create(n, 32);
      for (int i = 0; i < n; i++) memcpy(descriptors.ptr(i), &d[(size_t)i * 32], 32);
    }
    return mono;
  }

  int GetLevels() { return nlevels; }
  float GetScaleFactor() { return (float)scaleFactor; }
  int GetNFeatures() { return nfeatures; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

 protected:
  std::shared_ptr<FrontEndContext> fe_;
  int eye_;
  int nfeatures, nlevels;
  double scaleFactor;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

// The fields of MapPoint that the projection search reads or writes (reference include/MapPoint.h:170-181 and
// the getters used in Frame::isInFrustum). A full MapPoint class can expose the same names.
struct MapPoint {
  float mWorldPos[3] = {0, 0, 0}, mNormalVector[3] = {0, 0, 1};
  float mfMinDistance = 0, mfMaxDistance = 0;
  unsigned char mDescriptor[32] = {0};
  int nObs = 1;
  bool mbBad = false;
  unsigned long mnLastFrameSeen = ~0ul;
  // tracking scratch written by the search
  bool mbTrackInView = false, mbTrackInViewR = false;
  float mTrackProjX = -1, mTrackProjY = -1, mTrackProjXR = 0, mTrackDepth = 0, mTrackViewCos = 0;
  float mTrackProjXR_r = 0, mTrackProjYR = 0, mTrackDepthR = 0, mTrackViewCosR = 0;
  int mnTrackScaleLevel = -1, mnTrackScaleLevelR = -1;
  // row in the persistent device-side store (MapStore below); -1 = not inserted. The setters of a full MapPoint
  // (SetWorldPos, UpdateNormalAndDepth, ComputeDistinctiveDescriptors) set mbStoreDirty.
  int mnStoreRow = -1;
  bool mbStoreDirty = true;
  bool isBad() const { return mbBad; }
  int Observations() const { return nObs; }
};

class Frame {
 public:
  Frame(std::shared_ptr<FrontEndContext> fe, unsigned long id) : mnId(id), fe_(fe) {}
  std::shared_ptr<FrontEndContext> context() const { return fe_; }

  // Frame::ComputeStereoMatches (src/Frame.cc:835-1005): fills mvuRight / mvDepth for the frame extracted last.
  void ComputeStereoMatches() {
    ft_check(ft_stereo_match(fe_->get()));
    fetchStereo(false);
  }
  // Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1231-1271)
  void ComputeStereoFishEyeMatches() {
    ft_check(ft_stereo_match_fisheye(fe_->get()));
    fetchStereo(true);
  }
  void SetPose(const float Rcw[9], const float tcw[3]) {
    memcpy(mRcw, Rcw, sizeof(mRcw)); memcpy(mtcw, tcw, sizeof(mtcw));
    ft_check(ft_set_pose(fe_->get(), Rcw, tcw, nullptr, nullptr));
  }
  // Frame::ComputeBoW (src/Frame.cc:762-769) on the device-resident descriptors of the frame extracted last
  void ComputeBoW() {
    if (!mBowVec.empty() || !mpORBvocabulary) return;
    ft_check(ft_compute_bow(fe_->get(), mpORBvocabulary->get(), 4));
    const int cap = 2 * ft_max_keypoints(fe_->get());
    std::vector<int> node(cap);
    std::vector<uint32_t> ids(cap);
    std::vector<double> vals(cap);
    int nb = 0, n = 0;
    ft_check(ft_bow_download(fe_->get(), cap, nullptr, node.data(), ids.data(), vals.data(), &nb, &n));
    ORBVocabulary::fill(node.data(), n, ids.data(), vals.data(), nb, mBowVec, mFeatVec);
    mvFeatNode.assign(node.begin(), node.begin() + n);
  }
  ORBVocabulary* mpORBvocabulary = nullptr;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  std::vector<int> mvFeatNode;      // mFeatVec as one node per feature (-1 = none), the form the device search takes
  float mRcw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, mtcw[3] = {0, 0, 0};   // Tcw (Frame::SetPose)
  std::vector<bool> mvbOutlier;

  unsigned long mnId;
  int N = 0, Nleft = -1, Nright = -1;
  std::vector<ftcv::KeyPoint> mvKeys, mvKeysRight;
  ftcv::Mat mDescriptors, mDescriptorsRight;
  std::vector<float> mvuRight, mvDepth;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  std::vector<float> mvStereo3Dpoints;      // Nleft x 3
  std::vector<MapPoint*> mvpMapPoints;

 private:
  void fetchStereo(bool fisheye) {
    int nl = 0, nr = 0, ml = 0, mr = 0;
    ft_check(ft_frame_counts(fe_->get(), &nl, &nr, &ml, &mr));
    mvuRight.assign(nl, -1.f); mvDepth.assign(nl, -1.f);
    if (fisheye) {
      Nleft = nl; Nright = nr; N = nl + nr;
      mvLeftToRightMatch.assign(nl, -1); mvRightToLeftMatch.assign(nr, -1); mvStereo3Dpoints.assign((size_t)nl * 3, 0.f);
      ft_check(ft_frame_download(fe_->get(), 0, 0, nullptr, nullptr, nullptr, nullptr, mvuRight.data(), mvDepth.data(),
                                 mvLeftToRightMatch.data(), mvRightToLeftMatch.data(), mvStereo3Dpoints.data()));
    } else {
      N = nl; Nleft = Nright = -1;
      ft_check(ft_frame_download(fe_->get(), 0, 0, nullptr, nullptr, nullptr, nullptr, mvuRight.data(), mvDepth.data(),
                                 nullptr, nullptr, nullptr));
    }
    mvpMapPoints.assign(N, nullptr);
    mvbOutlier.assign(N, false);
  }
  std::shared_ptr<FrontEndContext> fe_;
};

// The fields of KeyFrame that SearchByBoW reads (reference include/KeyFrame.h): descriptors and keypoint angles over the
// (left, right) keypoints, the map-point matches and the FeatureVector made by KeyFrame::ComputeBoW (src/KeyFrame.cc:98-108).
struct KeyFrame {
  int N = 0, NLeft = -1;
  std::vector<unsigned char> mDescriptors;          // N x 32
  std::vector<float> mvAngles;                      // mvKeysUn[i].angle (mvKeys / mvKeysRight for two-camera rigs)
  std::vector<MapPoint*> mvpMapPoints;
  ORBVocabulary* mpORBvocabulary = nullptr;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  std::vector<int> mvFeatNode;
  const std::vector<MapPoint*>& GetMapPointMatches() const { return mvpMapPoints; }
  void ComputeBoW() {
    if (!mBowVec.empty() || !mpORBvocabulary) return;
    mpORBvocabulary->transform(mDescriptors.data(), N, mBowVec, mFeatVec, 4);
    mvFeatNode.assign(N, -1);
    for (const auto& e : mFeatVec) for (unsigned i : e.second) mvFeatNode[i] = (int)e.first;
  }
};

// Host side of the persistent device-side MapPoint store (ft_map_store_*): a MapPoint keeps one row for its lifetime;
// Flush() upserts the rows whose MapPoint changed since the last flush. Replaces the per-frame CudaMapPoint marshalling
// (reference src/Kernels/CudaWrappers/CudaMapPoint.cc:15-34, src/Tracking.cc:3595-3632).
class MapStore {
 public:
  MapStore(std::shared_ptr<FrontEndContext> fe, int capacity) : fe_(fe), cap_(capacity) {
    ft_check(ft_map_store_create(fe_->get(), capacity));
  }
  // gives pMP a row (no-op when it has one) and queues it for the next Flush()
  void Insert(MapPoint* pMP) {
    std::lock_guard<std::mutex> lk(mu_);     // LocalMapping inserts / culls while Tracking flushes
    if (pMP->mnStoreRow < 0) {
      if (!free_.empty()) { pMP->mnStoreRow = free_.back(); free_.pop_back(); }
      else if (next_ < cap_) pMP->mnStoreRow = next_++;
      else throw FtError(FT_ERR_CAPACITY, "MapStore: no free row");
      pMP->mbStoreDirty = true;
    }
    if (pMP->mbStoreDirty) pending_.push_back(pMP);
  }
  void Erase(MapPoint* pMP) {   // MapPoint::SetBadFlag / culling
    std::lock_guard<std::mutex> lk(mu_);
    if (pMP->mnStoreRow >= 0) free_.push_back(pMP->mnStoreRow);
    pMP->mnStoreRow = -1;
  }
  void Flush() {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<int> rows; std::vector<float> pos, nrm, mm; std::vector<unsigned char> desc;
    for (MapPoint* p : pending_) {
      if (!p->mbStoreDirty || p->mnStoreRow < 0) continue;
      rows.push_back(p->mnStoreRow);
      pos.insert(pos.end(), p->mWorldPos, p->mWorldPos + 3); nrm.insert(nrm.end(), p->mNormalVector, p->mNormalVector + 3);
      mm.push_back(p->mfMinDistance); mm.push_back(p->mfMaxDistance);
      desc.insert(desc.end(), p->mDescriptor, p->mDescriptor + 32);
      p->mbStoreDirty = false;
    }
    pending_.clear();
    if (!rows.empty())
      ft_check(ft_map_store_update(fe_->get(), (int)rows.size(), rows.data(), pos.data(), nrm.data(), mm.data(), desc.data()));
  }
  std::shared_ptr<FrontEndContext> context() const { return fe_; }

 private:
  std::shared_ptr<FrontEndContext> fe_;
  int cap_, next_ = 0;
  std::mutex mu_;
  std::vector<int> free_;
  std::vector<MapPoint*> pending_;
};

class ORBmatcher {
 public:
  static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;
  ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

  // Tracking::SearchLocalPoints loop 2 (isInFrustum, viewing-cos limit 0.5) + SearchByProjection #1.
  // Map points are snapshotted at call time (the back-end threads mutate them concurrently in the reference);
  // F.mvpMapPoints is updated in place and the function returns the reference's nmatches.
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                         const float thFarPoints = 50.0f) {
    const int M = (int)vpMapPoints.size(), N = F.N;
    std::vector<float> pos((size_t)M * 3), nrm((size_t)M * 3), mm((size_t)M * 2);
    std::vector<unsigned char> desc((size_t)M * 32);
    std::vector<int> flags(M);
    for (int i = 0; i < M; i++) {
      const MapPoint* p = vpMapPoints[i];
      memcpy(&pos[3 * (size_t)i], p->mWorldPos, 12); memcpy(&nrm[3 * (size_t)i], p->mNormalVector, 12);
      mm[2 * (size_t)i] = p->mfMinDistance; mm[2 * (size_t)i + 1] = p->mfMaxDistance;
      memcpy(&desc[32 * (size_t)i], p->mDescriptor, 32);
      flags[i] = ((p->isBad() || p->mnLastFrameSeen == F.mnId) ? 1 : 0) | (p->Observations() > 0 ? 2 : 0);
    }
    // F.mvpMapPoints as indices: map points of this call by index, any other map point as -2
    std::vector<int> holder(N, -1);
    std::vector<unsigned char> hobs(N, 0);
    std::vector<MapPoint*> foreign(N, nullptr);
    for (int i = 0; i < N; i++) {
      MapPoint* q = F.mvpMapPoints[i];
      if (!q) continue;
      foreign[i] = q; holder[i] = -2; hobs[i] = q->Observations() > 0;
    }
    std::vector<int> best((size_t)std::max(M, 1) * 2, -1);
    int nmatches = 0;
    ft_check(ft_search_local_points(F.context()->get(), M, pos.data(), nrm.data(), mm.data(), desc.data(), flags.data(), th,
                                    bFarPoints ? 1 : 0, thFarPoints, mfNNratio, holder.data(), hobs.data(), best.data(),
                                    &nmatches));
    for (int i = 0; i < N; i++) F.mvpMapPoints[i] = holder[i] >= 0 ? vpMapPoints[holder[i]] : (holder[i] == -2 ? foreign[i] : nullptr);
    if (M > 0) {   // the mTrack* scratch the reference leaves in every MapPoint
      std::vector<int> ti((size_t)M * 4);
      std::vector<float> tf((size_t)M * 9);
      ft_check(ft_debug_track(F.context()->get(), M, ti.data(), tf.data()));
      for (int i = 0; i < M; i++) {
        MapPoint* p = vpMapPoints[i];
        const float* f = &tf[9 * (size_t)i];
        p->mbTrackInView = ti[4 * (size_t)i] != 0; p->mbTrackInViewR = ti[4 * (size_t)i + 1] != 0;
        p->mnTrackScaleLevel = ti[4 * (size_t)i + 2]; p->mnTrackScaleLevelR = ti[4 * (size_t)i + 3];
        p->mTrackProjX = f[0]; p->mTrackProjY = f[1]; p->mTrackProjXR = f[2]; p->mTrackDepth = f[3]; p->mTrackViewCos = f[4];
        p->mTrackProjXR_r = f[5]; p->mTrackProjYR = f[6]; p->mTrackDepthR = f[7]; p->mTrackViewCosR = f[8];
      }
    }
    return nmatches;
  }

  // The same search with the local map named as rows of a MapStore: map points without a row or with pending changes
  // are inserted / flushed first, then 8 bytes per map point (row, flags) go to the device instead of 68.
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, MapStore& store, const float th = 3,
                         const bool bFarPoints = false, const float thFarPoints = 50.0f) {
    const int M = (int)vpMapPoints.size(), N = F.N;
    std::vector<int> rows(M), flags(M);
    for (int i = 0; i < M; i++) {
      MapPoint* p = vpMapPoints[i];
      if (p->mnStoreRow < 0 || p->mbStoreDirty) store.Insert(p);
    }
    store.Flush();
    for (int i = 0; i < M; i++) {
      const MapPoint* p = vpMapPoints[i];
      rows[i] = p->mnStoreRow;
      flags[i] = ((p->isBad() || p->mnLastFrameSeen == F.mnId) ? 1 : 0) | (p->Observations() > 0 ? 2 : 0);
    }
    std::vector<int> holder(N, -1);
    std::vector<unsigned char> hobs(N, 0);
    std::vector<MapPoint*> foreign(N, nullptr);
    for (int i = 0; i < N; i++) {
      MapPoint* q = F.mvpMapPoints[i];
      if (!q) continue;
      foreign[i] = q; holder[i] = -2; hobs[i] = q->Observations() > 0;
    }
    std::vector<int> best((size_t)std::max(M, 1) * 2, -1);
    int nmatches = 0;
    ft_check(ft_search_store(F.context()->get(), M, rows.data(), flags.data(), th, bFarPoints ? 1 : 0, thFarPoints, mfNNratio,
                             holder.data(), hobs.data(), best.data(), &nmatches));
    for (int i = 0; i < N; i++) F.mvpMapPoints[i] = holder[i] >= 0 ? vpMapPoints[holder[i]] : (holder[i] == -2 ? foreign[i] : nullptr);
    return nmatches;
  }

  // Frame-to-last-frame search of Tracking::TrackWithMotionModel (reference src/ORBmatcher.cc:1775-2085).
  // LastFrame needs mvKeys / mvKeysRight (octave, angle), mvpMapPoints, mvbOutlier and its pose; CurrentFrame must be
  // the frame the context extracted last, with SetPose() already called (mVelocity * mLastFrame.GetPose()).
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const int nLast = LastFrame.N, N = CurrentFrame.N;
    std::vector<float> pos((size_t)nLast * 3, 0.f), angle(nLast, 0.f);
    std::vector<unsigned char> desc((size_t)nLast * 32, 0);
    std::vector<int> octave(nLast, 0), flags(nLast, 1);
    std::vector<MapPoint*> src(nLast, nullptr);
    for (int i = 0; i < nLast; i++) {
      MapPoint* p = LastFrame.mvpMapPoints[i];
      const ftcv::KeyPoint& kp = (LastFrame.Nleft == -1 || i < LastFrame.Nleft) ? LastFrame.mvKeys[i]
                                                                                : LastFrame.mvKeysRight[i - LastFrame.Nleft];
      octave[i] = kp.octave; angle[i] = kp.angle;
      if (!p || LastFrame.mvbOutlier[i]) continue;
      src[i] = p;
      memcpy(&pos[3 * (size_t)i], p->mWorldPos, 12); memcpy(&desc[32 * (size_t)i], p->mDescriptor, 32);
      flags[i] = p->Observations() > 0 ? 2 : 0;
    }
    std::vector<int> holder(N, -1);
    std::vector<unsigned char> hobs(N, 0);
    std::vector<MapPoint*> foreign(N, nullptr);
    for (int i = 0; i < N; i++) {
      MapPoint* q = CurrentFrame.mvpMapPoints[i];
      if (!q) continue;
      foreign[i] = q; holder[i] = -2; hobs[i] = q->Observations() > 0;
    }
    int nmatches = 0;
    ft_check(ft_search_last_frame(CurrentFrame.context()->get(), nLast, pos.data(), desc.data(), octave.data(), angle.data(),
                                  flags.data(), LastFrame.mRcw, LastFrame.mtcw, th, bMono ? 1 : 0, mbCheckOrientation ? 1 : 0,
                                  holder.data(), hobs.data(), nullptr, &nmatches));
    for (int i = 0; i < N; i++)
      CurrentFrame.mvpMapPoints[i] = holder[i] >= 0 ? src[holder[i]] : (holder[i] == -2 ? foreign[i] : nullptr);
    return nmatches;
  }

  // Search matches between MapPoints in a KeyFrame and ORB in a Frame, constrained by the vocabulary (reference
  // src/ORBmatcher.cc:322-523; Tracking::TrackReferenceKeyFrame / Relocalization). F must be the frame the context
  // extracted last with ComputeBoW() done; pKF->ComputeBoW() done.
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    std::vector<unsigned char> has(pKF->N > 0 ? pKF->N : 1, 0);
    for (int i = 0; i < pKF->N; i++) has[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
    std::vector<int> match(F.N > 0 ? F.N : 1, -1);
    int nmatches = 0;
    ft_check(ft_search_by_bow(F.context()->get(), pKF->N, pKF->mDescriptors.data(), pKF->mvAngles.data(), pKF->mvFeatNode.data(),
                              has.data(), mfNNratio, mbCheckOrientation ? 1 : 0, match.data(), &nmatches));
    vpMapPointMatches.assign(F.N, nullptr);
    for (int i = 0; i < F.N; i++) if (match[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[match[i]];
    return nmatches;
  }

 protected:
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3
