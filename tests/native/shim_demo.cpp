// Drives the host-side mirror (fasttrack_b200/host/ft_shim.h) the way the reference's Frame constructor and
// Tracking::SearchLocalPoints do: two threads call the two extractors (Frame.cc:127-130), then
// ComputeStereoMatches, then ORBmatcher(0.8).SearchByProjection over a local map. Inputs/outputs are raw
// binary files exchanged with tests/test_gpu_shim.py.
#include <cstdio>
#include <fstream>
#include <thread>

#include "../../fasttrack_b200/host/ft_shim.h"

using namespace ORB_SLAM3;

template <typename T>
static std::vector<T> rd(const std::string& p) {
  std::ifstream f(p, std::ios::binary | std::ios::ate);
  if (!f) { fprintf(stderr, "shim_demo: cannot open %s\n", p.c_str()); exit(2); }
  std::vector<T> v((size_t)f.tellg() / sizeof(T));
  f.seekg(0); f.read((char*)v.data(), v.size() * sizeof(T));
  return v;
}
template <typename T>
static void wr(const std::string& p, const std::vector<T>& v) { std::ofstream f(p, std::ios::binary); f.write((const char*)v.data(), v.size() * sizeof(T)); }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  const int W = 752, H = 480;
  std::vector<unsigned char> L = rd<unsigned char>(dir + "/L.bin"), R = rd<unsigned char>(dir + "/R.bin");
  std::vector<float> cam = rd<float>(dir + "/cam.bin");   // fx fy cx cy bf
  ft_config cfg = {};
  cfg.device_id = 0; cfg.width = W; cfg.height = H; cfg.nfeatures = 1200; cfg.nlevels = 8; cfg.scale_factor = 1.2f;
  cfg.ini_th_fast = 20; cfg.min_th_fast = 7; cfg.camera_type = FT_CAM_PINHOLE;
  for (int i = 0; i < 4; i++) cfg.cam1[i] = cfg.cam2[i] = cam[i];
  cfg.bf = cam[4]; cfg.max_map_points = 25000;
  try {
    auto fe = std::make_shared<FrontEndContext>(cfg);
    ORBextractor exL(fe, 0), exR(fe, 1);
    Frame F(fe, 7);
    ftcv::Mat imL(H, W, L.data(), W), imR(H, W, R.data(), W), mask;
    std::vector<int> lap = {0, 0};
    int monoL = -2, monoR = -2;
    std::thread tl([&] { monoL = exL(imL, mask, F.mvKeys, F.mDescriptors, lap); });
    std::thread tr([&] { monoR = exR(imR, mask, F.mvKeysRight, F.mDescriptorsRight, lap); });
    tl.join(); tr.join();
    F.ComputeStereoMatches();
    // local map
    std::vector<float> pos = rd<float>(dir + "/mp_pos.bin"), nrm = rd<float>(dir + "/mp_normal.bin"), mm = rd<float>(dir + "/mp_minmax.bin");
    std::vector<unsigned char> dsc = rd<unsigned char>(dir + "/mp_desc.bin");
    std::vector<int> flg = rd<int>(dir + "/mp_flags.bin");
    const int M = (int)flg.size();
    std::vector<MapPoint> mps(M);
    std::vector<MapPoint*> vp(M);
    for (int i = 0; i < M; i++) {
      memcpy(mps[i].mWorldPos, &pos[3 * i], 12); memcpy(mps[i].mNormalVector, &nrm[3 * i], 12);
      mps[i].mfMinDistance = mm[2 * i]; mps[i].mfMaxDistance = mm[2 * i + 1];
      memcpy(mps[i].mDescriptor, &dsc[32 * i], 32);
      mps[i].mbBad = flg[i] & 1; mps[i].nObs = (flg[i] & 2) ? 3 : 0;
      vp[i] = &mps[i];
    }
    const float Rcw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tcw[3] = {0, 0, 0};
    F.SetPose(Rcw, tcw);
    ORBmatcher matcher(0.8f);
    const int nm = matcher.SearchByProjection(F, vp, 3.0f, false, 50.0f);
    // the same search over the persistent map store: identical assignments, also after a change + re-flush
    std::vector<MapPoint*> snapshot = F.mvpMapPoints;
    MapStore store(fe, 30000);
    int storeOk = 1;
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1) {   // the mapping side moves some points: both paths must agree again
        for (int i = 0; i < M; i += 7) { mps[i].mWorldPos[0] += 0.01f; mps[i].mbStoreDirty = true; }
        F.mvpMapPoints.assign(F.N, nullptr);
        matcher.SearchByProjection(F, vp, 3.0f, false, 50.0f);
        snapshot = F.mvpMapPoints;
      }
      F.mvpMapPoints.assign(F.N, nullptr);
      const int nms = matcher.SearchByProjection(F, vp, store, 3.0f, false, 50.0f);
      if (F.mvpMapPoints != snapshot) storeOk = 0;
      if (pass == 0 && nms != nm) storeOk = 0;
    }
    if (!storeOk) { fprintf(stderr, "shim_demo: MapStore search differs from the snapshot search\n"); return 3; }
    // restore the first result for the outputs below
    for (int i = 0; i < M; i += 7) { mps[i].mWorldPos[0] -= 0.01f; }
    F.mvpMapPoints.assign(F.N, nullptr);
    matcher.SearchByProjection(F, vp, 3.0f, false, 50.0f);
    // bag of words, when the test provides a vocabulary and a KeyFrame: Frame::ComputeBoW, KeyFrame::ComputeBoW,
    // ORBmatcher(0.7).SearchByBoW(pKF, F, matches) as Tracking::TrackReferenceKeyFrame drives them
    std::ifstream vocProbe(dir + "/voc.txt");
    if (vocProbe.good()) {
      ORBVocabulary voc(0);
      if (!voc.loadFromTextFile(dir + "/voc.txt")) { fprintf(stderr, "shim_demo: vocabulary: %s\n", ft_last_error()); return 4; }
      F.mpORBvocabulary = &voc;
      F.ComputeBoW();
      KeyFrame kf;
      kf.mDescriptors = rd<unsigned char>(dir + "/kf_desc.bin");
      kf.mvAngles = rd<float>(dir + "/kf_angle.bin");
      std::vector<unsigned char> kfHas = rd<unsigned char>(dir + "/kf_has.bin");
      kf.N = (int)kf.mvAngles.size();
      std::vector<MapPoint> kfMps(kf.N);
      kf.mvpMapPoints.assign(kf.N, nullptr);
      for (int i = 0; i < kf.N; i++) if (kfHas[i]) kf.mvpMapPoints[i] = &kfMps[i];
      kf.mpORBvocabulary = &voc;
      kf.ComputeBoW();
      ORBmatcher bowMatcher(0.7f, true);
      std::vector<MapPoint*> vpMatches;
      const int nbow = bowMatcher.SearchByBoW(&kf, F, vpMatches);
      std::vector<int> bowMatch(F.N, -1);
      for (int i = 0; i < F.N; i++) if (vpMatches[i]) bowMatch[i] = (int)(vpMatches[i] - kfMps.data());
      bowMatch.push_back(nbow);
      wr(dir + "/out_bow_match.bin", bowMatch);
      std::vector<double> bv;
      for (const auto& e : F.mBowVec) { bv.push_back((double)e.first); bv.push_back(e.second); }
      wr(dir + "/out_bow_vec.bin", bv);
    }
    // outputs
    std::vector<float> k;
    for (auto& kp : F.mvKeys) { k.push_back(kp.pt.x); k.push_back(kp.pt.y); k.push_back(kp.size); k.push_back(kp.angle); k.push_back(kp.response); k.push_back((float)kp.octave); }
    wr(dir + "/out_kL.bin", k);
    std::vector<unsigned char> d;
    for (int i = 0; i < F.mDescriptors.rows; i++) d.insert(d.end(), F.mDescriptors.ptr(i), F.mDescriptors.ptr(i) + 32);
    wr(dir + "/out_dL.bin", d);
    wr(dir + "/out_uRight.bin", F.mvuRight);
    wr(dir + "/out_depth.bin", F.mvDepth);
    std::vector<int> holder(F.N, -1);
    for (int i = 0; i < F.N; i++) if (F.mvpMapPoints[i]) holder[i] = (int)(F.mvpMapPoints[i] - mps.data());
    wr(dir + "/out_holder.bin", holder);
    std::vector<int> meta = {monoL, monoR, (int)F.mvKeys.size(), (int)F.mvKeysRight.size(), nm};
    wr(dir + "/out_meta.bin", meta);
    int inview = 0;
    for (auto& m : mps) inview += m.mbTrackInView;
    printf("shim_demo: nL=%zu nR=%zu mono=%d/%d matches=%d inView=%d\n", F.mvKeys.size(), F.mvKeysRight.size(), monoL, monoR, nm, inview);
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_demo failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
