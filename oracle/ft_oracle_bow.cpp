// ft_oracle_bow.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, see ft_oracle.h): bag-of-words side of the front-end.
//
// Restates, with file:line cites into /root/reference:
//   Frame::ComputeBoW                                  src/Frame.cc:762-769
//   TemplatedVocabulary::loadFromTextFile              Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1423
//   TemplatedVocabulary::transform (features -> Bow/FeatureVector)            :1127-1194
//   TemplatedVocabulary::transform (one feature, tree descent)                :1218-1260
//   FORB::distance                                     Thirdparty/DBoW2/DBoW2/FORB.cpp:80-99
//   BowVector::addWeight / addIfNotExist / normalize   Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-86
//   FeatureVector::addFeature                          Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:31-45
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)    src/ORBmatcher.cc:322-523
//   ORBmatcher::ComputeThreeMaxima                     src/ORBmatcher.cc:2210-2254
// The transform is pinned against the reference's own DBoW2 compiled into oracle/_ref (tests/test_oracle_bow.py), SearchByBoW
// against the reference's own function text compiled into oracle/_ref/libft_ref_frame.so (tests/test_oracle_ref_frame.py).
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#include "ft_oracle.h"

namespace fto {

static void parse_descriptor(const std::string& s, uint8_t* out) {   // FORB::fromString (FORB.cpp:119-135)
  std::stringstream ss(s);
  for (int i = 0; i < 32; ++i) {
    int n;
    ss >> n;
    if (!ss.fail()) out[i] = (uint8_t)n;
  }
}

bool Vocabulary::loadText(const char* path) {
  std::ifstream f;
  f.open(path);
  if (!f.is_open() || f.eof()) return false;
  std::string s;
  std::getline(f, s);
  std::stringstream ss;
  ss << s;
  int n1 = -1, n2 = -1;
  ss >> k; ss >> L; ss >> n1; ss >> n2;
  if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;
  scoring = n1; weighting = n2;
  parent.assign(1, 0); children.assign(1, {}); desc.assign(32, 0); weight.assign(1, 0.0); wordId.assign(1, 0);
  nWords = 0;
  // The reference declares `int pid; int nIsLeaf;` inside the loop without initialising them. For the empty line after
  // a trailing newline every extraction fails at the stream sentry, which leaves its target untouched, so the phantom
  // node takes whatever the stack slots hold: with gcc that is the previous line's parent and leaf flag (checked against
  // the reference's own code in oracle/_ref). Declaring them outside the loop restates that outcome.
  int pid = 0, nIsLeaf = 0;
  while (!f.eof()) {
    std::string snode;
    std::getline(f, snode);
    std::stringstream ssnode;
    ssnode << snode;
    const int nid = (int)parent.size();
    parent.push_back(0); children.emplace_back(); desc.resize(desc.size() + 32, 0); weight.push_back(0.0);
    wordId.push_back(0);
    ssnode >> pid;
    parent[nid] = pid;
    children[pid].push_back(nid);
    ssnode >> nIsLeaf;
    std::stringstream ssd;
    for (int iD = 0; iD < 32; iD++) {
      std::string sElement;
      ssnode >> sElement;
      ssd << sElement << " ";
    }
    parse_descriptor(ssd.str(), &desc[(size_t)nid * 32]);
    ssnode >> weight[nid];
    if (nIsLeaf > 0) wordId[nid] = nWords++;
  }
  return true;
}

void Vocabulary::fromArrays(int k_, int L_, int scoring_, int weighting_, int n, const int* parent_, const uint8_t* isLeaf,
                            const uint8_t* desc_, const double* weight_) {
  k = k_; L = L_; scoring = scoring_; weighting = weighting_;
  parent.assign(n + 1, 0); children.assign(n + 1, {}); desc.assign((size_t)(n + 1) * 32, 0); weight.assign(n + 1, 0.0);
  wordId.assign(n + 1, 0);
  nWords = 0;
  for (int i = 0; i < n; i++) {
    const int nid = i + 1;
    parent[nid] = parent_[i];
    children[parent_[i]].push_back(nid);
    std::memcpy(&desc[(size_t)nid * 32], desc_ + (size_t)i * 32, 32);
    weight[nid] = weight_[i];
    if (isLeaf[i]) wordId[nid] = nWords++;
  }
}

void Vocabulary::transformOne(const uint8_t* f, int levelsup, unsigned& word, double& w, unsigned& nid) const {
  const int nid_level = L - levelsup;
  if (nid_level <= 0) nid = 0;   // root
  int final_id = 0, current_level = 0;
  do {
    ++current_level;
    const std::vector<int>& nodes = children[final_id];
    final_id = nodes[0];
    double best_d = (double)descriptor_distance(f, &desc[(size_t)final_id * 32]);
    for (size_t j = 1; j < nodes.size(); ++j) {
      const int id = nodes[j];
      const double d = (double)descriptor_distance(f, &desc[(size_t)id * 32]);
      if (d < best_d) { best_d = d; final_id = id; }
    }
    if (current_level == nid_level) nid = (unsigned)final_id;
  } while (!children[final_id].empty());
  word = (unsigned)wordId[final_id];
  w = weight[final_id];
}

void voc_transform(const Vocabulary& v, const uint8_t* desc, int n, int levelsup, std::vector<int>& featNode,
                   std::vector<int>& featWord, std::vector<unsigned>& bowIds, std::vector<double>& bowVals) {
  std::map<unsigned, double> bv;
  featNode.assign(n, -1); featWord.assign(n, -1);
  bowIds.clear(); bowVals.clear();
  if (v.parent.size() <= 1) return;   // empty()
  // mustNormalize (ScoringObject.h:74-89): every scoring but DOT_PRODUCT normalises, L2_NORM with the L2 norm
  const bool must = v.scoring != 5;
  const bool l2 = v.scoring == 1;
  const bool tf = v.weighting == 0 || v.weighting == 1;   // TF_IDF or TF
  for (int i = 0; i < n; i++) {
    unsigned id = 0, nid = 0; double w = 0;
    v.transformOne(desc + 32 * (size_t)i, levelsup, id, w, nid);
    featWord[i] = (int)id;
    if (w > 0) {   // not stopped
      auto it = bv.lower_bound(id);
      if (it != bv.end() && it->first == id) { if (tf) it->second += w; }   // addWeight / addIfNotExist
      else bv.insert(it, std::make_pair(id, w));
      featNode[i] = (int)nid;   // addFeature: ascending feature index inside a node by construction
    }
  }
  if (tf && !bv.empty() && !must) {
    const double nd = (double)bv.size();
    for (auto& e : bv) e.second /= nd;
  }
  if (must) {   // BowVector::normalize
    double norm = 0.0;
    if (!l2) { for (auto& e : bv) norm += std::fabs(e.second); }
    else { for (auto& e : bv) norm += e.second * e.second; norm = std::sqrt(norm); }
    if (norm > 0.0) for (auto& e : bv) e.second /= norm;
  }
  for (auto& e : bv) { bowIds.push_back(e.first); bowVals.push_back(e.second); }
}

static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
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

int search_by_bow(int nKF, const uint8_t* kfDesc, const float* kfAngle, const int* kfNode, const uint8_t* kfHasMp,
                  int nF, const uint8_t* fDesc, const float* fAngle, const int* fNode, int fNleft, float nnratio,
                  bool checkOrientation, std::vector<int>& match) {
  const int TH_LOW = 50, HISTO_LENGTH = 30;
  // the two FeatureVectors (std::map<NodeId, std::vector<unsigned>>)
  std::map<unsigned, std::vector<unsigned>> fvKF, fvF;
  for (int i = 0; i < nKF; i++) if (kfNode[i] >= 0) fvKF[(unsigned)kfNode[i]].push_back((unsigned)i);
  for (int i = 0; i < nF; i++) if (fNode[i] >= 0) fvF[(unsigned)fNode[i]].push_back((unsigned)i);
  match.assign(nF, -1);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  auto KFit = fvKF.begin(), KFend = fvKF.end();
  auto Fit = fvF.begin(), Fend = fvF.end();
  auto vote = [&](int idxKF, int idxF) {
    float rot = kfAngle[idxKF] - fAngle[idxF];
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    rotHist[bin].push_back(idxF);
  };
  while (KFit != KFend && Fit != Fend) {
    if (KFit->first == Fit->first) {
      const std::vector<unsigned>& vIndicesKF = KFit->second;
      const std::vector<unsigned>& vIndicesF = Fit->second;
      for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
        const unsigned realIdxKF = vIndicesKF[iKF];
        if (!kfHasMp[realIdxKF]) continue;
        const uint8_t* dKF = kfDesc + 32 * (size_t)realIdxKF;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        int bestDist1R = 256, bestIdxFR = -1, bestDist2R = 256;
        for (size_t iF = 0; iF < vIndicesF.size(); iF++) {
          const unsigned realIdxF = vIndicesF[iF];
          if (match[realIdxF] >= 0) continue;
          const int dist = descriptor_distance(dKF, fDesc + 32 * (size_t)realIdxF);
          if (fNleft == -1) {
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
            else if (dist < bestDist2) bestDist2 = dist;
          } else {
            if ((int)realIdxF < fNleft && dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
            else if ((int)realIdxF < fNleft && dist < bestDist2) bestDist2 = dist;
            if ((int)realIdxF >= fNleft && dist < bestDist1R) { bestDist2R = bestDist1R; bestDist1R = dist; bestIdxFR = (int)realIdxF; }
            else if ((int)realIdxF >= fNleft && dist < bestDist2R) bestDist2R = dist;
          }
        }
        if (bestDist1 <= TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            match[bestIdxF] = (int)realIdxKF;
            if (checkOrientation) vote((int)realIdxKF, bestIdxF);
            nmatches++;
          }
          if (bestDist1R <= TH_LOW) {   // `ratio || true` in the reference (:451)
            match[bestIdxFR] = (int)realIdxKF;
            if (checkOrientation) vote((int)realIdxKF, bestIdxFR);
            nmatches++;
          }
        }
      }
      ++KFit; ++Fit;
    } else if (KFit->first < Fit->first) {
      KFit = fvKF.lower_bound(Fit->first);
    } else {
      Fit = fvF.lower_bound(KFit->first);
    }
  }
  if (checkOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { match[idx] = -1; nmatches--; }
    }
  }
  return nmatches;
}

}  // namespace fto

// ---- flat C entry points for ctypes ----
using namespace fto;
extern "C" {

void* fto_voc_load_text(const char* path) {
  Vocabulary* v = new Vocabulary();
  if (!v->loadText(path)) { delete v; return nullptr; }
  return v;
}
void* fto_voc_from_arrays(int k, int L, int scoring, int weighting, int n, const int* parent, const uint8_t* isLeaf,
                          const uint8_t* desc, const double* weight) {
  Vocabulary* v = new Vocabulary();
  v->fromArrays(k, L, scoring, weighting, n, parent, isLeaf, desc, weight);
  return v;
}
void fto_voc_free(void* v) { delete (Vocabulary*)v; }
void fto_voc_info(void* v_, int* out /* k L scoring weighting nodes(with root) words */) {
  Vocabulary* v = (Vocabulary*)v_;
  out[0] = v->k; out[1] = v->L; out[2] = v->scoring; out[3] = v->weighting; out[4] = (int)v->parent.size();
  out[5] = v->nWords;
}
// arrays of the loaded tree without the root (node i+1 -> entry i): what ft_vocabulary_create takes
void fto_voc_arrays(void* v_, int* parent, uint8_t* isLeaf, uint8_t* desc, double* weight) {
  Vocabulary* v = (Vocabulary*)v_;
  const int n = (int)v->parent.size() - 1;
  for (int i = 0; i < n; i++) {
    parent[i] = v->parent[i + 1];
    isLeaf[i] = v->children[i + 1].empty() ? 1 : 0;
    weight[i] = v->weight[i + 1];
  }
  std::memcpy(desc, v->desc.data() + 32, (size_t)n * 32);
}
int fto_voc_transform(void* v_, const uint8_t* desc, int n, int levelsup, int* featNode, int* featWord, unsigned* bowIds,
                      double* bowVals, int bowCap) {
  std::vector<int> fn, fw; std::vector<unsigned> ids; std::vector<double> vals;
  voc_transform(*(Vocabulary*)v_, desc, n, levelsup, fn, fw, ids, vals);
  std::memcpy(featNode, fn.data(), sizeof(int) * n);
  if (featWord) std::memcpy(featWord, fw.data(), sizeof(int) * n);
  for (size_t i = 0; i < ids.size() && (int)i < bowCap; i++) { bowIds[i] = ids[i]; bowVals[i] = vals[i]; }
  return (int)ids.size();
}
int fto_search_by_bow(int nKF, const uint8_t* kfDesc, const float* kfAngle, const int* kfNode, const uint8_t* kfHasMp,
                      int nF, const uint8_t* fDesc, const float* fAngle, const int* fNode, int fNleft, float nnratio,
                      int checkOri, int* match) {
  std::vector<int> m;
  const int nm = search_by_bow(nKF, kfDesc, kfAngle, kfNode, kfHasMp, nF, fDesc, fAngle, fNode, fNleft, nnratio,
                               checkOri != 0, m);
  std::memcpy(match, m.data(), sizeof(int) * nF);
  return nm;
}

}  // extern "C"
