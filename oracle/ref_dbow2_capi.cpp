// ref_dbow2_capi.cpp -- C entry points over the REFERENCE's own DBoW2 (TEST INFRASTRUCTURE ONLY).
//
// Compiled by oracle/Makefile together with the reference's Thirdparty/DBoW2/DBoW2/{FORB,BowVector,FeatureVector,
// ScoringObject}.cpp, from where they lie under /root/reference, against the stand-in headers in oracle/ref_stubs,
// into oracle/_ref/libft_ref_dbow2.so. It pins the oracle's restatement of the vocabulary transform
// (Frame::ComputeBoW -> ORBVocabulary::transform, reference src/Frame.cc:762-769) against the code the reference runs.
// Nothing of the reference is copied into this repository.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

// The k-means training path (TemplatedVocabulary::create, virtual and therefore instantiated) refers to DUtils::Random,
// whose .cpp needs a header the reference does not ship (DUtils/Timestamp.h); training is never called here.
namespace DUtils {
void Random::SeedRandOnce() {}
int Random::RandomInt(int min, int) { return min; }
}  // namespace DUtils

// the reference's typedef (include/ORBVocabulary.h:28-29)
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

extern "C" {

void* ftref_voc_load_text(const char* path) {
  ORBVocabulary* v = new ORBVocabulary();
  if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
  return v;
}
void ftref_voc_free(void* h) { delete static_cast<ORBVocabulary*>(h); }
int ftref_voc_words(void* h) { return (int)static_cast<ORBVocabulary*>(h)->size(); }

// Frame::ComputeBoW: transform(descriptors, BowVec, FeatVec, levelsup). Outputs: per-feature node id (-1 = not in the
// FeatureVector) from the FeatureVector map; BowVector as (ids, values) in map order; returns the number of words.
int ftref_voc_transform(void* h, const unsigned char* desc, int n, int levelsup, int* feat_node, unsigned* bow_ids,
                        double* bow_vals, int bow_cap, int* featvec_order /* [n] feature indices in map order */,
                        int* n_featvec) {
  ORBVocabulary* voc = static_cast<ORBVocabulary*>(h);
  std::vector<cv::Mat> feats(n);
  for (int i = 0; i < n; ++i) {
    feats[i].create(1, 32, CV_8U);
    std::memcpy(feats[i].ptr<unsigned char>(), desc + 32 * (size_t)i, 32);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  voc->transform(feats, bv, fv, levelsup);
  for (int i = 0; i < n; ++i) feat_node[i] = -1;
  int k = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
    for (size_t j = 0; j < it->second.size(); ++j) {
      feat_node[it->second[j]] = (int)it->first;
      featvec_order[k++] = (int)it->second[j];
    }
  *n_featvec = k;
  int w = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++w)
    if (w < bow_cap) { bow_ids[w] = it->first; bow_vals[w] = it->second; }
  return w;
}

// single-feature transform: word id, weight, node id `levelsup` levels above the word
void ftref_voc_transform_one(void* h, const unsigned char* desc, int levelsup, unsigned* word, double* weight,
                             unsigned* node) {
  ORBVocabulary* voc = static_cast<ORBVocabulary*>(h);
  cv::Mat f(1, 32, CV_8U);
  std::memcpy(f.ptr<unsigned char>(), desc, 32);
  DBoW2::BowVector bv; DBoW2::FeatureVector fv;
  std::vector<cv::Mat> one(1, f);
  voc->transform(one, bv, fv, levelsup);
  *word = bv.empty() ? 0xFFFFFFFFu : bv.begin()->first;
  *weight = bv.empty() ? 0.0 : bv.begin()->second;
  *node = fv.empty() ? 0xFFFFFFFFu : fv.begin()->first;
}

}  // extern "C"
