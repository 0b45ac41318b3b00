// ref_bow_build.cpp — compiles the reference's own vendored DBoW2 (third_part/DBoW2, UNMODIFIED, included from
// where it lies under /root/reference) against oracle/shim so the restated vocabulary transform in tb_oracle.c
// can be pinned to it.  ORACLE / TEST INFRASTRUCTURE ONLY; output goes to oracle/_ref/libtb_ref_bow.so.
// The vocabulary is loaded by the reference's loadFromTextFile (TemplatedVocabulary.h:1338-1420) and features are
// transformed through its public API exactly as Frame::SetBow does (src/types/Frame.cpp:267-270).
#include "cvshim.h"
#include "cvshim_filestorage.h"
#include "DBoW2/FORB.cpp"
#include "DBoW2/BowVector.cpp"
#include "DBoW2/FeatureVector.cpp"
#include "DBoW2/ScoringObject.cpp"
#include "DUtils/Random.cpp"
#include "DUtils/Timestamp.cpp"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;  // include/types/Frame.h:22

extern "C" {

void *tbref_vocab_load_text(const char *path)
{
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromTextFile(path) || v->empty()) { delete v; return nullptr; }
    return v;
}
void tbref_vocab_free(void *h) { delete (ORBVocabulary *)h; }
int tbref_vocab_size(void *h) { return (int)((ORBVocabulary *)h)->size(); }

// Frame::SetBow: voc->transform(descriptors, bowVec, featVec, levelsup).  Outputs: per-feature node id (-1 when the
// feature is absent from the FeatureVector), per-feature word id (public single-feature transform), and the
// BowVector as (ids, values) in map order.  Returns the BowVector size.
int tbref_set_bow(void *h, const unsigned char *desc, int n, int levelsup, int *node_id, int *word_id,
                  int *bow_ids, double *bow_vals, int bow_cap)
{
    ORBVocabulary *v = (ORBVocabulary *)h;
    std::vector<cv::Mat> feats;
    for (int i = 0; i < n; i++) feats.push_back(cv::Mat(1, 32, CV_8U, (void *)(desc + 32 * (size_t)i), 32).clone());
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    v->transform(feats, bv, fv, levelsup);
    for (int i = 0; i < n; i++) { node_id[i] = -1; word_id[i] = (int)v->transform(feats[i]); }
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t k = 0; k < it->second.size(); k++) node_id[it->second[k]] = (int)it->first;
    int m = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++m)
        if (m < bow_cap) { bow_ids[m] = (int)it->first; bow_vals[m] = it->second; }
    return m;
}

}  // extern "C"
