// sift_ref.cpp -- ORACLE-SIDE TOOLING (test infrastructure): the step BEFORE the path, used only to produce the
// BASELINE config-1 inputs (SIFT regions of the reference's bundled data/imageData images) with the reference's
// OWN wrapper: SIFTDetector<unsigned char>(img, feats, descs, is_zoom=false, root_sift=true, 0.04f)
// (libs/feature/include/mvg/feature/sift.hpp:67-134, as called at apps/compute_matches/compute_matches.cpp:209-211)
// over the vendored VLFeat subset (3rdparty/sift/vl/*.c).  Compiled by oracle/build_sift_ref.sh from the sources
// where they lie; only tests/golden/make_golden_imagedata.py loads the result.
#include <cstdint>
#include <cstring>
#include <vector>

#include "mvg/feature/sift.hpp"

using namespace mvg::feature;
using namespace mvg::image;

extern "C" int ref_sift_u8(const unsigned char* gray, int w, int h, int is_zoom, int root_sift, float contrast,
                           float* feats_out, unsigned char* descs_out, int cap) {
  Image<unsigned char> img(w, h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) img(y, x) = gray[(size_t)y * w + x];
  std::vector<ScalePointFeature> feats;
  std::vector<Descriptor<unsigned char, 128> > descs;
  SIFTDetector<unsigned char>(img, feats, descs, is_zoom != 0, root_sift != 0, contrast);
  const int n = (int)feats.size();
  for (int k = 0; k < n && k < cap; ++k) {
    feats_out[4 * k + 0] = feats[k].x();
    feats_out[4 * k + 1] = feats[k].y();
    feats_out[4 * k + 2] = feats[k].scale();
    feats_out[4 * k + 3] = feats[k].orientation();
    memcpy(descs_out + (size_t)k * 128, descs[k].getData(), 128);
  }
  return n;
}
