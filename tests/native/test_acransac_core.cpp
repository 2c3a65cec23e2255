// Host-side unit test of csrc/acransac_core.cuh (the scalar building blocks of the GPU AC-RANSAC filter) against the
// reference's OWN code (oracle/_ref/libmvgref_geom.so = the reference's headers + Eigen 3.2.2 compiled in place): glibc
// rand() stream, RandomSample, NormalizePoints, SevenPointSolver::Solve and EpipolarDistanceError, compared BIT FOR BIT.
// Built and run by tests/test_oracle.py when the reference tree is present.  TEST INFRASTRUCTURE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <algorithm>
#include <cmath>
#include <utility>
#include "../../3dreconstruction_b200/csrc/acransac_core.cuh"
#define MVG_SORT_STATS 1
#include "../../3dreconstruction_b200/csrc/stdsort_restated.cuh"

extern "C" {
int ref_seven_point(const double* x1, const double* x2, double* F);
void ref_normalize(const float* pts, int n, int width, int height, double* out, double* T);
double ref_epipolar_error(const double* F, double x1, double y1, double x2, double y2);
int ref_four_point(const double* x1, const double* x2, double* H);
double ref_homography_error(const double* H, double x1, double y1, double x2, double y2);
void ref_rand(unsigned seed, int n, unsigned* out);
void ref_random_sample7(unsigned seed, int skip, int n, int* out);
}

using namespace mvgcuda::geo;

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, 8) == 0; }

int main() {
  int bad = 0;
  // 1. rand() stream
  for (unsigned seed : {1u, 0u, 12345u, 4000000000u}) {
    std::vector<unsigned> want(200000);
    ref_rand(seed, (int)want.size(), want.data());
    GlibcRand g;
    glibc_srand(g, seed);
    for (size_t i = 0; i < want.size(); ++i)
      if (glibc_rand(g) != want[i]) { std::printf("rand mismatch seed %u at %zu\n", seed, i); ++bad; break; }
  }
  // 2. RandomSample
  {
    std::vector<unsigned> stream(100000);
    ref_rand(1, (int)stream.size(), stream.data());
    for (int t = 0; t < 2000; ++t) {
      const int n = 8 + (t * 37) % 900, skip = t * 7;
      int want[7], got[7];
      ref_random_sample7(1, skip, n, want);
      random_sample<7>(stream.data() + skip, n, got);
      if (std::memcmp(want, got, sizeof want)) { std::printf("sample mismatch n %d skip %d\n", n, skip); ++bad; break; }
    }
  }
  // 3. normalisation
  std::mt19937 rng(7);
  {
    std::uniform_real_distribution<float> ux(0.f, 4000.f), uy(0.f, 3000.f);
    for (int wh = 0; wh < 4; ++wh) {
      const int w = wh == 0 ? 4000 : wh == 1 ? 1416 : wh == 2 ? 641 : 3001, h = wh == 0 ? 3000 : wh == 1 ? 1064 : wh == 2 ? 477 : 2003;
      const int n = 500;
      std::vector<float> pts(2 * n);
      for (int i = 0; i < n; ++i) { pts[2 * i] = ux(rng); pts[2 * i + 1] = uy(rng); }
      std::vector<double> want(2 * n);
      double T[9];
      ref_normalize(pts.data(), n, w, h, want.data(), T);
      const Normalizer N = make_normalizer(w, h);
      if (!same_bits(N.d, T[0]) || !same_bits(N.d, T[4]) || !same_bits(N.tx, T[2]) || !same_bits(N.ty, T[5])) { std::printf("normalizer mismatch %d x %d\n", w, h); ++bad; }
      for (int i = 0; i < n; ++i) {
        double x, y;
        normalize_point(N, pts[2 * i], pts[2 * i + 1], x, y);
        if (!same_bits(x, want[2 * i]) || !same_bits(y, want[2 * i + 1])) { std::printf("normalised point mismatch %d\n", i); ++bad; break; }
      }
    }
  }
  // 4. seven-point solver + 5. residual: realistic normalised coordinates (|x| < 1), incl. near-degenerate samples
  {
    std::uniform_real_distribution<double> u(-0.6, 0.6);
    long models = 0, bits_diff = 0, count_diff = 0, one = 0, three = 0, err_diff = 0;
    double W[81], V[81];
    for (int t = 0; t < 20000; ++t) {
      double x1[14], x2[14];
      for (int i = 0; i < 14; ++i) { x1[i] = u(rng); x2[i] = u(rng); }
      if (t % 5 == 1) for (int i = 0; i < 14; ++i) x2[i] = x1[i] + 0.01 * u(rng);          // near-identity motion
      if (t % 5 == 2) { x1[2] = x1[0]; x1[3] = x1[1]; }                                        // duplicated point
      if (t % 5 == 3) for (int i = 0; i < 7; ++i) x1[2 * i + 1] = 0.3 * x1[2 * i] + 0.1;       // collinear points
      double Fw[27], Fg[27];
      const int nw = ref_seven_point(x1, x2, Fw);
      const int ng = seven_point_models(x1, x2, W, V, Fg);
      if (nw != ng) { ++count_diff; continue; }
      (nw == 1 ? one : three) += nw > 0;
      for (int k = 0; k < 9 * nw; ++k) { ++models; if (!same_bits(Fw[k], Fg[k])) ++bits_diff; }
      for (int k = 0; k < nw; ++k) {
        const double a = u(rng), b = u(rng), c = u(rng), d = u(rng);
        if (!same_bits(ref_epipolar_error(Fw + 9 * k, a, b, c, d), epipolar_error(Fw + 9 * k, a, b, c, d))) ++err_diff;
      }
    }
    std::printf("seven-point: %ld coefficients compared, %ld differ in bits, %ld samples with a different model count "
                "(1-root %ld, 3-root %ld), residual mismatches %ld\n", models, bits_diff, count_diff, one, three, err_diff);
    if (bits_diff || count_diff || err_diff) ++bad;
  }
  // 6. four-point homography solver (QR-preconditioned Jacobi SVD of the 16x9 action matrix) + its residual
  {
    std::uniform_real_distribution<double> u(-0.6, 0.6);
    long coeffs = 0, bits_diff = 0, err_diff = 0;
    double A[144], W[81], V[81];
    for (int t = 0; t < 20000; ++t) {
      double x1[8], x2[8];
      for (int i = 0; i < 8; ++i) { x1[i] = u(rng); x2[i] = u(rng); }
      if (t % 5 == 1) for (int i = 0; i < 8; ++i) x2[i] = x1[i] + 0.01 * u(rng);   // near-identity
      if (t % 5 == 2) { x1[2] = x1[0]; x1[3] = x1[1]; }                              // duplicated point
      if (t % 5 == 3) for (int i = 0; i < 4; ++i) x1[2 * i + 1] = 0.3 * x1[2 * i] + 0.1;  // collinear
      double Hw[9], Hg[9];
      ref_four_point(x1, x2, Hw);
      four_point_model(x1, x2, A, W, V, Hg);
      for (int k = 0; k < 9; ++k) { ++coeffs; if (!same_bits(Hw[k], Hg[k])) ++bits_diff; }
      const double a = u(rng), b = u(rng), c = u(rng), d = u(rng);
      if (!same_bits(ref_homography_error(Hw, a, b, c, d), homography_error(Hw, a, b, c, d))) ++err_diff;
    }
    std::printf("four-point: %ld coefficients compared, %ld differ in bits, residual mismatches %ld\n", coeffs, bits_diff, err_diff);
    if (bits_diff || err_diff) ++bad;
  }
  // 7. std::sort restated (csrc/stdsort_restated.cuh) against THIS toolchain's std::sort on (residual, index) arrays with
  //    NaNs in them -- no strict weak ordering, the result is whatever the introsort steps leave.  The real sort runs inside
  //    a padded buffer whose guard entries stop an unguarded scan that leaves the range (cases where the restatement
  //    reports it left the array are not compared: the reference would have read foreign memory there).
  {
    typedef std::pair<double, size_t> EI;
    std::uniform_real_distribution<double> u(0.0, 1.0);
    long cases = 0, differ = 0, left = 0, with_nan = 0, heap_path = 0;
    for (int t = 0; t < 6000; ++t) {
      const int n = t < 200 ? 1 + t % 40 : 17 + (t * 131) % 3000;
      const int kind = t % 6;
      std::vector<double> e(n);
      for (int i = 0; i < n; ++i) e[i] = kind == 4 ? std::floor(8 * u(rng)) : u(rng);   // kind 4: many equal residuals
      const double share = kind == 0 ? 0.0 : kind == 1 ? 0.02 : kind == 2 ? 0.2 : kind == 3 ? 0.7 : 0.1;
      int nn = 0;
      for (int i = 0; i < n; ++i) if (u(rng) < share) { e[i] = std::nan(""); ++nn; }
      if (kind == 5) for (int i = 0; i < n; ++i) if (u(rng) < 0.1) e[i] = INFINITY;
      const int pad = 64;
      std::vector<EI> buf(n + 2 * pad);
      for (int i = 0; i < pad; ++i) { buf[i] = EI(-INFINITY, 0); buf[pad + n + i] = EI(INFINITY, (size_t)-1); }
      for (int i = 0; i < n; ++i) buf[pad + i] = EI(e[i], (size_t)i);
      std::sort(buf.begin() + pad, buf.begin() + pad + n);
      std::vector<double> ge(e);
      std::vector<int> gi(n);
      for (int i = 0; i < n; ++i) gi[i] = i;
      const int la = libstdcxx_sort(ge.data(), gi.data(), n);
      ++cases; with_nan += nn > 0;
      if (la) { ++left; continue; }
      bool same = true;
      for (int i = 0; i < n && same; ++i) same = (int)buf[pad + i].second == gi[i] && same_bits(buf[pad + i].first, ge[i]);
      if (!same) { if (differ < 5) std::printf("std::sort restatement differs: case %d n %d kind %d NaNs %d\n", t, n, kind, nn); ++differ; }
    }
    // adversarial for the depth limit (heap sort branch): organ-pipe / median-of-3 killer-ish inputs
    for (int t = 0; t < 40; ++t) {
      const int n = 200 + 97 * t;
      std::vector<double> e(n);
      for (int i = 0; i < n; ++i) e[i] = (t & 1) ? (double)((i * 7919) % 13) : (i < n / 2 ? i : n - i);
      if (t % 4 >= 2) for (int i = 0; i < n; i += 9) e[i] = std::nan("");
      std::vector<EI> buf(n + 128);
      for (int i = 0; i < 64; ++i) { buf[i] = EI(-INFINITY, 0); buf[64 + n + i] = EI(INFINITY, (size_t)-1); }
      for (int i = 0; i < n; ++i) buf[64 + i] = EI(e[i], (size_t)i);
      std::sort(buf.begin() + 64, buf.begin() + 64 + n);
      std::vector<int> gi(n);
      for (int i = 0; i < n; ++i) gi[i] = i;
      const int la = libstdcxx_sort(e.data(), gi.data(), n);
      ++cases; ++heap_path;
      if (la) { ++left; continue; }
      bool same = true;
      for (int i = 0; i < n && same; ++i) same = (int)buf[64 + i].second == gi[i];
      if (!same) { ++differ; std::printf("std::sort restatement differs on structured case %d\n", t); }
    }
    // median-of-3 killer sequences (Musser 1997): every partition peels off two elements, the depth limit 2 floor(log2 n)
    // is reached and the rest of the range goes through the heap-sort branch -- with and without NaNs sprinkled in
    for (int t = 0; t < 24; ++t) {
      const int n = 2 * (300 + 211 * t), k = n / 2;
      std::vector<double> e(n);
      for (int i = 1; i <= k; ++i) {
        if (i & 1) { e[i - 1] = i; e[i] = k + i; }
        e[k + i - 1] = 2 * i;
      }
      if (t % 3 == 1) for (int i = 5; i < n; i += 97) e[i] = std::nan("");
      if (t % 3 == 2) for (int i = 0; i < n; i += 3) e[i] = std::nan("");
      std::vector<EI> buf(n + 128);
      for (int i = 0; i < 64; ++i) { buf[i] = EI(-INFINITY, 0); buf[64 + n + i] = EI(INFINITY, (size_t)-1); }
      for (int i = 0; i < n; ++i) buf[64 + i] = EI(e[i], (size_t)i);
      std::sort(buf.begin() + 64, buf.begin() + 64 + n);
      std::vector<int> gi(n);
      for (int i = 0; i < n; ++i) gi[i] = i;
      const int la = libstdcxx_sort(e.data(), gi.data(), n);
      ++cases;
      if (la) { ++left; continue; }
      bool same = true;
      for (int i = 0; i < n && same; ++i) same = (int)buf[64 + i].second == gi[i];
      if (!same) { ++differ; std::printf("std::sort restatement differs on killer sequence %d\n", t); }
    }
    std::printf("std::sort restated: %ld arrays (%ld with NaNs), %ld differ, %ld left the array (not compared), heap-sort branch taken %ld times\n",
                cases, with_nan, differ, left, mvgcuda::geo::g_ss_heap_sorts);
    if (mvgcuda::geo::g_ss_heap_sorts == 0) { std::printf("the depth-limit branch was never exercised\n"); ++bad; }
    if (differ) ++bad;
  }
  std::printf(bad ? "ACRANSAC CORE FAILED\n" : "ACRANSAC CORE OK\n");
  return bad ? 1 : 0;
}
