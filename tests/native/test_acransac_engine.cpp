// Host-side test of the speculative ACRANSAC engine (csrc/acransac_engine.cuh: ranges of iterations evaluated with the
// current sampling set, then accounted for in order) against the reference's own sequential ACRANSAC
// (oracle/_ref/libmvgref_geom.so): inlier lists IN ORDER, minNFA and errorMax bit for bit, and the number of rand() values
// consumed, on planted-geometry pairs (meaningful model found early), pure-noise pairs (no model: the reserve logic),
// tiny pairs and threshold-free runs.  TEST INFRASTRUCTURE.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <random>
#include <vector>

#include "../../3dreconstruction_b200/csrc/acransac_engine.cuh"
#include "../../3dreconstruction_b200/csrc/stdsort_restated.cuh"

extern "C" {
int ref_acransac_f(const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, double precision, int iterations,
                   unsigned seed, int* inliers, double* out);
int ref_acransac_h(const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, double precision, int iterations,
                   unsigned seed, int* inliers, double* out);
int ref_acransac_at(char model, const float* xI, const float* xJ, int n, int wI, int hI, int wJ, int hJ, double precision, int iterations,
                    unsigned seed, long skip, int verbose, int* inliers, double* out);
void ref_logc(int n, float* logc_n, float* logc_k);
}
using namespace mvgcuda::geo;

struct Iter { double nfa; int n_inl, model, n_models; double F[27]; };
struct HostRange {
  const std::vector<Iter>& r;
  int first_below(int lo, int hi, double thr) const { for (int i = lo; i < hi; ++i) if (r[i].nfa < thr) return i; return -1; }
  int argmin_first(int lo, int hi) const { int a = lo; for (int i = lo + 1; i < hi; ++i) if (r[i].nfa < r[a].nfa) a = i; return a; }
  double nfa(int i) const { return r[i].nfa; }
  int n_inl(int i) const { return r[i].n_inl; }
  int model(int i) const { return r[i].model; }
};

// The candidates of one model as the product builds them (acransac_kernels.cuh: model_candidates_warp): residuals <=
// max_threshold sorted by (residual, index); with a NaN among the residuals, std::sort of ALL of them restated step by step
// (stdsort_restated.cuh) and the length bestNFA's scan reaches.
static long g_nan_models = 0;
static int candidates(const PairGeo& P, const double* F, std::vector<Cand>& list) {
  list.clear();
  std::vector<double> all(P.n);
  bool any_nan = false;
  for (int i = 0; i < P.n; ++i) {
    const double e = P.sample == kSampleH ? homography_error(F, P.x1[2 * i], P.x1[2 * i + 1], P.x2[2 * i], P.x2[2 * i + 1])
                                          : epipolar_error(F, P.x1[2 * i], P.x1[2 * i + 1], P.x2[2 * i], P.x2[2 * i + 1]);
    all[i] = e;
    any_nan = any_nan || e != e;
    if (e <= P.max_threshold) list.push_back(Cand{e, i});
  }
  if (any_nan) {
    ++g_nan_models;
    std::vector<int> idx(P.n);
    for (int i = 0; i < P.n; ++i) idx[i] = i;
    libstdcxx_sort(all.data(), idx.data(), P.n);
    const int m = nfa_scan_end(all.data(), P.n, P.sample, P.max_threshold);
    list.resize(m);
    for (int i = 0; i < m; ++i) list[i] = Cand{all[i], idx[i]};
    return m;
  }
  std::sort(list.begin(), list.end(), cand_less);
  return (int)list.size();
}

struct Result { std::vector<int> inliers; double nfa, err_max; long used; };

static Result run_engine(char model, const std::vector<float>& xI, const std::vector<float>& xJ, int n, int wI, int hI, int wJ, int hJ,
                         double precision, int iterations, unsigned seed, long skip = 0, bool verbose = false, int first_chunk = 0) {
  const int sample = model == 'h' ? kSampleH : kSampleF;
  std::vector<double> x1(2 * n), x2(2 * n);
  const Normalizer N1 = make_normalizer(wI, hI), N2 = make_normalizer(wJ, hJ);
  for (int i = 0; i < n; ++i) {
    normalize_point(N1, xI[2 * i], xI[2 * i + 1], x1[2 * i], x1[2 * i + 1]);
    normalize_point(N2, xJ[2 * i], xJ[2 * i + 1], x2[2 * i], x2[2 * i + 1]);
  }
  std::vector<float> lcn(n + 1), lck(n + 1);
  make_logc_n(n, lcn.data());
  make_logc_k(sample, n, lck.data());
  PairGeo P;
  P.n = n; P.x1 = x1.data(); P.x2 = x2.data();
  P.max_threshold = precision == ac_inf() ? ac_inf() : precision * N2.d * N2.d;
  const double D = sqrt(wJ * (double)wJ + hJ * (double)hJ), A = wJ * (double)hJ;
  // estimator_acransac_kernel_adaptator.h:53-61: point to line 2 D / A / N2(0,0), point to point pi / A / N2(0,0)^2
  P.logalpha0 = model == 'h' ? log10(M_PI / A / (N2.d * N2.d)) : log10(2.0 * D / A / N2.d);
  P.loge0 = n > sample ? log10((double)(model == 'h' ? 1 : 3) * (size_t)(n - sample)) : 0.0;
  P.logc_n = lcn.data(); P.logc_k = lck.data();
  P.sample = sample; P.mult_error = model == 'h' ? 1.0 : 0.5;
  GlibcRand g;
  glibc_srand(g, seed);
  std::vector<uint32_t> stream((size_t)iterations * sample + 16);
  for (long i = 0; i < skip; ++i) glibc_rand(g);
  for (auto& v : stream) v = glibc_rand(g);

  AcState S;
  ac_init(S, n, iterations, sample);
  std::vector<Iter> res(iterations + 8);
  std::vector<int> vec_index(n);
  for (int i = 0; i < n; ++i) vec_index[i] = i;
  int cur_it = -1, cur_model = 0;
  std::vector<Cand> list;
  double W[81], V[81];
  while (!S.done) {
    // (first_chunk > 0: only the head of the first phase is evaluated before the first accounting, as the product does)
    const bool head = first_chunk > 0 && S.iter == 0 && S.reserve > 0 && S.extend_to == 0 && first_chunk < ac_range_end(S);
    const int hi = head ? first_chunk : ac_range_end(S);
    for (int it = S.iter; it < hi; ++it) {
      int s[7];
      if (model == 'h') random_sample<4>(&stream[(size_t)4 * it], S.n_index, s);
      else random_sample<7>(&stream[(size_t)7 * it], S.n_index, s);
      double a[14], b[14];
      for (int k = 0; k < sample; ++k) {
        const int id = vec_index[s[k]];
        a[2 * k] = x1[2 * id]; a[2 * k + 1] = x1[2 * id + 1]; b[2 * k] = x2[2 * id]; b[2 * k + 1] = x2[2 * id + 1];
      }
      Iter& R = res[it];
      if (model == 'h') { double A[144]; four_point_model(a, b, A, W, V, R.F); R.n_models = 1; }
      else R.n_models = seven_point_models(a, b, W, V, R.F);
      R.nfa = ac_inf(); R.n_inl = 0; R.model = 0;
      for (int k = 0; k < R.n_models; ++k) {
        const int m = candidates(P, R.F + 9 * k, list);
        double v; int kb;
        best_nfa_scalar(P, list.data(), m, v, kb);
        if (v < R.nfa) { R.nfa = v; R.n_inl = kb; R.model = k; }
      }
    }
    HostRange HR{res};
    ac_account(S, HR, head ? hi : -1);
    if (verbose) std::printf("  engine: accounted up to %d of %d, minNFA %.17g (it %d, %d inliers), reserve %d\n", S.iter, S.iter_num, S.min_nfa, S.best_it, S.n_inl, S.reserve);
    if (S.index_it != cur_it || S.index_model != cur_model) {
      cur_it = S.index_it; cur_model = S.index_model;
      candidates(P, res[cur_it].F + 9 * cur_model, list);
      vec_index.resize(S.n_index);
      for (int i = 0; i < S.n_index; ++i) vec_index[i] = list[i].i;
    }
  }
  Result out;
  out.nfa = S.min_nfa; out.err_max = ac_inf(); out.used = (long)sample * S.iter_num;
  if (S.min_nfa < 0.0) {  // ACRANSAC clears the inliers of a non-meaningful model (:240-241); the 2.5 x 7 floor is Fit's
    candidates(P, res[S.best_it].F + 9 * S.best_model, list);
    for (int i = 0; i < S.n_inl; ++i) out.inliers.push_back(list[i].i);
    out.err_max = sqrt(list[S.n_inl - 1].e) / N2.d;  // unormalizeError
  }
  return out;
}

// Debug aid: one pair from a file written by tests/tools/geo_h_debug.py --dump:  int32 {model, n, wI, hI, wJ, hJ, skip, iterations}, float xI[2n], xJ[2n]
static int run_file(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return 2;
  int h[8];
  if (std::fread(h, 4, 8, f) != 8) return 2;
  const int n = h[1];
  std::vector<float> xI(2 * n), xJ(2 * n);
  if (std::fread(xI.data(), 4, 2 * n, f) != (size_t)2 * n || std::fread(xJ.data(), 4, 2 * n, f) != (size_t)2 * n) return 2;
  std::fclose(f);
  std::vector<int> want(n + 1);
  double o[3];
  const int nw = ref_acransac_at((char)h[0], xI.data(), xJ.data(), n, h[2], h[3], h[4], h[5], 4.0, h[7], 1, h[6], 1, want.data(), o);
  const Result got = run_engine((char)h[0], xI, xJ, n, h[2], h[3], h[4], h[5], 4.0, h[7], 1, h[6], true);
  std::printf("reference: %d inliers, minNFA %.17g, used %.0f | engine: %zu inliers, minNFA %.17g, used %ld\n", nw, o[1], o[2], got.inliers.size(), got.nfa, got.used);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1) return run_file(argv[1]);
  std::mt19937 rng(11);
  int bad = 0, cases = 0, meaningful = 0;
  // log tables
  for (int n : {8, 9, 15, 16, 100, 361, 1000}) {
    std::vector<float> a(n + 1), b(n + 1), c(n + 1), d(n + 1);
    ref_logc(n, a.data(), b.data());
    make_logc_n(n, c.data());
    make_logc_k(7, n, d.data());
    if (std::memcmp(a.data(), c.data(), 4 * (n + 1)) || std::memcmp(b.data(), d.data(), 4 * (n + 1))) { std::printf("logc tables differ at n = %d\n", n); ++bad; }
  }
  for (int t = 0; t < 120; ++t) {
    const char model = t < 60 ? 'f' : 'h';
    const int kind = t % 6;
    int n = kind == 5 ? (model == 'h' ? 3 : 5) + t % 4 : 30 + (t * 53) % 400;
    const int wI = 4000, hI = 3000, wJ = kind == 2 ? 1416 : 4000, hJ = kind == 2 ? 1064 : 3000;
    std::uniform_real_distribution<float> ux(0.f, (float)wI), uy(0.f, (float)hI), u01(0.f, 1.f), noise(-1.5f, 1.5f);
    std::vector<float> xI(2 * n), xJ(2 * n);
    // planted geometry: points on a few depth planes seen by a translated + slightly rotated camera; a share of outliers
    const float inlier_share = kind == 1 ? 0.f : kind == 3 ? 0.25f : 0.6f;
    for (int i = 0; i < n; ++i) {
      const float x = ux(rng), y = uy(rng);
      xI[2 * i] = x; xI[2 * i + 1] = y;
      if (u01(rng) < inlier_share) {
        const float depth = model == 'h' ? 6.f + 0.0004f * x - 0.0003f * y : 4.f + 6.f * u01(rng);  // H: one slanted plane
        const float fx = 3000.f, cx = wI / 2.f, cy = hI / 2.f;
        const float X = (x - cx) / fx * depth, Y = (y - cy) / fx * depth, Z = depth;
        const float th = 0.05f, Xc = std::cos(th) * X + std::sin(th) * Z - 1.0f, Zc = -std::sin(th) * X + std::cos(th) * Z + 0.2f, Yc = Y + 0.1f;
        xJ[2 * i] = std::min((float)wJ, std::max(0.f, (fx * Xc / Zc + cx) * wJ / wI + noise(rng)));
        xJ[2 * i + 1] = std::min((float)hJ, std::max(0.f, (fx * Yc / Zc + cy) * hJ / hI + noise(rng)));
      } else {
        xJ[2 * i] = u01(rng) * wJ; xJ[2 * i + 1] = u01(rng) * hJ;
      }
    }
    if (t % 12 == 6 || t % 12 == 8)  // duplicated correspondences: identical residuals (ties go to the lower index), degenerate samples
      for (int i = 0; i < n / 5; ++i) { const int d = n - 1 - i; xI[2 * d] = xI[2 * i]; xI[2 * d + 1] = xI[2 * i + 1]; xJ[2 * d] = xJ[2 * i]; xJ[2 * d + 1] = xJ[2 * i + 1]; }
    if (t % 12 == 3 || t % 12 == 9 || t % 12 == 10)  // features clipped to the image border: collinear triples, 0 / 0 residuals under degenerate models
      for (int i = 0; i < n; ++i) {
        if (i % 5 == 0) xI[2 * i] = 0.f;
        if (i % 7 == 0) xJ[2 * i + 1] = 0.f;
        if (i % 11 == 0) { xI[2 * i] = 0.f; xI[2 * i + 1] = 0.f; }
      }
    const double precision = kind == 4 ? ac_inf() : 4.0;
    const int iterations = t % 7 == 3 ? 1024 : 4096;
    const unsigned seed = 1 + t;
    std::vector<int> want(n + 1);
    double o[3];
    const int nw = (model == 'h' ? ref_acransac_h : ref_acransac_f)(xI.data(), xJ.data(), n, wI, hI, wJ, hJ, precision, iterations, seed, want.data(), o);
    const Result got = run_engine(model, xI, xJ, n, wI, hI, wJ, hJ, precision, iterations, seed, 0, false, t % 3 == 0 ? 0 : (t % 3 == 1 ? 192 : 7));
    ++cases;
    meaningful += nw > 0;
    bool ok = nw == (int)got.inliers.size() && std::equal(got.inliers.begin(), got.inliers.end(), want.begin());
    if (n > (model == 'h' ? kSampleH : kSampleF)) ok = ok && std::memcmp(&o[1], &got.nfa, 8) == 0;  // (the reference returns (0, 0) without looking when nData <= 7)
    if (nw > 0) ok = ok && std::memcmp(&o[0], &got.err_max, 8) == 0;
    if (o[2] >= 0) ok = ok && (long)o[2] == got.used;
    if (!ok) {
      ++bad;
      std::printf("case %d model %c kind %d n %d: reference %d inliers nfa %.17g used %.0f | engine %zu inliers nfa %.17g used %ld\n", t, model, kind, n, nw, o[1], o[2],
                  got.inliers.size(), got.nfa, got.used);
    }
  }
  std::printf("%ld models had NaN residuals (std::sort restated step by step)\n", g_nan_models);
  std::printf("%d pairs (%d with a meaningful model): %s\n", cases, meaningful, bad ? "MISMATCH" : "inliers, order, minNFA, errorMax and rand() consumption identical");
  std::printf(bad ? "ACRANSAC ENGINE FAILED\n" : "ACRANSAC ENGINE OK\n");
  return bad ? 1 : 0;
}
