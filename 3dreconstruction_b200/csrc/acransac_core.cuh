// acransac_core.cuh -- scalar building blocks of the AC-RANSAC fundamental-matrix filter, written so that
// the SAME source compiles for the host (g++, tests/native/test_acransac_core.cpp, where it is compared bit for bit with
// the reference's own classes) and for the device (nvcc -fmad=false: no FMA contraction, so IEEE double +,-,*,/,sqrt give
// the bits the reference's x86-64 SSE2 build gives).
//
//   glibc rand()                         random_sampling.h:46-58 draws from the C library's default generator (TYPE_3,
//                                        additive feedback x[i] = x[i-3] + x[i-31]); never seeded by the reference == srand(1)
//   RandomSample                         random_sampling.h:46-58
//   NormalizePoints                      conditioning.cpp:46-56,58-69 (image-size preconditioner)
//   SevenPointSolver::Solve              solver_fundamental_kernel.cpp:11-66: 9x9 epipolar system -> Nullspace2 (numeric.h:248-262,
//                                        Eigen 3.2.2 JacobiSVD<Matrix<double,9,9>>, two-sided Jacobi, JacobiSVD.h:823-916 + Jacobi.h)
//                                        -> det(F1 + a F2) = 0 -> SolveCubicPolynomial (poly.h:24-115)
//   EpipolarDistanceError::Error         solver_fundamental_kernel.h:98-108 (expression order of Eigen's fixed-size products)
//   bestNFA term                         estimator_acransac.h:73-96
//
// Transcendentals: the cubic solver calls acos / cos / pow; CUDA's and glibc's implementations differ in the last bit for
// ~10 % of the roots, and the ORDER of the reference's output depends on rounding noise (the seven sampled points have
// residuals of ~1e-32 and lead the sorted inlier list), so a model that is ACCEPTED must have the bits the reference's C
// library gives.  The GPU evaluates every iteration with CUDA's functions and has the few models that matter (the ones an
// ACRANSAC decision hinges on) recomputed with roots from the host's libm (solve_cubic on the host, acransac_kernels.cuh).
// log10 inside the NFA only enters comparisons between different models and stays on the device.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MVG_GEO_HD __host__ __device__ inline
#else
#define MVG_GEO_HD inline
#endif

namespace mvgcuda {
namespace geo {

// ------------------------------------------------------------------------------------------ glibc rand()
// random_r.c, TYPE_3 (degree 31, separation 3): after srandom(seed) the state is r[0..30] from the Lehmer generator
// (16807 mod 2^31-1, Schrage), the first 310 outputs are discarded, and every output is  (x[i] = x[i-31] + x[i-3]) >> 1.
struct GlibcRand {
  uint32_t x[31];  // the state table of random_r.c
  int f, r;        // front / rear index (fptr = &state[3], rptr = &state[0] after srandom)
};

MVG_GEO_HD uint32_t glibc_rand(GlibcRand& g) {  // random_r: *fptr += *rptr; result = *fptr >> 1; advance both
  const uint32_t v = (g.x[g.f] += g.x[g.r]);
  g.f = g.f == 30 ? 0 : g.f + 1;
  g.r = g.r == 30 ? 0 : g.r + 1;
  return v >> 1;
}

MVG_GEO_HD void glibc_srand(GlibcRand& g, unsigned seed) {  // srandom_r
  int32_t word = seed == 0 ? 1 : (int32_t)seed;
  g.x[0] = (uint32_t)word;
  for (int i = 1; i < 31; ++i) {
    const long hi = word / 127773, lo = word % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    word = (int32_t)w;
    g.x[i] = (uint32_t)word;
  }
  g.f = 3;
  g.r = 0;
  for (int i = 0; i < 310; ++i) (void)glibc_rand(g);
}

// ------------------------------------------------------------------------------------------ RandomSample
// random_sampling.h:46-58: X sorted distinct values of [0, n) from X consecutive rand() values.
template <int X>
MVG_GEO_HD void random_sample(const uint32_t* r, int n, int* s) {
  for (int i = 0; i < X; ++i) {
    int v = (int)((r[i] >> 3) % (uint32_t)(n - i));
    int j = 0;
    for (; j < i && v >= s[j]; ++j) ++v;
    for (int k = i; k > j; --k) s[k] = s[k - 1];
    s[j] = v;
  }
}

// ------------------------------------------------------------------------------------------ normalisation
// PreconditionerFromPoints(width, height, T) (conditioning.cpp:46-56): T = [d 0 tx; 0 d ty; 0 0 1] with
// d = 1 / sqrt(double(width * height)), tx = double(-.5f * width) * d (a FLOAT product in the reference), ty = -.5 * height * d.
struct Normalizer { double d, tx, ty; };
MVG_GEO_HD Normalizer make_normalizer(int width, int height) {
  Normalizer N;
  N.d = 1.0 / sqrt(static_cast<double>(width * height));
  N.tx = static_cast<double>(-.5f * static_cast<float>(width)) * N.d;
  N.ty = -.5 * height * N.d;
  return N;
}
// applyTransformationToPoints (conditioning.cpp:58-69): out = T * (x, y, 1), then out / out(2); out(2) == 1 exactly and the
// zero entries of T contribute exact zeros, so  x' = (d * x + 0 * y) + tx  in Eigen's left-to-right accumulation.
MVG_GEO_HD void normalize_point(const Normalizer& N, float x, float y, double& ox, double& oy) {
  const double xd = static_cast<double>(x), yd = static_cast<double>(y);
  ox = ((N.d * xd + 0.0 * yd) + N.tx * 1.0) / 1.0;
  oy = ((0.0 * xd + N.d * yd) + N.ty * 1.0) / 1.0;
}

// ------------------------------------------------------------------------------------------ 9x9 Jacobi SVD (V only)
// Eigen 3.2.2 JacobiSVD<Matrix<double,9,9>>(A, ComputeFullV).  W and V are 9x9 COLUMN-major (element (r, c) at [r + 9 c]).
// On return V holds matrixV() with the columns sorted by descending singular value (step 4 of JacobiSVD::compute).
MVG_GEO_HD double eig_hypot(double x, double y) {  // MathFunctions.h:285-302
  const double ax = fabs(x), ay = fabs(y);
  const double p = ax < ay ? ay : ax;  // std::max(_x, _y): _x unless _x < _y
  if (p == 0.0) return 0.0;
  const double q = ay < ax ? ay : ax;  // std::min(_x, _y): _x unless _y < _x
  const double qp = q / p;
  return p * sqrt(1.0 + qp * qp);
}

struct Rot { double c, s; };

// real_2x2_jacobi_svd (JacobiSVD.h:413-441) on the 2x2 block (p, q) of W
MVG_GEO_HD void real_2x2_jacobi_svd(double m00, double m01, double m10, double m11, Rot& j_left, Rot& j_right) {
  Rot rot1;
  const double t = m00 + m11;
  const double d = m10 - m01;
  if (t == 0.0) {
    rot1.c = 0.0;
    rot1.s = d > 0.0 ? 1.0 : -1.0;
  } else {
    const double t2d2 = eig_hypot(t, d);
    rot1.c = fabs(t) / t2d2;
    rot1.s = d / t2d2;
    if (t < 0.0) rot1.s = -rot1.s;
  }
  // m.applyOnTheLeft(0, 1, rot1): skipped entirely when c == 1 && s == 0 (Jacobi.h:302-303)
  if (!(rot1.c == 1.0 && rot1.s == 0.0)) {
    const double x0 = m00, x1 = m01, y0 = m10, y1 = m11;
    m00 = rot1.c * x0 + rot1.s * y0;  m10 = -rot1.s * x0 + rot1.c * y0;
    m01 = rot1.c * x1 + rot1.s * y1;  m11 = -rot1.s * x1 + rot1.c * y1;
  }
  // j_right->makeJacobi(m, 0, 1) == makeJacobi(m00, m01, m11) (Jacobi.h:82-113)
  if (m01 == 0.0) {
    j_right.c = 1.0;
    j_right.s = 0.0;
  } else {
    const double tau = (m00 - m11) / (2.0 * fabs(m01));
    const double w = sqrt(tau * tau + 1.0);
    const double tt = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    const double sign_t = tt > 0.0 ? 1.0 : -1.0;
    const double n = 1.0 / sqrt(tt * tt + 1.0);
    j_right.s = -sign_t * (m01 / fabs(m01)) * fabs(tt) * n;
    j_right.c = n;
  }
  // *j_left = rot1 * j_right->transpose()   (Jacobi.h:51-56,59: transpose = (c, -s))
  const double oc = j_right.c, os = -j_right.s;
  j_left.c = rot1.c * oc - rot1.s * os;
  j_left.s = rot1.c * os + rot1.s * oc;
}

// apply_rotation_in_the_plane(x, y, j) (Jacobi.h:294-427): x' = c x + s y, y' = -s x + c y; no-op for the identity
MVG_GEO_HD void rotate_pair(double& x, double& y, double c, double s) {
  const double xi = x, yi = y;
  x = c * xi + s * yi;
  y = -s * xi + c * yi;
}

// jacobi_svd9_sweeps: steps 2-4 of JacobiSVD::compute on a 9x9 work matrix W with V already initialised (identity for a
// square input, the column permutation of the QR preconditioner for a tall one).
MVG_GEO_HD void jacobi_svd9_sweeps(double* W, double* V);
MVG_GEO_HD void jacobi_svd9_v(double* W, double* V) {
  for (int i = 0; i < 81; ++i) V[i] = 0.0;
  for (int i = 0; i < 9; ++i) V[i + 9 * i] = 1.0;
  jacobi_svd9_sweeps(W, V);
}
MVG_GEO_HD void jacobi_svd9_sweeps(double* W, double* V) {
  const double precision = 2.0 * DBL_EPSILON;
  const double consider_as_zero = 2.0 * 4.9406564584124654e-324;  // 2 * denorm_min
  double scale = 0.0;
  for (int i = 0; i < 81; ++i) { const double a = fabs(W[i]); if (a > scale) scale = a; }
  if (scale == 0.0) scale = 1.0;
  const double inv = 1.0 / scale;  // operator/= multiplies by 1/scale (SelfCwiseBinaryOp.h:181-193)
  for (int i = 0; i < 81; ++i) W[i] *= inv;
  bool finished = false;
  while (!finished) {
    finished = true;
    for (int p = 1; p < 9; ++p) {
      for (int q = 0; q < p; ++q) {
        const double app = fabs(W[p + 9 * p]), aqq = fabs(W[q + 9 * q]);
        const double mx = app < aqq ? aqq : app;             // std::max(a, b): a unless a < b
        const double pm = precision * mx;
        const double threshold = consider_as_zero < pm ? pm : consider_as_zero;
        const double apq = fabs(W[p + 9 * q]), aqp = fabs(W[q + 9 * p]);
        const double off = apq < aqp ? aqp : apq;
        if (off > threshold) {
          finished = false;
          Rot jl, jr;
          real_2x2_jacobi_svd(W[p + 9 * p], W[p + 9 * q], W[q + 9 * p], W[q + 9 * q], jl, jr);
          if (!(jl.c == 1.0 && jl.s == 0.0))  // workMatrix.applyOnTheLeft(p, q, j_left): rows p, q
            for (int i = 0; i < 9; ++i) rotate_pair(W[p + 9 * i], W[q + 9 * i], jl.c, jl.s);
          // applyOnTheRight(p, q, j_right) == rotation j_right.transpose() = (c, -s) on columns p, q
          if (!(jr.c == 1.0 && -jr.s == 0.0)) {
            for (int i = 0; i < 9; ++i) rotate_pair(W[i + 9 * p], W[i + 9 * q], jr.c, -jr.s);
            for (int i = 0; i < 9; ++i) rotate_pair(V[i + 9 * p], V[i + 9 * q], jr.c, -jr.s);
          }
        }
      }
    }
  }
  // steps 3-4: singular values = |diagonal|, selection sort in descending order (first maximum wins), stop at an exact zero
  double sv[9];
  for (int i = 0; i < 9; ++i) sv[i] = fabs(W[i + 9 * i]);
  for (int i = 0; i < 9; ++i) {
    int pos = 0;
    double best = sv[i];
    for (int k = 1; k < 9 - i; ++k) if (sv[i + k] > best) { best = sv[i + k]; pos = k; }
    if (best == 0.0) break;
    if (pos) {
      pos += i;
      const double t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
      for (int r = 0; r < 9; ++r) { const double u = V[r + 9 * pos]; V[r + 9 * pos] = V[r + 9 * i]; V[r + 9 * i] = u; }
    }
  }
}

// ------------------------------------------------------------------------------------------ cubic (poly.h:24-115)
MVG_GEO_HD int solve_cubic(const double* P /* ascending powers */, double* roots) {
  if (P[0] == 0.0) return 0;
  const double a = P[2] / P[3], b = P[1] / P[3], c = P[0] / P[3];
  const double q = a * a - 3 * b;
  const double r = 2 * a * a * a - 9 * a * b + 27 * c;
  const double Q = q / 9, R = r / 54;
  const double Q3 = Q * Q * Q, R2 = R * R;
  const double CR2 = 729 * r * r, CQ3 = 2916 * q * q * q;
  if (R == 0 && Q == 0) {
    roots[0] = roots[1] = roots[2] = -a / 3;
    return 3;
  } else if (CR2 == CQ3) {
    const double sqrtQ = sqrt(Q);
    if (R > 0) { roots[0] = -2 * sqrtQ - a / 3; roots[1] = sqrtQ - a / 3; roots[2] = sqrtQ - a / 3; }
    else       { roots[0] = -sqrtQ - a / 3; roots[1] = -sqrtQ - a / 3; roots[2] = 2 * sqrtQ - a / 3; }
    return 3;
  } else if (CR2 < CQ3) {
    const double sqrtQ = sqrt(Q);
    const double sqrtQ3 = sqrtQ * sqrtQ * sqrtQ;
    const double theta = acos(R / sqrtQ3);
    const double norm = -2 * sqrtQ;
    double x0 = norm * cos(theta / 3) - a / 3;
    double x1 = norm * cos((theta + 2.0 * 3.14159265358979323846) / 3) - a / 3;
    double x2 = norm * cos((theta - 2.0 * 3.14159265358979323846) / 3) - a / 3;
    if (x0 > x1) { const double t = x0; x0 = x1; x1 = t; }
    if (x1 > x2) {
      const double t = x1; x1 = x2; x2 = t;
      if (x0 > x1) { const double u = x0; x0 = x1; x1 = u; }
    }
    roots[0] = x0; roots[1] = x1; roots[2] = x2;
    return 3;
  }
  const double sgnR = (R >= 0 ? 1 : -1);
  const double A = -sgnR * pow(fabs(R) + sqrt(R2 - Q3), 1.0 / 3.0);
  const double B = Q / A;
  roots[0] = A + B - a / 3;
  return 1;
}

// ------------------------------------------------------------------------------------------ 7-point solver
// x1, x2: the 7 sampled (normalised) correspondences, [7][2].  W, V: 81 doubles of scratch each.
// det(F1 + a F2) as a cubic in a, ascending powers (solver_fundamental_kernel.cpp:37-59; F1 = Map<RMat3>(f1): F1(r, c) = f1[3 r + c]).
MVG_GEO_HD void cubic_from_null_vectors(const double* f1, const double* f2, double* P) {
  const double a = f1[0], j = f2[0], b = f1[1], k = f2[1], c = f1[2], l = f2[2], d = f1[3], m = f2[3], e = f1[4], n = f2[4],
               f = f1[5], o = f2[5], g = f1[6], p = f2[6], h = f1[7], q = f2[7], i = f1[8], r = f2[8];
  P[0] = a * e * i + b * f * g + c * d * h - a * f * h - b * d * i - c * e * g;
  P[1] = a * e * r + a * i * n + b * f * p + b * g * o + c * d * q + c * h * m + d * h * l + e * i * j + f * g * k -
         a * f * q - a * h * o - b * d * r - b * i * m - c * e * p - c * g * n - d * i * k - e * g * l - f * h * j;
  P[2] = a * n * r + b * o * p + c * m * q + d * l * q + e * j * r + f * k * p + g * k * o + h * l * m + i * j * n -
         a * o * q - b * m * r - c * n * p - d * k * r - e * l * p - f * j * q - g * l * n - h * j * o - i * k * m;
  P[3] = j * n * r + k * o * p + l * m * q - j * o * q - k * m * r - l * n * p;
}

// Step 1 (pure IEEE arithmetic, bit-identical on every machine): the two null vectors f1 = V.col(8), f2 = V.col(7)
// (numeric.h:252-253) and the coefficients P of det(F1 + a F2), ascending powers.
MVG_GEO_HD void seven_point_basis(const double* x1, const double* x2, double* W, double* V, double* P) {
  for (int i = 0; i < 81; ++i) W[i] = 0.0;
  for (int i = 0; i < 7; ++i) {  // EncodeEpipolarEquation, solver_fundamental_kernel.h:56-69 (rows 7, 8 stay zero)
    const double u1 = x1[2 * i], v1 = x1[2 * i + 1], u2 = x2[2 * i], v2 = x2[2 * i + 1];
    W[i + 9 * 0] = u2 * u1; W[i + 9 * 1] = u2 * v1; W[i + 9 * 2] = u2;
    W[i + 9 * 3] = v2 * u1; W[i + 9 * 4] = v2 * v1; W[i + 9 * 5] = v2;
    W[i + 9 * 6] = u1;      W[i + 9 * 7] = v1;      W[i + 9 * 8] = 1.0;
  }
  jacobi_svd9_v(W, V);
  cubic_from_null_vectors(V + 9 * 8, V + 9 * 7, P);
}
// Step 2: the models F1 + root * F2 (solver_fundamental_kernel.cpp:62-64); F(r, c) at [3 r + c].
MVG_GEO_HD void models_from_roots(const double* f1, const double* f2, const double* roots, int nr, double* F) {
  for (int kk = 0; kk < nr; ++kk)
    for (int t = 0; t < 9; ++t) F[9 * kk + t] = f1[t] + roots[kk] * f2[t];
}
// Both steps with the cubic solved in place (acos / cos / pow of whichever C library the caller links).
MVG_GEO_HD int seven_point_models(const double* x1, const double* x2, double* W, double* V, double* F) {
  double P[4], roots[3];
  seven_point_basis(x1, x2, W, V, P);
  const int nr = solve_cubic(P, roots);
  models_from_roots(V + 9 * 8, V + 9 * 7, roots, nr, F);
  return nr;
}

// ------------------------------------------------------------------------------------------ 4-point homography solver
// homography::FourPointSolver::Solve (solver_homography_kernel.cpp:34-58): 16x9 action matrix (two rows per point, rows
// 8..15 zero) -> Nullspace -> JacobiSVD<Matrix<double,16,9>>(L, ComputeFullV).  More rows than columns: Eigen first runs
// its QR preconditioner, ColPivHouseholderQR (ColPivHouseholderQR.h compute(), Householder.h), then the two-sided Jacobi
// on R with V initialised to the column permutation.  The sums inside are SSE2-vectorised in the reference, so their
// association is fixed by Eigen's code and the 16-byte alignment of the operands (element (r, c) of the 16x9 column-major
// matrix is aligned iff r is even):
//   fixed 16-vector squaredNorm   Redux.h redux_vec_unroller: binary tree over the 8 packets, then lane 0 + lane 1
//   dynamic tail squaredNorm      Redux.h:202-253 (LinearVectorizedTraversal, NoUnrolling) on the abs2 EXPRESSION, which has no
//                                 direct access: packets start at element 0 whatever the address (unaligned loads), two
//                                 packet accumulators, lane 0 + lane 1, then the odd tail scalar
//   essential^T * bottom          coefficient-based product (all dimensions "small"): a plain sequential dot per column
struct P2 { double a, b; };
MVG_GEO_HD P2 p2add(P2 x, P2 y) { return P2{x.a + y.a, x.b + y.b}; }
MVG_GEO_HD P2 p2sq(const double* v) { return P2{v[0] * v[0], v[1] * v[1]}; }

// squaredNorm of the dynamic block x[0 .. size) whose element 0 sits at an address of parity `odd` (1: misaligned)
MVG_GEO_HD double eig_sqnorm_dyn(const double* x, int size, int odd) {
  const int aligned_start = odd < size ? odd : size;
  const int aligned_size2 = ((size - aligned_start) / 4) * 4;
  const int aligned_size = ((size - aligned_start) / 2) * 2;
  const int aligned_end2 = aligned_start + aligned_size2, aligned_end = aligned_start + aligned_size;
  double res;
  if (aligned_size) {
    P2 r0 = p2sq(x + aligned_start);
    if (aligned_size > 2) {
      P2 r1 = p2sq(x + aligned_start + 2);
      for (int i = aligned_start + 4; i < aligned_end2; i += 4) { r0 = p2add(r0, p2sq(x + i)); r1 = p2add(r1, p2sq(x + i + 2)); }
      r0 = p2add(r0, r1);
      if (aligned_end > aligned_end2) r0 = p2add(r0, p2sq(x + aligned_end2));
    }
    res = r0.a + r0.b;
    for (int i = 0; i < aligned_start; ++i) res = res + x[i] * x[i];
    for (int i = aligned_end; i < size; ++i) res = res + x[i] * x[i];
  } else {
    res = x[0] * x[0];
    for (int i = 1; i < size; ++i) res = res + x[i] * x[i];
  }
  return res;
}
// essential^T * bottom: with a 16x9 maximum size Eigen's product selector classifies all three dimensions as "small"
// (GeneralProduct.h product_size_category: MaxSize is not Dynamic) and evaluates the product coefficient by coefficient,
// the inner dimension being dynamic without vectorisation: res = l0 r0; res += l_i r_i in order (CoeffBasedProduct.h).
MVG_GEO_HD double eig_coeff_dot(const double* lhs, const double* rhs, int depth) {
  double res = lhs[0] * rhs[0];
  for (int i = 1; i < depth; ++i) res += lhs[i] * rhs[i];
  return res;
}

// x1, x2: the 4 sampled (normalised) correspondences, [4][2].  A: 144 doubles of scratch (16x9), W, V: 81 each.
// H: the model, H(r, c) at [3 r + c] (Map<RMat3> of the null vector V.col(8)).
MVG_GEO_HD void four_point_qr(const double* x1, const double* x2, double* A, double* W, double* V);
MVG_GEO_HD void four_point_model(const double* x1, const double* x2, double* A, double* W, double* V, double* H) {
  four_point_qr(x1, x2, A, W, V);
  jacobi_svd9_sweeps(W, V);
  for (int t = 0; t < 9; ++t) H[t] = V[t + 9 * 8];
}
// the action matrix and its QR preconditioner: leaves the Jacobi work matrix in W and the column permutation in V
MVG_GEO_HD void four_point_qr(const double* x1, const double* x2, double* A, double* W, double* V) {
  for (int i = 0; i < 144; ++i) A[i] = 0.0;
  for (int i = 0; i < 4; ++i) {  // BuildActionMatrix, solver_homography_kernel.cpp:12-32
    const double xx = x1[2 * i], xy = x1[2 * i + 1], yx = x2[2 * i], yy = x2[2 * i + 1];
    int j = 2 * i;
    A[j + 16 * 0] = xx; A[j + 16 * 1] = xy; A[j + 16 * 2] = 1.0;
    A[j + 16 * 6] = -yx * xx; A[j + 16 * 7] = -yx * xy; A[j + 16 * 8] = -yx;
    ++j;
    A[j + 16 * 3] = xx; A[j + 16 * 4] = xy; A[j + 16 * 5] = 1.0;
    A[j + 16 * 6] = -yy * xx; A[j + 16 * 7] = -yy * xy; A[j + 16 * 8] = -yy;
  }
  // ---- ColPivHouseholderQR<Matrix<double,16,9>>::compute
  const int rows = 16, cols = 9;
  double col_sq[9], temp[9];
  int transp[9], perm[9];
  for (int k = 0; k < cols; ++k) {  // fixed-size aligned column: tree over the 8 packets
    const double* c = A + 16 * k;
    const P2 t = p2add(p2add(p2add(p2sq(c), p2sq(c + 2)), p2add(p2sq(c + 4), p2sq(c + 6))),
                       p2add(p2add(p2sq(c + 8), p2sq(c + 10)), p2add(p2sq(c + 12), p2sq(c + 14))));
    col_sq[k] = t.a + t.b;
  }
  double mx = col_sq[0];
  for (int k = 1; k < cols; ++k) if (col_sq[k] > mx) mx = col_sq[k];
  const double threshold_helper = mx * (DBL_EPSILON * DBL_EPSILON) / static_cast<double>(rows);
  int nonzero_pivots = cols;
  for (int k = 0; k < cols; ++k) {
    int big = 0;
    double big_v = col_sq[k];
    for (int j = 1; j < cols - k; ++j) if (col_sq[k + j] > big_v) { big_v = col_sq[k + j]; big = j; }
    big += k;
    big_v = eig_sqnorm_dyn(A + k + 16 * big, rows - k, 0);
    col_sq[big] = big_v;
    if (big_v < threshold_helper * static_cast<double>(rows - k)) {
      nonzero_pivots = k;
      for (int c = k; c < cols; ++c)  // bottomRightCorner(rows-k, cols-k).triangularView<StrictlyLower>().setZero()
        for (int r = c + 1; r < rows; ++r) A[r + 16 * c] = 0.0;
      break;
    }
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < rows; ++r) { const double t = A[r + 16 * k]; A[r + 16 * k] = A[r + 16 * big]; A[r + 16 * big] = t; }
      const double t = col_sq[k]; col_sq[k] = col_sq[big]; col_sq[big] = t;
    }
    // makeHouseholderInPlace on col(k).tail(rows - k)
    double* ck = A + k + 16 * k;
    const int n = rows - k;
    const double tail_sq = n == 1 ? 0.0 : eig_sqnorm_dyn(ck + 1, n - 1, 0);
    const double c0 = ck[0];
    double tau, beta;
    if (tail_sq == 0.0) {
      tau = 0.0; beta = c0;
      for (int r = 1; r < n; ++r) ck[r] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.0) beta = -beta;
      const double den = c0 - beta;
      for (int r = 1; r < n; ++r) ck[r] = ck[r] / den;
      tau = (beta - c0) / beta;
    }
    ck[0] = beta;
    // applyHouseholderOnTheLeft on bottomRightCorner(rows - k, cols - k - 1) with essential = col(k).tail(rows - k - 1)
    const int nc = cols - k - 1;
    if (nc > 0) {
      const double* ess = ck + 1;
      for (int j = 0; j < nc; ++j) temp[j] = eig_coeff_dot(ess, A + (k + 1) + 16 * (k + 1 + j), n - 1);
      for (int j = 0; j < nc; ++j) temp[j] += A[k + 16 * (k + 1 + j)];
      for (int j = 0; j < nc; ++j) A[k + 16 * (k + 1 + j)] -= temp[j] * tau;
      for (int j = 0; j < nc; ++j)
        for (int r = 0; r < n - 1; ++r) A[(k + 1 + r) + 16 * (k + 1 + j)] -= temp[j] * (ess[r] * tau);
    }
    for (int j = 0; j < nc; ++j) col_sq[k + 1 + j] -= A[k + 16 * (k + 1 + j)] * A[k + 16 * (k + 1 + j)];
  }
  for (int i = 0; i < cols; ++i) perm[i] = i;
  for (int k = 0; k < nonzero_pivots; ++k) { const int t = perm[k]; perm[k] = perm[transp[k]]; perm[transp[k]] = t; }
  // ---- JacobiSVD: work matrix = upper triangle of R, V = the permutation (V(perm[i], i) = 1)
  for (int c = 0; c < 9; ++c)
    for (int r = 0; r < 9; ++r) W[r + 9 * c] = r <= c ? A[r + 16 * c] : 0.0;
  for (int i = 0; i < 81; ++i) V[i] = 0.0;
  for (int i = 0; i < 9; ++i) V[perm[i] + 9 * i] = 1.0;
}

// homography::AsymmetricError::Error (solver_homography_kernel.h:32-38): || x2 - dehomogenise(H (x1, 1)) ||^2
MVG_GEO_HD double homography_error(const double* H, double x1, double y1, double x2, double y2) {
  const double hx = (H[0] * x1 + H[1] * y1) + H[2] * 1.0;
  const double hy = (H[3] * x1 + H[4] * y1) + H[5] * 1.0;
  const double hz = (H[6] * x1 + H[7] * y1) + H[8] * 1.0;
  const double dx = x2 - hx / hz, dy = y2 - hy / hz;
  return dx * dx + dy * dy;
}

// ------------------------------------------------------------------------------------------ residual
// EpipolarDistanceError::Error (solver_fundamental_kernel.h:98-108): F_x = F * (x1, 1); Square(F_x . (x2, 1)) / |F_x.head<2>|^2.
// Eigen's fixed-size evaluation order: matrix-vector rows accumulate left to right, the 3-term dot is e0 + (e1 + e2).
MVG_GEO_HD double epipolar_error(const double* F, double x1, double y1, double x2, double y2) {
  const double fx0 = (F[0] * x1 + F[1] * y1) + F[2] * 1.0;
  const double fx1 = (F[3] * x1 + F[4] * y1) + F[5] * 1.0;
  const double fx2 = (F[6] * x1 + F[7] * y1) + F[8] * 1.0;
  const double dot = fx0 * x2 + (fx1 * y2 + fx2 * 1.0);
  return (dot * dot) / (fx0 * fx0 + fx1 * fx1);
}

// One term of bestNFA (estimator_acransac.h:86-91); mult_error = 0.5 for point-to-line residuals (F), 1.0 for
// point-to-point ones (H) (estimator_acransac_kernel_adaptator.h:92):
//   logalpha = logalpha0 + mult_error * log10(e + FLT_MIN);  NFA = loge0 + logalpha * (k - sample_size) + logc_n[k] + logc_k[k]
MVG_GEO_HD double nfa_term(double logalpha0, double loge0, double mult_error, double err, int k, int sample_size, float logc_n_k, float logc_k_k) {
  const double logalpha = logalpha0 + mult_error * log10(err + static_cast<double>(FLT_MIN));
  return loge0 + logalpha * static_cast<double>(k - sample_size) + logc_n_k + logc_k_k;
}

}  // namespace geo
}  // namespace mvgcuda
