// CPU check of 3dreconstruction_b200/csrc/rbtree_dedup.cuh (the product's device code, compiled for the host here) against the
// real container the reference uses: std::set range-constructed with the reference's FULL ordering predicate over
// (x1, y1, x2, y2) (indexed_match_decorator.h:33-53, 90-104), libstdc++.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "../../3dreconstruction_b200/csrc/rbtree_dedup.cuh"

struct Elem { float x1, y1, x2, y2; int pos; };
struct Less {  // the reference's predicate, expression by expression
  bool operator()(const Elem& a, const Elem& b) const {
    if (a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2) return false;
    if (a.x1 < b.x1) return a.y1 < b.y1;
    else if (a.x1 > b.x1) return a.y1 < b.y1;
    return a.x1 < b.x1;
  }
};

int main(int argc, char** argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 20000;
  std::mt19937 rng(12345);
  long long total = 0, kept = 0;
  for (int c = 0; c < cases; ++c) {
    const int n = (c % 50 == 0) ? (int)(rng() % 3000) : (int)(rng() % 120);
    const int alpha = 2 + (int)(rng() % (c % 3 == 0 ? 6 : 400));  // small alphabets: many equal x / y / (x, y)
    std::vector<Elem> v(n);
    for (int i = 0; i < n; ++i) {
      Elem e;
      e.x1 = (float)(rng() % alpha) * 0.5f; e.y1 = (float)(rng() % alpha) * 0.25f;
      e.x2 = (float)(rng() % alpha); e.y2 = (float)(rng() % alpha);
      if (i > 0 && rng() % 9 == 0) { const int src = rng() % i; e = v[src]; }  // exact duplicates
      if (c % 97 == 0 && rng() % 40 == 0) e.x1 = std::nanf("");               // unordered keys
      e.pos = i;
      v[i] = e;
    }
    std::set<Elem, Less> s(v.begin(), v.end());
    std::vector<mvgcuda::RbNode> nd(n + 1);
    std::vector<int> out(n + 1);
    for (int i = 0; i < n; ++i) { nd[i].x = v[i].x1; nd[i].y = v[i].y1; }
    const int m = mvgcuda::rbtree_dedup(nd.data(), n, out.data());
    {  // the 16-bit node form (what the shared-memory kernel runs) must agree with the 32-bit one
      std::vector<mvgcuda::RbNode16> nd16(n + 1);
      std::vector<unsigned short> out16(n + 1);
      for (int i = 0; i < n; ++i) { nd16[i].x = v[i].x1; nd16[i].y = v[i].y1; }
      const int m16 = mvgcuda::rbtree_dedup(nd16.data(), n, out16.data());
      if (m16 != m) { printf("case %d: 16-bit nodes keep %d, 32-bit %d\n", c, m16, m); return 1; }
      for (int i = 0; i < m; ++i) if (out16[i] != out[i]) { printf("case %d: 16-bit nodes differ at %d\n", c, i); return 1; }
    }
    if (m != (int)s.size()) { printf("case %d: size %d != %zu\n", c, m, s.size()); return 1; }
    int q = 0;
    for (const Elem& e : s) {
      if (out[q] != e.pos) { printf("case %d: element %d is %d, std::set has %d\n", c, q, out[q], e.pos); return 1; }
      ++q;
    }
    total += n; kept += m;
  }
  printf("rbtree_dedup == std::set on %d cases, %lld elements, %lld kept\n", cases, total, kept);
  return 0;
}
