// CPU check of experiments/rbtree_dedup.h against the real container the reference uses:
// std::set<DecoratedMatch, DecoratedLess> range-constructed (indexed_match_decorator.h:90-104), libstdc++.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "../../experiments/rbtree_dedup.h"

using mvgcuda::DecoratedKey;

struct Elem { DecoratedKey k; int pos; };
struct Less { bool operator()(const Elem& a, const Elem& b) const { return mvgcuda::decorated_less(a.k, b.k); } };

int main(int argc, char** argv) {
  const int cases = argc > 1 ? atoi(argv[1]) : 20000;
  std::mt19937 rng(12345);
  long long total = 0, kept = 0;
  for (int c = 0; c < cases; ++c) {
    const int n = (c % 50 == 0) ? (int)(rng() % 3000) : (int)(rng() % 120);
    const int alpha = 2 + (int)(rng() % (c % 3 == 0 ? 6 : 400));  // small alphabets: many equal x / y / (x, y)
    std::vector<Elem> v(n);
    std::vector<DecoratedKey> keys(n);
    for (int i = 0; i < n; ++i) {
      DecoratedKey k;
      k.x1 = (float)(rng() % alpha) * 0.5f; k.y1 = (float)(rng() % alpha) * 0.25f;
      k.x2 = (float)(rng() % alpha); k.y2 = (float)(rng() % alpha);
      if (i > 0 && rng() % 9 == 0) k = keys[rng() % i];           // exact duplicates
      keys[i] = k;
      v[i] = Elem{k, i};
    }
    std::set<Elem, Less> s(v.begin(), v.end());
    std::vector<int> parent(n + 1), left(n + 1), right(n + 1), out(n + 1);
    std::vector<unsigned char> red(n + 1);
    const int m = mvgcuda::rbtree_dedup(keys.data(), n, parent.data(), left.data(), right.data(), red.data(), out.data());
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
