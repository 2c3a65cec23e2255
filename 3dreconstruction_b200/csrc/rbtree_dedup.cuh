// rbtree_dedup.cuh -- what  std::set<IndexedMatchDecorator<float>>(first, last)  (libstdc++) leaves behind, restated on
// index-based nodes so that it runs on the GPU, one thread per image pair (SURVEY.md 8(a) row 13:
// IndexedMatchDecorator<float>::getDeduplicated, indexed_match_decorator.h:90-104).
//
// The reference's ordering predicate (indexed_match_decorator.h:33-53) is not a strict weak order, so the surviving
// elements and their order are DEFINED by the container's algorithm: hinted unique insertion at end() for every element
// of the range (_M_insert_range_unique -> _M_get_insert_hint_unique_pos(end(), k) -> _M_get_insert_unique_pos(k),
// stl_tree.h) and the red-black rebalancing of _Rb_tree_insert_and_rebalance (tree.cc).  This file restates exactly those
// steps (no pointers, no allocation, no STL); tests/native/test_rbtree_dedup.cpp checks it against the real std::set with
// the reference's full predicate on millions of elements, tests/test_gpu_parity.py against the oracle end to end.
//
// The predicate only ever looks at the LEFT feature (x1, y1):
//     if (all four coordinates equal) return false;
//     if (a.x1 < b.x1) return a.y1 < b.y1;   if (a.x1 > b.x1) return a.y1 < b.y1;   return a.x1 < b.x1;
// When x1 differs the first test cannot fire; when x1 is equal (or unordered) the last line returns false whatever the
// first test said.  So less(a, b) == ((a.x1 < b.x1) || (a.x1 > b.x1)) && (a.y1 < b.y1), and a node needs 8 bytes of key.
#pragma once
#if defined(__CUDACC__)
#define MVG_HD __host__ __device__ __forceinline__
#else
#define MVG_HD inline
#endif

namespace mvgcuda {

struct alignas(16) RbNode {  // global-memory form, 32 bytes: the descent reads one 16-byte half per level
  static constexpr int kNil = -1;
  float x, y;                // (x1, y1) of the match's left feature
  int left, right;
  int parent, red;
  int pad0, pad1;
};

struct alignas(16) RbNode16 {  // shared-memory form, 16 bytes: pairs with fewer than 65,535 matches
  static constexpr int kNil = 0xFFFF;
  float x, y;
  unsigned short left, right, parent, red;
};

MVG_HD bool decorated_less_xy(float ax, float ay, float bx, float by) {
  return ((ax < bx) || (ax > bx)) && (ay < by);
}

// nd[0 .. n-1]: node k is element k of the input with x / y already filled in; nd[n] is the header (parent = root,
// left = leftmost, right = rightmost).  Writes the input positions of the surviving elements in set (in-order) order to
// out[0 ..] and returns their number.
template <typename Node, typename OutT>
MVG_HD int rbtree_dedup(Node* __restrict__ nd, int n, OutT* __restrict__ out) {
  const int H = n, NIL = Node::kNil;
  nd[H].parent = NIL; nd[H].left = H; nd[H].right = H; nd[H].red = 1;
  int count = 0;

#define MVG_ROTATE_LEFT(x_)                                                     \
  {                                                                             \
    const int x = (x_);                                                         \
    const int y = nd[x].right;                                                  \
    const int yl = nd[y].left;                                                  \
    nd[x].right = yl;                                                           \
    if (yl != NIL) nd[yl].parent = x;                                           \
    const int xp = nd[x].parent;                                                \
    nd[y].parent = xp;                                                          \
    if (x == nd[H].parent) nd[H].parent = y;                                    \
    else if (x == nd[xp].left) nd[xp].left = y;                                 \
    else nd[xp].right = y;                                                      \
    nd[y].left = x;                                                             \
    nd[x].parent = y;                                                           \
  }
#define MVG_ROTATE_RIGHT(x_)                                                    \
  {                                                                             \
    const int x = (x_);                                                         \
    const int y = nd[x].left;                                                   \
    const int yr = nd[y].right;                                                 \
    nd[x].left = yr;                                                            \
    if (yr != NIL) nd[yr].parent = x;                                           \
    const int xp = nd[x].parent;                                                \
    nd[y].parent = xp;                                                          \
    if (x == nd[H].parent) nd[H].parent = y;                                    \
    else if (x == nd[xp].right) nd[xp].right = y;                               \
    else nd[xp].left = y;                                                       \
    nd[y].right = x;                                                            \
    nd[x].parent = y;                                                           \
  }

  for (int k = 0; k < n; ++k) {
    const float kx = nd[k].x, ky = nd[k].y;
    int y = NIL;  // parent for the insertion as returned by _M_get_insert_hint_unique_pos(end(), k); NIL: equivalent key present
    bool insert_left_forced = false;
    const int rm = nd[H].right;
    if (count > 0 && decorated_less_xy(nd[rm].x, nd[rm].y, kx, ky)) {
      y = rm;  // larger than the rightmost element: append
    } else {
      // _M_get_insert_unique_pos
      int cur = nd[H].parent;
      y = H;
      bool comp = true;
      while (cur != NIL) {
        y = cur;
        comp = decorated_less_xy(kx, ky, nd[cur].x, nd[cur].y);
        cur = comp ? nd[cur].left : nd[cur].right;
      }
      int j = y;
      bool decided = false;
      if (comp) {
        if (j == nd[H].left) decided = true;  // j == begin(): insert before it
        else {                                  // --j  (_Rb_tree_decrement; j is never the header here)
          if (nd[j].left != NIL) {
            int t = nd[j].left;
            while (nd[t].right != NIL) t = nd[t].right;
            j = t;
          } else {
            int t = nd[j].parent;
            while (j == nd[t].left) { j = t; t = nd[t].parent; }
            j = t;
          }
        }
      }
      if (!decided && !decorated_less_xy(nd[j].x, nd[j].y, kx, ky)) y = NIL;  // equivalent to j: dropped
    }
    if (y == NIL) continue;
    // _M_insert_ + _Rb_tree_insert_and_rebalance
    const bool insert_left = insert_left_forced || y == H || decorated_less_xy(kx, ky, nd[y].x, nd[y].y);
    const int z = k;
    nd[z].parent = y; nd[z].left = NIL; nd[z].right = NIL; nd[z].red = 1;
    if (insert_left) {
      nd[y].left = z;  // also sets leftmost when y is the header
      if (y == H) { nd[H].parent = z; nd[H].right = z; }
      else if (y == nd[H].left) nd[H].left = z;
    } else {
      nd[y].right = z;
      if (y == nd[H].right) nd[H].right = z;
    }
    int c = z;
    while (c != nd[H].parent && nd[nd[c].parent].red) {
      const int cp = nd[c].parent;
      const int pp = nd[cp].parent;
      if (cp == nd[pp].left) {
        const int u = nd[pp].right;
        if (u != NIL && nd[u].red) {
          nd[cp].red = 0; nd[u].red = 0; nd[pp].red = 1;
          c = pp;
        } else {
          if (c == nd[cp].right) { c = cp; MVG_ROTATE_LEFT(c) }
          nd[nd[c].parent].red = 0; nd[pp].red = 1;
          MVG_ROTATE_RIGHT(pp)
        }
      } else {
        const int u = nd[pp].left;
        if (u != NIL && nd[u].red) {
          nd[cp].red = 0; nd[u].red = 0; nd[pp].red = 1;
          c = pp;
        } else {
          if (c == nd[cp].left) { c = cp; MVG_ROTATE_RIGHT(c) }
          nd[nd[c].parent].red = 0; nd[pp].red = 1;
          MVG_ROTATE_LEFT(pp)
        }
      }
    }
    nd[nd[H].parent].red = 0;
    ++count;
  }
#undef MVG_ROTATE_LEFT
#undef MVG_ROTATE_RIGHT

  // in-order walk from the leftmost node (_Rb_tree_increment)
  int m = 0;
  if (count > 0) {
    int c = nd[H].left;
    while (c != H) {
      out[m++] = static_cast<OutT>(c);
      if (nd[c].right != NIL) {
        c = nd[c].right;
        while (nd[c].left != NIL) c = nd[c].left;
      } else {
        int p = nd[c].parent;
        while (p != H && c == nd[p].right) { c = p; p = nd[p].parent; }
        c = p;  // the header when climbing out of the rightmost node: the walk ends
      }
    }
  }
  return m;
}

}  // namespace mvgcuda
