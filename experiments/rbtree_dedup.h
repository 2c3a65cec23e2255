// rbtree_dedup.h -- device-compatible restatement of what
//     std::set<DecoratedMatch, DecoratedLess>(first, last)      (libstdc++, GCC 13)
// leaves behind, for moving the coordinate de-duplication (IndexedMatchDecorator::getDeduplicated,
// indexed_match_decorator.h:33-53,90-104) from the host pool to the GPU in a later round.
// The reference's comparator is not a strict weak order, so the surviving elements and their order are DEFINED by the
// container's algorithm: hinted unique insertion at end() for every element of the range
// (_M_insert_range_unique -> _M_get_insert_hint_unique_pos(end(), k) -> _M_get_insert_unique_pos(k), stl_tree.h) and
// the red-black rebalancing of _Rb_tree_insert_and_rebalance (tree.cc).  This file restates exactly those steps on
// index-based nodes (no pointers, no allocation, no STL), so the same code compiles for host and device.
// Status: NOT wired into the product.  tests/native/test_rbtree_dedup.cpp checks it against the real std::set on
// millions of elements (CPU); one thread per pair on the GPU is the intended use.
#pragma once
#if defined(__CUDACC__)
#define MVG_HD __host__ __device__
#else
#define MVG_HD
#endif

namespace mvgcuda {

struct DecoratedKey {  // left feature (x1, y1), right feature (x2, y2) of a match
  float x1, y1, x2, y2;
};

// IndexedMatchDecorator::operator< (indexed_match_decorator.h:33-53), same expression order
MVG_HD inline bool decorated_less(const DecoratedKey& a, const DecoratedKey& b) {
  if (a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2) return false;
  if (a.x1 < b.x1) return a.y1 < b.y1;
  if (a.x1 > b.x1) return a.y1 < b.y1;
  return a.x1 < b.x1;
}

// Node k of the tree is element k of the input; slot n is the header (parent = root, left = leftmost, right = rightmost).
// parent / left / right: int[n + 1], red: unsigned char[n + 1], out: int[n].  Returns the number of surviving elements
// and writes their input positions in set (in-order) order to out.
MVG_HD inline int rbtree_dedup(const DecoratedKey* key, int n, int* parent, int* left, int* right, unsigned char* red,
                               int* out) {
  const int H = n, NIL = -1;
  parent[H] = NIL; left[H] = H; right[H] = H; red[H] = 1;
  int count = 0;

  auto rotate_left = [&](int x) {
    const int y = right[x];
    right[x] = left[y];
    if (left[y] != NIL) parent[left[y]] = x;
    parent[y] = parent[x];
    if (x == parent[H]) parent[H] = y;
    else if (x == left[parent[x]]) left[parent[x]] = y;
    else right[parent[x]] = y;
    left[y] = x;
    parent[x] = y;
  };
  auto rotate_right = [&](int x) {
    const int y = left[x];
    left[x] = right[y];
    if (right[y] != NIL) parent[right[y]] = x;
    parent[y] = parent[x];
    if (x == parent[H]) parent[H] = y;
    else if (x == right[parent[x]]) right[parent[x]] = y;
    else left[parent[x]] = y;
    right[y] = x;
    parent[x] = y;
  };

  for (int k = 0; k < n; ++k) {
    int x = NIL, y = NIL;  // (x, y) as returned by _M_get_insert_hint_unique_pos(end(), k); y == NIL: equivalent key present
    if (count > 0 && decorated_less(key[right[H]], key[k])) {
      y = right[H];  // larger than the rightmost element: append
    } else {
      // _M_get_insert_unique_pos
      int cur = parent[H];
      y = H;
      bool comp = true;
      while (cur != NIL) {
        y = cur;
        comp = decorated_less(key[k], key[cur]);
        cur = comp ? left[cur] : right[cur];
      }
      int j = y;
      bool decided = false;
      if (comp) {
        if (j == left[H]) decided = true;  // j == begin(): insert at (x = NIL, y)
        else {                               // --j  (_Rb_tree_decrement; j is never the header here)
          if (left[j] != NIL) {
            int t = left[j];
            while (right[t] != NIL) t = right[t];
            j = t;
          } else {
            int t = parent[j];
            while (j == left[t]) { j = t; t = parent[t]; }
            j = t;
          }
        }
      }
      if (!decided && !decorated_less(key[j], key[k])) y = NIL;  // equivalent to j: dropped
    }
    if (y == NIL) continue;
    // _M_insert_ + _Rb_tree_insert_and_rebalance
    const bool insert_left = (x != NIL) || y == H || decorated_less(key[k], key[y]);
    const int z = k;
    parent[z] = y; left[z] = NIL; right[z] = NIL; red[z] = 1;
    if (insert_left) {
      left[y] = z;  // also sets leftmost when y is the header
      if (y == H) { parent[H] = z; right[H] = z; }
      else if (y == left[H]) left[H] = z;
    } else {
      right[y] = z;
      if (y == right[H]) right[H] = z;
    }
    int c = z;
    while (c != parent[H] && red[parent[c]]) {
      const int pp = parent[parent[c]];
      if (parent[c] == left[pp]) {
        const int u = right[pp];
        if (u != NIL && red[u]) {
          red[parent[c]] = 0; red[u] = 0; red[pp] = 1;
          c = pp;
        } else {
          if (c == right[parent[c]]) { c = parent[c]; rotate_left(c); }
          red[parent[c]] = 0; red[pp] = 1;
          rotate_right(pp);
        }
      } else {
        const int u = left[pp];
        if (u != NIL && red[u]) {
          red[parent[c]] = 0; red[u] = 0; red[pp] = 1;
          c = pp;
        } else {
          if (c == left[parent[c]]) { c = parent[c]; rotate_right(c); }
          red[parent[c]] = 0; red[pp] = 1;
          rotate_left(pp);
        }
      }
    }
    red[parent[H]] = 0;
    ++count;
  }

  // in-order walk from the leftmost node (_Rb_tree_increment)
  int m = 0;
  if (count > 0) {
    int c = left[H];
    while (c != H) {
      out[m++] = c;
      if (right[c] != NIL) {
        c = right[c];
        while (left[c] != NIL) c = left[c];
      } else {
        int p = parent[c];
        while (p != H && c == right[p]) { c = p; p = parent[p]; }
        // libstdc++: "if (x->right != y) x = y" -- the header case: climbing out of the rightmost node ends the walk
        c = (p == H) ? H : p;
      }
    }
  }
  return m;
}

}  // namespace mvgcuda
