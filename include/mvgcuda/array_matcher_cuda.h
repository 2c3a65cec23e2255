// array_matcher_cuda.h -- drop-in CUDA back-end behind the reference's array-level matcher API.
//
//   mvg::feature::ArrayMatcherCuda<unsigned char, Metric>  :  mvg::feature::ArrayMatcher<unsigned char, Metric>
//       (libs/feature/include/mvg/feature/matching_interface.h:16-64)
//
// It stands where ArrayMatcherBruteForce<unsigned char, SquaredEuclideanDistanceVectorized<unsigned char>>
// (matcher_brute_force.h:23-141, selected at compute_matches.cpp:226-227) stands today; results are bit-identical
// (same distances, same indices incl. the std::partial_sort tie behaviour of indexed_sort.h:52-66).
// Header-only; needs the reference's headers on the include path and libmvgcuda (include/mvgcuda.h) at link time.
//
// Behaviour kept from the BF matcher:
//   * Build(ptr, rows, dim): false if rows < 1 (matcher_brute_force.h:43-46).  Unlike BF, which borrows `ptr`
//     (Eigen::Map, :47-48), the rows are COPIED -- to HBM, once, by Build (mvgcuda_db_create); every search uploads only
//     its queries -- so the caller may free them.
//   * SearchNeighbours APPENDS k (index, distance) per query (push_back, :128-131); returns false and prints
//     "Too much asked nearest neighbors" when k > rows or nq < 1 (:107-110); re-entrant after Build (the reference
//     calls it from several OpenMP threads on one object, matcher_all_in_memory.h:84-107) -- calls are serialised
//     on the context by a mutex.
//   * DistanceType is float (Accumulator<unsigned char>::Type, metric.h:9-10).
// Only k = 1 and k = 2 and dimension 128 are accelerated -- what the path uses (NNN__ = 2, matcher_all_in_memory.h:102).
// Ties: k = 2 reproduces the two-slot machine std::partial_sort(first, first + 2, last) amounts to (indexed_sort.h:52-66);
// k = 1 returns the FIRST minimum, as partial_sort(first, first + 1, last) does (strict < in __heap_select).
// There is no CPU fallback: without a B200 the calls fail (return false) with the CUDA error on std::cerr.
#ifndef MVGCUDA_ARRAY_MATCHER_CUDA_H_
#define MVGCUDA_ARRAY_MATCHER_CUDA_H_

#include <algorithm>
#include <iostream>
#include <memory>
#include <mutex>
#include <vector>

#include "mvg/feature/matching_interface.h"
#include "mvg/feature/metric.h"
#include "mvgcuda.h"

namespace mvg {
namespace feature {

// One process-wide context per device, shared by all matcher objects.
class MvgCudaContextPool {
 public:
  static mvgcuda_ctx* get(int device, std::mutex** mtx) {
    static std::mutex pool_mtx;
    static std::vector<std::pair<mvgcuda_ctx*, std::mutex*> > pool;
    std::lock_guard<std::mutex> lock(pool_mtx);
    if ((int)pool.size() <= device) pool.resize(device + 1, std::make_pair((mvgcuda_ctx*)NULL, (std::mutex*)NULL));
    if (!pool[device].first) {
      mvgcuda_ctx* c = NULL;
      if (mvgcuda_create(device, &c) != MVGCUDA_OK) {
        std::cerr << "mvgcuda: " << mvgcuda_last_error(NULL) << std::endl;
        return NULL;
      }
      pool[device] = std::make_pair(c, new std::mutex());
    }
    *mtx = pool[device].second;
    return pool[device].first;
  }
};

template <typename Scalar = unsigned char, typename Metric = SquaredEuclideanDistanceVectorized<Scalar> >
class ArrayMatcherCuda : public ArrayMatcher<Scalar, Metric> {
  static_assert(sizeof(Scalar) == 1, "ArrayMatcherCuda handles 8-bit descriptors (the u8 SIFT path) only");

 public:
  typedef typename Metric::ResultType DistanceType;

  explicit ArrayMatcherCuda(int device = 0, int tie_mode = MVGCUDA_TIE_REFERENCE) : device_(device), tie_mode_(tie_mode) {}
  virtual ~ArrayMatcherCuda() { Release(); }
  ArrayMatcherCuda(const ArrayMatcherCuda&) = delete;
  ArrayMatcherCuda& operator=(const ArrayMatcherCuda&) = delete;

  bool Build(const Scalar* dataset, int rows_num, int dimension) {
    Release();
    if (rows_num < 1) return false;
    if (dimension != MVGCUDA_DIM) {
      std::cerr << "ArrayMatcherCuda: dimension must be " << MVGCUDA_DIM << std::endl;
      return false;
    }
    std::mutex* mtx = NULL;
    mvgcuda_ctx* ctx = MvgCudaContextPool::get(device_, &mtx);
    if (!ctx) return false;
    const uint8_t* rows = reinterpret_cast<const uint8_t*>(dataset);
    // the 2-NN kernel needs two db rows; a single-row db (k == 1 only) is kept as two copies of that row
    std::vector<uint8_t> twice;
    int n = rows_num;
    if (rows_num == 1) {
      twice.assign(rows, rows + MVGCUDA_DIM);
      twice.insert(twice.end(), rows, rows + MVGCUDA_DIM);
      rows = twice.data();
      n = 2;
    }
    std::lock_guard<std::mutex> lock(*mtx);
    if (mvgcuda_db_create(ctx, rows, n, &db_) != MVGCUDA_OK) {
      std::cerr << "mvgcuda: " << mvgcuda_last_error(ctx) << std::endl;
      db_ = NULL;
      return false;
    }
    rows_ = rows_num;
    return true;
  }

  // std::min_element semantics: first minimum == lowest index (matcher_brute_force.h:61-89)
  bool SearchNeighbour(const Scalar* query, int* indice, DistanceType* distance) {
    if (rows_ < 1) return false;
    std::vector<int> vi;
    std::vector<DistanceType> vd;
    if (!Search(query, 1, &vi, &vd, 1, MVGCUDA_TIE_LOWEST_INDEX)) return false;
    *indice = vi[0];
    *distance = vd[0];
    return true;
  }

  bool SearchNeighbours(const Scalar* query, int query_num, std::vector<int>* vec_indice,
                        std::vector<DistanceType>* vec_distance, size_t nearest_neighbor_num) {
    if (nearest_neighbor_num > (size_t)rows_ || query_num < 1) {
      std::cerr << "Too much asked nearest neighbors" << std::endl;
      return false;
    }
    if (nearest_neighbor_num < 1 || nearest_neighbor_num > 2) {
      std::cerr << "ArrayMatcherCuda: only 1 or 2 nearest neighbours are supported" << std::endl;
      return false;
    }
    // k == 1: partial_sort(first, first + 1, last) keeps the first minimum, i.e. the lowest index
    return Search(query, query_num, vec_indice, vec_distance, (int)nearest_neighbor_num,
                  nearest_neighbor_num == 1 ? (int)MVGCUDA_TIE_LOWEST_INDEX : tie_mode_);
  }

 private:
  void Release() {
    if (db_) {
      std::mutex* mtx = NULL;
      mvgcuda_ctx* ctx = MvgCudaContextPool::get(device_, &mtx);
      if (ctx) {
        std::lock_guard<std::mutex> lock(*mtx);
        mvgcuda_db_destroy(ctx, db_);
      }
      db_ = NULL;
    }
    rows_ = 0;
  }

  bool Search(const Scalar* query, int nq, std::vector<int>* vi, std::vector<DistanceType>* vd, int k, int tie) {
    std::mutex* mtx = NULL;
    mvgcuda_ctx* ctx = MvgCudaContextPool::get(device_, &mtx);
    if (!ctx || !db_) return false;
    std::vector<int32_t> idx(2 * (size_t)nq);
    std::vector<float> dist(2 * (size_t)nq);
    int rc;
    {
      std::lock_guard<std::mutex> lock(*mtx);  // re-entrant after Build, like the reference (matcher_all_in_memory.h:84-107)
      rc = mvgcuda_db_knn2(ctx, db_, reinterpret_cast<const uint8_t*>(query), nq, tie, idx.data(), dist.data());
      if (rc != MVGCUDA_OK) std::cerr << "mvgcuda: " << mvgcuda_last_error(ctx) << std::endl;
    }
    if (rc != MVGCUDA_OK) return false;
    for (int q = 0; q < nq; ++q)
      for (int n = 0; n < k; ++n) {
        vd->push_back(static_cast<DistanceType>(dist[2 * q + n]));
        vi->push_back(rows_ == 1 ? 0 : idx[2 * q + n]);
      }
    return true;
  }

  int device_, tie_mode_;
  mvgcuda_db* db_ = NULL;
  int rows_ = 0;
};

}  // namespace feature
}  // namespace mvg

#endif  // MVGCUDA_ARRAY_MATCHER_CUDA_H_
