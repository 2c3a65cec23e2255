#!/usr/bin/env bash
# build_ref.sh -- ORACLE L0 build recipe (test infrastructure).
#
# Compiles oracle/ref_driver.cpp against the reference's OWN headers, where they lie under
# $MVG_REF (default /root/reference), into oracle/_ref/ (git-ignored, travels to the GPU box):
#   oracle/_ref/libmvgref.so      single-threaded == the reference's shipped configuration
#                                 (USE_OPENMP is defined nowhere in its build)
#   oracle/_ref/libmvgref_omp.so  the reference's own OpenMP mode (-DUSE_OPENMP -fopenmp,
#                                 matcher_all_in_memory.h:87-89)
# The reference's own CMake tree is NOT run (it does not configure under CMake 4 and libs/base/src is
# Windows-only); only g++ on the header-only hot path + two portable .cpp files.
# No reference source is copied into the repository: the two generated stubs and the 2-token
# `typename` patch of matcher_brute_force.h (lines 77 and 80 do not compile under g++) live in a
# temp directory that is deleted afterwards.
set -euo pipefail
REF="${MVG_REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/libs/feature/include" ]; then
  echo "build_ref.sh: reference tree not found at $REF (nothing built)" >&2
  exit 3
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/mvg/feature"
# (1) stubs for the two CMake-generated headers (parse-files/config.h.in, version.h.in)
cat > "$TMP/mvg/config.h" <<'EOF'
#ifndef MVG_CONFIG_H
#define MVG_CONFIG_H
#define MVG_OS_LINUX
#define MVG_VERSION_POSTFIX ""
#define MVG_WORD_SIZE 64
#endif
EOF
cat > "$TMP/mvg/version.h" <<'EOF'
#ifndef MVG_VERSION_H
#define MVG_VERSION_H
#define MVG_VERSION_STR "0.1.0"
#endif
EOF
# (2) typename patch on a temp copy, placed earlier on the include path
sed 's/std::vector<DistanceType>::const_iterator/typename &/' \
  "$REF/libs/feature/include/mvg/feature/matcher_brute_force.h" > "$TMP/mvg/feature/matcher_brute_force.h"
INC=(-I"$TMP" -I"$REF/libs/base/include" -I"$REF/libs/feature/include" -I"$REF/3rdparty/eigen3"
     -I"$REF/3rdparty/flann/src/cpp" -I"$REF/libs/base/src")
SRC=("$HERE/ref_driver.cpp" "$REF/libs/base/src/utils/file_system.cpp" "$REF/libs/base/src/utils/wildcard.cpp")
# -O2, no -ffast-math, no -D_GLIBCXX_PARALLEL (would silently change std::partial_sort)
CXXFLAGS=(-std=c++11 -O2 -fPIC -shared -w -DORACLE_WITH_FLANN)
g++ "${CXXFLAGS[@]}" "${INC[@]}" "${SRC[@]}" -o "$OUT/libmvgref.so"
g++ "${CXXFLAGS[@]}" -DUSE_OPENMP -fopenmp "${INC[@]}" "${SRC[@]}" -o "$OUT/libmvgref_omp.so"
# (3) the step after the path: the reference's AC-RANSAC geometric filter, for golden matches.f/h.txt
GEOM_INC=("${INC[@]}" -I"$REF/libs/multiview/include" -I"$REF/libs/camera/include" -I"$REF/libs/image/include")
GEOM_SRC=("$HERE/ref_geom_driver.cpp" "$REF/libs/base/src/utils/file_system.cpp" "$REF/libs/base/src/utils/wildcard.cpp"
          "$REF/libs/multiview/src/solver_fundamental_kernel.cpp" "$REF/libs/multiview/src/solver_homography_kernel.cpp"
          "$REF/libs/multiview/src/conditioning.cpp" "$REF/libs/base/src/math/numeric.cpp" "$REF/libs/camera/src/projection.cpp")
if g++ -std=c++11 -O2 -fPIC -shared -w -Wl,--no-undefined "${GEOM_INC[@]}" "${GEOM_SRC[@]}" -o "$OUT/libmvgref_geom.so" 2> "$TMP/geom.log"; then
  echo "built $OUT/libmvgref_geom.so"
else
  echo "build_ref.sh: geometric-filter oracle NOT built:" >&2; head -20 "$TMP/geom.log" >&2
fi
echo "built $OUT/libmvgref.so $OUT/libmvgref_omp.so from $REF"
