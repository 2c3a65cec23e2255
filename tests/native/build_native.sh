#!/usr/bin/env bash
# Build the native test binaries into build/native/ (git-ignored, travels to the GPU box):
#   test_adaptors   needs the reference headers ($MVG_REF, default /root/reference) -- same prelude/stubs as oracle/build_ref.sh
set -euo pipefail
REF="${MVG_REF:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$ROOT/build/native"
mkdir -p "$OUT"
if [ ! -d "$REF/libs/feature/include" ]; then echo "build_native.sh: no reference tree at $REF, skipping test_adaptors" >&2; exit 0; fi
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/mvg/feature"
printf '#ifndef MVG_CONFIG_H\n#define MVG_CONFIG_H\n#define MVG_OS_LINUX\n#define MVG_VERSION_POSTFIX ""\n#define MVG_WORD_SIZE 64\n#endif\n' > "$TMP/mvg/config.h"
printf '#ifndef MVG_VERSION_H\n#define MVG_VERSION_H\n#define MVG_VERSION_STR "0.1.0"\n#endif\n' > "$TMP/mvg/version.h"
sed 's/std::vector<DistanceType>::const_iterator/typename &/' "$REF/libs/feature/include/mvg/feature/matcher_brute_force.h" > "$TMP/mvg/feature/matcher_brute_force.h"
g++ -std=c++11 -O2 -w -pthread -I"$TMP" -I"$ROOT/include" -I"$REF/libs/base/include" -I"$REF/libs/feature/include" -I"$REF/3rdparty/eigen3" \
    -I"$REF/libs/base/src" -I"$REF/libs/multiview/include" -I"$REF/libs/camera/include" -I"$REF/libs/image/include" \
    "$ROOT/tests/native/test_adaptors.cpp" "$REF/libs/base/src/utils/file_system.cpp" "$REF/libs/base/src/utils/wildcard.cpp" \
    "$REF/libs/multiview/src/solver_fundamental_kernel.cpp" "$REF/libs/multiview/src/solver_homography_kernel.cpp" "$REF/libs/multiview/src/conditioning.cpp" "$REF/libs/base/src/math/numeric.cpp" \
    "$REF/libs/camera/src/projection.cpp" \
    -L"$ROOT/3dreconstruction_b200" -lmvgcuda -Wl,-rpath,'$ORIGIN/../../3dreconstruction_b200' -o "$OUT/test_adaptors"
echo "built $OUT/test_adaptors"
