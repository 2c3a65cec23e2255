#!/usr/bin/env bash
# Oracle-side tooling: build oracle/_ref/libmvgsift.so = the reference's SIFT wrapper + vendored VLFeat subset,
# compiled from the sources where they lie under $MVG_REF.  A temp copy of pixel_types.h fixes the MSVC-tolerated
# typo at libs/image/include/mvg/image/pixel_types.h:104 (`+ *0.59*g()`), as noted in SURVEY.md 8(d).
set -euo pipefail
REF="${MVG_REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"; mkdir -p "$OUT"
[ -d "$REF/3rdparty/sift/vl" ] || { echo "no reference tree at $REF" >&2; exit 3; }
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/mvg/image" "$TMP/obj"
printf '#ifndef MVG_CONFIG_H\n#define MVG_CONFIG_H\n#define MVG_OS_LINUX\n#define MVG_VERSION_POSTFIX ""\n#define MVG_WORD_SIZE 64\n#define MVG_HAS_JPEG 0\n#define MVG_HAS_PNG 0\n#endif\n' > "$TMP/mvg/config.h"
printf '#ifndef MVG_VERSION_H\n#define MVG_VERSION_H\n#define MVG_VERSION_STR "0.1.0"\n#endif\n' > "$TMP/mvg/version.h"
sed 's/+ \*0\.59\*g()/+ 0.59*g()/' "$REF/libs/image/include/mvg/image/pixel_types.h" > "$TMP/mvg/image/pixel_types.h"
for f in generic host imopv imopv_sse2 mathop mathop_sse2 random sift; do
  gcc -std=gnu99 -O2 -fPIC -w -msse2 -DVL_DISABLE_AVX -I"$REF/3rdparty/sift/vl" -c "$REF/3rdparty/sift/vl/$f.c" -o "$TMP/obj/$f.o"
done
g++ -std=c++11 -O2 -fPIC -shared -w -I"$TMP" -I"$REF/libs/base/include" -I"$REF/libs/feature/include" -I"$REF/libs/image/include" \
    -I"$REF/3rdparty/eigen3" -I"$REF/3rdparty/sift/vl" -I"$REF/3rdparty" "$HERE/sift_ref.cpp" "$TMP"/obj/*.o -o "$OUT/libmvgsift.so"
echo "built $OUT/libmvgsift.so"
