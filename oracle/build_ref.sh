#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- not part of the product path.
#
# Builds the *unmodified* reference implementation of the low-level descriptor
# hot path (TSampleAnalyser + vendored LibXtract / Aubio / libresample) straight
# from the sources where they lie under $AFEC_REF (default /root/reference).
# Nothing is copied into this repository: the compiler is pointed at the
# reference files by path, and every output (objects, archives, the harness
# binary `afec_ref`) goes to oracle/_ref/ which is git-ignored.
#
# Recipe follows SURVEY.md section 8(c). The reference's own cmake build is not
# run (its prebuilt 3rdParty .a files are git-LFS stubs in this checkout).
#
# Usage: oracle/build_ref.sh [-j N]
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${AFEC_REF:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [[ "${1:-}" == "-j" ]]; then JOBS="$2"; fi

if [[ ! -d "$REF/Source/Crawler/FeatureExtraction" ]]; then
  echo "build_ref: reference tree not found at $REF (nothing to do)" >&2
  exit 3
fi

mkdir -p "$OUT/obj/xtract" "$OUT/obj/aubio" "$OUT/obj/resample" \
         "$OUT/obj/boostser" "$OUT/obj/afec" "$OUT/shim/sys" "$OUT/aubiocfg"

T="$REF/3rdParty"
S="$REF/Source"

# ---- shims (our own files, not reference code) --------------------------------
: > "$OUT/shim/sys/sysctl.h"       # glibc >= 2.32 dropped <sys/sysctl.h>
cat > "$OUT/aubiocfg/config.h" <<'EOF'
#define HAVE_STDLIB_H 1
#define HAVE_STDIO_H 1
#define HAVE_COMPLEX_H 1
#define HAVE_MATH_H 1
#define HAVE_STRING_H 1
#define HAVE_LIMITS_H 1
#define HAVE_STDARG_H 1
#define HAVE_ERRNO_H 1
#define HAVE_C99_VARARGS_MACROS 1
#define HAVE_AUBIO_DOUBLE 1
#define HAVE_MEMCPY_HACKS 1
EOF
cat > "$OUT/shim/iconv_shim.c" <<'EOF'
/* the reference links GNU libiconv; glibc's iconv has the same semantics */
#include <iconv.h>
#include <stddef.h>
void* libiconv_open(const char* to, const char* from) { return (void*)iconv_open(to, from); }
size_t libiconv(void* cd, char** in, size_t* inleft, char** out, size_t* outleft)
{ return iconv((iconv_t)cd, in, inleft, out, outleft); }
int libiconv_close(void* cd) { return iconv_close((iconv_t)cd); }
EOF

# compile helper: cc_one <compiler> <flags-file> <src> <objdir>
cc_one() {
  local cc="$1" flags="$2" src="$3" objdir="$4"
  local base; base="$(echo "$src" | sed -e "s#^$REF/##" -e 's#[/ ]#_#g')"
  local obj="$objdir/${base%.*}.o"
  if [[ ! -f "$obj" || "$src" -nt "$obj" ]]; then
    # shellcheck disable=SC2046
    $cc $(cat "$flags") -c "$src" -o "$obj" || { echo "FAILED: $src" >&2; return 1; }
  fi
}
export -f cc_one
export REF

par() { # par <compiler> <flags-file> <objdir>  (sources on stdin)
  xargs -P "$JOBS" -I{} bash -c 'cc_one "$0" "$1" "{}" "$2"' "$1" "$2" "$3"
}

# ---- 1. LibXtract -------------------------------------------------------------
echo "-O3 -fPIC -fcommon -std=c99 -w -DUSE_OOURA -I$T/LibXtract/Dist/include -I$T/LibXtract/Dist/src" > "$OUT/obj/xtract.flags"
ls "$T"/LibXtract/Dist/src/*.c "$T"/LibXtract/Dist/src/ooura/*.c \
   "$T"/LibXtract/Dist/src/dywapitchtrack/*.c "$T"/LibXtract/Dist/src/c-ringbuf/*.c \
  | par gcc "$OUT/obj/xtract.flags" "$OUT/obj/xtract"

# ---- 2. Aubio (double precision, Ooura FFT) ----------------------------------------
echo "-O3 -fPIC -std=c99 -w -DHAVE_CONFIG_H -I$OUT/aubiocfg -I$T/Aubio/Dist/src" > "$OUT/obj/aubio.flags"
find "$T/Aubio/Dist/src" -name '*.c' -not -path '*/io/*' \
  | par gcc "$OUT/obj/aubio.flags" "$OUT/obj/aubio"

# ---- 3. libresample -----------------------------------------------------------
echo "-O3 -fPIC -w -I$T/Resample/Dist/src -I$T/Resample/Dist/include" > "$OUT/obj/resample.flags"
ls "$T"/Resample/Dist/src/resample.c "$T"/Resample/Dist/src/resamplesubs.c "$T"/Resample/Dist/src/filterkit.c \
  | par gcc "$OUT/obj/resample.flags" "$OUT/obj/resample"

# ---- 4. boost_serialization (SampleAnalyser.o instantiates model serializers) --------
echo "-O2 -fPIC -w -std=c++11 -I$T/Boost/Dist" > "$OUT/obj/boostser.flags"
ls "$T"/Boost/Dist/libs/serialization/src/*.cpp | grep -v -e xml -e '/w[a-z_]*\.cpp$' -e utf8_codecvt -e codecvt_null \
  | par g++ "$OUT/obj/boostser.flags" "$OUT/obj/boostser"

# ---- 5. AFEC libraries (CMake Release = -O3 -DNDEBUG, no fast-math) ----------------
INC="-I$OUT/shim -I$S/Core -I$S/Crawler"
for d in Boost/Dist Aubio/Dist/src LibXtract/Dist/include Resample/Dist/include Msgpack/Dist/include \
         Shark/Dist/include OpenBLAS/Dist Sqlite/Dist/src LightGBM/Dist/include Iconv/Dist/include \
         OggVorbis/Dist/include Ogg/Dist/include Flac/Dist/include ZLib/Dist Mpg123/Dist/src/libmpg123; do
  INC="$INC -I$T/$d"
done
COMMON="-O3 -std=c++11 -fPIC -DMRelease -DMArch_X64 -DMCompiler_GCC -DMLinux -DNDEBUG -w $INC"
for proj in CoreTypes AudioTypes CoreFileFormats; do
  echo "$COMMON -DM$proj -I$S/Core/$proj/Source" > "$OUT/obj/$proj.flags"
  ls "$S"/Core/$proj/Source/*.cpp | grep -v -e '/Mac' -e '/Win' -e PrecompiledHeader \
    | par g++ "$OUT/obj/$proj.flags" "$OUT/obj/afec"
done
echo "$COMMON -DMFeatureExtraction -I$S/Crawler/FeatureExtraction/Source" > "$OUT/obj/FeatureExtraction.flags"
ls "$S"/Crawler/FeatureExtraction/Source/*.cpp | grep -v PrecompiledHeader \
  | par g++ "$OUT/obj/FeatureExtraction.flags" "$OUT/obj/afec"
# the one C file of CoreFileFormats
echo "-O2 -fPIC -w $INC" > "$OUT/obj/c.flags"
ls "$S"/Core/CoreFileFormats/Source/*.c 2>/dev/null | par gcc "$OUT/obj/c.flags" "$OUT/obj/afec" || true
gcc -O2 -fPIC -c "$OUT/shim/iconv_shim.c" -o "$OUT/obj/iconv_shim.o"

# ---- 6. harness (our file: oracle/ref_harness.cpp) ---------------------------------
g++ $COMMON -DMFeatureExtraction -c "$HERE/ref_harness.cpp" -o "$OUT/obj/ref_harness.o"

SQLITE="$(ls /usr/lib/x86_64-linux-gnu/libsqlite3.so.0 2>/dev/null || true)"
g++ -o "$OUT/afec_ref" "$OUT/obj/ref_harness.o" "$OUT"/obj/afec/*.o "$OUT"/obj/xtract/*.o \
    "$OUT"/obj/aubio/*.o "$OUT"/obj/resample/*.o "$OUT"/obj/boostser/*.o "$OUT/obj/iconv_shim.o" \
    $SQLITE -lpthread -ldl -lrt -Wl,--unresolved-symbols=ignore-all
echo "built $OUT/afec_ref"
