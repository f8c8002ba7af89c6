#!/bin/bash
# ORACLE build recipe (test infrastructure): compiles the REFERENCE's own partitioner hecmw_part1
# (hecmw1/tools/partitioner + hecmw1/src/common, plain C) from the sources where they lie under the
# reference tree.  Nothing is copied into the repository; objects and the binary go to oracle/_ref/
# (git-ignored, shipped to the GPU box by gpurun).  Serial build (-DHECMW_SERIAL), METIS absent =>
# METHOD=RCB only.  The reference's own build system (cmake) is not used.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
mkdir -p "$OUT/obj"
COMMON=$REF/hecmw1/src/common
PART=$REF/hecmw1/tools/partitioner
CFLAGS="-O2 -w -DHECMW_SERIAL -I$COMMON -I$PART -I$REF/hecmw1/src/hecmw"
objs=""
for f in "$COMMON"/*.c "$PART"/*.c; do
  b=$(basename "$f" .c)
  case "$b" in
    hecmw_io_nastran) continue ;;        # does not compile (unterminated string); the reference ships a dummy
    varray_test) continue ;;
  esac
  o="$OUT/obj/$b.o"
  gcc $CFLAGS -c "$f" -o "$o" || { echo "skip $b"; continue; }
  objs="$objs $o"
done
gcc -o "$OUT/hecmw_part1" $objs -lm
echo "built $OUT/hecmw_part1"
