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

# ---- the same partitioner WITH METIS (METHOD=KMETIS / PMETIS), checker of the product's fistr_b200_part_metis.  No METIS
# ships with the reference; the CUDA toolkit's libmetis_static.a (METIS 5, 64-bit idx_t) is used through the 32-bit front in
# oracle/metis_shim (the reference's code passes "int *" arrays).  Skipped when the toolkit has no such archive.
METIS_A=""
for p in /usr/local/cuda/lib64/libmetis_static.a /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a; do
  [ -f "$p" ] && METIS_A="$p" && break
done
if [ -n "$METIS_A" ]; then
  SHIM=$HERE/metis_shim
  objcopy --redefine-sym METIS_PartGraphKway=FB64_METIS_PartGraphKway \
          --redefine-sym METIS_PartGraphRecursive=FB64_METIS_PartGraphRecursive "$METIS_A" "$OUT/obj/libmetis64_renamed.a"
  mobjs=""
  for f in "$PART"/hecmw_partition.c "$PART"/hecmw_part_get_control.c; do
    b=$(basename "$f" .c)
    gcc $CFLAGS -DHECMW_PART_WITH_METIS -I"$SHIM" -c "$f" -o "$OUT/obj/${b}_metis.o"
  done
  gcc $CFLAGS -I"$SHIM" -c "$SHIM/metis32_adapter.c" -o "$OUT/obj/metis32_adapter.o"
  for o in $objs; do
    case "$(basename "$o" .o)" in
      hecmw_partition) mobjs="$mobjs $OUT/obj/hecmw_partition_metis.o" ;;
      hecmw_part_get_control) mobjs="$mobjs $OUT/obj/hecmw_part_get_control_metis.o" ;;
      *) mobjs="$mobjs $o" ;;
    esac
  done
  gcc -o "$OUT/hecmw_part1_metis" $mobjs "$OUT/obj/metis32_adapter.o" "$OUT/obj/libmetis64_renamed.a" -lm
  echo "built $OUT/hecmw_part1_metis"
fi
