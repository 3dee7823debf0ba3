#!/bin/bash
# Build a variant of libqgd_b200.so with extra -D flags for A/B runs on the GPU box:
#   tools/build_variant.sh <name> "<extra nvcc flags>" [objects...]
#       ->  quantumgatedesign.jl_b200/csrc/variants/libqgd_b200_<name>.so
# Only the listed objects (default: qgd_fast_m4, the order-8 register-operator sweeps the bench runs) are recompiled
# with the flags; every other object comes from the default in-tree build (run `make` there first).  Object names are
# the Makefile's: qgd_fast_m<M> / qgd_fast_s_m<M> (qgd_fast_unit.cu), qgd_inst_el<EL> (qgd_inst_unit.cu), or a plain unit.
# Select the variant at run time with QGD_B200_LIB=<path> (development only; the default library is csrc/libqgd_b200.so).
set -e
name=$1; extra=$2; shift 2 || true
units=${@:-qgd_fast_m4}
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/quantumgatedesign.jl_b200/csrc
bld=/tmp/qgd_variant_$name
rm -rf $bld; mkdir -p $bld $src/variants
objs=""
for o in $src/*.o; do
  b=$(basename $o .o); skip=0
  for u in $units; do [ "$b" = "$u" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for u in $units; do
  case $u in
    qgd_fast_s_m*) file=qgd_fast_unit.cu; defs="-DQGD_FAST_M=${u#qgd_fast_s_m} -DQGD_FAST_STRICT=1";;
    qgd_fast_m*) file=qgd_fast_unit.cu; defs="-DQGD_FAST_M=${u#qgd_fast_m}";;
    qgd_inst_el*) file=qgd_inst_unit.cu; defs="-DQGD_EL=${u#qgd_inst_el}";;
    *) file=$u.cu; defs="";;
  esac
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $defs $extra -Xptxas -v -c $src/$file -o $bld/$u.o 2> $bld/$u.ptxas.log || { tail -20 $bld/$u.ptxas.log; exit 1; }
  objs="$objs $bld/$u.o"
done
nvcc -shared -o $src/variants/libqgd_b200_$name.so $objs -lcudart -ldl 2>/dev/null
echo "built $src/variants/libqgd_b200_$name.so"
