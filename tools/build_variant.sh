#!/bin/bash
# Build a variant of libqgd_b200.so with extra -D flags for A/B runs on the GPU box:
#   tools/build_variant.sh <name> "<extra nvcc flags>"   ->  quantumgatedesign.jl_b200/csrc/variants/libqgd_b200_<name>.so
# Select it at run time with QGD_B200_LIB=<path> (development only; the default library is csrc/libqgd_b200.so).
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/quantumgatedesign.jl_b200/csrc
bld=/tmp/qgd_variant_$name
rm -rf $bld; mkdir -p $bld/quantumgatedesign.jl_b200 $root/quantumgatedesign.jl_b200/csrc/variants
cp -r $root/include $bld/include
mkdir -p $bld/quantumgatedesign.jl_b200/csrc
cp $src/*.cu $src/*.cuh $src/*.h $src/Makefile $bld/quantumgatedesign.jl_b200/csrc/
make -C $bld/quantumgatedesign.jl_b200/csrc -j12 NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $extra" > $bld/make.log 2>&1 || { tail -20 $bld/make.log; exit 1; }
cp $bld/quantumgatedesign.jl_b200/csrc/libqgd_b200.so $src/variants/libqgd_b200_$name.so
echo "built $src/variants/libqgd_b200_$name.so"
