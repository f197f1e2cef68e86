#!/bin/bash
# Experiment builds of libmpegb200.so (same C-ABI; select one with MPEGB200_LIB=mpeg_b200/variants/libX.so).
# usage: tools/build_variants.sh name "extra nvcc flags" [name "flags" ...]
#   e.g. tools/build_variants.sh exp "-DMPEGB200_EXPERIMENTS" occ7 "-DMPEGB200_EXPERIMENTS -DMPEGB200_EXP_COEF_ALIAS"
set -e
cd "$(dirname "$0")/../mpeg_b200/csrc"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  env -u CC -u CXX make -s B=_build_$name OUT=../variants/lib$name.so EXTRA="$flags" > /dev/null
  echo "built variants/lib$name.so ($flags)"
done
