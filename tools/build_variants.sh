#!/bin/bash
# Experiment builds of libmpegb200.so with other window-box heights (row phases of the staging follow from them).
# usage: tools/build_variants.sh "20 10" "18 9" ...   ->  mpeg_b200/variants/libmpegb200_L20C10.so ...
set -e
cd "$(dirname "$0")/../mpeg_b200/csrc"
mkdir -p ../variants
for v in "$@"; do
  set -- $v
  name="L$1C$2"
  env -u CC -u CXX make -s B=_build_$name OUT=../variants/libmpegb200_$name.so \
      EXTRA="-DMPEGB200_LUMA_BOX_ROWS=$1 -DMPEGB200_CHROMA_BOX_ROWS=$2" > /dev/null
  echo "built variants/libmpegb200_$name.so"
done
