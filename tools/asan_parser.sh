#!/bin/bash
# Host parser under AddressSanitizer + UBSan on mutated copies of the reference's clips.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_build
python - <<'PY'
import sys
sys.path.insert(0, "tests")
from test_host_parser import crafted_negative_address_stream
open("tools/_build/crafted.es", "wb").write(crafted_negative_address_stream())
PY
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=all -pthread \
    tools/asan_parser.cpp mpeg_b200/csrc/host_parser.cpp mpeg_b200/csrc/coeff_pack.cpp -o tools/_build/asan_parser
ASAN_OPTIONS=detect_leaks=1 tools/_build/asan_parser tests/golden/test.mpeg1video tests/golden/test.mp2 "${1:-6000}" tools/_build/crafted.es
