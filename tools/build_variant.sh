#!/bin/bash
# A/B builds of libbhmel.so: tools/build_variant.sh <name> [extra nvcc flags...]  ->  build/libbhmel_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
python -c "from beatheritage_b200 import build; build.generate()"
/usr/local/cuda/bin/nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
  -o build/libbhmel_$name.so beatheritage_b200/csrc/bhmel.cu
echo build/libbhmel_$name.so
