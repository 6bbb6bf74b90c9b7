#!/bin/bash
# Build a kernel-parameter variant of libnxsearch.so for A/B runs:
#   scripts/build_variant.sh NAME -DST_CWARPS=12 -DST_SLOTS=4 ...
# -> nxsearch_b200/lib/variants/NAME/libnxsearch.so  (select with NXSB_LIBRARY)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/nxsearch_b200/lib/variants/$name
mkdir -p "$out"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-fvisibility=hidden -I"$root/include" "$@" \
    -c "$root/nxsearch_b200/csrc/gpu/engine.cu" -o "$out/engine.o"
objs=$(ls "$root"/nxsearch_b200/lib/obj/*.c.o | grep -v "corpus.c.o\|querytools.c.o")
nvcc -shared -o "$out/libnxsearch.so" $objs "$out/engine.o" -cudart static -lpthread -lm
echo "$out/libnxsearch.so"
