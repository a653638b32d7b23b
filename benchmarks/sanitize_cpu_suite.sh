#!/bin/bash
# Runs the Python-driven CPU suite against AddressSanitizer + UBSan builds of the host libraries (libkvbm_physical.so,
# libkvbm_router.so) and of the oracle, via LD_PRELOAD of the sanitizer runtimes.  The in-tree libraries are put back
# (byte-identical) afterwards.  No GPU needed.   usage: benchmarks/sanitize_cpu_suite.sh
set -u
cd "$(dirname "$0")/.."
TMP=$(mktemp -d)
for f in dynamo_b200/libkvbm_physical.so dynamo_b200/libkvbm_router.so oracle/libkvbm_oracle.so; do cp "$f" "$TMP/$(basename $f)"; done
restore() { for f in dynamo_b200/libkvbm_physical.so dynamo_b200/libkvbm_router.so oracle/libkvbm_oracle.so; do cp "$TMP/$(basename $f)" "$f"; touch "$f"; done; rm -rf "$TMP"; }
trap restore EXIT
XXH=$(python -c "import pyarrow,os;print(os.path.join(pyarrow.get_include(),'arrow','vendored','xxhash'))")
SAN="-O1 -g -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -pthread"
g++ -std=c++17 $SAN -I include -I /usr/local/cuda/include dynamo_b200/csrc/host/transfer_manager.cpp dynamo_b200/csrc/host/multicast.cpp \
    -L dynamo_b200 -lkvbm_kernels -L /usr/local/cuda/lib64 -lcudart -ldl -Wl,-rpath,'$ORIGIN' -o dynamo_b200/libkvbm_physical.so || exit 1
g++ -std=c++17 $SAN -I "$XXH" -I include -o dynamo_b200/libkvbm_router.so dynamo_b200/csrc/router/radix_tree.cpp dynamo_b200/csrc/router/kv_events.cpp || exit 1
gcc -std=c11 $SAN -o oracle/libkvbm_oracle.so oracle/kvbm_oracle.c -lm || exit 1
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
  python -m pytest tests -q -m "not gpu" -p no:cacheprovider -k "not sanitizer and not asan and not c_program and not plain_c"
