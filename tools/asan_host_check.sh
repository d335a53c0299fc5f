#!/bin/bash
# Host side of the library under AddressSanitizer (no GPU needed): the containers, builders, error paths and the
# C drop-in driver, leak detection on.  Builds a private copy under $1 (default /tmp/storm_asan); nothing in the
# repository is touched.      tools/asan_host_check.sh [workdir]
set -eu
root=$(cd "$(dirname "$0")/.." && pwd)
work=${1:-/tmp/storm_asan}
rm -rf "$work" && mkdir -p "$work/obj"
for f in "$root"/stormbitmaps_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC -Xcompiler -fsanitize=address \
       -Xcompiler -fno-omit-frame-pointer -I "$root/include" -I "$root/stormbitmaps_b200/csrc" -c "$f" -o "$work/obj/$(basename "$f" .cu).o" &
done
wait
nvcc -shared -o "$work/libstorm_b200.so" "$work"/obj/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fsanitize=address -lasan
gcc -std=c99 -O1 -g -fsanitize=address -I "$root/include" "$root/tests/drivers/dropin_driver.c" -L "$work" -lstorm_b200 \
    -Wl,-rpath,"$work" -o "$work/dropin_driver"
for a in "4096 40 300 7" "65536 300 6553 42" "65536 60 150 3" "524288 50 30000 5" "1000 130 128 9" "131072 64 1 11"; do
  ASAN_OPTIONS=detect_leaks=1:protect_shadow_gap=0 "$work/dropin_driver" $a > "$work/driver.out" 2> "$work/driver.err" \
    || { echo "driver failed on: $a"; tail -n 20 "$work/driver.err"; exit 1; }
  if grep -q "ERROR: \(Address\|Leak\)Sanitizer" "$work/driver.err"; then echo "sanitizer report on: $a"; tail -n 30 "$work/driver.err"; exit 1; fi
done
echo "driver: 6 argument sets clean (no ASan or leak report)"
# the Python host-side tests against the same build
mkdir -p "$work/repo" && cp -r "$root/stormbitmaps_b200" "$root/tests" "$root/oracle" "$root/include" "$root/tools" "$work/repo/"
cp "$work/libstorm_b200.so" "$work/repo/stormbitmaps_b200/libstorm_b200.so"
python - "$work/repo/stormbitmaps_b200/build.py" <<'P'
import sys
p = sys.argv[1]
s = open(p).read().replace("def build(force: bool = False, verbose: bool = False) -> str:",
                           "def build(force: bool = False, verbose: bool = False) -> str:\n    return LIB\ndef _build_real(force: bool = False, verbose: bool = False) -> str:")
open(p, "w").write(s)
P
cd "$work/repo"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 \
  python -m pytest tests/test_abi.py -x -q -k "host or builder or loudly or exported or thread_pool" 2>&1 | tail -n 3
