#!/bin/bash
# Host code of libkzp_b200.so under AddressSanitizer + UndefinedBehaviorSanitizer (no GPU needed).
# Builds an instrumented copy of the library in a scratch directory (nvcc -Xcompiler -fsanitize=address,undefined for
# every translation unit's host half, linked by g++) and runs the CPU test suite against it with the sanitizer
# runtimes preloaded into Python. Covers the C ABI's host entry points: zkey / wtns loaders, host field and curve
# arithmetic, proof assembly, decimal printing, pairing + verifier, witness packing (SSE), pool checkout queue, CLI.
#   scripts/host_sanitizer_run.sh            -> prints the pytest summary and the number of sanitizer reports
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
WORK="${1:-/tmp/kzp_host_sanitizers}"
rm -rf "$WORK" && mkdir -p "$WORK"
(cd "$ROOT" && tar --exclude=.git --exclude=gpurun_out --exclude=profiles -cf - .) | (cd "$WORK" && tar xf -)
cd "$WORK"
python - <<'PY'
p = "keyless-zk-proofs_b200/build.py"
s = open(p).read()
s = s.replace('"-Xcompiler", "-fPIC,-fvisibility=default",',
              '"-Xcompiler", "-fPIC,-fvisibility=default,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer",')
open(p, "w").write(s)
PY
# compile with the copy's build.py (its nvcc link step cannot find libasan: expected), then link with g++
python keyless-zk-proofs_b200/build.py --force >/dev/null 2>&1 || true
cd keyless-zk-proofs_b200
g++ -shared -o libkzp_b200.so build/*.o -fsanitize=address,undefined -L/usr/local/cuda/lib64 -lcudart_static -lpthread -ldl -lrt
g++ -O2 -std=c++17 csrc/cli_main.cpp -o kzp_prove -L. -lkzp_b200 '-Wl,-rpath,$ORIGIN' -fsanitize=address,undefined
cd ..
python - <<'PY'
import importlib.util, os
spec = importlib.util.spec_from_file_location("b", "keyless-zk-proofs_b200/build.py")
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
open(os.path.join(m.BUILD, "stamp.sha256"), "w").write(m._digest())  # the instrumented library is "current"
PY
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0
# the two tests that build their own sanitizer binaries are left out: a preloaded ASan runtime cannot host a TSan child
python -m pytest tests -q -s -m "not gpu" -p no:cacheprovider -k "not thread_sanitizer and not under_sanitizers" > "$WORK/run.log" 2>&1 || true
tail -1 "$WORK/run.log"
echo "UBSan reports: $(grep -c 'runtime error' "$WORK/run.log" || true)   ASan reports: $(grep -c 'ERROR: AddressSanitizer' "$WORK/run.log" || true)"
