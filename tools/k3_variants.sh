#!/bin/bash
# K3 build variants side by side in ONE gpurun call (round 2, second session):
#   gpurun --timeout 900 -- 'bash tools/k3_variants.sh <tag> <variant> [<variant> ...]'
# For the product library ("main") and every libxrs_<variant>.so: CUDA-event kernel times of C3 and C5
# (tools/bench_configs.py, no profiler); then the reprojection GPU tests + the full-size C3 / C5 tests
# with the product library.
set -u
TAG=${1:-k3v}; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for v in main "$@"; do
    if [ "$v" = main ]; then unset XRS_LIB; else export XRS_LIB=$PWD/xcube_resampling_b200/libxrs_$v.so; fi
    timeout 200 python tools/bench_configs.py c3 c5 --no-cpu --reps 5 > "$OUT/$v.jsonl" 2> "$OUT/$v.err"
    echo "== $v rc=$?" | tee -a "$OUT/summary.txt"
    python - "$OUT/$v.jsonl" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    print(f'{d.get("ms", float("nan")):8.3f} ms  frac {d.get("frac") or 0:.3f}  {d["config"]}' if "ms" in d else d)
PY
done
unset XRS_LIB
timeout 500 python -m pytest tests/test_reproject_gpu.py tests/test_multigpu_gpu.py tests/test_zz_full_size_gpu.py -q -m gpu -k "not c2 and not c4 and not c1" --maxfail=10 > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest.log"
