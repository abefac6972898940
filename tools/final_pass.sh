#!/bin/bash
# Round-end measurement pass on the B200 box: tests, smoke, bench lines, ncu launch list and
# --set full captures (summarised to text; the .ncu-rep files are deleted before gpurun merges
# gpurun_out/ back).  Usage: gpurun --timeout 1500 -- 'bash tools/final_pass.sh <tag>'
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
B="python bench.py"
NCU="ncu --clock-control none"

python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/status.txt"

$B --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "ref rc=$?" | tee -a "$OUT/status.txt"
$B --steps 5 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/status.txt"
$B --steps 5 --warmup 3 --two-step --no-e2e --no-cpu > "$OUT/bench_two_step.json" 2> "$OUT/bench_two_step.err"
python tools/bench_configs.py c1 c3 c4 c5 > "$OUT/configs_kernel_times.jsonl" 2> "$OUT/configs.err"; echo "configs rc=$?" | tee -a "$OUT/status.txt"

# launch list of the bench command (per-launch gpu__time_duration.sum)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$OUT/bench_launches.csv" \
    $B --steps 2 --warmup 3 --no-e2e --no-cpu > "$OUT/ncu_launches.log" 2>&1

cap() {  # name, kernel regex, count, command...
    local name=$1 regex=$2 count=$3; shift 3
    $NCU --set full --import-source on -k "regex:$regex" -c "$count" -f -o "$OUT/$name" "$@" > "$OUT/ncu_$name.log" 2>&1
    if [ -f "$OUT/$name.ncu-rep" ]; then
        python tools/ncu_summary.py "$OUT/$name.ncu-rep" "$OUT/${name}_ncu_full.txt" >> "$OUT/ncu_$name.log" 2>&1
        if [ "$name" = k0_k1 ]; then
            python tools/ncu_by_line.py "$OUT/$name.ncu-rep" xcube_resampling_b200/_obj/rectify_ij.o _ZN3xrs10k1_scatterENS_6IjGeomE 40 "k1_scatter(" > "$OUT/k1_scatter_by_line.txt" 2>> "$OUT/ncu_$name.log"
        fi
        rm -f "$OUT/$name.ncu-rep"
    fi
}
cap k0_k1 'k1_|k0_tile' 5 $B --steps 1 --warmup 1 --no-e2e --no-cpu
cap k2_fused 'k2_gather' 2 $B --steps 1 --warmup 1 --no-e2e --no-cpu
cap k2_two_step 'k2_gather|k1_resolve' 4 $B --steps 1 --warmup 1 --no-e2e --no-cpu --two-step
cap k3_c3 'k3_reproject' 3 python tools/bench_configs.py c3 --one-shot
cap k5_c1_c4 'k5_|k4_' 16 python tools/bench_configs.py c1 c4 --one-shot
ls -la "$OUT"
cat "$OUT/status.txt"
