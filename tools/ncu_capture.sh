#!/bin/bash
# One `ncu --set full` capture on the GPU box, summarised there (the .ncu-rep files are too large to
# travel back): gpurun -- 'bash tools/ncu_capture.sh <name> <kernel regex> <skip> <count> <object.o> <mangled kernel> <cmd...>'
# Runs <cmd> once without ncu first (it must exit 0), then under ncu; writes into gpurun_out/:
#   <name>_ncu_full.txt  per-launch metrics (tools/ncu_summary.py)
#   <name>_by_line.txt   instructions / stall samples per source line (tools/ncu_by_line.py)
#   <name>_opcodes.txt   warp instructions per SASS opcode (tools/ncu_source_hot.py)
set -u
NAME=$1; REGEX=$2; SKIP=$3; COUNT=$4; OBJ=$5; KERNEL=$6; shift 6
OUT=gpurun_out
mkdir -p $OUT
timeout 300 "$@" > $OUT/${NAME}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${NAME}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -s "$SKIP" -c "$COUNT" -f -o $OUT/$NAME "$@" > $OUT/${NAME}_ncu.log 2>&1
if [ -f $OUT/$NAME.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/$NAME.ncu-rep $OUT/${NAME}_ncu_full.txt > /dev/null 2>> $OUT/${NAME}_ncu.log
    python tools/ncu_by_line.py $OUT/$NAME.ncu-rep "$OBJ" "$KERNEL" 45 > $OUT/${NAME}_by_line.txt 2>> $OUT/${NAME}_ncu.log
    python tools/ncu_source_hot.py $OUT/$NAME.ncu-rep > $OUT/${NAME}_opcodes.txt 2>> $OUT/${NAME}_ncu.log
    rm -f $OUT/$NAME.ncu-rep
    echo "$NAME: captured"
else
    echo "$NAME: no report"; tail -5 $OUT/${NAME}_ncu.log
fi
