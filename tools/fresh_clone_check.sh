#!/bin/bash
# No GPU needed.  Clone HEAD into a scratch directory, build everything with __graft_entry__.build(),
# compare the SASS of the fresh libxrs.so with the in-tree library (the one the GPU runs loaded; -lineinfo
# embeds source paths, so the files differ byte-wise while the code must not), run the CPU suite there.
#   bash tools/fresh_clone_check.sh [scratch-dir]
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=${1:-/tmp/xrs_clonecheck}
rm -rf "$TMP"
git clone -q "$ROOT" "$TMP"
(cd "$TMP" && python -c "import __graft_entry__ as g; g.build(); print('fresh build ok')")
sass() { cuobjdump -sass "$1" | grep -v "^\s*//" | grep -v "$TMP" | grep -v "$ROOT" | md5sum | cut -d' ' -f1; }
A=$(sass "$TMP/xcube_resampling_b200/libxrs.so")
B=$(sass "$ROOT/xcube_resampling_b200/libxrs.so")
echo "fresh   $A"
echo "in-tree $B"
[ "$A" = "$B" ] && echo "SASS identical" || { echo "SASS DIFFERS: the in-tree libxrs.so is not built from HEAD"; exit 1; }
(cd "$TMP" && python -m pytest tests/ -x -q -m "not gpu" | tail -2)
rm -rf "$TMP"
