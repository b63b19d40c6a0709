#!/bin/bash
# Recipe for oracle/_ref: the UNMODIFIED reference, staged where it can travel to the GPU box.
#
# The reference is pure Python (no build step).  /root/reference does not exist on the GPU box, so the files the
# checker needs are staged -- byte for byte, nothing edited -- under oracle/_ref/reference/.  oracle/_ref/ is listed in
# .gitignore (reference sources never enter this repository's history) but not in .gpurunignore, so it ships with the
# snapshot exactly like a compiled oracle/_ref/*.so would.  TEST INFRASTRUCTURE ONLY: tests/, bench.py's
# `--impl reference` / cpu_baseline legs and smoke() may import it (through oracle/ref_loader.py); the product never does.
#
#   usage: oracle/make_ref.sh [REFERENCE_DIR]      (default /root/reference)
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref/reference"
if [ ! -d "$SRC/EigenTrajectory" ]; then
  echo "make_ref: $SRC is not a reference checkout; leaving $DST as it is" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/datasets"
# the path itself (descriptor / normaliser / anchors / k-means / model), its callers (trainer, metrics, loader, config)
# and the predictor plugins that sit behind the hook seam (config 4 runs baseline/sgcn unmodified)
for d in EigenTrajectory utils baseline config; do
  (cd "$SRC" && find "$d" -type f \( -name '*.py' -o -name '*.json' -o -name '*.yml' -o -name '*.yaml' \) -print0) |
    (cd "$SRC" && xargs -0 cp --parents -t "$DST")
done
cp "$SRC/trainval.py" "$SRC/LICENSE" "$DST/"
# config 4 data: zara1 (train + val for the descriptor initialisation, test for the evaluation loop), 2.8 MB of text
cp -r "$SRC/datasets/zara1" "$DST/datasets/zara1"
(cd "$SRC" && find EigenTrajectory utils baseline config trainval.py datasets/zara1 -type f \
   \( -name '*.py' -o -name '*.json' -o -name '*.yml' -o -name '*.yaml' -o -name '*.txt' \) -print0 | sort -z | xargs -0 sha256sum) \
  > "$HERE/_ref/MANIFEST.sha256"
echo "make_ref: staged $(find "$DST" -type f | wc -l) files from $SRC into $DST"
