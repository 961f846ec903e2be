#!/bin/bash
# developer tool, runs ON the GPU box: A/B of the fast-mode gather kernel with / without the skipped loads (variant library
# built with tools/build_variant.sh noskip -DAAI_FAST_SKIP_LOADS=0), then the fast-mode GPU tests
cd "$(dirname "$0")/.."
V=area_average_interpolation_b200/csrc/gpurun_variants/noskip.so
for i in 1 2; do
    timeout 100 python tools/dev_bin.py --skip-checks --time 2>&1 | grep gather | head -1
    timeout 100 python tools/dev_bin.py --skip-checks --time --lib $V 2>&1 | grep gather | head -1 | sed "s/gather/gather (no skip)/"
done
for c in 2 3; do
    timeout 60 python tools/profile_target.py --config $c --mode 2 --arith f32 --steps 100
    timeout 60 python tools/profile_target.py --config $c --mode 2 --arith f32 --steps 100 --lib $V | sed "s/$/ (no skip)/"
done
timeout 400 python -m pytest tests -x -q -m gpu -k "fast or staged or golden or batch" 2>&1 | tail -3
