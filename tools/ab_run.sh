#!/bin/bash
# developer tool, runs ON the GPU box (under gpurun): A/B of every library under csrc/gpurun_variants/ against the
# default build -- device time of configs 4 / 2 / 3 (tools/profile_target.py) and, with --parity, the FP32 parity
# tests with the variant loaded through --lib / --aai-lib.
#   tools/build_variant.sh row2 -DAAI_EXP_ROW2=1; tools/build_variant.sh ry -DAAI_EXP_RY_INC=1; ...
#   gpurun --timeout 400 -- 'tools/ab_run.sh --parity > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log'
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
parity=0; [ "$1" = "--parity" ] && parity=1
run() {  # $1 = label, $2 = library path ('' = default build)
    for c in 4 2 3; do
        echo -n "$1 "
        timeout 60 python tools/profile_target.py --config $c --arith f32 --steps 50 ${2:+--lib $2} | tail -1
    done
    if [ $parity = 1 ] && [ -n "$2" ]; then
        echo -n "$1 parity: "
        timeout 200 python -m pytest tests --aai-lib $2 -m gpu -q -x -k "f32 or random or full_size or batch" 2>&1 | tail -1
    fi
}
run default ""
for so in area_average_interpolation_b200/csrc/gpurun_variants/*.so; do
    [ -e "$so" ] || continue
    run "$(basename "$so" .so)" "$root/$so"
done
run default ""
