#!/bin/bash
# GPU box: sweep the k_anyhit tuning knobs on the 4K c4_open frame (class times from rt580_main --bench)
D=$(python -c "import bench; print(bench.scene_dir('c4_open'))")
for cfg in "24 8 12" "24 4 12" "24 2 12" "24 0 12" "12 4 12" "48 4 12" "24 6 16" "24 4 16"; do
  set -- $cfg
  RT580_AH_STEPS=$1 RT580_AH_MIN_SEARCH=$2 RT580_AH_BLOCKS_PER_SM=$3 ./580-raytracer_b200/rt580_main c4_open.json 3840 2160 /tmp/o.ppm $D 16 4 --bench 2 --no-ppm 2>/dev/null | grep "ao_tree\|shadow_tree\|^mean" | tr '\n' ' ' | sed "s/^/steps $1 min $2 bpsm $3 -> /"; echo
done
