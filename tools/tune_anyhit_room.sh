D=$(python -c "import bench; print(bench.scene_dir('c4_room'))")
for cfg in "48 16 12" "24 6 12" "24 8 12" "48 8 12"; do
  set -- $cfg
  RT580_AH_STEPS=$1 RT580_AH_MIN_SEARCH=$2 RT580_AH_BLOCKS_PER_SM=$3 ./580-raytracer_b200/rt580_main c4_room.json 3840 2160 /tmp/o.ppm $D 16 4 --bench 3 --no-ppm 2>/dev/null | grep "ao_tree\|shadow_tree\|^mean" | tr '\n' ' ' | sed "s/ \+/ /g" | sed "s/^/steps $1 min $2 bpsm $3 -> /"; echo
done
