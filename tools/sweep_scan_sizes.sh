#!/usr/bin/env bash
# Device A/B of the scan kernels over frame sizes and sample counts: where do lanes sharing a pixel (RTCU_SCAN_DIRECT=G, dynamic
# pixel-sized work items) beat the thread-per-pixel grid (first frame of a view: RTCU_TILE_ORDER=0)?
# gpurun -- 'bash tools/sweep_scan_sizes.sh'
set -u
mkdir -p gpurun_out
out=gpurun_out/sweep_scan_sizes.txt
: > "$out"
for cfg in ${SWEEP_CONFIGS:-c1 c2}; do
  for size in ${SWEEP_SIZES:-320x240 640x360 800x600 1280x720 1920x1080 2560x1440 3840x2160}; do
    for spp in ${SWEEP_SPP:-16 30 64 256}; do
      line="$cfg $size spp=$spp"
      for knobs in ${SWEEP_KNOBS:-RTCU_TILE_ORDER=0 RTCU_SCAN_DIRECT=4 RTCU_SCAN_DIRECT=8 RTCU_SCAN_DIRECT=16}; do
        ms=$(env $knobs timeout 120 python tests/tools/run_configs.py $cfg --reps 3 --size $size --spp $spp 2>/dev/null | python -c 'import json,sys; print(json.loads(sys.stdin.readline())["kernel_ms"])')
        line="$line  $knobs=$ms"
      done
      echo "$line" | tee -a "$out"
    done
  done
done
