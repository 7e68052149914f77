#!/usr/bin/env bash
# Measurements that were prepared on the CPU and are waiting for a device.  One GPU, about two minutes:
#   gpurun --timeout 300 -- 'bash tools/measure_pending.sh'
# Results land in gpurun_out/pending_*.jsonl (copy what is worth keeping into profiles/).
set -u
mkdir -p gpurun_out
out=gpurun_out/pending_builder_variants.jsonl
: > "$out"
# 1. builder / collapse variants (bvh.h, pack_bvh4): kernel ms, node visits and sphere-test slots on the BVH configs
for knobs in "A=0" "RTCU_BVH_LEAF_COST=1" "RTCU_BVH_COLLAPSE=sah" "RTCU_BVH_COLLAPSE=sah RTCU_BVH_LEAF_COST=1" "RTCU_BVH_SWEEP=512" \
             "RTCU_BVH_SWEEP=512 RTCU_BVH_COLLAPSE=sah"; do
    echo "{\"variant\": \"$knobs\"}" >> "$out"
    env $knobs timeout 120 python tests/tools/run_configs.py c3 c4 c5slice --reps 3 >> "$out" 2>&1
done
# 2. what a scene change costs per builder thread count (the final library)
timeout 60 python tools/time_scene_upload.py 1 4 8 16 > gpurun_out/pending_scene_upload.jsonl 2>&1
# 3. parity of every variant on the device: the BVH tests under each knob set (the CPU replay already passes for all of them)
for knobs in "RTCU_BVH_LEAF_COST=1" "RTCU_BVH_COLLAPSE=sah" "RTCU_BVH_SWEEP=512 RTCU_BVH_COLLAPSE=sah RTCU_BVH_LEAF_COST=1"; do
    echo "== $knobs" >> gpurun_out/pending_variant_parity.log
    env $knobs timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "bvh" 2>&1 | tail -2 >> gpurun_out/pending_variant_parity.log
done
tail -n 40 "$out"
cat gpurun_out/pending_variant_parity.log
