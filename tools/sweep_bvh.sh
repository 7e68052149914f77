#!/usr/bin/env bash
# Device A/B of the BVH kernel variants on the BVH configs (kernel ms, node visits, sphere-test slots):
#   gpurun --timeout 600 -- 'bash tools/sweep_bvh.sh "<knob set>" "<knob set>" ...'    (a knob set is "NAME=value NAME=value")
# Results: gpurun_out/sweep_bvh.jsonl (one {"variant": ...} line, then one line per config).
set -u
mkdir -p gpurun_out
out=gpurun_out/sweep_bvh.jsonl
: > "$out"
configs="${SWEEP_CONFIGS:-c3 c4 c5slice}"
for knobs in "$@"; do
    echo "{\"variant\": \"$knobs\"}" >> "$out"
    env $knobs timeout 300 python tests/tools/run_configs.py $configs --reps 3 >> "$out" 2>&1
done
python - <<'PY'
import json
v = None
for line in open("gpurun_out/sweep_bvh.jsonl"):
    try:
        d = json.loads(line)
    except Exception:
        print(line.rstrip()[:200]); continue
    if "variant" in d:
        v = d["variant"]; continue
    print(f'{v:58s} {d["config"]:8s} {d["kernel_ms"]:9.3f} ms  nodes/seg {d["node_visits"]/max(d["segments"],1):6.3f}  slots/seg {d["sphere_tests"]/max(d["segments"],1):6.3f}')
PY
