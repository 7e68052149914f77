#!/usr/bin/env bash
# Device A/B of the scan-kernel variants on the small configs: gpurun -- 'bash tools/sweep_scan.sh "<knob set>" ...'
set -u
mkdir -p gpurun_out
out=gpurun_out/sweep_scan.jsonl
: > "$out"
for knobs in "$@"; do
    echo "{\"variant\": \"$knobs\"}" >> "$out"
    env $knobs timeout 300 python tests/tools/run_configs.py ${SWEEP_CONFIGS:-c1 c2 c2mg} --reps 5 >> "$out" 2>&1
done
python - <<'PY'
import json
v = None
for line in open("gpurun_out/sweep_scan.jsonl"):
    try:
        d = json.loads(line)
    except Exception:
        print(line.rstrip()[:200]); continue
    if "variant" in d:
        v = d["variant"]; continue
    print(f'{v:40s} {d["config"]:8s} {d["kernel_ms"]:8.3f} ms   e2e {d["e2e_ms"]:8.3f} ms  launches {d["launches"]}')
PY
