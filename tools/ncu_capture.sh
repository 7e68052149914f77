#!/usr/bin/env bash
# One `ncu --set full` capture of the dominant kernel of each BASELINE config (first frame of tests/tools/run_configs.py; C5 = one
# GPU's 512-sample share at N = 8), after the same command has run once without ncu.  One GPU:
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r2g'
# then here: python tools/kernel_facts.py <git sha> c1=gpurun_out/r2g_c1.ncu-rep ... c5=gpurun_out/r2g_c5n8.ncu-rep
set -u
tag=${1:-cap}
mkdir -p gpurun_out
for c in ${CAPTURE_CONFIGS:-c1 c2 c3 c4 c5n8}; do
    timeout 300 python tests/tools/run_configs.py $c --reps 1 > gpurun_out/${tag}_${c}_plain.log 2>&1 || { echo "$c: plain run failed"; continue; }
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -c 1 -f -o gpurun_out/${tag}_${c} \
        python tests/tools/run_configs.py $c --reps 1 > gpurun_out/${tag}_${c}_ncu.log 2>&1
    echo "$c: $(grep -c '==PROF==' gpurun_out/${tag}_${c}_ncu.log) profiler lines, $(ls -la gpurun_out/${tag}_${c}.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
