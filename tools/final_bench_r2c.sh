# evidence refresh after a BVH-kernel change: GPU tests, C3 / C4 headline runs, the default bench line, ncu captures of the BVH configs
#   gpurun --timeout 900 -- 'bash tools/final_bench_r2c.sh <tag>'
set -u
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; tail -2 gpurun_out/${tag}_gputests.log
for c in "c3 40" "c4 20"; do
  set -- $c
  timeout 300 python bench.py --config $1 --steps $2 --warmup 5 --no-configs > gpurun_out/bench_${tag}_$1_n1.json 2> gpurun_out/bench_${tag}_$1.err
  python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_$1_n1.json'));print('$1',d['ms_per_step'],d['value'],d['e2e']['value'],d.get('parity_max_lsb'),d['roofline']['frac'])"
done
SECONDS=0; timeout 600 python bench.py > gpurun_out/bench_${tag}_c5_n1.json 2> gpurun_out/bench_${tag}_c5.err
echo "default bench wall: ${SECONDS} s"
python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_c5_n1.json'));print('c5',d['ms_per_step'],d['value'],d['e2e']['value'],d.get('parity_max_lsb'),d['roofline']['frac'],d['cpu_baseline']['value'], {k:v['ms_per_step'] for k,v in d['configs'].items()})"
CAPTURE_CONFIGS="c3 c4 c5n8" bash tools/ncu_capture.sh ${tag}
