set -u
mkdir -p gpurun_out
for c in "c1 400" "c2 200"; do
  set -- $c
  timeout 300 python bench.py --config $1 --steps $2 --warmup 5 --no-configs > gpurun_out/bench_r2i_$1_n1.json 2> gpurun_out/bench_r2i_$1.err
  python -c "import json;d=json.load(open('gpurun_out/bench_r2i_$1_n1.json'));print('$1',d['ms_per_step'],d['value'],d['e2e']['value'],d.get('parity_max_lsb'),d['roofline']['frac'],d['cpu_baseline']['value'])"
done
CAPTURE_CONFIGS="c1 c2" bash tools/ncu_capture.sh r2i
