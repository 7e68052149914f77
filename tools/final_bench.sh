set -u
mkdir -p gpurun_out
for c in "c1 400" "c2 200" "c3 40" "c4 20"; do
  set -- $c
  timeout 600 python bench.py --config $1 --steps $2 --warmup 5 --no-configs > gpurun_out/bench_r2b_$1_n1.json 2> gpurun_out/bench_r2b_$1.err
  python -c "import json;d=json.load(open('gpurun_out/bench_r2b_$1_n1.json'));print('$1',d['ms_per_step'],d['value'],d['e2e']['value'],d.get('parity_max_lsb'),d['roofline']['frac'],d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)"
done
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/bench_r2b_c5_n1.json 2> gpurun_out/bench_r2b_c5.err
grep "Elapsed (wall" gpurun_out/bench_r2b_c5.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2b_c5_n1.json'));print('c5',d['ms_per_step'],d['value'],d['e2e']['value'],d.get('parity_max_lsb'),d['roofline']['frac'],d['cpu_baseline']['value'], {k:v['ms_per_step'] for k,v in d['configs'].items()})"
