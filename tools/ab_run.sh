set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo rc=$?
for T in 9 12; do python bench.py --steps 20 --warmup 5 --e2e-threads $T --trials 2 --no-cpu-baseline --no-library-baseline > gpurun_out/s3_e2e_T$T.json 2>/dev/null; done
python - <<'P'
import json
for f in ('s3_bench','s3_e2e_T9','s3_e2e_T12'):
    d=json.load(open(f'gpurun_out/{f}.json')); print(f, 'value %.1f M best %.1f M e2e %.1f M clk %s' % (d['value']/1e6, d['best']/1e6, d['e2e']['value']/1e6, d['clocks']['sm_mhz']), d['trials_ms'], d['roofline']['traffic'])
P
