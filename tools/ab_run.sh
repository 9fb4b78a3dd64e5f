python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo rc=$?
for sa in 2 0; do python bench.py --config wide --steps 20 --warmup 5 --stage-ahead $sa --no-cpu-baseline --no-library-baseline > gpurun_out/s8_wide_sa$sa.json 2> gpurun_out/s8_wide_sa$sa.err; echo rc=$?; done
python bench.py --config wide --per-gpu-batch 8192 --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/s8_wide8192.json 2> gpurun_out/s8_wide8192.err; echo rc=$?
python - <<'P'
import json
for f in ('s8_bench','s8_wide_sa2','s8_wide_sa0','s8_wide8192'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); c=d['clocks']; r=d['roofline']
        print(f, 'value %.2f M best %.2f M e2e %.2f M frac %.3f burst %.3f sust %.3f' % (d['value']/1e6, d['best']/1e6, d['e2e']['value']/1e6, r['frac'], r['frac_of_burst'], r['frac_of_sustained']), d['trials_ms'], c['sm_mhz_in_kernel_by_trial'], c['power_w'])
    except Exception as e:
        print(f, 'failed', e)
P
tail -3 gpurun_out/s8_wide_sa2.err
