python -m pytest tests/test_gpu_topk.py tests/test_gpu_inference.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_topk.py 2>&1 | tail -11
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; echo rc=$?
python - <<'P'
import json
d=json.load(open('gpurun_out/s10_bench.json')); c=d['clocks']; r=d['roofline']
print('value %.2f M best %.2f M e2e %.2f M frac %.3f' % (d['value']/1e6, d['best']/1e6, d['e2e']['value']/1e6, r['frac']), d['trials_ms'], c)
print({k:v for k,v in d['gpu_library_baseline'].items() if k!='what'})
print(r.get('library_gemms_only_tflops'), r.get('frac_of_library_gemms_only'))
P
