set -x
./tools/ubench_discard 0 8 4; ./tools/ubench_discard 1 8 4
ncu --cache-control none --metrics dram__bytes_write.sum,dram__bytes_read.sum --csv --log-file gpurun_out/discard_ncu_mode0.csv ./tools/ubench_discard 0 8 2 > /dev/null
ncu --cache-control none --metrics dram__bytes_write.sum,dram__bytes_read.sum --csv --log-file gpurun_out/discard_ncu_mode1.csv ./tools/ubench_discard 1 8 2 > /dev/null
PBG_DISCARD=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_inference.py -m gpu -x -q 2>&1 | tail -3
for d in 0 1 0 1; do PBG_DISCARD=$d python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/ab_discard_$d.json 2>/dev/null; python - <<P
import json; d=json.load(open('gpurun_out/ab_discard_$d.json')); print('discard=$d value %.1f M best %.1f M e2e %.1f M clk %s trials %s' % (d['value']/1e6, d['best']/1e6, d['e2e']['value']/1e6, d['clocks']['sm_mhz'], d['trials_ms']))
P
done
