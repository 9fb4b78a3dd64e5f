N=$1
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/final_n${N}_$name.json 2> gpurun_out/final_n${N}_$name.err; echo "$name rc=$?"; }
nproc
run default
if [ "$N" = "8" ]; then run e2eT6 --e2e-threads 6 --trials 1; fi
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/final_n${N}_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); c=d['clocks']
            print(f, 'value %.1f M best %.1f M e2e %.1f M frac %.3f' % (d['value']/1e6, d['best']/1e6, d['e2e']['value']/1e6, d['roofline']['frac']), d['trials_ms'], c['sm_mhz_in_kernel_by_trial'], c['power_w'], d.get('exchange_check','')[:40], d['e2e']['api'][-120:])
P
