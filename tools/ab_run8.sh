N=8
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --trials 1 --min-ms 10 "$@" > gpurun_out/x_n${N}_$name.json 2> gpurun_out/x_n${N}_$name.err; echo "$name rc=$?"; }
export PBG_HOST_SYNC=block
run blockT12 --e2e-threads 12
run blockT8 --e2e-threads 8
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/x_n8_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, 'value %.1f M e2e %.1f M' % (d['value']/1e6, d['e2e']['value']/1e6), d['e2e']['steps'], d['e2e']['seconds'])
P
