#!/bin/bash
# usage: tools/bench_sweep.sh OUT "args1" "args2" ...   -- one bench.py run per argument string, one summary line each
out=$1; shift
for a in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline $a 2>>$out.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('%-58s value %6.1f M  best %6.1f M  burst %.3f  sust %.3f  us/launch %6.1f  e2e %6.1f M  clk %s %s  trials %s' % (sys.argv[1], d['value']/1e6, d['best']/1e6,
      r['whole_step']['frac'], r['whole_step']['frac_of_sustained'], r['us_per_launch'], d['e2e']['value']/1e6, d['clocks']['sm_mhz'], ','.join(d['clocks']['reasons']), d['trials_ms']))" "$a" >> $out
done
cat $out
