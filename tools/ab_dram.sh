set -x
for d in 0 1; do
PBG_DISCARD=$d ncu --cache-control none -k regex:pbg_pass2 --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/dram_steady_discard$d.csv python tools/dram_steady.py 6 8 4096 > gpurun_out/dram_steady_$d.log 2>&1
tail -2 gpurun_out/dram_steady_$d.log
python tools/dram_steady.py --summarise gpurun_out/dram_steady_discard$d.csv
done
