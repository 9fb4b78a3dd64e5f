# DRAM traffic per pass launch in rotation (profiles/dram_steady_r2.txt): fused gather and the stage-next form, discards on / off
for mode in "" stage2; do for d in 0 1; do
PBG_DISCARD=$d ncu --cache-control none -k regex:pbg_pass2 --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/dram_steady_${mode:-fused}_discard$d.csv python tools/dram_steady.py 6 8 4096 $mode > gpurun_out/dram_steady_${mode:-fused}_$d.log 2>&1
tail -1 gpurun_out/dram_steady_${mode:-fused}_$d.log
python tools/dram_steady.py --summarise gpurun_out/dram_steady_${mode:-fused}_discard$d.csv
done; done
