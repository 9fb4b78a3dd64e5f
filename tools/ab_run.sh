python -m pytest tests/test_gpu_topk.py tests/test_gpu_inference.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
python tools/diag_topk.py 2>&1 | tail -9 | cut -c1-60
python tools/bench_topk.py 2>&1 | tail -10
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_topk_r2d.csv python tools/run_topk_once.py > /dev/null 2>&1
