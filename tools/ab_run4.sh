bash tools/ab_run.sh 4
nvidia-smi topo -m 2>&1 | head -16
echo "--- H2D on 4 GPUs at once (tools/pcie_bw.py, one process per GPU)"
for g in 0 1 2 3; do CUDA_VISIBLE_DEVICES=$g python tools/pcie_bw.py > gpurun_out/pcie_g$g.log 2>&1 & done; wait
for g in 0 1 2 3; do echo "GPU $g:"; grep "1 stream" gpurun_out/pcie_g$g.log; done
echo "--- H2D on GPU 0 alone"
CUDA_VISIBLE_DEVICES=0 python tools/pcie_bw.py | grep "1 stream"
