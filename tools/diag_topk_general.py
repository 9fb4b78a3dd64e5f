"""The general top-k path (exact fp32 SIMT GEMM over row chunks + one selection CTA per row): where does its time go?
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/diag_topk_general.py E K B"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
import torch.nn.functional as F
import modular_prot_b_gan as m
dev = torch.device("cuda:0")
E, k, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(5)
t, q = torch.randn(65536, E, generator=g).to(dev), torch.randn(B, E, generator=g).to(dev)
for _ in range(2):
    s, i = m.cosine_topk(q, t, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s, i = m.cosine_topk(q, t, k); e1.record(); torch.cuda.synchronize()
ours = e0.elapsed_time(e1)
fn = lambda: torch.matmul(F.normalize(q, dim=-1), F.normalize(t, dim=-1).T).topk(k, dim=1)
fn(); torch.cuda.synchronize()
e0.record(); ts, ti = fn(); e1.record(); torch.cuda.synchronize()
print(f"E={E} k={k} B={B}: this library {ours * 1e3:.0f} us, torch lines {e0.elapsed_time(e1) * 1e3:.0f} us, max |score diff| {float((s - ts).abs().max()):.1e}")
