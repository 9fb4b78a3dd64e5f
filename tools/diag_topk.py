"""Time and flag report of cosine_topk for a few (B, k) at N = 65536 (diagnostics)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
import modular_prot_b_gan as m
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator().manual_seed(3)
t = torch.randn(N, 128, generator=g).to(dev)
for B, k in ((256, 10), (256, 17), (256, 32), (256, 64), (1024, 64), (4096, 17), (4096, 40), (4096, 64), (16384, 40)):
    q = torch.randn(B, 128, generator=g).to(dev)
    for _ in range(3):
        m.cosine_topk(q, t, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.cosine_topk(q, t, k)
    e1.record(); torch.cuda.synchronize()
    eng = m._TOPK_ENGINES[(0, 128)]
    n = eng.topk_last_flagged()
    print(f"B={B:6d} k={k:3d}: {e0.elapsed_time(e1) / 10 * 1e3:9.1f} us   flagged {n}: {eng.topk_flag_report}", flush=True)
