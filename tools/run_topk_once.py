"""A few entity-scoring calls at one size (the command ncu wraps)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
import modular_prot_b_gan as m
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator().manual_seed(1)
q, t = torch.randn(B, 128, generator=g).to(dev), torch.randn(65536, 128, generator=g).to(dev)
for _ in range(4):
    s, i = m.cosine_topk(q, t, 10)
torch.cuda.synchronize()
eng = m._TOPK_ENGINES[(0, 128)]
print("ok", float(s.sum()), "flagged:", eng.topk_last_flagged(), eng.topk_flag_report)
