"""Per-CTA clock64 breakdown of the tensor-core Linear kernels (run on the B200 box).  Writes gpurun_out/trace.log."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch

from pbg import synth
import modular_prot_b_gan as m

out = open(ROOT / "gpurun_out" / "trace.log", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); out.write(s + "\n"); out.flush()


dev = torch.device("cuda:0")
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
eng = m.make_fused_engine(G.to(dev), D.to(dev))
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cases = [("G.L0", 0, 0, 320), ("G.L1", 0, 1, 1024), ("G.L2", 0, 2, 1024), ("D.L0", 1, 0, 384), ("D.L1", 1, 1, 1024)]
for name, model, layer, K in cases:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    for _ in range(3):
        eng.linear_bf16(model, layer, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.linear_bf16(model, layer, a)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    eng.debug_trace(True)
    eng.linear_bf16(model, layer, a)
    t = eng.debug_trace(False).double()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0]
    P(f"{name} M={M} K={K}: {us:.2f} us/launch (back-to-back), {t.shape[0]} CTAs traced")
    P(f"   total clk/CTA (epi end - begin)    : mean {(t[:, 8] - t0).mean():.0f}  max {(t[:, 8] - t0).max():.0f}")
    P(f"   first operands landed after         : mean {(t[:, 5] - t0).mean():.0f}")
    P(f"   mma loop (first full -> last commit): mean {(t[:, 6] - t[:, 5]).mean():.0f}  = {((t[:, 6] - t[:, 5]) / t[:, 9]).mean():.0f} clk per k-block ({t[:, 9].mean():.0f} k-blocks)")
    P(f"   mma thread waiting on full barriers : mean {t[:, 3].mean():.0f}   on tmem_empty {t[:, 4].mean():.0f}")
    P(f"   producer waiting on empty barriers  : mean {t[:, 1].mean():.0f}   producer done at {(t[:, 2] - t0).mean():.0f}")
    P(f"   epilogue: last acc ready at {(t[:, 10] - t0).mean():.0f}, wait sum {t[:, 7].mean():.0f}, epilogue tail {(t[:, 8] - t[:, 10]).mean():.0f}")
