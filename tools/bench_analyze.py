"""analyze_relations: one batched discriminator pass vs the reference's per-(head, tail, relation) call + .item() loop
(pro_b_gan_infer.py:290-318), both over the CUDA modules on one GPU.  Appends to gpurun_out/bench_analyze.log."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from torch import nn

from pbg import synth
from pbg.analyze import analyze_relations_batched
import modular_prot_b_gan as m

dev = torch.device("cuda:0")
_, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
D = D.to(dev).eval()
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
rel = nn.Embedding.from_pretrained(rel_w)
out = open(ROOT / "gpurun_out" / "bench_analyze.log", "a")
for H, T in ((4, 4), (32, 32), (128, 128)):
    heads, tails = list(range(10, 10 + H)), list(range(500, 500 + T))
    for _ in range(3):
        analyze_relations_batched(D, node_emb, rel, heads, tails, 5)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        analyze_relations_batched(D, node_emb, rel, heads, tails, 5)
    torch.cuda.synchronize(); tb = (time.perf_counter() - t0) / 5
    # the reference's loop, bounded to at most 4 x 4 pairs and scaled
    hs, ts = heads[:4], tails[:4]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        for h in hs:
            for t in ts:
                he, te = node_emb[h:h + 1], node_emb[t:t + 1]
                for r in range(rel_w.shape[0]):
                    s = D(he, rel(torch.tensor([r], device=dev)), te).item()
                    torch.sigmoid(torch.tensor(s)).item()
    tl = (time.perf_counter() - t0) * (H * T) / (len(hs) * len(ts))
    line = (f"analyze_relations {H}x{T} pairs x {rel_w.shape[0]} relations = {H * T * rel_w.shape[0]} triplets: "
            f"batched {tb * 1e3:.2f} ms (host dict building included), per-call loop {tl * 1e3:.0f} ms"
            f"{' (scaled from 4x4)' if H * T > 16 else ''} -> {tl / tb:.0f}x")
    print(line, flush=True); out.write(line + "\n")
