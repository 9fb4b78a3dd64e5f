"""BASELINE configs[4] (width-scaled model: E = 256, H = 4096) on one GPU: time per pass and TFLOP/s of the pair kernel."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m

dev = torch.device("cuda:0")
E, Z, H, N = 256, 64, 4096, 65536
FLOP = 2 * ((2 * E + Z) * H + H * H + H * E) + 2 * (3 * E * H + H * (H // 2) + H // 2)   # 63 442 944 per sample
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator, E, Z, H, H)
eng = m.make_fused_engine(G.to(dev), D.to(dev))
node_emb, rel_w = (t.to(dev) for t in synth.make_tables(N, 64, E))
for B in (1024, 4096, 8192):
    trip, z = synth.make_triplets(B, N, 64).to(dev), synth.make_latents(B, Z).to(dev)
    out = {"gen_out": torch.empty(B, E, dtype=torch.bfloat16, device=dev), "gen_scores": torch.empty(B, device=dev),
           "logits": torch.empty(B, device=dev), "probs": torch.empty(B, device=dev)}

    def run():
        eng.score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True, want_disc=True,
                           precision="bf16", out_dtype=torch.bfloat16, out=out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"wide model (E={E}, H={H}) B={B}: {us:.1f} us/pass = {B / us:.2f} M samples/s = {B * FLOP / us / 1e6:.0f} TFLOP/s "
          f"({B * FLOP / us / 1e6 / 1644.4:.2f} of the measured burst peak)", flush=True)
