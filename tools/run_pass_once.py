"""A few fused passes at one batch size (the command ncu wraps)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
CTAS = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # SMs per pass (0 = all)
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
eng = m.make_fused_engine(G.to(dev), D.to(dev), ctas=CTAS)
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
trip, z = synth.make_triplets(B).to(dev), synth.make_latents(B).to(dev)
for _ in range(6):
    res = eng.score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True, want_disc=True,
                             precision="bf16", out_dtype=torch.bfloat16)
torch.cuda.synchronize()
eng.check_indices()
print("ok", float(res["logits"].sum()))
