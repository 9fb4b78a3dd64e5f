"""Smoke-sized passes for compute-sanitizer (memcheck / racecheck / synccheck), SURVEY.md:244:
    compute-sanitizer --tool memcheck python tools/sanitize_pass.py
B = 16 and 256 in both precisions, a ragged bf16 batch, a staged request, result mirrors and one top-k call; every
result is also checked against the first run of its inputs, so a tool that perturbs timing still verifies values."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m

dev = torch.device("cuda:0")
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
eng = m.make_fused_engine(G.to(dev), D.to(dev), ctas=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
node_emb, rel_w = (t.to(dev) for t in synth.make_tables(num_entities=4096))
kw = dict(want_gen_out=True, want_gen_scores=True, want_disc=True)
for B in (16, 256, 300):
    trip, z = synth.make_triplets(B, num_entities=4096).to(dev), synth.make_latents(B).to(dev)
    for prec in ("fp32", "bf16"):
        a = eng.score_triplets(node_emb, rel_w, trip, z, precision=prec, **kw)
        b = eng.score_triplets(node_emb, rel_w, trip, z, precision=prec, **kw)
        eng.check_indices()
        assert all(torch.equal(a[k], b[k]) for k in a), (B, prec)
        print(f"B={B} {prec}: ok, logits sum {float(a['logits'].sum()):.4f}", flush=True)
    eng.stage_triplets(0, node_emb, rel_w, trip, z)
    c = eng.score_staged(0, **kw)
    eng.check_indices()
    assert all(torch.equal(a[k], c[k]) for k in a), (B, "staged")
    mir = {k: torch.zeros_like(v) for k, v in a.items()}
    eng.set_result_mirrors(**{k: [v.data_ptr()] for k, v in mir.items()})
    d = eng.score_triplets(node_emb, rel_w, trip, z, precision="bf16", **kw)
    torch.cuda.synchronize()
    eng.set_result_mirrors()
    assert all(torch.equal(mir[k], d[k]) for k in d), (B, "mirror")
    print(f"B={B} staged + mirror: ok", flush=True)
q = torch.randn(64, 128, generator=torch.Generator().manual_seed(5)).to(dev)
s, i = m.cosine_topk(q, node_emb, 10)
torch.cuda.synchronize()
print("topk ok", float(s.sum()), int(i.sum()), flush=True)
print("sanitize_pass done")
