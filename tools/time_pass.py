"""Time of one pass alone on the device (CUDA graph of 50 back-to-back passes on one stream):
    [PBG_LIB_PATH=...] python tools/time_pass.py CTAS STAGED B [B ...]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m

dev = torch.device("cuda:0")
ctas, staged = int(sys.argv[1]), int(sys.argv[2])
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
ZERO = int(os.environ.get("TIME_ZERO", "0"))   # 1: all-zero weights and tables (does the MMA rate depend on the data?)
if ZERO:
    with torch.no_grad():
        for p_ in list(G.parameters()) + list(D.parameters()):
            p_.zero_()
eng = m.make_fused_engine(G.to(dev), D.to(dev), ctas=ctas)
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
if ZERO:
    node_emb.zero_(); rel_w.zero_()
for B in [int(x) for x in sys.argv[3:]]:
    trip, z = synth.make_triplets(B).to(dev), synth.make_latents(B).to(dev)
    if ZERO:
        z.zero_()
    out = {"gen_out": torch.empty(B, 128, dtype=torch.bfloat16, device=dev), "gen_scores": torch.empty(B, device=dev),
           "logits": torch.empty(B, device=dev), "probs": torch.empty(B, device=dev)}
    kw = dict(want_gen_out=True, want_gen_scores=True, want_disc=True, out_dtype=torch.bfloat16, out=out)
    eng.reserve(B, "bf16", 2)
    if staged:
        eng.stage_triplets(0, node_emb, rel_w, trip, z)

    def run():
        if staged:
            eng.score_staged(0, **kw)          # the staging kernel is not part of this time
        else:
            eng.score_triplets(node_emb, rel_w, trip, z, precision="bf16", **kw)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    g, side, n = torch.cuda.CUDAGraph(), torch.cuda.Stream(dev), 50
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                run()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    print(f"{os.environ.get('PBG_LIB_PATH', 'default'):>48s}  ctas={ctas or 'all'} staged={staged} zero={ZERO} B={B}: {best:7.2f} us/pass = "
          f"{B * 4850688 / best / 1e6:6.0f} TFLOP/s", flush=True)
