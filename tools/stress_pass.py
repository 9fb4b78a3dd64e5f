"""Repeated fused passes over a mix of batch sizes; every result is compared with the first run of its size
(the pass is deterministic), so protocol races show up as hangs (the in-kernel hang guard traps) or mismatches."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sizes = [int(x) for x in sys.argv[2:]] or [1, 16, 100, 256, 1000, 4096, 5000, 8192, 32768]
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
eng = m.make_fused_engine(G.to(dev), D.to(dev))
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
ref = {}
for rep in range(reps):
    for B in sizes:
        trip, z = synth.make_triplets(B).to(dev), synth.make_latents(B).to(dev)
        res = eng.score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True, want_disc=True,
                                 precision="bf16", out_dtype=torch.bfloat16)
        torch.cuda.synchronize()
        cur = {k: v.clone() for k, v in res.items() if torch.is_tensor(v)}
        if B not in ref:
            ref[B] = cur
        else:
            for k, v in cur.items():
                if not torch.equal(v, ref[B][k]):
                    print(f"MISMATCH rep {rep} B {B} {k}: max diff {(v.float() - ref[B][k].float()).abs().max().item():.3e}", flush=True)
    print("rep", rep, "ok", flush=True)
eng.check_indices()
print("stress done")
