"""DRAM traffic of the pass in rotation (the command ncu wraps):
    ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv \
        python tools/dram_steady.py [LANES=6] [ROUNDS=8] [B=4096] [stage2]
LANES engines (one workspace each, like bench.py's lanes) take turns over distinct input batches.  A single cold launch
shows 10 MB of DRAM traffic (inputs, outputs, cold weights: profiles/ncu_pass2_*.txt) because its dirty activation lines
are still in L2 when it ends; in rotation the lanes' workspaces (6 x 30 MB) push each other's dead activations out of
the 126 MB L2, which is where the write-backs show -- and what PBG_DISCARD=1 removes.
Summarise with:  python tools/dram_steady.py --summarise X.csv"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]

if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    import csv
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in csv.reader(open(sys.argv[2])):
        if len(r) > 10 and r[0].isdigit() and "pbg_pass2_kernel" in r[4]:
            per.setdefault(int(r[0]), {})[r[-3]] = float(r[-1].replace(",", "")) * mult[r[-2]]
    ids = sorted(per)
    tail = ids[len(ids) // 2:]   # the second half of the launches: every workspace has been round the rotation
    rd = sum(per[i]["dram__bytes_read.sum"] for i in tail) / len(tail) / 1e6
    wr = sum(per[i]["dram__bytes_write.sum"] for i in tail) / len(tail) / 1e6
    print(f"{sys.argv[2]}: {len(ids)} pass launches; last {len(tail)}: DRAM read {rd:.1f} MB, write {wr:.1f} MB per launch")
    sys.exit(0)

import torch
from pbg import synth
import modular_prot_b_gan as m
dev = torch.device("cuda:0")
LANES = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
G, D = G.to(dev), D.to(dev)
engs = [m.make_fused_engine(G, D, ctas=0) for _ in range(LANES)]
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
batches = [(synth.make_triplets(B, seed=100 + i).to(dev), synth.make_latents(B, seed=200 + i).to(dev)) for i in range(2 * LANES)]
outs = [{"gen_out": torch.empty(B, 128, dtype=torch.bfloat16, device=dev), "gen_scores": torch.empty(B, device=dev),
         "logits": torch.empty(B, device=dev), "probs": torch.empty(B, device=dev)} for _ in range(LANES)]
STAGE2 = len(sys.argv) > 4 and sys.argv[4] == "stage2"   # bench.py's default: every pass gathers the lane's next request
kw = dict(want_gen_out=True, want_gen_scores=True, want_disc=True, out_dtype=torch.bfloat16)
n = 0
if STAGE2:
    for l in range(LANES):
        engs[l].reserve(B, "bf16", 2)
        engs[l].stage_triplets(0, node_emb, rel_w, *batches[l % len(batches)])
for r in range(ROUNDS):
    for l in range(LANES):
        if STAGE2:
            nxt = batches[(n + LANES) % len(batches)]; n += 1
            engs[l].score_staged(r & 1, out=outs[l], stage_next=(node_emb, rel_w, *nxt), **kw)
        else:
            trip, z = batches[n % len(batches)]; n += 1
            engs[l].score_triplets(node_emb, rel_w, trip, z, precision="bf16", out=outs[l], **kw)
torch.cuda.synchronize()
for e in engs:
    e.check_indices()
print("ok", n, "passes", float(outs[0]["logits"].sum()))
