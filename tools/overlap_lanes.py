"""Throughput of back-to-back independent passes issued round-robin on S streams, each with its own ctx."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
G, D = G.to(dev), D.to(dev)
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
P = 24
pool = []
for i in range(P):
    trip, z = synth.make_triplets(B, seed=4321 + i).to(dev), synth.make_latents(B, seed=1234 + i).to(dev)
    out = {"gen_out": torch.empty(B, 128, dtype=torch.bfloat16, device=dev), "gen_scores": torch.empty(B, device=dev),
           "logits": torch.empty(B, device=dev), "probs": torch.empty(B, device=dev)}
    pool.append((trip, z, out))
for S in [int(x) for x in (sys.argv[2:] or ['1', '2', '3', '4'])]:
    engines = [m.make_fused_engine(G, D) for _ in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]

    def issue(i):
        s = i % S
        trip, z, out = pool[i % P]
        with torch.cuda.stream(streams[s]):
            engines[s].score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True, want_disc=True,
                                      precision="bf16", out_dtype=torch.bfloat16, out=out)
    for i in range(2 * S):
        issue(i)
    torch.cuda.synchronize()
    # one graph holding P*S steps spread over the S streams (fork / join inside the capture)
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    n = P * 2
    with torch.cuda.stream(cap):
        with torch.cuda.graph(g, stream=cap):
            ev = torch.cuda.Event(); ev.record(cap)
            for s in streams:
                s.wait_event(ev)
            for i in range(n):
                issue(i)
            for s in streams:
                e = torch.cuda.Event(); e.record(s); cap.wait_event(e)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (3 * n) * 1e3
    print(f"B={B} streams={S}: {us:.2f} us/step = {B / us:.1f} M samples/s = {B * 4850688 / us / 1e6:.0f} TFLOP/s", flush=True)
    del g, engines
