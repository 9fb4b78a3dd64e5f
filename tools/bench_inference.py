"""Request-level timing of score_triplets at the CLI boundary (JSON text in, result dict out) on one GPU:
  script-style : json.loads -> torch.tensor(list) -> gathers -> G -> cosine -> D.score_triplets -> .tolist()
                 (the steps of pro_b_gan_infer.py:182-209 over the CUDA modules: the module-level drop-in)
  fused host   : pbg.inference.FusedInference.score_triplets(text)  (C parser, one fused pass, pinned staging)
Appends to gpurun_out/bench_inference.log."""
import json
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
import torch.nn.functional as F

from pbg import synth
from pbg.inference import FusedInference
import modular_prot_b_gan as m

out = open(ROOT / "gpurun_out" / "bench_inference.log", "a")


def P(s):
    print(s, flush=True); out.write(s + "\n"); out.flush()


with tempfile.TemporaryDirectory() as d:
    path = str(Path(d) / "ckpt.pt")
    torch.save(synth.make_checkpoint(m.ModularGenerator, m.ModularDiscriminator), path)
    t0 = time.perf_counter()
    inf = FusedInference(path, "cuda")
    P(f"checkpoint ingest: {(time.perf_counter() - t0) * 1e3:.1f} ms total; " +
      ", ".join(f"{k} {v:.1f} ms" for k, v in inf.ingest_ms.items()))
    inf2 = FusedInference(path, "cuda")
    P("second ingest (CUDA context warm): " + ", ".join(f"{k} {v:.1f} ms" for k, v in inf2.ingest_ms.items()))
    del inf2

dev = inf.device
G, D, node_emb, rel_w = inf.generator, inf.discriminator, inf.node_emb, inf.rel_weight


def script_style(text):
    triplets = json.loads(text)
    with torch.no_grad():
        tt = torch.tensor(triplets, device=dev)
        h, r, t = node_emb[tt[:, 0]], rel_w[tt[:, 1]], node_emb[tt[:, 2]]
        pred = G(h, r)
        res = {"triplets": triplets, "generator_scores": F.cosine_similarity(pred, t, dim=1).cpu().numpy().tolist()}
        lo, pr = D.score_triplets(node_emb, rel_w, tt)
        res["discriminator_logits"], res["discriminator_probabilities"] = lo.tolist(), pr.tolist()
    return res


def timeit(f, n):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for B in (16, 4096, 32768):
    text = json.dumps(synth.make_triplets(B).tolist())
    n = 200 if B <= 16 else 20 if B <= 4096 else 5
    a = timeit(lambda: script_style(text), n)
    b = timeit(lambda: inf.score_triplets(text), n)
    trip = synth.make_triplets(B)
    c = timeit(lambda: inf.score_triplets(trip), n)
    from pbg import hostio
    d = timeit(lambda: json.dumps(script_style(text), indent=2), max(2, n // 4))          # JSON text in -> JSON text out
    e = timeit(lambda: hostio.dumps_results(inf.score_triplets(text, as_arrays=True), indent=2), max(2, n // 4))
    P(f"score_triplets CLI round trip (JSON text in -> JSON text out, indent=2), B={B:6d}: script-style {d:8.2f} ms | "
      f"fused host + C reader / writer {e:7.2f} ms ({d / e:.1f}x)")
    P(f"score_triplets request, B={B:6d}: script-style over the CUDA modules {a:8.2f} ms | fused host from JSON text {b:7.2f} ms "
      f"({a / b:.1f}x) | fused host from an int64 tensor {c:7.2f} ms ({B / c / 1e3:.2f} M triplets/s; latent draw on the CPU included)")
