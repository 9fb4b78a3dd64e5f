"""Generates the golden fixtures in this directory.  Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

What it pins
------------
The reference ships no model, no tests and no golden vectors, so the fixtures are outputs of
  (1) the *unmodified* /root/reference/pro_b_gan_infer.py, driven through its own main() (CLI, argparse, JSON
      result schema :153-165, :190-211) on CPU fp32, with the oracle restatement injected as the missing
      ``modular_prot_b_gan`` module and as the undefined ``Generator`` / ``Discriminator`` names (:41, :93-94), on
      a synthetic checkpoint in the reference's wire format (:74-112)                      -> config1_*.json
  (2) the oracle modules called directly on the same seeded tensors (B = 16)               -> config1_tensors.pt
BASELINE.json configs[0]: "pro_b_gan_infer.py as shipped, batch=16, random-init weights, fixed-seed latents,
fp32 on CPU".  Everything is regenerated from the frozen seeds in pbg/synth.py; only small outputs are committed.
"""
from __future__ import annotations

import contextlib
import io
import json
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
for p in (str(ROOT / "pro-b-gan_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import prot_b_gan_oracle as oracle  # noqa: E402
from pbg import launcher, synth  # noqa: E402

REFERENCE_SCRIPT = "/root/reference/pro_b_gan_infer.py"
B = 16
TOP_K = 10
SIMILAR_QUERIES = [0, 7, 40000, 65535]
ANALYZE_HEADS, ANALYZE_TAILS = [3, 40000, 65535], [17, 1234]


def run_cli(argv, ckpt_path):
    """ref.main() with sys.argv set, stdout captured; returns the JSON document the script printed."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        launcher.run_main(REFERENCE_SCRIPT, ["--checkpoint_path", ckpt_path, "--device", "cpu", *argv],
                          model_module=oracle)
    text = buf.getvalue()
    return json.loads(text[text.index("{"):])


def main():
    torch.set_num_threads(1)  # one thread: reduction order inside the CPU GEMMs is then machine-independent
    ckpt = synth.make_checkpoint(oracle.ModularGenerator, oracle.ModularDiscriminator)
    trip = synth.make_triplets(B)
    triplets = trip.tolist()
    pairs = [[t[0], t[1]] for t in triplets]
    with tempfile.TemporaryDirectory() as d:
        path = str(Path(d) / "synthetic_ckpt.pt")
        torch.save(ckpt, path)
        # each CLI run constructs fresh modules -> the generator's latent stream restarts at seed 1234
        score = run_cli(["--task", "score_triplets", "--input_triplets", json.dumps(triplets)], path)
        pred = run_cli(["--task", "predict_tails", "--input_pairs", json.dumps(pairs), "--top_k", str(TOP_K)], path)
        info = run_cli(["--task", "model_info"], path)
        similar = run_cli(["--task", "similar_entities", "--input_entities", json.dumps(SIMILAR_QUERIES), "--top_k",
                           str(TOP_K)], path)
    (HERE / "config1_score_triplets.json").write_text(json.dumps(score, indent=1))
    (HERE / "config1_predict_tails.json").write_text(json.dumps(pred, indent=1))
    info["checkpoint_path"] = "synthetic_ckpt.pt"   # the temp directory differs per run
    (HERE / "config1_model_info.json").write_text(json.dumps(info, indent=1))
    (HERE / "config1_similar_entities.json").write_text(json.dumps(similar, indent=1))

    # analyze_relations is accepted by the CLI but never dispatched (a quirk the tests preserve), so the method of
    # the unmodified class is called directly: 3 heads x 2 tails x all 64 relations, top 5        -> :264-318
    with tempfile.TemporaryDirectory() as d2:
        path2 = str(Path(d2) / "synthetic_ckpt.pt")
        torch.save(ckpt, path2)
        ref = launcher.load_reference_script(REFERENCE_SCRIPT, model_module=oracle)
        with contextlib.redirect_stdout(io.StringIO()):
            inf = ref.ProtBGANInference(path2, "cpu")
            rel = inf.analyze_relations(ANALYZE_HEADS, ANALYZE_TAILS, top_k=5)
    (HERE / "config1_analyze_relations.json").write_text(json.dumps(rel, indent=1))

    # direct tensors, B = 16
    G, D = synth.make_models(oracle.ModularGenerator, oracle.ModularDiscriminator)
    node_emb, rel_w = synth.make_tables()
    z = synth.make_latents(B)
    with torch.no_grad():
        h, r, t = node_emb[trip[:, 0]], rel_w[trip[:, 1]], node_emb[trip[:, 2]]
        g = G(h, r, z)
        d = D(h, r, t)
        cs = F.cosine_similarity(g, t, dim=1)
    torch.save({"triplets": trip, "z": z, "h": h, "r": r, "t": t, "gen_out": g, "logits": d,
                "probs": torch.sigmoid(d), "gen_scores": cs}, HERE / "config1_tensors.pt")
    # the CLI and the direct call must agree (same seeds, same latents): this pins the seam itself
    assert torch.allclose(torch.tensor(score["discriminator_logits"]), d, atol=1e-6)
    assert torch.allclose(torch.tensor(score["generator_scores"]), cs, atol=1e-6)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
