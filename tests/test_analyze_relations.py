"""Batched analyze_relations against the fixture the unmodified reference method produced
(tests/golden/make_golden.py -> config1_analyze_relations.json; pro_b_gan_infer.py:264-318).
CPU leg: the host function over the oracle modules.  GPU leg: the same function over the CUDA modules (one fused
discriminator pass for all head x tail x relation triplets), fp32 at 1e-4 and bf16 at 2e-2 relative."""
import json

import pytest
import torch
from torch import nn

from conftest import GOLDEN

HEADS, TAILS = [3, 40000, 65535], [17, 1234]


@pytest.fixture(scope="module")
def gold_rel():
    return json.loads((GOLDEN / "config1_analyze_relations.json").read_text())


def compare(res, gold, tol, gap):
    assert res["metadata"] == gold["metadata"]
    assert len(res["relation_analysis"]) == len(gold["relation_analysis"]) == len(HEADS) * len(TAILS)
    for a, b in zip(res["relation_analysis"], gold["relation_analysis"]):
        assert (a["head_entity"], a["tail_entity"]) == (b["head_entity"], b["tail_entity"])      # pair order :291-292
        ra, rb = a["top_relations"], b["top_relations"]
        assert len(ra) == len(rb)
        pb = [x["probability"] for x in rb]
        assert all(ra[i]["probability"] >= ra[i + 1]["probability"] for i in range(len(ra) - 1))   # descending :310
        for i, (x, y) in enumerate(zip(ra, rb)):
            assert abs(x["discriminator_score"] - y["discriminator_score"]) <= tol
            assert abs(x["probability"] - y["probability"]) <= tol
            clear = all(abs(pb[i] - pb[j]) > gap for j in (i - 1, i + 1) if 0 <= j < len(pb))
            if clear:
                assert x["relation_id"] == y["relation_id"]
            assert isinstance(x["relation_id"], int) and isinstance(x["probability"], float)


def test_batched_analyze_relations_oracle_modules_match_reference_fixture(gold_rel, oracle_models, tables):
    from pbg.analyze import analyze_relations_batched
    _, D = oracle_models
    node_emb, rel_w = tables
    res = analyze_relations_batched(D, node_emb, nn.Embedding.from_pretrained(rel_w), HEADS, TAILS, top_k=5,
                                    model_hit10=gold_rel["metadata"]["model_hit10"])
    compare(res, gold_rel, 1e-5, 2e-5)
    json.dumps(res)                                            # the CLI json.dump()s the result (:421)


def test_batched_analyze_relations_edge_cases(oracle_models, tables):
    from pbg.analyze import analyze_relations_batched
    _, D = oracle_models
    node_emb, rel_w = tables
    rel = nn.Embedding.from_pretrained(rel_w)
    assert analyze_relations_batched(D, node_emb, rel, [], TAILS)["relation_analysis"] == []
    assert analyze_relations_batched(D, node_emb, rel, HEADS, [])["relation_analysis"] == []
    res = analyze_relations_batched(D, node_emb, rel, [5], [5], top_k=1000)           # top_k > R: every relation, once
    top = res["relation_analysis"][0]["top_relations"]
    assert sorted(x["relation_id"] for x in top) == list(range(rel_w.shape[0]))
    # ties keep relation-id order, as the reference's stable sort: a constant discriminator ranks 0, 1, 2, ...
    class Flat:
        def score_triplets(self, node_emb, rel_emb, trip):
            z = torch.zeros(trip.shape[0]); return z, torch.sigmoid(z)
    top = analyze_relations_batched(Flat(), node_emb, rel, [1, 2], [3], top_k=4)["relation_analysis"][1]["top_relations"]
    assert [x["relation_id"] for x in top] == [0, 1, 2, 3] and top[0]["probability"] == 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol,gap", [("fp32", 1e-4, 2e-4), ("bf16", None, None)])
def test_batched_analyze_relations_cuda_modules_match_reference_fixture(gold_rel, synth, tables, prec, tol, gap):
    import modular_prot_b_gan as m
    from pbg.analyze import analyze_relations_batched
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    _, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
    D = D.to(dev).eval()
    D.precision = prec
    node_emb, rel_w = (t.to(dev) for t in tables)
    eng = D._get_engine()
    n0 = eng.launch_count
    res = analyze_relations_batched(D, node_emb, nn.Embedding.from_pretrained(rel_w), HEADS, TAILS, top_k=5,
                                    model_hit10=gold_rel["metadata"]["model_hit10"])
    assert eng.launch_count > n0                                # the CUDA path ran (one pass, not 384 calls)
    if prec == "bf16":
        smax = max(abs(x["discriminator_score"]) for p in gold_rel["relation_analysis"] for x in p["top_relations"])
        tol = 2e-2 * max(smax, 0.05); gap = 2 * tol
    compare(res, gold_rel, tol, gap)
