"""GPU tests of the fused host class (pbg.inference.FusedInference) against the JSON documents the UNMODIFIED reference
script printed for the same synthetic checkpoint (tests/golden/make_golden.py): every public method of
ProtBGANInference (pro_b_gan_infer.py:118-338), same inputs, same result dictionaries, same error types.
fp32 mode: floats within 1e-4; bf16 mode: within 2e-2 of the largest magnitude; ids exact wherever the fixture's
neighbouring scores differ by more than the tolerance."""
import json

import pytest
import torch

from conftest import GOLDEN
from test_analyze_relations import HEADS, TAILS, compare as compare_relations

pytestmark = pytest.mark.gpu
SIMILAR_QUERIES = [0, 7, 40000, 65535]


def gold_json(name):
    return json.loads((GOLDEN / name).read_text())


@pytest.fixture(scope="module")
def ckpt_path(tmp_path_factory, oracle, synth):
    path = tmp_path_factory.mktemp("ckpt") / "synthetic_ckpt.pt"
    torch.save(synth.make_checkpoint(oracle.ModularGenerator, oracle.ModularDiscriminator), path)
    return str(path)


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def inf(request, ckpt_path):
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device (there is no CPU fallback)"
    from pbg.inference import FusedInference
    return FusedInference(ckpt_path, "cuda", precision=request.param)


def tol_for(inf, values):
    if inf.precision == "fp32":
        return 1e-4
    return 2e-2 * max(max(abs(v) for v in values), 0.05)


def close(a, b, tol):
    assert len(a) == len(b)
    assert max(abs(x - y) for x, y in zip(a, b)) <= tol


def test_score_triplets_matches_reference_cli_document(inf):
    want = gold_json("config1_score_triplets.json")
    inf.generator.reseed()                                   # each CLI run starts the latent stream afresh
    n0 = inf.engine.launch_count
    got = inf.score_triplets(want["triplets"], method="both")
    if inf.precision == "bf16":
        assert inf.engine.launch_count - n0 == 1             # gather + G + cosine + D: ONE launch
    assert got["triplets"] == want["triplets"] and got["metadata"] == want["metadata"]
    assert list(got) == list(want)                           # same keys in the same order
    for k in ("generator_scores", "discriminator_logits", "discriminator_probabilities"):
        assert all(isinstance(v, float) for v in got[k])
        close(got[k], want[k], tol_for(inf, want[k]))
    json.dumps(got)


def test_score_triplets_accepts_cli_text_and_arrays_identically(inf):
    want = gold_json("config1_score_triplets.json")
    docs = []
    for form in (want["triplets"], json.dumps(want["triplets"]), torch.tensor(want["triplets"]),
                 torch.tensor(want["triplets"]).numpy(), [tuple(t) for t in want["triplets"]]):
        inf.generator.reseed()
        docs.append(inf.score_triplets(form))
    for d in docs[1:]:
        for k in ("generator_scores", "discriminator_logits", "discriminator_probabilities"):
            assert d[k] == docs[0][k]                        # bit-identical floats
        assert [list(t) for t in d["triplets"]] == want["triplets"]


def test_score_triplets_methods(inf):
    trip = gold_json("config1_score_triplets.json")["triplets"]
    inf.generator.reseed()
    both = inf.score_triplets(trip, "both")
    inf.generator.reseed()
    d_only = inf.score_triplets(trip, "discriminator")       # draws no latents (:199 is skipped)
    g_only = inf.score_triplets(trip, "generator")           # so this sees the stream from its start
    assert "generator_scores" not in d_only and "discriminator_logits" not in g_only
    assert d_only["discriminator_logits"] == both["discriminator_logits"]
    assert g_only["generator_scores"] == both["generator_scores"]
    other = inf.score_triplets(trip, "neither")
    assert list(other) == ["triplets", "metadata"] and other["metadata"]["method"] == "neither"


def test_predict_tails_matches_reference_cli_document(inf):
    want = gold_json("config1_predict_tails.json")
    pairs = [[t[0], t[1]] for t in gold_json("config1_score_triplets.json")["triplets"]]
    inf.generator.reseed()
    got = inf.predict_tails(pairs, top_k=10, return_scores=True)
    assert got["metadata"] == want["metadata"] and list(got) == list(want)
    flat = [s for row in want["scores"] for s in row]
    tol = 1e-4 if inf.precision == "fp32" else 2e-2
    for grow, wrow, gs, ws in zip(got["predictions"], want["predictions"], got["scores"], want["scores"]):
        close(gs, ws, tol)
        for c in range(len(wrow)):
            clear = all(abs(ws[c] - ws[j]) > 2 * tol for j in (c - 1, c + 1) if 0 <= j < len(ws))
            if clear and inf.precision == "fp32":
                assert grow[c] == wrow[c]
        if inf.precision == "bf16":                          # the generator's bf16 error may reorder near-ties: same set
            assert len(set(grow) & set(wrow)) >= len(wrow) - 3
    inf.generator.reseed()
    assert "scores" not in inf.predict_tails(json.dumps(pairs), top_k=10)
    assert max(abs(s) for s in flat) <= 1.0 + 1e-5


def test_find_similar_entities_matches_reference_cli_document(inf):
    want = gold_json("config1_similar_entities.json")
    for form in (SIMILAR_QUERIES, json.dumps(SIMILAR_QUERIES)):
        got = inf.find_similar_entities(form, top_k=10)
        assert got["metadata"] == want["metadata"]
        for g, w in zip(got["similar_entities"], want["similar_entities"]):
            assert g["query_entity"] == w["query_entity"]
            assert g["similar_entities"] == w["similar_entities"]          # no model in this path: ids exact
            assert w["query_entity"] not in g["similar_entities"] and len(g["similar_entities"]) == 10
            close(g["similarity_scores"], w["similarity_scores"], 2e-6)


def test_analyze_relations_matches_reference_method_document(inf):
    want = gold_json("config1_analyze_relations.json")
    got = inf.analyze_relations(HEADS, TAILS, top_k=5)
    if inf.precision == "fp32":
        compare_relations(got, want, 1e-4, 2e-4)
    else:
        smax = max(abs(x["discriminator_score"]) for p in want["relation_analysis"] for x in p["top_relations"])
        compare_relations(got, want, 2e-2 * max(smax, 0.05), 4e-2 * max(smax, 0.05))
    assert inf.analyze_relations(json.dumps(HEADS), json.dumps(TAILS), 5)["relation_analysis"][0]["head_entity"] == HEADS[0]


def test_get_model_info_matches_reference_cli_document(inf, ckpt_path):
    want = gold_json("config1_model_info.json")
    got = inf.get_model_info()
    assert got["checkpoint_path"] == ckpt_path and got["device"].startswith("cuda")
    for k in ("model_architecture", "training_performance"):
        assert got[k] == want[k]
    assert list(got) == list(want)


def test_negative_ids_wrap_like_tensor_indexing_and_bad_ids_raise(inf):
    """node_emb[idx] wraps -N..-1 (:186, :188); nn.Embedding does not (:187); anything else out of range raises."""
    N, R = inf.num_entities, inf.num_relations
    inf.generator.reseed()
    a = inf.score_triplets([[-1, 5, -N], [3, R - 1, -7]])
    inf.generator.reseed()
    b = inf.score_triplets([[N - 1, 5, 0], [3, R - 1, N - 7]])
    for k in ("generator_scores", "discriminator_logits", "discriminator_probabilities"):
        assert a[k] == b[k]
    assert a["triplets"] == [[-1, 5, -N], [3, R - 1, -7]]
    for bad in ([[0, -1, 1]], [[N, 0, 1]], [[0, R, 1]], [[0, 0, -N - 1]]):
        with pytest.raises(IndexError):
            inf.score_triplets(bad)
    with pytest.raises(IndexError):
        inf.predict_tails([[N, 0]])
    with pytest.raises(IndexError):
        inf.find_similar_entities([N])
    with pytest.raises(IndexError):
        inf.analyze_relations([N], [0])
    with pytest.raises(IndexError):
        inf.score_triplets([])                               # torch.tensor([])[:, 0] in the reference
    with pytest.raises(ValueError):
        inf.score_triplets([[0, 1, 2], [3, 4]])
    with pytest.raises(ValueError):
        inf.score_triplets("[[0, 1, 2], [3, 4.5, 6]]")
    inf.generator.reseed()
    assert inf.score_triplets([[0, 1, 2]])["metadata"]["num_triplets"] == 1      # the ctx is still healthy


def test_large_request_from_cli_text_matches_oracle(inf, oracle_models, tables, synth):
    """B = 5000 (ragged against the 256-row blocks) through JSON text, against the oracle on the same latents."""
    import torch.nn.functional as F
    B = 5000
    trip = synth.make_triplets(B)
    inf.generator.reseed()
    got = inf.score_triplets(json.dumps(trip.tolist()))
    Go, Do = oracle_models
    node_emb, rel_w = tables
    with torch.no_grad():
        h, r, t = node_emb[trip[:, 0]], rel_w[trip[:, 1]], node_emb[trip[:, 2]]
        g = Go(h, r, synth.make_latents(B))
        cs, d = F.cosine_similarity(g, t, dim=1), Do(h, r, t)
    tol = 1e-4 if inf.precision == "fp32" else 2e-2
    assert (torch.tensor(got["generator_scores"]) - cs).abs().max().item() <= tol
    assert (torch.tensor(got["discriminator_logits"]) - d).abs().max().item() <= tol * max(d.abs().max().item(), 1.0)
    assert (torch.tensor(got["discriminator_probabilities"]) - torch.sigmoid(d)).abs().max().item() <= tol


def test_missing_checkpoint_and_cpu_device_fail_loudly(tmp_path, ckpt_path):
    from pbg.inference import FusedInference
    with pytest.raises(FileNotFoundError):
        FusedInference(str(tmp_path / "nope.pt"), "cuda")
    with pytest.raises(RuntimeError):
        FusedInference(ckpt_path, "cpu")


def test_as_arrays_and_the_c_writer_reproduce_the_json_document(inf):
    """score_triplets / predict_tails with as_arrays=True + pbg.hostio.dumps_results give the very text
    json.dumps(results, indent=2) prints for the list form (the reference's output step, :505-508)."""
    import numpy as np
    from pbg import hostio
    trip = gold_json("config1_score_triplets.json")["triplets"]
    inf.generator.reseed()
    lists = inf.score_triplets(trip)
    inf.generator.reseed()
    arrays = inf.score_triplets(trip, as_arrays=True)
    assert list(arrays) == list(lists) and isinstance(arrays["generator_scores"], np.ndarray)
    assert hostio.dumps_results(arrays, indent=2) == json.dumps(lists, indent=2)
    inf.generator.reseed()
    again = inf.score_triplets(trip[:5], as_arrays=True)            # the arrays own their memory: a later call does not
    assert arrays["discriminator_logits"].tolist() == lists["discriminator_logits"] and len(again["triplets"]) == 5
    pairs = [[t[0], t[1]] for t in trip]
    inf.generator.reseed()
    pl = inf.predict_tails(pairs, top_k=10, return_scores=True)
    inf.generator.reseed()
    pa = inf.predict_tails(pairs, top_k=10, return_scores=True, as_arrays=True)
    assert hostio.dumps_results(pa, indent=2) == json.dumps(pl, indent=2)
