"""GPU parity tests (``-m gpu``) of the entity scoring + top-k path against the reference's own
torch lines (oracle.cosine_topk = pro_b_gan_infer.py:146-151 / :231-236, CPU fp32).

Bar: indices bit-exact wherever the reference's own fp32 scores separate the candidates by more than fp32 summation
noise (adjacent gap > 1e-5: cuBLAS / MKL / this kernel sum the 128 products in different orders); scores within 2e-6."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SCORE_ATOL = 2e-6
GAP = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked tests need a CUDA device (there is no CPU fallback to test)")
    return torch.device("cuda:0")


def check(queries, table, k, dev, min_exact=0.9):
    import modular_prot_b_gan as m
    from oracle import prot_b_gan_oracle as oracle
    ref_s, ref_i = oracle.cosine_topk(queries, table, min(k + 1, table.shape[0]))
    got_s, got_i = m.cosine_topk(queries.to(dev), table.to(dev), k)
    torch.cuda.synchronize()
    got_s, got_i = got_s.cpu(), got_i.cpu()
    assert got_s.shape == (queries.shape[0], k) and got_i.dtype == torch.int64
    assert (got_s - ref_s[:, :k]).abs().max().item() <= SCORE_ATOL
    gaps = (ref_s[:, :-1] - ref_s[:, 1:]).min(dim=1).values if ref_s.shape[1] > 1 else torch.full((queries.shape[0],), 1.0)
    clear = gaps > GAP
    assert clear.float().mean().item() >= min_exact, "test inputs too degenerate to pin indices"
    assert torch.equal(got_i[clear], ref_i[clear, :k]), "top-k indices differ on rows the reference separates clearly"
    # the remaining rows: every returned index must carry (within noise) the score the reference has at that rank
    full = torch.nn.functional.normalize(queries[~clear], dim=-1) @ torch.nn.functional.normalize(table, dim=-1).T
    assert (full.gather(1, got_i[~clear]) - ref_s[~clear, :k]).abs().max().item() <= 10 * SCORE_ATOL if (~clear).any() else True


@pytest.mark.parametrize("B,N,k", [(4096, 65536, 10), (300, 1000, 5), (1, 257, 16), (17, 65000, 11), (256, 4096, 1)])
def test_topk_matches_reference_lines(dev, B, N, k):
    g = torch.Generator().manual_seed(B * 31 + N)
    check(torch.randn(B, 128, generator=g), torch.randn(N, 128, generator=g), k, dev)


def test_topk_large_k_and_other_width_take_the_general_path(dev):
    """k > 16 or E != 128: exact fp32 scores (SIMT GEMM over row chunks) + one selection CTA per row (k > 64
    used to be refused and k = 17..64 scanned the table once per query)."""
    g = torch.Generator().manual_seed(5)
    check(torch.randn(50, 128, generator=g), torch.randn(3000, 128, generator=g), 40, dev, min_exact=0.5)  # k > 16 (41 gaps per row)
    check(torch.randn(33, 64, generator=g), torch.randn(2000, 64, generator=g), 10, dev)         # E != 128
    check(torch.randn(300, 128, generator=g), torch.randn(65000, 128, generator=g), 64, dev, min_exact=0.3)
    check(torch.randn(40, 128, generator=g), torch.randn(5000, 128, generator=g), 200, dev, min_exact=0.0)   # values only: 201 gaps per row
    check(torch.randn(64, 256, generator=g), torch.randn(4096, 256, generator=g), 10, dev)       # the wide model's E
    import modular_prot_b_gan as m
    with pytest.raises(NotImplementedError):
        m.cosine_topk(torch.randn(4, 128).to(dev), torch.randn(2000, 128).to(dev), 513)


def test_topk_k17_to_64_on_the_filter_path(dev):
    """k = 17 .. 64 stays on the tensor-core filter when the table is large enough (N >= 512 k): cut-off from the
    sampled group keys (16 kept per list, 4096-row chunks so that a row has >= 10 lists), up to 1024 rescored candidates."""
    import modular_prot_b_gan as m
    g = torch.Generator().manual_seed(77)
    q, t = torch.randn(4100, 128, generator=g), torch.randn(65536, 128, generator=g)   # two chunks, the second ragged
    check(q, t, 40, dev, min_exact=0.5)
    eng = [e for (_, width), e in m._TOPK_ENGINES.items() if width == 128][0]
    n_flagged = eng.topk_last_flagged()
    assert n_flagged == 0, eng.topk_flag_report     # the filter proved every row of the last chunk
    check(torch.randn(300, 128, generator=g), torch.randn(40000, 128, generator=g), 64, dev, min_exact=0.3)
    check(torch.randn(64, 128, generator=g), torch.randn(9000, 128, generator=g), 17, dev, min_exact=0.6)


def test_topk_general_path_histogram_select_and_its_fallback(dev):
    """The general path's selection (topk_select_kernel: 4096-bin histogram, collect, bitonic sort): large k, a crowded
    threshold bin that needs the second histogram level (scores within 5e-4 of each other), and a row the kernel cannot
    take -- thousands of IDENTICAL scores at the cut -- which must come back from the per-thread-list kernel with the
    lowest indices first."""
    import modular_prot_b_gan as m
    g = torch.Generator().manual_seed(123)
    check(torch.randn(70, 128, generator=g), torch.randn(3000, 128, generator=g), 300, dev, min_exact=0.0)
    check(torch.randn(9, 128, generator=g), torch.randn(20000, 128, generator=g), 512, dev, min_exact=0.0)
    base = torch.randn(1, 128, generator=g)
    near = base + 2e-3 * torch.randn(6000, 128, generator=g)           # all cosines with `base` within ~1e-5 of each other
    check(base + 0.05 * torch.randn(8, 128, generator=g), near, 70, dev, min_exact=0.0)
    dup = torch.cat([base.repeat(3000, 1), torch.randn(500, 128, generator=g)])
    s, i = m.cosine_topk(base.to(dev), dup.to(dev), 100)
    assert torch.equal(i.cpu()[0], torch.arange(100)), "equal scores: lower index first"
    assert (s.cpu() - 1.0).abs().max().item() < 1e-5


def test_topk_k16_on_the_filter_path_and_ragged_sizes(dev):
    g = torch.Generator().manual_seed(21)
    check(torch.randn(1000, 128, generator=g), torch.randn(65536, 128, generator=g), 16, dev)
    check(torch.randn(257, 128, generator=g), torch.randn(4100, 128, generator=g), 3, dev)      # 17 tiles: 3 sample tiles
    check(torch.randn(5, 128, generator=g), torch.randn(200001, 128, generator=g), 10, dev)


def test_topk_near_duplicate_table_falls_back_to_the_exact_scan(dev):
    """Scores closer together than the bf16 error bound: the filter cannot prove its candidates, the exact scan must
    take over and still rank like the reference."""
    g = torch.Generator().manual_seed(9)
    base = torch.randn(1, 128, generator=g)
    table = base + 0.3 * torch.randn(5000, 128, generator=g)
    queries = base + 0.3 * torch.randn(64, 128, generator=g)
    check(queries, table, 10, dev, min_exact=0.6)   # >= 40 entities within the bf16 bound of the 10th score: the proof must fail


def test_find_similar_entities_form(dev):
    """queries are table rows, k + 1 results, the entity itself first (pro_b_gan_infer.py:231-247)."""
    import modular_prot_b_gan as m
    g = torch.Generator().manual_seed(11)
    table = torch.randn(20000, 128, generator=g).to(dev)
    ids = torch.tensor([0, 5, 19999, 1234], device=dev)
    s, i = m.cosine_topk(table[ids], table, 11)
    assert torch.equal(i[:, 0], ids)
    assert (s[:, 0] - 1.0).abs().max().item() < 1e-5


def test_topk_table_changes_are_seen_and_k_too_large_raises(dev):
    import modular_prot_b_gan as m
    g = torch.Generator().manual_seed(13)
    table = torch.randn(1000, 128, generator=g).to(dev)
    q = torch.randn(8, 128, generator=g).to(dev)
    _, i0 = m.cosine_topk(q, table, 3)
    table[int(i0[0, 0])] = -q[0]           # in-place edit bumps the tensor version: the prepared copy must be rebuilt
    _, i1 = m.cosine_topk(q, table, 3)
    assert int(i1[0, 0]) != int(i0[0, 0])
    with pytest.raises(RuntimeError):
        m.cosine_topk(q, table, 1001)


def test_predict_tails_pipeline_reproduces_the_reference_scripts_golden_output(dev):
    """Generator (fp32 mode) + fused scoring against the fixture the unmodified reference script produced on CPU
    (tests/golden/config1_predict_tails.json): same top-10 indices, scores within 1e-5."""
    import json
    from pathlib import Path
    import modular_prot_b_gan as m
    from pbg import synth
    gold_dir = Path(__file__).resolve().parent / "golden"
    want = json.loads((gold_dir / "config1_predict_tails.json").read_text())
    gold = torch.load(gold_dir / "config1_tensors.pt")
    G, _ = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
    G = G.to(dev).eval()
    node_emb, _ = synth.make_tables()
    G.precision = "fp32"
    with torch.no_grad():
        pred = G(gold["h"].to(dev), gold["r"].to(dev), gold["z"].to(dev))
    scores, idx = m.cosine_topk(pred, node_emb.to(dev), want["metadata"]["top_k"])
    ws = torch.tensor(want["scores"])
    assert (scores.cpu() - ws).abs().max().item() <= 1e-5
    clear = (ws[:, :-1] - ws[:, 1:]).min(dim=1).values > 1e-5   # 15 of the 16 fixture rows
    assert clear.float().mean().item() > 0.8
    assert idx.cpu()[clear].tolist() == torch.tensor(want["predictions"])[clear].tolist()
