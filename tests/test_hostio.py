"""Host-side ingest: the C parser for the CLI's JSON index arrays and the list / array converters
must produce exactly what the reference's ``json.loads`` + ``torch.tensor`` produce (pro_b_gan_infer.py:485-501,
:135-136, :182, :226), and refuse what they refuse.  No GPU needed: pbg_parse_index_rows touches no device."""
import json

import numpy as np
import pytest
import torch

from pbg import hostio


@pytest.mark.parametrize("rows,cols", [(0, 3), (1, 3), (16, 3), (4097, 3), (33, 2), (7, 1)])
def test_parser_matches_json_loads_plus_torch_tensor(rows, cols):
    g = torch.Generator().manual_seed(rows * 10 + cols)
    want = torch.randint(-5, 1 << 40, (rows, cols), generator=g)
    lists = want.tolist() if cols > 1 else want[:, 0].tolist()
    for text in (json.dumps(lists), json.dumps(lists, indent=2), json.dumps(lists, separators=(",", ":"))):
        ref = torch.tensor(json.loads(text), dtype=torch.int64).reshape(rows, cols)      # the reference's two steps
        got = hostio.parse_index_rows(text, cols)
        assert got.dtype == torch.int64 and torch.equal(got, ref)
        assert torch.equal(hostio.index_rows(text.encode(), cols), ref)
    assert torch.equal(hostio.index_rows(lists, cols), want)
    small = (want % 65536).to(torch.int32)
    assert torch.equal(hostio.index_rows(small.numpy(), cols), small.long())          # narrower integer arrays widen
    assert torch.equal(hostio.index_rows(want, cols), want)


def test_parser_int64_limits_and_whitespace():
    got = hostio.parse_index_rows(" [\n 9223372036854775807 ,\t-9223372036854775808, 0, -0 ]\r\n", 1)
    assert got[:, 0].tolist() == [9223372036854775807, -9223372036854775808, 0, 0]


@pytest.mark.parametrize("text,cols", [
    ("[[1,2,3],[4,5]]", 3), ("[[1,2,3,4]]", 3), ("[1.5]", 1), ("[1e3]", 1), ("[[1,2,3]] x", 3), ("[[01,2,3]]", 3),
    ("[[1,2,9223372036854775808]]", 3), ("", 3), ("[[1,2,3],]", 3), ("[1,2,3]", 3), ("[[1,2,3]", 3), ("[[1,2,\"3\"]]", 3),
    ("[[1,2,null]]", 3), ("{\"a\": 1}", 1), ("[+1]", 1), ("[[1 2 3]]", 3),
])
def test_parser_rejects_what_the_reference_path_rejects(text, cols):
    """Each of these makes json.loads or torch.tensor(...)-then-index fail in the reference; none may yield ids."""
    with pytest.raises(ValueError):
        hostio.parse_index_rows(text, cols)


def test_list_inputs_keep_the_reference_error_types():
    with pytest.raises(ValueError):                     # torch.tensor([[0, 1, 2], [3, 4]]) -> ValueError
        hostio.index_rows([[0, 1, 2], [3, 4]], 3)
    with pytest.raises(IndexError):                     # float ids: torch indexes with a float tensor -> IndexError
        hostio.index_rows([[0.5, 1, 2]], 3)
    with pytest.raises(IndexError):
        hostio.index_rows(torch.tensor([[0.5, 1, 2]]), 3)
    with pytest.raises(ValueError):                     # pairs where triplets are expected
        hostio.index_rows([[0, 1]], 3)
    assert hostio.index_rows([], 3).shape == (0, 3) and hostio.index_rows((), 1).shape == (0, 1)
    assert hostio.index_rows([(1, 2, 3)], 3).tolist() == [[1, 2, 3]]                      # tuples, as the type hints say
