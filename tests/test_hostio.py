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


def test_parser_fuzz_against_json_loads():
    """Random well-formed texts (random whitespace, signs, widths) parse to what json.loads + torch.tensor give; random
    single-character corruptions either still mean the same integers to json.loads or are rejected by both."""
    import random
    rnd = random.Random(20261018)
    ws = ["", " ", "  ", "\n", "\t", "\r\n "]
    for trial in range(300):
        cols = rnd.choice([1, 2, 3])
        rows = rnd.randrange(0, 12)
        vals = [[rnd.choice([0, 1, -1, rnd.randrange(-10**6, 10**6), rnd.randrange(-2**63, 2**63)]) for _ in range(cols)]
                for _ in range(rows)]
        def num(v):
            return rnd.choice(ws) + str(v) + rnd.choice(ws)
        if cols == 1:
            text = rnd.choice(ws) + "[" + ",".join(num(r[0]) for r in vals) + "]" + rnd.choice(ws)
        else:
            text = rnd.choice(ws) + "[" + ",".join(rnd.choice(ws) + "[" + ",".join(num(v) for v in r) + "]" + rnd.choice(ws)
                                                   for r in vals) + "]" + rnd.choice(ws)
        if rows == 0:
            text = rnd.choice(ws) + "[" + rnd.choice(ws) + "]"
        want = torch.tensor(vals, dtype=torch.int64).reshape(rows, cols)
        assert torch.equal(hostio.parse_index_rows(text, cols), want), text
        if not text.strip("[] \n\t\r"):
            continue
        pos = rnd.randrange(len(text))
        bad = text[:pos] + rnd.choice("x.,[]- 0e\"") + text[pos + 1:]
        try:
            ref = json.loads(bad)
            ref_t = torch.tensor(ref, dtype=torch.int64)
            ok_ref = ref_t.numel() == 0 and cols != 1 or (ref_t.dim() == (1 if cols == 1 else 2) and (cols == 1 or ref_t.shape[1] == cols))
            if any(isinstance(x, float) for x in (ref if cols == 1 else sum((r if isinstance(r, list) else [r] for r in ref), []))):
                ok_ref = False
        except Exception:
            ok_ref = False
        try:
            got = hostio.parse_index_rows(bad, cols)
        except ValueError:
            got = None
        if got is not None:       # whatever the C parser accepts, json.loads accepts with the same integers
            assert ok_ref, (bad, got.tolist())
            assert got.reshape(-1).tolist() == ref_t.reshape(-1).tolist(), bad


def _special_floats():
    import struct
    vals = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e-4, 9.9999e-5, 1e-5, 1.5e-7, 1e15, 1e16, 9.999999e15, 1.2345678e17, 3.4028235e38,
            1.17549435e-38, 1e-45, 123456.789, 100.0, 1e7, 16777216.0, 0.30000001192092896, float("inf"), float("-inf"), float("nan")]
    vals += [struct.unpack("<f", struct.pack("<I", b))[0] for b in (0x00000001, 0x007FFFFF, 0x00800000, 0x3F800001, 0x7F7FFFFF, 0x4B000000, 0x38D1B717)]
    return np.asarray(vals, dtype=np.float32)


def test_result_writer_reproduces_json_dumps_byte_for_byte():
    """pbg_format_f32_json / pbg_format_i64_json + dumps_results == json.dumps(results with .tolist(), indent=2), the
    reference's output step (pro_b_gan_infer.py:505-508): Python float repr of every fp32 value, same layout."""
    rng = np.random.default_rng(7)
    f = np.concatenate([_special_floats(), rng.standard_normal(3000).astype(np.float32),
                        (rng.standard_normal(2000) * 10.0 ** rng.integers(-30, 30, 2000)).astype(np.float32),
                        rng.integers(0, 2**32, 3000, dtype=np.uint64).astype(np.uint32).view(np.float32)])   # any bit pattern
    ints = np.concatenate([rng.integers(-2**63, 2**63 - 1, 500), np.asarray([0, -1, 1, 2**63 - 1, -2**63, 65535])])
    res = {
        "triplets": rng.integers(0, 65536, (257, 3)),
        "metadata": {"num_triplets": 257, "method": "both", "model_hit10": 0.4242, "nested": {"a": [1, 2.5, None]}},
        "generator_scores": f,
        "predictions": ints[:500].reshape(50, 10),
        "scores": torch.from_numpy(f[:4000].reshape(400, 10).copy()),
        "empty": np.zeros(0, dtype=np.float32),
        "empty_rows": np.zeros((0, 3), dtype=np.int64),
        "a_list": [1, 2, 3],
        "a_string": "x\"y\u00e9",
    }
    ref = {k: (v.tolist() if isinstance(v, (np.ndarray, torch.Tensor)) else v) for k, v in res.items()}
    for indent in (2, 1, 4, 0):
        assert hostio.dumps_results(res, indent=indent) == json.dumps(ref, indent=indent)
    assert hostio.dumps_results(res, indent=-1) == json.dumps(ref)
    assert hostio.dumps_results({}) == json.dumps({}, indent=2)
    back = json.loads(hostio.dumps_results({"x": f[np.isfinite(f)]}))["x"]
    assert np.array_equal(np.asarray(back, dtype=np.float32), f[np.isfinite(f)])          # and it round-trips exactly


def test_result_writer_sizes_its_buffer():
    import ctypes as C
    from pbg import cabi
    lib = cabi.load()
    a = np.asarray([1.5, -2.25e-7, 3.0], dtype=np.float32)
    need = lib.pbg_format_f32_json(C.c_void_p(a.ctypes.data), 3, 0, 2, 0, None, 0)
    want = json.dumps(a.tolist(), indent=2)
    assert need == len(want)
    small = C.create_string_buffer(4)
    assert lib.pbg_format_f32_json(C.c_void_p(a.ctypes.data), 3, 0, 2, 0, small, 4) == need   # never writes past cap
    buf = C.create_string_buffer(need)
    assert lib.pbg_format_f32_json(C.c_void_p(a.ctypes.data), 3, 0, 2, 0, buf, need) == need and buf.raw[:need].decode() == want
