"""The gob restatement against the known answers published in encoding/gob's package documentation
(the only golden vectors that exist for the reference's data path: the reference itself has no
tests), value-exact round trips (what bounce.go:105,133 assert), and the restated TCP path."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as O


def _put(fn, v):
    buf = (ctypes.c_uint8 * 16)()
    n = getattr(O.lib(), fn)(buf, v)
    return bytes(buf[:n]).hex()


def _golden():
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "gob_doc_vectors.json")) as f:
        return json.load(f)


def test_golden_fixture_vectors():
    """Every entry of tests/golden/gob_doc_vectors.json (transcribed from the gob documentation)."""
    g = _golden()
    for v, want in g["uint"].items():
        assert _put("gob_put_uint", int(v)) == want
    for v, want in g["int"].items():
        assert _put("gob_put_int", int(v)) == want
    for v, want in g["float"].items():
        assert _put("gob_put_float", float(v)) == want
    names = (ctypes.c_char_p * 2)(b"X", b"Y")
    ids = (ctypes.c_int * 2)(2, 2)
    buf = (ctypes.c_uint8 * 128)()
    n = O.lib().gob_put_struct_typedef(buf, b"Point", 65, 2, names, ids)
    assert bytes(buf[:n]).hex() == g["point_type_descriptor"]
    # the value message: length, type id 65, field 0 = 22, field 1 = 33, end of struct
    val = bytes.fromhex(g["point_value_22_33"])
    assert val == bytes([7]) + bytes.fromhex(_put("gob_put_int", 65)) + b"\x01" + bytes.fromhex(_put("gob_put_int", 22)) + b"\x01" + bytes.fromhex(_put("gob_put_int", 33)) + b"\x00"
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "splitmix64.json")) as f:
        sm = json.load(f)
    for i, h in enumerate(sm["seed0"]):
        assert O.lib().oracle_splitmix64(0, i) == int(h, 16)
    for i, h in enumerate(sm["seed_b2000000"]):
        assert O.lib().oracle_splitmix64(0xB2000000, i) == int(h, 16)


def test_documented_scalar_encodings():
    assert _put("gob_put_uint", 0) == "00"
    assert _put("gob_put_uint", 7) == "07"
    assert _put("gob_put_uint", 256) == "fe0100"
    assert _put("gob_put_int", -129) == "fe0101"      # (^i << 1) | 1
    assert _put("gob_put_float", 17.0) == "fe3140"    # byte-reversed float64 bits
    assert _put("gob_put_int", 22) == "2c" and _put("gob_put_int", 33) == "42"
    assert _put("gob_put_int", -65) == "ff81" and _put("gob_put_int", 65) == "ff82"


def test_documented_point_type_descriptor():
    """type Point struct{X, Y int}: the 32-byte descriptor message of the package documentation."""
    names = (ctypes.c_char_p * 2)(b"X", b"Y")
    ids = (ctypes.c_int * 2)(2, 2)
    buf = (ctypes.c_uint8 * 128)()
    n = O.lib().gob_put_struct_typedef(buf, b"Point", 65, 2, names, ids)
    want = "1f" "ff81" "03" "01" "01" "05" "506f696e74" "01" "ff82" "00" "01" "02" \
           "01" "01" "58" "01" "04" "00" "01" "01" "59" "01" "04" "00" "00" "00"
    assert bytes(buf[:n]).hex() == want


def test_slice_streams_have_the_documented_shape():
    s = O.gob_encode(np.array([17.0]))
    # descriptor: len, -65, SliceT(field 1 => delta 2), CommonType{Name "[]float64", Id 65}, Elem float(4 -> 08)
    assert s.hex() == "17ff8102010109" + b"[]float64".hex() + "01ff82000108000007ff820001fe3140"
    assert O.gob_encode(b"abc").hex() == "060a0003616263"  # []byte is predefined id 5, no descriptor
    assert O.gob_encode("hi").hex() == "050c00026869"      # string is id 6


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.uint8])
def test_round_trip_is_value_exact(dtype):
    rng = np.random.default_rng(1)
    for count in (0, 1, 10, 100, 1000, 12345):
        if dtype == np.uint8:
            x = rng.integers(0, 256, count, dtype=np.uint8)
        elif dtype == np.int64:
            x = rng.integers(-2**63, 2**63 - 1, count, dtype=np.int64)
        else:
            x = rng.standard_normal(count).astype(dtype)
        y = O.gob_decode(O.gob_encode(x), dtype, count)
        assert y.size == count and np.array_equal(x.view(np.uint8), y.view(np.uint8))
    if dtype in (np.float64, np.float32):
        fi = np.finfo(dtype)
        edge = np.array([0.0, -0.0, np.inf, -np.inf, fi.tiny, fi.smallest_subnormal, fi.max, -fi.max, 1e38 if dtype == np.float32 else 1e300], dtype=dtype)
        y = O.gob_decode(O.gob_encode(edge), dtype, edge.size)
        assert np.array_equal(edge.view(np.uint8), y.view(np.uint8))
        assert np.isnan(O.gob_decode(O.gob_encode(np.array([np.nan], dtype=dtype)), dtype, 1)[0])
    if dtype == np.int64:
        edge = np.array([0, 1, -1, 63, 64, -64, -65, 2**63 - 1, -2**63], dtype=np.int64)
        assert np.array_equal(O.gob_decode(O.gob_encode(edge), dtype, edge.size), edge)


def test_string_round_trip_and_capacity():
    assert O.gob_decode(O.gob_encode('"Hello node 1, I\'m node 0"'), str, 64) == '"Hello node 1, I\'m node 0"'
    with pytest.raises(ValueError):
        O.gob_decode(O.gob_encode(np.arange(10, dtype=np.int64)), np.int64, 5)


def test_random_floats_cost_about_nine_bytes_each():
    # gob wire ~9 B/elem for random doubles
    x = np.random.default_rng(0).random(10000)
    assert 8.5 < len(O.gob_encode(x)) / x.size < 9.1


@pytest.mark.parametrize("n", [1, 2, 4])
def test_restated_tcp_path_delivers_exact_values(n):
    """bounce's own assertions (bounce.go:105, 133): what comes back equals what was sent."""
    cnt = 4099
    if n >= 2:
        for dt in (np.float64, np.int64, np.float32):
            secs, out = O.ref_bench(O.COLL_PINGPONG, dt, n, cnt, iters=2, warmup=1)
            assert np.array_equal(out, O.fill(dt, 0xB2000000, cnt)) and secs > 0
    for dt in (np.float32, np.float64, np.int64):
        secs, out = O.ref_bench(O.COLL_ALLREDUCE, dt, n, cnt, iters=1, warmup=1)
        ins = [O.fill(dt, 0xB2000000 + r, cnt) for r in range(n)]
        want = O.allreduce(ins, order=O.ORDER_F64)
        if dt == np.int64:
            assert np.array_equal(out, want)
        else:
            assert np.allclose(out, want, rtol=1e-6, atol=0)
    secs, out = O.ref_bench(O.COLL_ALLGATHER, np.int64, n, 1000, iters=1, warmup=1)
    assert np.array_equal(out, O.allgather([O.fill(np.int64, 0xB2000000 + r, 1000) for r in range(n)]))
    secs, out = O.ref_bench(O.COLL_BCAST, np.float32, n, 1000, iters=1, warmup=1)
    assert np.array_equal(out, O.fill(np.float32, 0xB2000000, 1000))
    secs, out = O.ref_bench(O.COLL_ALLREDUCE_NAIVE, np.float64, n, 1000, iters=1, warmup=1)
    assert np.array_equal(out, O.allreduce([O.fill(np.float64, 0xB2000000 + r, 1000) for r in range(n)], order=O.ORDER_RANK))


# ---- an independent restatement of the same published rules, in Python, to cross-check oracle/gob.c
def _py_uint(u):
    if u < 128:
        return bytes([u])
    b = u.to_bytes((u.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def _py_int(i):
    return _py_uint(((~i) << 1 | 1) & (2**64 - 1) if i < 0 else i << 1)


def _py_float(f):
    import struct
    return _py_uint(int.from_bytes(struct.pack("<d", float(f)), "big"))  # float64 bits, byte-reversed


def _py_slice_stream(arr):
    name, elem, enc = {
        np.dtype(np.float64): (b"[]float64", 4, _py_float), np.dtype(np.float32): (b"[]float32", 4, _py_float),
        np.dtype(np.int64): (b"[]int64", 2, lambda v: _py_int(int(v)))}[arr.dtype]
    td = _py_int(-65) + b"\x02\x01\x01" + _py_uint(len(name)) + name + b"\x01" + _py_int(65) + b"\x00\x01" + _py_int(elem) + b"\x00\x00"
    body = _py_int(65) + b"\x00" + _py_uint(arr.size) + b"".join(enc(v) for v in arr)
    return _py_uint(len(td)) + td + _py_uint(len(body)) + body


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
def test_c_encoder_matches_independent_python_encoder(dtype):
    rng = np.random.default_rng(3)
    for count in (0, 1, 7, 200):
        if dtype == np.int64:
            x = rng.integers(-2**63, 2**63 - 1, count, dtype=np.int64)
            if count >= 7:
                x[:7] = [0, 1, -1, 63, 64, -64, -65]
        else:
            x = (rng.standard_normal(count) * 10.0 ** rng.integers(-30, 30, count)).astype(dtype)
            if count >= 7:
                x[:7] = [0.0, -0.0, 1.0, 17.0, 0.5, np.inf, -np.inf]
        assert O.gob_encode(x) == _py_slice_stream(x), (dtype, count)
    raw = bytes(range(256)) * 3
    assert O.gob_encode(raw) == _py_uint(len(_py_int(5) + b"\x00" + _py_uint(len(raw)) + raw)) + _py_int(5) + b"\x00" + _py_uint(len(raw)) + raw
