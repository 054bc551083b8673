"""The library's host text parser (gsevt_parse_int_table, csrc/events_io.cu) against what the reference's parser does
with the same bytes: split() on whitespace and int() every token (utils/event_camera/event.py:11-39)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H  # noqa: F401  (sys.path for the package)
from gsevt import lib


def _parse(raw: bytes, threads=1, capacity=None):
    L = lib.load()
    buf = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8)
    cap = len(raw) // 2 + 1 if capacity is None else capacity
    out = np.empty(max(cap, 1), np.int64)
    n = int(L.gsevt_parse_int_table(buf.ctypes.data if buf.size else None, buf.size, out.ctypes.data, cap, threads))
    return n, out


def _reference(raw: bytes):
    return np.array([int(t) for t in raw.split()], dtype=np.int64)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_parser_matches_split_int(threads):
    rng = np.random.default_rng(5)
    n = 400_000   # > 1 MiB of text, so that the multi-threaded slicing is exercised
    tab = np.stack([np.sort(rng.integers(0, 10**12, n)), rng.integers(0, 640, n), rng.integers(0, 480, n), rng.integers(0, 2, n)], 1)
    raw = "".join(f"{a} {b} {c} {d}\n" for a, b, c, d in tab.tolist()).encode()
    assert len(raw) > (1 << 20)
    got_n, got = _parse(raw, threads)
    assert got_n == 4 * n and np.array_equal(got[:got_n].reshape(-1, 4), tab)


def test_parser_whitespace_signs_and_missing_final_newline():
    raw = b"  12\t-7\r\n+3   0\n\n\n9223372036854775 -1\x0b5\x0c6"
    n, out = _parse(raw)
    assert np.array_equal(out[:n], _reference(raw))
    assert _parse(b"")[0] == 0 and _parse(b" \n\t ")[0] == 0


def test_size_query_and_capacity():
    L = lib.load()
    raw = b"1 2 3 4\n5 6 7 8\n"
    buf = np.frombuffer(raw, dtype=np.uint8)
    assert int(L.gsevt_parse_int_table(buf.ctypes.data, buf.size, None, 0, 1)) == 8
    n, _ = _parse(raw, capacity=7)
    assert n == -3   # GSEVT_ENOMEM
    assert b"integers" in L.gsevt_last_error()


@pytest.mark.parametrize("bad", [b"1 2 x 4\n", b"1 2.5 3 4\n", b"1 2 3 4a\n", b"1 - 3 4\n", b"1_000 2 3 4\n"])
def test_malformed_token_is_an_error_like_int(bad):
    if b"_" not in bad:   # (int("1_000") is legal Python, but not something an event file holds)
        with pytest.raises(ValueError):
            _reference(bad)
    n, _ = _parse(bad)
    assert n == -1   # GSEVT_EINVAL
    assert b"not an integer" in lib.load().gsevt_last_error()


def test_load_events_from_txt_packets_and_tail(tmp_path):
    """Fixed-count packets, the incomplete tail dropped (event.py:25-37), columns in the file's order ts x y p."""
    from utils.event_camera.event import load_events_from_txt
    rng = np.random.default_rng(1)
    n = 1050
    tab = np.stack([np.sort(rng.integers(0, 10**6, n)), rng.integers(0, 64, n), rng.integers(0, 48, n), rng.integers(0, 2, n)], 1)
    p = tmp_path / "events.txt"
    p.write_text("".join(f"{a} {b} {c} {d}\n" for a, b, c, d in tab.tolist()))
    arrs = load_events_from_txt(str(p), 100)
    assert len(arrs) == 10
    ts, x, y, pol = arrs[3].columns()
    assert np.array_equal(ts, tab[300:400, 0]) and np.array_equal(x, tab[300:400, 1]) and np.array_equal(pol, tab[300:400, 3])
    (tmp_path / "bad.txt").write_text("1 2 3\n")
    with pytest.raises(ValueError):
        load_events_from_txt(str(tmp_path / "bad.txt"), 1)
