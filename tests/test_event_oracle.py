"""Pins the event-side oracle (oracle/event_oracle.py) against (a) OpenCV itself — the third-party library
the reference calls (utils/event_camera/event.py:121-124) — and (b) golden frames produced by the
reference's own event.py in tests/golden/make_event_golden.py.  CPU only."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import event_oracle as eo

GOLD = os.path.join(H.GOLDEN, "events_96x64.npz")


def _frame(x, y, p, W, Hh):
    f = np.zeros((Hh, W), np.float32)
    np.add.at(f, (np.asarray(y, int), np.asarray(x, int)), np.where(np.asarray(p) != 0, 1, -1).astype(np.float32))
    return f


@pytest.mark.parametrize("W,Hh,n,seed", [(640, 480, 30000, 0), (96, 64, 500, 1), (320, 240, 100000, 2)])
def test_stages_match_opencv_bit_exactly(W, Hh, n, seed):
    import cv2
    rng = np.random.default_rng(seed)
    x, y, p = rng.integers(0, W, n), rng.integers(0, Hh, n), rng.integers(0, 2, n)
    s = W / 640.0
    K = np.array([327.32749 * s, 0, 304.97749 * s, 0, 327.46184 * s, 235.37621 * s, 0, 0, 1.0]).reshape(3, 3)
    D = np.array([-0.031982, 0.041966, -0.000507, -0.001031, 0.0])
    cnt = eo.accumulate(x, y, p, W, Hh)
    f = _frame(x, y, p, W, Hh)
    assert np.array_equal(cnt.astype(np.float32), f)                                   # E0
    und = cv2.undistort(f, K, D)
    assert H.bits_equal(eo.undistort(f, K, D), und)                                     # E1
    blur = cv2.GaussianBlur(und, (9, 9), 0, borderType=cv2.BORDER_REPLICATE)
    assert H.bits_equal(eo.gaussian_blur9(und), blur)                                   # E2
    assert H.bits_equal(eo.l2_normalize(blur), cv2.normalize(blur, None))               # E3
    s_ref, u_ref = eo.event_frame(x, y, p, W, Hh, K, D)
    assert H.bits_equal(s_ref[0], cv2.normalize(blur, None)) and H.bits_equal(u_ref, np.abs(s_ref))
    for l, lvl in enumerate(eo.pyramid(s_ref[0])):                                      # pyramid == INTER_NEAREST
        ref = cv2.resize(s_ref[0], (int(W * 0.5 ** l), int(Hh * 0.5 ** l)), interpolation=cv2.INTER_NEAREST)
        assert H.bits_equal(lvl, ref)


@pytest.mark.parametrize("ksize", [1, 3, 5, 7, 9])
def test_dyadic_kernel_sizes_match_opencv_bit_exactly(ksize):
    """Event.gaussian_kernel_size other than the configs' 9: the generic blur (OpenCV's coefficients from the committed
    table, row pass left to right with FMA, symmetric column pass) against cv2.GaussianBlur on event frames, and the
    committed table against cv2.getGaussianKernel itself.  Up to 9 taps the coefficients are multiples of 1/256 and the
    blur is exact; from 11 on cv2's result depends on its vector-body / scalar-tail split (seen here: 11 taps agree at
    widths 96 and 320 and differ in the tail columns at width 33), so those sizes are rejected by the product."""
    import cv2
    assert np.array_equal(eo.gauss_taps(ksize), cv2.getGaussianKernel(ksize, 0, cv2.CV_32F).ravel().astype(np.float32))
    for W, Hh, n, seed in ((96, 64, 20000, 0), (320, 240, 30000, 1), (33, 17, 3000, 2)):
        rng = np.random.default_rng(seed)
        x, y, p = rng.integers(0, W, n), rng.integers(0, Hh, n), rng.integers(0, 2, n)
        s = W / 640.0
        K = np.array([327.32749 * s, 0, 304.97749 * s, 0, 327.46184 * s, 235.37621 * s, 0, 0, 1.0]).reshape(3, 3)
        D = np.array([-0.031982, 0.041966, -0.000507, -0.001031, 0.0])
        und = cv2.undistort(_frame(x, y, p, W, Hh), K, D)
        blur = cv2.GaussianBlur(und, (ksize, ksize), 0, borderType=cv2.BORDER_REPLICATE)
        assert H.bits_equal(eo.gaussian_blur(und, ksize), blur), (ksize, W, Hh)
        s_ref, _ = eo.event_frame(x, y, p, W, Hh, K, D, ksize=ksize)
        assert H.bits_equal(s_ref[0], cv2.normalize(blur, None))
    if ksize == 9:
        assert H.bits_equal(eo.gaussian_blur(und, 9), eo.gaussian_blur9(und))


@pytest.mark.parametrize("W,Hh", [(346, 260), (641, 479), (100, 75), (640, 480)])
def test_pyramid_is_opencv_nearest_on_sizes_not_divisible_by_four(W, Hh):
    """cv2.resize(INTER_NEAREST) samples floor(dst * src / dst_size): on 346 x 260 (DAVIS346) level 2 is NOT frame[::4, ::4]."""
    import cv2
    rng = np.random.default_rng(W)
    f = rng.normal(size=(Hh, W)).astype(np.float32)
    for l, lvl in enumerate(eo.pyramid(f)):
        ref = cv2.resize(f, (int(W * 0.5 ** l), int(Hh * 0.5 ** l)), interpolation=cv2.INTER_NEAREST)
        assert lvl.shape == ref.shape and H.bits_equal(lvl, ref), (W, Hh, l)
    if (W, Hh) == (346, 260):
        assert not np.array_equal(eo.pyramid(f)[2], f[::4, ::4][:65, :86])   # the case the old restatement got wrong


def test_negative_event_coordinates_wrap_like_numpy():
    """frame[y, x] += polarity with numpy indexing (reference event.py:118-120): -1 is the last row / column."""
    cnt = eo.accumulate([-1, 3, -96], [-1, -64, 2], [1, 0, 1], 96, 64)
    assert cnt[63, 95] == 1 and cnt[0, 3] == -1 and cnt[2, 0] == 1 and np.abs(cnt).sum() == 3
    with pytest.raises(IndexError):
        eo.accumulate([-97], [0], [1], 96, 64)


def test_empty_and_degenerate_frames():
    K = np.array([50.0, 0, 48, 0, 50.0, 32, 0, 0, 1.0]).reshape(3, 3)
    D = np.zeros(5)
    s, u = eo.event_frame([], [], [], 96, 64, K, D)
    assert s.shape == (1, 64, 96) and not s.any() and not u.any()          # zero norm -> zero frame, no NaN
    # +1 and -1 on the same pixel cancel exactly
    s, _ = eo.event_frame([5, 5], [7, 7], [1, 0], 96, 64, K, D)
    assert not s.any()


def test_matches_the_references_own_event_code():
    """Golden frames made by the reference's load_events_from_txt + EventFrame (numpy loop + OpenCV)."""
    g = np.load(GOLD)
    tab, W, Hh, NPK = g["table"], int(g["W"]), int(g["H"]), int(g["NPK"])
    pk = eo.packetise(tab, NPK)
    assert len(pk) == int(g["n_packets"]) == 2                              # the 700-event tail is dropped
    for i, pkt in enumerate(pk):
        assert eo.packet_duration(pkt) == g["durations"][i] and eo.packet_time(pkt) == g["times"][i]
        s, u = eo.event_frame(pkt[:, 1], pkt[:, 2], pkt[:, 3], W, Hh, g["K"], g["D"])
        assert H.bits_equal(s, g[f"sign_{i}"]) and H.bits_equal(u, g[f"unsign_{i}"])
        for l, lvl in enumerate(eo.pyramid(s[0])):
            assert H.bits_equal(lvl, g[f"sign_{i}_L{l}"])


def test_host_event_parser_matches_reference_semantics(tmp_path):
    """utils.event_camera.event (the product's host mirror): same packets, durations, mid-times as the
    reference produced for the golden file; Event objects materialise on demand."""
    from utils.event_camera.event import EventArray, Event, load_events_from_txt
    g = np.load(GOLD)
    path = tmp_path / "events.txt"
    np.savetxt(path, g["table"], fmt="%d", delimiter=" ")
    arrays = load_events_from_txt(str(path), int(g["NPK"]))
    assert len(arrays) == 2 and all(a.size() == int(g["NPK"]) for a in arrays)
    assert [a.duration() for a in arrays] == list(g["durations"]) and [a.time() for a in arrays] == list(g["times"])
    e0 = arrays[0].events[0]
    assert (e0.ts, e0.x, e0.y, e0.polarity) == tuple(int(v) for v in g["table"][0])
    assert len(load_events_from_txt(str(path), int(g["NPK"]), array_nums=1)) == 1
    later = load_events_from_txt(str(path), 1000, start_time=int(g["table"][3000, 0]))
    assert later[0].columns()[0][0] >= int(g["table"][3000, 0])
    a = EventArray()
    assert a.size() == 0 and a.duration() == 0
    a.callback(Event(x=1, y=2, ts=10, polarity=1)); a.callback(Event(x=3, y=4, ts=30, polarity=0))
    assert a.size() == 2 and a.duration() == 20 / 1e6 and a.time() == 20 / 1e6
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.txt"
        bad.write_text("1 2 3\n4 5 6\n")
        load_events_from_txt(str(bad), 1)
