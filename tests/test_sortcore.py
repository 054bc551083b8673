"""CPU tests of the bucket binning scheme (csrc/bucketbin.cu) through the host/device code it is built from
(csrc/sortcore.cuh, compiled here with g++ via tests/sortcore_host.cpp): the 8-key network, the merge-path rounds,
and the whole data flow — unstable scatter of the visible pairs into (1 << s)^2-tile buckets, merge sort on
(depth bits << 32 | Gaussian index), stable filter of every sorted bucket into its tiles — against the reference's
ordering rule: a tile's list holds the Gaussians whose rect covers it, ordered by (depth bits, Gaussian index)
(dgr/cuda_rasterizer/rasterizer_impl.cu:70-138, 303-321).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sortcore") / "libsortcore_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "sortcore_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.sc_sort.restype = C.c_int
    lib.sc_sort.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.sc_bin_view.restype = C.c_longlong
    lib.sc_bin_view.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p]
    return lib


def test_network_sorts_every_zero_one_input(sc):
    assert sc.sc_network_failures() == 0


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 63, 64, 65, 1000, 4096, 5777, 12345])
@pytest.mark.parametrize("threads", [32, 512])
def test_merge_rounds_sort(sc, n, threads):
    rng = np.random.default_rng(n * 7 + threads)
    # few distinct depth values: long runs of equal depth bits whose order is decided by the index
    depth = rng.integers(0x3E4CCCCD, 0x3E4CCCCD + max(1, n // 20 + 1), size=n, dtype=np.uint64)
    ids = rng.permutation(max(n, 1))[:n].astype(np.uint64)
    keys = (depth << np.uint64(32)) | ids
    want = np.sort(keys)
    got = keys.copy()
    sc.sc_sort(got.ctypes.data, n, threads)
    assert np.array_equal(got, want)


def _reference_lists(rect, depth, gx, gy, y0, y1):
    x0, yy0, x1, yy1 = rect & 255, rect >> 8 & 255, rect >> 16 & 255, rect >> 24
    order = np.lexsort((np.arange(rect.size), depth))   # depth bits, then index
    lists, ranges, total = [], np.zeros((gx * gy, 2), np.uint32), 0
    for ty in range(gy):
        for tx in range(gx):
            cov = (rect != 0) & (x0 <= tx) & (tx < x1) & (yy0 <= ty) & (ty < yy1)
            ids = order[cov[order]]
            if ids.size:
                ranges[ty * gx + tx] = (total, total + ids.size)
                total += ids.size
                lists.append(ids)
    return (np.concatenate(lists).astype(np.uint32) if lists else np.zeros(0, np.uint32)), ranges


def _scene(rng, P, gx, gy, y0, y1, big=3):
    cx, cy = rng.integers(0, gx, P), rng.integers(0, gy, P)
    w, h = rng.integers(1, 4, P), rng.integers(1, 4, P)
    w[:big], h[:big] = gx, gy          # screen-filling splats
    x0, x1 = np.clip(cx - w // 2, 0, gx), np.clip(cx - w // 2 + w, 0, gx)
    ry0, ry1 = np.clip(cy - h // 2, y0, y1), np.clip(cy - h // 2 + h, y0, y1)   # the projection clips rects to the strip
    vis = (x1 > x0) & (ry1 > ry0) & (rng.random(P) < 0.8)
    rect = np.where(vis, x0 | ry0 << 8 | x1 << 16 | ry1 << 24, 0).astype(np.uint32)
    depth = rng.integers(0x3E4CCCCD, 0x3E4CCCCD + 40, P).astype(np.uint32)       # many depth ties
    return rect, depth


@pytest.mark.parametrize("s", [0, 1])
@pytest.mark.parametrize("gx,gy,y0,y1", [(10, 8, 0, 8), (11, 7, 0, 7), (20, 15, 3, 10), (20, 15, 4, 5), (5, 3, 0, 0)])
def test_bucket_binning_gives_the_reference_lists(sc, s, gx, gy, y0, y1):
    rng = np.random.default_rng(gx * 100 + gy * 10 + s + y0)
    P = 3000
    rect, depth = _scene(rng, P, gx, gy, y0, y1)
    want_list, want_ranges = _reference_lists(rect, depth, gx, gy, y0, y1)
    for trial in range(2):
        order = rng.permutation(P).astype(np.int32)   # arrival order of the scatter's atomics: must not matter
        got = np.zeros(max(1, want_list.size + 16), np.uint32)
        ranges = np.zeros((gx * gy, 2), np.uint32)
        n = sc.sc_bin_view(P, rect.ctypes.data, depth.ctypes.data, order.ctypes.data, gx, gy, s, y0, y1, 64, got.ctypes.data, ranges.ctypes.data)
        assert n == want_list.size
        assert np.array_equal(got[:n], want_list)
        assert np.array_equal(ranges, want_ranges)
