"""CPU model of the engine's binning scheme (csrc/preprocess.cu + csrc/tilebin.cu) against the
reference's ordering rule, on random inputs with many depth ties.

Reference (rasterizer_impl.cu:70-138, 303-311): per view, every visible Gaussian i emits one key
(tile << 32 | depth bits) per tile of its rect (row-major), in index order; a STABLE sort on the key gives the
per-tile lists, i.e. inside a tile the order is (depth bits, Gaussian index).

Engine: visible (view, Gaussian) pairs compacted in index order -> stable sort on (view << 31 | depth bits) -> the
instance sequence (pairs in that order, rect row-major) is cut into chunks of CH instances that never straddle the view
boundary -> per-chunk per-tile counts -> prefix over chunks and tiles -> every instance goes to
ranges[tile].x + base[chunk][tile] + rank, where the rank comes from 8 slices per chunk walked in order, 32 instances
per step, equal tiles inside a step ranked by lane.  The model follows those steps literally (small CH so that chunk,
slice and step boundaries all occur) and must reproduce the reference's lists and ranges exactly.
"""
import numpy as np
import pytest


def reference_lists(depth_bits, rects, gx, tiles):
    """depth_bits[P] (0 = culled), rects[P] = (x0, y0, x1, y1).  Returns (point_list, ranges[tiles, 2])."""
    keys, vals = [], []
    for i, (d, (x0, y0, x1, y1)) in enumerate(zip(depth_bits, rects)):
        if d == 0:
            continue
        for y in range(y0, y1):
            for x in range(x0, x1):
                keys.append(((y * gx + x) << 32) | int(d))
                vals.append(i)
    keys, vals = np.array(keys, np.uint64), np.array(vals, np.uint32)
    order = np.argsort(keys, kind="stable")
    keys, vals = keys[order], vals[order]
    ranges = np.zeros((tiles, 2), np.uint32)
    t = (keys >> np.uint64(32)).astype(np.int64)
    for k in range(len(keys)):
        if k == 0 or t[k] != t[k - 1]:
            ranges[t[k], 0] = k
        if k == len(keys) - 1 or t[k] != t[k + 1]:
            ranges[t[k], 1] = k + 1
    return vals, ranges


def engine_lists(depth_bits2, rects2, P, gx, tiles, CH=64):
    """Both views at once, as the engine does.  depth_bits2[2][P], rects2[2][P].  Returns per view (list, ranges) with
    the ranges rebased to the view's own list, like gsevt_engine_binning."""
    SLICES, STEP = 8, 4          # 8 slices per chunk; a 'warp step' of 4 lanes keeps the model small but multi-step
    assert CH % (SLICES * STEP) == 0
    # compaction in index order, sort key view << 31 | depth
    ids = [v * P + i for v in (0, 1) for i in range(P) if depth_bits2[v][i] != 0]
    key = np.array([(j // P) << 31 | int(depth_bits2[j // P][j % P]) for j in ids], np.uint64)
    ids = np.array(ids, np.int64)[np.argsort(key, kind="stable")]
    # instance sequence
    inst_tile, inst_id, inst_view = [], [], []
    for j in ids:
        v, i = divmod(int(j), P)
        x0, y0, x1, y1 = rects2[v][i]
        for y in range(y0, y1):
            for x in range(x0, x1):
                inst_tile.append(y * gx + x)
                inst_id.append(i)
                inst_view.append(v)
    inst_tile, inst_id, inst_view = np.array(inst_tile, np.int64), np.array(inst_id, np.uint32), np.array(inst_view, np.int64)
    N0 = int((inst_view == 0).sum())
    total = len(inst_tile)
    chunks = []                  # (view, begin, end): never straddle the view boundary
    for v, lo, hi in ((0, 0, N0), (1, N0, total)):
        for b in range(lo, hi, CH):
            chunks.append((v, b, min(b + CH, hi)))
    hist = np.zeros((len(chunks), tiles), np.int64)
    for c, (v, b, e) in enumerate(chunks):
        np.add.at(hist[c], inst_tile[b:e], 1)
    # prefix over the chunks of a view, per tile; totals; ranges over (view, tile)
    base = np.zeros_like(hist)
    tot = np.zeros((2, tiles), np.int64)
    for c, (v, b, e) in enumerate(chunks):
        base[c] = tot[v]
        tot[v] += hist[c]
    flat = tot.reshape(-1)
    start = np.concatenate([[0], np.cumsum(flat)[:-1]])
    ranges = np.where(flat[:, None] > 0, np.stack([start, start + flat], 1), 0).reshape(2, tiles, 2)
    values = np.full(total, 0xFFFFFFFF, np.uint32)
    for c, (v, b, e) in enumerate(chunks):
        n = e - b
        sl = CH // SLICES
        cnt = np.zeros((SLICES, tiles), np.int64)
        for li in range(n):
            cnt[li // sl, inst_tile[b + li]] += 1
        cursor = np.cumsum(cnt, 0) - cnt                      # exclusive prefix over the slices
        for w in range(SLICES):                               # every 'warp' walks its slice in order
            for s0 in range(w * sl, min((w + 1) * sl, n), STEP):
                lanes = list(range(s0, min(s0 + STEP, (w + 1) * sl, n)))
                seen = {}
                for li in lanes:                              # equal tiles inside a step: rank by lane
                    t = inst_tile[b + li]
                    r = seen.get(t, 0)
                    seen[t] = r + 1
                    values[ranges[v, t, 0] + base[c, t] + cursor[w, t] + r] = inst_id[b + li]
                for t, k in seen.items():
                    cursor[w, t] += k
    out = []
    for v in (0, 1):
        first = 0 if v == 0 else N0
        count = N0 if v == 0 else total - N0
        rr = np.where(ranges[v][:, 1:2] > ranges[v][:, 0:1], ranges[v] - first, 0).astype(np.uint32)
        out.append((values[first:first + count], rr))
    return out


@pytest.mark.parametrize("seed,P,gx,gy,CH", [(0, 60, 4, 3, 32), (1, 200, 5, 4, 64), (2, 300, 8, 6, 128), (3, 97, 3, 3, 32)])
def test_counting_partition_reproduces_the_reference_order(seed, P, gx, gy, CH):
    rng = np.random.default_rng(seed)
    tiles = gx * gy
    depth_bits2, rects2 = [], []
    for v in (0, 1):
        # few distinct depths -> many ties, which the index order must break; ~25 % culled
        d = rng.integers(1, 12, P).astype(np.uint32) + np.uint32(0x3F000000)
        d[rng.random(P) < 0.25] = 0
        x0 = rng.integers(0, gx, P); y0 = rng.integers(0, gy, P)
        x1 = np.minimum(gx, x0 + rng.integers(1, 4, P)); y1 = np.minimum(gy, y0 + rng.integers(1, 4, P))
        big = rng.random(P) < 0.05                              # a few screen-filling Gaussians
        x0[big], y0[big], x1[big], y1[big] = 0, 0, gx, gy
        depth_bits2.append(d)
        rects2.append(list(zip(x0.tolist(), y0.tolist(), x1.tolist(), y1.tolist())))
    ours = engine_lists(depth_bits2, rects2, P, gx, tiles, CH)
    for v in (0, 1):
        ref_list, ref_ranges = reference_lists(depth_bits2[v], rects2[v], gx, tiles)
        assert np.array_equal(ours[v][0], ref_list), f"view {v}: per-tile lists differ"
        assert np.array_equal(ours[v][1], ref_ranges), f"view {v}: tile ranges differ"


def test_empty_and_single_view_inputs():
    P, gx, gy = 16, 2, 2
    none = [np.zeros(P, np.uint32), np.zeros(P, np.uint32)]
    rects = [[(0, 0, 1, 1)] * P, [(0, 0, 1, 1)] * P]
    for v, (lst, rr) in enumerate(engine_lists(none, rects, P, gx, gx * gy, 32)):
        assert len(lst) == 0 and not rr.any()
    only1 = [np.zeros(P, np.uint32), np.full(P, 0x3F800000, np.uint32)]   # nothing visible in view 0, all ties in view 1
    ours = engine_lists(only1, rects, P, gx, gx * gy, 32)
    ref_list, ref_ranges = reference_lists(only1[1], rects[1], gx, gx * gy)
    assert len(ours[0][0]) == 0 and np.array_equal(ours[1][0], ref_list) and np.array_equal(ours[1][1], ref_ranges)
    assert np.array_equal(ref_list, np.arange(P, dtype=np.uint32))          # ties keep the index order
