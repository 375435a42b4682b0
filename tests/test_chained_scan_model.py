"""CPU: a line-by-line Python model of the look-back of `scan_chained_kernel` (metrics_b200/csrc/curve.cu) under EVERY kind of
interleaving the GPU run only samples: each predecessor tile is observed either as an aggregate or as an inclusive prefix
(tile 0 always a prefix), in windows of 32 with arbitrary "not published yet" stalls.  The carries the model arrives at —
positives before the tile, group ends before it, TP / FP at the last group end before it — must equal the directly computed
ones.  This pins the state algebra (the `pending` resolution of an aggregate's local TP through the positives accumulated on
the way, prefix words overriding, tiles without any group end); the kernel itself is held to the oracle by test_curves_gpu.py.
"""
import numpy as np
import pytest

TILE = 4096


def tile_aggregates(labels, ends, n_tiles):
    """Per tile: positives, group ends, local (pos1, tp) of its last group end (0, 0 if none)."""
    npos, nb, pos1, tp = [], [], [], []
    for t in range(n_tiles):
        lab, e = labels[t * TILE:(t + 1) * TILE], ends[t * TILE:(t + 1) * TILE]
        npos.append(int(lab.sum()))
        nb.append(int(e.sum()))
        idx = np.flatnonzero(e)
        if idx.size:
            pos1.append(int(idx[-1]) + 1)
            tp.append(int(lab[: idx[-1] + 1].sum()))
        else:
            pos1.append(0)
            tp.append(0)
    return npos, nb, pos1, tp


def lookback(tile, seen_as_prefix, not_ready_until, agg, pre):
    """The warp-0 loop of the kernel for `tile` > 0.  agg / pre: per tile (sums hi, sums lo, last hi, last lo) as published;
    seen_as_prefix[j]: which generation this reader observes for tile j; not_ready_until[j]: how many polls tile j stays
    unpublished (models the spin).  Returns (sp, sb, f_tp, f_pos1)."""
    sp = sb = f_tp = f_pos1 = 0
    found = pending = False
    p_tp_l = p_pos1_g = p_acc = 0
    j = tile - 1
    polls = {}
    while True:
        lanes = []
        for lane in range(32):
            jj = j - lane
            if jj < 0:
                lanes.append(("P", 0, 0, 0, 0))  # in front of the segment: an empty prefix
                continue
            polls[jj] = polls.get(jj, 0) + 1
            if polls[jj] <= not_ready_until[jj]:
                lanes.append((None, 0, 0, 0, 0))
            elif seen_as_prefix[jj]:
                lanes.append(("P",) + pre[jj])
            else:
                lanes.append(("A",) + agg[jj])
        first_not = next((i for i, l in enumerate(lanes) if l[0] is None), 32)
        usable = range(first_not)
        stop = next((i for i in usable if lanes[i][0] == "P"), -1)
        consumed = range(stop + 1) if stop >= 0 else usable
        incl_np = np.cumsum([lanes[i][1] if i in consumed else 0 for i in range(32)])
        tot_nb = sum(lanes[i][2] for i in consumed)
        if not found and not pending:
            hb = [i for i in consumed if lanes[i][3] != 0]
            if hb:
                ls = hb[0]
                if lanes[ls][0] == "P":
                    found, f_tp, f_pos1 = True, lanes[ls][4], lanes[ls][3]
                else:
                    pending, p_tp_l = True, lanes[ls][4]
                    p_pos1_g = (j - ls) * TILE + lanes[ls][3]
                    p_acc = sp + int(incl_np[ls])
        sp += int(incl_np[31])
        sb += tot_nb
        if stop >= 0:
            break
        j -= len(consumed)
    if pending:
        f_tp, f_pos1 = sp - p_acc + p_tp_l, p_pos1_g
    return sp, sb, f_tp, f_pos1


@pytest.mark.parametrize("seed", range(12))
def test_lookback_carries_equal_the_direct_ones(seed):
    g = np.random.default_rng(seed)
    n_tiles = int(g.integers(2, 90))
    n = n_tiles * TILE - int(g.integers(0, TILE))
    labels = (g.random(n_tiles * TILE) < g.choice([0.02, 0.5, 0.9])).astype(np.int64)
    labels[n:] = 0
    # group ends: dense, sparse (long tie runs spanning several tiles), or none at all in stretches
    density = g.choice([0.9, 0.01, 0.0003])
    ends = g.random(n_tiles * TILE) < density
    ends[n:] = False
    ends[n - 1] = True
    if seed % 3 == 0:
        ends[TILE // 2: 5 * TILE] = False  # several tiles without any group end
    npos, nb, pos1, tp = tile_aggregates(labels, ends, n_tiles)
    agg = [(npos[t], nb[t], pos1[t], tp[t]) for t in range(n_tiles)]
    # inclusive prefixes as the kernel publishes them
    pre, cp, cb, last_pos1, last_tp = [], 0, 0, 0, 0
    for t in range(n_tiles):
        if pos1[t]:
            last_pos1, last_tp = t * TILE + pos1[t], cp + tp[t]
        cp, cb = cp + npos[t], cb + nb[t]
        pre.append((cp, cb, last_pos1, last_tp))
    cum_lab = np.concatenate([[0], np.cumsum(labels)])
    for tile in range(1, n_tiles):
        seen = g.random(n_tiles) < g.choice([0.0, 0.1, 0.5, 1.0])
        seen[0] = True
        stall = np.where(g.random(n_tiles) < 0.2, g.integers(1, 4, n_tiles), 0)
        sp, sb, f_tp, f_pos1 = lookback(tile, seen, stall, agg, pre)
        start = tile * TILE
        assert sp == cum_lab[start] and sb == int(ends[:start].sum())
        idx = np.flatnonzero(ends[:start])
        want_pos1 = int(idx[-1]) + 1 if idx.size else 0
        assert f_pos1 == want_pos1, (tile, f_pos1, want_pos1)
        assert f_tp == (int(cum_lab[want_pos1]) if want_pos1 else 0)
