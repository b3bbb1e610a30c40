"""Shared helpers for the parity tests."""
import numpy as np


def halves_equal(a, b):
    """Compare reference-layout half-node arrays: index/leaf word exactly, xyz by value (-0 == +0)."""
    return (
        np.array_equal(a["ib"], b["ib"])
        and np.array_equal(a["x"], b["x"])
        and np.array_equal(a["y"], b["y"])
        and np.array_equal(a["z"], b["z"])
    )


def assert_tree_equal(got, want, check_keys=True):
    assert got["root"] == want["root"], (got["root"], want["root"])
    if check_keys:
        assert np.array_equal(got["keys"].astype(np.uint64), want["keys"].astype(np.uint64)), "Morton keys differ"
    assert np.array_equal(got["primitive_indices"], want["primitive_indices"]), "sorted primitive order differs"
    assert np.array_equal(got["parents"], want["parents"]), "node_parents differ"
    for name in ("node_lowers", "node_uppers"):
        for f in ("ib", "x", "y", "z"):
            bad = np.flatnonzero(got[name][f] != want[name][f])
            assert bad.size == 0, f"{name}.{f} differs at {bad[:8]} (of {bad.size})"


def visible_nodes(tree):
    """Indices of nodes reachable from the root without descending below packed leaves."""
    lo, hi = tree["node_lowers"], tree["node_uppers"]
    out, stack = [], [tree["root"]]
    while stack:
        c = stack.pop()
        out.append(c)
        if not (lo["ib"][c] >> 31):
            stack.append(int(lo["ib"][c] & 0x7FFFFFFF))
            stack.append(int(hi["ib"][c] & 0x7FFFFFFF))
    return np.array(sorted(out))


def assert_results_equal(got, want, fields):
    for f in fields:
        g, w = np.asarray(got[f]), np.asarray(want[f])
        bad = np.flatnonzero((g != w).reshape(len(g), -1).any(axis=1))
        assert bad.size == 0, f"field {f}: {bad.size} mismatches, first at {bad[:5]}: got {g[bad[:5]]} want {w[bad[:5]]}"


def random_boxes(n, seed=123, extent=10.0, size=1.0):
    rng = np.random.default_rng(seed)
    lo = (rng.random((n, 3)) * extent).astype(np.float32)
    hi = (lo + rng.random((n, 3)).astype(np.float32) * size).astype(np.float32)
    return lo, hi


def visible_mask(tree):
    """Boolean mask over nodes: reachable from the root without descending below a packed leaf.  Level-by-level and
    vectorised, for trees too large for the Python loop of visible_nodes()."""
    lo, hi = tree["node_lowers"]["ib"], tree["node_uppers"]["ib"]
    mask = np.zeros(len(lo), bool)
    frontier = np.array([tree["root"]], np.int64)
    while frontier.size:
        mask[frontier] = True
        inner = frontier[(lo[frontier] >> 31) == 0]
        frontier = np.concatenate([lo[inner] & 0x7FFFFFFF, hi[inner] & 0x7FFFFFFF]).astype(np.int64)
    return mask
