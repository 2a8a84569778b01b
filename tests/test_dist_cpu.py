"""world_size-2 / 3 gloo tests on CPU of the host-side logic of the partitioned path: the slab split of the particles
(hot_b200/dist.py), the library's shared-page tables (hot_share_tables = the code dist_after_sort runs after its all-gather) and
the exchange protocol (pack -> neighbour send / recv -> add in ascending rank order) emulated over gloo."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from hot_b200.dist import split_slabs

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def launch(kind, world, out_dir, timeout=300, env=None):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), kind, str(r), str(world), str(port), str(out_dir)],
                              env=dict(os.environ, **(env or {})))
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=timeout) == 0
    return [np.load(os.path.join(out_dir, f"rank{r}.npz")) for r in range(world)]


def test_split_slabs_properties():
    rng = np.random.default_rng(1)
    X = rng.random((10007, 3))
    for world in (1, 2, 3, 8):
        parts = split_slabs(X, world)
        allp = np.concatenate(parts)
        assert len(allp) == len(X) and len(np.unique(allp)) == len(X)                 # a partition
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1            # balanced
        for a, b in zip(parts, parts[1:]):
            assert X[a, 1].max() <= X[b, 1].min()                                      # slabs along y
        assert all((np.diff(p) > 0).all() for p in parts)                              # original order kept inside a slab


@pytest.mark.parametrize("world", [2, 3])
def test_shared_page_tables_and_exchange_protocol_gloo(tmp_path, world):
    res = launch("cpu", world, tmp_path)
    assert all(bool(r["ok"]) for r in res)
    assert all(int(r["n_shared"]) > 0 and int(r["n_nbr"]) == world - 1 for r in res)


# ---- ghost ring / page authority of the partitioned assembled-matrix path (host logic of dist.cu, no device) -----------------
XMASK, YMASK, ZMASK = 0x9249249249249800, 0x4924924924924600, 0x2492492492492180   # SPGrid_Mask<7,.,3,12> (fp64 GridState)


def _spread(v, mask):
    out, bit = 0, 0
    for k in range(64):
        if mask >> k & 1:
            out |= ((v >> bit) & 1) << k
            bit += 1
    return out


def _page_id(px, py, pz):
    """page id (offset >> 12) of the 2x4x4-node page with page coordinates (px, py, pz)"""
    return (_spread(2 * px, XMASK) | _spread(4 * py, YMASK) | _spread(4 * pz, ZMASK)) >> 12


def _lists(sets):
    world = len(sets)
    counts = np.array([len(s) for s in sets], dtype=np.int32)
    maxp = int(counts.max())
    allp = np.full((world, maxp), 0xffffffff, dtype=np.uint32)
    for r, s in enumerate(sets):
        allp[r, :len(s)] = np.sort(np.array(sorted(s), dtype=np.uint32))
    return world, maxp, counts, np.ascontiguousarray(allp)


def test_ghost_ring_and_page_authority_host_logic():
    import ctypes as C
    from hot_b200._lib import load_library
    lib = load_library()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    # three slabs along y with one page layer of overlap, a notch so that some blocks are split between ranks
    coords = [set() for _ in range(3)]
    for r in range(3):
        for px in range(1, 7):
            for py in range(4 * r + 1, 4 * r + 6):
                for pz in range(1, 4):
                    if r == 1 and px == 6 and py == 5:
                        continue
                    coords[r].add((px, py, pz))
    sets = [{_page_id(*c) for c in cs} for cs in coords]
    inv = {_page_id(*c): c for cs in coords for c in cs}
    world, maxp, counts, allp = _lists(sets)
    union = set().union(*sets)
    ext = []
    for r in range(world):
        n = C.c_int()
        assert lib.hot_halo_pages(r, world, maxp, vp(counts), vp(allp), C.byref(n), None) == 0
        out = np.zeros(max(1, n.value), dtype=np.uint32)
        assert lib.hot_halo_pages(r, world, maxp, vp(counts), vp(allp), C.byref(n), vp(out)) == 0
        e = set(int(x) for x in out[:n.value])
        ext.append(e)
        assert not (e & sets[r]) and e <= union                              # ghost pages: not mine, active somewhere
        assert list(out[:n.value]) == sorted(e)
        shared = {p for p in sets[r] if any(p in sets[q] for q in range(world) if q != r)}
        assert shared
        for p in shared:                                                     # the 27-neighbourhood of a shared page is held
            x, y, z = inv[p]
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        q = _page_id(x + dx, y + dy, z + dz)
                        if q in union:
                            assert q in sets[r] or q in e
        private = sets[r] - shared
        assert all(all(_page_id(inv[p][0] + dx, inv[p][1] + dy, inv[p][2] + dz) not in e for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1))
                   for p in private if all(_page_id(inv[p][0] + dx, inv[p][1] + dy, inv[p][2] + dz) not in shared
                                           for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)))   # nothing grows far from the seam
    pids = np.array(sorted(union), dtype=np.uint32)
    auth = np.zeros(len(pids), dtype=np.int32)
    assert lib.hot_page_authority(world, maxp, vp(counts), vp(allp), len(pids), vp(pids), vp(auth)) == 0
    a = dict(zip((int(p) for p in pids), (int(v) for v in auth)))
    for p in union:
        x, y, z = inv[p]
        partner = _page_id(x ^ 1, y, z)                                      # the other 2-node half of the 4^3 Gauss-Seidel block
        holders = [r for r in range(world) if p in sets[r]]
        both = [r for r in holders if partner in sets[r]]
        assert a[p] in holders                                               # the authority activates the page itself
        if both:
            assert a[p] == min(both) and a[partner] == a[p]                  # one rank sweeps the whole block
        else:
            assert a[p] == min(holders)
