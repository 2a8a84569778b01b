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
