"""world_size-2 / 3 gloo tests on CPU of the host-side logic of the partitioned path (hot_b200/dist.py): the balanced
contiguous cut of the page groups and the interface-only exchange protocol (pack -> all-reduce -> unpack)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from hot_b200.dist import split_groups

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def launch(kind, world, out_dir, timeout=300):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), kind, str(r), str(world), str(port), str(out_dir)])
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=timeout) == 0
    return [np.load(os.path.join(out_dir, f"rank{r}.npz")) for r in range(world)]


def test_split_groups_properties():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        sizes = rng.integers(1, 300, size=1000)
        first = np.concatenate([[0], np.cumsum(sizes)]).tolist()
        cut = split_groups(first, first[-1], world)
        assert cut[0] == 0 and cut[-1] == len(sizes) and all(a <= b for a, b in zip(cut, cut[1:]))
        loads = [first[cut[r + 1]] - first[cut[r]] for r in range(world)]
        assert sum(loads) == first[-1]
        assert max(loads) - min(loads) <= 2 * sizes.max()            # balanced up to one group
    # degenerate: fewer groups than ranks
    cut = split_groups([0, 10, 20], 20, 8)
    assert cut[0] == 0 and cut[-1] == 2 and all(a <= b for a, b in zip(cut, cut[1:]))


@pytest.mark.parametrize("world", [2, 3])
def test_interface_exchange_protocol_gloo(tmp_path, world):
    res = launch("cpu", world, tmp_path)
    assert all(bool(r["ok"]) for r in res)
    assert all((r["cut"] == res[0]["cut"]).all() for r in res)
    assert res[0]["n_iface"] > 0
