"""Cuts the part of the reference's Poisson-disk tile that the C1 / C2 scenes need and stores it as hot_b200/data/poisson_brick.npz.

The reference samples every analytic level set from Data/MpmParticles/particles-1000k.dat (1 084 359 float32 points in
[-60, 60]^3, minimum distance 1; Lib/Ziran/Math/Geometry/PoissonDisk.h:185-222).  /root/reference does not exist on the GPU box and
the whole tile is 13 MB, so the brick 0 <= x, z <= 60 (all y) travels instead: sampleFromPeriodicData only ever uses tile points
with 0 <= t_d <= side_d / min_distance on an axis whose side is shorter than 60 min_distance, which holds for x and z of the C1 box
and the C2 bar.  File order is kept (the reference's sample order follows it).  Run here, with /root/reference present:
    python tests/golden/make_poisson_brick.py
"""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = "/root/reference/Data/MpmParticles/particles-1000k.dat"

with open(SRC, "rb") as f:
    count, elem = np.frombuffer(f.read(16), dtype=np.uint64)
    pts = np.frombuffer(f.read(), dtype=np.float32).reshape(-1, 3)
assert count == len(pts) == 1084359 and elem == 12
keep = (pts[:, 0] >= 0) & (pts[:, 2] >= 0)
out = os.path.join(ROOT, "hot_b200", "data", "poisson_brick.npz")
np.savez_compressed(out, points=pts[keep], tile_count=np.uint64(count), x_range=np.array([0.0, 60.0]), z_range=np.array([0.0, 60.0]))
print(out, int(keep.sum()), "points", os.path.getsize(out) / 1e6, "MB")
