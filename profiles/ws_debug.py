"""Phase stamps of the persistent scatter (HOT_WS_DEBUG) and A/B timing of the scatter skeletons on one scene.
usage: python profiles/ws_debug.py [poisson|jitter] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hot_b200
from hot_b200 import scenes

kind = sys.argv[1] if len(sys.argv) > 1 else "poisson"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = scenes.config_c2() if kind == "poisson" else scenes.block((22, 165, 22), 0.12 / 22, ppc=12, origin_cells=(16, 16, 16), rho=2000.0, E=1e5, nu=0.3, seed=0)
stream = torch.cuda.current_stream()
sim = hot_b200.MpmSimulationB200(sc["dx"], device=0, stream=stream.cuda_stream)
sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
sim.sortParticlesAndPolluteGrid()
nn = sim.particlesToGrid()
print(kind, "particles", len(sc["mass"]), "nodes", nn, file=sys.stderr)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_(); sim.particlesToGrid(); sim.gridToParticles(0.0, want_flags=False)
sim.timing(2)
for _ in range(reps):
    flush.zero_(); sim.particlesToGrid(); sim.gridToParticles(0.0, want_flags=False)
torch.cuda.synchronize()
kt = sim.get_timings()
print({k: round(1e3 * v[0] / max(v[1], 1), 2) for k, v in kt.items() if v[1]}, "us per launch", file=sys.stderr)
