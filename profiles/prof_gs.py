"""profiling driver: builds the C2 hierarchy and runs a few V-cycles (used under ncu; never a bench value)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hot_b200
from hot_b200 import scenes
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
sc, _ = bench.make_workload(wl)
sim = hot_b200.MpmSimulationB200(sc["dx"])
sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
sim.set_dt_gravity(bench.SOLVER_DT, (0, 0, 0))
sim.sortParticlesAndPolluteGrid(); sim.particlesToGrid()
sim.gridToParticles(0.0)
bc = bench.end_cap_bc(sim.get_id2coord())
sim.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
sim.backupStrain(); sim.updateState()
sim.multiply(np.ones((sim.num_nodes, 3)))
r = sim.computeResidual()
sim.buildMatrix(True); sim.buildMultigrid(levels=3)
sim.vcycle(r)
print("vcycle ms", sim.vcycle_bench(3))
sim.gridToParticles(0.0)
