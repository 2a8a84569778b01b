"""profiling driver: builds the C2 hierarchy and times one smoother call per level (hot_op_bench); HOT_GX_DIAG etc. are read by the library"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hot_b200
import bench

sc, _ = bench.make_workload("c2")
sim = hot_b200.MpmSimulationB200(sc["dx"])
sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
sim.set_dt_gravity(bench.SOLVER_DT, (0, 0, 0))
sim.sortParticlesAndPolluteGrid(); sim.particlesToGrid()
sim.gridToParticles(0.0)
bc = bench.end_cap_bc(sim.get_id2coord())
sim.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
sim.backupStrain(); sim.updateState()
sim.buildMatrix(True); sim.buildMultigrid(levels=3, smoother=5, coarseSolver=2, Ainv=1, times=1)
print("gs_smooth ms per level", [round(sim.op_bench("smooth", 10, level=l), 4) for l in range(3)], "env", {k: v for k, v in os.environ.items() if k.startswith("HOT_")})
