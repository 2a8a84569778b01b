"""whole implicit substeps on the C2 workload with the reference's solver configurations (tog.sh): wall-clock per substep and the
solver log.  python profiles/prof_substep.py [steps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hot_b200
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc, _ = bench.make_workload("c2")
out = {}
for name, kw in (("HOT (L-BFGS + 3-level MG, tog.sh:38)", dict(lsolver=3, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, bcproject=1, usecn=1)),
                 ("PN-MGPCG (tog.sh:49)", dict(lsolver=2, mg_level=3, smoother=5, coarse_solver=2, project=1, linesearch=1, bcproject=1, usecn=1)),
                 ("PN-PCG matrix-free (tog.sh:25)", dict(lsolver=2, matfree=1, mg_level=1, project=1, linesearch=1, bcproject=0, usecn=1))):
    sim = hot_b200.MpmSimulationB200(sc["dx"])
    sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    dt = bench.SOLVER_DT
    sim.set_dt_gravity(dt, (0.0, 0.0, 0.0))
    rows = []
    for it in range(steps):
        t0 = time.perf_counter()
        sim.sortParticlesAndPolluteGrid()
        n = sim.particlesToGrid()
        if it == 0:
            bc = bench.end_cap_bc(sim.get_id2coord())
        else:
            bc = bench.end_cap_bc(sim.get_id2coord())
        sim.set_bc(bc, P=np.zeros((len(bc), 9)), dv_bc=np.zeros((len(bc), 3)))
        log = sim.backwardEulerStep(**kw)
        sim.gridToParticles(dt)
        ms = 1e3 * (time.perf_counter() - t0)
        rows.append({"substep_ms": ms, "nodes": n, "converged": bool(log["converged"]), "iterations": int(log["iterations"]),
                     "linear_iterations": int(log.get("total_linear_iterations", 0)), "residual_first_last": [float(log["residual_norm"][0]), float(log["residual_norm"][-1])]})
    out[name] = rows
print(json.dumps(out, indent=1))
