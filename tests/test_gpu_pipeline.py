"""Pipelined state exchange of the C ABI (hot_upload_state_async / hot_commit_state / hot_download_state_async / hot_wait_download)
against the plain set / get path: identical particle state after a transfer step, also when the particles are already sorted."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
import hot_b200
from hot_b200 import scenes

pytestmark = pytest.mark.gpu


def test_pipelined_state_exchange_matches_set_get():
    sc = scenes.block((10, 9, 7), 1.0 / 64, ppc=8, seed=3)
    n = len(sc["mass"])
    dt = 1e-3
    ref = hot_b200.MpmSimulationB200(sc["dx"], device=0)
    ref.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    ref.sortParticlesAndPolluteGrid(); ref.particlesToGrid(); ref.gridToParticles(dt)
    want1 = ref.get_particles(gradV=False)
    ref.sortParticlesAndPolluteGrid(); ref.particlesToGrid(); ref.gridToParticles(dt)
    want2 = ref.get_particles(gradV=False)

    sim = hot_b200.MpmSimulationB200(sc["dx"], device=0)
    sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    sim.sortParticlesAndPolluteGrid()      # particles now live in sorted order: the upload has to follow orig_id
    host_in = [torch.from_numpy(np.ascontiguousarray(sc[k])).pin_memory() for k in ("X", "V", "C", "F")]
    host_out = [torch.empty((n, c), dtype=torch.float64).pin_memory() for c in (3, 3, 9, 9)]
    ip, op = [t.data_ptr() for t in host_in], [t.data_ptr() for t in host_out]
    sim.upload_state_async(ip)
    sim.commit_state()
    sim.sortParticlesAndPolluteGrid(); sim.particlesToGrid(); sim.gridToParticles(dt, want_flags=False)
    sim.download_state_async(op)
    sim.wait_download()
    for k, t in zip(("X", "V", "C", "F"), host_out):
        np.testing.assert_allclose(t.numpy(), want1[k], rtol=0, atol=1e-13 * np.abs(want1[k]).max())
    # second step: feed the first step's output back through the pipeline while a download is still draining
    for a, b in zip(host_in, host_out):
        a.copy_(b)
    sim.upload_state_async(ip)
    sim.commit_state()
    sim.sortParticlesAndPolluteGrid(); sim.particlesToGrid(); sim.gridToParticles(dt, want_flags=False)
    sim.download_state_async(op)
    sim.wait_download()
    for k, t in zip(("X", "V", "C", "F"), host_out):
        np.testing.assert_allclose(t.numpy(), want2[k], rtol=0, atol=1e-12 * np.abs(want2[k]).max())


def test_upload_twice_without_commit_is_an_error():
    sc = scenes.block((4, 4, 4), 1.0 / 64, ppc=4, seed=0)
    sim = hot_b200.MpmSimulationB200(sc["dx"], device=0)
    sim.set_particles(sc["X"], sc["V"], sc["mass"], sc["C"], sc["F"], sc["vol"], sc["mu"], sc["lam"])
    host_in = [torch.from_numpy(np.ascontiguousarray(sc[k])).pin_memory() for k in ("X", "V", "C", "F")]
    ip = [t.data_ptr() for t in host_in]
    sim.upload_state_async(ip)
    with pytest.raises(Exception):
        sim.upload_state_async(ip)
    sim.commit_state()
