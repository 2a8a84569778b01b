"""The A/B switches of the kernels (measured alternatives kept selectable through environment variables) must stay
parity-green too: every variant re-runs the operator parity tests in a fresh process, because the switches are read once
per process.  Default configuration: plane scatter (compact records) for P2G and the force / Hessian scatters, column scatter for the CN tolerance, TMA-staged
G2P / Hessian gather, Gauss-Seidel colour phases in block-inverse form (stream through per-warp rings of TMA bulk copies, one launch per colour phase with programmatic dependent launch), stream-based residual update."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    ({"HOT_SCATTER": "column"}, "test_gpu_transfer.py test_gpu_force.py", "p2g_parity or update_state_residual_multiply or cn_tolerance"),
    ({"HOT_SCATTER": "plane"}, "test_gpu_transfer.py test_gpu_force.py", "p2g_parity or update_state_residual_multiply or cn_tolerance"),
    ({"HOT_G2P_TMA": "0", "HOT_HG_TMA": "0", "HOT_PF_DIST": "0"}, "test_gpu_transfer.py test_gpu_force.py",
     "g2p_parity or update_state_residual_multiply"),
    ({"HOT_GS_COOP": "1"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),       # block-inverse form, one cooperative launch per sweep
    ({"HOT_GX_PDL": "0"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),        # colour phases as fully serialised launches
    ({"HOT_GX_CLUSTER": "2"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),    # a block swept by a cluster of 2 CTAs (DSMEM)
    ({"HOT_GS_STREAM_UPDATE": "0"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),
    ({"HOT_GS_INV": "0"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),        # substitution form, TMA ring
    ({"HOT_GS_INV": "0", "HOT_GS_COOP": "0"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),
    ({"HOT_GS_INV": "0", "HOT_GS_RING": "0"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),   # per-lane stream loads
    ({"HOT_GS_INV": "0", "HOT_GS_RING": "0", "HOT_GS_COOP": "1"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"),
    ({"HOT_GS_INV": "0", "HOT_GS_STREAM": "0"}, "test_gpu_matrix.py", "smoother_parity or vcycle_parity"), # fixed 125-slot rows
]


@pytest.mark.gpu
@pytest.mark.parametrize("env,files,expr", VARIANTS, ids=["-".join(f"{k}={v}" for k, v in e.items()) for e, _, _ in VARIANTS])
def test_variant_parity(env, files, expr):
    e = dict(os.environ, **env)
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", expr] + [os.path.join(ROOT, "tests", f) for f in files.split()]
    r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
