"""Multi-rank parity: one process per rank under torchrun (one GPU each when the box has them; on a 1-GPU box the
ranks SHARE the device -- CUDA-IPC and the device-side barriers work across processes of one GPU, so the whole
distributed host logic, the peer pointers and the barrier protocol are exercised there too, only NVLink is not),
AmrSim after lambrexInitParallel on every rank, compared BIT FOR BIT with the same problem run alone on one GPU
(tools/amr_dist_check.py): the uniform path with level 0 stored as one slab per rank and the face exchange fused
into the step kernel (peer stores over NVLink), the per-box distributed AMR path (2 and 3 levels, mid-run regrid),
the transition between the two when refinement appears, and the per-rank input / output API (profiles, local
arrays, local fields)."""
import ctypes
import os
import subprocess
import sys

import pytest

from lambrex_b200 import lbx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gpus():
    n = ctypes.c_int(0)
    lbx.lib().lbx_device_count(ctypes.byref(n))
    return n.value


def _torchrun(script, nproc, port, *args):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0 and nproc > _gpus() and ("busy or unavailable" in (r.stdout + r.stderr) or "exclusive" in (r.stdout + r.stderr).lower()):
        pytest.skip("the GPU is in an exclusive compute mode: %d ranks cannot share it" % nproc)
    return r


@pytest.mark.parametrize("coupling", ["rohde", "subcycle"])
def test_distributed_amrsim_bit_equal_single_gpu(coupling):
    """2 ranks: every rank owns boxes of every level.  Runs on one GPU too (the two ranks share it)."""
    r = _torchrun("amr_dist_check.py", 2, 29541 if coupling == "rohde" else 29542, *(["--subcycle"] if coupling == "subcycle" else []))
    assert r.returncode == 0 and "AMR_DIST_CHECK_OK" in r.stdout, r.stdout[-4000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-equal=True") >= 8, r.stdout[-4000:]


def test_eight_rank_layout_bit_equal_single_gpu():
    """8 ranks (sharing the visible GPUs): the layout of the 8-GPU runs -- interleaved ownership of a hierarchy's
    levels, ranks that own no box of a level, the same BoxArray distributed in two ways in one process (a uniform
    run's slabs, then a hierarchy's interleaved runs: plans are cached per ownership), equal barrier counts on
    every rank."""
    r = _torchrun("amr_dist_check.py", 8, 29544)
    assert r.returncode == 0 and "AMR_DIST_CHECK_OK" in r.stdout, r.stdout[-4000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-equal=True") >= 8 and "BARRIER COUNTS DIFFER" not in r.stdout, r.stdout[-4000:]


def test_bench_multi_gpu_goes_through_amrsim():
    """bench.py --gpus 2 on a reduced grid: the JSON line names AmrSim as the API and conserves mass."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "6", "--warmup", "3",
           "--grid-multi", "128"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert "AmrSim" in line["config"]["api"] and line["n_gpus"] == 2 and line["gpu_launches"] >= 12
    assert abs(line["check"]["total_mass_over_cells"] - 1.0) < 1e-12
    assert abs(line["check"]["ux_amplitude_over_U"] - line["check"]["ux_amplitude_expected"]) < 1e-3
