"""P ranks on P GPUs of one box versus 1 rank: forces by atom, site sums, lambda trajectories,
bookkeeping totals (SURVEY.md §4 test plan, §8e).  Needs >= 2 GPUs (gpurun --gpus 2)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,mode", [(2, "pair"), (2, "bonded"), (2, "ljstates"), (2, "ewald"), (4, "pair"), (8, "pair")])
def test_multi_rank_matches_single_rank(built, world, mode):
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + {"pair": 0, "bonded": 50, "ljstates": 70, "ewald": 90}[mode]),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), "1.0" if world <= 2 else "2.0", "70" if mode == "ljstates" else "40", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-4000:])
    out = json.loads(line[0][len("MGPU_RESULT "):])
    w = out["worst"]
    assert w["f"] <= 1e-10 and w["dudl"] <= 1e-10 and w["e"] <= 1e-10 and w["lam"] <= 1e-8, out
    assert out["nlocal"] == out["ref_nlocal"] and out["neighbors"] == out["ref_neighbors"]
    assert out["titr"] == out["ref_titr"] and out["builds"] == out["ref_builds"] and out["builds"] > 2
