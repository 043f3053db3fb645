"""CPU (gloo, world_size 2): the multi-rank host logic of bench.py -- disjoint per-rank inputs, max-over-ranks timing,
whole-job aggregate.  The data path itself has no collective (independent 30 s chunks per GPU)."""
import json
import os
import subprocess
import sys

from wbtest import ROOT


def test_two_ranks_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert j["world"] == 2
    a, b = sorted(j["ranks"], key=lambda x: x["rank"])
    assert a["digest"] != b["digest"]                    # different chunks on different ranks
    assert a["dt"] == b["dt"]                            # every rank sees the same (max) time
    assert abs(a["dt"] - max(a["dt_local"], b["dt_local"])) < 1e-9
    assert b["dt_local"] > a["dt_local"]
    assert abs(j["value"] - 2.0 / a["dt"]) < 1e-6        # whole-job audio seconds / max time
