"""N > 1 path on CPU: two gloo ranks, contigs sharded by LPT, exact reductions, merge on rank 0."""
import os
import socket
import subprocess
import sys

import pytest

from phaser_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_plan_is_balanced_and_complete():
    w = [248, 242, 198, 190, 181, 170, 159, 145, 138, 133, 135, 133, 114, 107, 101, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    plan = shard.plan_shards(w, 8)
    assert sorted(c for p in plan for c in p) == list(range(24))
    loads = [sum(w[c] for c in p) for p in plan]
    assert max(loads) <= 1.15 * sum(w) / 8


@pytest.mark.parametrize("case", ["rna_two_bams", "rna_conflict", "quirks"])
def test_two_rank_sharded_run_matches_reference(case):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "shard_worker.py"), case], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED PARITY OK" in outs[0]
