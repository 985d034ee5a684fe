"""CPU, world_size 2 on gloo: the shard-by-game partition and the whole-job aggregation bench.py
uses at N > 1 (no data-path collective exists to test)."""
import os
import subprocess
import sys
import textwrap

import harness as H

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from kami_b200 import parallel as P
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    red = P.Reducer(dist, "cpu")
    games = P.local_games(37, world, rank)
    assert all(P.game_to_rank(g, world) == rank for g in games)
    total = red.sum(len(games))
    red.barrier()
    value, units, ms = P.whole_job_throughput(red, local_units=1000.0 * (rank + 1), local_ms=10.0 * (rank + 1))
    seeds = red.sum(P.per_rank_seed(7, rank))
    if rank == 0:
        print(json.dumps({"total": total, "value": value, "units": units, "ms": ms, "seeds": seeds}))
    dist.destroy_process_group()
''')


def test_two_rank_partition_and_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % H.ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-800:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["total"] == 37                      # every game owned exactly once
    assert r["units"] == 3000.0 and r["ms"] == 20.0
    assert abs(r["value"] - 3000.0 / 0.020) < 1e-6   # all units / MAX over ranks of the device time
    assert r["seeds"] == 7 + 7 + 1000003


def test_partition_is_exact():
    from kami_b200 import parallel as P

    for world in (1, 2, 4, 8):
        seen = sorted(g for r in range(world) for g in P.local_games(8192, world, r))
        assert seen == list(range(8192))
        assert all(len(P.local_games(8192, world, r)) == 8192 // world for r in range(world))
