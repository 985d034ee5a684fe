"""CPU, world_size 2 on gloo: the shard-by-game partition and the whole-job aggregation bench.py uses at N > 1
(the self-play path has no data-path collective), and the gradient averaging of the data-parallel training step
(the one collective of the system; NCCL on the GPUs, the same all-reduce on gloo here) checked against the
training oracle: the mean of two ranks' mini-batch gradients."""
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
    # data-parallel training: each rank differentiates its own mini-batch (training oracle), the flat gradient
    # vectors are all-reduced and averaged, every replica takes the same SGD step
    sys.path.insert(0, os.path.join(%r, "oracle")); sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, harness as H, nn_oracle as NO, train_oracle as TO
    F, R, n = 64, 1, 6
    params = NO.init_params(F, R, seed=5)
    envs = H.sample_positions(n, seed=70 + rank)
    obs = np.stack([e.observe() for e in envs])
    pi, z = TO.synthetic_targets(n, 80 + rank, [e.actions() for e in envs])
    _, loss, grads = TO.train_step(params, obs, pi, z, F, R, 0.0)
    names = [k for k, _ in NO.param_order(F, R) if TO.trainable(k)]
    flat = np.concatenate([grads[k].reshape(-1) for k in names]).astype(np.float32)
    avg = P.average_host_gradients(dist, flat.copy())
    other = np.zeros_like(flat)
    import torch
    gathered = [torch.zeros(flat.size) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(flat))
    want = sum(g.numpy() for g in gathered) / world
    gerr = float(np.abs(avg - want).max())
    if rank == 0:
        print(json.dumps({"total": total, "value": value, "units": units, "ms": ms, "seeds": seeds, "grad_err": gerr,
                          "grad_norm": float(np.linalg.norm(avg))}))
    dist.destroy_process_group()
''')


def test_two_rank_partition_and_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (H.ROOT, H.ROOT, H.ROOT))
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
    assert r["grad_err"] <= 1e-7 and r["grad_norm"] > 0   # all-reduce(sum) / world == mean of the ranks' gradients


def test_partition_is_exact():
    from kami_b200 import parallel as P

    for world in (1, 2, 4, 8):
        seen = sorted(g for r in range(world) for g in P.local_games(8192, world, r))
        assert seen == list(range(8192))
        assert all(len(P.local_games(8192, world, r)) == 8192 // world for r in range(world))
