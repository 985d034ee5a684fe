"""Generates tests/golden/arena_ref.json from the UNMODIFIED reference arena (kami/evaluate.cpp compiled in
oracle/_ref/libkami_ref_arena.so, oracle/ref_shim/ref_arena.cpp) driven with the injected pseudo-network.
Run in the build container (needs /root/reference at oracle build time):  python tests/golden/make_arena_golden.py"""
import ctypes as C
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libkami_ref_arena.so"))
L.ref_arena_opt_int.argtypes = [C.c_char_p, C.c_int]
L.ref_arena_opt_float.argtypes = [C.c_char_p, C.c_float]
L.ref_arena_run.argtypes = [C.c_uint, C.c_uint32, C.c_uint32, C.c_char_p, C.c_int]

# options.def.yml values of the keys MCTS reads; noise off (mcts.h:97: the only non-deterministic term)
BASE = dict(cpuct=1.5, force_expand_unvisited=0, unvisited_node_value_pct=50, bootstrap_weight=20, bootstrap_window=1600,
            bootstrap_amp_pct=75, scale_cpuct_by_actions=0)
CASES = [
    dict(seed=1, batch=8, games=6, nodes=12, target=54, salt_cur=11, salt_cd=23),
    dict(seed=2, batch=8, games=8, nodes=20, target=54, salt_cur=5, salt_cd=9),
    dict(seed=3, batch=4, games=4, nodes=8, target=75, salt_cur=100, salt_cd=7),
    dict(seed=4, batch=6, games=5, nodes=16, target=30, salt_cur=31, salt_cd=32),
    dict(seed=5, batch=8, games=8, nodes=32, target=50, salt_cur=77, salt_cd=78),
    dict(seed=6, batch=8, games=7, nodes=24, target=60, salt_cur=3, salt_cd=4),
    dict(seed=7, batch=2, games=2, nodes=10, target=50, salt_cur=8, salt_cd=1),
    dict(seed=8, batch=8, games=4, nodes=64, target=50, salt_cur=21, salt_cd=12),
]
out = []
for c in CASES:
    for k, v in BASE.items():
        (L.ref_arena_opt_float if k == "cpuct" else L.ref_arena_opt_int)(k.encode(), v)
    L.ref_arena_opt_float(b"mcts_noise_weight", 0.0)
    L.ref_arena_opt_int(b"evaluate_batch", c["batch"])
    L.ref_arena_opt_int(b"evaluate_games", c["games"])
    L.ref_arena_opt_int(b"evaluate_nodes", c["nodes"])
    L.ref_arena_opt_int(b"evaluate_target_pct", c["target"])
    buf = C.create_string_buffer(1 << 16)
    n = L.ref_arena_run(c["seed"], c["salt_cur"], c["salt_cd"], buf, len(buf))
    assert n > 0
    text = buf.value.decode()
    games = [(float(m.group(1)), int(m.group(2))) for m in re.finditer(r"game \d+ of \d+ \[([-\d.e+]+)\]: score (-?\d+)%", text)]
    verdict = int(re.search(r"VERDICT (\d)", text).group(1))
    assert "EXCEPTION" not in text, text
    c = dict(c, options=dict(BASE, mcts_noise_weight=0.0), games_log=games, verdict=verdict,
             ending=("early pass" if "finished evaluating early" in text else "early fail" if "aborting evaluation" in text else "full"))
    print(c["seed"], len(games), verdict, c["ending"], games)
    out.append(c)
json.dump({"generator": "tests/golden/make_arena_golden.py", "source": "oracle/_ref/libkami_ref_arena.so = /root/reference/kami/evaluate.cpp (unmodified) + injected pseudo-network",
           "cases": out}, open(os.path.join(HERE, "arena_ref.json"), "w"), indent=1)
