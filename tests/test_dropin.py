"""Drop-in boundary: the reference's OWN test programs (test/*.cpp, unmodified) compiled against
this repo's kami/*.h and linked to libkami_b200.so (kami/Makefile).  CPU: they build.  GPU: the
reference's one deterministic test (test/encoding.cpp) prints byte-identical output."""
import hashlib
import json
import os
import subprocess

import pytest

import harness as H

DROPIN = os.path.join(H.ROOT, "kami", "_dropin")
TESTS = ["encoding", "mcts", "bench", "nn", "nncuda", "selfplay", "play", "nndisk", "nntrain"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference sources not present")
def test_reference_tests_compile_against_dropin_headers():
    subprocess.check_call(["make", "-C", os.path.join(H.ROOT, "kami"), "dropin"], stdout=subprocess.DEVNULL)
    for t in TESTS:
        assert os.path.exists(os.path.join(DROPIN, "test_" + t)), t


@pytest.mark.gpu
def test_reference_encoding_test_output_is_byte_identical(kb):
    exe = os.path.join(DROPIN, "test_encoding")
    if not os.path.exists(exe):
        pytest.skip("kami/_dropin not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-400:]
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoding_game.json")))
    assert hashlib.sha256(out.stdout).hexdigest() == g["sha256"]


@pytest.mark.gpu
def test_reference_nn_and_mcts_programs_run(kb):
    """test/nncuda.cpp (NN::infer throughput loop) and test/bench.cpp (select/expand/pick/push
    loop) run against the CUDA path.  test/bench.cpp and test/mcts.cpp pass a 4184-float policy
    array where PSIZE is 4672 (stale test, out-of-bounds reads for actions
    >= 4184), so their behaviour past the first moves is undefined in the reference too; only a
    prefix of the run is checked."""
    exe = os.path.join(DROPIN, "test_nncuda")
    if not os.path.exists(exe):
        pytest.skip("kami/_dropin not built")
    out = subprocess.run([exe], capture_output=True, timeout=180)
    assert out.returncode == 0, out.stderr.decode()[-400:]
    text = out.stdout.decode()
    assert text.count("pred/s") == 5 and "aborting" not in text
    try:
        out = subprocess.run([os.path.join(DROPIN, "test_bench")], capture_output=True, timeout=60)
        text = out.stdout.decode()
    except subprocess.TimeoutExpired as e:
        text = (e.stdout or b"").decode()
    assert "Observations / second" in text


@pytest.mark.gpu
def test_reference_nntrain_program_runs(kb):
    """test/nntrain.cpp, unmodified: NN::train on 512 random samples (8 epochs of 64 mini-batches of 8, options
    defaults: 256 filters, 2 residual blocks) through the CUDA training step.  The data is pure noise (5 % of the
    4672 policy targets set to 1, uniform inputs), so the loss only has to stay finite and near its start."""
    import re

    exe = os.path.join(DROPIN, "test_nntrain")
    if not os.path.exists(exe):
        pytest.skip("kami/_dropin not built")
    out = subprocess.run([exe], capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-400:]
    text = out.stdout.decode()
    ep = re.findall(r"Epoch (\d+)/8: loss ([-+.e\d]+) => ([-+.e\d]+), 64 batches", text)
    assert len(ep) == 8 and "Generated model 1" in text, text[-600:]
    first = float(ep[0][1])
    assert all(0.88 * first < float(a) < 1.12 * first and 0.88 * first < float(b) < 1.12 * first for _, a, b in ep)
    avg = re.search(r"average loss ([-+.e\d]+) to ([-+.e\d]+) over 8 epochs", text)
    assert avg and float(avg.group(2)) <= 1.02 * float(avg.group(1))  # measured: 12541 -> 12234


@pytest.mark.gpu
def test_arena_eval_smoke(kb):
    """kami::eval (kami/evaluate.h over the reference's algorithm, evaluate.cpp:10-160) on two small random
    networks: the stale-generation bail-out, then a real 3-game arena through select / infer x2 / expand."""
    exe = os.path.join(DROPIN, "eval_smoke")
    if not os.path.exists(exe):
        pytest.skip("kami/_dropin not built")
    out = subprocess.run([exe], capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-400:]
    text = out.stdout.decode()
    assert "model was updated during evaluation, skipping!" in text and "stale verdict 0" in text
    assert "evaluating model generation 1 over 3 games" in text and "arena verdict" in text
    assert "game 1 of 3" in text


@pytest.mark.gpu
def test_selfplay_train_arena_loop_smoke(kb):
    """The whole loop of kami.cpp on small settings (kami/tests/selfplay_smoke.cpp): device self-play fills the
    replay buffer, Selfplay::training_main (selfplay.cpp:215-304) clones, trains (CUDA step), runs the arena."""
    exe = os.path.join(DROPIN, "selfplay_smoke")
    if not os.path.exists(exe):
        pytest.skip("kami/_dropin not built")
    out = subprocess.run([exe], capture_output=True, timeout=400)
    text = out.stdout.decode()
    assert out.returncode == 0, (out.stderr.decode()[-400:], text[-400:])
    assert "TRAIN 0: training generation 0 with 48 trajectories" in text, text[-800:]
    assert "EVAL 0: evaluating model generation 1" in text
    assert ("candidate accepted" in text) or ("candidate rejected" in text)


@pytest.mark.gpu
def test_env_pgn_matches_reference_movetext(kb):
    """Env::pgn() (kami/env.h, SAN written on this library's rules) against the movetext the UNMODIFIED reference prints
    through the vendored thc library (tests/golden/pgn_games.json, made by tests/golden/make_pgn_golden.py): 16 seeded
    games, 2 716 plies with mates, checks, en-passant and ordinary captures, under-promotions, castling and file / rank
    disambiguation.  Games with a queen "promotion" are not in the fixture: the reference's own output is undefined
    there (the thc board promotes, the reference's game does not)."""
    import json

    exe = os.path.join(DROPIN, "pgn_check")
    if not os.path.exists(exe):
        pytest.skip("kami/_dropin not built")
    games = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pgn_games.json")))
    text = "".join("%d %s\n" % (len(g["actions"]), " ".join(map(str, g["actions"]))) for g in games)
    out = subprocess.run([exe], input=text.encode(), capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-400:]
    got = [ln[4:] for ln in out.stdout.decode().splitlines() if ln.startswith("PGN ")]
    assert len(got) == len(games)
    for g, mine in zip(games, got):
        assert mine == g["pgn"], (mine[-120:], g["pgn"][-120:])


@pytest.mark.gpu
def test_pool_hands_over_a_finished_game(kb):
    """Selfplay::get_next_pgn's device half (kb_pool_request_game / kb_pool_take_game): the action list of the next
    game that finishes replays legally on an Env to a terminal position."""
    import numpy as np
    import harness as H
    import nn_oracle as NO
    from kami_b200 import api

    net = kb.NN(64, 1)
    net.load_blob(NO.pack_blob(NO.init_params(64, 1, seed=3), 64, 1))
    pool = kb.TreePool(64, 1 << 14, api.tree_cfg(noise_weight=0.05, selfplay_nodes=4, seed=5, alpha_initial=1.0, alpha_final=1.0, **H.DEF_YML))
    assert pool.take_game() is None  # nothing requested
    pool.request_game()
    game = None
    for _ in range(400):
        pool.step(net, 16)
        game = pool.take_game()
        if game is not None:
            break
    assert game is not None and len(game) > 4
    env = kb.Env()
    for a in game:
        assert int(a) in [int(x) for x in env.actions()]
        env.push(int(a))
    assert env.terminal()[0]
    assert pool.take_game() is None  # handed over once
