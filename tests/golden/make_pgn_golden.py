"""Generates tests/golden/pgn_games.json from the UNMODIFIED reference (run in the dev container):

    make -C oracle ref && python tests/golden/make_pgn_golden.py

Seeded random legal games are played on the reference Env (oracle/_ref/libkami_ref_core.so) to a terminal position and
printed by the reference's own Env::pgn() -- SAN through the vendored thc library (oracle/_ref/libkami_ref_pgn.so).
Each record holds the action list and the reference movetext.  Games in which a pawn reaches the last rank without
promoting are left out: thc's private board promotes there while the reference's game does not, later moves
fail thc's Move::TerseIn and the reference prints (or crashes on) uninitialised moves.
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(n_games=160, keep=16):
    # only this library is loaded: two copies of the reference in one process share the static-init guard of
    # env.h:25-39 and the second copy's lookup tables stay uninitialised
    P = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libkami_ref_pgn.so"))
    games = []
    for g in range(n_games):
        arr = (C.c_int * 4096)()
        n = P.ref_random_game(C.c_ulonglong(1000 + g), 35, arr, 4096)
        assert n > 0
        buf = C.create_string_buffer(1 << 16)
        q3 = C.c_int()
        m = P.ref_game_pgn(arr, n, buf, len(buf), C.byref(q3))
        if m == -3:
            continue  # Q3 event: the reference's own output is undefined (see the shim)
        assert m > 0, m
        why = C.create_string_buffer(64)
        P.ref_game_reason(arr, n, why, 64)
        games.append({"actions": list(arr[:n]), "pgn": buf.value.decode(), "reason": why.value.decode()})
    # keep a spread: games that end in mate or stalemate, that contain under-promotions or castling first, then the
    # shortest of the rest
    games.sort(key=lambda r: ("mated" not in r["reason"], "=" not in r["pgn"], "O-O" not in r["pgn"], len(r["actions"])))
    out = games[:keep]
    if "--check" in sys.argv:  # tests/test_pgn_golden.py: the committed fixture is what the reference prints today
        have = json.load(open(os.path.join(HERE, "pgn_games.json")))
        assert have == out, "tests/golden/pgn_games.json differs from the reference's output"
        print("pgn fixture matches the compiled reference:", len(out), "games")
        return
    json.dump(out, open(os.path.join(HERE, "pgn_games.json"), "w"))
    print(len(games), "games without a Q3 event of", n_games)
    for r in out:
        print(len(r["actions"]), "plies,", r["reason"], "|", r["pgn"][:60], "...", r["pgn"][-60:])


if __name__ == "__main__":
    main()
