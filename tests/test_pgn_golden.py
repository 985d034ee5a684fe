"""CPU: the committed PGN fixture (tests/golden/pgn_games.json) is exactly what the UNMODIFIED reference's Env::pgn() prints
for the same seeded games (oracle/_ref/libkami_ref_pgn.so: reference Env + vendored thc).  Runs in a subprocess: two copies of
the reference in one process share the static-initialisation guard of env.h:25-39, so the PGN library cannot be loaded next to
libkami_ref_core.so.  Skipped where oracle/_ref was never built."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libkami_ref_pgn.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libkami_ref_pgn.so not built (make -C oracle ref)")
def test_pgn_fixture_is_the_reference_output():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_pgn_golden.py"), "--check"], capture_output=True, timeout=600)
    assert out.returncode == 0, (out.stdout.decode()[-300:], out.stderr.decode()[-600:])
    assert b"pgn fixture matches the compiled reference" in out.stdout
