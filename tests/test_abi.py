"""CPU: the C-ABI library builds, loads, and exports every symbol include/kami_b200.h declares;
compute entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

import harness as H


def header_symbols():
    src = open(os.path.join(H.ROOT, "include", "kami_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import kami_b200
    from kami_b200 import api

    so = kami_b200.lib_path()
    assert os.path.exists(so), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    declared = header_symbols()
    assert len(declared) >= 60
    exported = subprocess.check_output(["nm", "-D", "--defined-only", so]).decode()
    exported = set(re.findall(r" T (kb_[a-z0-9_]+)", exported))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    assert sorted(api.ABI_SYMBOLS) == declared  # the ctypes table covers the header exactly
    L = kami_b200.lib()
    for s in declared:
        assert hasattr(L, s)


def test_struct_layouts_match_header():
    from kami_b200 import api

    assert C.sizeof(api.Position) == 80 and api.POSITION_DTYPE.itemsize == 80
    assert C.sizeof(api.TreeCfg) == 72
    # the oracle's wire export and the product's position share one layout
    e = H.OracleEnv()
    p = api.as_positions(e.export()[None, :])
    assert p["castle"][0] == 0xF and p["ep"][0] == 0xFF and p["ctm"][0] == 0 and p["key"][0] == e.key()
    assert int(p["white"][0]) == 0xFFFF and int(p["pieces"][0][0]) == 0x00FF00000000FF00


def test_blob_size_matches_oracle_param_order():
    import kami_b200
    import nn_oracle as NO

    L = kami_b200.lib()
    for F, R in ((64, 2), (256, 2), (256, 20), (128, 3)):
        n = sum(int(np.prod(s)) for _, s in NO.param_order(F, R))
        assert L.kb_net_blob_floats(F, R) == n


def test_no_cpu_fallback():
    import kami_b200

    if kami_b200.device_count() > 0:
        return  # on the GPU box the gpu-marked tests cover the compute path
    L = kami_b200.lib()
    assert L.kb_init(0) != 0
    assert b"no CPU path" in L.kb_last_error()
    h = C.c_void_p()
    assert L.kb_env_create(C.byref(h)) != 0  # compute entry points refuse to run without a device


def test_sass_has_tcgen05_and_tma():
    import kami_b200

    sass = subprocess.run(["cuobjdump", "-sass", kami_b200.lib_path()], capture_output=True, text=True)
    if sass.returncode != 0:
        return
    for mnemonic in ("UTCHMMA", "UBLKCP", "LDTM"):
        assert mnemonic in sass.stdout, mnemonic
