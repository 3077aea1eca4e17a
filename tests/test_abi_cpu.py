"""CPU suite: the C-ABI library loads, exports every symbol include/mm2b200.h declares, and refuses to compute without a GPU."""
import os
import re

import pytest


def test_library_exports_every_declared_symbol(mm2):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "mm2b200.h")).read()
    declared = set(re.findall(r"\b(mm2_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = mm2.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(mm2.ABI_SYMBOLS)
    diag = open(os.path.join(root, "include", "mm2b200_diag.h")).read()
    ddecl = set(re.findall(r"\b(mm2_[a-z0-9_]+)\s*\(", diag))
    assert ddecl == set(mm2.DIAG_SYMBOLS) and not [s for s in ddecl if not hasattr(L, s)]
    assert not (ddecl & declared)   # diagnostics stay out of the drop-in header


def test_no_cpu_fallback(mm2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mm2.Mm2Error) as e:
        mm2.Context(0)
    assert e.value.code == mm2.MM2_E_CUDA


def test_product_does_not_touch_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "minimap2_rs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("checked against the oracle", ""), os.path.join(dp, f)


def test_default_params_match_reference_values(mm2):
    p = mm2.default_chain_params(15)   # main.rs:105-123
    assert (p.max_dist_x, p.max_dist_y, p.bw, p.max_chain_iter, p.min_chain_score, p.min_cnt) == (5000, 5000, 500, 5000, 40, 3)
    assert (p.max_chain_skip, p.max_drop, p.bw_long, p.rmq_rescue_size) == (25, 500, 20000, 1000)
    import numpy as np
    assert np.float32(p.chn_pen_gap) == np.float32(np.float32(0.01) * np.float32(0.8)) * np.float32(15)
    assert mm2.apply_preset("map-hifi", 10, 15) == (10, 19) and mm2.apply_preset("sr", 10, 15) == (11, 21)
    o = mm2.default_map_opts()
    assert (o.w, o.k, o.max_gap, o.min_cnt, o.min_chain_score, o.best_n) == (10, 15, 5000, 3, 40, 5)


def test_pack_reads_host_helper(mm2):
    """mm2_pack_reads (host only): nt4.rs:2-10 codes, 4 per byte, + the positions of everything else"""
    import numpy as np
    rng = np.random.default_rng(3)
    cat = rng.choice(np.frombuffer(b"ACGTacgtNnRY", dtype=np.uint8), 10_003).astype(np.uint8)
    packed, n_pos = mm2.pack_reads(cat)
    nt4 = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        nt4[ch] = i
        nt4[ch + 32] = i
    codes = nt4[cat]
    assert (n_pos == np.nonzero(codes == 4)[0]).all()
    got = (packed[np.arange(cat.size) // 4] >> (2 * (np.arange(cat.size) % 4)).astype(np.uint8)) & 3
    assert (got[codes < 4] == codes[codes < 4]).all() and (got[codes == 4] == 0).all()
    p2, n2 = mm2.pack_reads(np.frombuffer(b"N" * 5000, dtype=np.uint8))     # more positions than the first capacity guess
    assert n2.size == 5000
