"""CPU: the C-ABI library loads and exports every symbol include/minimarl_b200.h declares, and the
ctypes binding covers exactly that set (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "minimarl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_are_exported_and_bound():
    import minimarl_b200._lib as L
    lib = L.lib()
    declared = _declared()
    assert declared, "no declarations parsed"
    exported = set(re.findall(r" T (mm_\w+)", subprocess.check_output(["nm", "-D", L.LIB_PATH], text=True)))
    assert declared <= exported, declared - exported
    assert exported <= declared, f"exported but not declared in the header: {exported - declared}"
    assert set(L.SIGNATURES) == declared
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mm_version() >= 1
    assert lib.mm_status_string(0) == b"ok"


def test_cfg_struct_layouts_match_header():
    import minimarl_b200._lib as L
    # int32 x4 + float[2][3] + float[3] + float[8]
    assert ctypes.sizeof(L.CheckersCfg) == 4 * 4 + 4 * (6 + 3 + 8)
    # int32 x5 + float x2 + float[3] + float[7]
    assert ctypes.sizeof(L.SwitchCfg) == 4 * 5 + 4 * (2 + 3 + 7)


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    import minimarl_b200._lib as L
    lib = L.lib()
    cfg = L.CheckersCfg()
    cfg.n_envs, cfg.max_steps = 0, 100
    assert lib.mm_checkers_reset(None, None, None, ctypes.byref(cfg), None) == -1
    assert lib.mm_gae_returns(None, None, None, None, None, 0, 0, 0.99, 0.95, 1, 0, 1.0, 0.0, None) == -1
    # the entry points added in round 2 validate before they launch as well (NULL pointers, unsupported shapes)
    assert lib.mm_clip_adam_step(None, None, None, None, None, None, None, 10, 10, 1e-3, None, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert lib.mm_qmixer_bwd(None, None, None, None, None, None, None, None, 1, 128, 94, 2, None) == -1
    assert lib.mm_qmixer_bwd_dims(95, (ctypes.c_int32 * 4)()) == -3            # odd state width: unsupported
    assert lib.mm_ln_head_fwd(None, None, None, None, None, None, None, None, 10, 64, 5, None) == -1
    assert lib.mm_gru_wgrad(None, None, None, None, None, None, 1, None, None, 0, 4, 8, 64, None) == -1
    assert lib.mm_per_sample_dev(None, None, 8, 4, None, 2, None, None, None, None) == -1
    assert lib.mm_replay_gather_stacked(*([None] * 13), 1, 1, 1, 2, 4, 8, 2, None) == -1
    assert lib.mm_gru_wgrad_ctas(128 * 1000) >= 1 and lib.mm_qmixer_bwd_ctas(1) == 1
    with pytest.raises(RuntimeError, match="bad argument"):
        L.check(-1)


def test_product_path_refuses_cpu_tensors():
    import torch
    import minimarl_b200._lib as L
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.ptr(torch.zeros(4))
    from minimarl_b200 import BatchedCheckers
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedCheckers(4, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "minimarl_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(d, f)
                assert "oracle/" not in src, os.path.join(d, f)
