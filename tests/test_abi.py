"""The C-ABI library builds, loads and exports every symbol include/reacher_b200.h declares; without a GPU the compute
entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os

import pytest
import torch

from reacherdistilation_b200 import _lib, build


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    L = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    assert set(syms) == set(_lib._SIGS), "ctypes table out of sync with the header"
    assert L.rb_version() >= 100


def test_static_queries_need_no_gpu():
    L = _lib.lib()
    assert L.rb_policy_param_count(2) == 22 + 11 * 64 + 64 + 64 * 64 + 64 + 64 * 2 + 2 + 2
    assert L.rb_policy_param_count(4) == 5212 and L.rb_policy_param_count(3) == -1
    assert L.rb_student_param_count(_lib.STUDENT_MLP) == 24380
    assert L.rb_student_param_count(_lib.STUDENT_POLICY64) == 5212
    assert L.rb_student_input_dim(_lib.STUDENT_MLP) == 16 and L.rb_student_input_dim(_lib.STUDENT_POLICY64) == 11


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure():
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.rb_env_create(C.byref(h), 16, 0, 0, 0)
    assert rc == -2 and b"CUDA" in L.rb_last_error()
    from reacherdistilation_b200.env import VecReacher
    with pytest.raises(_lib.ReacherB200Error):
        VecReacher(num_envs=4)


def test_bad_arguments_are_rejected():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.rb_env_create(C.byref(h), 0, 0, 0, 0) == -1
    assert L.rb_env_create(None, 4, 0, 0, 0) == -1
    assert L.rb_env_create(C.byref(h), 16, 0, 0, 2 ** 32 - 4) == -1      # global env id would overflow
    assert b"num_envs" in L.rb_last_error() or b"overflow" in L.rb_last_error()


def test_host_side_param_count_matches_the_library_and_teacher_loader(tmp_path):
    """teacher.policy_param_count restates rb_policy_param_count so that building / loading a teacher vector maps no CUDA library (the
    reference arm of bench.py relies on that); load_teacher_params reads a flat vector or the baselines variable names."""
    import numpy as np
    from reacherdistilation_b200 import teacher as T
    L = _lib.lib()
    for nout in (2, 4):
        assert T.policy_param_count(nout) == L.rb_policy_param_count(nout)
    p = T.init_policy_params(seed=5)
    np.savez(tmp_path / "flat.npz", params=p)
    q, desc = T.load_teacher_params(teacher_ckpt=str(tmp_path / "flat.npz"))
    assert np.array_equal(p, q) and "flat" in desc
    assert T.load_teacher_params() == (None, None)
    with pytest.raises(ValueError):
        T.load_teacher_params(teacher_params=p[:-1])
    with pytest.raises(_lib.ReacherB200Error):
        T.TeacherAgent(restore=True)                       # the reference always restores teacher.ckpt: never silently randomised
