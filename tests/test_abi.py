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
