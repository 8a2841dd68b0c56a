"""ctypes binding of libreacher_b200.so -- the only way the Python host layer reaches the GPU.

There is no CPU fallback: if the library is missing it is (re)built with nvcc; if that fails, or no CUDA device is
visible when a compute entry point is called, the call raises.
"""
import ctypes as C
import os
import re

from . import build as _build

HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "reacher_b200.h")


class ReacherB200Error(RuntimeError):
    pass


_lib = None

_vp, _fp, _u8p, _i32p, _u32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (device or host)
_SIGS = {
    "rb_last_error": (C.c_char_p, []),
    "rb_version": (C.c_int, []),
    "rb_device_count": (C.c_int, []),
    "rb_sm_count": (C.c_int, [C.c_int]),
    "rb_mode_available": (C.c_int, [C.c_int]),
    "rb_student_mode_available": (C.c_int, [C.c_int]),
    "rb_env_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_uint64, C.c_int, C.c_uint32]),
    "rb_env_destroy": (C.c_int, [_vp]),
    "rb_env_num_envs": (C.c_int64, [_vp]),
    "rb_env_reset": (C.c_int, [_vp, _fp, _vp]),
    "rb_env_step": (C.c_int, [_vp, _fp, _fp, _fp, _u8p, _vp]),
    "rb_env_reset_host": (C.c_int, [_vp, _fp]),
    "rb_env_step_host": (C.c_int, [_vp, _fp, _fp, _fp, _u8p]),
    "rb_env_serve_policy": (C.c_int, [_vp, _fp, C.c_int]),
    "rb_env_serve_policy_fwd": (C.c_int, [_vp, _fp, _fp]),
    "rb_env_act_step_host": (C.c_int, [_vp, _fp, _fp, _fp, _u8p, _fp]),
    "rb_env_get_state": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _i32p, _u32p, _fp, _vp]),
    "rb_env_set_state": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _i32p, _u32p, _fp, _vp]),
    "rb_env_observe": (C.c_int, [_vp, _fp, _vp]),
    "rb_env_rollout_random": (C.c_int, [_vp, C.c_int, C.c_uint32, _fp, _fp, _fp, _u8p, _vp]),
    "rb_policy_param_count": (C.c_int64, [C.c_int]),
    "rb_policy_fwd": (C.c_int, [_fp, C.c_int, _fp, C.c_int64, _fp, C.c_int, _vp]),
    "rb_policy_fwd_host": (C.c_int, [_fp, C.c_int, _fp, C.c_int64, _fp, C.c_int, C.c_int]),
    "rb_env_rollout_policy": (C.c_int, [_vp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _u8p, C.c_int, _vp]),
    "rb_env_rollout_policy_host": (C.c_int, [_vp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _u8p, C.c_int]),
    "rb_env_rollout_policy_host_ex": (C.c_int, [_vp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _u8p, _vp, _fp, C.c_int]),
    "rb_env_rollout_policy_host_begin": (C.c_int, [_vp, _fp, C.c_int, C.c_int, _fp, _vp, _fp, C.c_int]),
    "rb_env_rollout_policy_host_wait": (C.c_int, [_vp]),
    "rb_env_set_host_transport": (C.c_int, [_vp, C.c_int]),
    "rb_env_rollout_buffer": (C.c_int, [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                              C.POINTER(C.c_int)]),
    "rb_student_param_count": (C.c_int64, [C.c_int]),
    "rb_student_input_dim": (C.c_int, [C.c_int]),
    "rb_student_workspace_bytes": (C.c_int64, [C.c_int, C.c_int64, C.c_int]),
    "rb_student_mlp_input": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_uint32, _fp, _vp]),
    "rb_student_fwd": (C.c_int, [C.c_int, _fp, _fp, C.c_int64, _fp, C.c_int, _vp]),
    "rb_student_fwd_ws": (C.c_int, [C.c_int, _fp, _fp, C.c_int64, _fp, _vp, C.c_int, _vp]),
    "rb_student_loss_grad": (C.c_int, [C.c_int, _fp, _fp, _fp, C.c_int64, C.c_int, _fp, _fp, _vp, C.c_int, _vp]),
    "rb_student_step": (C.c_int, [C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, _fp, _fp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_int, _vp]),
    "rb_student_step_dp": (C.c_int, [C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, _fp, _fp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float, C.c_int, C.c_int, _vp, _vp, C.c_uint32, _vp]),
    "rb_debug_student_timers": (C.c_int, [_vp]),
    "rb_adam_step": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    "rb_dagger_create": (C.c_int, [C.POINTER(C.c_void_p), _vp, C.c_int, C.c_float]),
    "rb_dagger_destroy": (C.c_int, [_vp]),
    "rb_dagger_observe": (C.c_int, [_vp, _fp, C.c_uint32, _fp, _fp, _fp, _fp, C.c_int, _vp]),
    "rb_dagger_invalidate_teacher": (C.c_int, [_vp]),
    "rb_dagger_get_state": (C.c_int, [_vp, _fp, _fp, _fp, _vp]),
    "rb_dagger_set_state": (C.c_int, [_vp, _fp, _fp, _fp, _vp]),
    "rb_dagger_set_clock": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "rb_dagger_step": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _fp, _fp, _fp, _fp, _u8p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "rb_dagger_act": (C.c_int, [_vp, _fp, _fp, _fp, _u8p, _vp]),
    "rb_dagger_wait_loss": (C.c_int, [_vp, C.c_uint32, C.POINTER(C.c_float)]),
    "rb_lstm_param_count": (C.c_int64, []),
    "rb_lstm_steps": (C.c_int, []),
    "rb_lstm_units": (C.c_int, []),
    "rb_lstm_workspace_bytes": (C.c_int64, [C.c_int64]),
    "rb_lstm_fwd": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int64, _fp, _fp, _vp, _vp]),
    "rb_lstm_loss_grad": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _fp, _fp, _fp, _vp, _vp]),
    "rb_lstm_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "rb_lstm_ctx_destroy": (C.c_int, [_vp]),
    "rb_lstm_ctx_set_clock": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp]),
    "rb_lstm_step": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_int, _fp, _fp, _vp, C.c_float,
                               C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _vp]),
    "rb_lstm2_param_count": (C.c_int64, [C.POINTER(C.c_int)]),
    "rb_lstm2_workspace_bytes": (C.c_int64, [C.POINTER(C.c_int), C.c_int64]),
    "rb_lstm2_fwd": (C.c_int, [C.POINTER(C.c_int), _fp, _fp, _fp, _fp, C.c_int64, _fp, _fp, _fp, _vp, _vp]),
    "rb_lstm2_loss_grad": (C.c_int, [C.POINTER(C.c_int), _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                     _fp, _fp, _fp, _fp, _vp, _vp]),
    "rb_lstm2_step": (C.c_int, [_vp, C.POINTER(C.c_int), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_int, _fp, _fp,
                                _fp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _vp]),
    "rb_gemm_bf16x3": (C.c_int, [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _fp,
                                 C.c_int, _fp, C.c_int64, _vp]),
    "rb_gemm_set_cta_packing": (C.c_int, [C.c_int]),
    "rb_dense_param_count": (C.c_int64, [C.c_int, C.POINTER(C.c_int)]),
    "rb_dense_workspace_bytes": (C.c_int64, [C.c_int, C.POINTER(C.c_int), C.c_int64]),
    "rb_dense_fwd": (C.c_int, [_fp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _fp, C.c_int64, _fp, _vp, _vp]),
    "rb_dense_loss_grad": (C.c_int, [_fp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _fp, _fp, C.c_int64, C.c_int, _fp, _fp, _vp, _vp]),
    "rb_vf_targets": (C.c_int, [_fp, C.c_int64, C.c_int, C.c_float, _fp, _vp]),
    "rb_dataset_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int64, C.c_int]),
    "rb_dataset_destroy": (C.c_int, [_vp]),
    "rb_dataset_write": (C.c_int, [_vp, _fp, _fp, _fp, _fp, C.c_int, _vp]),
    "rb_dataset_flush": (C.c_int, [_vp]),
    "rb_dataset_num_episodes": (C.c_int64, [_vp]),
    "rb_dataset_num_available": (C.c_int64, [_vp]),
    "rb_dataset_episode_len": (C.c_int, [_vp]),
    "rb_dataset_generations": (C.c_int64, [_vp]),
    "rb_dataset_export_host": (C.c_int, [_vp, C.c_int64, _fp, _fp, _fp, _fp, _u8p]),
    "rb_dataset_ring_rows": (C.c_int64, [_vp]),
    "rb_dataset_save_host": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _u8p, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "rb_dataset_load_host": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _u8p, C.c_int, C.c_int64]),
    "rb_dataset_training_batch": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _i32p, _i32p, _vp]),
    "rb_dataset_test_batch": (C.c_int, [_vp, _fp, C.c_int, _fp, _fp, _fp, _vp]),
}

MODE_FP32, MODE_TC = 0, 1
STUDENT_POLICY64, STUDENT_MLP = 0, 1
LOSS_KL_ST, LOSS_KL_TS, LOSS_MSE, LOSS_MSE_ACTION = 0, 1, 2, 3


def header_symbols():
    """Every function name declared in include/reacher_b200.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path):
            path = _build.build()
        L = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise ReacherB200Error("libreacher_b200 error %d: %s" % (rc, lib().rb_last_error().decode()))


def ptr(t):
    """Raw address of a torch tensor (device or host) or numpy array; None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous(), "tensor must be contiguous"
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
