"""reacherdistilation_b200 -- B200-native hot path of winstonww/ReacherDistilation (batched Reacher-v2 rollout feeding the
teacher->student distillation loop).  Module names mirror the reference's `distilation` package."""
from . import config  # noqa: F401
from ._lib import (LOSS_KL_ST, LOSS_KL_TS, LOSS_MSE, LOSS_MSE_ACTION, MODE_FP32, MODE_TC, STUDENT_MLP, STUDENT_POLICY64, ReacherB200Error)  # noqa: F401

__all__ = ["config", "env", "teacher", "student_nn", "loss", "mlp_train", "lstm_train", "lstm2_train", "dataset", "main", "vf_train"]
