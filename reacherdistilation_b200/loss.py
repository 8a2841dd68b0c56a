"""kl_loss (/root/reference src/distilation/loss.py:3-13): KL(student || teacher) of diagonal Gaussians given as pdflat,
SUMMED over every leading dimension.  On the hot path the loss is fused into the student backward kernel
(StudentNet.loss_grad); this standalone form evaluates loss only, through the same kernel's forward half."""
import torch

from ._lib import LOSS_KL_ST, LOSS_KL_TS  # noqa: F401


def kl_loss(s_pdflat_batch, t_pdflat_batch, pdtype=None, reverse=False):
    """Device tensors [..., 4] -> scalar tensor.  Plain elementwise torch on already-computed pdflats (not on the training
    path -- training uses the fused CUDA loss+grad); kept for API parity with the reference's loss module."""
    s, t = s_pdflat_batch.reshape(-1, 4).float(), t_pdflat_batch.reshape(-1, 4).float()
    if reverse:
        s, t = t, s
    ms, ls, mt, lt = s[:, :2], s[:, 2:], t[:, :2], t[:, 2:]
    return (lt - ls + (torch.exp(2 * ls) + (ms - mt) ** 2) / (2 * torch.exp(2 * lt)) - 0.5).sum()
