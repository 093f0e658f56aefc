"""Drop-in replacement for posetimation/loss/mse_loss.py and the loss combination of
engine/core/functions/alignment_mi_function_term6_1.py:119-148."""
import torch
import torch.nn as nn

from . import ops


class JointMSELoss(nn.Module):
    """posetimation/loss/mse_loss.py:13-40: sum_j MSE_mean(pred_j*w_j, gt_j*w_j) [/ num_joints],
    as one fused reduction."""

    def __init__(self, use_target_weight: bool = True, divided_num_joints=True):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.divided_num_joints = divided_num_joints

    def forward(self, output, target, target_weight):
        tw = target_weight if self.use_target_weight else None
        if torch.is_grad_enabled() and output.requires_grad:
            from .autograd import JointMSEFunction
            loss = JointMSEFunction.apply(output, target, tw)
        else:
            loss = ops.joint_mse(output, target, tw)
        if not self.divided_num_joints:
            loss = loss * output.shape[1]
        return loss


def combine_losses(mse, mi, w_mse=1.0, alpha=0.5, beta=0.1):
    """alignment_mi_function_term6_1.py:119-148 on device scalars (no .item() syncs):
    loss = MSE*w + alpha*( -beta*mi1 + beta*mi2 + mi3 - mi4 + mi5 - mi6 )."""
    return mse * w_mse + alpha * (-beta * mi[0] + beta * mi[1] + mi[2] - mi[3] + mi[4] - mi[5])
