"""TEST INFRASTRUCTURE ONLY -- torch restatement (any device) of the reference's image-space regularisers, checked
on the CPU against golden vectors produced by the reference's own functions (tests/golden/make_golden_regularizers.py).

  normal2curv               matcha/dm_utils/rendering.py:392-406
  compute_depth_order_loss  matcha/dm_regularization/depth.py:142-222 (the pixel shifts are an argument here, so a
                            test can hand both sides the same draw)
"""
import torch


def normal2curv(normal, mask):
    # rendering.py:392-406
    n = normal.permute([1, 2, 0])
    m = mask.permute([1, 2, 0])
    n = torch.nn.functional.pad(n[None], [0, 0, 1, 1, 1, 1], mode='replicate')
    m = torch.nn.functional.pad(m[None].to(torch.float32), [0, 0, 1, 1, 1, 1], mode='replicate').to(torch.bool)
    n_c = n[:, 1:-1, 1:-1, :] * m[:, 1:-1, 1:-1, :]
    n_u = (n[:, :-2, 1:-1, :] - n_c) * m[:, :-2, 1:-1, :]
    n_l = (n[:, 1:-1, :-2, :] - n_c) * m[:, 1:-1, :-2, :]
    n_b = (n[:, 2:, 1:-1, :] - n_c) * m[:, 2:, 1:-1, :]
    n_r = (n[:, 1:-1, 2:, :] - n_c) * m[:, 1:-1, 2:, :]
    curv = (n_u + n_l + n_b + n_r)[0]
    curv = curv.permute([2, 0, 1]) * mask
    return curv.norm(1, 0, True)


def depth_order_loss(depth, prior_depth, pixel_shifts, scene_extent=1., normalize_loss=True, log_space=False, log_scale=20.,
                     reduction="mean"):
    # depth.py:168-213 with the shifts handed in
    height, width = depth.squeeze().shape
    pixel_coords = torch.stack(torch.meshgrid(torch.arange(height, device=depth.device), torch.arange(width, device=depth.device),
                                              indexing='ij'), dim=-1).view(-1, 2)
    shifted = (pixel_coords + pixel_shifts).clamp(min=torch.tensor([0, 0], device=depth.device),
                                                  max=torch.tensor([height - 1, width - 1], device=depth.device))
    shifted_depth = depth.squeeze()[shifted[:, 0], shifted[:, 1]].reshape(depth.shape)
    shifted_prior = prior_depth.squeeze()[shifted[:, 0], shifted[:, 1]].reshape(depth.shape)
    diff = (depth - shifted_depth) / scene_extent
    prior_diff = (prior_depth - shifted_prior) / scene_extent
    if normalize_loss:
        prior_diff = prior_diff / prior_diff.detach().abs().clamp(min=1e-8)
    loss = -(diff * prior_diff).clamp(max=0)
    if log_space:
        loss = torch.log(1. + log_scale * loss)
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss
