"""GMM style-space losses with the reference's call signatures (gmm.py:13-22, 33-41)."""
import torch

from . import ops


def _cat(x):
    return torch.cat(list(x), dim=1) if isinstance(x, (list, tuple)) else x


def gmm_kl_distance_sp(pred_mus, pred_sigma, mus, sigma):
    """sum_i mean_b sum_k 0.5*(log(sigma/exp(lv_i)) + (exp(lv_i) + (mu_i - c[:, i])^2)/sigma - 1)   (gmm.py:13-22).
    pred_mus / pred_sigma: lists of num_cls [B, c_dim] tensors or already concatenated [B, num_cls*c_dim]."""
    mu, lv = _cat(pred_mus), _cat(pred_sigma)
    ncls = mus.shape[1]
    return ops.gmm_kl(mu, lv, mus, float(sigma), ncls, mu.shape[1] // ncls)


def gmm_earth_mover_distance_sp(pred_mus, mus):
    """sum_i mean_b sum_k |mu_i - c[:, i]|  (gmm.py:33-41), the `dist_mode: em` alternative."""
    mu = _cat(pred_mus)
    ncls = mus.shape[1]
    cdim = mu.shape[1] // ncls
    target = mus.float().repeat_interleave(cdim, dim=1).contiguous()
    return ops.l1_loss(mu.contiguous().float(), target) * float(mu.shape[1])
