"""Label / sampling helpers with the reference's signatures (tools.py:40-47, 65-70)."""
import torch

from . import ops


def asign_label(label, c_dim=None, mode='CelebA', normalize=True):
    """{0,1} attribute labels -> GMM component means {-1,+1} (tools.py:40-47)."""
    if mode not in ('CelebA', 'CUB200'):
        raise NotImplementedError("one-hot label modes are not used by configs/celeba_faces.yaml")
    out = label.clone()
    return out * 2.0 - 1.0 if normalize else out


def dist_sampling_split(mu, c_dim=8, stddev=0.5, device=None, eps=None):
    """z ~ N(mu_component, stddev) in the attribute-major layout of tools.py:65-70:
    z[b, j*c_dim + k] = mu[b, j] + stddev * eps[0, k, b, j].  `eps` defaults to a standard-normal draw of
    shape (1, c_dim, B, num_cls) from torch's generator (the stream Normal.sample consumes); tests pass it
    explicitly to make the draw identical to the oracle's."""
    if eps is None:
        eps = torch.randn(1, c_dim, mu.shape[0], mu.shape[1], device=mu.device, dtype=torch.float32)
    return ops.gmm_sample(mu, eps, float(stddev), c_dim)
