"""TEST-ONLY dense torch checker (device agnostic) used by the GPU tests for sizes where the numpy
oracle is too slow: the same closed forms as oracle/contrastive_oracle.py written with torch ops."""
import torch


def row_stats(x, y, s_eff):
    """x[M,D], y[N,D] (any float dtype) -> lse[M] (nats), mu[M], var[M] in float64."""
    z = x.double() @ y.double().t()
    l = s_eff * z
    lse = torch.logsumexp(l, dim=1)
    p = torch.exp(l - lse[:, None])
    mu = (p * z).sum(1)
    var = (p * z * z).sum(1) - mu * mu
    return z, lse, mu, var


def clip_loss_and_grads(img, txt, s):
    """Single-rank symmetric InfoNCE via autograd in float64 (reference loss.py:132-155)."""
    i = img.double().clone().requires_grad_(True)
    t = txt.double().clone().requires_grad_(True)
    sc = torch.tensor(float(s), dtype=torch.float64, device=img.device, requires_grad=True)
    l = sc * i @ t.t()
    lab = torch.arange(i.shape[0], device=img.device)
    loss = 0.5 * (torch.nn.functional.cross_entropy(l, lab) + torch.nn.functional.cross_entropy(l.t(), lab))
    loss.backward()
    return loss.detach(), i.grad, t.grad, sc.grad
