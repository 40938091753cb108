"""TEST-ONLY stand-in for ITPAL `cpp_projection` (C++/Armadillo/NLopt, not installable offline,
unpinned upstream).  NOT ITPAL: it forwards to the oracle's fp64 restated 1-D dual solve so that the
reference's unmodified kl_projection_layer.py can be imported and run for golden generation.
Only the op the shipped configs reach (policy_type gnn_diag + proj_type kl) is provided."""
import numpy as np
import torch

from oracle.projection import kl_diag_cov_solve


class BatchedDiagCovOnlyProjection:
    def __init__(self, batch_shape, dim, max_eval=1000):
        self.batch_shape, self.dim, self.max_eval = batch_shape, dim, max_eval
        self._saved = None

    def forward(self, eps, old_cov, cov):
        eps = np.asarray(eps, dtype=np.float64).reshape(-1)
        assert np.all(eps == eps[0])
        c = torch.as_tensor(np.asarray(cov, dtype=np.float64))
        o = torch.as_tensor(np.asarray(old_cov, dtype=np.float64))
        c_t, eta = kl_diag_cov_solve(c, o, float(eps[0]))
        self._saved = (c, o, eta)
        return c_t.numpy()

    def backward(self, d_cov):
        c, o, eta = self._saved
        g = torch.as_tensor(np.asarray(d_cov, dtype=np.float64))
        e = eta[..., None]
        D = e / o + 1.0 / c
        c_t = (e + 1.0) / D
        dct_dc = (e + 1.0) / (D * D * c * c)
        dct_deta = (1.0 / c - 1.0 / o) / (D * D)
        dkl_dct = 0.5 * (1.0 / o - 1.0 / c_t)
        num = (g * dct_deta).sum(-1, keepdim=True)
        den = (dkl_dct * dct_deta).sum(-1, keepdim=True)
        active = (eta > 0)[..., None]
        den = torch.where(active, den, torch.ones_like(den))
        grad = torch.where(active, g * dct_dc - num * (dkl_dct * dct_dc) / den, g)
        return grad.numpy()
