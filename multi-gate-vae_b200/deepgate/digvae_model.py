"""Variational pieces (reference digvae_model.py:105-151, trainer.py:145-148).

Upstream's ``--model DG_VAE`` wiring is broken end to end, so the class
keeps the reference's constructor, parameter names (fc_{s,t}_{mu,logstd}) and ``sample`` contract
-- it stashes s_mu / s_logstd / t_mu / t_logstd on ``self`` -- and adds ``kl_loss`` with the
trainer's formula.  The elementwise reparameterisation and the KL reduction run in the fused
CUDA kernel (csrc/vae_func.cu); the four 64x64 Linear layers are plain library GEMMs.
"""
import torch

from . import ops
from .digae_layer import DirectedInnerProductDecoder

EPS = 1e-15
MAX_LOGSTD = 10


class DirectedGVAE(torch.nn.Module):
    def __init__(self, encoder, dim_hidden, decoder=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = DirectedInnerProductDecoder() if decoder is None else decoder
        self.dim_hidden = dim_hidden
        self.fc_s_mu = torch.nn.Linear(dim_hidden, dim_hidden)
        self.fc_s_logstd = torch.nn.Linear(dim_hidden, dim_hidden)
        self.fc_t_mu = torch.nn.Linear(dim_hidden, dim_hidden)
        self.fc_t_logstd = torch.nn.Linear(dim_hidden, dim_hidden)
        self.kl = None

    def sample(self, s, t, eps_s=None, eps_t=None, hf=None, tt_pair_index=None, tt_sim=None):
        """z = mu + exp(logstd) * eps for s and t.  Noise may be injected (parity tests); when the
        func-loss inputs are given the same launch also produces it (``self.func_loss``)."""
        self.s_mu, self.s_logstd = self.fc_s_mu(s), self.fc_s_logstd(s)
        self.t_mu, self.t_logstd = self.fc_t_mu(t), self.fc_t_logstd(t)
        if eps_s is None:
            eps_s = torch.randn_like(self.s_mu)
        if eps_t is None:
            eps_t = torch.randn_like(self.t_mu)
        mu = torch.stack([self.s_mu, self.t_mu])
        ls = torch.stack([self.s_logstd, self.t_logstd])
        eps = torch.stack([eps_s, eps_t])
        z_s, z_t, self.kl, self.func_loss = ops.vae_func_loss(mu, ls, eps, hf, tt_pair_index, tt_sim)
        return z_s, z_t

    def kl_loss(self):
        """trainer.py:145-148: sum over {s,t} of -0.5/N * mean_i sum_d (1 + 2 logstd - mu^2 - exp(logstd)^2);
        produced by the last ``sample`` call."""
        if self.kl is None:
            raise RuntimeError("call sample() first")
        return self.kl

    def encode(self, *args, **kwargs):
        return self.encoder(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.decoder(*args, **kwargs)
