"""Squeeze-and-Excitation block -- drop-in for the reference's ``models.se_blocks.SEBlock``
(reference: src/models/se_blocks.py:4-19; same constructor, parameter names and shapes).

Inside ``BetaVAE`` the SE arithmetic never runs through this module's ``forward``: it is
fused with GroupNorm + ReLU in ``bvae_gn_relu_se_{fwd,bwd}`` (see models/beta_vae.py); the
sub-modules here only own the parameters so ``state_dict`` keys (``fc.0.weight`` ...) and
``model.apply(weight_init)`` behave exactly like the reference's."""
import torch
import torch.nn as nn


class SEBlock(nn.Module):
    def __init__(self, channels, reduction=16):
        super().__init__()
        hidden = max(1, channels // reduction)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channels, hidden),
            nn.ReLU(inplace=True),
            nn.Linear(hidden, channels),
            nn.Sigmoid(),
        )

    @property
    def gate_params(self):
        """(w1 [R,C], b1 [R], w2 [C,R], b2 [C]) as the fused kernel consumes them."""
        return self.fc[0].weight, self.fc[0].bias, self.fc[2].weight, self.fc[2].bias

    def forward(self, x):
        """The SE arithmetic exists only fused with GroupNorm + ReLU (``bvae_gn_relu_se_fwd``); there is no ATen or
        CPU composition behind this module, so a stand-alone call fails loudly instead of silently running one."""
        raise NotImplementedError(
            "models.se_blocks.SEBlock (bvae_b200) owns the SE parameters; its arithmetic runs inside the fused "
            "GroupNorm+ReLU+SE kernel of BetaVAE's blocks and is not available as a stand-alone module call")
