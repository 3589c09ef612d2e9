"""The discriminator calls of the reference's loss module, packed.

`model/losses/loss_module.py` evaluates its discriminator -- a `TiTokEncoder(out_channels=1)` with 4 register tokens
(loss_module.py:41-48, 96-101) -- six times per training step: twice in the generator step (real, fake; :140-153) and four
times in the discriminator step (real, fake, real + noise, fake + noise; :166-214), each time as its own launch
sequence over B clips. Clips never interact inside the encoder (block-diagonal attention, per-row norms), so the
forwards of one step are ONE packed batch here: one plan, one launch sequence (~25 launches instead of ~100), one
backward. The loss arithmetic is the reference's, expression for expression.

Drop-in use inside a `ReconstructionLoss`-like module:

    disc = PackedDiscriminator(disc_model, disc_tokens=4)
    g_loss = disc.generator_loss(target, recon)                       # [B], loss_module.py:140-153
    total, logs = disc.discriminator_loss(target, recon, gp_weight=0.1, gp_noise=0.1, centering_weight=0.01)   # :166-214
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


class PackedDiscriminator:
    def __init__(self, disc_model: torch.nn.Module, disc_tokens: int = 4):
        self.disc_model = disc_model
        self.disc_tokens = int(disc_tokens)

    def logits(self, groups: Sequence[Sequence[torch.Tensor]]) -> List[torch.Tensor]:
        """One packed encoder call for several lists of clips; returns one [B_i] logit vector per list
        (`disc_wrapper`, loss_module.py:96-101: mean over the register tokens)."""
        clips = [c for g in groups for c in g]
        out = self.disc_model(clips, [self.disc_tokens] * len(clips))  # [sum(B) * disc_tokens, 1]
        per_clip = out.view(len(clips), -1).mean(-1)
        return list(torch.split(per_clip, [len(g) for g in groups]))

    def set_requires_grad(self, flag: bool) -> None:
        for p in self.disc_model.parameters():
            p.requires_grad = flag

    # ---- generator step (loss_module.py:140-153) ---------------------------------------------------------------
    def generator_loss(self, target: Sequence[torch.Tensor], recon: Sequence[torch.Tensor]) -> torch.Tensor:
        """softplus(-(D(fake) - D(real))) per clip; discriminator parameters frozen, gradient flows to `recon`."""
        self.set_requires_grad(False)
        logits_real, logits_fake = self.logits([[t.detach().contiguous() for t in target], [r.contiguous() for r in recon]])
        return F.softplus(-(logits_fake - logits_real))

    # ---- discriminator step (loss_module.py:166-214) -----------------------------------------------------------
    def discriminator_loss(self, target: Sequence[torch.Tensor], recon: Sequence[torch.Tensor], gp_weight: float,
                           gp_noise: float, centering_weight: float,
                           noise: Optional[Sequence[torch.Tensor]] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """Relativistic loss + noise-based R1 / R2 penalties (finite differences of the logits: no second-order path) +
        centering. `noise` (one tensor per clip, already scaled by gp_noise) may be passed for reproducibility; by default
        it is drawn like the reference does (`randn_like(x) * gp_noise`, the same noise for the real and the fake clip).

        The reference marks the detached clips `requires_grad_(True)` (:169-170) although nothing reads their gradients;
        here they stay plain inputs, which skips the pixel-gradient kernels and leaves the loss and every parameter
        gradient unchanged."""
        self.set_requires_grad(True)
        target = [t.detach().contiguous() for t in target]
        recon = [r.detach().contiguous() for r in recon]
        groups = [target, recon]
        if gp_weight > 0.0:
            if noise is None:
                noise = [torch.randn_like(x) * gp_noise for x in target]
            groups += [[x + n for x, n in zip(target, noise)], [x + n for x, n in zip(recon, noise)]]
        lg = self.logits(groups)
        logits_real, logits_fake = lg[0], lg[1]
        loss_dict: Dict[str, torch.Tensor] = {}
        logits_relative = logits_real - logits_fake
        d_loss = F.softplus(-logits_relative)
        loss_dict["d_loss"] = d_loss
        loss_dict["logits_relative"] = logits_relative
        gradient_penalty = 0.0
        if gp_weight > 0.0:
            r1_penalty = (logits_real - lg[2]) ** 2
            r2_penalty = (logits_fake - lg[3]) ** 2
            loss_dict["r1_penalty"] = r1_penalty
            loss_dict["r2_penalty"] = r2_penalty
            gradient_penalty = r1_penalty + r2_penalty
        centering_loss = 0.0
        if centering_weight > 0.0:
            centering_loss = ((logits_real + logits_fake) ** 2) / 2
            loss_dict["centering_loss"] = centering_loss
        total_loss = (d_loss + ((gp_weight / gp_noise ** 2 * gradient_penalty) if gp_weight > 0.0 else 0.0)
                      + (centering_weight * centering_loss)).mean()
        loss_dict["total_loss"] = total_loss
        return total_loss, {"disc/" + k: v.clone().mean().detach() for k, v in loss_dict.items()}
