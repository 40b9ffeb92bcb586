"""Model description shared by the host-side mirror and the C-ABI.

Mirrors the constructor arguments of the reference ``DiffusionGPT``
(beso/agents/diffusion_agents/k_diffusion/score_gpts.py:121-139) and the
``sigma_data`` of ``GCDenoiser`` (score_wrappers.py:26-29).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class ModelConfig:
    obs_dim: int
    act_dim: int
    window: int           # obs_seq_len  (W)
    goal_len: int         # goal_seq_len (G)
    d: int                # embed_dim
    n_layers: int
    n_heads: int
    sigma_data: float = 0.5
    linear_output: bool = True
    goal_conditioned: bool = True

    @property
    def G(self) -> int:
        return self.goal_len if self.goal_conditioned else 0

    @property
    def block_size(self) -> int:      # score_gpts.py:148
        return self.G + 2 * self.window + 1

    @property
    def seq_size(self) -> int:        # score_gpts.py:150 (last pos_emb row is never read)
        return self.G + self.window + 1

    @property
    def head_dim(self) -> int:
        return self.d // self.n_heads

    def n_tokens(self, t: int | None = None) -> int:
        t = self.window if t is None else t
        return 1 + self.G + 2 * t

    def fwd_flops_per_seq(self, t: int | None = None) -> float:
        """Algorithmic FLOPs of one model evaluation of one sequence."""
        W = self.window if t is None else t
        T, d, G = self.n_tokens(W), self.d, self.G
        f = self.n_layers * (24 * T * d * d + 4 * T * T * d)
        f += 2 * d * (self.obs_dim * (W + G) + self.act_dim * W + 1)
        f += 2 * W * d * self.act_dim
        return float(f)

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """Parameters in ``nn.Module.parameters()`` order of the reference
        (verified against the reference and its checkpoints)."""
        d, P = self.d, "inner_model."
        out = [(P + "pos_emb", (1, self.seq_size, d)),
               (P + "tok_emb.weight", (d, self.obs_dim)), (P + "tok_emb.bias", (d,))]
        for l in range(self.n_layers):
            b = f"{P}blocks.{l}."
            out += [(b + "ln1.weight", (d,)), (b + "ln1.bias", (d,)),
                    (b + "ln2.weight", (d,)), (b + "ln2.bias", (d,))]
            for n in ("key", "query", "value", "proj"):
                out += [(b + f"attn.{n}.weight", (d, d)), (b + f"attn.{n}.bias", (d,))]
            out += [(b + "mlp.0.weight", (4 * d, d)), (b + "mlp.0.bias", (4 * d,)),
                    (b + "mlp.2.weight", (d, 4 * d)), (b + "mlp.2.bias", (d,))]
        out += [(P + "ln_f.weight", (d,)), (P + "ln_f.bias", (d,)),
                (P + "sigma_emb.weight", (d, 1)), (P + "sigma_emb.bias", (d,)),
                (P + "action_emb.weight", (d, self.act_dim)), (P + "action_emb.bias", (d,))]
        if self.linear_output:
            out += [(P + "action_pred.weight", (self.act_dim, d)), (P + "action_pred.bias", (self.act_dim,))]
        else:
            out += [(P + "action_pred.0.weight", (100, d)), (P + "action_pred.0.bias", (100,)),
                    (P + "action_pred.2.weight", (self.act_dim, 100)), (P + "action_pred.2.bias", (self.act_dim,))]
        return out

    def n_params(self) -> int:
        n = 0
        for _, s in self.param_shapes():
            k = 1
            for v in s:
                k *= v
            n += k
        return n


# BASELINE.md / SURVEY.md 8d shapes
K256 = ModelConfig(obs_dim=60, act_dim=9, window=10, goal_len=2, d=256, n_layers=4, n_heads=4)
B256 = ModelConfig(obs_dim=16, act_dim=2, window=10, goal_len=1, d=256, n_layers=4, n_heads=4)
T16 = ModelConfig(obs_dim=60, act_dim=9, window=7, goal_len=1, d=256, n_layers=4, n_heads=4)
# frozen configs beside the shipped checkpoints (trained_models/*/.hydra/config.yaml)
KITCHEN_CKPT = ModelConfig(obs_dim=30, act_dim=9, window=4, goal_len=2, d=360, n_layers=6, n_heads=6)
BLOCKPUSH_CKPT = ModelConfig(obs_dim=10, act_dim=2, window=5, goal_len=1, d=240, n_layers=4, n_heads=12)
