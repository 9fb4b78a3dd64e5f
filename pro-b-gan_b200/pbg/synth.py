"""Synthetic checkpoint in the reference's wire format (no sample checkpoint is shipped).

Schema follows what ``ProtBGANInference._load_checkpoint`` reads
(/root/reference/pro_b_gan_infer.py:74-112): ``args`` {embed_dim, noise_dim,
hidden_dim} :77-80, ``node_emb`` Tensor[N,E] :83, ``rel_emb`` {'weight': [R,E]}
:85/:103, ``generator`` / ``discriminator`` state_dicts :97-98 and the optional
``best_val_hit10`` / ``best_epoch`` / ``training_history`` :110-112.

Seeds are the ones frozen in SURVEY.md 8d so every run (oracle, CUDA path, bench,
golden fixtures) sees identical tensors:
  G/D weights  torch.manual_seed(0), PyTorch default Linear init
  BatchNorm    non-trivial stats from a CPU generator seeded 1 (default stats would
               make the load-time fold a near no-op and hide bugs)
  node_emb     N(0,1), seed 2          rel_emb  N(0,1), seed 3
  latents      N(0,1), CPU generator seed 1234, materialised once
  indices      randint, CPU generator seed 4321

The module classes are passed in, so the same function builds the checkpoint with
the CUDA-backed modules (bench, product) or with the oracle modules (tests).
"""
from __future__ import annotations

import torch
import torch.nn as nn

NUM_ENTITIES = 65536
NUM_RELATIONS = 64


def randomize_batchnorm(module: nn.Module, seed: int = 1) -> None:
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, nn.BatchNorm1d):
                n = m.num_features
                m.weight.copy_(torch.rand(n, generator=g) + 0.5)          # U(0.5, 1.5)
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)           # N(0, 0.1)
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)   # N(0, 0.1)
                m.running_var.copy_(torch.rand(n, generator=g) + 0.5)     # U(0.5, 1.5)


def make_models(gen_cls, disc_cls, embed_dim=128, noise_dim=64, hidden_dim=1024, g_hidden=None):
    """Build (G, D) on CPU with the frozen seeds; both returned in eval() mode."""
    torch.manual_seed(0)
    G = gen_cls(embed_dim, noise_dim) if g_hidden is None else gen_cls(embed_dim, noise_dim, g_hidden)
    D = disc_cls(embed_dim, hidden_dim)
    randomize_batchnorm(G, seed=1)
    return G.eval(), D.eval()


def make_tables(num_entities=NUM_ENTITIES, num_relations=NUM_RELATIONS, embed_dim=128):
    node_emb = torch.randn(num_entities, embed_dim, generator=torch.Generator().manual_seed(2))
    rel_w = torch.randn(num_relations, embed_dim, generator=torch.Generator().manual_seed(3))
    return node_emb, rel_w


def make_latents(batch: int, noise_dim: int = 64, seed: int = 1234) -> torch.Tensor:
    return torch.randn(batch, noise_dim, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def make_triplets(batch: int, num_entities=NUM_ENTITIES, num_relations=NUM_RELATIONS, seed: int = 4321) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    h = torch.randint(0, num_entities, (batch,), generator=g)
    r = torch.randint(0, num_relations, (batch,), generator=g)
    t = torch.randint(0, num_entities, (batch,), generator=g)
    return torch.stack([h, r, t], dim=1).contiguous()  # [B,3] int64


def make_checkpoint(gen_cls, disc_cls, embed_dim=128, noise_dim=64, hidden_dim=1024, g_hidden=None,
                    num_entities=NUM_ENTITIES, num_relations=NUM_RELATIONS) -> dict:
    G, D = make_models(gen_cls, disc_cls, embed_dim, noise_dim, hidden_dim, g_hidden)
    node_emb, rel_w = make_tables(num_entities, num_relations, embed_dim)
    return {
        "args": {"embed_dim": embed_dim, "noise_dim": noise_dim, "hidden_dim": hidden_dim},
        "node_emb": node_emb,
        "rel_emb": {"weight": rel_w},
        "generator": {k: v.clone() for k, v in G.state_dict().items()},
        "discriminator": {k: v.clone() for k, v in D.state_dict().items()},
        "best_val_hit10": 0.4242,
        "best_epoch": 7,
        "training_history": {},
    }
