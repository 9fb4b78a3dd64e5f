"""ORACLE -- test infrastructure, not product code.  PARITY UNPINNED.

CPU restatement of the module the reference imports but does not ship:
``from modular_prot_b_gan import ModularGenerator, ModularDiscriminator``
(/root/reference/pro_b_gan_infer.py:41).  The reference contains no model source,
no tests, no golden vectors and no sample checkpoint, so the
layer graph below is this repository's own frozen specification ("parity unpinned":
nothing in the reference pins the arithmetic of Generator / Discriminator).  What
the reference *does* pin, and what this file follows line by line, is the
boundary:

  * ``Generator(embed_dim, noise_dim)``                       pro_b_gan_infer.py:93
  * ``Generator.forward(h_emb, r_emb) -> [B, E]``             pro_b_gan_infer.py:143, :201
  * ``Discriminator(embed_dim, hidden_dim)``                  pro_b_gan_infer.py:94
  * ``Discriminator.forward(h, r, t) -> [B]`` logits          pro_b_gan_infer.py:301 (.item() at B=1)
  * ``Discriminator.score_triplets(node_emb, rel_emb, trip)`` pro_b_gan_infer.py:207
        -> (logits[B], probs[B]), prob = sigmoid(logit)       pro_b_gan_infer.py:302, :399-400
  * nn.Module plumbing: .to / .load_state_dict (strict) / .eval   pro_b_gan_infer.py:93-98, :106-107

Only stock ``torch.nn`` layers are used so the arithmetic is unarguably PyTorch's.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file, and only as the checker.

Frozen layer graph (also recorded in BASELINE.json ``oracle_spec`` and DESIGN.md):

  Generator      x = cat[h, r, z]                     [B, 2E+Z]   (E=128, Z=64 -> 320)
                 Linear(2E+Z, H) -> BatchNorm1d(H) -> LeakyReLU(0.2)       (H=1024)
                 Linear(H, H)    -> BatchNorm1d(H) -> LeakyReLU(0.2)
                 Linear(H, E)    -> Tanh                                   -> [B, E]
  Discriminator  x = cat[h, r, t]                     [B, 3E]     (384)
                 Linear(3E, H)   -> LeakyReLU(0.2)                         (H=1024)
                 Linear(H, H/2)  -> LeakyReLU(0.2)
                 Linear(H/2, 1)  -> squeeze(-1)                            -> [B] logits
"""
from __future__ import annotations

import torch
import torch.nn as nn

LEAKY_SLOPE = 0.2
DEFAULT_G_HIDDEN = 1024
LATENT_SEED = 1234  # latents from a CPU torch.Generator seeded 1234


class ModularGenerator(nn.Module):
    """(head emb, relation emb[, latent]) -> predicted tail embedding.

    Constructor takes exactly the two positional arguments the reference passes
    (pro_b_gan_infer.py:93); the hidden width is an internal default because the
    reference never hands ``hidden_dim`` to the generator.
    """

    def __init__(self, embed_dim: int, noise_dim: int, hidden_dim: int = DEFAULT_G_HIDDEN):
        super().__init__()
        self.embed_dim, self.noise_dim, self.hidden_dim = embed_dim, noise_dim, hidden_dim
        self.net = nn.Sequential(
            nn.Linear(2 * embed_dim + noise_dim, hidden_dim),
            nn.BatchNorm1d(hidden_dim),
            nn.LeakyReLU(LEAKY_SLOPE),
            nn.Linear(hidden_dim, hidden_dim),
            nn.BatchNorm1d(hidden_dim),
            nn.LeakyReLU(LEAKY_SLOPE),
            nn.Linear(hidden_dim, embed_dim),
            nn.Tanh(),
        )
        # forward() takes no noise argument in the reference (:143, :201): the latent
        # is drawn inside, from a module-owned *CPU* generator so that the same seed
        # gives the same latents whatever device the module lives on.
        self._latent_gen = torch.Generator(device="cpu")
        self._latent_gen.manual_seed(LATENT_SEED)

    def reseed(self, seed: int = LATENT_SEED) -> None:
        self._latent_gen.manual_seed(seed)

    def sample_latent(self, batch: int) -> torch.Tensor:
        return torch.randn(batch, self.noise_dim, generator=self._latent_gen, dtype=torch.float32)

    def forward(self, h_emb: torch.Tensor, r_emb: torch.Tensor, z: torch.Tensor | None = None) -> torch.Tensor:
        if z is None:
            z = self.sample_latent(h_emb.shape[0]).to(h_emb.device)
        return self.net(torch.cat([h_emb, r_emb, z], dim=-1))


class ModularDiscriminator(nn.Module):
    """(head, relation, tail) embeddings -> one real/fake logit per triplet."""

    def __init__(self, embed_dim: int, hidden_dim: int):
        super().__init__()
        self.embed_dim, self.hidden_dim = embed_dim, hidden_dim
        self.net = nn.Sequential(
            nn.Linear(3 * embed_dim, hidden_dim),
            nn.LeakyReLU(LEAKY_SLOPE),
            nn.Linear(hidden_dim, hidden_dim // 2),
            nn.LeakyReLU(LEAKY_SLOPE),
            nn.Linear(hidden_dim // 2, 1),
        )

    def forward(self, h_emb: torch.Tensor, r_emb: torch.Tensor, t_emb: torch.Tensor) -> torch.Tensor:
        # 1-D [B]: .item() must work at B=1 (:301) and results[...][0] must be a float (:399-400)
        return self.net(torch.cat([h_emb, r_emb, t_emb], dim=-1)).squeeze(-1)

    def score_triplets(self, node_emb: torch.Tensor, rel_emb: nn.Embedding, triplets: torch.Tensor):
        # Called with the raw tables + the [B,3] int64 index tensor (:207); gathers the
        # same way the script does at :186-188 (advanced indexing / nn.Embedding call).
        h = node_emb[triplets[:, 0]]
        r = rel_emb(triplets[:, 1])
        t = node_emb[triplets[:, 2]]
        logits = self.forward(h, r, t)
        return logits, torch.sigmoid(logits)  # prob = sigmoid(logit), :302


# The script's globals must contain these two names when ProtBGANInference.__init__
# runs (pro_b_gan_infer.py:93-94); pro-b-gan_b200/pbg/launcher.py and tests/golden/make_golden.py do the injection.
Generator = ModularGenerator
Discriminator = ModularDiscriminator


def cosine_topk(queries: torch.Tensor, table: torch.Tensor, k: int):
    """The entity-scoring tail of predict_tails / find_similar_entities, verbatim torch ops
    (pro_b_gan_infer.py:146-151, :231-236): normalise both sides, full similarity matrix, topk.
    Unlike the G / D graph above this part IS pinned by the reference: these are its own lines."""
    import torch.nn.functional as F
    q_norm = F.normalize(queries, dim=-1)
    t_norm = F.normalize(table, dim=-1)
    similarities = torch.matmul(q_norm, t_norm.T)
    return similarities.topk(k, dim=1)
