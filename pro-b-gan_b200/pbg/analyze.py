"""Batched ``analyze_relations``.

The reference walks |H| x |T| x R single-row discriminator calls, each followed by ``.item()`` -- one launch and one
device sync per (head, tail, relation) (pro_b_gan_infer.py:290-318).  Here the |H| x |T| x R triplets are built once on
the device, scored by ONE discriminator pass (``Discriminator.score_triplets``, the fused kernel on the CUDA modules),
and ranked per (head, tail) pair on the device; the only host transfer is the [pairs, top_k] result.

Result schema and ordering are the reference's: per pair the ``top_k`` relations by probability, descending, ties in
relation-id order (its ``list.sort(key=probability, reverse=True)`` is stable); ``probability`` is the fp32 sigmoid of
the fp32 logit (:302-303)."""
from __future__ import annotations

from typing import Any, Dict, List

import torch


def analyze_relations_batched(discriminator, node_emb: torch.Tensor, rel_emb, head_ids: List[int], tail_ids: List[int],
                              top_k: int = 5, model_hit10=None) -> Dict[str, Any]:
    """Drop-in for ``ProtBGANInference.analyze_relations`` given its ``discriminator`` / ``node_emb`` / ``rel_emb``."""
    dev = node_emb.device
    rel_w = rel_emb.weight if hasattr(rel_emb, "weight") else rel_emb
    R = int(rel_w.shape[0])
    H, T = len(head_ids), len(tail_ids)
    results: Dict[str, Any] = {
        "relation_analysis": [],
        "metadata": {"num_head_entities": H, "num_tail_entities": T, "top_k": top_k, "model_hit10": model_hit10},
    }
    if H == 0 or T == 0:
        return results
    with torch.no_grad():
        heads = torch.tensor(head_ids, dtype=torch.int64, device=dev)
        tails = torch.tensor(tail_ids, dtype=torch.int64, device=dev)
        # pair-major, relation-minor: row (i * T + j) * R + r  <->  (head_ids[i], r, tail_ids[j])   (:291-296 loop order)
        trip = torch.stack([heads.repeat_interleave(T * R),
                            torch.arange(R, device=dev).repeat(H * T),
                            tails.repeat_interleave(R).repeat(H)], dim=1).contiguous()
        logits, _ = discriminator.score_triplets(node_emb, rel_emb, trip)        # one pass instead of H*T*R calls
        logits = logits.float().view(H * T, R)
        probs = torch.sigmoid(logits)                                            # fp32, as :303
        k = min(top_k, R)
        order = torch.sort(probs, dim=1, descending=True, stable=True).indices[:, :k]
        # one D2H of [pairs, k] x 3; .tolist() yields the Python int / float the reference's .item() calls produce
        top_logits = logits.gather(1, order).tolist()
        top_probs = probs.gather(1, order).tolist()
        order = order.tolist()
    for i, head_id in enumerate(head_ids):
        for j, tail_id in enumerate(tail_ids):
            p = i * T + j
            results["relation_analysis"].append({
                "head_entity": head_id,
                "tail_entity": tail_id,
                "top_relations": [{"relation_id": rid, "discriminator_score": sc, "probability": pr}
                                  for rid, sc, pr in zip(order[p], top_logits[p], top_probs[p])],
            })
    return results
