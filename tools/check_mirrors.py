"""Multi-GPU check of the result mirrors (run under torchrun, one rank per GPU):
every rank scores its shard of a batch and its pass writes the rows into every peer's window; after a barrier each
rank compares the assembled buffer with an NCCL all-gather of the same results.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_mirrors.py"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from pbg import synth, shard
import modular_prot_b_gan as m

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
E, B = 128, 4096
Bg = B * world
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
eng = m.make_fused_engine(G.to(dev), D.to(dev), ctas=48)
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
lo, hi = shard.shard_bounds(Bg, world, rank)
blk = B * (2 * E + 12)
sym = symm_mem.empty(world * blk, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
sym.zero_()
torch.cuda.synchronize(); dist.barrier()
mine = sym[rank * blk:(rank + 1) * blk]
f32 = mine[B * 2 * E:].view(torch.float32)
out = {"gen_out": mine[:B * 2 * E].view(torch.bfloat16).view(B, E), "gen_scores": f32[:B], "logits": f32[B:2 * B], "probs": f32[2 * B:3 * B]}
peers = [r for r in range(world) if r != rank]
base = [hdl.buffer_ptrs[r] + rank * blk for r in peers]
mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
modes = ["mirrors (one unicast store per peer)"] + (["multicast (multimem.st, NVSwitch replicates)"] if mc else [])
if rank == 0:
    print(f"world {world}: multicast mapping {'available' if mc else 'NOT available'}", flush=True)
ok = True
for mode in modes:
  sym.zero_(); torch.cuda.synchronize(); dist.barrier()
  if mode.startswith("mirrors"):
    eng.set_result_multicast()
    eng.set_result_mirrors(gen_out=base, gen_scores=[b + B * 2 * E for b in base], logits=[b + B * 2 * E + 4 * B for b in base],
                           probs=[b + B * 2 * E + 8 * B for b in base])
  else:
    eng.set_result_mirrors()
    b0 = mc + rank * blk
    eng.set_result_multicast(gen_out=b0, gen_scores=b0 + B * 2 * E, logits=b0 + B * 2 * E + 4 * B, probs=b0 + B * 2 * E + 8 * B)
  for rep in range(3):
    trip = synth.make_triplets(Bg, seed=4321 + rep)[lo:hi].contiguous().to(dev)
    z = synth.make_latents(Bg, seed=1234 + rep)[lo:hi].contiguous().to(dev)
    eng.score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True, want_disc=True,
                       precision="bf16", out_dtype=torch.bfloat16, out=out)
    torch.cuda.synchronize(); dist.barrier()
    ref = torch.empty(world * blk, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(ref, mine.clone())
    torch.cuda.synchronize()
    same = torch.equal(ref, sym)
    ok &= same
    print(f"rank {rank} {mode} rep {rep}: assembled buffer {'matches' if same else 'DIFFERS from'} the all-gather "
          f"({int((ref != sym).sum())} bytes differ), logits sum {float(out['logits'].sum()):.4f}", flush=True)
    dist.barrier()
eng.set_result_mirrors(); eng.set_result_multicast()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
