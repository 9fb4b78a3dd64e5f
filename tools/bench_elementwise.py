"""The HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json: hbm_gbs):
   stage_rows_kernel (gather + concat + bf16 cast of one request, pbg_stage_triplets) and topk_prepare_kernel (row norms +
   normalised bf16 copy of the entity table, pbg_topk_prepare).  Algorithmic bytes per row:
     stage:   read 3 E fp32 + Z fp32 + 24 B of ids, write (2E + Z) + 3E bf16 + E fp32   = 1816 + 1920 B at E = 128, Z = 64
     prepare: read E fp32 (twice, the second pass hits L1 / L2), write E bf16 + 4 B      =  512 +  260 B at E = 128"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch
from pbg import synth
import modular_prot_b_gan as m

dev = torch.device("cuda:0")
pk = ROOT / "MEASURED_PEAKS.json"
peak = json.loads(pk.read_text())["hbm_gbs"] if pk.exists() else 6650.0
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
eng = m.make_fused_engine(G.to(dev), D.to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=20, cold=True):
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()          # 256 MiB > L2: the next launch reads from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for N in (65536, 1 << 20):
    node_emb, rel_w = (t.to(dev) for t in synth.make_tables(N, 64, 128))
    for B in (4096, 65536):
        trip, z = synth.make_triplets(B, N, 64).to(dev), synth.make_latents(B).to(dev)
        eng.reserve(B, "bf16", 2)
        us = timed(lambda: eng.stage_triplets(0, node_emb, rel_w, trip, z))
        by = B * (1816 + 1920)
        print(f"stage_rows_kernel   N={N:8d} B={B:6d}: {us:8.1f} us  {by / us / 1e3:7.1f} GB/s = {by / us / 1e3 / peak:.2f} of the measured "
              f"copy bandwidth ({peak:.0f} GB/s); L2 flushed before every launch", flush=True)
    us = timed(lambda: eng._lib.pbg_topk_prepare(eng._h, node_emb.data_ptr(), N, torch.cuda.current_stream().cuda_stream))
    by = N * (512 + 260)
    print(f"topk_prepare_kernel N={N:8d}          : {us:8.1f} us  {by / us / 1e3:7.1f} GB/s = {by / us / 1e3 / peak:.2f} of the measured copy bandwidth",
          flush=True)
