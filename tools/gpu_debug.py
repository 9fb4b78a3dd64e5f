"""GPU bring-up diagnostic (run on the B200 box): per-layer tensor-core GEMM vs a torch fp32 product of the same
bf16-rounded operands, with an error map that localises descriptor / swizzle / pipeline mistakes.
Writes gpurun_out/debug.log.  Uses the oracle only as the checker."""
import os
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pro-b-gan_b200"))
sys.path.insert(0, str(ROOT))
import torch

from oracle import prot_b_gan_oracle as oracle
from pbg import synth
import modular_prot_b_gan as m

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
log = open(OUT / "debug.log", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


def errmap(got, ref, name, rb=32, cb=64):
    d = (got.float() - ref.float()).abs()
    P(f"[{name}] shape={tuple(got.shape)} max_abs={d.max().item():.4e} ref_max={ref.abs().max().item():.4e} "
      f"nan={torch.isnan(got.float()).sum().item()}")
    if d.max().item() > 0.05 * max(ref.abs().max().item(), 1e-6) and got.dim() == 2:
        R, Cc = d.shape
        rows = min(R, 256)
        mp = d[:rows].reshape(rows // rb if rows >= rb else 1, -1, Cc)
        mp = mp.amax(1)  # [rowblocks, C]
        ncb = (Cc + cb - 1) // cb
        for i in range(mp.shape[0]):
            P("   rows %4d.. :" % (i * rb), " ".join("%7.1e" % mp[i, j * cb:(j + 1) * cb].max().item() for j in range(ncb)))


def main():
    dev = torch.device("cuda:0")
    P(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    Go, Do = synth.make_models(oracle.ModularGenerator, oracle.ModularDiscriminator)
    G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
    G, D = G.to(dev), D.to(dev)
    eng = m.make_fused_engine(G, D)
    P("engine created; launches so far", eng.launch_count)
    slope = 0.2
    gl = [(w.to(dev), b.to(dev)) for w, b in G.folded_layers()]
    dl = [(w.to(dev), b.to(dev)) for w, b in D.folded_layers()]

    def ref_lin(a_bf, w, b):
        return a_bf.float() @ w.bfloat16().float().T + b

    for M in (128, 200, 4096):
        torch.manual_seed(M)
        # generator layers
        a0 = (torch.randn(M, 320, device=dev) * 0.5).bfloat16()
        o0 = eng.linear_bf16(0, 0, a0); torch.cuda.synchronize()
        r0 = torch.nn.functional.leaky_relu(ref_lin(a0, *gl[0]), slope)
        errmap(o0, r0, f"G.L0 M={M} K=320 N=1024")
        a1 = r0.bfloat16()
        o1 = eng.linear_bf16(0, 1, a1); torch.cuda.synchronize()
        r1 = torch.nn.functional.leaky_relu(ref_lin(a1, *gl[1]), slope)
        errmap(o1, r1, f"G.L1 M={M} K=1024 N=1024")
        a2 = r1.bfloat16()
        o2 = eng.linear_bf16(0, 2, a2); torch.cuda.synchronize()
        r2 = torch.tanh(ref_lin(a2, *gl[2]))
        errmap(o2, r2, f"G.L2 M={M} K=1024 N=128 tanh")
        # discriminator layers
        d0 = (torch.randn(M, 384, device=dev) * 0.5).bfloat16()
        p0 = eng.linear_bf16(1, 0, d0); torch.cuda.synchronize()
        q0 = torch.nn.functional.leaky_relu(ref_lin(d0, *dl[0]), slope)
        errmap(p0, q0, f"D.L0 M={M} K=384 N=1024")
        d1 = q0.bfloat16()
        p1 = eng.linear_bf16(1, 1, d1); torch.cuda.synchronize()
        q1 = torch.nn.functional.leaky_relu(ref_lin(d1, *dl[1]), slope) @ dl[2][0].reshape(-1) + dl[2][1]
        errmap(p1.reshape(-1, 1), q1.reshape(-1, 1), f"D.L1+rowdot M={M} K=1024 N=512")

    # whole network, both precisions, vs the CPU oracle
    node_emb, rel_w = synth.make_tables()
    for B in (16, 256, 4096):
        trip = synth.make_triplets(B)
        z = synth.make_latents(B)
        with torch.no_grad():
            h, r, t = node_emb[trip[:, 0]], rel_w[trip[:, 1]], node_emb[trip[:, 2]]
            g_ref = Go(h, r, z)
            d_ref = Do(h, r, t)
            cs_ref = torch.nn.functional.cosine_similarity(g_ref, t, dim=1)
        for prec in ("fp32", "bf16"):
            res = eng.score_triplets(node_emb.to(dev), rel_w.to(dev), trip.to(dev), z.to(dev), want_gen_out=True,
                                     want_gen_scores=True, want_disc=True, precision=prec)
            torch.cuda.synchronize()
            eng.check_indices()
            errmap(res["gen_out"].cpu(), g_ref, f"G full B={B} {prec}")
            errmap(res["logits"].cpu().reshape(-1, 1), d_ref.reshape(-1, 1), f"D logits B={B} {prec}")
            errmap(res["gen_scores"].cpu().reshape(-1, 1), cs_ref.reshape(-1, 1), f"G cosine B={B} {prec}")
            errmap(res["probs"].cpu().reshape(-1, 1), torch.sigmoid(d_ref).reshape(-1, 1), f"D probs B={B} {prec}")
    P("launches", eng.launch_count)


try:
    main()
    P("DEBUG DONE")
except Exception:
    P(traceback.format_exc())
    sys.exit(1)
