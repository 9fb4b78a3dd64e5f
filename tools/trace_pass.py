"""Per-CTA / per-item timeline of the fused pass kernel (run on the B200 box).  Appends to gpurun_out/trace_pass.log."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "pro-b-gan_b200"), str(ROOT)]
import torch

from pbg import synth
import modular_prot_b_gan as m

out = open(ROOT / "gpurun_out" / "trace_pass.log", "a")
KIND = {0: "G_L0", 1: "D_L0", 2: "G_L1", 3: "D_L1", 4: "G_L2", 5: "GATH"}


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True); out.write(s + "\n"); out.flush()


dev = torch.device("cuda:0")
G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
import os
CTAS = int(os.environ.get("TRACE_CTAS", "0"))   # SMs per pass (0 = all), e.g. 48 = bench.py's lane width
eng = m.make_fused_engine(G.to(dev), D.to(dev), ctas=CTAS)
node_emb, rel_w = (t.to(dev) for t in synth.make_tables())
for B in [int(x) for x in (sys.argv[1:] or ["4096"])]:
    trip, z = synth.make_triplets(B).to(dev), synth.make_latents(B).to(dev)
    outb = {"gen_out": torch.empty(B, 128, dtype=torch.bfloat16, device=dev), "gen_scores": torch.empty(B, device=dev),
            "logits": torch.empty(B, device=dev), "probs": torch.empty(B, device=dev)}

    STAGED = int(os.environ.get("TRACE_STAGED", "0"))   # 1: the rows are staged by pbg_stage_triplets (gather phase off)

    def run():
        if STAGED:
            eng.stage_triplets(0, node_emb, rel_w, trip, z)
            eng.score_staged(0, want_gen_out=True, want_gen_scores=True, want_disc=True, out_dtype=torch.bfloat16, out=outb)
        else:
            eng.score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True, want_disc=True,
                               precision="bf16", out_dtype=torch.bfloat16, out=outb)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    n = 50
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                run()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (2 * n) * 1e3
    eng.debug_trace(True)
    run()
    ti64 = eng.debug_trace(False)
    ti64 = ti64[ti64[:, 0] > 0]
    t = ti64.double()
    hdr, items = t[:, :16], t[:, 16:240].reshape(t.shape[0], 56, 4)
    raw0 = ti64[:, 16:240].reshape(t.shape[0], 56, 4)[:, :, 0]
    c0 = hdr[:, 0:1]
    P(f"pass B={B} on {CTAS or 'all'} SMs: {us:.2f} us/step (graph of {n}) = {B / us:.1f} M samples/s = {B * 4850688 / us / 1e6:.0f} TFLOP/s; {t.shape[0]} CTAs")
    P(f"   CTA lifetime (begin -> epilogue end): mean {(hdr[:, 8] - hdr[:, 0]).mean():.0f} max {(hdr[:, 8] - hdr[:, 0]).max():.0f} clk;"
      f"  begin skew (globaltimer) {(hdr[:, 14].max() - hdr[:, 14].min()):.0f} ns")
    P(f"   per CTA: items {hdr[:, 12].mean():.1f}  k-blocks {hdr[:, 9].mean():.1f} (max {hdr[:, 9].max():.0f})  -> MMA floor {hdr[:, 9].mean() * 512:.0f} clk")
    P(f"   producer: dep wait {hdr[:, 11].mean():.0f} (max {hdr[:, 11].max():.0f})  empty wait {hdr[:, 1].mean():.0f}")
    P(f"   mma: full wait {hdr[:, 3].mean():.0f} (max {hdr[:, 3].max():.0f})  tmem wait {hdr[:, 4].mean():.0f}")
    ln = t[:, 238].sum().clamp_min(1)
    P(f"   leader's loads that the MMA waited for: {int(t[:, 238].sum())} of {int(hdr[:, 9].sum())} k-blocks, issue -> both CTAs' data landed: mean {t[:, 237].sum() / ln:.0f} clk, max {t[:, 239].max():.0f}")
    P(f"   epilogue warp 2: acc wait {hdr[:, 7].mean():.0f}  busy {hdr[:, 10].mean():.0f}")
    ph = t[:, 240:256]
    nchunk = ph[:, 3].clamp_min(1); ntile = ph[:, 5].clamp_min(1)
    P(f"   mma: waiting for an item {hdr[:, 15][hdr[:, 9] > 0].mean():.0f}")
    P(f"   STORE epilogue phases (warp 2, clk per 64-col chunk): staging free {(ph[:, 4] / nchunk).mean():.0f}  tmem ld+wait {(ph[:, 0] / nchunk).mean():.0f}"
      f"  bias+math+STS {(ph[:, 1] / nchunk).mean():.0f}  store {(ph[:, 2] / nchunk).mean():.0f}")
    P(f"      of bias+math+STS: first half math {(ph[:, 11] / nchunk).mean():.0f}  second TMEM wait {(ph[:, 12] / nchunk).mean():.0f}")
    if ph[:, 13].max() > 0:
        P(f"   kernel body (globaltimer): first CTA entry -> roles start {(hdr[:, 14].min() - ph[:, 14][ph[:, 14] > 0].min()) / 1e3:.2f} us (entry skew {(ph[:, 14].max() - ph[:, 14][ph[:, 14] > 0].min()) / 1e3:.2f}),"
          f" -> last epilogue end ~{(hdr[:, 8] - hdr[:, 0]).max() / 1.92e3:.2f} us later, -> counters reset at {(ph[:, 13].max() - ph[:, 14][ph[:, 14] > 0].min()) / 1e3:.2f} us")
    tsel = t[:, 236] > 0
    if tsel.any():
        tt = t[tsel]
        P(f"   TANH epilogue (warp 2, last tile): start -> tmem loaded {(tt[:, 233] - tt[:, 232]).mean():.0f}  tanh {(tt[:, 234] - tt[:, 233]).mean():.0f}"
          f"  cosine {(tt[:, 235] - tt[:, 234]).mean():.0f} (tail dots {(tt[:, 245] - tt[:, 234]).mean():.0f}, partner wait {(tt[:, 255] - tt[:, 245]).mean():.0f},"
          f" finish + store {(tt[:, 235] - tt[:, 255]).mean():.0f})  output {(tt[:, 236] - tt[:, 235]).mean():.0f}")
    pfn = ph[:, 8].clamp_min(1)
    P(f"   tile hand-off (warp 2, clk per tile): bulk-store completion wait {(ph[:, 6] / pfn).mean():.0f}  whole arrival {(ph[:, 7] / pfn).mean():.0f}")
    gsel = ph[:, 9] > 0
    if gsel.any():
        P(f"   phase-0 gather (warp 2): start@ {(ph[:, 9] - hdr[:, 0])[gsel].mean():.0f}  loads+stores issued@ {(ph[:, 10] - hdr[:, 0])[gsel].mean():.0f}"
          f"  arrived@ {(hdr[:, 5] - hdr[:, 0])[gsel].mean():.0f}")
    kind = raw0 & 0xff
    mblk = (raw0 >> 8) & 0xfff
    valid = raw0 != 0
    valid[:, 54:] = False
    claim_r = (((raw0 >> 20) - (ti64[:, 0:1] & ((1 << 43) - 1))) & ((1 << 43) - 1)).double()  # clock64 << 20 keeps 43 bits
    for k, name in KIND.items():
        sel = valid & (kind == k)
        if sel.sum() == 0:
            continue
        cl = claim_r[sel]
        dep = items[:, :, 1][sel]; accr = items[:, :, 2][sel]; done = items[:, :, 3][sel]
        c0s = c0.expand(-1, 56)[sel]
        depw = torch.where(dep > 0, dep - c0s - cl, torch.zeros_like(dep))
        P(f"   {name}: n={int(sel.sum()):5d} claim@ {cl.mean():8.0f} (min {cl.min():.0f} max {cl.max():.0f})  claim->mma0 {depw.mean():7.0f} (max {depw.max():.0f})"
          f"  claim->acc {((accr - c0s) - cl).mean():7.0f}  epilogue {(done - accr).mean():6.0f} (max {(done - accr).max():.0f})  done@ max {(done - c0s).max():.0f}")
    if B <= 4096:
        for cta in (0, 1, 10, 20, 30, 46) if CTAS else (0, 60, 140):
            if cta >= t.shape[0]:
                continue
            rows = []
            for i in range(56):
                if not valid[cta, i]:
                    break
                rows.append(f"{KIND[int(kind[cta, i])]}[m{int(mblk[cta, i])}] claim {claim_r[cta, i]:.0f} mma0 {max(items[cta, i, 1] - c0[cta, 0], 0):.0f} acc {items[cta, i, 2] - c0[cta, 0]:.0f} done {items[cta, i, 3] - c0[cta, 0]:.0f}")
            P(f"   CTA {cta} (end {hdr[cta, 8] - hdr[cta, 0]:.0f}): " + " | ".join(rows))
