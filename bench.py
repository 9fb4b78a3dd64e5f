#!/usr/bin/env python
"""bench.py -- throughput of the PRO-B-GAN inference hot path (generator forward + discriminator scoring).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one canonical generator + discriminator pass (ProtBGANInference.score_triplets,
pro_b_gan_infer.py:186-209, at the tensor boundary) over one batch of synthetic triplets:
gather h/r/t rows -> G(h, r, z) -> cosine(pred, t) -> D(h, r, t) -> sigmoid.  One sample = one triplet.

Workload (BASELINE.json configs[2] at N = 1, configs[3] at N = 8): bf16 tensor-core mode, 4096 triplets per GPU
per step, batch-index sharded (weak scaling: 4096 x N triplets per step), outputs re-assembled on every rank
with an all-gather over NVLink at N > 1.  E = 128, Z = 64, H = 1024, 65536 entities, 64 relations, random-init
weights with the frozen seeds of pbg/synth.py.

Lanes: one 4096-triplet pass is a chain of dependent layers (gather -> L0 -> L1 -> L2) that cannot keep 148 SMs busy
for its whole duration, so the steps -- independent requests -- are issued round-robin on `--lanes` streams, each
with its own engine (ctx) whose passes occupy `--ctas` SMs; passes of different lanes run side by side.
`--lanes 1` is the single-stream, full-width configuration.

Printed JSON (one line, rank 0):
  value      samples/s, whole job, inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e        same metric through the C-ABI host entry point pbg_score_triplets_host (synchronous), one host thread
             per lane: per step H2D of the triplets + latents from pinned host memory and D2H of the result the
             reference's score_triplets returns (generator scores, discriminator logits and probabilities,
             pro_b_gan_infer.py:204-209)
  roofline   dominant kernel vs the measured bf16 tensor peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (oracle/prot_b_gan_oracle.py, a port: the reference ships no model) on the
             box's host cores, bounded sample
`--impl reference` times that CPU oracle alone as the reference arm (the reference is CPU PyTorch code).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT / "pro-b-gan_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

E, Z, H = 128, 64, 1024
NUM_ENTITIES, NUM_RELATIONS = 65536, 64
FLOP_G = 2 * ((2 * E + Z) * H + H * H + H * E)           # 3 014 656 per sample  
FLOP_D = 2 * (3 * E * H + H * (H // 2) + (H // 2) * 1)   # 1 836 032 per sample
FLOP_SAMPLE = FLOP_G + FLOP_D                            # 4 850 688
L2_BYTES = 126 * 2 ** 20
METRIC = "generator+discriminator samples/sec (score_triplets pass, bf16 tensor-core mode)"


def measured_peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def workload_name(batch: int, world: int) -> str:
    return (f"score_triplets G+D pass, bf16, {batch} triplets/GPU/step x {world} GPU = {batch * world} per step "
            f"(BASELINE configs[2] per GPU; configs[3] at 8 GPUs), E={E} Z={Z} H={H}, "
            f"{NUM_ENTITIES} entities, {NUM_RELATIONS} relations")


# ------------------------------------------------------------------------------------------ CPU oracle leg
def cpu_oracle_pass_factory(batch: int):
    """Returns (fn, cores): fn() runs one oracle G+D pass over `batch` triplets on the CPU (fp32, all threads)."""
    import torch.nn.functional as F
    from oracle import prot_b_gan_oracle as oracle  # checker / baseline only, never the product path
    from pbg import synth
    G, D = synth.make_models(oracle.ModularGenerator, oracle.ModularDiscriminator)
    node_emb, rel_w = synth.make_tables(NUM_ENTITIES, NUM_RELATIONS, E)
    rel_emb = torch.nn.Embedding(NUM_RELATIONS, E)
    rel_emb.load_state_dict({"weight": rel_w})
    trip, z = synth.make_triplets(batch), synth.make_latents(batch)

    def fn():
        with torch.no_grad():
            h, r, t = node_emb[trip[:, 0]], rel_emb(trip[:, 1]), node_emb[trip[:, 2]]   # :186-188
            pred = G(h, r, z)                                                            # :201
            cs = F.cosine_similarity(pred, t, dim=1)                                     # :202
            logits, probs = D.score_triplets(node_emb, rel_emb, trip)                    # :207
        return pred, cs, logits, probs

    return fn, torch.get_num_threads()


def time_cpu_oracle(batch: int, budget_s: float, max_reps: int = 200):
    fn, cores = cpu_oracle_pass_factory(batch)
    fn()  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    reps = 0
    while reps < max_reps and (time.perf_counter() - t0 < budget_s or reps < 3):
        fn()
        reps += 1
    dt = time.perf_counter() - t0
    return batch * reps / dt, cores, reps, dt


def run_reference(args, rank: int, world: int) -> None:
    """Reference arm: the reference's own implementation of the path is CPU PyTorch; its model module is not
    shipped, so the oracle port is what runs (kind = "port").  Rank 0 only."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every core
    fn, cores = cpu_oracle_pass_factory(args.batch)
    t0 = time.perf_counter(); fn(); t_one = time.perf_counter() - t0
    # bound the whole run to ~2 minutes: shrink the per-step sample if K full batches would take longer
    sample = args.batch
    total = (args.steps + args.warmup) * t_one
    if total > 120.0:
        sample = max(256, int(args.batch * 120.0 / total) // 256 * 256)
        fn, cores = cpu_oracle_pass_factory(sample)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, world), "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} triplets per step x {args.steps} steps, torch {torch.__version__} fp32, "
                                   f"{cores} threads of {os.cpu_count()} logical cores"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank: int, local_rank: int, world: int) -> None:
    import torch.distributed as dist
    import modular_prot_b_gan as m
    from pbg import synth, shard
    from pbg.clocks import ClockSampler

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    Bg = B * world
    K, W = args.steps, args.warmup
    # Lanes: a 4096-triplet pass is a chain of dependent layers and cannot keep 148 SMs busy for its whole duration,
    # so independent steps run side by side -- one engine (ctx) + stream per lane, each pass on `ctas` SMs.
    S = max(1, args.lanes)
    num_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    # default width: a third of the device per pass (more lanes than that keep a queue of CTAs behind every SM)
    ctas = args.ctas if args.ctas > 0 else (0 if S == 1 else max(2, (num_sms // min(S, 3)) // 2 * 2))
    G, D = synth.make_models(m.ModularGenerator, m.ModularDiscriminator)
    G, D = G.to(dev), D.to(dev)
    engines = [m.make_fused_engine(G, D, ctas=ctas) for _ in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    node_emb, rel_w = (t.to(dev) for t in synth.make_tables(NUM_ENTITIES, NUM_RELATIONS, E))

    # ---- input pool: distinct pre-staged batches whose footprint exceeds L2, visited round-robin
    per_batch = B * (3 * 8 + Z * 4 + E * 2 + 3 * 4)
    P = max(8, math.ceil(1.6 * L2_BYTES / per_batch))
    P = (P + S - 1) // S * S          # pool entry i always runs on lane i % S
    lo, hi = shard.shard_bounds(Bg, world, rank)
    # one contiguous result block per pool entry and rank: [gen_out bf16 B x E | scores | logits | probs fp32 B each].
    # N > 1: the blocks of all ranks for one step form that step's assembled output [world, block] on EVERY rank.
    #   exchange "p2p"  (default): the buffers live in symmetric memory and every pass writes its rows straight into
    #                   the peers' copies from the kernel epilogues (result mirrors, NVLink stores; no collective)
    #   exchange "nccl": one all-gather of the block per step on a per-lane communicator
    blk_bytes = B * (2 * E + 12)
    exchange = "none"
    peer_ptrs = None
    if world > 1:
        exchange = args.exchange
        if exchange == "p2p":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                sym = symm_mem.empty(P * world * blk_bytes, dtype=torch.uint8, device=dev)
                hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
                peer_ptrs = [int(x) for x in hdl.buffer_ptrs]
            except Exception as ex:  # symmetric memory unavailable on this box: fall back to the collective
                if rank == 0:
                    print(f"bench.py: symmetric memory unavailable ({ex}); using the NCCL all-gather", file=sys.stderr)
                exchange = "nccl"
        # every rank must take the same path
        flag = torch.tensor([1 if exchange == "p2p" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange, peer_ptrs = "nccl", None
    if exchange != "p2p":
        sym = torch.empty(P * max(world, 1) * blk_bytes, dtype=torch.uint8, device=dev)
    pool, mirrors = [], []
    for i in range(P):
        trip = synth.make_triplets(Bg, NUM_ENTITIES, NUM_RELATIONS, seed=4321 + i)[lo:hi].contiguous().to(dev)
        z = synth.make_latents(Bg, Z, seed=1234 + i)[lo:hi].contiguous().to(dev)
        off = (i * world + rank) * blk_bytes
        blk = sym[off:off + blk_bytes]
        f32 = blk[B * 2 * E:].view(torch.float32)
        out = {"gen_out": blk[:B * 2 * E].view(torch.bfloat16).view(B, E),
               "gen_scores": f32[0:B], "logits": f32[B:2 * B], "probs": f32[2 * B:3 * B], "block": blk,
               "assembled": sym[i * world * blk_bytes:(i + 1) * world * blk_bytes]}
        pool.append((trip, z, out))
        if exchange == "p2p":
            base = [peer_ptrs[r] + off for r in range(world) if r != rank]
            mirrors.append({"gen_out": base, "gen_scores": [b + B * 2 * E for b in base],
                            "logits": [b + B * 2 * E + 4 * B for b in base], "probs": [b + B * 2 * E + 8 * B for b in base]})
    if exchange == "nccl":
        # one communicator per lane: collectives of different lanes run on different streams, and NCCL requires the
        # collectives of ONE communicator to execute in the same order on every rank
        lane_pg = [dist.new_group(ranks=list(range(world)), backend="nccl") for _ in range(S)]

    def compute(i: int):
        """One pass, issued on the current stream."""
        trip, z, out = pool[i % P]
        if exchange == "p2p":
            engines[i % S].set_result_mirrors(**mirrors[i % P])
        engines[i % S].score_triplets(node_emb, rel_w, trip, z, want_gen_out=True, want_gen_scores=True,
                                      want_disc=True, precision="bf16", out_dtype=torch.bfloat16, out=out)
        if exchange == "nccl":  # reassemble the outputs on every rank (north_star: NVLink all-gather)
            dist.all_gather_into_tensor(out["assembled"], out["block"], group=lane_pg[i % S])
        return out

    def step(i: int):
        with torch.cuda.stream(streams[i % S]):
            compute(i)

    # ---- optional CUDA graphs: one graph per pool entry (pass + all-gather), captured and replayed on its lane's stream
    use_graphs = bool(args.graphs) and exchange != "nccl"   # NCCL inside per-lane graphs hung on this box: eager there
    graphs = []
    for lane in range(S):
        with torch.cuda.stream(streams[lane]):
            compute(lane)
    torch.cuda.synchronize()
    for e in engines:
        e.check_indices()
    l0 = engines[0].launch_count
    with torch.cuda.stream(streams[0]):
        compute(0)
    torch.cuda.synchronize()
    launches_per_step = engines[0].launch_count - l0
    if use_graphs:
        for i in range(P):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(streams[i % S]):
                with torch.cuda.graph(g, stream=streams[i % S]):
                    compute(i)
            graphs.append(g)
        torch.cuda.synchronize()

        def step(i: int):  # noqa: F811
            with torch.cuda.stream(streams[i % S]):
                graphs[i % P].replay()

    main = torch.cuda.current_stream(dev)

    def fork():
        e = torch.cuda.Event(); e.record(main)
        for s in streams:
            s.wait_event(e)

    def join():
        for s in streams:
            e = torch.cuda.Event(); e.record(s); main.wait_event(e)

    fork()
    for i in range(W):
        step(i)
    join()
    torch.cuda.synchronize(); barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize(); barrier()
        ev0.record(main)
        fork()
        for i in range(K):
            step(W + i)
        join()
        ev1.record(main)
        torch.cuda.synchronize(); barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    for e in engines:
        e.check_indices()
    value = Bg * K / (ms * 1e-3)
    gpu_launches = launches_per_step * K

    # ---- e2e: host buffers through the C-ABI host entry point (synchronous: H2D, pass, D2H, one sync per call);
    #      one host thread per lane keeps `S` calls in flight, each on its own ctx, like a multi-threaded server
    import threading
    hp = []
    for i in range(min(P, 16 * S)):
        trip = synth.make_triplets(Bg, NUM_ENTITIES, NUM_RELATIONS, seed=9000 + i)[lo:hi].contiguous().pin_memory()
        z = synth.make_latents(Bg, Z, seed=9500 + i)[lo:hi].contiguous().pin_memory()
        hp.append((trip, z))
    # a synchronous call spends most of its life in PCIe copies and the completion wake-up, so the server model is
    # more calls in flight than passes fit on the device: T host threads (default one per lane; --e2e-threads), one ctx each
    T = args.e2e_threads if args.e2e_threads > 0 else S   # measured: 6 / 12 / 18 threads give 117 / 121 / 116 M samples/s (PCIe-bound)
    for e in engines:
        e.set_result_mirrors()   # host results need no re-assembly: every rank's caller receives its own shard
    e2e_engines = engines + [m.make_fused_engine(G, D, ctas=ctas) for _ in range(max(0, T - S))]
    h_out = [(None, torch.empty(B).pin_memory(), torch.empty(B).pin_memory(), torch.empty(B).pin_memory())
             for _ in range(T)]  # score_triplets returns scores / logits / probabilities, not the predicted embeddings
    Ke = min(K, 2000)

    def e2e_worker(lane: int, first: int, last: int):
        torch.cuda.set_device(local_rank)
        h_gen, h_sc, h_lg, h_pb = h_out[lane]
        for i in range(first + lane, last, T):
            trip, z = hp[i % len(hp)]
            e2e_engines[lane].score_triplets_host(node_emb, rel_w, trip, z, h_gen, h_sc, h_lg, h_pb, precision="bf16")

    def e2e_run(first: int, last: int):
        ts = [threading.Thread(target=e2e_worker, args=(lane, first, last)) for lane in range(T)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    e2e_run(0, max(3 * T, min(W, 10)))
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    e2e_run(0, Ke)  # every call returns after the D2H of its results
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": Bg * Ke / e2e_s, "unit": "samples/s",
           "h2d_bytes_per_step": B * (3 * 8 + Z * 4) * world, "d2h_bytes_per_step": B * 3 * 4 * world,
           "steps": Ke, "api": f"pbg_score_triplets_host (C ABI, pinned host buffers, one sync per call), "
                               f"{T} host thread(s), one ctx each"}

    # ---- roofline of the dominant kernel: per-kernel CUDA events on its launch stream, the same lanes in flight
    peaks = measured_peaks()
    prof_steps = min(K, 60 * S)
    for e in engines:
        e.profile_enable(True); e.profile_read()
    fork()
    for i in range(prof_steps):
        with torch.cuda.stream(streams[i % S]):
            compute(i)
    join()
    torch.cuda.synchronize()
    prof = {}
    for e in engines:
        for k, v in e.profile_read().items():
            a = prof.setdefault(k, [0.0, 0])
            a[0] += v[0]; a[1] += v[1]
        e.profile_enable(False)
    flops = {"g_l0": 2 * B * (2 * E + Z) * H, "g_l1": 2 * B * H * H, "g_l2": 2 * B * H * E,
             "d_l0": 2 * B * 3 * E * H, "d_l1": 2 * B * H * (H // 2) + 2 * B * (H // 2)}
    flops["pass"] = B * FLOP_SAMPLE  # the fused kernel runs the whole G + D pass
    kinds = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in prof.items() if v[1] > 0}
    dom = max((k for k in kinds if k in flops), key=lambda k: prof[k][0])
    ach = flops[dom] / (kinds[dom]["ms_per_launch"] * 1e-3) / 1e12
    step_tflops = FLOP_SAMPLE * value / world / 1e12
    peak = peaks["bf16_burst"]  # timed regions here last well under a second: burst figure
    share = (ctas if ctas > 0 else num_sms) / num_sms
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get(dom)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": step_tflops, "peak": peak, "unit": "TFLOP/s",
                "frac": step_tflops / peak, "traffic": traffic, "peak_source": peaks["source"] + ", burst figure",
                "note": f"achieved = algorithmic flops of the {S} launches in flight / average launch duration x "
                        f"overlap, taken as flops per sample x measured throughput of the timed region; one launch "
                        f"alone: see per_launch (it occupies {ctas if ctas > 0 else num_sms} of {num_sms} SMs)",
                "flops_per_launch": flops[dom], "us_per_launch": kinds[dom]["ms_per_launch"] * 1e3,
                "per_launch": {"achieved": ach, "sm_share": share, "peak_share": peak * share, "frac_of_share": ach / (peak * share)},
                "whole_step": {"achieved": step_tflops, "frac": step_tflops / peak,
                               "frac_of_sustained": step_tflops / peaks["bf16_sustained"],
                               "flops_per_sample": FLOP_SAMPLE},
                "per_kernel_us": {k: round(v["ms_per_launch"] * 1e3, 3) for k, v in kinds.items()}}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, reps, dt = time_cpu_oracle(B, budget_s=12.0)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": f"{reps} passes of {B} triplets in {dt:.1f} s, oracle fp32, torch {torch.__version__}, "
                             f"{cores} threads of {os.cpu_count()} logical cores"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(B, world), "global_batch": Bg, "parallelism": f"dp{world}",
                       "l2": f"inputs rotate over {P} distinct pre-staged batches ({P * per_batch / 2**20:.0f} MiB "
                             f"> 126 MiB L2); no flush", "cuda_graphs": bool(use_graphs),
                       "lanes": S, "ctas_per_pass": ctas if ctas > 0 else num_sms,
                       "collective": {"none": "none", "p2p": "none: every pass writes its result rows into all peers' symmetric-memory "
                                      "windows from its epilogues (NVLink stores)", "nccl": "all-gather of outputs (NCCL, "
                                      "one per step)"}[exchange]},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=4096, help="triplets per GPU per step")
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=6, help="independent passes in flight (one ctx + stream each)")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p", help="N > 1: how the outputs are re-assembled")
    ap.add_argument("--e2e-threads", type=int, default=0, help="host threads of the e2e leg (0: one per lane)")
    ap.add_argument("--ctas", type=int, default=0, help="SMs per pass (0: all SMs / lanes, in whole CTA pairs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
