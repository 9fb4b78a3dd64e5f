"""cuBLAS bf16 (torch.matmul) on THIS pass's GEMM shapes, sustained: the library's rate for the five matmuls of one
4096-triplet pass alone -- no gather, bias, activation, cosine or hand-off -- as a like-for-like ceiling beside the
8192^3 figure in MEASURED_PEAKS.json.  Four streams, one CUDA graph each (like bench.py's library baseline), regions of
>= 60 ms, the first region dropped (the power cap has not engaged yet)."""
import statistics, sys
import torch
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
shapes = {"G.L0": (320, 1024), "D.L0": (384, 1024), "G.L1": (1024, 1024), "D.L1": (1024, 512), "G.L2": (1024, 128)}
S = 4

def run(names, label):
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    ops = []
    for s in range(S):
        xs = {n: torch.randn(B, shapes[n][0], device=dev, dtype=torch.bfloat16) for n in names}
        ws = {n: torch.randn(shapes[n][1], shapes[n][0], device=dev, dtype=torch.bfloat16) for n in names}
        outs = {n: torch.empty(B, shapes[n][1], device=dev, dtype=torch.bfloat16) for n in names}
        ops.append((xs, ws, outs))
    flop = sum(2.0 * B * shapes[n][0] * shapes[n][1] for n in names)
    reps = 40
    graphs = []
    for s in range(S):
        xs, ws, outs = ops[s]
        with torch.cuda.stream(streams[s]):
            for n in names:
                torch.matmul(xs[n], ws[n].T, out=outs[n])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(streams[s]):
            with torch.cuda.graph(g, stream=streams[s]):
                for _ in range(reps):
                    for n in names:
                        torch.matmul(xs[n], ws[n].T, out=outs[n])
        graphs.append(g)
    torch.cuda.synchronize()
    def region(n_replays):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for s in streams:
            s.wait_stream(main)
        for _ in range(n_replays):
            for s in range(S):
                with torch.cuda.stream(streams[s]):
                    graphs[s].replay()
        for s in streams:
            main.wait_stream(s)
        e1.record(main); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    ms = region(2)
    n_rep = max(2, int(60.0 / (ms / 2)) + 1)
    t = [region(n_rep) for _ in range(5)]
    per = lambda x: x / (n_rep * S * reps)          # ms per pass-equivalent
    med = statistics.median(t[1:])
    print(f"{label:34s} first region {flop / per(t[0]) / 1e9:7.1f} TFLOP/s   sustained {flop / per(med) / 1e9:7.1f} TFLOP/s "
          f"= {B / per(med) / 1e3:6.1f} M samples/s-equivalent   ({per(med) * 1e3:.2f} us per {B} rows)", flush=True)

for n in shapes:
    run([n], f"{n} {B}x{shapes[n][0]}x{shapes[n][1]}")
run(list(shapes), "all five GEMMs of one pass")
