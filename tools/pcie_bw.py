"""Host -> device copy bandwidth from pinned memory at the e2e leg's block size (bench.py: 1 146 880 B per 4096-triplet
request) and at 64 MB: the ceiling of `e2e` (280 B per sample in) on this box."""
import torch
dev = torch.device("cuda:0")
for nbytes in (1146880, 64 << 20):
    n = max(8, min(400, (256 << 20) // nbytes))
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(min(n, 16))]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(min(n, 16))]
    for streams in (1, 2):
        ss = [torch.cuda.Stream(dev) for _ in range(streams)]
        for _ in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in ss:
                s.wait_stream(torch.cuda.current_stream())
            for i in range(n):
                with torch.cuda.stream(ss[i % streams]):
                    ds[i % len(ds)].copy_(hs[i % len(hs)], non_blocking=True)
            for s in ss:
                torch.cuda.current_stream().wait_stream(s)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"H2D {nbytes:>9d} B x {n} copies on {streams} stream(s): {nbytes * n / ms * 1e-6:6.1f} GB/s "
              f"({ms / n * 1e3:6.1f} us per copy) -> {nbytes * n / ms * 1e-6 / 280 * 1e3:6.1f} M samples/s at 280 B/sample", flush=True)
