"""Summarise an `ncu --set full` report (run here, no GPU needed): key metrics per captured launch into a text file
under profiles/ and the dominant kernel's DRAM traffic per launch into profiles/traffic.json (read by bench.py).
    python tools/ncu_summary.py gpurun_out/prof_pass2_r1c.ncu-rep profiles/ncu_pass2_r1c.txt pass"""
import csv, io, json, subprocess, sys
from pathlib import Path

rep, out, kind = sys.argv[1], Path(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
lines, traffic = [], []
for r in rows[2:]:
    d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
    lines.append(f"kernel {d['Kernel Name'][0]}  (launch id {d['ID'][0]})")
    for w in want:
        if w in d and d[w][0] != "":
            lines.append(f"   {w} = {d[w][0]} {d[w][1]}")
    def to_bytes(key):
        v, u = d[key]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic.append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
    lines.append(f"   dram traffic (read + write) = {traffic[-1] / 1e6:.3f} MB per launch")
out.write_text("\n".join([f"source: {rep} (ncu --set full --clock-control none; cold caches, serialised launches)"] + lines) + "\n")
tf = out.parent / "traffic.json"
cur = json.loads(tf.read_text()) if tf.exists() else {}
cur[kind] = sum(traffic) / len(traffic)
tf.write_text(json.dumps(cur, indent=1) + "\n")
print("\n".join(lines))
