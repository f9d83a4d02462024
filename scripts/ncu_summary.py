"""Markdown summary of `ncu --set full` reports: duration, DRAM traffic, tensor-pipe / memory utilisation."""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__inst_executed_pipe_tc.sum", "tensor-core instructions"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor pipe active cycles (per TPC)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts (LSU)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smem wavefronts (tensor core)"),
]

def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"## {path}: no data"); continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            print(f"## {d.get('Kernel Name', ('?',))[0][:110]}\n\nsource: `{path}` (ncu --set full --clock-control none)\n")
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                hit = [h for h in d if h.endswith(k) or h == k]
                if hit:
                    v, u = d[hit[0]]
                    print(f"| {label} (`{k}`) | {v} {u} |")
            print()

if __name__ == "__main__":
    main(sys.argv[1:])
