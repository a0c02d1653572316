"""Turns ncu --set full reports (.ncu-rep, brought back in gpurun_out/) into the records committed under profiles/:
    python scripts/ncu_export.py <tag> <report.ncu-rep> [...]
writes profiles/<tag>_<report>.csv (one row per captured launch: every raw metric that matters for the roofline reading --
duration, DRAM / L2 bytes, tensor-pipe / issue / LSU activity, registers, shared memory, warp-stall breakdown) and merges
{kernel name -> dram__bytes_read.sum + dram__bytes_write.sum per launch} into profiles/r2_ncu_traffic.json (read by bench.py)."""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|"
                  r"lts__t_bytes\.sum$|lts__t_sectors_op_(red|atom)\.sum$|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"l1tex__lsu_writeback_active\.avg\.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts\.sum$|l1tex__data_pipe_lsu_wavefronts\.avg\.pct|"
                  r"launch__registers_per_thread$|launch__shared_mem_per_block_dynamic|launch__occupancy_limit|smsp__inst_executed\.sum$|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_.*|smsp__pcsamp_warps_issue_stalled_)")

def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u)

def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
        name = os.path.splitext(os.path.basename(rep))[0]
        out = os.path.join(ROOT, "profiles", f"{tag}_{name}.csv")
        with open(out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + [f"launch{k}" for k in range(len(data))])
            for i in cols:
                w.writerow([hdr[i], units[i]] + [r[i] for r in data])
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = {}
        for r in data:
            k = re.sub(r"[<(].*", "", r[ki].replace("void ", "")).strip()
            per.setdefault(k, []).append(float(r[ri]) * unit_scale(units[ri]) + float(r[wi]) * unit_scale(units[wi]))
        for k, v in per.items():
            key = k if name.endswith("map") or k not in traffic else f"{k}@{name}"
            traffic[key] = sum(v) / len(v)
        print(out, {k: round(sum(v) / len(v) / 1e6, 2) for k, v in per.items()}, "MB/launch")
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)

if __name__ == "__main__":
    main()
