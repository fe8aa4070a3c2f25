"""Summarise an `ncu --set full` report of the dominant GEMM: the per-launch metrics the bench line and DESIGN.md quote.

    python profiles/summarise_ncu.py gpurun_out/prof_rNN.ncu-rep profiles/ncu_rNN_summary.txt [profiles/ncu_dominant.json]
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).  The optional JSON (dram bytes of the launch
with the largest grid x duration, i.e. f = h W2^T) is what bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys

WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(rep, out_txt, out_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines, best = [f"# {rep}: ncu --set full --clock-control none (cold caches, serialised replays)"], None
    for k, r in enumerate(data):
        lines.append(f"## launch {k}: {r[idx['Kernel Name']]}")
        vals = {}
        for w in WANT:
            if w in idx:
                vals[w] = num(r[idx[w]])
                lines.append(f"{w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
        rd, wr = vals.get("dram__bytes_read.sum") or 0, vals.get("dram__bytes_write.sum") or 0
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
        rd *= scale.get(units[idx["dram__bytes_read.sum"]], 1.0)
        wr *= scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        sect = vals.get("lts__t_sectors_srcunit_tex_op_read.sum") or 0
        lines.append(f"{'derived: L2->SM read bytes (sectors x 32) / DRAM read bytes':70s} {32 * sect / max(rd, 1):16.2f}")
        if best is None or rd > best["dram_bytes_read"]:
            best = {"dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "lts_tex_read_bytes": int(32 * sect),
                    "duration_us_under_ncu": vals.get("gpu__time_duration.sum"),
                    "tensor_pipe_active_pct": vals.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "kernel": r[idx["Kernel Name"]], "source": out_txt}
    with open(out_txt, "w") as f:
        f.write("\n".join(lines) + "\n")
    if out_json:
        with open(out_json, "w") as f:
            json.dump(best, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:])
