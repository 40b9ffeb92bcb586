"""Summarise an .ncu-rep (one kernel launch) into a small JSON + markdown for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_fast_kernel
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or h == "Kernel Name":
            d[h] = {"value": v, "unit": u}
    json.dump(d, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary: {d.get('Kernel Name', {}).get('value', '?')}\n\nsource: `{rep}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in d:
                f.write(f"| {k} | {d[k]['value']} | {d[k]['unit']} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
