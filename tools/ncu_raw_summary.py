"""Key metrics per kernel launch from an `ncu --page raw --csv` export.
Usage: python tools/ncu_raw_summary.py x_raw.csv [...] -> JSON on stdout"""
import csv, json, sys
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_inst",
    "lts__t_bytes.sum": "l2_bytes",
}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}
out = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for i, h in enumerate(hdr):
            if h in KEYS and r[i] not in ("", "n/a"):
                v = float(r[i].replace(",", ""))
                u = units[i]
                if KEYS[h].endswith("_bytes") or KEYS[h] == "duration_us":
                    v *= UNIT.get(u, 1)
                d[KEYS[h]] = round(v, 3)
        res.append(d)
    out[path.split("/")[-1]] = res
print(json.dumps(out, indent=1))
