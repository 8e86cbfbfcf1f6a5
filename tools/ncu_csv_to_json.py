"""`ncu -i x.ncu-rep --page raw --csv` -> compact JSON (one object per launch) with the metrics the
roofline discussion uses.  python tools/ncu_csv_to_json.py raw.csv out.json"""
import csv, json, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_fp64_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__cycles_elapsed.max",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    # the header is the first row whose first cell is "ID"; the next row holds the units
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[h], rows[h + 1]
    out = []
    for r in rows[h + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        o = {"Kernel Name": d.get("Kernel Name"), "Grid Size": d.get("Grid Size"), "Block Size": d.get("Block Size"), "units": {}}
        for k in KEEP:
            if k in d and d[k] != "":
                try:
                    o[k] = float(d[k].replace(",", ""))
                except ValueError:
                    o[k] = d[k]
                o["units"][k] = units[names.index(k)]
        out.append(o)
    json.dump(out, open(dst, "w"), indent=0)
    for o in out:
        t = o.get("gpu__time_duration.sum")
        rd, wr = o.get("dram__bytes_read.sum"), o.get("dram__bytes_write.sum")
        print(f'{str(o["Kernel Name"])[:60]:60s} {t} {o["units"].get("gpu__time_duration.sum")} rd {rd} wr {wr} {o["units"].get("dram__bytes_read.sum")}')


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
