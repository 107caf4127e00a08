"""profiles/rNN_traffic.json from the ncu metrics pass of a round (the same run that gives the launch list):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -s <warm-up launches> -c <N> --csv --log-file gpurun_out/launches.csv python bench.py ...
    python tools/ncu_traffic.py gpurun_out/launches.csv c4 profiles/r02_traffic.json "r02 <what was captured>"

Launches are grouped by bench.py's kernel slots (row_forward_kernel, column_kernel, row_inverse_kernel,
row_inverse_forward_fused_kernel); traffic_bytes_per_launch = mean over the captured launches of a slot of
dram__bytes_read.sum + dram__bytes_write.sum, which bench.py reports as roofline.traffic for the dominant kernel.
"""
import csv
import json
import sys

SLOTS = (("row_inv_fwd_fused", "row_inverse_forward_fused_kernel"), ("row_fwd", "row_forward_kernel"),
         ("row_forward_kernel", "row_forward_kernel"), ("row_inv", "row_inverse_kernel"),
         ("row_inverse_kernel", "row_inverse_kernel"), ("col_", "column_kernel"), ("column_kernel", "column_kernel"))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0}


def slot_of(name):
    for key, slot in SLOTS:
        if key in name:
            return slot
    return None


def main(path, workload, out, source):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        if r[ix["ID"]] == "ID":
            continue
        rec = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
        val = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        rec[r[ix["Metric Name"]]] = val
    per = {}
    for _, rec in sorted(launches.items()):
        slot = slot_of(rec["name"])
        if slot is None:
            continue
        p = per.setdefault(slot, {"n": 0, "bytes": 0.0, "ms": 0.0, "fma": 0.0, "each": []})
        rd, wr = rec.get("dram__bytes_read.sum", 0.0), rec.get("dram__bytes_write.sum", 0.0)
        p["n"] += 1
        p["bytes"] += rd + wr
        p["ms"] += rec.get("gpu__time_duration.sum", 0.0)
        p["fma"] += rec.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        p["each"].append(f"{rd / 1e9:.2f}+{wr / 1e9:.2f} GB in {rec.get('gpu__time_duration.sum', 0.0):.3f} ms")
    try:
        doc = json.load(open(out))
    except Exception:
        doc = {}
    doc["comment"] = ("DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the ncu metrics pass of "
                      "this round's launch list; written by tools/ncu_traffic.py, read by bench.py for roofline.traffic")
    doc["source"] = source
    doc[workload] = {
        slot: {"traffic_bytes_per_launch": p["bytes"] / p["n"], "launches_captured": p["n"],
               "ncu_ms_per_launch": p["ms"] / p["n"], "fp32_pipe_busy_frac": (p["fma"] / p["n"] / 100.0) if p["fma"] else None,
               "launches": "; ".join(p["each"][:8])}
        for slot, p in per.items()}
    json.dump(doc, open(out, "w"), indent=1)
    for slot, p in per.items():
        print(f"{slot:36s} {p['n']:3d} launches  {p['bytes'] / p['n'] / 1e9:7.3f} GB/launch  {p['ms'] / p['n']:7.3f} ms  "
              f"fma pipe {p['fma'] / p['n']:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
