"""ncu CSV (one step of bench.py under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`)
-> profiles/r2_traffic.json: DRAM bytes and time per kernel, and the per-step total of the counting kernels that
bench.py reports as roofline.traffic.   python tools/ncu_traffic.py gpurun_out/x_traffic.csv profiles/r2_traffic.json"""
import csv, json, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iname, imetric, iunit, ival = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
iid = hdr.index("ID")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6}
per = defaultdict(lambda: defaultdict(float)); launches = defaultdict(set)
for r in rows[1:]:
    name = re.sub(r"<.*", "", r[iname].replace("void ", "").replace("mfkc::", ""))
    v = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1)
    per[name][r[imetric]] += v
    launches[name].add(r[iid])
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, one step of bench.py (cfg2, 1 GPU) after one warm-up step",
       "kernels": {}, "per_step_bytes": {}}
for name, m in per.items():
    out["kernels"][name] = {"launches": len(launches[name]), "dram_read_bytes": m.get("dram__bytes_read.sum", 0), "dram_write_bytes": m.get("dram__bytes_write.sum", 0),
                            "time_ms_under_ncu": m.get("gpu__time_duration.sum", 0)}
# the capture window need not coincide with one step: per-launch averages x the launches one cfg2 step makes (20 batches)
per_step = {"extract_skm_kernel": 20, "mark_read_ends_kernel": 20, "ovf_place_kernel": 1, "bin_count_kernel": 1, "drain_heavy_kernel": 0,
            "rs_hist_kernel": 8, "rs_chunk_kernel": 8, "rs_base_kernel": 8, "rs_offsets_kernel": 8, "rs_scatter_kernel": 8, "records_kernel": 1}
def short(k):
    return k.split("(")[0]
tot_c = tot_a = 0.0
counting = []
for k, v in out["kernels"].items():
    n = per_step.get(short(k), v["launches"])
    b = (v["dram_read_bytes"] + v["dram_write_bytes"]) / max(1, v["launches"]) * n
    v["launches_per_step"] = n; v["dram_bytes_per_step"] = b
    tot_a += b
    if short(k) in ("extract_skm_kernel", "bin_count_kernel", "drain_heavy_kernel", "ovf_place_kernel", "mark_read_ends_kernel"):
        counting.append(short(k)); tot_c += b
out["per_step_bytes"]["counting_kernels"] = counting
out["per_step_bytes"]["total_counting"] = tot_c
out["per_step_bytes"]["total_all"] = tot_a
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["per_step_bytes"]))
