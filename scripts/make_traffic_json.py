"""dev tool: profiles/r2_traffic.json (what bench.py's roofline.traffic reads) from the committed ncu captures' raw pages.
    python scripts/make_traffic_json.py gpurun_out/r2_iter24_shipped.ncu-rep gpurun_out/r2_iter01_shipped.ncu-rep"""
import csv, io, json, subprocess, sys
def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        def val(m):
            i = hdr.index(m)
            return float(d[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1)
        k = d[hdr.index("Kernel Name")]
        out.append(("k_pass_a" if "k_pass_a" in k else "k_pass_b_wl" if "k_pass_b_wl" in k else "k_solve",
                    {"dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
                     "gpu_time_us": round(val("gpu__time_duration.sum"), 2)}))
    return out
settled, search = launches(sys.argv[1]), launches(sys.argv[2])
t = {k: dict(v, source="profiles/r2_iter24_shipped_ncu_summary.txt (ncu --set full --clock-control none; iteration 24 of the BASELINE "
                        "point-to-plane trajectory, 32 objects, shipped kernels)") for k, v in settled}
t["search_bound_iteration_1"] = {k: v for k, v in search}
t["search_bound_iteration_1"]["source"] = "profiles/r2_iter01_shipped_ncu_summary.txt (iteration 1: every point is searched)"
json.dump(t, open("profiles/r2_traffic.json", "w"), indent=1)
print(json.dumps(t, indent=1))
