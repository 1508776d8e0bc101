"""dev tool: per-iteration table (us, DRAM MB) from an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` launch list of
k_pass_a / k_pass_b_wl / k_solve.   python scripts/launch_table.py gpurun_out/r2_launches_32obj.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
d = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
    else:
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    d.setdefault(int(r[ii]), {"k": r[ki]})[r[mi]] = v
ids = sorted(d)
name = lambda k: "A" if "k_pass_a" in k else ("B" if "k_pass_b" in k else "S")
print("iter   A us  (rd MB, wr MB)     B us  (rd MB, wr MB)     S us    sum us")
it, cur = 0, {}
tot = {"A": 0.0, "B": 0.0, "S": 0.0}
for i in ids:
    n = name(d[i]["k"])
    cur[n] = d[i]
    tot[n] += d[i]["gpu__time_duration.sum"]
    if n == "S":
        g = lambda n, m: cur.get(n, {}).get(m, 0.0)
        print("%3d  %6.1f  (%6.1f, %5.1f)   %6.1f  (%6.1f, %5.1f)   %5.1f   %6.1f" % (
            it, g("A", "gpu__time_duration.sum"), g("A", "dram__bytes_read.sum"), g("A", "dram__bytes_write.sum"),
            g("B", "gpu__time_duration.sum"), g("B", "dram__bytes_read.sum"), g("B", "dram__bytes_write.sum"),
            g("S", "gpu__time_duration.sum"), sum(g(x, "gpu__time_duration.sum") for x in "ABS")))
        it += 1; cur = {}
s = sum(tot.values())
print("share of the trajectory: A %.1f %%  B %.1f %%  S %.1f %%   (%.0f us in %d iterations; ncu serialises the launches: shares, not absolutes)"
      % (100 * tot["A"] / s, 100 * tot["B"] / s, 100 * tot["S"] / s, s, it))
