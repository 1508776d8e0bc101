"""dev tool: warp-instruction share per source REGION (named line ranges) for each profiled launch of an .ncu-rep.
usage: python scripts/ncu_regions.py rep.ncu-rep regions.txt   (regions: `file first last name` per line)"""
import csv, io, subprocess, sys
rep, regf = sys.argv[1], sys.argv[2]
regions = []
for l in open(regf):
    l = l.split("#")[0].split()
    if len(l) == 4:
        regions.append((l[0], int(l[1]), int(l[2]), l[3]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
launches, cur_file, cur = [], None, None
last_fn_file = None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) == 2 and r[0] == "Function Name":
        # a new launch starts when the file sequence restarts (first file seen again)
        if cur is None or (cur_file in cur["files"]):
            cur = {"files": set(), "rows": []}
            launches.append(cur)
        cur["files"].add(cur_file)
    elif len(r) > 8 and r[0].isdigit() and cur is not None:
        try:
            cur["rows"].append((cur_file, int(r[0]), int(r[7]), int(r[8])))
        except ValueError:
            pass
for i, L in enumerate(launches):
    tot = sum(x[2] for x in L["rows"]) or 1
    agg = {}
    for f, line, inst, tinst in L["rows"]:
        name = "other:" + f
        for rf, a, b, n in regions:
            if rf == f and a <= line <= b:
                name = n
                break
        w, t = agg.get(name, (0, 0))
        agg[name] = (w + inst, t + tinst)
    print("== launch %d: %d warp instructions" % (i, tot))
    for n, (w, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("   %5.1f%%  %12d  lanes %.1f  %s" % (100.0 * w / tot, w, t / max(w, 1), n))
