"""dev tool: per-kernel counts of the SASS mnemonics that show what the hardware is asked to do (FP64 tensor-core MMA,
TMA bulk copies and their mbarriers, 256-bit loads, cluster / programmatic-launch control), from the built library.
    python scripts/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "visma_b200", "libvisma_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["DMMA", "UBLKCP", "SYNCS", "LDG.E.ENL2.256", "UCGABAR", "ACQBULK", "LDGDEPBAR", "BAR.SYNC", "FFMA", "DFMA", "SHFL", "ATOMS", "ATOMG", "RED", "CCTL"]
cur, counts, samples, total = None, collections.defaultdict(collections.Counter), collections.defaultdict(dict), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if not m or cur is None:
        continue
    ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
    total[cur] += 1
    for w in want:
        if ins.startswith(w) or (w == "SYNCS" and ins.startswith("SYNCS")):
            counts[cur][w] += 1
            samples[cur].setdefault(w, ins[:90])
print("# SASS evidence, libvisma_b200.so (cuobjdump -sass, sm_100a).  Columns: kernel, instructions, then mnemonic counts.")
print("# DMMA = FP64 tensor-core MMA (estimator Gram rows); UBLKCP = cp.async.bulk (TMA 1-D copies) with SYNCS mbarrier ops;")
print("# LDG.E.ENL2.256 = one 256-bit load per 32-byte scene record; UCGABAR = cluster barrier (k_solve's DSMEM reduce);")
print("# ACQBULK / LDGDEPBAR appear where griddepcontrol (programmatic dependent launch) is used.")
for k in sorted(total, key=lambda k: -total[k]):
    if not counts[k]:
        continue
    print("%-52s %6d  %s" % (k[:52], total[k], "  ".join("%s %d" % (w, c) for w, c in counts[k].most_common())))
print("\n# one sample line per (kernel, mnemonic) of interest")
for k in sorted(samples):
    for w in ("DMMA", "UBLKCP", "SYNCS", "LDG.E.ENL2.256", "UCGABAR"):
        if w in samples[k]:
            print("%-40s %s" % (k[:40], samples[k][w]))
