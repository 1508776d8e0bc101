#!/bin/bash
# One GPU-box call that produces the round's evidence (run through gpurun from the repo root):
#   tests, smoke, the bench line, the ncu launch list of the same command, render / KNN-sweep benches,
#   ncu captures of the new kernels and a compute-sanitizer pass.  Outputs: gpurun_out/<tag>_*.
tag=${1:-r1s3}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 400 $out/${tag}_bench_1gpu.json
python scripts/bench_render.py > $out/${tag}_render_bench.json 2> $out/${tag}_render.err; tail -c 500 $out/${tag}_render_bench.json
python scripts/bench_knn_sweep.py > $out/${tag}_knn_sweep.json 2> $out/${tag}_knn_sweep.err
python - <<PY
import json
for f in ("$out/${tag}_knn_sweep.json", "$out/${tag}_knn_sweep_seq.json"):
    try:
        for r in json.load(open(f))["rows"]:
            print(f[-18:], r["N"], "%.1f us" % (1e3 * r["query_ms_incl_sort"]), {k: "%.1f us %.0f GB/s" % (1e3 * v["ms"], v["GBps"]) for k, v in r["exhaustive"].items()}, r["spot_check_ok"])
    except Exception as e:
        print(f, "FAILED", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_ncu.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_tris" -s 1 -c 1 -o $out/${tag}_ktris python scripts/bench_render.py > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_bf_knn1|k_knn1_wpq" -c 4 -o $out/${tag}_knn python scripts/profile_bruteforce.py > /dev/null 2>&1
for tool in memcheck racecheck initcheck; do
  echo "== $tool"; timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_new_kernels.py 2>&1 | grep -v "^$" | tail -16
done > $out/${tag}_compute_sanitizer.txt 2>&1
grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out/${tag}_compute_sanitizer.txt
ls -la $out | grep ${tag}
