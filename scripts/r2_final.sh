#!/bin/bash
# round 2, final evidence call (one box): full GPU tests, smoke, the bench line of both arms, the ncu launch list of the
# bench command, full captures of the shipped pass kernels (iterations 1 and 24), compute-sanitizer over the changed kernels.
tag=${1:-r2}; out=gpurun_out; mkdir -p $out
date -u +%T
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
date -u +%T
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 300 $out/${tag}_bench_1gpu.json; echo
date -u +%T
timeout 600 python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; tail -c 400 $out/${tag}_bench_reference_arm.json; echo
date -u +%T
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_ncu.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
N_ITER=30 STRIDE=1 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/${tag}_launches_32obj.csv python scripts/profile_traj.py > /dev/null 2>&1
N_ITER=30 STRIDE=8 timeout 300 ncu --metrics $M --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" --csv --log-file $out/${tag}_launches_4obj.csv python scripts/profile_traj.py > /dev/null 2>&1
N_ITER=3 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -s 3 -c 3 -o $out/${tag}_iter01_final python scripts/profile_traj.py > /dev/null 2>&1
N_ITER=26 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_pass_a|k_pass_b_wl|k_solve" -s 72 -c 3 -o $out/${tag}_iter24_final python scripts/profile_traj.py > /dev/null 2>&1
date -u +%T
for tool in memcheck racecheck; do
  echo "== $tool"; timeout 500 compute-sanitizer --tool $tool python scripts/sanitize_new_kernels.py 2>&1 | grep -v "^$" | tail -12
done > $out/${tag}_compute_sanitizer.txt 2>&1
grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out/${tag}_compute_sanitizer.txt
date -u +%T
ls -la $out | grep " ${tag}_"
