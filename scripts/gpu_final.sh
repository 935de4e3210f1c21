# round-end check on one B200: full GPU suite, the bench line (all configs + CPU arm), the reference arm,
# the launch list of the bench step and one full ncu capture of the headline rollout kernel
set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider) > gpurun_out/final_tests.log 2>&1
tail -6 gpurun_out/final_tests.log | cut -c1-300
M=$(python -c "import sys; sys.path.insert(0,'scripts'); import rollout_counts as r; print(r.METRICS)")
for i in 0 1 2 3 4; do timeout 200 ncu --metrics $M --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 --csv --log-file gpurun_out/counts_cfg$i.csv python scripts/prof_cfg.py $i 3 > /dev/null 2>&1; done
python scripts/rollout_counts.py gpurun_out/counts_cfg0.csv gpurun_out/counts_cfg1.csv gpurun_out/counts_cfg2.csv gpurun_out/counts_cfg3.csv gpurun_out/counts_cfg4.csv; cp profiles/rollout_counts.json gpurun_out/rollout_counts.json
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; tail -c 600 gpurun_out/final_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rollout|update|trajbar|shift|weights|ybar|split|bars" -c 300 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --only --no-cpu-baseline > gpurun_out/final_ncu_list.log 2>&1
python scripts/launch_list.py gpurun_out/final_launches.csv 2>/dev/null | head -14
timeout 300 ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 -f -o gpurun_out/r02_final2_cfg1 python scripts/prof_cfg.py 1 3 > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
