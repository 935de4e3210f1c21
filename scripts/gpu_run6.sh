set -x
mkdir -p gpurun_out
for c in 1 2 4 0; do python scripts/prof_cfg.py $c 3 --time; done > gpurun_out/time6.log 2>&1
for c in 1 2; do DIAL_NO_MIDSYNC=1 python scripts/prof_cfg.py $c 3 --time; done >> gpurun_out/time6.log 2>&1
for c in 1 2; do DIAL_WPC=7 python scripts/prof_cfg.py $c 3 --time; done >> gpurun_out/time6.log 2>&1
grep cfg gpurun_out/time6.log
python -m pytest tests/test_gpu_at_size.py tests/test_gpu_parity.py tests/test_custom_env.py -q --tb=line -p no:cacheprovider -k "not baseline_size-3" > gpurun_out/tests6.log 2>&1
tail -6 gpurun_out/tests6.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for c in 1 2; do
  ncu --metrics $M --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 --csv --log-file gpurun_out/inst_cfg$c.csv python scripts/prof_cfg.py $c 2 > /dev/null 2>&1
  grep -E "inst_executed|time_duration" gpurun_out/inst_cfg$c.csv | awk -F'","' '{print $(NF-2), $NF}'
done
