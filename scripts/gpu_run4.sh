set -x
mkdir -p gpurun_out
for lib in "" dial_mpc_b200/csrc/exp/libdial_b200_old.so dial_mpc_b200/csrc/exp/libdial_b200_noxch.so; do
  for c in 1 2 4 0; do DIAL_B200_LIB=$lib python scripts/prof_cfg.py $c 3 --time; done
done > gpurun_out/time4.log 2>&1
cat gpurun_out/time4.log
python -m pytest tests/test_gpu_at_size.py tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider > gpurun_out/tests4.log 2>&1
tail -12 gpurun_out/tests4.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for c in 1 2; do
  ncu --metrics $M --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 --csv --log-file gpurun_out/inst_cfg$c.csv python scripts/prof_cfg.py $c 2 > /dev/null 2>&1
  grep -E "inst_executed|time_duration" gpurun_out/inst_cfg$c.csv | awk -F'","' '{print $(NF-2), $NF}'
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench4_ref.json 2> gpurun_out/bench4_ref.err; tail -c 700 gpurun_out/bench4_ref.json
