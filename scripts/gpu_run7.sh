set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests7.log 2>&1
tail -15 gpurun_out/tests7.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -c 300 gpurun_out/bench7.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench7_ref.json 2> gpurun_out/bench7_ref.err
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum
for c in 0 1 2 3 4; do
  ncu --metrics $M --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 --csv --log-file gpurun_out/counts_cfg$c.csv python scripts/prof_cfg.py $c 2 > gpurun_out/counts_cfg$c.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --only --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
