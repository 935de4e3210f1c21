set -x
mkdir -p gpurun_out
for c in 1 2 3 4; do python scripts/prof_cfg.py $c 3 --time; done > gpurun_out/time9.log 2>&1
for se in 2 3; do for c in 1 2; do DIAL_SYNC_EVERY=$se python scripts/prof_cfg.py $c 3 --time; done; done >> gpurun_out/time9.log 2>&1
grep cfg gpurun_out/time9.log
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests9.log 2>&1
tail -15 gpurun_out/tests9.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench9.json 2> gpurun_out/bench9.err
tail -c 300 gpurun_out/bench9.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err
DIAL_NO_FUSED_UPDATE=1 python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline > gpurun_out/bench9_nofused.json 2> /dev/null
ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 -o gpurun_out/r02_final_cfg1 python scripts/prof_cfg.py 1 2 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 -o gpurun_out/r02_final_cfg3 python scripts/prof_cfg.py 3 2 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 11 -c 1 -o gpurun_out/r02_final_cfg2 python scripts/prof_cfg.py 2 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
