# round-2 re-entry check on a B200: full GPU suite, bench line (all configs), launch list, then A/B timing
# of launch-shape knobs on the final kernels (rollout kernel only, CUDA events, 20 reps)
set -x
mkdir -p gpurun_out
L=gpurun_out/exp16.log
(time timeout 800 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider) > gpurun_out/tests16_full.log 2>&1
tail -15 gpurun_out/tests16_full.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/b16.json 2> gpurun_out/b16.err; tail -c 1200 gpurun_out/b16.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches16.csv python bench.py --steps 2 --warmup 1 --only --no-cpu-baseline > gpurun_out/b16_ncu.log 2>&1
python scripts/launch_list.py gpurun_out/launches16.csv | head -30
: > $L
t() { echo "## $*" >> $L; env "$@" timeout 120 python scripts/prof_cfg.py $CFG 3 --time >> $L 2>&1; }
for CFG in 1 2; do
  t X=base
  t DIAL_SPLIT_SYNC=1
  t DIAL_SPLIT_SYNC=2
  t DIAL_SYNC_EVERY=2
  t DIAL_SYNC_EVERY=3
  t DIAL_PROF_NOISE=0
  t DIAL_B200_LIB=dial_mpc_b200/csrc/exp/libdial_b200_t448.so
  t DIAL_NO_LOCKSTEP=1
done
grep -v "^+" $L | cut -c1-250
