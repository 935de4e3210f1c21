# A/B timing of launch-shape knobs on the final kernels (rollout kernel only, CUDA events, 20 reps)
set -x
mkdir -p gpurun_out
L=gpurun_out/exp16.log
: > $L
t() { echo "## $*" >> $L; env "$@" timeout 300 python scripts/prof_cfg.py $CFG 3 --time >> $L 2>&1; }
for CFG in 1 2; do
  t X=base
  t DIAL_SYNC_EVERY=2
  t DIAL_SYNC_EVERY=3
  t DIAL_SYNC_EVERY=6
  t DIAL_NO_LOCKSTEP=1
  t DIAL_PROF_NOISE=0
  t DIAL_PROF_NOISE=0 DIAL_NO_LOCKSTEP=1
  t DIAL_B200_LIB=dial_mpc_b200/csrc/exp/libdial_b200_t448.so
  t DIAL_WPC=7
  t DIAL_WPC=1
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "seq_jump_randomize or randomize_tasks_one or update_stage or golden or device_loop" > gpurun_out/tests16.log 2>&1
tail -25 gpurun_out/tests16.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline > gpurun_out/b16.json 2> gpurun_out/b16.err; tail -c 1500 gpurun_out/b16.json
CFG=3
t X=base
t DIAL_PROF_NOISE=0
t DIAL_B200_LIB=dial_mpc_b200/csrc/exp/libdial_b200_t448.so
grep -v "^+" $L | cut -c1-250
