# staggered lock-step experiment (rollout kernel only, CUDA events, 20 reps) + launch list of the bench step
set -x
mkdir -p gpurun_out
L=gpurun_out/exp17.log
: > $L
t() { echo "## $*" >> $L; env "$@" timeout 120 python scripts/prof_cfg.py $CFG 3 --time >> $L 2>&1; }
for CFG in 1 2; do
  t X=base
  t DIAL_STAGGER=50
  t DIAL_STAGGER=150
  t DIAL_STAGGER=400
  t DIAL_STAGGER=1000
  t DIAL_STAGGER=100 DIAL_STAGGER_PHASES=4
  t DIAL_STAGGER=300 DIAL_STAGGER_PHASES=4
done
CFG=3
t X=base
t DIAL_STAGGER=300
grep -v "^+" $L | cut -c1-250
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rollout|update|trajbar|shift|weights|ybar|split|bars" -c 300 --csv --log-file gpurun_out/launches17.csv python bench.py --steps 2 --warmup 1 --only --no-cpu-baseline > gpurun_out/b17_ncu.log 2>&1
python scripts/launch_list.py gpurun_out/launches17.csv 2>/dev/null | head -20
