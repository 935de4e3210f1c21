set -x
mkdir -p gpurun_out
L=gpurun_out/exp20.log
: > $L
t() { echo "## $*" >> $L; env "$@" timeout 120 python scripts/prof_cfg.py $CFG 3 --time >> $L 2>&1; }
for CFG in 1 2 0 4; do
  t X=base
  t DIAL_B200_LIB=dial_mpc_b200/csrc/exp/libdial_b200_peel.so
done
grep -v "^+" $L | cut -c1-250
