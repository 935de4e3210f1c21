set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "randomize or fused_update or obs_and_done or golden" > gpurun_out/tests13.log 2>&1
tail -25 gpurun_out/tests13.log
python bench.py --only --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench13.json 2> gpurun_out/bench13.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench13.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])
PY
