set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_custom_env.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests24.log 2>&1; tail -12 gpurun_out/tests24.log | cut -c1-300
cat gpurun_out/parity/single_step_pincher.json
