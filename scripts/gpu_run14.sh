set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "randomize or fused_update" > gpurun_out/tests14.log 2>&1
tail -25 gpurun_out/tests14.log | cut -c1-400
