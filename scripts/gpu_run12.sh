set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "obs_and_done or fused_update or env_step_sequence or device_loop or deploy" > gpurun_out/tests12.log 2>&1
tail -15 gpurun_out/tests12.log
python bench.py > gpurun_out/bench12.json 2> gpurun_out/bench12.err
tail -c 3000 gpurun_out/bench12.json
