set -x
mkdir -p gpurun_out
timeout 120 scripts/probes/icache_probe > gpurun_out/icache_probe.jsonl 2>&1; cat gpurun_out/icache_probe.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "update_stage or golden or device_loop or fused_update" > gpurun_out/tests18.log 2>&1; tail -5 gpurun_out/tests18.log | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rollout|update|trajbar|shift|weights|ybar|split|bars" -c 300 --csv --log-file gpurun_out/launches18.csv python bench.py --steps 2 --warmup 1 --only --no-cpu-baseline > gpurun_out/b18_ncu.log 2>&1
python scripts/launch_list.py gpurun_out/launches18.csv 2>/dev/null | head -12
timeout 300 python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline > gpurun_out/b18.json 2> gpurun_out/b18.err; tail -c 2500 gpurun_out/b18.json
