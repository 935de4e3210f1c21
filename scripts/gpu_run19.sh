set -x
mkdir -p gpurun_out
timeout 200 scripts/probes/icache_probe > gpurun_out/icache_probe2.jsonl 2>&1; tail -34 gpurun_out/icache_probe2.jsonl
