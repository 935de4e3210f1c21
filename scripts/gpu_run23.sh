set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > gpurun_out/tests23_multi.log 2>&1; tail -4 gpurun_out/tests23_multi.log | cut -c1-300
python -c "import json; d=json.load(open('gpurun_out/parity/p2p_exchange.json')); print(d['graph_vs_eager'], d['graph_vs_eager_per_step'])"
