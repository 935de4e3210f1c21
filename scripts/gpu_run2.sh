set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > gpurun_out/tests_multi.log 2>&1
tail -15 gpurun_out/tests_multi.log
cat gpurun_out/parity/p2p_exchange.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --only > gpurun_out/bench2_p2p.json 2> gpurun_out/bench2_p2p.err
DIAL_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --only > gpurun_out/bench2_nccl.json 2> gpurun_out/bench2_nccl.err
tail -c 400 gpurun_out/bench2_p2p.err
python -c "
import json
for f in ('p2p','nccl'):
    d=json.load(open('gpurun_out/bench2_%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['phases_us_per_reverse_once'], d['config']['parallelism'])
"
