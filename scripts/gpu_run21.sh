# 2-GPU sanity of the sharded path (peer-memory exchange, update kernel) after the round-2 kernel revisions
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > gpurun_out/tests21_multi.log 2>&1; tail -4 gpurun_out/tests21_multi.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 --only > gpurun_out/s21_2.json 2> gpurun_out/s21_2.err; tail -c 1500 gpurun_out/s21_2.json
timeout 200 python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline > gpurun_out/s21_1.json 2> gpurun_out/s21_1.err; tail -c 300 gpurun_out/s21_1.json
