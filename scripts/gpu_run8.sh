set -x
mkdir -p gpurun_out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 "$@"; }
run 8 > gpurun_out/scale8.json 2> gpurun_out/scale8.err
DIAL_EXCHANGE=nccl run 8 --only > gpurun_out/scale8_nccl.json 2> gpurun_out/scale8_nccl.err
run 4 --only > gpurun_out/scale4.json 2> gpurun_out/scale4.err
run 2 --only > gpurun_out/scale2.json 2> gpurun_out/scale2.err
python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline > gpurun_out/scale1.json 2> gpurun_out/scale1.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > gpurun_out/tests_multi8.log 2>&1
tail -3 gpurun_out/tests_multi8.log
python - <<'PY'
import json
for f in ("scale1","scale2","scale4","scale8","scale8_nccl"):
    try:
        txt=open("gpurun_out/%s.json"%f).read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(f, "%.4e"%d["value"], round(d["ms_per_step"],4), "%.4e"%d["e2e"]["value"], d["phases_us_per_reverse_once"], d["config"]["parallelism"][:50])
        for k,v in d.get("other_configs",{}).items(): print("   ",k,"%.4e"%v["value"], round(v["ms_per_step"],4), v["config"]["Nsample_total"])
    except Exception as e: print(f, "ERR", e)
PY
