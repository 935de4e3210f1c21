set -x
mkdir -p gpurun_out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 5 "$@"; }
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > gpurun_out/tests_multi15.log 2>&1
tail -3 gpurun_out/tests_multi15.log
run 8 > gpurun_out/s15_8.json 2> gpurun_out/s15_8.err
run 4 --only > gpurun_out/s15_4.json 2> gpurun_out/s15_4.err
run 2 --only > gpurun_out/s15_2.json 2> gpurun_out/s15_2.err
python bench.py --steps 20 --warmup 5 --only --no-cpu-baseline > gpurun_out/s15_1.json 2> gpurun_out/s15_1.err
python - <<'PY'
import json
for f in ("s15_1","s15_2","s15_4","s15_8"):
    try:
        txt=open("gpurun_out/%s.json"%f).read(); d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(f, "%.4e"%d["value"], round(d["ms_per_step"],4), "%.4e"%d["e2e"]["value"], d["phases_us_per_reverse_once"], d.get("exchange_wait_us_per_reverse_once"))
        for k,v in d.get("other_configs",{}).items(): print("   ",k,"%.4e"%v["value"], round(v["ms_per_step"],4), v["config"]["Nsample_total"], v.get("exchange_wait_us_per_reverse_once"))
    except Exception as e: print(f, "ERR", e)
PY
