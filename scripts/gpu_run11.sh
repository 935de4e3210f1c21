set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests11.log 2>&1
tail -8 gpurun_out/tests11.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench11_ref.json 2> gpurun_out/bench11_ref.err
tail -c 900 gpurun_out/bench11_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke11.log 2>&1; tail -2 gpurun_out/smoke11.log
