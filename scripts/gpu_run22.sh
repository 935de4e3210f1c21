# bisect of the 2-rank graph-vs-eager difference: session-start tree, HEAD, HEAD without the fused update kernel
set -x
mkdir -p gpurun_out
run() { ( cd $1 && shift && env "$@" timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551 tests/dist_p2p_check.py 2>&1 | grep P2PCHECK | cut -c1-900 ); }
echo "== old (59f6baf)"; run scratch/old X=1
echo "== HEAD"; run . X=1
echo "== HEAD, three-kernel update in the graph"; run . DIAL_NO_FUSED_UPDATE=1
echo "== HEAD, NCCL exchange off-graph"; run . DIAL_NO_LOCKSTEP=1
