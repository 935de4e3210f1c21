# usage: [GPUS=2] scripts/gpurun_retry.sh <timeout> <command...> — retries while the pod answers "transient" (nothing charged)
T=$1; shift
G=${GPUS:+--gpus $GPUS}
for i in 1 2 3 4 5 6 7 8 9 10; do
  out=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -80
  echo "$out" | grep -q "status=transient\|status=busy" || exit 0
  sleep 120
done
