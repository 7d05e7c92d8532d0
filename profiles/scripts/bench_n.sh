#!/bin/bash
# bench.py at N ranks of one box, as the driver launches it: bash profiles/scripts/bench_n.sh <N> <tag> [extra bench args]
cd "$(dirname "$0")/../.." || exit 1
N=$1; TAG=$2; shift 2
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 3 "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
fi
echo "rc=$?"; tail -3 gpurun_out/${TAG}.err; python profiles/summarize_bench.py gpurun_out/${TAG}.json 2>/dev/null | head -40
