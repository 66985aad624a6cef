#!/bin/bash
# Round evidence on an N-GPU box (gpurun --gpus 8): bench at N = 2, 4, 8 (both arms at 8), the 2-rank GPU tests, the merge-step
# latency (peer windows vs NCCL), the per-stage scan profile.  Everything lands in gpurun_out/ev/.
cd "$(dirname "$0")/.."
O=gpurun_out/ev; mkdir -p $O
NG=${1:-8}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) "${@:2}"; }
timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_dist_gpu.py -q -m gpu > $O/pytest_gpu_multi.log 2>&1; tail -2 $O/pytest_gpu_multi.log
for n in 2 4 $NG; do
    [ $n -le $NG ] || continue
    run $n bench.py --gpus $n --steps 20 --warmup 5 2> $O/bench_n$n.err | grep "^{" > $O/bench_n$n.json
    python -c "
import json; d=json.load(open('$O/bench_n$n.json')); print($n, 'value', d['value'], 'e2e', d['e2e']['value'], 'strong c2', d['strong']['c2'].get('efficiency'), d['strong']['c2'].get('efficiency_graph'), 'c2x8', d['strong']['c2x8'].get('efficiency'), 'scan', d['scan']['seconds_per_map_pair'], d['scan_large']['seconds_per_map_pair'])"
done
run $NG bench.py --impl reference --gpus $NG --steps 20 --warmup 5 2> $O/bench_reference_arm_n$NG.err | grep "^{" > $O/bench_reference_arm_n$NG.json
for n in 2 $NG; do
    run $n tools/time_peer.py 2>/dev/null | grep "^{" > $O/peer_latency_n$n.json; cat $O/peer_latency_n$n.json
    run $n tools/profile_scan.py 2>/dev/null | grep "^{" > $O/scan_profile_n$n.json
done
ls -la $O
