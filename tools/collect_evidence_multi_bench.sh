#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/ev; mkdir -p $O
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29700 + $1)) "${@:2}"; }
for n in 2 4 8; do
    run $n bench.py --gpus $n --steps 20 --warmup 5 2> $O/bench_n$n.err | grep "^{" > $O/bench_n$n.json
    python -c "
import json; d=json.load(open('$O/bench_n$n.json')); print($n, 'value', d['value'], 'e2e', d['e2e']['value'], 'strong c2', d['strong']['c2'].get('efficiency'), d['strong']['c2'].get('efficiency_graph'), 'c2x8', d['strong']['c2x8'].get('efficiency'), 'scan', d['scan']['seconds_per_map_pair'], d['scan_large']['seconds_per_map_pair'])"
done
run 8 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2> $O/bench_reference_arm_n8.err | grep "^{" > $O/bench_reference_arm_n8.json
run 8 tools/profile_scan.py 2>/dev/null | grep "^{" > $O/scan_profile_n8.json
