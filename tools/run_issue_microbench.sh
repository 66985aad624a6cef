#!/bin/bash
# Runs the issue-slot microbenchmark on the GPU box and leaves the table under gpurun_out/ (copy into profiles/ to keep).
set -e
mkdir -p gpurun_out
SM_MHZ=$(nvidia-smi --query-gpu=clocks.max.sm --format=csv,noheader,nounits | head -1)
./tools/bin/issue_microbench "$SM_MHZ" > gpurun_out/issue_microbench.txt 2>&1
grep -v '^#JSON' gpurun_out/issue_microbench.txt | grep -v '^ *$'
