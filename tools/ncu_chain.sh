#!/bin/bash
# One `ncu --set full` capture of the materialising Compton chain kernel at C2 size (default library or ALPK_LIB_PATH).
# usage: tools/ncu_chain.sh OUTNAME [kernel-regex]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=${2:-chain_rows_kernel}
ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 30 --launch-count 1 \
    -o gpurun_out/$1 -f python tools/time_chain.py > gpurun_out/$1.log 2>&1
tail -2 gpurun_out/$1.log
