#!/bin/bash
# Times the Compton chain (tools/time_chain.py) for the default library and every tools/variants/libalpk_*.so.
# usage: tools/run_variants.sh [names...]   (default: all variants); ROWS="0 1 2" selects ALPK_CHAIN_ROWS modes per lib
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/variants.txt
: > $OUT
run() {   # lib-path-or-empty, rows-mode
    if [ -n "$1" ]; then export ALPK_LIB_PATH=$1; else unset ALPK_LIB_PATH; fi
    ALPK_CHAIN_ROWS=$2 python tools/time_chain.py 2>&1 | tail -1 | sed "s/^/rows=$2 /" >> $OUT
}
for m in ${ROWS:-0 1 2}; do run "" $m; done
if [ $# -gt 0 ]; then LIBS=$(for n in "$@"; do echo tools/variants/libalpk_$n.so; done); else LIBS=$(ls tools/variants/libalpk_*.so 2>/dev/null); fi
for l in $LIBS; do for m in ${VROWS:-1}; do run $PWD/$l $m; done; done
cat $OUT
