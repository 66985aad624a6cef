#!/bin/bash
# usage: tools/build_variant.sh NAME "-DALPK_X=1 ..."   -> tools/variants/libalpk_NAME.so (kernel-variant experiments)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
ALPK_LIB_PATH=$PWD/tools/variants/libalpk_$NAME.so ALPK_NVCC_EXTRA="$*" python -c "
from alplib_b200 import _native
print(_native.build(force=True))"
