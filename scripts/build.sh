#!/bin/bash
# rebuild libfvla.so in-tree and show register/spill lines of one source's kernels: scripts/build.sh [source-stem]
cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'vla-from-fastvlm_b200')
from vla_fastvlm import _native
print(_native.build(verbose=False))" 2>&1 | grep -v "warning\|Remark\|^$\|\^" | tail -8
if [ -n "$1" ]; then grep "Used\|spill" vla-from-fastvlm_b200/csrc/build/$1.ptxas.log; fi
