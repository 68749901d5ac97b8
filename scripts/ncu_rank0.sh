#!/bin/bash
# torchrun --no-python helper: rank 0 runs bench.py under ncu with a single-pass metric list (no
# kernel replay: the exchange kernel spins on its peers and cannot be replayed), the other ranks run
# it plainly.  Usage: NCU_METRICS=... NCU_OUT=... torchrun --no-python ... scripts/ncu_rank0.sh <bench args>
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics "$NCU_METRICS" --clock-control none -k regex:exchange_epilogue -c 24 --csv \
       --log-file "$NCU_OUT" python bench.py "$@"
else
  exec python bench.py "$@"
fi
