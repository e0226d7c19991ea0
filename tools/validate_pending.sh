#!/usr/bin/env bash
# First GPU call of the next round: run everything that was written after the round-1 GPU budget was spent.
#   1 GPU : SRG_TEST_UNVALIDATED=1 pytest (fast / two-order PPR normalisers), the all-ones shortcut of the host pipeline
#   2 GPUs: native NCCL handle at world size 2;  4 GPUs: copy / push_tma on the 2 x 2 grid
# usage: gpurun --gpus 2 -- 'bash tools/validate_pending.sh'
set -u
export SRG_TEST_UNVALIDATED=1
python -m pytest tests/test_fast_ppr.py tests/test_dist_native.py tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -15
for v in 0 1; do
  SRG_ONES_SHORTCUT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('SRG_ONES_SHORTCUT=$v e2e ms', d['e2e']['ms_per_step'], 'checksum', d['e2e']['checksum'])"
done
