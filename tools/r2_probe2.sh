#!/bin/bash
# round-2 probe 2 (ONE GPU): floor sharing parity + A/B timing against the sample pass
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or ip or fp32 or tie" > $O/r2p2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2p2_pytest.log
tail -5 $O/r2p2_pytest.log
for o in "share_floor=1" "share_floor=0,sample_rows=65536" "share_floor=0" "share_floor=1,sample_rows=65536"; do
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py f16 1250000 384 1024 10 20
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py i8 1250000 384 1024 10 20
done > $O/r2p2_ab_shard.log 2>&1
for o in "share_floor=1" "share_floor=0,sample_rows=65536"; do
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py f16 10000000 384 1024 10 40
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py i8 10000000 384 1024 10 40
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py i8 12500000 384 16 100 20
done > $O/r2p2_ab_10m.log 2>&1
cat $O/r2p2_ab_shard.log $O/r2p2_ab_10m.log
