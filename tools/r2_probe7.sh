#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or ip" > $O/r2p7_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2p7_pytest.log
tail -5 $O/r2p7_pytest.log
timeout 500 python tools/gemm_stalls.py shard > $O/r2p7_stalls.log 2>&1; cat $O/r2p7_stalls.log
for o in "gemm_warm=8" "gemm_warm=0" "gemm_warm=4" "gemm_warm=16"; do
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py f16 1250000 384 1024 10 20
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py i8 1250000 384 1024 10 20
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py f16 10000000 384 1024 10 40
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py i8 10000000 384 1024 10 40
  CRS_OPTS=$o timeout 300 python tools/step_breakdown.py i8 12500000 384 16 100 20
done > $O/r2p7_ab.log 2>&1
cat $O/r2p7_ab.log
