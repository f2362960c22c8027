#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rebinned or sequence") > $O/r02h_pytest_gpu.log 2>&1
B="timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --particles 300000"
for rb in 0 1; do
  $B --poly-order 4 --rebin $rb > $O/r02h_vmec_k4_rebin$rb.json 2>> $O/r02h_err.log
  $B --poly-order 3 --rebin $rb > $O/r02h_vmec_k3_rebin$rb.json 2>> $O/r02h_err.log
done
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants --no-e2e --particles 1000000 --poly-order 4 --rebin 1 > $O/r02h_vmec_k4_rebin1_1M.json 2>> $O/r02h_err.log
tail -5 $O/r02h_pytest_gpu.log
