#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > $O/r02g_pytest_gpu.log 2>&1
tail -5 $O/r02g_pytest_gpu.log
