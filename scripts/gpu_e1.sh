#!/bin/bash
# A/B of the grand-product grid shaping (HG_GP_BALANCE / HG_GP_MIN_TPG) on the Lasso node, after the GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/e1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/e1_tests.log
for cfg in "HG_GP_BALANCE=0" "HG_GP_BALANCE=1 HG_GP_MIN_TPG=2" "HG_GP_BALANCE=1 HG_GP_MIN_TPG=4" "HG_GP_BALANCE=1 HG_GP_MIN_TPG=8"; do
  env $cfg timeout 200 python scripts/dev_gp_grid.py 2>&1 | tail -1
done | tee gpurun_out/e1_grid.log
