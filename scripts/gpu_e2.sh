#!/bin/bash
# A/B of the grand-product mid stages (HG_GP_MID_LOG / HG_GP_MID_TPG) on the Lasso node, after the GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/e2_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/e2_tests.log
for cfg in "HG_GP_MID_LOG=0" "HG_GP_MID_LOG=11" "HG_GP_MID_LOG=13" "HG_GP_MID_LOG=13 HG_GP_MID_TPG=4" "HG_GP_MID_LOG=13 HG_GP_MID_TPG=13" "HG_GP_MID_LOG=15" "HG_GP_MID_LOG=16 HG_GP_MID_TPG=13"; do
  env $cfg timeout 200 python scripts/dev_gp_grid.py 2>&1 | tail -1
done | tee gpurun_out/e2_grid.log
