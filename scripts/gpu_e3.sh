#!/bin/bash
# A/B of the grouped grand-product tail kernel (HG_GP_TAIL_GROUPS / HG_GP_TAIL_LOG) on the Lasso node, after the GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/e3_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/e3_tests.log
for cfg in "HG_GP_TAIL_GROUPS=1" "HG_GP_TAIL_GROUPS=2" "HG_GP_TAIL_GROUPS=4" "HG_GP_TAIL_GROUPS=8" "HG_GP_TAIL_GROUPS=4 HG_GP_TAIL_LOG=7" "HG_GP_TAIL_GROUPS=8 HG_GP_TAIL_LOG=7" "HG_GP_TAIL_GROUPS=8 HG_GP_TAIL_LOG=8" "HG_GP_TAIL_GROUPS=13 HG_GP_TAIL_LOG=8" "HG_GP_TAIL_GROUPS=13 HG_GP_TAIL_LOG=9"; do
  env $cfg timeout 200 python scripts/dev_gp_grid.py 2>&1 | tail -1
done | tee gpurun_out/e3_grid.log
