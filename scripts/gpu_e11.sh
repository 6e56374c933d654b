#!/bin/bash
# late sweep of the grid-shape knobs of the grand-product pipeline with the balanced term groups in place (Lasso node, dev_gp_grid.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "HG_GP_MAXBX=4" "HG_GP_MAXBX=2" "HG_GP_MAXBX=8" "HG_GP_TARGET=0.125" "HG_GP_TARGET=0.5" "HG_GP_TARGET=1.0" "HG_FUSED_CPS=8" "HG_FUSED_CPS=32" "HG_GP_MIN_TPG=6" "HG_GP_TAIL_GROUPS=13"; do
  env $cfg timeout 100 python scripts/dev_gp_grid.py 2>&1 | tail -1
done | tee gpurun_out/e11_grid.log
