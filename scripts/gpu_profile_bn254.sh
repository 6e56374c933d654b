#!/bin/bash
# ncu --set full of the BN254 grand-product fold kernels (rounds 1 and 2 of the second proof of a bench run), turned into text on the box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_gp_fold_multi" -s 14 -c 2 -o gpurun_out/r2bn_fold \
    python bench.py --field bn254 --steps 1 --warmup 3 --no-cpu-baseline --inflight 1 --pool 1 > gpurun_out/r2bn_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2bn_fold.ncu-rep --page raw > gpurun_out/r2bn_fold_raw.txt 2>&1
rm -f gpurun_out/r2bn_fold.ncu-rep
grep -c "k_gp_fold_multi" gpurun_out/r2bn_fold_raw.txt
