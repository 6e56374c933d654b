#!/bin/bash
# round-2 profile captures (one B200): launch list of one proof, ncu --set full of the top kernels, racecheck / memcheck of smoke().
# The .ncu-rep files are turned into text on the box and removed (gpurun brings back at most 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --inflight 1 --pool 1 > gpurun_out/r2_ncu_bench.log 2>&1; echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, count
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r2_$1 python scripts/dev_gp_grid.py > gpurun_out/r2_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
    ncu -i gpurun_out/r2_$1.ncu-rep --page raw > gpurun_out/r2_$1_raw.txt 2>&1
    ncu -i gpurun_out/r2_$1.ncu-rep --page source --csv > gpurun_out/r2_$1_source.csv 2>&1
    rm -f gpurun_out/r2_$1.ncu-rep
}
cap fused "k_hash_rw_up_r0|k_tree_up_r0" 26 2
cap fold "k_gp_fold_multi" 26 3
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2_racecheck.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_memcheck.log
du -sh gpurun_out
