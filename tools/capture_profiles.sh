#!/bin/bash
# Run ON the GPU box (under gpurun): bench lines, the ncu launch list of the default bench
# command and `ncu --set full` captures of the top kernels, reduced to CSV (the .ncu-rep files
# stay on the box: gpurun_out/ is capped at 64 MiB).  Each profile follows a clean run of the
# same command without ncu.
set -u
out=gpurun_out
python bench.py --no-secondary --steps 5 > $out/r02_bench_c2_nosec.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/r02_launches_bench.csv python bench.py --no-secondary --steps 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_european_wide -c 1 -f \
    -o /tmp/heston python bench.py --no-secondary --steps 2 > /dev/null 2>&1
ncu -i /tmp/heston.ncu-rep --page raw --csv > $out/r02_ncu_heston_raw.csv
ncu -i /tmp/heston.ncu-rep --page source --csv > $out/r02_ncu_heston_source.csv 2>/dev/null
python bench.py --workload c4 --steps 5 > $out/r02_bench_c4.json 2>/dev/null || exit 1
ncu --set full --clock-control none -k regex:k_lsmc -c 2 -f -o /tmp/lsmc \
    python bench.py --workload c4 --steps 2 > /dev/null 2>&1
ncu -i /tmp/lsmc.ncu-rep --page raw --csv > $out/r02_ncu_lsmc_raw.csv
ls -la $out/r02_*
