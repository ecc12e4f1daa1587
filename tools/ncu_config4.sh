#!/bin/bash
# ncu captures of the config-4 chain kernel (relay pass: 256-thread blocks; redo round: 1024-thread blocks) for profiles/.
cd "${GRAFT_REPO_ROOT:-.}"
ncu --set full --clock-control none --import-source on -k regex:stc007_chain_kernel -s 1 -c 3 -o gpurun_out/r2_config4_chain -f python tools/time_config4.py 200 > gpurun_out/r2_ncu_c4.log 2>&1
ncu -i gpurun_out/r2_config4_chain.ncu-rep --page raw --csv > gpurun_out/r2_config4_chain_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
