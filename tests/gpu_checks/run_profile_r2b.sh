#!/bin/bash
# round-2 final evidence: ncu --set full of the forward stencil kernels (TMA-staged sdw_fwd_v6, tdw_fwd, se_pool) and the temporal
# backward at the six C2 block shapes (kbench --ncu: one launch each)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"sdw_fwd|tdw_|se_pool" -o gpurun_out/r2_fwd_final -f \
  python tests/gpu_checks/kbench.py sdw_fwd tdw_fwd tdw_bwd se_pool --ncu > gpurun_out/ncu_fwd_final.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_fwd_final.log
ncu -i gpurun_out/r2_fwd_final.ncu-rep --page raw --csv > gpurun_out/r2_fwd_final_raw.csv 2>/dev/null
python tests/gpu_checks/ncu_summary.py gpurun_out/r2_fwd_final_raw.csv gpurun_out/r2_ncu_full_fwd_final.csv \
  "ncu --set full --clock-control none on tests/gpu_checks/kbench.py sdw_fwd tdw_fwd tdw_bwd se_pool --ncu (C2 shapes, batch 32), final round-2 code"
python tests/gpu_checks/kbench.py > gpurun_out/kbench_r2_final.txt 2>&1; grep -c . gpurun_out/kbench_r2_final.txt
rm -f gpurun_out/*.ncu-rep
