#!/bin/bash
O=gpurun_out/${1:-r2e}
mkdir -p $O
echo "== ncu full mlpbwd_tc"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_mlpbwd_tc -s 2 -c 1 -f -o $O/mlpbwd_tc python tools/profile_step.py --steps 4 > $O/ncu_mlpbwd_tc.log 2>&1
tail -2 $O/ncu_mlpbwd_tc.log
python tools/ncu_mix.py $O/mlpbwd_tc.ncu-rep > $O/mlpbwd_tc_summary.txt 2>&1; cat $O/mlpbwd_tc_summary.txt
ncu -i $O/mlpbwd_tc.ncu-rep --page details 2>/dev/null | grep -E "Block Limit|Theoretical|Achieved Occ|Registers Per|Dynamic Shared|Bank|bank" | head -20
