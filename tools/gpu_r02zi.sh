#!/bin/bash
# Round 2, session 4: ncu --set full of every kernel of one CAE pass (592 utterances) in the final state
mkdir -p gpurun_out
timeout 200 python tools/prof_cae_small.py > gpurun_out/prof_cae_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/cae_launches.csv python tools/prof_cae_small.py > gpurun_out/ncu_cae_list.log 2>&1
echo "list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:cae_|conv_tc_kernel|xt_prep" -s 22 -c 11 -f -o gpurun_out/prof_cae_final python tools/prof_cae_small.py > gpurun_out/ncu_cae_final.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_cae_final.ncu-rep --page raw --csv > gpurun_out/prof_cae_final_raw.csv 2>/dev/null
python tools/launch_share.py gpurun_out/cae_launches.csv | head -20
