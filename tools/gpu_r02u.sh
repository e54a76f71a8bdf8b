#!/bin/bash
mkdir -p gpurun_out
EER_N=0 timeout 120 python tools/prof_all_small.py > gpurun_out/prof_plain.log 2>&1 &&
EER_N=0 timeout 600 ncu --metrics lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_prefetch.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:"cnn1d_fused_kernel" -s 1 -c 1 --csv --log-file gpurun_out/c1d_l2.csv python tools/prof_all_small.py > gpurun_out/ncu_c1d.log 2>&1
echo "ncu exit $?"
cat gpurun_out/c1d_l2.csv | tail -12
