#!/bin/bash
# ncu full capture of the FP32-mode band 3 kernel (which unit saturates at 16 resident warps?)
mkdir -p gpurun_out
EXP_OPTS="precision=32" timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_kernel -s 2 -c 1 -o gpurun_out/prof_f32b3 python tools/dev/exp_bands.py > gpurun_out/ncu_f32.log 2>&1
ncu -i gpurun_out/prof_f32b3.ncu-rep --page details > gpurun_out/r01_v7_fp32_band3_ncu_details.txt 2>/dev/null
ncu -i gpurun_out/prof_f32b3.ncu-rep --page raw --csv > gpurun_out/r01_v7_fp32_band3_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_f32b3.ncu-rep --page source --csv > gpurun_out/prof_f32b3_source.csv 2>/dev/null
gzip -f gpurun_out/prof_f32b3_source.csv; rm -f gpurun_out/prof_f32b3.ncu-rep
grep -E "Duration|Throughput|Hit Rate|Issue Slots|Executed Ipc|Registers Per|Theoretical Active|Achieved Active|Eligible|No Eligible|Stall|stall" gpurun_out/r01_v7_fp32_band3_ncu_details.txt | head -40
