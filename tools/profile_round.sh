#!/bin/bash
# ncu evidence of a round (one GPU): launch lists with DRAM bytes for both smoothers, full-set captures of the
# dominant kernels.  Outputs in gpurun_out/; condensed into profiles/ with tools/ncu_launches.py afterwards.
mkdir -p gpurun_out
R='regex:^k_(st|rb3|jr3|fix|coarse_gemv|jacobi|residual|prolong|colour)'
for sm in jacobi rbgs; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k "$R" --csv --log-file gpurun_out/r2_launches_raw_$sm.csv python tools/gpu_probe.py --smoother $sm --cycles 2 \
      > gpurun_out/r2_launches_$sm.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rb3 -c 8 -o gpurun_out/r2_rb3_final \
    python tools/gpu_probe.py --smoother rbgs --cycles 1 > gpurun_out/r2_ncu_rb3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_(st3|jr3)' -c 4 -o gpurun_out/r2_st3_final \
    python tools/gpu_probe.py --cycles 1 > gpurun_out/r2_ncu_st3.log 2>&1
ls -la gpurun_out/r2_launches_raw_* gpurun_out/*final*
