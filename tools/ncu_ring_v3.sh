S=125000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:felsenstein_walk -s 2 -c 1 -f -o gpurun_out/r2_ring_v3_cfg4_$S python bench.py --sites $S --steps 2 --warmup 1 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/r2_ring_v3_ncu_$S.err
ncu -i gpurun_out/r2_ring_v3_cfg4_$S.ncu-rep --page raw --csv > gpurun_out/r2_ring_v3_cfg4_${S}_raw.csv
ncu -i gpurun_out/r2_ring_v3_cfg4_$S.ncu-rep --page source --csv > gpurun_out/r2_ring_v3_cfg4_${S}_source.csv
