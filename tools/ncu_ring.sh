set -x
# clocks / power while the ring A/B runs at full size (one sample every 100 ms)
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 100 > gpurun_out/r2_ring_clocks.csv &
SMI=$!
timeout 280 python tools/ab_ring.py --cases cfg4:1000000,cfg4:500000,cfg4:250000 --reps 8 --tag clocks > gpurun_out/r2_ab_ring_sizes.json 2> gpurun_out/r2_ab_ring_sizes.err
kill $SMI
for S in 1000000 125000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:felsenstein_walk -s 2 -c 1 -f -o gpurun_out/r2_ring_cfg4_$S python bench.py --sites $S --steps 2 --warmup 1 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/r2_ring_ncu_$S.err
  ncu -i gpurun_out/r2_ring_cfg4_$S.ncu-rep --page raw --csv > gpurun_out/r2_ring_cfg4_${S}_raw.csv
  ncu -i gpurun_out/r2_ring_cfg4_$S.ncu-rep --page source --csv > gpurun_out/r2_ring_cfg4_${S}_source.csv
done
ls -la gpurun_out/*.ncu-rep
