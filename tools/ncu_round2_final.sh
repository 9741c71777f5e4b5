set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/ab_ring.py --cases cfg4:1000000,cfg4:125000,cfg3:100000,cfg5:2000000,cfg2:4000000,cfg1:4000000 --tag final > gpurun_out/r2d_ab_ring_final.json 2> gpurun_out/r2d_ab_ring_final.err
# launch list of the bench command (per-launch times under ncu are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_cfg4_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2d_bench_under_ncu.json 2> gpurun_out/r2d_ncu_list.err
for S in 1000000 500000 250000 125000; do
  ncu --set full --clock-control none --import-source on -k regex:felsenstein_walk -s 2 -c 1 -f -o gpurun_out/r2d_walk_cfg4_$S python bench.py --sites $S --steps 2 --warmup 1 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/r2d_ncu_$S.err
  ncu -i gpurun_out/r2d_walk_cfg4_$S.ncu-rep --page raw --csv > gpurun_out/r2d_walk_cfg4_${S}_raw.csv
done
ncu -i gpurun_out/r2d_walk_cfg4_125000.ncu-rep --page source --csv > gpurun_out/r2d_walk_cfg4_125000_source.csv
python bench.py > gpurun_out/r2d_bench_cfg4_n1.json 2> gpurun_out/r2d_bench_cfg4_n1.err
tail -c 600 gpurun_out/r2d_bench_cfg4_n1.json
