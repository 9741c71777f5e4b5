set -x
# ncu captures of the kernels behind the other BASELINE configs (end of round 2)
for W in cfg2 cfg3 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:felsenstein_walk -s 3 -c 1 -f -o gpurun_out/r2d_walk_$W python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline > /dev/null 2> gpurun_out/r2d_ncu_$W.err
  ncu -i gpurun_out/r2d_walk_$W.ncu-rep --page raw --csv > gpurun_out/r2d_walk_${W}_raw.csv
done
ls -la gpurun_out/r2d_walk_cfg*.ncu-rep
