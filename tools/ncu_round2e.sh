set -x
# ncu captures at the end of round 2 (last session): the kernels behind cfg2 / cfg3 / cfg5 after the launch-shape change
# (cfg3: two columns per thread), the two-level final reduction and the asynchronous staging of the small-tree kernel,
# and the column-per-thread kernel of a model-gradient evaluation (compile-time K = 4) on the cfg3 shape
for W in cfg2 cfg3 cfg5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:felsenstein_walk -s 3 -c 1 -f -o gpurun_out/r2e_walk_$W python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline > /dev/null 2> gpurun_out/r2e_ncu_$W.err
  ncu -i gpurun_out/r2e_walk_$W.ncu-rep --page raw --csv > gpurun_out/r2e_walk_${W}_raw.csv
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:felsenstein_walk_generic -s 1 -c 1 -f -o gpurun_out/r2e_model_gradient_cfg3 python tools/model_gradient_probe.py --cases cfg3:100000 --reps 1 > /dev/null 2> gpurun_out/r2e_ncu_mg.err
ncu -i gpurun_out/r2e_model_gradient_cfg3.ncu-rep --page raw --csv > gpurun_out/r2e_model_gradient_cfg3_raw.csv
ls -la gpurun_out/r2e_*.ncu-rep
