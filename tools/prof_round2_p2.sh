# ncu captures of K3/K4 on the four part-2 workloads of bench.py (one launch each, 65536 sites)
set -x
for w in c2_germline c2_continuous c5_trio c3_tumor_normal; do
  ncu --set full --clock-control none --import-source on -k regex:posterior_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_k3_$w python tools/p2_sweep.py $w --sites 65536 --steps 2 --check 0 > gpurun_out/r2_ncu_k3_$w.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_k3_$w.ncu-rep > gpurun_out/r2_ncu_k3_posterior_${w}_summary.txt 2>/dev/null
done
ls -la gpurun_out/r2_k3_*.ncu-rep gpurun_out/r2_ncu_k3_posterior_*_summary.txt
