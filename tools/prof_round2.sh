set -x
ncu --set full --clock-control none --import-source on -k regex:besthit_kernel -c 1 -f -o gpurun_out/r2_k2_final python tools/as_trace.py --steps 1 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairhmm_band_kernel -c 4 -f -o gpurun_out/r2_k1b_final python tools/as_trace.py --steps 1 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairhmm_litband_kernel -c 1 -f -o gpurun_out/r2_litband_final python tools/as_trace.py --steps 1 > gpurun_out/r2_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:obs_wire_decode_kernel -c 1 -f -o gpurun_out/r2_wire_final python -m pytest tests/test_posterior_gpu.py -m gpu -q -k wire > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
