# end-of-round-2 evidence: tests, bench lines (own arm + reference arm), smoke(), launch list; with
# "ncu" as first argument also the captures of the kernels changed last (K2 with the word-restricted
# stores, the anti-diagonal kernel)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; tail -2 gpurun_out/r2f_tests.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err || exit 1
python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2f_ncu_bench.log 2>&1
if [ "$1" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:besthit_kernel -c 1 -f -o gpurun_out/r2f_k2 python tools/as_trace.py --steps 1 > gpurun_out/r2f_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairhmm_diag_kernel -c 2 -f -o gpurun_out/r2f_diag python -m pytest tests/test_pairhmm_gpu.py -m gpu -q -k "long_banded" > gpurun_out/r2f_ncu_b.log 2>&1
fi
ls -la gpurun_out/r2f_*
