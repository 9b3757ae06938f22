set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 1200 gpurun_out/bench10.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench10_ref.json 2> gpurun_out/bench10_ref.err; head -c 300 gpurun_out/bench10_ref.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
python tests/gpu_ncu_gw.py KL2 > gpurun_out/plain_gwk2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gw_kernel -s 1 -c 1 -o gpurun_out/r01e_gw_kl2 -f python tests/gpu_ncu_gw.py KL2 > gpurun_out/ncu_gwk2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
