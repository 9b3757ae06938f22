set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -c 1500 gpurun_out/bench9.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err; cat gpurun_out/bench9_ref.json
python bench.py --workload config4 --files 296 --batch 148 --steps 2 --warmup 3 > gpurun_out/c4_batch2.json 2>/dev/null; cat gpurun_out/c4_batch2.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
python tests/gpu_ncu_batch.py 148 > gpurun_out/plain_batch.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'cl_batch_kernel|gw_kernel|k1_tile_write_batch' -s 3 -c 3 -o gpurun_out/r01d_batch -f python tests/gpu_ncu_batch.py 148 > gpurun_out/ncu_batch.log 2>&1
python tests/gpu_ncu_gw.py BIC > gpurun_out/plain_gwb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gw_kernel -s 1 -c 1 -o gpurun_out/r01d_gw_bic_split -f python tests/gpu_ncu_gw.py BIC > gpurun_out/ncu_gwb.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
