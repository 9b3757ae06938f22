set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/r02_ncu_plain_bench.json 2> gpurun_out/r02_ncu_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/r02_ncu_launches.log 2>&1
python profiles/drivers/cluster_time.py c3 > gpurun_out/r02_c3_plain.txt 2>&1 || exit 2
ncu --set full --clock-control none --import-source on -k 'regex:k5_direct|cl_fill_pairs|cl_small_loop|cl_self_logdet' -c 4 -f -o gpurun_out/r02_c3 python profiles/drivers/cluster_time.py c3 > gpurun_out/r02_c3_ncu.log 2>&1
python profiles/drivers/ncu_gw.py KL2 > gpurun_out/r02_kl2_plain.txt 2>&1 || exit 3
ncu --set full --clock-control none --import-source on -k regex:gw_kernel -s 1 -c 1 -f -o gpurun_out/r02_gw_kl2 python profiles/drivers/ncu_gw.py KL2 > gpurun_out/r02_gw_kl2_ncu.log 2>&1
python profiles/drivers/ncu_batch.py > gpurun_out/r02_batch_plain.txt 2>&1 || exit 4
ncu --set full --clock-control none --import-source on -k 'regex:gw_kernel|cl_batch_kernel' -c 2 -f -o gpurun_out/r02_batch python profiles/drivers/ncu_batch.py > gpurun_out/r02_batch_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r02_launches.csv > gpurun_out/r02_launches_bench.txt 2>&1
for n in c3 gw_kl2 batch; do python scripts/ncu_summary.py report gpurun_out/r02_$n.ncu-rep > gpurun_out/r02_${n}_ncu.txt 2>&1; done
# the binary reports exceed what gpurun brings back: the text summaries are what is kept
rm -f gpurun_out/r02_c3.ncu-rep gpurun_out/r02_gw_kl2.ncu-rep gpurun_out/r02_batch.ncu-rep
ls -la gpurun_out | tail -12
