# End-of-round evidence on one B200 (gpurun -- 'bash tests/evidence.sh'): GPU test suite, parity report, default bench line,
# DCGAN / 64x64 bench lines, ncu launch lists (step and batch sweep, with the tensor-pipe counter) and one `ncu --set full`
# capture of the top kernels.  Outputs land in gpurun_out/r2z_*; tests/ncu_by_kernel.py / ncu_summary.py turn them into profiles/.
set -x
P=gpurun_out/r2z
timeout 1500 python -m pytest tests -m gpu -q > ${P}_tests.log 2>&1; tail -3 ${P}_tests.log
timeout 900 python tests/parity_report.py > ${P}_parity_report.txt 2>${P}_parity.err; tail -2 ${P}_parity.err
timeout 600 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2>${P}_bench.err; cut -c1-300 ${P}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2>${P}_bench_reference.err; cut -c1-300 ${P}_bench_reference.json
timeout 600 python bench.py --workload 64x64 --steps 10 --warmup 3 --no-cpu-baseline > ${P}_bench_64x64.json 2>/dev/null; cut -c1-200 ${P}_bench_64x64.json
timeout 300 python tests/sweep_critic.py > ${P}_sweep_critic_step.jsonl 2>/dev/null; cat ${P}_sweep_critic_step.jsonl | cut -c1-200
for B in 64 512 4096; do timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${P}_sweep_launches_B$B.csv python tests/profile_step.py $B > /dev/null 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches.csv python tests/profile_step.py 64 > /dev/null 2>&1; grep -c tc:: ${P}_launches.csv
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${P}_launches_cifar_dcgan.csv python tests/profile_dcgan.py cifar > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o ${P}_top python tests/profile_kernels.py > ${P}_ncu.log 2>&1; tail -2 ${P}_ncu.log
ls -la gpurun_out | grep r2z
