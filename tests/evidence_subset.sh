set -x
P=gpurun_out/r2y
timeout 1500 python -m pytest tests -m gpu -q > ${P}_tests.log 2>&1; tail -3 ${P}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2>${P}_bench.err; cut -c1-300 ${P}_bench.json
timeout 900 python tests/parity_report.py > ${P}_parity_report.txt 2>${P}_parity.err; tail -2 ${P}_parity.err
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${P}_launches.csv python tests/profile_step.py 64 > /dev/null 2>&1; grep -c tc:: ${P}_launches.csv
ls -la gpurun_out | grep r2y
