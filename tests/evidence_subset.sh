# The ncu part of tests/evidence.sh alone (launch list of the eager ResNet steps + one `ncu --set full` capture of the top kernels):
#   gpurun -- 'bash tests/evidence_subset.sh'   ->  gpurun_out/r2w_*
set -x
P=gpurun_out/r2w
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file ${P}_launches.csv python tests/profile_step.py 64 > /dev/null 2>&1; grep -c tc:: ${P}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o ${P}_top python tests/profile_kernels.py > ${P}_ncu.log 2>&1; tail -2 ${P}_ncu.log
ncu -i ${P}_top.ncu-rep --page raw --csv > ${P}_top_raw.csv 2>/dev/null; wc -l ${P}_top_raw.csv
ls -la gpurun_out | grep r2w
