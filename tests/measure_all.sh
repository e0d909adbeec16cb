#!/bin/bash
# One-call measurement plan for a fresh GPU box (run from the repo root, e.g.
#   gpurun --timeout 600 -- 'bash tests/measure_all.sh'
# ): everything lands under gpurun_out/.  Nothing here is a bench value except the bench.py lines.
set -u
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; ( timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?" >> "gpurun_out/$name.log" ); tail -2 "gpurun_out/$name.log"; }

# 1. correctness first: the full GPU suite (~30 s), then the staged N4 row (layer norm, CT_gan_64x64.py)
run 240 m1_gpu_suite python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8
CTGAN_STAGED=1 run 180 m2_staged_64x64 python -m pytest tests/test_staged_64x64_gpu.py -m gpu -q -p no:cacheprovider

# 2. the three bench lines (ResNet = BASELINE.json's metric)
( timeout 120 python bench.py > gpurun_out/m3_bench_resnet.json 2> gpurun_out/m3_bench_resnet.err )
( timeout 60 python bench.py --workload cifar > gpurun_out/m3_bench_cifar.json 2> gpurun_out/m3_bench_cifar.err )
( timeout 60 python bench.py --workload mnist > gpurun_out/m3_bench_mnist.json 2> gpurun_out/m3_bench_mnist.err )
cut -c1-200 gpurun_out/m3_bench_*.json

# 3. per-graph replay times and launch lists (ncu numbers are never bench values)
run 60 m4_graph_times python tests/graph_times.py 64
run 150 m5_ncu_resnet ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_resnet.csv python tests/profile_step.py 64
run 150 m6_ncu_cifar ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_cifar_dcgan.csv python tests/profile_dcgan.py cifar
python tests/summarize_launches.py gpurun_out/launches_resnet.csv 40 > gpurun_out/launches_resnet_by_kernel.txt 2>&1
python tests/summarize_launches.py gpurun_out/launches_cifar_dcgan.csv 40 > gpurun_out/launches_cifar_dcgan_by_kernel.txt 2>&1
head -12 gpurun_out/launches_resnet_by_kernel.txt
