#!/bin/bash
# tools/collect_profiles.sh TAG — on the GPU box: every capture profiles/README.md refers to, into gpurun_out/TAG_*.
# (one GPU; never under torchrun)
T=$1
B="python bench.py --no-cpu-baseline --sections none"
# launch list of the benchmark command: share of the step per kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $B --steps 3 --warmup 3 > gpurun_out/${T}_launch_bench.log 2>&1
# full capture, config 3 (launches 0-3 are the untimed stats passes, 4-6 warm-up)
ncu --set full --import-source on --clock-control none -k vmis_predict_kernel --launch-skip 5 -c 1 -f -o gpurun_out/${T}_full $B --steps 1 --warmup 3 > gpurun_out/${T}_full.log 2>&1
# config 2: 1 M interactions / 50 k items, launches of 1024 sessions
ncu --set full --import-source on --clock-control none -k vmis_predict_kernel --launch-skip 8 -c 1 -f -o gpurun_out/${T}_config2 $B --workload synthetic-1M-50k --batch 1024 --steps 4 --warmup 3 > gpurun_out/${T}_config2.log 2>&1
# config 4: 582 M interactions / 6.5 M items, device-built
ncu --set full --import-source on --clock-control none -k vmis_predict_kernel --launch-skip 5 -c 1 -f -o gpurun_out/${T}_config4 $B --workload synthetic-582M-6.5M --batch 524288 --steps 1 --warmup 3 > gpurun_out/${T}_config4.log 2>&1
