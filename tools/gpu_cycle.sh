#!/bin/bash
# tools/gpu_cycle.sh TAG — one GPU round trip of the tuning loop: parity tests, a short bench line, one ncu full capture.
TAG=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_tests.log; cut -c1-200 gpurun_out/${TAG}_bench.json
if [ "$2" != "noprof" ]; then
ncu --set full --import-source on --clock-control none -k vmis_predict_kernel --launch-skip 5 -c 1 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
fi
