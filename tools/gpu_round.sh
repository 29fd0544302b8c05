#!/bin/bash
# One GPU-box visit: tests, smoke, bench, ncu launch list, ncu full capture of the main kernels.  Usage: tools/gpu_round.sh TAG [kernel-regex]
TAG=${1:-rX}
KRE=${2:-'regex:layer_kernel|stream_finish|stream_prepare|loader_gather'}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$KRE" --launch-skip 12 -c 8 -f -o gpurun_out/${TAG}_full python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
