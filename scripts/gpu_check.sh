#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list, ncu full capture of the two hot kernels.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh TAG [tests] [bench] [list] [full]'
TAG=${1:-x}; shift
WHAT="${*:-tests bench list full}"
O=gpurun_out; mkdir -p $O
python -c "import torch; print(torch.cuda.get_device_name(0))" > $O/dev_$TAG.log 2>&1
for w in $WHAT; do
case $w in
tests) timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_$TAG.log;;
bench) timeout 600 python bench.py > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"; cat $O/bench_$TAG.log; tail -3 $O/bench_$TAG.err;;
list) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1; echo "list rc=$?";;
full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:scaml_fit_kernel -s 3 -c 1 -f -o $O/fit_$TAG \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-posterior > $O/ncu_fit_$TAG.log 2>&1; echo "full-fit rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:scaml_predict_kernel -s 1 -c 1 -f -o $O/pred_$TAG \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_pred_$TAG.log 2>&1; echo "full-pred rc=$?";;
esac
done
