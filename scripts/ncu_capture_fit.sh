#!/bin/bash
# Re-capture of the fit kernel alone (after a change to scaml_fit.cuh): `ncu --set full` over scripts/ncu_driver.py restricted
# to scaml_fit_kernel (launch 0 = config 3, 1 = factorize mode, 2 = config 4 block), exported on the box as text.
#   gpurun --timeout 1200 -- 'bash scripts/ncu_capture_fit.sh TAG'
# then here:  python profiles/make_traffic_json.py <all-kernel raw csv of the round> gpurun_out/ncu_TAG/fit_raw.csv
TAG=${1:-x}; O=gpurun_out/ncu_$TAG; mkdir -p $O
timeout 300 python scripts/ncu_driver.py > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
SCAML_NCU_ONCE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:scaml_fit_kernel -f -o /tmp/fitk \
  python scripts/ncu_driver.py > $O/driver.log 2>&1
tail -2 $O/driver.log
ncu -i /tmp/fitk.ncu-rep --page raw --csv > $O/fit_raw.csv 2>/dev/null
for i in 0 2; do
  ncu -i /tmp/fitk.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > /tmp/src_$i.csv 2>/dev/null
  python profiles/ncu_lines.py /tmp/src_$i.csv 40 > $O/fit_${i}_hotspots.txt
done
ls -la $O
