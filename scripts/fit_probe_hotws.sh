#!/bin/bash
# Timing-only probes of the fit kernel on the FINAL kernel (WRONG results by design; experiments only):
#  (1) -DSCAML_FIT_PROBE_HOTWS: all CTAs share 32 workspace slots (~20 MB, L2-resident) -- what does the DRAM spill of the
#      270 MB tile / kappa workspace cost?
#  (2) -DSCAML_ABLATE: phase ablations (scripts/fit_ablate.py) re-measured on the final kernel
L=scalable-meta-learning-with-gaussian-processes_b200/csrc/libscaml_b200_hotws.so
O=gpurun_out/fit_probe_hotws.txt; : > $O
for shape in "4096 6 256 6" "2048 2 512 10"; do
  echo "== $shape" >> $O
  echo -n "product kernel:                      " >> $O; timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
  echo -n "probe: 32 shared workspace slots:    " >> $O; SCAML_LIB=$L timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
done
echo "== ablations (config 3), bits as in scripts/fit_ablate.py / ABL() in scaml_fit.cuh" >> $O
timeout 400 python scripts/fit_ablate.py 0 8 16 24 33 1 32 6 2 4 128 256 512 1024 4096 8192 16384 32768 65536 63 255 1023 8191 131071 2>&1 | grep ablate >> $O
cat $O
