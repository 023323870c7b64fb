#!/bin/bash
# Timing-only probe (WRONG results by design): what would a FOURTH co-resident CTA of the fit kernel buy?
# build: python -c "import sys; sys.path.insert(0,'scalable-meta-learning-with-gaussian-processes_b200'); import build; build.build_variant('probe4', ['-DSCAML_FIT_PROBE4'])"
# run:   gpurun --timeout 600 -- 'bash scripts/fit_probe4.sh'
# The probe build aliases the D^-1 tiles with the staging area (44 KB of shared memory per CTA), caps the kernel at 128
# registers and ignores pivot failures: instruction stream and memory traffic per evaluation as in the product kernel.
L=scalable-meta-learning-with-gaussian-processes_b200/csrc/libscaml_b200_probe4.so
O=gpurun_out/fit_probe4.txt; : > $O
for shape in "4096 6 256 6" "2048 2 512 10" "1776 2 384 6"; do
  echo "== $shape" >> $O
  echo -n "product kernel, 3 CTAs/SM, 168 regs:  " >> $O; timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
  echo -n "probe build, 3 CTAs/SM, 128 regs:     " >> $O; SCAML_LIB=$L SCAML_FIT_CTAS_PER_SM=3 timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
  echo -n "probe build, 4 CTAs/SM, 128 regs:     " >> $O; SCAML_LIB=$L timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
done
cat $O
