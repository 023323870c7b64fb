#!/bin/bash
# A/B of the mask-free epilogue loops on interior super-tiles (results bit-identical in every variant).
# Variants (build.build_variant(tag, [defines])):
#   default       gradient epilogue only (what ships)
#   nointerior    -DSCAML_FIT_NOINTERIOR                          masked loops only
#   int_both      -DSCAML_FIT_INTERIOR_ASM                        both epilogues
#   int_asmonly   -DSCAML_FIT_INTERIOR_ASM -DSCAML_FIT_NOINTERIOR assembly epilogue only
D=scalable-meta-learning-with-gaussian-processes_b200/csrc
O=gpurun_out/fit_interior_ab.txt; : > $O
for shape in "4096 6 256 6" "2048 2 512 10" "1776 2 384 6"; do
  echo "== $shape" >> $O
  for v in default nointerior int_both int_asmonly; do
    printf "%-14s " $v >> $O
    if [ $v = default ]; then timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
    elif [ -f $D/libscaml_b200_$v.so ]; then SCAML_LIB=$D/libscaml_b200_$v.so timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
    else echo "(variant not built)" >> $O; fi
  done
done
cat $O
