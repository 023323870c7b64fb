D=scalable-meta-learning-with-gaussian-processes_b200/csrc
O=gpurun_out/fit_trismem_ab.txt; : > $O
for shape in "4096 6 256 6" "2048 2 512 10" "1776 2 384 6"; do
  echo "== $shape" >> $O
  for rep in 1 2; do
  for v in default trismem; do
    printf "%-10s " $v >> $O
    if [ $v = default ]; then timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O
    else SCAML_LIB=$D/libscaml_b200_$v.so timeout 200 python scripts/fit_bench.py $shape 2>&1 | tail -1 >> $O; fi
  done; done
done
cat $O
