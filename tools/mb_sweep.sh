#!/bin/bash
# microbench sweep over library variants under build/variants (scratch experiments)
run() { echo "== $1"; ENF_B200_LIB=$2 python tools/microbench.py --spec "$3" --D "$4" --N 20000000 --iters 10 2>&1 | tail -1 | cut -c1-140; }
for spec in "hh4,jo,cs:16" "cc,jo,hh4,ss:32" "cs:16" "hh4:16"; do
  s=${spec%%:*}; d=${spec#*:}
  run "main $s D=$d" $PWD/euclidiannormalizingflows.jl_b200/libenf_b200.so $s $d
  for v in 8b 8d 8e; do run "v$v $s D=$d" $PWD/build/variants/libenf_v$v.so $s $d; done
done
