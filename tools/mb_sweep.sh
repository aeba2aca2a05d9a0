#!/bin/bash
# microbench sweep over the in-tree library and any variant libraries under build/variants (scratch experiments:
# make -C euclidiannormalizingflows.jl_b200/csrc OUT=$PWD/build/variants/libenf_NAME.so OBJDIR=$PWD/build/obj_NAME EXTRA="-D...")
run() { echo "== $1"; ENF_B200_LIB=$2 python tools/microbench.py --spec "$3" --D "$4" --N 20000000 --iters 10 2>&1 | tail -1 | cut -c1-140; }
for spec in "hh4,jo,cs:16" "cc,jo,hh4,ss:32" "cs:16" "hh4:16"; do
  s=${spec%%:*}; d=${spec#*:}
  run "main $s D=$d" $PWD/euclidiannormalizingflows.jl_b200/libenf_b200.so $s $d
  for lib in build/variants/libenf_*.so; do [ -e "$lib" ] && run "$(basename $lib .so) $s D=$d" $PWD/$lib $s $d; done
done
