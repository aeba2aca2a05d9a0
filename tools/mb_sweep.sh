for v in 1 2; do ENF_B200_LIB=$PWD/build/variants/libenf_g$v.so python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what grad 2>&1 | tail -1 | cut -c1-110; done
python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what grad 2>&1 | tail -1 | cut -c1-110
