python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what grad 2>&1 | tail -1 | cut -c1-110
ENF_B200_LIB=$PWD/build/variants/libenf_g4.so python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what grad 2>&1 | tail -1 | cut -c1-110
python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what negll 2>&1 | tail -1 | cut -c1-110
python tools/microbench.py --spec hh4,jo,cs --what fwd_ladj 2>&1 | tail -1 | cut -c1-110
