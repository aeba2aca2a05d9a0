python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 4; do
  echo "== static variant $v"; ENF_STATIC_VARIANT=$v python tools/microbench.py --spec hh4,jo,cs --what fwd_ladj 2>&1 | tail -1 | cut -c1-110
done
export ENF_NO_STATIC=1
for spec in hh4,jo,cs jo,cs cc,ji,hh4 ss; do python tools/microbench.py --spec $spec --what fwd_ladj 2>&1 | tail -1 | cut -c1-110; done
python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what fwd_ladj 2>&1 | tail -1 | cut -c1-110
python tools/microbench.py --spec hh4,jo,cs --what fwd 2>&1 | tail -1 | cut -c1-110
python tools/microbench.py --spec cc,jo,hh4,ss --D 32 --N 10000000 --what grad 2>&1 | tail -1 | cut -c1-110
