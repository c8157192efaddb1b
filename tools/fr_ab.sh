set -x
cd /root/repo
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "feature_refine or fr_" 2>&1 | tail -5
timeout -s KILL 200 python tools/fr_time.py 2>&1 | tail -4
for v in h10 h9 p1h12 c8p2; do
  JDET_B200_LIB=tools/_build/ab/lib$v.so JDET_B200_ALLOW_STALE_LIB=1 timeout -s KILL 200 python tools/fr_time.py 2>&1 | tail -3 | head -2
done
