set -x
cd /root/repo
timeout -s KILL 200 python tools/fr_time.py 2>&1 | tail -3 | head -1
for v in a16p2 a16p2b2 h3 b4 p2 h6; do
  JDET_B200_LIB=tools/_build/ab/lib$v.so JDET_B200_ALLOW_STALE_LIB=1 timeout -s KILL 200 python tools/fr_time.py 2>&1 | tail -3 | head -1
done
