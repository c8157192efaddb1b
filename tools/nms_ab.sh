set -x
cd /root/repo
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "nms or pack or record or detect or graph or stress or heads or network" 2>&1 | tail -15
timeout -s KILL 200 python tools/nms_time.py 2>&1 | tail -7
JDET_NMS_NO_GRID=1 timeout -s KILL 200 python tools/nms_time.py 2>&1 | tail -7
