cd /root/repo
for S in g5 g6 g7; do TDG_LIB=$PWD/tagdust_b200/libtagdust_b200_$S.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "label_parity or golden" 2>&1 | tail -1; done
bash scripts/gpu_ab.sh g4 g5 g6 g7 2>&1 | tail -10
