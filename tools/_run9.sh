timeout 100 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k tiny-phi --tb=short 2>&1 | grep -v "^$" | tail -25
echo ---- legacy attention
LMRS_B200_ATT_CLUSTER=0 timeout 100 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k tiny-phi --tb=line 2>&1 | tail -4
