timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -8
timeout 120 python tools/attn_probe.py llama-3.2-1b 1 64,576,1000,1500,2000,2100
export QUICK_CONFIGS='[{}, {"LMRS_B200_ATT_GROUPS": "1"}]'
timeout 300 python tools/quick_decode.py llama-3.2-1b 1 512 64
LMRS_B200_SO=lm.rs_b200/lmrs_b200/liblmrs_b200_trace.so timeout 300 python tools/timeline.py llama-3.2-1b 1 512 2>&1 | tail -8
