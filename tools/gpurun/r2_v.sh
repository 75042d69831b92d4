timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "head or fused_trainer" 2>&1 | tail -1
export GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so
GCNB_HEAD_TRACE=1 timeout 300 python tools/time_head.py 2>&1 | tail -10
