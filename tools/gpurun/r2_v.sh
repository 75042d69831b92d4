export GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so
GCNB_HEAD_TRACE=1 timeout 300 python tools/time_head.py 2>&1 | tail -4
