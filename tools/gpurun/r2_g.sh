touch gcn_fmri_decoding_b200/csrc/head_fused.cu
GCNB_NVCC_EXTRA=-DGCNB_TRACE bash gcn_fmri_decoding_b200/csrc/build.sh > /dev/null 2>&1
GCNB_HEAD_TRACE=1 timeout 120 python tools/time_head.py
touch gcn_fmri_decoding_b200/csrc/head_fused.cu gcn_fmri_decoding_b200/csrc/cheb_fwd_umma.cu
bash gcn_fmri_decoding_b200/csrc/build.sh > /dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -x -q -k "dropout_matches or head_step or fused_trainer" 2>&1 | tail -3
