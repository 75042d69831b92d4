touch gcn_fmri_decoding_b200/csrc/cheb_fwd_umma.cu
GCNB_NVCC_EXTRA=-DGCNB_TRACE bash gcn_fmri_decoding_b200/csrc/build.sh > /dev/null 2>&1
for d in 0; do echo debug=$d; GCNB_UMMA_DEBUG=$d timeout 120 python tools/umma_trace.py f1 | cut -c1-260; done
touch gcn_fmri_decoding_b200/csrc/cheb_fwd_umma.cu
