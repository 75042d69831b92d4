mkdir -p gpurun_out
touch gcn_fmri_decoding_b200/csrc/cheb_fwd_umma.cu
GCNB_NVCC_EXTRA=-DGCNB_TRACE bash gcn_fmri_decoding_b200/csrc/build.sh > /dev/null 2>&1
for d in 0 3; do echo debug=$d; GCNB_UMMA_DEBUG=$d timeout 120 python tools/umma_trace.py f1 | cut -c1-260 | grep -v item; done
echo conv2; timeout 120 python tools/umma_trace.py f2 | cut -c1-260 | grep -v item
touch gcn_fmri_decoding_b200/csrc/cheb_fwd_umma.cu
bash gcn_fmri_decoding_b200/csrc/build.sh > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_tests.log; echo "tests rc=$?"
tail -3 gpurun_out/r2d_tests.log
WHICH="f1 f2 s1" timeout 300 python tools/time_layers.py 2>&1 | tail -2 | tee gpurun_out/r2d_layers.log
