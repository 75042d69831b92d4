mkdir -p gpurun_out
export GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so
for w in f1; do
for d in 0 4; do
echo "== $w debug $d"
GCNB_UMMA_DEBUG=$d timeout 300 python tools/umma_trace.py $w 2>&1 | grep -v "^  sw" | tee -a gpurun_out/r2_trace5_$w.log
done
done
