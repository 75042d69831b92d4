mkdir -p gpurun_out
for d in 1 5 9 13 21 0 4 20; do
echo "== f1 debug $d"
GCNB_UMMA_DEBUG=$d timeout 300 python tools/umma_trace.py f1 2>&1 | grep -E "sparse0|sw 0|sw19"
done
