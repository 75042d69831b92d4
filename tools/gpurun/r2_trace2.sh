mkdir -p gpurun_out
for d in 1 0; do
echo "== f1 debug $d"
GCNB_UMMA_DEBUG=$d timeout 300 python tools/umma_trace.py f1 2>&1 | grep -E "MMA|sw 0|sw19|sparse0"
done
