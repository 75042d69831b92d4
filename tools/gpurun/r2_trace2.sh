mkdir -p gpurun_out
for d in 0 1; do
echo "== debug $d"
GCNB_UMMA_DEBUG=$d timeout 300 python tools/umma_trace.py f1 2>&1 | tail -26
done
echo "== f2"
timeout 300 python tools/umma_trace.py f2 2>&1 | tail -26
