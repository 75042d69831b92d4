# 2 GPUs: DP tests (peer-memory + NCCL), bench at N=2 and N=1, traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2u_tests.log; tail -1 gpurun_out/r2u_tests.log
for w in f1 f2; do
GCNB_LIB_PATH=$PWD/gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so timeout 300 python tools/umma_trace.py $w 2>&1 | grep "prologue\|sparse0" 
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2u_n1.json 2> gpurun_out/r2u_n1.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2u_n2.json 2> gpurun_out/r2u_n2.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --strong > gpurun_out/r2u_n2s.json 2> gpurun_out/r2u_n2s.err; echo rc=$?
python - <<'P'
import json
for n in ("n1","n2","n2s"):
    try:
        d=json.loads(open("gpurun_out/r2u_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["config"].get("allreduce"), d["config"]["global_batch"], [(k["op"][:9], round(k["us"],1)) for k in d["roofline"]["kernels"]])
    except Exception as e:
        print(n, "fail", e); print(open("gpurun_out/r2u_%s.err"%n).read()[-1500:])
P
