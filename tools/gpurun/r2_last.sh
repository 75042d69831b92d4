mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2l_tests.log; tail -2 gpurun_out/r2l_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["gpu_launches_per_step"])
P
