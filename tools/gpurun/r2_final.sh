# round-2 final captures: GPU tests, smoke, bench line of every config + the reference arm, launch list, ncu --set full
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02f_tests.log; tail -2 gpurun_out/r02f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for c in 2 3a 3a1 3b 1 5; do
  timeout 900 python bench.py --config $c > gpurun_out/r02f_bench_cfg$c.json 2> gpurun_out/r02f_bench_cfg$c.err; echo "config $c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02f_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 27 -c 18 --csv --log-file gpurun_out/r02f_launches.csv python tools/prof_step.py 5 > gpurun_out/r02f_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -s 27 -c 9 -o gpurun_out/r02f_step -f python tools/prof_step.py 5 > gpurun_out/r02f_step_ncu.log 2>&1; echo "step ncu rc=$?"
python - <<'P'
import json
for c in ("2","3a","3a1","3b","1","5"):
    try:
        d=json.loads(open("gpurun_out/r02f_bench_cfg%s.json"%c).read().strip().splitlines()[-1])
        print(c, round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(c, "fail", e)
P
