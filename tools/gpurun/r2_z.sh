mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "layer_stack or fused_trainer_with_dropout or golden_layer or input_pipeline or saved_basis or head_step or large_graph" > gpurun_out/r2z_synccheck.log 2>&1; echo "synccheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2z_synccheck.log | head -4
grep -o "[a-z_0-9]*\.cu[h]*:[0-9]*" gpurun_out/r2z_synccheck.log | sort | uniq -c | sort -rn | head -8
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
