# last GPU seconds of the round: the full GPU suite on the final build (after the tap-image commit)
mkdir -p gpurun_out
timeout 41 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2fin_tests.log 2>&1; echo rc=$?
tail -4 gpurun_out/r2fin_tests.log
