timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TAG=stagger WHICH="f1 f2" python tools/time_layers.py
for d in 1 2 4; do TAG="debug=$d" GCNB_FWD_DEBUG=$d WHICH="f1 f2" python tools/time_layers.py; done
