timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TAG=now WHICH="f1 f2" python tools/time_layers.py
TAG="debug=4" GCNB_FWD_DEBUG=4 WHICH="f1 f2" python tools/time_layers.py
TAG="debug=7" GCNB_FWD_DEBUG=7 WHICH="f1 f2" python tools/time_layers.py
