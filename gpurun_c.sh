timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TAG=base python tools/time_layers.py
TAG=fwd_maxw14 GCNB_FWD_MAXW=14 WHICH="f1 f2" python tools/time_layers.py
TAG=fwd_maxw20 GCNB_FWD_MAXW=20 WHICH="f1 f2" python tools/time_layers.py
TAG=bwd_S WHICH="b2" GCNB_BWD_S=1 python tools/time_layers.py
TAG=bwd_S2ws2 WHICH="b2" GCNB_BWD_S=2 GCNB_BWD_WS=2 python tools/time_layers.py
TAG=bwd_S2ws1 WHICH="b2" GCNB_BWD_S=2 GCNB_BWD_WS=1 python tools/time_layers.py
