for d in 0 1 2 3 4 7; do TAG="debug=$d" GCNB_FWD_DEBUG=$d python tools/time_fwd.py; done
TAG="nostage" GCNB_FWD_NOSTAGE=1 python tools/time_fwd.py
