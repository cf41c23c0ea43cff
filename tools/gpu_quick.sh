python -m pytest tests -m gpu -q -W ignore 2>&1 | tail -4
PA_PROF_ELEMS=592 python tools/profile_chain.py 2>&1 | tail -2
PA_PROF_LEN=cfg2 PA_PROF_ELEMS=2048 PA_PROF_ITERS=200 python tools/profile_chain.py 2>&1 | tail -2
