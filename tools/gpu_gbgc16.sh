#!/bin/bash
mkdir -p gpurun_out/gbgc16
PA_LIB=$PWD/phyloacc_b200/libpa_old.so python tools/gbgc_host_ab.py /tmp/old.npz 2>&1 | tail -1 | tee -a gpurun_out/gbgc16/ab.txt
python tools/gbgc_host_ab.py /tmp/new.npz 2>&1 | tail -1 | tee -a gpurun_out/gbgc16/ab.txt
python - <<'PY' | tee -a gpurun_out/gbgc16/ab.txt
import numpy as np
a, b = np.load("/tmp/old.npz"), np.load("/tmp/new.npz")
for k in a.files:
    same = np.array_equal(a[k], b[k], equal_nan=True)
    print(k, "bit-identical" if same else "DIFFERENT: max abs diff %g" % np.nanmax(np.abs(a[k].astype(float) - b[k].astype(float))))
PY
