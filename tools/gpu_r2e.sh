mkdir -p gpurun_out/r2e
python -m pytest tests/test_gpu_outputs.py tests/test_gpu_posterior.py tests/test_gpu_chain.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2e/pytest.txt
cat gpurun_out/r2e/pytest.txt
# our binary's files for the post-processing fixture (tests/golden/post_consume)
python - <<'PY'
import os, shutil, subprocess, sys
sys.path.insert(0, os.getcwd())
import phyloacc_b200 as pab
from tests._util import Case
case = Case("ratite43")
r = subprocess.run([pab.BIN_PATH, case.cfg], capture_output=True, text=True)
dst = "gpurun_out/r2e/post_consume"
shutil.rmtree(dst, ignore_errors=True)
shutil.copytree(os.path.join(case.dir, "out"), dst)
open(os.path.join(dst, "stdout.log"), "w").write(r.stdout)
print(sorted(os.listdir(dst)))
PY
