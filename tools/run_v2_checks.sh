#!/bin/bash
cd "$(dirname "$0")/.."
for k in 1 2 3 4 5 6; do
python - <<'PY'
import sys, json
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_gpu_reference_api as t
import re
body = open('tests/test_gpu_reference_api.py').read()
# run the trainer scenario and print the worst tensor
m = re.search(r'def test_reference_trainer_step.*?res = _run\(r"""(.*?)"""\)', body, re.S)
res = t._run(m.group(1))
print(round(res['worst'], 5), res['worst_key'], res['worst_ref'])
PY
done
