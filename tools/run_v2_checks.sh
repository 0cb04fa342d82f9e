#!/bin/bash
cd "$(dirname "$0")/.."
for k in 1 2 3; do timeout 900 python -m pytest "tests/test_gpu_train_step.py::test_graphed_step_equals_eager_step" -q 2>&1 | grep -E "^E |passed|failed" | head -8; done
