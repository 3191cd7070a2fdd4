#!/bin/bash
# First-contact GPU run: smoke, every GPU test function in its own process (a faulted CUDA context
# then costs one test, not the suite), then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -3 gpurun_out/smoke.log
: > gpurun_out/pytest.log
for t in $(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep '::' | sed 's/\[.*//' | sort -u); do
  timeout 400 python -m pytest "$t" -q -x --timeout 300 -p no:cacheprovider >> gpurun_out/pytest.log 2>&1
  echo "$t exit=$?"
done
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -40
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -5 gpurun_out/bench.log
