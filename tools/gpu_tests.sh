#!/bin/bash
# smoke + the whole -m gpu suite on the current build
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log); tail -4 gpurun_out/r02_pytest.log | cut -c1-200
