#!/bin/bash
# final tree: GPU tests + smoke
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-250 | tee gpurun_out/r2w_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
