#!/bin/bash
# Builds the matcher with cycle counters (-DMIMC3CU_PROFILE), prints the per-node phase breakdown, then
# restores the normal build.  Run on the GPU box: bash scripts/prof_run.sh [c1|c2s] [ocw list]
set -e
cd "$(dirname "$0")/.."
MIMC3CU_NVCC_EXTRA="-DMIMC3CU_PROFILE" python -c "from mimc3_b200 import build; build.build(force=True)"
python scripts/quick_time.py "${1:-c2s}" v2 "${2:-7,15,30,40}" 2>&1 | grep -E "prof|ocw"
python -c "from mimc3_b200 import build; build.build(force=True)"
