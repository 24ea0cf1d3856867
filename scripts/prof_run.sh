#!/bin/bash
# Per-node phase breakdown of the matcher (cycle counters, -DMIMC3CU_PROFILE).  The profiling variant of the
# library is built beforehand with `python -m mimc3_b200.build --variant prof -DMIMC3CU_PROFILE`.
# Run on the GPU box: bash scripts/prof_run.sh [c1|c2s|c4s] [ocw list]
cd "$(dirname "$0")/.."
MIMC3CU_LIB=$PWD/mimc3_b200/libmimc3cu_prof.so python scripts/quick_time.py "${1:-c2s}" v2 "${2:-7,15,30,40}" 2>&1 | grep -E "prof|ocw"
