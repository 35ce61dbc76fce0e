#!/bin/bash
# Run HERE before a gpurun call: product library + the two instrumented builds the GPU-side tools load
# (tools/timeline.py: -DMON_TIMELINE -> ro_map_b200/_build_tl/, tools/tc_phase_times.py: -DMON_TC_STAMPS -> ro_map_b200/_build_tl2/).
set -e
cd "$(dirname "$0")/.."
python -m ro_map_b200.build | tail -1
MON_EXTRA_NVCC_FLAGS=-DMON_TIMELINE python tools/timeline.py --build | tail -1
MON_EXTRA_NVCC_FLAGS=-DMON_TC_STAMPS python - <<'PY'
from pathlib import Path
from ro_map_b200 import build
d = Path("ro_map_b200/_build_tl2").resolve()
d.mkdir(exist_ok=True)
build.BUILD, build.LIB = d, d / "libmon_b200_stamps.so"
print(build.build(force=True))
PY
