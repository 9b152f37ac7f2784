#!/bin/bash
# Builds the native host of the time-loop seam against the CUDA product library:
#   axisem_b200/axisem_b200_solver  (rpath $ORIGIN -> libaxisem_b200.so next to it)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
CXX="${CXX:-g++}"
OUT="${OUT:-$HERE/../axisem_b200_solver}"
"$CXX" -O2 -std=c++17 -Wall -Wextra -I"$ROOT/include" -o "$OUT" \
    "$HERE/main.cpp" "$HERE/time_loop.cpp" "$HERE/modules.cpp" "$HERE/meshdb.cpp" "$HERE/receivers.cpp" "$HERE/rundir.cpp" \
    "$HERE/precomp.cpp" "$HERE/mapping.cpp" "$HERE/background_models.cpp" \
    -L"$HERE/.." -laxisem_b200 -Wl,-rpath,'$ORIGIN'
echo "built $OUT"
# mesher database -> module variables (no device code)
"$CXX" -O2 -std=c++17 -Wall -Wextra -I"$ROOT/include" -o "$HERE/../axisem_b200_meshdb2axbp" \
    "$HERE/meshdb2axbp.cpp" "$HERE/meshdb.cpp" "$HERE/modules.cpp" "$HERE/spectral.cpp" \
    "$HERE/precomp.cpp" "$HERE/mapping.cpp" "$HERE/background_models.cpp"
echo "built $HERE/../axisem_b200_meshdb2axbp"
# spectral basis and background models (the first pieces of the native pre-processing)
"$CXX" -O2 -std=c++17 -Wall -Wextra -o "$HERE/../axisem_b200_hosttool" \
    "$HERE/hosttool.cpp" "$HERE/spectral.cpp" "$HERE/background_models.cpp"
echo "built $HERE/../axisem_b200_hosttool"
# post-processing of the solver output (radiation factors, rotation, STF convolution)
"$CXX" -O2 -std=c++17 -Wall -Wextra -o "$HERE/../axisem_b200_postproc" "$HERE/postproc_main.cpp" "$HERE/postprocess.cpp" "$HERE/rundir.cpp"
echo "built $HERE/../axisem_b200_postproc"
# the native pre-computation as a stand-alone step: MESHER databases -> complete containers
"$CXX" -O2 -std=c++17 -Wall -Wextra -o "$HERE/../axisem_b200_precomp" "$HERE/precomp_main.cpp" "$HERE/precomp.cpp" "$HERE/receivers.cpp" \
    "$HERE/mapping.cpp" "$HERE/background_models.cpp" "$HERE/meshdb.cpp" "$HERE/modules.cpp"
echo "built $HERE/../axisem_b200_precomp"
