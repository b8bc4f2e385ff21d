#!/bin/bash
# A/B of the split-precision chain kernel settings (one process per setting: the library reads DUDF_TCX_* once)
out=${1:-gpurun_out/tcx_ab.txt}
: > $out
run() { env "$@" timeout 200 python tools/tcx_check.py grid train >> $out 2>&1; }
run DUDF_TCX_DBG=0
run DUDF_TCX_DBG=64
run DUDF_TCX_DBG=0 DUDF_TCX_SINCOS=poly
run DUDF_TCX_DBG=64 DUDF_TCX_SINCOS=poly
grep "probe\|train tcx3" $out
