#!/bin/bash
# Where does the split-precision query kernel spend its time?  One process per setting (the library reads DUDF_TCX_* once).
out=${1:-gpurun_out/tcx_probe.txt}
: > $out
run() { env "$@" timeout 120 python tools/tcx_check.py grid >> $out 2>&1; }
run DUDF_TCX_SINCOS=poly
run DUDF_TCX_SINCOS=mufu
run DUDF_TCX_SINCOS=poly DUDF_TCX_CLUSTER=1
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=1
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=2
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=4
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=8
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=9
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=6
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=11
run DUDF_TCX_SINCOS=mufu DUDF_TCX_DBG=13
grep probe $out
