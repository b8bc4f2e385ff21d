# pipeline probe of the tcgen05 query kernel for every cluster size (see tools/pipe_probe.py)
for cl in 1 2 4; do echo "== cluster $cl"; DUDF_TC_CLUSTER=$cl timeout 120 python tools/pipe_probe.py 256 2>&1 | grep -E "full|neither|no epilogue|no MMA" ; done
