#!/bin/bash
# per-kernel CUDA-event times of the conforming step (bench.py --light, roofline.kernel_ms) + a parity spot check
python bench.py --steps 20 --warmup 3 --light 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],4), 'sustained', round(d['aux']['sustained_ms_per_step'],4), 'kernels', d['roofline']['kernel_ms'])"
python -m pytest -q tests/test_gpu_tcx3.py -m gpu -k "loss_terms or ragged_and" 2>&1 | tail -2
