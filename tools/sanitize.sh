#!/usr/bin/env bash
# compute-sanitizer over the small GPU parity tests (race detection / sanitizers).
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# Run it on a GPU box (e.g. `gpurun --timeout 1500 -- 'tools/sanitize.sh racecheck > gpurun_out/racecheck.log 2>&1'`).
# The tests it selects are the ones that finish in seconds natively; under the sanitizer expect minutes.  The TMA /
# mbarrier kernels are covered by memcheck and synccheck; racecheck follows shared-memory accesses of ordinary loads
# and stores (the tiled deposit's atomics, the FFT exchange buffers), not the async-proxy copies.
# NOT run in rounds 1-2 (the GPU budget went to parity, bench and ncu); the bitwise checks that stand in for it are
# the fixed-point deposit tests (run-to-run, permutation and decomposition invariance) and the bit-identical transport
# variants of tests/mgpu_worker.py.
set -euo pipefail
tool="${1:-memcheck}"
expr="${2:-test_deposit or test_potential or test_fused_step or test_tiled or test_sort or test_power_spectrum}"
cd "$(dirname "$0")/.."
exec compute-sanitizer --tool "$tool" --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$expr"
