#!/bin/bash
# compute-sanitizer over the final kernels (v12): memcheck on a subset that touches every kernel family, racecheck on the shared-memory users
mkdir -p gpurun_out
SEL="castle_small_gi or gi_frames_bit_exact or ray_batch_small_domain or frames_in_flight or ray_batch_many_instances or teapot_bounces or wide_bvh or tile_candidate"
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gi.py -m gpu -x -q -k "$SEL" > gpurun_out/r02_memcheck_pytest.log 2>&1
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_gi.py -m gpu -x -q -k "castle_small_gi or gi_frames_bit_exact or ray_batch_many_instances" > gpurun_out/r02_racecheck_pytest.log 2>&1
for f in gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log gpurun_out/r02_memcheck_pytest.log gpurun_out/r02_racecheck_pytest.log; do tail -n 3 $f; done
