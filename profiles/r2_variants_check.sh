#!/bin/bash
# GPU check of the kernel variants (VF_NODE_WARP, VF_P2_WARP): parity tests of both settings,
# A/B timing, then the rest of the GPU suite (benchmark-size tests last / separately).
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernel_variants.py -q --durations=5 > gpurun_out/r2_variants_tests.log 2>&1
echo "variants rc=$?" >> gpurun_out/r2_variants_tests.log
timeout 120 python profiles/variants_ab.py 1e6 6 20 > gpurun_out/r2_variants_ab.json 2> gpurun_out/r2_variants_ab.err
timeout 200 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_p2.py tests/test_gpu_forward.py tests/test_gpu_fluid_static.py tests/test_gpu_partition.py tests/test_postprocess.py tests/test_gpu_gridsolve.py -m gpu -q --durations=8 > gpurun_out/r2_suite_rest.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2_suite_rest.log
tail -3 gpurun_out/r2_variants_tests.log; cat gpurun_out/r2_variants_ab.json; tail -3 gpurun_out/r2_suite_rest.log
