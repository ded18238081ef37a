#!/bin/bash
# mid-round evidence on 1 GPU (run under gpurun): full GPU suite, smoke, default bench, memcheck of the newest kernels, the example
TAG=${1:-r02_quick}
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/${TAG}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_regularizers.py::test_region_laplacian_matches_reference_golden tests/test_gpu_regularizers.py::test_pcd_smoothing_matches_reference_golden tests/test_gpu_regularizers.py::test_l2norm_matches_reference_golden tests/test_gpu_raster.py::test_pair_list_overflow_is_repaired_and_async_mode_reports_it -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_memcheck.txt
python examples/train_step.py --steps 20 > gpurun_out/${TAG}_example.txt 2>&1
cat gpurun_out/${TAG}_tests.txt; tail -1 gpurun_out/${TAG}_smoke.txt; tail -c 300 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_memcheck.txt; tail -4 gpurun_out/${TAG}_example.txt
