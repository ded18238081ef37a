#!/bin/bash
# end-of-round evidence on 1 GPU (run under gpurun): tests, smoke, default bench (both arms), sanitizers
TAG=${1:-r02_final}
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/${TAG}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
SAN="tests/test_gpu_raster.py::test_sh_paths_forward_backward tests/test_gpu_raster.py::test_empty_and_all_culled tests/test_gpu_dropin.py::test_fused_deform_kernels_equal_the_separate_ones tests/test_gpu_dropin.py::test_avatar_step_matches_autograd_path tests/test_gpu_dropin.py::test_cuda_graph_replay_equals_eager_launches tests/test_gpu_lbs.py::test_against_reference_golden tests/test_gpu_output.py::test_frame_to_uint8_bit_exact tests/test_gpu_image_loss.py::test_ragged_size_no_mask_and_cpu_refusal tests/test_gpu_regularizers.py::test_region_laplacian_matches_reference_golden tests/test_gpu_regularizers.py::test_pcd_smoothing_matches_reference_golden tests/test_gpu_regularizers.py::test_l2norm_matches_reference_golden"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SAN -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_memcheck.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_dropin.py::test_fused_deform_kernels_equal_the_separate_ones tests/test_gpu_raster.py::test_sh_paths_forward_backward tests/test_gpu_regularizers.py::test_l2norm_matches_reference_golden tests/test_gpu_regularizers.py::test_pcd_smoothing_matches_reference_golden -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_racecheck.txt
python examples/train_step.py --steps 20 > gpurun_out/${TAG}_example.txt 2>&1
cat gpurun_out/${TAG}_tests.txt; tail -1 gpurun_out/${TAG}_smoke.txt; tail -c 400 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_memcheck.txt; tail -3 gpurun_out/${TAG}_racecheck.txt; tail -3 gpurun_out/${TAG}_example.txt
