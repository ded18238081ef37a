#!/bin/bash
# Run the emulated-kernel tests (tests/cuda_emu: kernel sources compiled for the host) under a compiler sanitizer:
#   tools/emu_sanitize.sh address   out-of-bounds / use-after-free in the kernel code (shared arrays are globals with red zones)
#   tools/emu_sanitize.sh thread    data races between the threads of a block (blocks run one after the other)
# CPU only; complements compute-sanitizer memcheck / racecheck on the GPU (profiles/*_memcheck.txt, *_racecheck.txt).
SAN=${1:-address}
shift
TESTS=${@:-tests/test_raster_emulated.py tests/test_lbs_kernels_emulated.py tests/test_fused_kernels_emulated.py tests/test_reg_kernels_emulated.py tests/test_stream_kernels_emulated.py}
LIB=$(gcc -print-file-name=lib$([ "$SAN" = thread ] && echo tsan || echo asan).so)
export SGS_EMU_SANITIZE=$SAN
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1
export TSAN_OPTIONS=halt_on_error=0:report_signal_unsafe=0
# -s: sanitizer reports go to stderr; pytest's capture would swallow those of passing tests
LD_PRELOAD=$LIB python -m pytest $TESTS -x -q -s -p no:cacheprovider 2>&1 | grep -v '^$' | tail -60
