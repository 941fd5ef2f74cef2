#!/bin/bash
# The host half (host.cpp: RNG, plan builder, planner threads, ingest, writers, proto dump, contributions) under
# AddressSanitizer + UBSan, then ThreadSanitizer, driven by the CPU test suite.  No GPU needed.  The in-tree library is
# replaced by an instrumented build for the run and rebuilt afterwards.
#   tools/sanitize_host.sh            (last run: 0 reports from either pass)
set -u
cd "$(dirname "$0")/.."
LIB=go_gsgp_b200/libgsgp_b200.so
SRC="go_gsgp_b200/csrc/gsgp_b200.cu go_gsgp_b200/csrc/host.cpp"
TESTS="tests/test_plan_channel.py tests/test_example_host.py tests/test_host_driver.py tests/test_ingest.py tests/test_program_builder.py tests/test_capi_symbols.py tests/test_golden.py tests/test_proto_dump.py tests/test_contributions.py"
run() {   # $1 = -fsanitize flags, $2 = libraries to preload, $3 = linker libs, $4 = log
    nvcc -O1 -g -std=c++17 -shared -Xcompiler -fPIC,$1,-fno-omit-frame-pointer -gencode arch=compute_100a,code=sm_100a \
        -o $LIB $SRC -ldl -lz $3 || return 1
    LD_PRELOAD="$2" ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
        TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" \
        python -m pytest $TESTS -q -s -m "not gpu" -k "not streaming_pass" -p no:cacheprovider > "$4" 2>&1
    tail -1 "$4"
    echo "reports: $(grep -c 'runtime error\|ERROR: AddressSanitizer\|WARNING: ThreadSanitizer' "$4")"
}
run "-fsanitize=address,-fsanitize=undefined" "$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" "-Xlinker -lasan -Xlinker -lubsan" /tmp/gsgp_asan.log
run "-fsanitize=thread" "$(gcc -print-file-name=libtsan.so)" "-Xlinker -ltsan" /tmp/gsgp_tsan.log
python -m go_gsgp_b200.build --force > /dev/null && echo "production library rebuilt"
