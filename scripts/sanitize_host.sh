#!/bin/bash
# Host logic under AddressSanitizer / UBSan (no GPU needed): builds an instrumented copy of the library OUT OF TREE, points the
# CPU tests at it (WB200_LIB) for the tests that drive this library's host code (grammar, VAD, DTW, dequantisers, sampler, KV, vocabulary, timestamps, the scripted
# whisper_full flows incl. the 3-thread lock-step driver) and restores the original.   usage: scripts/sanitize_host.sh address|undefined|thread
set -e
SAN=${1:-address}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/wb200_san_$SAN; mkdir -p $W
sed -e "s#^OBJDIR   := build#OBJDIR   := $W/build#" -e "s#-I../include -Icsrc#-I$ROOT/include -I$ROOT/whisper.cpp_b200/csrc#" \
    -e "s#csrc/#$ROOT/whisper.cpp_b200/csrc/#g" -e "s#\.\./include/\*\.h#$ROOT/include/*.h#" \
    -e "s#-fvisibility=hidden,-pthread#-fvisibility=hidden,-pthread,-fsanitize=$SAN,-fno-omit-frame-pointer,-g#" -e "s#-O3#-O1#" \
    -e "s#-Xlinker --no-undefined#-Xcompiler -fsanitize=$SAN#" $ROOT/whisper.cpp_b200/Makefile > $W/Makefile
cp $ROOT/whisper.cpp_b200/exports.map $W/
make -C $W -j8 libwhisper_b200.so > $W/build.log 2>&1
export WB200_LIB=$W/libwhisper_b200.so      # the Python wrapper loads this build instead of the in-tree library (nothing is swapped)
case $SAN in address) RT=asan;; undefined) RT=ubsan;; thread) RT=tsan;; esac
RT=$(gcc -print-file-name=lib$RT.so)
cd $ROOT
LD_PRELOAD=$RT TSAN_OPTIONS=halt_on_error=0:report_signal_unsafe=0 ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:verify_asan_link_order=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=0 \
  python -m pytest tests/test_grammar_cpu.py tests/test_vad_cpu.py tests/test_dtw_cpu.py tests/test_dequant_cpu.py tests/test_sampler_cpu.py tests/test_kv_cpu.py \
  tests/test_vocab_cpu.py tests/test_token_timestamps_cpu.py tests/test_full_scripted_cpu.py tests/test_beam_predraw_cpu.py tests/test_pool_cpu.py -x -q 2>&1 | grep -i "passed\|failed\|ERROR: AddressSanitizer\|runtime error\|WARNING: ThreadSanitizer" | sort | uniq -c
