#!/bin/sh
# compute-sanitizer is closed on the GPU pool, so memory safety of the kernels and of the host layer
# is checked here instead: the product sources compiled against the SIMT emulator with
# AddressSanitizer, driven through the C ABI by the parity cases.
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/rtk_asan
g++ -x c++ -std=c++17 -O1 -g -fPIC -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -mfma -w \
    -DRTK_SIMT_EMU=1 -DSIMT_IMPL=1 -include tests/emu/simt.h -c rtk_b200/csrc/rtk_device.cu -o /tmp/rtk_asan/dev.o
gcc -O1 -g -fPIC -fsanitize=address -std=gnu11 -c rtk_b200/csrc/rtk_host.c -o /tmp/rtk_asan/host.o
gcc -O1 -g -fPIC -fsanitize=address -std=gnu11 -c rtk_b200/csrc/rtk_place.c -o /tmp/rtk_asan/place.o
g++ -shared -fsanitize=address -o /tmp/rtk_asan/librtk_emu_asan.so /tmp/rtk_asan/dev.o /tmp/rtk_asan/host.o /tmp/rtk_asan/place.o -lpthread -lm
cat > /tmp/rtk_asan/run.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from rtk_b200 import api
from oracle import orc
import parity_cases as pc
lib = api.Library('/tmp/rtk_asan/librtk_emu_asan.so')
assert lib.rtk_cuda_init(0) == 0
pc.case_kats(lib, orc)
pc.case_edge_scenes(lib, orc)
for mode in (0, 1):
    pc.case_config(lib, orc, "C3", 0.004, 600, mode=mode)
pc.case_config(lib, orc, "C4", 0.0005, 500, mode=1)
pc.case_mesh_formats(lib, orc)
pc.case_api_semantics(lib, orc)
pc.case_ties(lib, orc)
pc.case_wavefront(lib, orc, pc.HostDevice())
pc.case_refit(lib, orc, pc.HostDevice())
pc.case_deep_stack(lib, orc, pc.HostDevice())
pc.case_host_batch_chunks(lib, orc, nrays=20000, chunk_log2=12)
pc.case_triangle_filter(lib, orc, pc.HostDevice())
pc.case_instancing(lib, orc, pc.HostDevice())
pc.case_threads(lib, orc)
print("asan run clean")
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 python /tmp/rtk_asan/run.py
