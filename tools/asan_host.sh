#!/bin/sh
# AddressSanitizer + UndefinedBehaviorSanitizer over the host adapters (fun::b200_receiver_chain, fun::b200_rx,
# fun::b200_receiver, host_capi.cpp), compiled together with the CPU test double of the C ABI (tests/fake/fake_b200rx.cpp:
# the reference's blocks behind the b200rx_* entry points) and driven by tests/fake/asan_main.cpp: five call sizes x three
# pipeline depths, by-value and pointer calls, pinned caller buffers (direct path), the block on a tagged stream, the
# receiver loop with pause / resume.  Needs oracle/_ref/libfunref.so (make -C oracle ref).  Leak checking is off: the
# reference's own fft objects are never freed (fft.cpp:34-39).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=${TMPDIR:-/tmp}/b200_asan
mkdir -p "$W"
cd "$ROOT"
python - "$W" <<'PY'
import os, sys
W = sys.argv[1]
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import numpy as np
from oracle import bind
from stress_receive import make_stream
ref = bind.ref()
rng = np.random.default_rng(5)
x = np.concatenate([make_stream(ref, rng)[0] for _ in range(6)])
x.astype(np.complex128).tofile(os.path.join(W, "stream.bin"))
s, t = ref.sync(x, chunk=4096)
n = (len(t) // 4096) * 4096
np.ascontiguousarray(s[:n]).astype(np.complex128).tofile(os.path.join(W, "synced.bin"))
np.ascontiguousarray(t[:n]).astype(np.uint8).tofile(os.path.join(W, "tags.bin"))
os._exit(0)
PY
H=fun_ofdm_b200/host
# a g++ that ships libasan (the distribution's; a toolchain under /opt may not)
GXX=/usr/bin/g++; [ -x "$GXX" ] || GXX=g++
sed "s#/tmp/asan/#$W/#g" tests/fake/asan_main.cpp > "$W/main.cpp"
$GXX -std=c++17 -g -O1 -fsanitize=address,undefined -fno-omit-frame-pointer -pthread -Iinclude \
    tests/fake/fake_b200rx.cpp $H/b200_rx.cpp $H/b200_receiver_chain.cpp $H/b200_receiver.cpp $H/host_capi.cpp "$W/main.cpp" \
    -o "$W/asan_test" -ldl
B200RX_FAKE_REF="$ROOT/oracle/_ref/libfunref.so" ASAN_OPTIONS=detect_leaks=0 "$W/asan_test" 2>&1 | grep -v "^Invalid CRC"
