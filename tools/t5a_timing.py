"""Phase timing of attn_fwd_t5_kernel (debug build with -DT5A_TIMING, see csrc/window_attn.cu):
clock64 stamps of thread 0 / CTA 0 for its 4th item.

    cd underwater-image-restoration_b200/csrc
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DT5A_TIMING -c window_attn.cu -o /tmp/wa_dbg.o
    nvcc -shared -o build/libuwr_dbg.so $(ls build/*.o | grep -v window_attn.o) /tmp/wa_dbg.o
    UWR_B200_LIB=$PWD/build/libuwr_dbg.so python ../../tools/t5a_timing.py        # on the GPU box
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from uwr import ops, _lib
B, H, heads, hd = 16, 256, 2, 32
C = heads * hd
qkv = torch.randn(B * H * H, 3 * C, device="cuda")
table = torch.randn(225, heads, device="cuda") * 0.02
w = torch.ones(2, device="cuda")
ops.set_attn_tcgen05(True)
for _ in range(3):
    ops.window_attn_fwd(qkv, 0, qkv, C, 2 * C, table, w, B, H, H, heads, hd, 4, hd ** -0.5)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
f = _lib.lib.uwr_attn_t5_debug
f.restype = ctypes.c_int
assert f(buf) == 0
t = list(buf)[:7]
names = ["wait loads", "lo split + fence + sync", "S MMA (12) + commit wait", "issue next loads + tmem ld + softmax + P write + sync",
         "O MMA (16) + commit wait", "tmem ld + store + sync"]
for i, n in enumerate(names):
    print(f"{n:58s} {t[i + 1] - t[i]:7d} cycles")
print(f"{'item total':58s} {t[6] - t[0]:7d} cycles")
