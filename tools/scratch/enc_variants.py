"""Timing experiment (not a product path): variants of encode_mask_planes_kernel built as separate libraries and timed on
the same box, checked against the product library's output.  Build the variants first, e.g.
    F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -cudart static -I include -shared"
    SRC="matrix0_b200/csrc/encode_kernels.cu matrix0_b200/csrc/capi_common.cu"
    nvcc $F -o tools/scratch/libenc_final.so $SRC -lcuda
    nvcc $F -DENC_EXP_STORES_ONLY -o tools/scratch/libenc_stores.so $SRC -lcuda     # store pattern without the move logic
    nvcc $F -DENCW_MIN_BLOCKS=6 -o tools/scratch/libenc_mb6.so $SRC -lcuda          # occupancy variants
then `gpurun -- python tools/scratch/enc_variants.py` (writes gpurun_out/enc_variants.json)."""
import ctypes, glob, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from matrix0_b200 import _native
lib = _native.lib()
n = 1 << 20
dev = torch.device("cuda")
pos = torch.empty((n, 9), dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
_native.check(lib.m0_random_playouts(pos.data_ptr(), n, 1234, 120, s))
planes = torch.empty((n, 19, 8, 8), dtype=torch.float32, device=dev)
mask = torch.empty((n, 4672), dtype=torch.uint8, device=dev)
_native.check(lib.m0_encode_positions(pos.data_ptr(), n, planes.data_ptr(), mask.data_ptr(), 0, 0, 0, s))
torch.cuda.synchronize()
ref_m = mask.clone(); ref_p = planes.clone()
out = {}
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "libenc_*.so"))):
    L = ctypes.CDLL(path)
    L.m0_encode_positions.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_void_p]
    f = lambda: L.m0_encode_positions(pos.data_ptr(), n, planes.data_ptr(), mask.data_ptr(), None, None, None, s)
    for _ in range(3): assert f() == 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    out[os.path.basename(path)] = {"ms": ms, "GBps": n * 9608 / ms / 1e6, "same": bool(torch.equal(mask, ref_m) and torch.equal(planes, ref_p))}
    print(os.path.basename(path), out[os.path.basename(path)], flush=True)
json.dump(out, open("gpurun_out/enc_variants.json", "w"), indent=1)
