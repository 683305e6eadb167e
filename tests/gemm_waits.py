"""Debug: per-role wait-cycle breakdown of the 1-CTA dequant-GEMM (needs a library built with
-DSAMQ_PROFILE_WAITS; not part of the test suite)."""
import ctypes, os, sys
os.environ["SAMQ_GEMM"] = "1cta"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sam_quantization_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import rand_packed, dev
M, K, N = (int(v) for v in sys.argv[1:4])
epi = _lib.EPI_GELU if len(sys.argv) > 4 and sys.argv[4] == "gelu" else 0
d = torch.device("cuda:0")
qw, qz, sc, _ = rand_packed(K, N, 4, 128, seed=1)
tq, tz, ts = dev(qw, d), dev(qz, d), dev(sc, d)
x = torch.randn(M, K, device=d).half(); y = torch.empty(M, N, device=d, dtype=torch.float16)
for _ in range(3): ops.qlinear(x, tq, tz, ts, 4, 128, epilogue=epi, out=y)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (148 * 32))()
lib = _lib.load(); lib.samq_debug_read_waits.argtypes = [ctypes.c_void_p]
assert lib.samq_debug_read_waits(buf) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 32).astype(np.float64)
names = {0: "mma total", 1: "mma wait acc_empty", 2: "mma wait a_full", 3: "mma wait x_full", 4: "deq0 total", 5: "deq0 wait w_full",
         6: "deq0 wait a_empty", 7: "deq0 st+wait", 8: "deq1 total", 9: "deq1 wait w_full", 10: "deq1 wait a_empty", 11: "deq1 st+wait",
         16: "deq0 pre(consts)", 17: "deq0 lds", 18: "deq0 math", 20: "deq1 pre", 21: "deq1 lds", 22: "deq1 math", 12: "epi total", 13: "epi wait acc_full", 14: "tma wait w_empty", 15: "tma wait x_empty"}
tiles = (N // 128) * ((M + 191) // 192); kbs = tiles * (K // 64) / 148
print(f"k-blocks per CTA {kbs:.0f}; cycles per k-block (mma total) {a[:,0].mean()/kbs:.0f}")
for i, n in names.items():
    print(f"{n:22s} mean {a[:, i].mean():12.0f} cyc  per-kblock {a[:, i].mean()/kbs:7.1f}  ({100*a[:, i].mean()/a[:, 0].mean():5.1f}% of mma total)")
