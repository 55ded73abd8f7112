"""tcgen05.mma issue/throughput probe: cycles per 128 x N x 16 f16 MMA issued back to back by one thread (GPU box only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ggml_experiments_b200 as G
L = G.lib_ggml()
L.ggml_b200_debug_mma_rate.restype = ctypes.c_float
L.ggml_b200_debug_mma_rate.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_float)]
for ctas in (1, 2):
    for n in (64, 128, 256):
        for ce in (0, -1):
            iss = ctypes.c_float(0)
            tot = L.ggml_b200_debug_mma_rate(n, 2048, ctas, ce, ctypes.byref(iss))
            ideal = 128 * n * 16 * 2 / 8192.0
            print(f"ctas/SM={ctas} N={n:3d} {'thread-0 issue (divergent)' if ce == 0 else 'warp-uniform elect issue':>28}: {iss.value:7.1f} cycles/MMA to issue, {tot:7.1f} incl. completion (math alone {ideal:5.1f})", flush=True)
for n in (96, 128):
    for ce in (-1, -4, -8, -12, -24):
        iss = ctypes.c_float(0)
        tot = L.ggml_b200_debug_mma_rate(n, 2304, 1, ce, ctypes.byref(iss))
        print(f"N={n:3d} uniform issue, {'no commits' if ce == -1 else f'async commit every {-ce} MMAs, 2 accumulators'}: {iss.value:7.1f} cycles/MMA to issue, {tot:7.1f} incl. completion", flush=True)
