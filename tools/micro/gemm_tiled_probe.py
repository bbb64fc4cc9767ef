#!/usr/bin/env python
"""EXPERIMENT: is the tcgen05 GEMM bound by how its operand tiles are fetched?  Same GEMM ([700 x 6000 x 1500] and the
projection shape, K-major A and B) with (a) 2-D TMA tensor loads for both operands, (b) B as a pre-tiled / pre-swizzled image
fetched with 1-D bulk copies, (c) A and B both pre-tiled.  Results are checked against torch."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from zaremba_b200 import _lib
lib = _lib.load()
lib.zrb_gemm_f16_tiled.restype = C.c_int
lib.zrb_gemm_f16_tiled.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p]


def tiled(X, even):
    """X [rows, K] half -> [K/64][rows/128][128][64] with chunk c of row r at c ^ (r % 8)"""
    rows, K = X.shape
    nt = (rows + 127) // 128
    if even: nt = (nt + 1) // 2 * 2
    kb = (K + 63) // 64
    P = torch.zeros(nt * 128, kb * 64, device=X.device, dtype=X.dtype); P[:rows, :K] = X
    T = P.view(nt, 128, kb, 8, 8).permute(2, 0, 1, 3, 4).contiguous()          # [kb][nt][128][chunk][8]
    r = torch.arange(128, device=X.device).view(1, 1, 128, 1, 1)
    c = torch.arange(8, device=X.device).view(1, 1, 1, 8, 1)
    src = (c ^ (r % 8)).expand(kb, nt, 128, 8, 8)                               # dest chunk c holds source chunk c ^ (r%8)
    return torch.gather(T, 3, src).contiguous(), nt


out = []
for (M, N, K) in [(700, 6000, 1500), (700, 10000, 1500)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).half(); B = torch.randn(N, K, device="cuda", generator=g).half()
    Kp = (K + 63) // 64 * 64
    Ab = torch.zeros(M, Kp, device="cuda", dtype=torch.half); Ab[:, :K] = A
    Bb = torch.zeros(N, Kp, device="cuda", dtype=torch.half); Bb[:, :K] = B
    At, a_nt = tiled(A, False); Bt, b_nt = tiled(B, True)
    want = A.float() @ B.float().t()
    Cc = torch.empty(M, N, device="cuda")
    res = {}
    for name, at, bt in (("tma2d", None, None), ("B_tiled", None, Bt), ("A_and_B_tiled", At, Bt)):
        call = lambda: lib.zrb_gemm_f16_tiled(Ab.data_ptr(), Kp, Bb.data_ptr(), Kp, at.data_ptr() if at is not None else None, a_nt,
                                              bt.data_ptr() if bt is not None else None, b_nt, Cc.data_ptr(), N, M, N, K, 1.0, None)
        assert call() == 0, lib.zrb_last_error()
        torch.cuda.synchronize()
        err = (Cc - want).abs().max().item() / want.abs().max().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5): call()
        e0.record()
        for _ in range(30): call()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        res[name] = {"us": round(us, 2), "tflops": round(2 * M * N * K / us / 1e6, 1), "rel_err": err}
    out.append({"shape": [M, N, K], **res})
print(json.dumps(out, indent=1))
