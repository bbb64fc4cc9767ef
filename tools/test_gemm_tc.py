#!/usr/bin/env python
"""GPU check of zrb_gemm_f16 (tcgen05): every operand-major combination, ragged shapes,
against torch fp32 matmul of the same fp16-rounded operands; also times each case.
usage: python tools/test_gemm_tc.py <a_mn> <b_mn>"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zaremba_b200 import _lib

def pad(n, m=64): return (n + m - 1) // m * m

def run(a_mn, b_mn):
    lib = _lib.load()
    res = []
    shapes = [(128, 128, 64), (128, 128, 128), (256, 384, 192), (700, 6000, 1500), (700, 1500, 6000), (6000, 1500, 700),
              (700, 10000, 1500), (10000, 1500, 700), (20, 6000, 1500), (5, 37, 16), (130, 70, 100), (1, 8, 8), (35, 200, 200)]
    for (M, N, K) in shapes:
        g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
        A = torch.randn(M, K, device="cuda", generator=g).half()
        B = torch.randn(N, K, device="cuda", generator=g).half()
        bias = torch.randn(N, device="cuda", generator=g)
        want = 0.5 * (A.float() @ B.float().t()) + bias
        if a_mn:
            Ab = torch.zeros(K, pad(M), device="cuda", dtype=torch.half); Ab[:, :M] = A.t(); lda = pad(M)
        else:
            Ab = torch.zeros(M, pad(K), device="cuda", dtype=torch.half); Ab[:, :K] = A; lda = pad(K)
        if b_mn:
            Bb = torch.zeros(K, pad(N), device="cuda", dtype=torch.half); Bb[:, :N] = B.t(); ldb = pad(N)
        else:
            Bb = torch.zeros(N, pad(K), device="cuda", dtype=torch.half); Bb[:, :K] = B; ldb = pad(K)
        C = torch.full((M, N), 7.0, device="cuda")
        rc = lib.zrb_gemm_f16(_lib.ptr(Ab), lda, a_mn, _lib.ptr(Bb), ldb, b_mn, _lib.ptr(C), N, M, N, K, 0.5, _lib.ptr(bias), 0, None)
        torch.cuda.synchronize()
        if rc != 0:
            res.append(dict(shape=[M, N, K], rc=rc, err=lib.zrb_last_error().decode())); continue
        err = (C - want).abs().max().item(); scale = want.abs().max().item()
        # accumulate flag
        rc = lib.zrb_gemm_f16(_lib.ptr(Ab), lda, a_mn, _lib.ptr(Bb), ldb, b_mn, _lib.ptr(C), N, M, N, K, 0.5, None, 1, None)
        torch.cuda.synchronize()
        err2 = (C - (2 * want - bias)).abs().max().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): lib.zrb_gemm_f16(_lib.ptr(Ab), lda, a_mn, _lib.ptr(Bb), ldb, b_mn, _lib.ptr(C), N, M, N, K, 0.5, None, 0, None)
        e0.record()
        for _ in range(20): lib.zrb_gemm_f16(_lib.ptr(Ab), lda, a_mn, _lib.ptr(Bb), ldb, b_mn, _lib.ptr(C), N, M, N, K, 0.5, None, 0, None)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res.append(dict(shape=[M, N, K], max_err=err, acc_err=err2, scale=scale, us=round(us, 2), tflops=round(2 * M * N * K / us / 1e6, 1),
                        ok=bool(err < 2e-3 * max(scale, 1) and err2 < 4e-3 * max(scale, 1))))
    return res

if __name__ == "__main__":
    a_mn, b_mn = int(sys.argv[1]), int(sys.argv[2])
    out = run(a_mn, b_mn)
    print(json.dumps({"a_mn": a_mn, "b_mn": b_mn, "cases": out}))
    print("ALL_OK" if all(c.get("ok") for c in out) else "SOME_FAILED")
