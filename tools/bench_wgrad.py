"""Micro-benchmark of spf_wgrad_tc_multi on colour-field-shaped jobs: bf16 x bf16 operands vs bf16 dZ x fp16 activations
(the fp16 operand is converted to bf16 in shared memory inside the kernel).  usage: python tools/bench_wgrad.py [rows]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spurfies_b200 import _lib  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_200_000
rows = rows // 128 * 128
units = rows // 8
shapes = [(256, 256), (256, 256), (128, 112)]
g = torch.Generator(device="cuda").manual_seed(0)
dz_src = [torch.randn(rows + 128, 256, device="cuda", generator=g) for _ in shapes]
count = torch.tensor([units], dtype=torch.int32, device="cuda")
for name, fmt, dt in (("bf16 x bf16", 3, torch.bfloat16), ("bf16 x fp16 (in-kernel conversion)", 1, torch.float16), ("fp16 x fp16", 0, torch.float16)):
    act = [torch.randn(rows + 128, lda, device="cuda", generator=g).to(dt) for lda, _ in shapes]
    dz = [d.to(torch.float16 if fmt == 0 else torch.bfloat16) for d in dz_src]
    arr = (_lib.WgradJob * len(shapes))()
    outs = []
    for i, (lda, N) in enumerate(shapes):
        dW, db = torch.zeros(256, N, device="cuda"), torch.zeros(256, device="cuda")
        outs.append((dW, db))
        arr[i].dz, arr[i].act, arr[i].dW, arr[i].db = dz[i].data_ptr(), act[i].data_ptr(), dW.data_ptr(), db.data_ptr()
        arr[i].lda, arr[i].N, arr[i].fmt = lda, N, fmt
    run = lambda: _lib.call("spf_wgrad_tc_multi", C.cast(arr, C.c_void_p), len(shapes), _lib.ptr(count), 8, units, None, _lib.stream())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = rows * sum(512 + lda * 2 for lda, _ in shapes) / 1e9
    print(f"{name:38s} {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  ({gb:.2f} GB)")
