"""Development tool: per-event SM-clock timeline of the gen-2 geometry kernel (library built with SPF_TIMELINE=1)."""
import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
import bench
from spurfies_b200 import _lib, fields
from spurfies_b200.fields import SlotSet, geo_sdf_raw
dev = torch.device('cuda', 0)
sc, model = bench.build_scene(dev, precision='bf16')
g = torch.Generator().manual_seed(0)
n = 160000
q = (sc["pts"][torch.randint(0, 100000, (n,), generator=g)] + 0.01 * torch.randn(n, 3, generator=g)).cuda().contiguous()
slots = SlotSet(model._grid().query_points(q, 8, 2.0))
pack = model._pack()
fields.set_precision("bf16")
buf = (C.c_ulonglong * (4 * 8192))()
for i in range(3):
    torch.cuda.synchronize()
    _lib.lib.spf_debug_timeline(buf, 8192)  # reset
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    geo_sdf_raw(pack, slots, q, model.neural_pts, model.neural_feats_geometry.detach(), 45.0, True, True)
    e1.record()
    torch.cuda.synchronize()
    print("V", slots.V, "pairs", slots.V * 8, "ms", e0.elapsed_time(e1))
for mode in (1, 2, 0):
    _lib.lib.spf_debug_mode(mode)
    for i in range(2):
        torch.cuda.synchronize()
        _lib.lib.spf_debug_timeline(buf, 8192)  # reset
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        geo_sdf_raw(pack, slots, q, model.neural_pts, model.neural_feats_geometry.detach(), 45.0, True, True)
        e1.record()
        torch.cuda.synchronize()
        print("debug mode", mode, "(1: no TMEM loads, 2: loads only, 0: normal) ms", e0.elapsed_time(e1))
nev = _lib.lib.spf_debug_timeline(buf, 8192)
ev = sorted([(buf[4*i+3], buf[4*i], buf[4*i+1], buf[4*i+2]) for i in range(nev)])
t0 = ev[0][0]
names = {0: "E.acc_seen", 1: "E.compute_done", 2: "E.signalled", 3: "E.gather_done", 4: "E.bar_passed", 10: "M.a_ready_seen", 11: "M.issued"}
print("events", nev)
for c, e, t, l in ev[:260]:
    print(f"{c - t0:9d} cyc  {'  ' if t == 0 else '                          '}tile{t} L{l} {names.get(e, e)}")
